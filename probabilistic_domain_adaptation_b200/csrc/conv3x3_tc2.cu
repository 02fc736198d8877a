// CTA-pair (cta_group::2) variant of the implicit-GEMM conv3x3 of conv3x3_tc.cu.
//
// Why: an M = 128 x N x K = 16 tcgen05.mma needs N/2 tensor cycles but fetches 4 KB (A) + N * 32 B (B) of operands
// through a shared-memory port that delivers 128 B/clk per SM (profiles/r01e_umma_issue_rate.md): N = 64 tiles are capped
// at 2/3 of the tensor rate, N = 128 tiles are exactly balanced -- and the TMA fill shares the port.  Two CTAs of one
// cluster (the two SMs of a TPC) issue ONE M = 256 MMA instead: each SM multiplies its own 128 pixels (its own A slabs)
// with the SAME weight tile, of which each CTA holds only half the rows (N/2) -- per SM the B fetch and the B fill are
// halved (N = 64: 40 instead of 48 port cycles per 32 tensor cycles; N = 128: 48 per 64: tensor-bound).
//
// Same work decomposition, slabs, tap reuse, epilogue and outputs as conv3x3_tc_kernel; what changes is the plumbing:
//   * cluster of 2 CTAs; a pair-unit = two adjacent pixel tiles x one block of BN output channels
//   * both CTAs run their TMA producer (own A slabs, own half of the weight tile); the transaction bytes of both land on
//     the LEADER's (rank 0) full barriers (cp.async.bulk.tensor ... .cta_group::2 with the barrier address mapped to rank
//     0); the leader's producer alone arrives on them, announcing the bytes of both CTAs
//   * only the leader's MMA warp issues tcgen05.mma.cta_group::2; its commits are multicast to the barriers of both CTAs
//     (smem slot release, accumulator-full)
//   * each CTA's epilogue drains its own TMEM (its 128 rows of the M = 256 accumulator); the peer's epilogue warps
//     release the accumulator buffer with remote arrives on the leader's barrier
#include <stdio.h>
#include <stdlib.h>

#include "conv.cuh"
#include "ptx.cuh"

#ifndef PDA_PDL_TRIGGER_EARLY
#define PDA_PDL_TRIGGER_EARLY 0
#endif
#ifndef PDA_CONV_WIDE_DEFAULT
#define PDA_CONV_WIDE_DEFAULT 1
#endif
#ifndef PDA_UPS_A_STAGES64
#define PDA_UPS_A_STAGES64 2
#endif
// pixels per slab row of the fused up-sampling mode (the taps read 10; 16 = the WIDE layout)
#ifndef PDA_UPS_ROW_PX
#define PDA_UPS_ROW_PX 10
#endif

namespace pda {

__device__ __forceinline__ void named_bar2(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory offset in CTA `rank` of this cluster
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// (default semantics, as a local arrive: the explicit .release.cluster form costs a MEMBAR.ALL + ERRBAR per arrive, which
// also waits for the warp's outstanding global stores; the TMEM reads this arrive publishes are ordered by
// tcgen05.wait::ld + tcgen05.fence::before_thread_sync)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA loads whose completion bytes may be signalled on a barrier of the peer CTA of the pair
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* m, uint32_t bar_cluster, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_pair(uint32_t dst, const CUtensorMap* m, uint32_t bar_cluster, int c0, int c1,
                                                 int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A[smem of both CTAs, 128 rows each] * B[smem, N/2 rows in each CTA]; issued by the leader
__device__ __forceinline__ void umma_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives (count 1) on the barrier at this shared-memory offset in BOTH CTAs once all earlier MMAs of this thread are done
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(bar), "h"(static_cast<uint16_t>(3))
      : "memory");
}

// WIDE (default): ONE slab of 10-pixel rows per 64-channel chunk serves all three kx taps (the tap's A operand starts kx
// pixels = kx * 128 B into the row; the 1280-byte pitch need not be a whole number of swizzle atoms) instead of three
// column-shifted 8-pixel slabs: one TMA load and 42 % of the shared-memory fill traffic per chunk, one more stage of
// look-ahead.  Bit-identical results.  Measured per layer at 4 x 1024^2 (profiles/r02_conv_wide_slab.md): +1..4 % on
// nine of the ten layer shapes (64 -> 64: 958 -> 997 TFLOP/s), -2.6 % on the single-chunk 64 -> 128 layer, which keeps
// the 8-pixel slabs; the first version with 16-pixel rows and two stages was neutral.  PDA_CONV_WIDE=0 selects the
// 8-pixel slabs everywhere.
template <int BN, int MT, bool RES, bool WIDE = false, bool UPS = false>
struct Conv2Cfg {
  static constexpr int SLAB_ROWS = 16 * MT + 2;
  // one slab row: 16 or 8 px x 128 B.  UPS: PDA_UPS_ROW_PX pixels -- the pitch need not be a whole number of 1024-byte
  // swizzle atoms because both TMA and the MMA unit derive the swizzle from shared-memory ADDRESS bits
  static constexpr int ROW_BYTES = (UPS || WIDE) ? PDA_UPS_ROW_PX * 128 : 1024;   // (WIDE used 16-pixel rows at first)
  // UPS: the low-resolution patch one slab interpolates from (rows x 7 px x 64 channels, unswizzled), double-buffered
  static constexpr int PATCH_ROWS = 8 * MT + 3;
  static constexpr int PATCH_PX = 7;
  static constexpr int PATCH_BYTES = PATCH_ROWS * PATCH_PX * 128;
  static constexpr int P_STAGES = UPS ? 2 : 0;
  static constexpr int A_TX = SLAB_ROWS * ROW_BYTES;             // bytes one TMA slab load delivers
  static constexpr int A_BYTES = (A_TX + 1023) & ~1023;          // stage pitch: whole swizzle atoms
  static constexpr int BH = BN / 2;                    // weight rows held by one CTA of the pair
  static constexpr int B_BYTES = BH * 128;             // this CTA's half of one (tap, chunk) weight tile
  // (UPS, N = 64: a third slab stage -- PDA_UPS_A_STAGES64 = 3, paid for with two of the eight weight stages -- measured
  // no faster: that layer is bound by the shared-memory port, which the software producer's reads and writes share)
  static constexpr int A_STAGES = UPS ? (MT == 2 ? (BN == 64 ? PDA_UPS_A_STAGES64 : 2) : 3) : (WIDE ? (MT == 2 ? 3 : 4) : 4);
  static constexpr int B_STAGES =
      RES ? 9 : (BN == 256 ? 4 : (MT == 2 ? (BN == 128 ? 5 : ((UPS && PDA_UPS_A_STAGES64 == 3) ? 6 : 8)) : 6));
  static constexpr int A_OFF = 0;
  static constexpr int B_OFF = A_STAGES * A_BYTES;
  static constexpr int STG_OFF = B_OFF + B_STAGES * B_BYTES;   // 8 epilogue warps x 4 KB output staging
  static constexpr int P_OFF = STG_OFF + 8 * 4096;
  static constexpr int BAR_OFF = P_OFF + P_STAGES * PATCH_BYTES;
  static constexpr int NBARS = 2 * A_STAGES + 2 * B_STAGES + 4 + P_STAGES + (UPS ? A_STAGES : 0);
  static constexpr int SLOT_OFF = BAR_OFF + NBARS * 8;
  static constexpr int BIAS_OFF = SLOT_OFF + 16;
  static constexpr int MAX_COUT = 512;
  static constexpr int TOTAL = BIAS_OFF + MAX_COUT * 4;
  static constexpr int DYN_BYTES = TOTAL + 1024;
  static constexpr int TMEM_COLS = 2 * MT * BN;
  static_assert(TMEM_COLS <= 512 && (TMEM_COLS & (TMEM_COLS - 1)) == 0, "TMEM columns");
  static_assert(DYN_BYTES <= 227 * 1024, "shared memory");
  static_assert(B_BYTES % 1024 == 0, "weight half-tile must be whole swizzle atoms");
  static_assert(A_BYTES % 1024 == 0 && PATCH_BYTES % 128 == 0, "stage alignment");
};

constexpr int CONV2_THREADS = 320;

// UPS: K segment 0 is the bilinear x2 up-sampling (align_corners = True: unet_blocks.py:51) of a LOW-RESOLUTION tensor
// p.up_src [B][H/2][W/2][c0]; it is never materialised.  TMA cannot interpolate, so the eight epilogue warps -- idle for
// ~90 % of a unit -- are the producer of those chunks:
//   * TMA stages the low-resolution patch a slab interpolates from ((8 MT + 3) rows x 7 px x 64 channels, unswizzled) in
//     a two-deep ring, issued two chunks ahead by the producer's thread 0 (tmA0 is the patch map in this mode)
//   * 240 of the 256 threads own one (pixel column, 8-channel group) of the slab and a run of consecutive rows; they blend
//     in fp32 in the order upsample2x_kernel uses (bit-identical to up-sampling first), write the slab swizzled the way
//     TMA would have written it, fence it for the async proxy, and thread 0 arrives on the stage's full barrier (count 2
//     in this mode: one arrival per CTA for every chunk, software-produced or TMA-loaded)
//   * slab rows are PDA_UPS_ROW_PX = 10 pixels (the three kx taps read pixels kx .. kx + 7), not the WIDE layout's 16
//   * per unit they first produce the unit's segment-0 chunks, then drain the PREVIOUS unit's accumulators
//   * the MMA warp hands a consumed stage back on the barrier of the producer that owns the stage's NEXT use (TMA warp or
//     epilogue warps): neither producer ever watches the other's uses go by (see the note at nchunk_sw)
// Measured at the 4 x 1024^2 up-path shapes (tools/conv_up_bench.py): 0.66 / 0.76 / 1.10 ms (upsample2x + conv) ->
// 0.57 / 0.60 / 0.88 ms.
template <int BN, int MT, bool RES, bool F16, bool WIDE, bool UPS>
__global__ void __launch_bounds__(CONV2_THREADS, 1)
conv3x3_tc2_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                   const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmOut,
                   const ConvArgs p) {
  using L = Conv2Cfg<BN, MT, RES, WIDE, UPS>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar0 = sbase + L::BAR_OFF;
  auto a_full = [&](int s) { return bar0 + 8u * s; };
  auto a_empty = [&](int s) { return bar0 + 8u * (L::A_STAGES + s); };
  auto b_full = [&](int s) { return bar0 + 8u * (2 * L::A_STAGES + s); };
  auto b_empty = [&](int s) { return bar0 + 8u * (2 * L::A_STAGES + L::B_STAGES + s); };
  auto acc_full = [&](int s) { return bar0 + 8u * (2 * L::A_STAGES + 2 * L::B_STAGES + s); };
  auto acc_empty = [&](int s) { return bar0 + 8u * (2 * L::A_STAGES + 2 * L::B_STAGES + 2 + s); };
  auto p_full = [&](int s) { return bar0 + 8u * (2 * L::A_STAGES + 2 * L::B_STAGES + 4 + s); };
  // UPS: "stage s is free and its NEXT use is a software-produced chunk" (a_empty(s): ... a TMA-loaded chunk)
  auto a_empty_sw = [&](int s) { return bar0 + 8u * (2 * L::A_STAGES + 2 * L::B_STAGES + 4 + L::P_STAGES + s); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + L::SLOT_OFF);
  float* bias_s = reinterpret_cast<float*>(smem + L::BIAS_OFF);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();     // 0 = leader (issues the MMAs), 1 = peer
  const int pair = blockIdx.x >> 1;
  const int npairs = gridDim.x >> 1;
  const int ctot = p.c0 + p.c1;
  const int chunks = ctot >> 6;
  const int n_blocks = p.cout / BN;
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const int mtiles = tiles_per_img * p.B;
  const int mpairs = (mtiles + 1) >> 1;
  const int units = mpairs * n_blocks;         // pair-units

  if (threadIdx.x == 0) {
    // full barriers are only used in the leader: ONE arrival per phase (the leader's arrive.expect_tx, which announces
    // the bytes of BOTH CTAs); the peer's TMA loads just complete their bytes on it.  Bytes of the peer that land before
    // the leader's expect_tx leave the transaction count negative while the arrival is still pending: no early phase
    // flip.  (A remote arrive of the peer's producer per stage -- mbarrier.arrive.release.cluster = MEMBAR + ERRBAR --
    // serialised its loop to one stage per ~1000 cycles: measured, the whole kernel ran at half speed.)
    for (int s = 0; s < L::A_STAGES; ++s) {
      mbar_init(a_full(s), UPS ? 2 : 1);
      mbar_init(a_empty(s), 1);
    }
    for (int s = 0; s < L::B_STAGES; ++s) {
      mbar_init(b_full(s), 1);
      mbar_init(b_empty(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(acc_full(s), 1);
      mbar_init(acc_empty(s), 16);  // one arrive per epilogue warp of BOTH CTAs (leader's barrier)
    }
    for (int s = 0; s < L::P_STAGES; ++s) mbar_init(p_full(s), 1);
    if (UPS)
      for (int s = 0; s < L::A_STAGES; ++s) mbar_init(a_empty_sw(s), 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmOut);
  }
  if (warp == 1) {
    tmem_alloc_pair(smem_u32(const_cast<uint32_t*>(tmem_slot)), L::TMEM_COLS);
    tmem_relinquish_pair();
  }
  // Programmatic dependent launch: the on-chip set-up above overlaps the tail of the previous kernel in the stream (when
  // this launch carries the attribute; otherwise both instructions are no-ops); no global memory is touched before the
  // wait -- the bias below may have been written by the optimizer kernel just before.
#if PDA_PDL_TRIGGER_EARLY
  griddep_launch();
#endif
  griddep_wait();
  if (warp >= 2) {
    for (int i = threadIdx.x - 64; i < p.cout; i += CONV2_THREADS - 64) bias_s[i] = p.bias ? p.bias[i] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's barriers are initialised before anyone signals them remotely
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // UPS: a slab stage is shared by two producers -- the TMA warp (chunks of segment 1) and the epilogue warps (the
  // interpolated chunks of segment 0).  A producer that merely WATCHED the other's uses of a stage go by on one shared
  // "empty" barrier could fall two phases behind (a parity wait cannot tell phase n from n + 2: premature pass) or, if it
  // observed every use, arrive late at a use that had already been consumed and wait for a completion only its own next
  // chunk could cause (dead-lock; seen when the kernel shared the GPU with another stream or a programmatically dependent
  // launch).  So the MMA warp releases a stage on the barrier of the producer that OWNS THE STAGE'S NEXT USE
  // (a_empty: TMA warp, a_empty_sw: epilogue warps): each producer sees exactly one completion per own use -- the plain
  // single-producer protocol, twice.  Chunk g of the CTA's sequence (g = unit index * chunks + ch) uses stage
  // g % A_STAGES and is software-produced iff g % chunks < c0 / 64.
  const int nchunk_sw = UPS ? (p.c0 >> 6) : 0;
  // first_free: stages whose very first use belongs to this producer (no release precedes it)
  auto first_free_mask = [&](bool sw) {
    uint32_t m = 0;
    for (int s2 = 0; s2 < L::A_STAGES; ++s2)
      if (((s2 % chunks) < nchunk_sw) == sw) m |= 1u << s2;
    return m;
  };
  // this CTA's pixel tile of pair-unit u (the peer of an odd tile count re-computes the last tile and discards it)
  auto unit_tile = [&](int u, int& nb, int& img, int& ty, int& tx, bool& ghost) {
    nb = u % n_blocks;
    int mtile = 2 * (u / n_blocks) + (int)rank;
    ghost = mtile >= mtiles;
    if (ghost) mtile = mtiles - 1;
    img = mtile / tiles_per_img;
    const int t = mtile - img * tiles_per_img;
    ty = t / p.tiles_x;
    tx = t - ty * p.tiles_x;
  };

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (both CTAs)
    const bool leader_lane = elect_one();
    int as = 0, bs = 0;
    uint32_t aph = 0, bph = 0;
    uint32_t ups_first = first_free_mask(false), ups_ph = 0;
    if (RES) {
      // this CTA's half of the whole weight matrix of the (single) n-block: 9 taps x 1 chunk, loaded once
      const uint32_t bar = mapa_shared(b_full(0), 0);
      if (leader_lane) {
        if (rank == 0) mbar_expect_tx(b_full(0), 2 * 9 * L::B_BYTES);
        for (int tap = 0; tap < 9; ++tap)
          tma_load_2d_pair(sbase + L::B_OFF + tap * L::B_BYTES, &tmB, bar, tap * ctot, (int)rank * L::BH);
      }
      __syncwarp();
    }
    for (int u = pair; u < units; u += npairs) {
      int nb, img, ty, tx;
      bool ghost;
      unit_tile(u, nb, img, ty, tx, ghost);
      const int x0 = tx * 8, y0 = ty * (16 * MT);
      const int n0 = nb * BN + (int)rank * L::BH;
      for (int ch = 0; ch < chunks; ++ch) {
        const int c = ch << 6;
        for (int kx = 0; kx < 3; ++kx) {
          if ((!WIDE || kx == 0) && UPS && c < p.c0) {
            // produced by the epilogue warps: not this warp's stage use
            if (++as == L::A_STAGES) { as = 0; aph ^= 1; }
          } else if (!WIDE || kx == 0) {
            if (UPS) {
              // own uses of this stage only (see the note at nchunk_sw)
              if ((ups_first >> as) & 1u) {
                ups_first &= ~(1u << as);
              } else {
                mbar_wait(a_empty(as), (ups_ph >> as) & 1u);
                ups_ph ^= 1u << as;
              }
            } else {
              mbar_wait(a_empty(as), aph ^ 1);
            }
            if (leader_lane) {
              const uint32_t bar = mapa_shared(a_full(as), 0);
              if (rank == 0) mbar_expect_tx(a_full(as), 2 * L::A_TX);
              else if (UPS) mbar_arrive_cluster(bar);  // count-2 barrier in this mode
              const uint32_t dst = sbase + L::A_OFF + as * L::A_BYTES;
              const int xs = WIDE ? x0 - 1 : x0 + kx - 1;
              if (c < p.c0)
                tma_load_4d_pair(dst, &tmA0, bar, c, xs, y0 - 1, img);
              else
                tma_load_4d_pair(dst, &tmA1, bar, c - p.c0, xs, y0 - 1, img);
            }
            __syncwarp();
            if (++as == L::A_STAGES) { as = 0; aph ^= 1; }
          }
          if (!RES) {
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
              mbar_wait(b_empty(bs), bph ^ 1);
              if (leader_lane) {
                const uint32_t bar = mapa_shared(b_full(bs), 0);
                if (rank == 0) mbar_expect_tx(b_full(bs), 2 * L::B_BYTES);
                tma_load_2d_pair(sbase + L::B_OFF + bs * L::B_BYTES, &tmB, bar, (ky * 3 + kx) * ctot + c, n0);
              }
              __syncwarp();
              if (++bs == L::B_STAGES) { bs = 0; bph ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer: the leader CTA only
    if (rank == 0) {
      const bool leader_lane = elect_one();
      constexpr uint32_t idesc = F16 ? umma_idesc_f16(256, BN) : umma_idesc_bf16(256, BN);
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      uint32_t it = 0;
      if (RES) {
        mbar_wait(b_full(0), 0);
        tc_fence_after();
      }
      for (int u = pair; u < units; u += npairs, ++it) {
        const uint32_t buf = it & 1;
        mbar_wait(acc_empty(buf), ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t dcol = tmem_base + buf * (MT * BN);
        for (int ch = 0; ch < chunks; ++ch) {
          for (int kx = 0; kx < 3; ++kx) {
            if (!WIDE || kx == 0) {
              mbar_wait(a_full(as), aph);
              tc_fence_after();
            }
            const uint32_t sa = sbase + L::A_OFF + as * L::A_BYTES + (WIDE ? kx * 128 : 0);
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
              uint32_t sb;
              if (RES) {
                sb = sbase + L::B_OFF + (ky * 3 + kx) * L::B_BYTES;
              } else {
                mbar_wait(b_full(bs), bph);
                tc_fence_after();
                sb = sbase + L::B_OFF + bs * L::B_BYTES;
              }
              if (leader_lane) {
                const uint64_t db = umma_desc_k_sw128(sb);
                const uint32_t acc = (ch | kx | ky) != 0 ? 1u : 0u;
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                  const uint32_t arow = sa + (16 * mt + ky) * L::ROW_BYTES;
                  // WIDE: arow is kx * 128 B into a 1024-byte swizzle atom; the rows continue linearly into the next atom
                  // of the same 16-pixel slab row.  The descriptor's base-offset field stays 0 (measured: the swizzle is
                  // applied to absolute shared-memory address bits; with the field set the result is wrong).
                  const uint64_t da = umma_desc_k_sw128(arow, L::ROW_BYTES);
#pragma unroll
                  for (int k = 0; k < 4; ++k)
                    umma_pair(dcol + mt * BN, da + 2 * k, db + 2 * k, idesc, (acc | k) != 0 ? 1u : 0u);
                }
                if (!RES) umma_commit_pair(b_empty(bs));
              }
              __syncwarp();
              if (!RES) {
                if (++bs == L::B_STAGES) { bs = 0; bph ^= 1; }
              }
            }
            if (!WIDE || kx == 2) {
              if (leader_lane) {
                // UPS: the stage goes back to the producer of its next use (chunk ch + A_STAGES of the sequence)
                const bool next_sw = UPS && ((ch + L::A_STAGES) % chunks) < nchunk_sw;
                umma_commit_pair(next_sw ? a_empty_sw(as) : a_empty(as));
              }
              __syncwarp();
              if (++as == L::A_STAGES) { as = 0; aph ^= 1; }
            }
          }
        }
        if (leader_lane) umma_commit_pair(acc_full(buf));
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue: 8 warps, one output pixel per thread
    constexpr int NG = BN / 64;            // 64-channel groups per accumulator
    constexpr int NSUB = MT * NG;          // [128 px x 64 ch] sub-tiles per unit
    const int ew = warp - 2;               // 0..7
    const int wg = ew >> 2;                // warpgroup: takes the sub-tiles with index % 2 == wg
    const int q = warp & 3;                // TMEM lane quarter this warp may read
    const int row = q * 32 + lane;
    const int lty = row >> 3, ltx = row & 7;
    const int Hp = p.H >> 1, Wp = p.W >> 1;
    uint8_t* stage = smem + L::STG_OFF + ew * 4096;
    const uint32_t stage_u32 = sbase + L::STG_OFF + ew * 4096;
    const int sw = lane & 7;
    auto drain = [&](int u, uint32_t it) {
      int nb, img, ty, tx;
      bool ghost;
      unit_tile(u, nb, img, ty, tx, ghost);
      const int n0 = nb * BN;
      const int x = tx * 8 + ltx;
      const uint32_t buf = it & 1;
      const uint32_t acc_empty_leader = mapa_shared(acc_empty(buf), 0);
      mbar_wait(acc_full(buf), (it >> 1) & 1);
      tc_fence_after();
      if (wg >= NSUB) {  // nothing to read for this warpgroup (MT = 1, BN = 64)
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(acc_empty_leader);
      }
#pragma unroll 1
      for (int sub = wg; sub < NSUB; sub += 2) {
        const int mt = sub / NG, g = sub - mt * NG;
        const int y = ty * (16 * MT) + mt * 16 + lty;
        const bool valid = !ghost && (y < p.H) && (x < p.W);
        const bool pool_owner = valid && !(lty & 1) && !(ltx & 1);
        __nv_bfloat16* pool_px =
            p.out_pool
                ? p.out_pool + ((static_cast<size_t>(img) * Hp + (y >> 1)) * Wp + (x >> 1)) * p.cout + n0 + g * 64
                : nullptr;
        if (p.out) {
          // the previous TMA store of this warp has finished reading the staging box
          if (lane == 0) tma_store_wait_read<0>();
          __syncwarp();
        }
#pragma unroll
        for (int cb = 0; cb < 2; ++cb) {
          uint32_t v[32];
          tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * (MT * BN) + mt * BN + g * 64 + cb * 32,
                    v);
          uint4 mk[4];
          if (p.mask) {  // issued before the TMEM wait so that both latencies overlap
            const uint4* mp = reinterpret_cast<const uint4*>(
                p.mask + ((static_cast<size_t>(img) * p.H + y) * p.W + x) * p.cout + n0 + g * 64 + cb * 32);
#pragma unroll
            for (int j = 0; j < 4; ++j) mk[j] = valid ? __ldg(mp + j) : make_uint4(0, 0, 0, 0);
          }
          tmem_ld_wait();
          if (cb == 1 && sub + 2 >= NSUB) {
            // last TMEM read of this warp for this unit: hand the accumulator buffer back to the (leader's) MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(acc_empty_leader);
          }
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float tv = __uint_as_float(v[j]) + bias_s[n0 + g * 64 + cb * 32 + j];
            f[j] = p.relu ? fmaxf(tv, 0.f) : tv;
          }
          if (F16) {
            float vmax = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) vmax = fmaxf(vmax, fabsf(f[j]));
            if (p.range_flag != nullptr && __any_sync(0xffffffffu, valid && !(vmax <= F16_MAX)) && lane == 0)
              atomicOr(p.range_flag, 1);
          }
          if (p.mask) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint32_t w4[4] = {mk[j].x, mk[j].y, mk[j].z, mk[j].w};
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const uint32_t h = (w4[e >> 1] >> ((e & 1) * 16)) & 0xffffu;
                if (h == 0u || h >= 0x8000u) f[8 * j + e] = 0.f;
              }
            }
          }
          if (p.out) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 o;
              o.x = pack_act2<F16>(f[8 * j + 0], f[8 * j + 1]);
              o.y = pack_act2<F16>(f[8 * j + 2], f[8 * j + 3]);
              o.z = pack_act2<F16>(f[8 * j + 4], f[8 * j + 5]);
              o.w = pack_act2<F16>(f[8 * j + 6], f[8 * j + 7]);
              *reinterpret_cast<uint4*>(stage + lane * 128 + (((cb * 4 + j) ^ sw) << 4)) = o;
            }
          }
          if (pool_px) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              float tv = valid ? f[j] : 0.f;
              tv += __shfl_xor_sync(0xffffffffu, tv, 1);
              tv += __shfl_xor_sync(0xffffffffu, tv, 8);
              f[j] = 0.25f * tv;
            }
            if (pool_owner) {
              uint4* dst = reinterpret_cast<uint4*>(pool_px + cb * 32);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                uint4 o;
                o.x = pack_act2<F16>(f[8 * j + 0], f[8 * j + 1]);
                o.y = pack_act2<F16>(f[8 * j + 2], f[8 * j + 3]);
                o.z = pack_act2<F16>(f[8 * j + 4], f[8 * j + 5]);
                o.w = pack_act2<F16>(f[8 * j + 6], f[8 * j + 7]);
                dst[j] = o;
              }
            }
          }
        }
        if (p.out) {
          fence_proxy_async_smem();  // staging written by the generic proxy, read by the TMA engine
          __syncwarp();
          if (lane == 0 && !ghost) {
            tma_store_4d(&tmOut, stage_u32, n0 + g * 64, tx * 8, ty * (16 * MT) + mt * 16 + q * 4, img);
            tma_store_commit();
          }
        }
      }
    };
    if (!UPS) {
      uint32_t it = 0;
      for (int u = pair; u < units; u += npairs, ++it) drain(u, it);
    } else {
      // ---------------------------------------------------------- software producer of the up-sampled K segment
      const int ptid = threadIdx.x - 64;                  // 0 .. 255
      const int h = p.H >> 1, w = p.W >> 1;               // low-resolution source
      const float rh = (p.H > 1) ? (float)(h - 1) / (float)(p.H - 1) : 0.f;
      const float rw = (p.W > 1) ? (float)(w - 1) / (float)(p.W - 1) : 0.f;
      const int nchunk0 = p.c0 >> 6;
      // The low-resolution patch a slab interpolates from is staged in shared memory by TMA (tmA0: unswizzled boxes of
      // PATCH_ROWS x 7 px x 64 channels), two software-produced chunks ahead: gathering straight from global memory left
      // this producer latency-bound (measured: the 192 -> 64 layer at 1024^2 ran at 58 % of the unfused conv's rate).
      // Thread ptid 0 issues the loads: it knows a patch buffer is free when all 256 threads have passed the named barrier
      // that ends a chunk, so the ring needs no empty barriers.
      auto issue_patch = [&](uint32_t j) {                // j = index in this CTA's sequence of software-produced chunks
        const uint32_t it2 = j / (uint32_t)nchunk0;
        const int ch2 = (int)(j - it2 * (uint32_t)nchunk0);
        const long long u2 = (long long)pair + (long long)it2 * npairs;
        if (u2 >= units) return;
        int nb, img, ty, tx;
        bool ghost;
        unit_tile((int)u2, nb, img, ty, tx, ghost);
        const int xa = max(tx * 8 - 1, 0), ya = max(ty * (16 * MT) - 1, 0);
        const uint32_t ps = j & 1;
        mbar_expect_tx(p_full(ps), L::PATCH_BYTES);
        tma_load_4d(sbase + L::P_OFF + ps * L::PATCH_BYTES, &tmA0, p_full(ps), ch2 << 6, (int)(rw * (float)xa),
                    (int)(rh * (float)ya), img);
      };
      if (ptid == 0) {
        issue_patch(0);
        issue_patch(1);
      }
      constexpr int RSEG = (L::SLAB_ROWS + 2) / 3;         // slab rows per thread: 3 row segments x 80 columns = 240 threads
      const int col_seg = ptid / 80, col_rem = ptid - col_seg * 80;
      const int col_px = col_rem >> 3, col_g = col_rem & 7;
      const int col_rows_begin = col_seg * RSEG;          // threads 240..255: past the last row, no items
      int as = 0;
      uint32_t aph = 0;
      uint32_t sw_first = first_free_mask(true), sw_ph = 0;
      uint32_t it = 0, j = 0;
      int prev_u = -1;
      for (int u = pair; u < units; u += npairs, ++it) {
        int nb, img, ty, tx;
        bool ghost;
        unit_tile(u, nb, img, ty, tx, ghost);
        const int x0 = tx * 8 - 1, y0 = ty * (16 * MT) - 1;   // image coordinates of slab (row 0, px 0)
        const int x_lo = (int)(rw * (float)max(x0, 0)), y_lo = (int)(rh * (float)max(y0, 0));  // patch origin
        // this thread's pixel column of the slab
        const int col_x = x0 + col_px;
        const bool x_in = col_x >= 0 && col_x < p.W;
        const float sx = rw * (float)col_x;
        const int x1 = x_in ? (int)sx : 0;
        const float lx1 = sx - (float)x1, lx0 = 1.f - lx1;
        const int x_step = (x1 < w - 1) ? 128 : 0;
        const int col_off = (min(max(x1 - x_lo, 0), L::PATCH_PX - 2) << 7) + (col_g << 4);
        for (int ch = 0; ch < nchunk0; ++ch, ++j) {
          const uint32_t ps = j & 1;
          mbar_wait(p_full(ps), (j >> 1) & 1);
          // own uses of this stage only (see the note at nchunk_sw): the stage's first use needs no release
          if ((sw_first >> as) & 1u) {
            sw_first &= ~(1u << as);
          } else {
            mbar_wait(a_empty_sw(as), (sw_ph >> as) & 1u);
            sw_ph ^= 1u << as;
          }
          uint8_t* slab = smem + L::A_OFF + as * L::A_BYTES;
          const uint8_t* patch = smem + L::P_OFF + ps * L::PATCH_BYTES;
          // A thread owns one (pixel column px of the 10 the taps read, 8-channel group g) and a run of consecutive slab
          // rows: the column's horizontal blend is loop-invariant, and the horizontally blended source rows are carried
          // from one slab row to the next (the source row advances by 0 or 1 per output row), so a 16-byte item costs
          // about half a source-row blend plus the vertical blend.  Same fp32 operations in the same order as
          // upsample2x_kernel (misc_kernels.cu): bit-identical to up-sampling first.
          if (col_rows_begin < L::SLAB_ROWS) {
            const uint8_t* pcol = patch + col_off;
            auto hblend = [&](int ry, float (&hrow)[8]) {
              const uint4 qa = *reinterpret_cast<const uint4*>(pcol + ry * (L::PATCH_PX << 7));
              const uint4 qb = *reinterpret_cast<const uint4*>(pcol + ry * (L::PATCH_PX << 7) + x_step);
              const uint32_t a4[4] = {qa.x, qa.y, qa.z, qa.w}, b4[4] = {qb.x, qb.y, qb.z, qb.w};
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const float2 va = unpack_act2<F16>(a4[k]), vb = unpack_act2<F16>(b4[k]);
                hrow[2 * k] = blend2(lx0, lx1, va.x, vb.x);
                hrow[2 * k + 1] = blend2(lx0, lx1, va.y, vb.y);
              }
            };
            float h0[8], h1[8];
            int cached_y1 = -2;
            const int r_end = min(col_rows_begin + RSEG, L::SLAB_ROWS);
            for (int r = col_rows_begin; r < r_end; ++r) {
              const int y = y0 + r;
              uint4 o = make_uint4(0u, 0u, 0u, 0u);             // outside the image: the conv's zero padding
              if (x_in && y >= 0 && y < p.H) {
                const float sy = rh * (float)y;
                const int y1 = (int)sy;
                const float l1 = sy - (float)y1, l0 = 1.f - l1;
                if (y1 != cached_y1) {
                  const int ry = min(y1 - y_lo, L::PATCH_ROWS - 2);
                  if (y1 == cached_y1 + 1) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) h0[k] = h1[k];
                  } else {
                    hblend(ry, h0);
                  }
                  if (y1 < h - 1) {
                    hblend(ry + 1, h1);
                  } else {
#pragma unroll
                    for (int k = 0; k < 8; ++k) h1[k] = h0[k];
                  }
                  cached_y1 = y1;
                }
                uint32_t o4[4];
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  o4[k] = pack_act2<F16>(blend2(l0, l1, h0[2 * k], h1[2 * k]), blend2(l0, l1, h0[2 * k + 1], h1[2 * k + 1]));
                o = make_uint4(o4[0], o4[1], o4[2], o4[3]);
              }
              // SWIZZLE_128B as TMA writes it: 16-byte chunk index XOR address bits 7..9 (the stage base is 1024-aligned)
              const int off = r * L::ROW_BYTES + col_px * 128;
              *reinterpret_cast<uint4*>(slab + off + ((col_g ^ ((off >> 7) & 7)) << 4)) = o;
            }
          }
          fence_proxy_async_smem();        // generic-proxy stores -> visible to the tensor core (async proxy)
          named_bar2(2, 256);
          if (ptid == 0) {
            if (rank == 0) mbar_arrive(a_full(as));
            else mbar_arrive_cluster(mapa_shared(a_full(as), 0));
            issue_patch(j + 2);            // into the patch buffer every thread has just finished reading
          }
          if (++as == L::A_STAGES) { as = 0; aph ^= 1; }
        }
        if (prev_u >= 0) drain(prev_u, it - 1);
        prev_u = u;
        // the chunks the TMA warp loads are not this producer's stage uses: only the stage index moves on
        for (int ch = nchunk0; ch < chunks; ++ch) {
          if (++as == L::A_STAGES) { as = 0; aph ^= 1; }
        }
      }
      if (prev_u >= 0) drain(prev_u, it - 1);
    }
    if (lane == 0) tma_store_wait<0>();  // shared memory must outlive the last store
  }

#if !PDA_PDL_TRIGGER_EARLY
  // this CTA's work is done: the next kernel of the stream may be scheduled onto the SMs that free up while the slowest
  // CTAs of this grid finish (it still waits for this grid's completion before touching memory).  (Triggering at the START
  // of the kernel is what first exposed the slab-ring protocol bug of the fused up-sampling mode -- see the note at
  // nchunk_sw; with the trigger here at most two grids overlap.)
  griddep_launch();
#endif
  // both CTAs are done with each other's shared memory / barriers and with their tensor memory
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc_pair(tmem_base, L::TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
template <int BN, int MT, bool RES, bool F16, bool WIDE, bool UPS = false>
static int launch_conv2_fmt(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b, const CUtensorMap& o,
                            const ConvArgs& args, cudaStream_t stream) {
  using L = Conv2Cfg<BN, MT, RES, WIDE, UPS>;
  auto kern = conv3x3_tc2_kernel<BN, MT, RES, F16, WIDE, UPS>;
  static int configured[64];
  static int max_clusters[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return PDA_ERR_CUDA;
  if (dyn_smem_attr_needed(configured, L::DYN_BYTES)) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::DYN_BYTES) != cudaSuccess)
      return PDA_ERR_CUDA;
    max_clusters[dev] = 0;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(CONV2_THREADS);
  cfg.dynamicSmemBytes = L::DYN_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (max_clusters[dev] == 0) {
    // how many CTA pairs can be resident at once (74 on a full B200: one per TPC)
    cfg.gridDim = dim3(148);
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n <= 0) return PDA_ERR_CUDA;
    max_clusters[dev] = n > 74 ? 74 : n;
    if (getenv("PDA_DEBUG")) fprintf(stderr, "[pda] conv3x3_tc2<%d,%d,%d,%d>: max active clusters %d\n", BN, MT, (int)RES, (int)F16, n);
  }
  const long long mtiles = (long long)args.tiles_x * args.tiles_y * args.B;
  const long long units = ((mtiles + 1) / 2) * (args.cout / BN);
  if (units > 0x3fffffffLL) return PDA_ERR_SHAPE;
  int cap = max_clusters[dev];
  if (cap > sm_budget(0) / 2) cap = sm_budget(0) / 2;
  const int pairs = (int)(units < cap ? units : cap);
  cfg.gridDim = dim3(2 * pairs);
  {
    // programmatic dependent launch, OFF by default (PDA_PDL=1 switches it on): this kernel's CTAs may be scheduled while
    // the previous kernel of the stream drains; the kernel itself waits (griddepcontrol.wait) before it touches global
    // memory.  Measured on the inference and the training step (same box, two passes each): no difference outside the
    // run-to-run noise of +-1 % -- the CUDA-graph replay already hides the launch latency and the prologue is ~1 us.
    static const int pdl = [] {
      const char* e = getenv("PDA_PDL");
      return e ? atoi(e) : 0;
    }();
    if (pdl) {
      attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[1].val.programmaticStreamSerializationAllowed = 1;
      cfg.numAttrs = 2;
    }
  }
  PDA_COUNT(1);
  if (cudaLaunchKernelEx(&cfg, kern, a0, a1, b, o, args) != cudaSuccess) return PDA_ERR_CUDA;
  return cudaGetLastError() == cudaSuccess ? PDA_OK : PDA_ERR_CUDA;
}

template <int BN, int MT, bool RES>
static int launch_conv2(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b, const CUtensorMap& o,
                        const ConvArgs& args, cudaStream_t stream) {
  if (args.up_src != nullptr) {
    if (RES) return PDA_ERR_SHAPE;  // (an up-sampled segment never meets the resident-weight configuration)
    return args.act_f16 ? launch_conv2_fmt<BN, MT, false, true, true, true>(a0, a1, b, o, args, stream)
                        : launch_conv2_fmt<BN, MT, false, false, true, true>(a0, a1, b, o, args, stream);
  }
  if (args.wide)
    return args.act_f16 ? launch_conv2_fmt<BN, MT, RES, true, true>(a0, a1, b, o, args, stream)
                        : launch_conv2_fmt<BN, MT, RES, false, true>(a0, a1, b, o, args, stream);
  return args.act_f16 ? launch_conv2_fmt<BN, MT, RES, true, false>(a0, a1, b, o, args, stream)
                      : launch_conv2_fmt<BN, MT, RES, false, false>(a0, a1, b, o, args, stream);
}

// same contract as conv3x3_tc (csrc/conv3x3_tc.cu); selected by it for shapes with at least two pixel tiles
int conv3x3_tc2(const void* src0, int c0, const void* src1, int c1, const void* wpacked, const float* bias, void* out,
                void* out_pool, const void* mask, int B, int H, int W, int cout, int relu, int bn_override,
                int act_f16, int* range_flag, cudaStream_t stream, const void* up_src) {
  // up_src != nullptr: segment 0 = bilinear x2 of up_src [B][H/2][W/2][c0] (src0 is ignored)
  if (up_src != nullptr && ((H & 1) || (W & 1) || c1 <= 0 || !src1)) return PDA_ERR_SHAPE;
  if (c0 <= 0 || (c0 & 63) || (c1 & 63) || (cout & 63) || cout > 512 || B <= 0 || H <= 0 || W <= 0)
    return PDA_ERR_SHAPE;
  if (out_pool && ((H & 1) || (W & 1))) return PDA_ERR_SHAPE;
  int bn = (bn_override == 64 || bn_override == 128 || bn_override == 256)
               ? bn_override
               : ((cout % 256 == 0) ? 256 : (cout % 128 == 0) ? 128 : 64);
  if (cout % bn) return PDA_ERR_SHAPE;
  const int mt = (H > 16 && bn != 256) ? 2 : 1;
  ConvArgs a;
  a.B = B; a.H = H; a.W = W; a.c0 = c0; a.c1 = c1; a.cout = cout; a.relu = relu;
  a.tile_w = 8;
  a.tile_h = 16 * mt;
  a.tiles_x = (W + 7) / 8;
  a.tiles_y = (H + a.tile_h - 1) / a.tile_h;
  a.bias = bias;
  a.out = static_cast<__nv_bfloat16*>(out);
  a.out_pool = static_cast<__nv_bfloat16*>(out_pool);
  a.mask = static_cast<const __nv_bfloat16*>(mask);
  a.act_f16 = act_f16;
  a.range_flag = act_f16 ? range_flag : nullptr;
  {
    static const int wide_mode = [] {
      const char* e = getenv("PDA_CONV_WIDE");  // 0: three 8-px slabs per chunk; 1: one 16-px slab per chunk
      return e ? atoi(e) : PDA_CONV_WIDE_DEFAULT;
    }();
    // one 10-pixel slab per chunk (default) except for the single-chunk 64 -> 128 layer, the one shape that measured
    // slower with it (profiles/r02_conv_wide_slab.md)
    a.wide = (wide_mode != 0 && !(c0 + c1 == 64 && cout != 64)) || up_src != nullptr;
    a.wide_base_offset = 0;
  }
  a.up_src = up_src;
  const int box_w = a.wide ? PDA_UPS_ROW_PX : 8;
  CUtensorMap tA0, tA1, tB;
  int r;
  if (up_src == nullptr) {
    r = make_act_tensor_map(&tA0, src0, B, H, W, c0, box_w, a.tile_h + 2, 64);
    if (r) return r;
  }
  if (c1 > 0) {
    r = make_act_tensor_map(&tA1, src1, B, H, W, c1, box_w, a.tile_h + 2, 64);
    if (r) return r;
    if (up_src != nullptr) {
      // the low-resolution source, read as unswizzled patches of (8 * mt + 3) rows x 7 px x 64 channels
      r = make_act_tensor_map(&tA0, up_src, B, H / 2, W / 2, c0, 7, 8 * mt + 3, 64, /*swizzle128=*/0);
      if (r) return r;
    }
  } else {
    tA1 = tA0;
  }
  r = make_mat_tensor_map(&tB, wpacked, 9LL * (c0 + c1), cout, 64, bn / 2);  // one CTA loads half the rows
  if (r) return r;
  CUtensorMap tO = tA0;  // unused when out == nullptr
  if (out) {
    r = make_act_tensor_map(&tO, out, B, H, W, cout, 8, 4, 64);
    if (r) return r;
  }
  const bool resident = (bn == 64 && cout == 64 && c0 + c1 == 64 && mt == 2 && up_src == nullptr);
  if (mt == 2) {
    if (bn == 128) return launch_conv2<128, 2, false>(tA0, tA1, tB, tO, a, stream);
    return resident ? launch_conv2<64, 2, true>(tA0, tA1, tB, tO, a, stream)
                    : launch_conv2<64, 2, false>(tA0, tA1, tB, tO, a, stream);
  }
  if (bn == 256) return launch_conv2<256, 1, false>(tA0, tA1, tB, tO, a, stream);
  if (bn == 128) return launch_conv2<128, 1, false>(tA0, tA1, tB, tO, a, stream);
  return launch_conv2<64, 1, false>(tA0, tA1, tB, tO, a, stream);
}

}  // namespace pda
