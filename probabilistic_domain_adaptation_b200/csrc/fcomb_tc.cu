// Tensor-core version of the fused Fcomb + sigmoid + cross-sample mean + consensus kernel.
//
// Reference: S x Fcomb.forward (/root/reference/prob_utils/my_models/probabilistic_unet.py:200-214) + the consensus
// arithmetic of prob_utils/my_trainer/mean_teacher_trainer.py:74-86.
//
// The kernel is COMPUTE bound (~1000 FLOP/B at S=16), so the two 64x64 layers run on tcgen05:
//   per 128-pixel tile:  H1 = F . (W1f_hi + W1f_lo)^T         128x64x64 MMA x2 (bf16 hi/lo split of the fp32 weights
//                                                            -> ~16-bit mantissa), accumulator in TMEM, read ONCE
//   per sample s:        A1_s = relu(H1 + bz_s) in packed fp16 (H1 and bz_s rounded to fp16 once; one HFMA2.RELU per
//                        two channels) -> tcgen05.st -> TMEM: the A operand of the next MMA never touches shared
//                        memory (no swizzled stores, no generic->async proxy fence)              (producer warps)
//                        H2_s = [A1_s | 1 1 0..] . [W2 | b2_hi b2_lo 0..]^T   128x64x80 fp16 MMA, A from TMEM for the
//                        first 64 k, the bias block (ones) from shared memory
//                        A3_s = relu(H2_s) rounded to packed fp16 (ONE cvt.rn.relu.f16x2 per two channels), written
//                        back over the first 32 columns of the H2 accumulator it came from (FC_L3_MMA, round 2)
//                        D3_s = A3_s . [w3_hi | w3_lo | 0..]^T   128x16x64 fp16 MMA, A from TMEM, D into columns 32..47
//                        of the same accumulator buffer; logit = D3[0] + D3[1] + b3
//                        p = sigmoid ; mean / consensus                                       (epilogue warps)
//                        (FC_L3_MMA = 0: the round-1 epilogue, logit = w3 . relu(H2_s) + b3 as 64 FMNMX + 64 FFMA per
//                        pixel-sample in fp32)
// bz_s = b1 + W1z . z_s is the per-(sample, image) bias that replaces the tiled-z concat (fcomb_bz_kernel).
//
// Warp-specialised persistent CTA (2 per SM): warps 0-3 = producers (own H1 in registers, write the A1 ring),
// warps 4-7 = epilogue (TMEM -> relu -> packed fp16 -> TMEM, later the two last-layer columns -> sigmoid / counters),
// warp 8 = control (TMA of the feature tile, all tcgen05.mma).  mbarrier pipelines for the F tile, the A1 ring of 2 and
// the H2 accumulator ring of 3 (all in TMEM; H1 shares its columns with the A1 ring) let the three roles run
// concurrently; nothing but the outputs leaves the SM.
// Measured: 1.62 ms -> 0.99 ms at 4 x 1024^2 px, S = 16 against the version that staged A1 in shared memory with fp32
// adds (profiles/r01e_fcomb_variants.md); the last layer as an MMA another 8 %; what paces the kernel now is the
// epilogue warps' per-sample chain of one tensor-memory load and one store round trip (~700 cycles, one warp per TMEM
// lane quarter): profiles/r02_fcomb_l3_variants.md lists every experiment with same-box A/B numbers.
#include <cuda_fp16.h>

#include "conv.cuh"
#include "ptx.cuh"

namespace pda {

constexpr int FCT = 64;             // feature / hidden channels
constexpr int FC_TILE = 128;        // pixels per tile == TMEM lanes
// Tensor memory per CTA (256 columns, two CTAs per SM): H2 accumulator ring of FC_H2_RING x 64 columns from 0; A1 operand
// ring of FC_A1_RING x 32 columns (fp16 x 2 per column) behind it; H1 (64 columns).
//   FC_H2_RING = 2 (default): H1 in its own 64 columns, its MMA prefetched in the middle of the previous tile.
//   FC_H2_RING = 3: three accumulator buffers (the epilogue converts sample s while the last-layer MMA of s - 1 is in
//   flight and the second-layer MMA of s + 1 already has a free buffer).  3 x 64 + 2 x 32 = 256 leaves no columns for H1:
//   it ALIASES the A1 ring (FC_H1_ALIAS) -- the layer-1 MMA of the next tile runs when the tile's last second-layer MMA
//   has read the ring, the producers copy H1 to registers and only then write A1 rows again.  Correct (same parity
//   numbers) and exactly as fast as two buffers (1.11-1.17 ms against 1.12-1.15 ms, same box): kept as an option.
// What was measured about this kernel's bound is collected in profiles/r02_fcomb_l3_variants.md.
#ifndef FC_H2_RING
#define FC_H2_RING 3
#endif
constexpr int FC_A1_RING = 2;
constexpr bool FC_H1_ALIAS = FC_H2_RING == 3;
static_assert(FC_H2_RING == 2 || FC_H2_RING == 3, "supported accumulator ring depths");
constexpr int FC_TMEM_COLS = 256;
constexpr int FC_CTAS_PER_SM = 2;
constexpr int FC_H2_COL = 0;
constexpr int FC_A1_COL = FC_H2_RING * FCT;
constexpr int FC_H1_COL = FC_H1_ALIAS ? FC_A1_COL : FC_A1_COL + FC_A1_RING * 32;
static_assert(FC_H1_COL + FCT <= FC_TMEM_COLS && FC_A1_COL + FC_A1_RING * 32 <= FC_TMEM_COLS, "tensor memory budget");
constexpr int FC_PROD_WARPS = 4;
constexpr int FC_EPI_WARPS = 4;
constexpr int FC_CTRL_WARP = FC_PROD_WARPS + FC_EPI_WARPS;
constexpr int FC_THREADS = 32 * (FC_CTRL_WARP + 1);

// |H1| and |bz| below this bound cannot overflow the packed-fp16 add relu(H1 + bz) (2 x 32000 < 65504); a launch that
// sees a larger value raises the overflow flag and the caller's stream re-runs the batch through the exact fp32 kernel
constexpr float FC_F16_SAFE = 32000.f;

// last layer w3[64]: the first FC_W3_REG entries live in the epilogue threads' registers, the rest is read from shared
// memory as warp-wide broadcast LDS.128 (measured at 4 x 1024^2, S = 16: all 64 from shared memory 1.03 ms, 48 in registers 1.08 ms, 64 in registers -- which spills under the 96-register cap of 2 CTAs/SM -- 1.10 ms; profiles/r02_fcomb_w3_variants.md)
#ifndef FC_W3_REG
#define FC_W3_REG 0
#endif
// 1: the last layer (64 -> 1) runs on the tensor core as well (see the header); 0: fp32 FMA epilogue
#ifndef FC_L3_MMA
#define FC_L3_MMA 1
#endif

// relu(H2) is rounded to fp16 in that mode: a launch whose bound on |H2| (max |A1| x max row-L1 of W2 + max |b2|, from
// fcomb_bz_kernel and the per-tile H1 maximum) is not below this raises the overflow flag -> exact fp32 re-run
constexpr float FC_H2_SAFE = 60000.f;

// -DFC_PROFILE: cycle counters of the three roles (CTA 0, lane 0 of one warp per role), read with pda_fcomb_profile_read
#ifdef FC_PROFILE
__device__ unsigned long long fc_prof[16];
// (only CTA 0 / lane 0 reads the clock: clock reads in every thread of every CTA slowed the kernel down 3x)
#define FC_T0() long long _t0 = (blockIdx.x == 0 && lane == 0) ? clock64() : 0
#define FC_ACC(i) do { if (blockIdx.x == 0 && lane == 0) { long long _t1 = clock64(); fc_prof[i] += (unsigned long long)(_t1 - _t0); _t0 = _t1; } } while (0)
#else
#define FC_T0() do {} while (0)
#define FC_ACC(i) do {} while (0)
#endif

struct FcombSmem {
  static constexpr int A_BYTES = FC_TILE * 128;                 // 128 rows x 64 x 2 B
  static constexpr int W_BYTES = FCT * 128;                     // 64 rows x 64 x 2 B
  static constexpr int F_OFF = 0;                               // feature tile (TMA)
  static constexpr int W1_OFF = A_BYTES;                        // bf16 hi part of W1[:, :64]
  static constexpr int W1L_OFF = W1_OFF + W_BYTES;              // bf16 lo part (w - hi)
  static constexpr int W2_OFF = W1L_OFF + W_BYTES;              // fp16 W2
  static constexpr int W2X_OFF = W2_OFF + W_BYTES;              // fp16 K-extension: col 0 = b2_hi, col 1 = b2_lo
  static constexpr int AX_OFF = W2X_OFF + W_BYTES;              // 8 rows x 128 B: cols 0,1 = 1.0 (aliased by all rows)
  static constexpr int W3T_OFF = AX_OFF + 1024;                 // fp16 last-layer tile: row 0 = w3 hi, row 1 = w3 lo, 16 rows
  static constexpr int BAR_OFF = W3T_OFF + 2048;
  static constexpr int NBARS = 4 + 2 * FC_A1_RING + 4 * FC_H2_RING;
  static constexpr int SLOT_OFF = BAR_OFF + NBARS * 8;
  static constexpr int BZ_CHUNK = 64;                           // samples whose layer-1 bias is staged at a time
  static constexpr int W3_OFF = SLOT_OFF + 16;                  // w3[64] fp32 (per-launch copy)
  static constexpr int BZ_OFF = W3_OFF + FCT * 4;               // bz[BZ_CHUNK][64] fp16
  static int bytes(int) { return BZ_OFF + BZ_CHUNK * FCT * 2 + 1024; }
};

__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
// relu(a + c) on packed halves in one instruction (fma.relu with b = 1.0)
__device__ __forceinline__ uint32_t add_relu_f16x2(uint32_t a, uint32_t c) {
  uint32_t r;
  asm("fma.rn.relu.f16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(0x3C003C00u), "r"(c));
  return r;
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
// D[tmem] (+)= A[tmem] * B[smem], kind::f16: the A operand (128 rows = lanes, K 16-bit elements packed two per
// 32-bit column, i.e. 8 columns per K = 16 step) is read from tensor memory
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// relu + round two fp32 to packed fp16 (lo = even k) in one F2FP
__device__ __forceinline__ uint32_t pack_relu_f16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ void tmem_ld2(uint32_t taddr, uint32_t (&v)[2]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(v[0]), "=r"(v[1]) : "r"(taddr));
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// fp32 [64][ld] weights (first 64 columns) -> K-major SWIZZLE_128B operand tiles in shared memory:
// hi + lo (w ~= hi + lo; bf16 pair, or fp16 pair when f16) when dst_lo != nullptr, else a single fp16 tile.
__device__ __forceinline__ void stage_weight_sw128(uint8_t* dst, uint8_t* dst_lo, const float* __restrict__ w, int ld,
                                                   bool f16 = false) {
  for (int i = threadIdx.x; i < FCT * 8; i += blockDim.x) {
    const int n = i >> 3, c = i & 7;  // row n, 16-byte chunk c (8 k-values)
    const float* src = w + n * ld + c * 8;
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float a = src[2 * j], b = src[2 * j + 1];
      if (dst_lo && f16) {
        const float ah = __half2float(__float2half_rn(a)), bh = __half2float(__float2half_rn(b));
        hi[j] = pack_f16x2(a, b);
        lo[j] = pack_f16x2(a - ah, b - bh);
      } else if (dst_lo) {
        const __nv_bfloat16 ah = __float2bfloat16(a), bh = __float2bfloat16(b);
        hi[j] = pack_bf16x2(a, b);
        lo[j] = pack_bf16x2(a - __bfloat162float(ah), b - __bfloat162float(bh));
      } else {
        hi[j] = pack_f16x2(a, b);
      }
    }
    const int off = n * 128 + ((c ^ (n & 7)) << 4);
    *reinterpret_cast<uint4*>(dst + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    if (dst_lo) *reinterpret_cast<uint4*>(dst_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

__global__ void __launch_bounds__(FC_THREADS, FC_CTAS_PER_SM)
fcomb_tc_kernel(const __grid_constant__ CUtensorMap tmF, const float* __restrict__ bzg, const float* __restrict__ w1,
                const float* __restrict__ w2, const float* __restrict__ b2, const float* __restrict__ w3,
                const float* __restrict__ b3, int* __restrict__ oflag, const float* __restrict__ bounds, int feat_f16,
                int P, int S, int L, int B,
                int tiles_per_img, int num_tiles, float upper, float lower, float* __restrict__ mean_prob,
                float* __restrict__ cons_weight, int64_t* __restrict__ cons_mask, float* __restrict__ logits,
                float* __restrict__ probs) {
  using M = FcombSmem;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar0 = sbase + M::BAR_OFF;
  const uint32_t f_full = bar0, f_empty = bar0 + 8, h1_full = bar0 + 16, h1_empty = bar0 + 24;
  // A1 ring slot i = sample counter % FC_A1_RING; accumulator buffer i = sample counter % FC_H2_RING
  auto a1_full = [&](int i) { return bar0 + 32 + 8u * i; };                                        // A1 written (4 producer warps)
  auto a1_empty = [&](int i) { return bar0 + 32 + 8u * (FC_A1_RING + i); };                         // second-layer MMA has read it
  auto h2_full = [&](int i) { return bar0 + 32 + 8u * (2 * FC_A1_RING + i); };                      // second-layer MMA done
  auto h2_empty = [&](int i) { return bar0 + 32 + 8u * (2 * FC_A1_RING + FC_H2_RING + i); };        // epilogue is done with the buffer
  auto a3_full = [&](int i) { return bar0 + 32 + 8u * (2 * FC_A1_RING + 2 * FC_H2_RING + i); };     // relu(H2) written back (4 warps)
  auto d3_full = [&](int i) { return bar0 + 32 + 8u * (2 * FC_A1_RING + 3 * FC_H2_RING + i); };     // last-layer MMA done
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + M::SLOT_OFF);
  float* bzs = reinterpret_cast<float*>(smem + M::BZ_OFF);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int kin = FCT + L;

  if (tid == 0) {
    mbar_init(f_full, 1);
    mbar_init(f_empty, 1);
    mbar_init(h1_full, 1);
    mbar_init(h1_empty, FC_PROD_WARPS);
    for (int i = 0; i < FC_A1_RING; ++i) {
      mbar_init(a1_full(i), 4);
      mbar_init(a1_empty(i), 1);
    }
    for (int i = 0; i < FC_H2_RING; ++i) {
      mbar_init(h2_full(i), 1);
      mbar_init(h2_empty(i), 4);
      mbar_init(a3_full(i), 4);
      mbar_init(d3_full(i), 1);
    }
    fence_mbar_init();
    tma_prefetch_desc(&tmF);
  }
  if (warp == FC_CTRL_WARP) {
    tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), FC_TMEM_COLS);
    tmem_relinquish();
  }
  stage_weight_sw128(smem + M::W1_OFF, smem + M::W1L_OFF, w1, kin, feat_f16 != 0);  // same format as the features
  stage_weight_sw128(smem + M::W2_OFF, nullptr, w2, FCT);
  // K-extension tiles: B rows n carry (b2_hi, b2_lo) in k = 0, 1; the A rows carry (1, 1).  Only the first 16 k
  // (two 16-byte chunks) of each 128-byte row are read by the K = 16 MMA.
  for (int i = tid; i < FCT * 8; i += blockDim.x) {
    const int n = i >> 3, c = i & 7;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (c == 0) {
      const float b = b2[n];
      const float bh = __half2float(__float2half_rn(b));
      v.x = pack_f16x2(bh, b - bh);
    }
    *reinterpret_cast<uint4*>(smem + M::W2X_OFF + n * 128 + ((c ^ (n & 7)) << 4)) = v;
  }
  if (tid >= 64 && tid < 64 + FCT) reinterpret_cast<float*>(smem + M::W3_OFF)[tid - 64] = w3[tid - 64];
  if (tid >= 128 && tid < 128 + 16 * 8) {
    // last-layer operand tile [16 n][64 k]: n = 0 -> fp16(w3), n = 1 -> fp16(w3 - fp16(w3)), other rows 0
    const int i = tid - 128, n = i >> 3, c = i & 7;
    uint32_t h[4] = {0u, 0u, 0u, 0u};
    if (n < 2) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float a = w3[c * 8 + 2 * j], b = w3[c * 8 + 2 * j + 1];
        const float ah = __half2float(__float2half_rn(a)), bh = __half2float(__float2half_rn(b));
        h[j] = n == 0 ? pack_f16x2(a, b) : pack_f16x2(a - ah, b - bh);
      }
    }
    *reinterpret_cast<uint4*>(smem + M::W3T_OFF + n * 128 + ((c ^ (n & 7)) << 4)) = make_uint4(h[0], h[1], h[2], h[3]);
  }
  if (tid < 64) {
    const int n = tid >> 3, c = tid & 7;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (c == 0) v.x = 0x3C003C00u;  // (1.0h, 1.0h)
    *reinterpret_cast<uint4*>(smem + M::AX_OFF + n * 128 + ((c ^ (n & 7)) << 4)) = v;
  }
  fence_proxy_async_smem();  // operand tiles were written by the generic proxy, read by the tensor core
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp < FC_PROD_WARPS) {
    // ================================================================ producers (lane quarter = warp)
    const int ptid = tid;  // 0 .. 255
    const uint32_t lane_addr = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    uint32_t a_it = 0, t_it = 0;
    int cur_b = -1, cur_chunk = -1;
    const float bz_max = __ldg(bounds + 0), w2_rowsum = __ldg(bounds + 1), b2_max = __ldg(bounds + 2);
    // stages bz[s0 .. s0 + BZ_CHUNK) of image b (S <= BZ_CHUNK: once per image; else once per chunk of samples)
    auto stage_bz = [&](int b, int chunk) {
      named_bar_sync(1, 32 * FC_PROD_WARPS);  // every producer is done with the previous contents
      const int s0 = chunk * M::BZ_CHUNK;
      const int ns = min(M::BZ_CHUNK, S - s0);
      for (int i = ptid; i < ns * (FCT / 4); i += 32 * FC_PROD_WARPS) {
        const int s = i / (FCT / 4), j4 = i - s * (FCT / 4);
        const float4 v = __ldg(reinterpret_cast<const float4*>(bzg + (static_cast<size_t>(s0 + s) * B + b) * FCT) + j4);
        reinterpret_cast<uint2*>(bzs)[i] = make_uint2(pack_f16x2(v.x, v.y), pack_f16x2(v.z, v.w));
      }
      named_bar_sync(1, 32 * FC_PROD_WARPS);
      cur_b = b;
      cur_chunk = chunk;
    };
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++t_it) {
      const int b = tile / tiles_per_img;
      if (b != cur_b || cur_chunk != 0) stage_bz(b, 0);
      // ---- H1 -> registers (kept for all samples) as 32 packed f16x2 (saturating)
      mbar_wait(h1_full, t_it & 1);
      tc_fence_after();
      uint32_t h1[FCT / 2];
      float hmax = 0.f;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t v[32];
        tmem_ld32(lane_addr + FC_H1_COL + half * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          h1[16 * half + i] = pack_f16x2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
          hmax = fmaxf(hmax, fmaxf(fabsf(__uint_as_float(v[2 * i])), fabsf(__uint_as_float(v[2 * i + 1]))));
        }
      }
      tc_fence_before();
      // fp16 range guards (once per tile): "not below the bound" also catches NaN
      //   |H1|, |bz| < FC_F16_SAFE: relu(H1 + bz) cannot overflow
      //   |H2| <= (max |H1| + max |bz|) * max_c sum_k |W2[c][k]| + max |b2| < FC_H2_SAFE   (bounds[]: fcomb_bz_kernel)
      const bool unsafe = !(hmax < FC_F16_SAFE) || (FC_L3_MMA && !((hmax + bz_max) * w2_rowsum + b2_max < FC_H2_SAFE));
      if (__any_sync(0xffffffffu, unsafe) && lane == 0) atomicOr(oflag, 1);
      __syncwarp();
      if (lane == 0) mbar_arrive(h1_empty);
      for (int s = 0; s < S; ++s, ++a_it) {
        if (s / M::BZ_CHUNK != cur_chunk) stage_bz(b, s / M::BZ_CHUNK);
        const uint32_t slot = a_it % FC_A1_RING;
        FC_T0();
        mbar_wait(a1_empty(slot), ((a_it / FC_A1_RING) & 1) ^ 1);  // the MMA that read this slot has completed
        tc_fence_after();
        if (warp == 0) FC_ACC(0);
        // A1_s = relu(H1 + bz_s): this thread's row of 64 halves = 32 TMEM columns of its lane
        const uint32_t bz_addr = sbase + M::BZ_OFF + (s % M::BZ_CHUNK) * FCT * 2;
        uint32_t a[32];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint4 bb = lds128(bz_addr + 16 * c);
          a[4 * c + 0] = add_relu_f16x2(h1[4 * c + 0], bb.x);
          a[4 * c + 1] = add_relu_f16x2(h1[4 * c + 1], bb.y);
          a[4 * c + 2] = add_relu_f16x2(h1[4 * c + 2], bb.z);
          a[4 * c + 3] = add_relu_f16x2(h1[4 * c + 3], bb.w);
        }
        // (computing the row BEFORE waiting for the slot was measured: no change, profiles/r02_fcomb_l3_variants.md)
        tmem_st32(lane_addr + FC_A1_COL + slot * 32, a);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(a1_full(slot));
        if (warp == 0) FC_ACC(1);
      }
    }
  } else if (warp < FC_CTRL_WARP) {
    // ================================================================ epilogue (lane quarter = warp % 4)
    const int q = warp & 3;
    const int prow = q * 32 + lane;  // pixel row of this thread inside the tile
    const uint32_t lane_addr = tmem + (static_cast<uint32_t>(q * 32) << 16);
    uint32_t e_it = 0;
    const float b3r = __ldg(b3);
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int b = tile / tiles_per_img;
      const int pix = (tile - b * tiles_per_img) * FC_TILE + prow;
      const bool valid = pix < P;
      const size_t gp = static_cast<size_t>(b) * P + pix;
      float psum = 0.f;
      int count = 0;
#if FC_L3_MMA
      // Two tensor-memory round trips (tcgen05.ld / st + wait) per sample: one wait for the loads of this sample's 64
      // accumulator columns AND an earlier sample's two last-layer columns, one for the store of relu(H2).
      // FC_LAG = FC_H2_RING - 1 samples are in flight behind the one being converted: with three buffers the last-layer
      // columns read in iteration s are those of sample s - 2, whose MMA finished a whole iteration ago -- with two
      // buffers (lag 1) the warp waits out that MMA's round trip at the top of every iteration (cycle counters of
      // -DFC_PROFILE: 340 of 1080 cycles per sample).
      constexpr int FC_LAG = FC_H2_RING - 1;
      auto finish_math = [&](int s, float d0, float d1) {
        const float logit = (d0 + d1) + b3r;
        const float pr = __fdividef(1.0f, 1.0f + __expf(-logit));
        psum += pr;
        count += (pr >= upper || pr <= lower) ? 1 : 0;
        if (valid) {
          if (logits) logits[(static_cast<size_t>(s) * B + b) * P + pix] = logit;
          if (probs) probs[(static_cast<size_t>(s) * B + b) * P + pix] = pr;
        }
      };
      for (int s = 0; s < S; ++s, ++e_it) {
        const uint32_t hb = e_it % FC_H2_RING;
        const bool due = s >= FC_LAG;                       // sample s - FC_LAG of this tile has its columns ready
        const uint32_t pit = e_it - FC_LAG, pb = pit % FC_H2_RING;
        FC_T0();
        mbar_wait(h2_full(hb), (e_it / FC_H2_RING) & 1);
        if (due) mbar_wait(d3_full(pb), (pit / FC_H2_RING) & 1);
        tc_fence_after();
        if (warp == 4) FC_ACC(2);
        uint32_t v0[32], v1[32], d[2] = {0u, 0u};
        tmem_ld32(lane_addr + FC_H2_COL + hb * FCT, v0);
        tmem_ld32(lane_addr + FC_H2_COL + hb * FCT + 32, v1);
        if (due) tmem_ld2(lane_addr + FC_H2_COL + pb * FCT + 32, d);
        tmem_ld_wait();
        if (due) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(h2_empty(pb));  // that sample's accumulator buffer can be overwritten
        }
        // A3 = relu(H2) as packed fp16 (packed in place: result i only overwrites inputs that are already consumed),
        // written over columns 0..31 of this thread's own accumulator row
#pragma unroll
        for (int i = 0; i < 16; ++i) v0[i] = pack_relu_f16x2(__uint_as_float(v0[2 * i]), __uint_as_float(v0[2 * i + 1]));
#pragma unroll
        for (int i = 0; i < 16; ++i) v0[16 + i] = pack_relu_f16x2(__uint_as_float(v1[2 * i]), __uint_as_float(v1[2 * i + 1]));
        tmem_st32(lane_addr + FC_H2_COL + hb * FCT, v0);
        if (due) finish_math(s - FC_LAG, __uint_as_float(d[0]), __uint_as_float(d[1]));   // under the store's latency
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(a3_full(hb));
        if (warp == 4) FC_ACC(3);
      }
      // the tile's last FC_LAG samples
      for (int s = (S > FC_LAG ? S - FC_LAG : 0); s < S; ++s) {
        const uint32_t pit = e_it - (uint32_t)(S - s), pb = pit % FC_H2_RING;
        FC_T0();
        mbar_wait(d3_full(pb), (pit / FC_H2_RING) & 1);
        tc_fence_after();
        if (warp == 4) FC_ACC(4);
        uint32_t d[2];
        tmem_ld2(lane_addr + FC_H2_COL + pb * FCT + 32, d);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(h2_empty(pb));
        finish_math(s, __uint_as_float(d[0]), __uint_as_float(d[1]));
        if (warp == 4) FC_ACC(5);
      }
#else
      // round-1 epilogue: logit = w3 . relu(H2) + b3 in fp32 (64 FMNMX + 64 FFMA per pixel-sample), w3 read from shared
      // memory as warp-wide broadcast LDS.128
      const uint32_t w3s_addr = sbase + M::W3_OFF;
      for (int s = 0; s < S; ++s, ++e_it) {
        const uint32_t hb = e_it % FC_H2_RING;
        mbar_wait(h2_full(hb), (e_it / FC_H2_RING) & 1);
        tc_fence_after();
        float l0 = b3r, l1 = 0.f, l2 = 0.f, l3 = 0.f;  // b3 + four independent chains
#pragma unroll
        for (int part = 0; part < 4; ++part) {
          uint32_t v[16];
          tmem_ld16(lane_addr + FC_H2_COL + hb * FCT + part * 16, v);
          tmem_ld_wait();
          if (part == 3) {
            // all four quarters are in registers: the accumulator buffer can be overwritten
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(h2_empty(hb));
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint4 t = lds128(w3s_addr + 4 * (part * 16 + 4 * i));
            l0 = fmaf(__uint_as_float(t.x), fmaxf(__uint_as_float(v[4 * i + 0]), 0.f), l0);
            l1 = fmaf(__uint_as_float(t.y), fmaxf(__uint_as_float(v[4 * i + 1]), 0.f), l1);
            l2 = fmaf(__uint_as_float(t.z), fmaxf(__uint_as_float(v[4 * i + 2]), 0.f), l2);
            l3 = fmaf(__uint_as_float(t.w), fmaxf(__uint_as_float(v[4 * i + 3]), 0.f), l3);
          }
        }
        const float logit = (l0 + l1) + (l2 + l3);
        const float pr = __fdividef(1.0f, 1.0f + __expf(-logit));
        psum += pr;
        count += (pr >= upper || pr <= lower) ? 1 : 0;
        if (valid) {
          if (logits) logits[(static_cast<size_t>(s) * B + b) * P + pix] = logit;
          if (probs) probs[(static_cast<size_t>(s) * B + b) * P + pix] = pr;
        }
      }
#endif
      if (valid) {
        if (mean_prob) mean_prob[gp] = psum / static_cast<float>(S);
        if (cons_weight) cons_weight[gp] = static_cast<float>(count) / static_cast<float>(S);
        if (cons_mask) cons_mask[gp] = (count == S) ? 1 : 0;
      }
    }
  } else {
    // ================================================================ control: TMA + every tcgen05.mma
    // warp-uniform control flow (addresses / descriptors stay in uniform registers); one elected lane issues
    const bool leader = elect_one();
    // F x W1 hi/lo: bf16 x bf16, or fp16 x fp16 when the trunk ran with fp16 activations
    const uint32_t idesc1 = feat_f16 ? umma_idesc_f16(128, FCT) : umma_idesc_bf16(128, FCT);
    constexpr uint32_t idesc2 = umma_idesc_f16(128, FCT);   // A1 (fp16) x W2 (fp16)
    constexpr uint32_t idesc3 = umma_idesc_f16(128, 16);    // A3 (fp16, TMEM) x [w3_hi | w3_lo | 0 ..] (fp16)
    const uint64_t dW1 = umma_desc_k_sw128(sbase + M::W1_OFF);
    const uint64_t dW1L = umma_desc_k_sw128(sbase + M::W1L_OFF);
    const uint64_t dW2 = umma_desc_k_sw128(sbase + M::W2_OFF);
    const uint64_t dW2X = umma_desc_k_sw128(sbase + M::W2X_OFF);
    const uint64_t dW3 = umma_desc_k_sw128(sbase + M::W3T_OFF);
    const uint64_t dF = umma_desc_k_sw128(sbase + M::F_OFF);
    const uint64_t dAX = umma_desc_k_sw128(sbase + M::AX_OFF, /*sbo_bytes=*/0);  // all 8-row groups alias one atom
    auto load_tile = [&](int tile) {
      const int b = tile / tiles_per_img;
      if (leader) {
        mbar_expect_tx(f_full, M::A_BYTES);
        tma_load_3d(sbase + M::F_OFF, &tmF, f_full, 0, (tile - b * tiles_per_img) * FC_TILE, b);
      }
      __syncwarp();
    };
    auto mma1 = [&](uint32_t t_it) {
      // H1 = F . (W1 hi + lo)^T once the tile has landed and the producers have drained the previous H1
      mbar_wait(f_full, t_it & 1);
      mbar_wait(h1_empty, (t_it & 1) ^ 1);
      tc_fence_after();
      if (leader) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem + FC_H1_COL, dF + 2 * k, dW1 + 2 * k, idesc1, k ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem + FC_H1_COL, dF + 2 * k, dW1L + 2 * k, idesc1, 1u);
        umma_commit(h1_full);
        umma_commit(f_empty);
      }
      __syncwarp();
    };
    // D3 = relu(H2) . w3: A from columns 0..31 of accumulator buffer (it % FC_H2_RING), D into its columns 32..47
    auto mma3 = [&](uint32_t it) {
      const uint32_t pb = it % FC_H2_RING;
      mbar_wait(a3_full(pb), (it / FC_H2_RING) & 1);
      tc_fence_after();
      if (leader) {
        const uint32_t ta = tmem + FC_H2_COL + pb * FCT;
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16_ts(ta + 32, ta + 8 * k, dW3 + 2 * k, idesc3, k ? 1u : 0u);
        umma_commit(d3_full(pb));
      }
      __syncwarp();
    };
    uint32_t a_it = 0, t_it = 0, m3_it = 0;   // m3_it: next sample whose last-layer MMA is due
    int tile = blockIdx.x;
    if (tile < num_tiles) {
      load_tile(tile);
      mma1(0);
    }
    for (; tile < num_tiles; tile += gridDim.x, ++t_it) {
      const int next = tile + gridDim.x;
      if (next < num_tiles) {
        mbar_wait(f_empty, t_it & 1);  // MMA1 of this tile has consumed the feature tile
        load_tile(next);
      }
      for (int s = 0; s < S; ++s, ++a_it) {
        const uint32_t slot = a_it % FC_A1_RING, hb = a_it % FC_H2_RING;
        FC_T0();
        mbar_wait(a1_full(slot), (a_it / FC_A1_RING) & 1);
        FC_ACC(8);
        mbar_wait(h2_empty(hb), ((a_it / FC_H2_RING) & 1) ^ 1);
        tc_fence_after();
        FC_ACC(9);
        if (leader) {
          const uint32_t d = tmem + FC_H2_COL + hb * FCT;
          const uint32_t ta = tmem + FC_A1_COL + slot * 32;  // K = 16 halves = 8 columns per MMA
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16_ts(d, ta + 8 * k, dW2 + 2 * k, idesc2, k ? 1u : 0u);
          umma_bf16(d, dAX, dW2X, idesc2, 1u);  // + b2 (hi + lo)
          umma_commit(h2_full(hb));
          umma_commit(a1_empty(slot));
        }
        __syncwarp();
#if FC_L3_MMA
        // the last layer of the previous sample (after this sample's second-layer MMA is in the queue, so that the
        // tensor pipe has work while the epilogue converts)
        FC_ACC(10);
        while (m3_it < a_it) mma3(m3_it++);
        FC_ACC(11);
#endif
        if (!FC_H1_ALIAS) {
          // next tile's H1 as soon as half of this tile's samples are issued (the producers copied H1 to registers at
          // the start of the tile, so its TMEM columns are free; the feature tile was prefetched above)
          if (s == (S >> 1) && next < num_tiles) mma1(t_it + 1);
        }
      }
      if (FC_H1_ALIAS && next < num_tiles) {
        // H1 shares its columns with the A1 ring: the second-layer MMAs of this tile's last samples must have read
        // their rows (their a1_empty commits) before the layer-1 MMA of the next tile overwrites them
        for (uint32_t back = 1; back <= (uint32_t)FC_A1_RING && back <= a_it; ++back) {
          const uint32_t it = a_it - back;
          mbar_wait(a1_empty(it % FC_A1_RING), (it / FC_A1_RING) & 1);
        }
        mma1(t_it + 1);
      }
#if FC_L3_MMA
      while (m3_it < a_it) mma3(m3_it++);   // the tile's last sample (the epilogue writes the tile's outputs after it)
#endif
    }
#if !FC_L3_MMA
    (void)mma3;
#endif
  }

  tc_fence_before();
  __syncthreads();
  if (warp == FC_CTRL_WARP) tmem_dealloc(tmem, FC_TMEM_COLS);
}

// bz[s][b][j] = b1[j] + sum_d W1[j][64 + d] * z[s][b][d]: the per-(sample, image) bias that replaces the tiled-z concat.
// ONE block, so that the same launch can also (re)initialise the overflow flag: *oflag = any |bz| outside the fp16-safe
// range (no separate memset; the tensor-core kernel ORs its own H1 range check into it afterwards).
__global__ void __launch_bounds__(1024)
fcomb_bz_kernel(const float* __restrict__ z, const float* __restrict__ w1, const float* __restrict__ b1,
                const float* __restrict__ w2, const float* __restrict__ b2, const float* __restrict__ w3,
                float* __restrict__ bz, int SB, int L, int* __restrict__ oflag, float* __restrict__ bounds) {
  // bounds[0] = max |bz|, [1] = max_c sum_k |W2[c][k]|, [2] = max |b2|: the tensor-core kernel bounds |H2| with them
  __shared__ unsigned int smax[3];
  if (threadIdx.x < 3) smax[threadIdx.x] = 0u;
  __syncthreads();
  int over = 0;
  float m = 0.f;
  for (int i = threadIdx.x; i < SB * FCT; i += blockDim.x) {
    const int sb = i / FCT, j = i - sb * FCT;
    float acc = b1[j];
    for (int d = 0; d < L; ++d) acc = fmaf(w1[j * (FCT + L) + FCT + d], z[sb * L + d], acc);
    bz[i] = acc;
    over |= !(fabsf(acc) < FC_F16_SAFE);
    m = fmaxf(m, fabsf(acc));
  }
  atomicMax(&smax[0], __float_as_uint(m));  // non-negative floats order like their bit patterns
  if (threadIdx.x < FCT) {
    float rs = 0.f;
    for (int k = 0; k < FCT; ++k) rs += fabsf(w2[threadIdx.x * FCT + k]);
    over |= !(rs < FC_F16_SAFE) || !(fabsf(w3[threadIdx.x]) < FC_F16_SAFE);  // (also NaN / inf / fp16-unsafe weights)
    atomicMax(&smax[1], __float_as_uint(rs));
    atomicMax(&smax[2], __float_as_uint(fabsf(b2[threadIdx.x])));
  }
  over = __syncthreads_or(over);
  if (threadIdx.x == 0) *oflag = over ? 1 : 0;
  if (threadIdx.x < 3) bounds[threadIdx.x] = __uint_as_float(smax[threadIdx.x]);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace pda

using namespace pda;

#ifdef FC_PROFILE
// [0] producer wait slot, [1] producer work, [2] epilogue wait H2, [3] convert, [4] wait D3, [5] finish,
// [8] control wait A1, [9] control wait H2 buffer, [10] issue MMA2, [11] last-layer MMA incl. its wait
extern "C" int pda_fcomb_profile_read(unsigned long long* out16, int reset) {
  if (cudaMemcpyFromSymbol(out16, fc_prof, sizeof(unsigned long long) * 16) != cudaSuccess) return PDA_ERR_CUDA;
  if (reset) {
    unsigned long long z[16] = {0};
    if (cudaMemcpyToSymbol(fc_prof, z, sizeof(z)) != cudaSuccess) return PDA_ERR_CUDA;
  }
  return PDA_OK;
}
#endif

// scratch (fp32 words, caller-allocated per call -> no state shared between launches / streams / graphs):
// [0] overflow flag (int), [1..3] range bounds (max |bz|, max row-L1 of W2, max |b2|), [4 ..) bz[S][B][64]
extern "C" long long pda_fcomb_scratch_floats(int S, int B) { return 4 + (long long)S * B * FCT; }

extern "C" int pda_fcomb_mc_consensus(const void* feat, const float* z, const float* w1, const float* b1,
                                      const float* w2, const float* b2, const float* w3, const float* b3, int B, int P,
                                      int S, int latent, float upper, float lower, float* mean_prob,
                                      float* cons_weight, int64_t* cons_mask, float* logits, float* probs,
                                      float* scratch, int feat_f16, void* stream) {
  if (!feat || !z || !w1 || !b1 || !w2 || !b2 || !w3 || !b3 || !scratch) return PDA_ERR_ARG;
  if (B <= 0 || P <= 0 || S <= 0 || latent <= 0) return PDA_ERR_SHAPE;
  if (S > 640) return PDA_ERR_SHAPE;  // the fp32 range-guard fallback stages bz[S][64] fp32 in shared memory
  const int smem = FcombSmem::bytes(S);
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(get_encode_tiled());
  if (!enc) return PDA_ERR_DRIVER;
  CUtensorMap tm;
  cuuint64_t dims[3] = {64, (cuuint64_t)P, (cuuint64_t)B};
  cuuint64_t strides[2] = {128, (cuuint64_t)P * 128};
  cuuint32_t box[3] = {64, FC_TILE, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(feat), dims, strides, box, estr,
          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return PDA_ERR_TENSORMAP;
  cudaStream_t st = (cudaStream_t)stream;
  static int configured[64];
  if (dyn_smem_attr_needed(configured, smem)) {
    if (cudaFuncSetAttribute(fcomb_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
      return PDA_ERR_CUDA;
  }
  const int tiles_per_img = (P + FC_TILE - 1) / FC_TILE;
  const long long num_tiles = (long long)tiles_per_img * B;
  if (num_tiles > 0x7fffffffLL || (long long)S * B * FCT > 0x7fffffffLL) return PDA_ERR_SHAPE;
  // persistent grid: exactly the number of CTAs that are resident at once (a partial second wave would serialise);
  // tensor memory (256 of the 512 columns per CTA) and shared memory decide
  int per_sm = (227 * 1024) / (smem + 1024);
  if (per_sm < 1) return PDA_ERR_SHAPE;
  if (per_sm > FC_CTAS_PER_SM) per_sm = FC_CTAS_PER_SM;
  const int grid = (int)(num_tiles < 148 * per_sm ? num_tiles : 148 * per_sm);
  int* oflag = reinterpret_cast<int*>(scratch);
  float* bz = scratch + 4;
  PDA_COUNT(2);
  fcomb_bz_kernel<<<1, 1024, 0, st>>>(z, w1, b1, w2, b2, w3, bz, S * B, latent, oflag, scratch + 1);
  fcomb_tc_kernel<<<grid, FC_THREADS, smem, st>>>(tm, bz, w1, w2, b2, w3, b3, oflag, scratch + 1, feat_f16, P, S, latent, B, tiles_per_img,
                                                  (int)num_tiles, upper, lower, mean_prob, cons_weight, cons_mask,
                                                  logits, probs);
  if (cudaGetLastError() != cudaSuccess) return PDA_ERR_CUDA;
  // fp16 range guard: when the flag is up, the exact fp32 kernel overwrites every output of this call (device-side
  // decision, no host synchronisation; its blocks exit at once otherwise)
  return fcomb_mc_fp32(feat, z, w1, b1, w2, b2, w3, b3, B, P, S, latent, upper, lower, mean_prob, cons_weight,
                       cons_mask, logits, probs, oflag, feat_f16, st);
}
