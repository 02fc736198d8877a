// Weight gradient of conv3x3 (pad 1) on tcgen05 tensor cores.
//
// Replaces the cuDNN wgrad that autograd runs for every nn.Conv2d(k=3) of
// /root/reference/prob_utils/my_models/unet_blocks.py:19-24 and probabilistic_unet.py:56-61 in backward.
//
//   dW[co][ky][kx][ci] = sum_{y,x} dZ[y][x][co] * X[y+ky-1][x+kx-1][ci]
//                      = sum_{y,x'} dZ[y][x'-kx+1][co] * X[y+ky-1][x'][ci]          (x' = x + kx - 1)
//
// GEMM view with the reduction over PIXELS as K.  Both operands are read straight from their NHWC tensors by TMA
// (SWIZZLE_128B) and fed to the MMA as MN-major operands (the channel dimension is contiguous): no transpose is ever
// materialised.  What the MMA shape has to respect (profiles/r01e_umma_issue_rate.md): an M = 128 x N x K = 16 MMA
// needs N/2 tensor cycles but (4 KB + N * 32 B) / 128 B/clk of shared-memory operand fetch -- N = 64 tiles are capped at
// 2/3 of the tensor rate by the operand fetch alone (and the TMA fill shares that port), N >= 128 are not; MN-major
// operands cost nothing extra.  So the taps go into N, not into M:
//   work item  = (64 output channels, 64 input channels): all nine taps
//   pixel tile = 16 rows x 8 px of one image = K 128, as 8 K-steps of 16 px (two rows)
//   B (N x K)  = X^T: ONE slab [18 rows][8 px][64 ci]; the three ky taps of an output row are the slab rows
//                y, y+1, y+2 = three 64-channel N blocks 1024 B apart  ->  N = 192
//   A (M x K)  = dZ^T: ONE buffer [16 rows][16 px][64 co] (px x0 - 1 ..; TMA zero-fills the border); tap kx reads each row
//                from pixel 2 - kx on (descriptor start address + (2 - kx) * 128 B).  M = 128 = the block pair
//                (kx 2 | kx 1) (LBO = 128 B: the same rows one pixel apart); a second MMA (M = 64) takes kx = 0
//   D          = [128 (kx, co)][192 (ky, ci)] + an M = 64 accumulator [64 co][192] = 384 TMEM columns (+16: bias sums)
// Two MMAs per K-step instead of five (plus a 16-column "ones" MMA for the bias gradient on the items of input chunk 0).
// Measured with in-kernel cycle counters (round 1e): the MMA warp waits ~130 cycles per step for data and spends ~2130
// cycles per step issuing against a full MMA queue, i.e. the tensor pipe runs 16 MMAs in 2130 cycles where the tensor
// rate alone would need 1536.  The shared-memory port explains it: 66 KB of TMA fill + 144 KB of operand fetch per step
// = 1640 cycles at 128 B/clk (skipping two of the three dZ slab loads: -7 %; skipping the kx = 2 MMA: -21 %).  Round 2
// loads dZ once (32 KB instead of 48 KB per step, four pipeline stages instead of three); the X slab is still fetched
// twice per K-step (M is capped at 128, the three kx need 192 rows): that half-rate M = 64 MMA is the remaining lever.
// Persistent stream-K schedule: the (item, pixel tile) steps are split into 148 equal contiguous ranges; a CTA flushes
// its accumulators with coalesced fp32 atomics (lanes = output channels) into a zero-initialised scratch
// [9 taps][ctot][cout] whenever its range crosses an item boundary.
#include "conv.cuh"
#include "ptx.cuh"

namespace pda {

__device__ __forceinline__ void named_bar(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

struct WgradArgs {
  int B, H, W;
  int c0, c1, cout;
  int tiles_x, tiles_y;
  int num_tiles;   // pixel tiles = B * tiles_x * tiles_y
  int n64;         // cout / 64
  int items;       // n64 * (c0 + c1) / 64
  float* scratch;  // [9][ctot][cout] fp32, all zero at launch
  float* dbias;    // [cout] fp32 scratch behind it, zero at launch, or nullptr: bias gradient = column sums of dZ, from
                   // one extra "ones" MMA on the items of input-channel chunk 0
  float* part;     // deterministic mode (else nullptr): per-CTA partial accumulators [grid][slots_per_cta][WG_SLOT_FLOATS],
                   // plain stores instead of atomics; wgrad_reduce_kernel sums them in CTA order
  int slots_per_cta;
};

// one partial-accumulator slot: the M = 128 accumulator [192 cols][128 rows], the M = 64 one [192 cols][64 rows], 64 bias sums
constexpr int WG_SLOT_FLOATS = 192 * 128 + 192 * 64 + 64;

struct WgradSmem {
  // ONE dZ buffer of 16-pixel rows (px x0 - 1 .. x0 + 14, TMA zero-fills outside the image) serves the three column
  // shifts: the A operand of tap kx starts (2 - kx) pixels = (2 - kx) * 128 B into a row (an unaligned start inside the
  // 1024-byte swizzle atom is fine: the swizzle acts on absolute address bits, profiles/r02_conv_wide_slab.md) and its
  // eight K rows run on into the second atom of the row.  Round 1 loaded three shifted 8-pixel slabs (48 KB per stage).
  static constexpr int DZ_ROW = 2048;                // one row of the dZ buffer: 16 px x 128 B
  static constexpr int DZ_BUF = 16 * DZ_ROW;         // [16 rows][16 px][64 co] bf16
  static constexpr int X_SLAB = 18 * 1024;           // [18 rows][8 px][64 ci] bf16
  static constexpr int X_OFF = DZ_BUF;
  static constexpr int STAGE_BYTES = DZ_BUF + X_SLAB;
  static constexpr int STAGES = 4;
  static constexpr int ONES_OFF = STAGES * STAGE_BYTES;  // 1 KB of bf16 ones: an MN-major B block whose K rows all alias
  static constexpr int BAR_OFF = ONES_OFF + 1024;
  static constexpr int SLOT_OFF = BAR_OFF + (2 * STAGES + 2) * 8;
  static constexpr int DYN_BYTES = SLOT_OFF + 16 + 1024;
  static constexpr int TMEM_COLS = 512;              // [0,192) kx 2|1, [192,384) kx 0, [384,400) bias column sums
  static constexpr int D2_COL = 192, BIAS_COL = 384;
};

__global__ void __launch_bounds__(192, 1)
wgrad3x3_tc_kernel(const __grid_constant__ CUtensorMap tmX0, const __grid_constant__ CUtensorMap tmX1,
                   const __grid_constant__ CUtensorMap tmDZ, const WgradArgs p) {
  using L = WgradSmem;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar_base = sbase + L::BAR_OFF;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (L::STAGES + s); };
  const uint32_t acc_full = bar_base + 8u * (2 * L::STAGES);
  const uint32_t acc_empty = acc_full + 8;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + L::SLOT_OFF);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ctot = p.c0 + p.c1;
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  // this CTA's contiguous range of (item, tile) steps
  const long long total = (long long)p.items * p.num_tiles;
  const long long s0 = total * blockIdx.x / gridDim.x, s1 = total * (blockIdx.x + 1) / gridDim.x;

  if (threadIdx.x == 0) {
    for (int s = 0; s < L::STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, 4);
    fence_mbar_init();
    tma_prefetch_desc(&tmX0);
    tma_prefetch_desc(&tmX1);
    tma_prefetch_desc(&tmDZ);
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), L::TMEM_COLS);
    tmem_relinquish();
  }
  if (threadIdx.x >= 64 && threadIdx.x < 128)
    reinterpret_cast<uint4*>(smem + L::ONES_OFF)[threadIdx.x - 64] =
        make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (warp-uniform, elected lane issues)
    const bool leader = elect_one();
    uint32_t it = 0;
    for (long long st = s0; st < s1; ++st, ++it) {
      const int item = (int)(st / p.num_tiles);
      const int t = (int)(st - (long long)item * p.num_tiles);
      const int nb = item % p.n64, ch = item / p.n64;
      const int img = t / tiles_per_img;
      const int tt = t - img * tiles_per_img;
      const int ty = tt / p.tiles_x, tx = tt - ty * p.tiles_x;
      const int x0 = tx * 8, y0 = ty * 16;
      const int s = it % L::STAGES;
      mbar_wait(empty_bar(s), ((it / L::STAGES) & 1) ^ 1);
      if (leader) {
        mbar_expect_tx(full_bar(s), L::STAGE_BYTES);
        const uint32_t sa = sbase + s * L::STAGE_BYTES;
        const int c = ch << 6;
        // dZ columns x0 - 1 .. x0 + 14 (out-of-image columns arrive as zeros)
        tma_load_4d(sa, &tmDZ, full_bar(s), nb << 6, x0 - 1, y0, img);
        if (c < p.c0)
          tma_load_4d(sa + L::X_OFF, &tmX0, full_bar(s), c, x0, y0 - 1, img);
        else
          tma_load_4d(sa + L::X_OFF, &tmX1, full_bar(s), c - p.c0, x0, y0 - 1, img);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    const bool leader = elect_one();
    constexpr uint32_t idesc = umma_idesc_bf16(128, 192, /*a_mn_major=*/1, /*b_mn_major=*/1);
    // the third kx and the bias sums have only 64 useful rows: M = 64 MMAs (same tensor time as M = 128 -- measured --
    // but half the A-operand fetch and half the multipliers switching).  An M = 64 accumulator keeps rows 16 i .. 16 i + 15
    // in TMEM lanes 32 i .. 32 i + 15 (profiles/r01e_umma_issue_rate.md)
    constexpr uint32_t idesc64 = umma_idesc_bf16(64, 192, 1, 1);
    constexpr uint32_t idesc_bias = umma_idesc_bf16(64, 16, 1, 1);
    uint32_t it = 0, flushes = 0;
    int cur_item = -1;
    for (long long st = s0; st < s1; ++st, ++it) {
      const int item = (int)(st / p.num_tiles);
      const bool first = item != cur_item;
      if (first) {
        if (cur_item >= 0) {
          if (leader) umma_commit(acc_full);  // previous item complete -> epilogue
          __syncwarp();
          ++flushes;
        }
        mbar_wait(acc_empty, (flushes & 1) ^ 1);  // accumulators drained (passes immediately the first time)
        tc_fence_after();
        cur_item = item;
      }
      const int s = it % L::STAGES;
      mbar_wait(full_bar(s), (it / L::STAGES) & 1);
      tc_fence_after();
      if (leader) {
        const uint32_t sa = sbase + s * L::STAGE_BYTES;
        // MN-major SWIZZLE_128B: 64-channel blocks LBO apart, 8-pixel K groups SBO apart.  dZ: tap kx reads the row
        // from pixel (2 - kx) on; the M = 128 operand is the block pair (kx 2 | kx 1) = the same rows one pixel
        // (LBO = 128 B) apart; a K group is one row of the 16-pixel buffer (SBO = 2048).  X: N blocks = the slab rows
        // y, y+1, y+2 (LBO = 1024), K groups = slab rows (SBO = 1024).
        const uint64_t da21 = umma_desc_mn_sw128(sa, 128, L::DZ_ROW);                    // M = (kx 2 | kx 1) x co
        const uint64_t da0 = umma_desc_mn_sw128(sa + 2 * 128, 0, L::DZ_ROW);             // M = 64: kx 0 x co
        const uint64_t da1 = umma_desc_mn_sw128(sa + 128, 0, L::DZ_ROW);                 // unshifted dZ (bias sums)
        const uint64_t db = umma_desc_mn_sw128(sa + L::X_OFF, 1024, 1024);               // N = (ky 0 | 1 | 2) x ci
        const uint64_t d1s = umma_desc_mn_sw128(sbase + L::ONES_OFF, 0, 0);              // all-ones K rows
        const bool bias_item = p.dbias != nullptr && item / p.n64 == 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          // 16 pixels per MMA = 2 image rows of 8 px: dZ advances 2 x 2048 B, X 2 x 1024 B (16-byte units)
          const uint32_t acc = (!first || k != 0) ? 1u : 0u;
          umma_bf16(tmem, da21 + 256 * k, db + 128 * k, idesc, acc);
          umma_bf16(tmem + L::D2_COL, da0 + 256 * k, db + 128 * k, idesc64, acc);
          if (bias_item) umma_bf16(tmem + L::BIAS_COL, da1 + 256 * k, d1s, idesc_bias, acc);  // sum_px dZ[px][co]
        }
        umma_commit(empty_bar(s));
      }
      __syncwarp();
    }
    if (cur_item >= 0) {
      if (leader) umma_commit(acc_full);
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------ epilogue: accumulator row = (kx of the pair, co)
    const int q = warp & 3;
    const int co_l = (q & 1) * 32 + lane;
    const uint32_t lane_addr = tmem + (static_cast<uint32_t>(q * 32) << 16);
    uint32_t flushes = 0;
    int cur_item = -1;
    const int first_item = (int)(s0 / p.num_tiles);
    // deterministic mode: the accumulators of (this CTA, item) go to the CTA's own slot with plain, coalesced stores
    auto flush_det = [&](int item) {
      const int ch = item / p.n64;
      mbar_wait(acc_full, flushes & 1);
      tc_fence_after();
      float* slot = p.part + (static_cast<size_t>(blockIdx.x) * p.slots_per_cta + (item - first_item)) * WG_SLOT_FLOATS;
#pragma unroll 1
      for (int part = 0; part < 2; ++part) {
        const bool live = part == 0 || lane < 16;
        float* dst0 = part == 0 ? slot + q * 32 + lane : slot + 192 * 128 + 16 * q + lane;
        const int rows = part == 0 ? 128 : 64;
#pragma unroll 1
        for (int c32 = 0; c32 < 6; ++c32) {
          uint32_t v[32];
          tmem_ld32(lane_addr + part * L::D2_COL + c32 * 32, v);
          tmem_ld_wait();
          if (live) {
#pragma unroll
            for (int j = 0; j < 32; ++j) dst0[static_cast<size_t>(c32 * 32 + j) * rows] = __uint_as_float(v[j]);
          }
        }
      }
      if (p.dbias != nullptr && ch == 0) {
        uint32_t v[16];
        tmem_ld16(lane_addr + L::BIAS_COL, v);
        tmem_ld_wait();
        if (lane < 16) slot[192 * 128 + 192 * 64 + 16 * q + lane] = __uint_as_float(v[0]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty);
      ++flushes;
    };
    auto flush = [&](int item) {
      if (p.part != nullptr) {
        flush_det(item);
        return;
      }
      const int nb = item % p.n64, ch = item / p.n64;
      mbar_wait(acc_full, flushes & 1);
      tc_fence_after();
      // lanes = consecutive output channels: every atomic instruction of a warp covers one 128-byte line
      float* base = p.scratch + (static_cast<size_t>(ch) << 6) * p.cout + (nb << 6) + co_l;
      const size_t tap_stride = static_cast<size_t>(ctot) * p.cout;
#pragma unroll 1
      for (int part = 0; part < 2; ++part) {
        // part 0: M = 128 accumulator, row = (kx, co) = this thread's lane.  part 1: M = 64 accumulator of kx = 0:
        // rows 16 q .. 16 q + 15 sit in the first 16 lanes of lane quarter q
        const int kx = part == 0 ? 2 - (q >> 1) : 0;   // accumulator rows: (kx 2 | kx 1) x co, then kx 0 x co
        const bool live = part == 0 || lane < 16;
        float* pbase = part == 0 ? base : base - co_l + 16 * q + lane;
#pragma unroll 1
        for (int c32 = 0; c32 < 6; ++c32) {
          const int ky = c32 >> 1, ci0 = (c32 & 1) * 32;
          uint32_t v[32];
          tmem_ld32(lane_addr + part * L::D2_COL + c32 * 32, v);
          tmem_ld_wait();
          if (live) {
            float* dst = pbase + (ky * 3 + kx) * tap_stride + static_cast<size_t>(ci0) * p.cout;
#pragma unroll
            for (int j = 0; j < 32; ++j) atomicAdd(dst + static_cast<size_t>(j) * p.cout, __uint_as_float(v[j]));
          }
        }
      }
      if (p.dbias != nullptr && ch == 0) {
        // every column of the ones-block accumulator (M = 64 layout) holds sum_px dZ[px][co]
        uint32_t v[16];
        tmem_ld16(lane_addr + L::BIAS_COL, v);
        tmem_ld_wait();
        if (lane < 16) atomicAdd(p.dbias + (nb << 6) + 16 * q + lane, __uint_as_float(v[0]));
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty);
      ++flushes;
    };
    for (long long st = s0; st < s1;) {
      const int item = (int)(st / p.num_tiles);
      if (cur_item >= 0 && item != cur_item) flush(cur_item);
      cur_item = item;
      // jump to the first step of the next item (or the end of the range)
      const long long nxt = (long long)(item + 1) * p.num_tiles;
      st = nxt < s1 ? nxt : s1;
    }
    if (cur_item >= 0) flush(cur_item);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, L::TMEM_COLS);
}

// scratch [9][ctot][cout] fp32 -> dW OIHW fp32 [cout][ctot][3][3] (written, or accumulated when accumulate != 0);
// the bias sums accumulated behind the scratch go to dbias.  clean != 0: everything that is read is ZEROED again, so that a
// scratch the caller keeps per layer is ready for the next launch without a memset (measured: the extra stores cost more
// than the memset they replace -- +4 % on the weight-gradient path -- so the Python host does not use it).
// A 32 (ci) x 32 (co) tile with all nine taps goes through shared memory: reads are coalesced along co, and for one co
// the 32 ci x 9 taps are 288 CONTIGUOUS floats of dW (grid: ctot / 32 x cout / 32; both are multiples of 64).
__global__ void __launch_bounds__(256)
wgrad_scatter_kernel(float* __restrict__ scratch, float* __restrict__ dw, int cout, int ctot, int accumulate,
                     float* __restrict__ dbias, int clean) {
  __shared__ float tile[9][32][33];
  const int ci0 = blockIdx.x * 32, co0 = blockIdx.y * 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (dbias != nullptr && blockIdx.x == 0) {
    if (threadIdx.x < 32) {
      const float v = scratch[9ull * cout * ctot + co0 + threadIdx.x];
      dbias[co0 + threadIdx.x] = accumulate ? dbias[co0 + threadIdx.x] + v : v;
      if (clean) scratch[9ull * cout * ctot + co0 + threadIdx.x] = 0.f;
    }
  }
#pragma unroll
  for (int tap = 0; tap < 9; ++tap)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int ci_l = warp + 8 * k;
      const size_t si = (static_cast<size_t>(tap) * ctot + ci0 + ci_l) * cout + co0 + lane;
      tile[tap][ci_l][lane] = scratch[si];
      if (clean) scratch[si] = 0.f;
    }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int co_l = warp + 8 * k;
    float* dst = dw + (static_cast<size_t>(co0 + co_l) * ctot + ci0) * 9;
#pragma unroll
    for (int e0 = 0; e0 < 288; e0 += 32) {
      const int e = e0 + lane;
      const int ci_l = e / 9, tap = e - 9 * ci_l;
      const float v = tile[tap][ci_l][co_l];
      dst[e] = accumulate ? dst[e] + v : v;
    }
  }
}

// Deterministic mode: dW OIHW (and the bias gradient) from the per-CTA partial slots, summed in CTA order (a fixed order
// for a fixed shape and SM budget: bit-identical from run to run; no memset, no atomics).  Same 32 (ci) x 32 (co) x 9 taps
// tiling through shared memory as wgrad_scatter_kernel.  The CTAs that hold a piece of item i are those whose step range
// [total c / G, total (c + 1) / G) meets [i T, (i + 1) T); inside CTA c the item sits in slot i - (first item of c).
// CI_T = input channels per block: 32, or 8 for layers with so few 32 x 32 tiles that the sum over up to 148 slots per item
// would run on a handful of SMs (64 -> 64: 4 blocks; measured 190 us for the reduction alone).
template <int CI_T>
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float* __restrict__ part, float* __restrict__ dw, float* __restrict__ dbias, int cout, int ctot,
                    int n64, int num_tiles, int grid_ctas, int slots_per_cta, int accumulate) {
  constexpr int KR = CI_T / 8;   // input channels per warp
  __shared__ float tile[9][CI_T][33];
  const int ci0 = blockIdx.x * CI_T, co0 = blockIdx.y * 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = (ci0 >> 6) * n64 + (co0 >> 6);
  const int ci_in = ci0 & 63, co_in = co0 & 63;
  const long long T = num_tiles, total = (long long)n64 * (ctot >> 6) * T;
  const long long lo = (long long)item * T, hi = lo + T;
  int c_first = (int)(lo * grid_ctas / total);
  while (c_first > 0 && total * c_first / grid_ctas > lo) --c_first;
  while (total * (c_first + 1) / grid_ctas <= lo) ++c_first;
  float acc[9][KR];
#pragma unroll
  for (int tap = 0; tap < 9; ++tap)
#pragma unroll
    for (int k = 0; k < KR; ++k) acc[tap][k] = 0.f;
  float bsum = 0.f;
  const bool bias_block = dbias != nullptr && blockIdx.x == 0 && threadIdx.x < 32;   // (ci0 = 0: an item of chunk 0)
  int c_end = c_first + 1;               // one past the last CTA whose range starts before the item ends
  while (c_end < grid_ctas && total * c_end / grid_ctas < hi) ++c_end;
  // (unrolled: the loads of four slots are in flight together -- the loop is latency-bound otherwise; the ADDS stay in
  // CTA order)
#pragma unroll 4
  for (int c = c_first; c < c_end; ++c) {
    const long long cs0 = total * c / grid_ctas;
    const float* slot = part + (static_cast<size_t>(c) * slots_per_cta + (item - (int)(cs0 / T))) * WG_SLOT_FLOATS;
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int ky = tap / 3, kx = tap - 3 * ky;
#pragma unroll
      for (int k = 0; k < KR; ++k) {
        const int col = ky * 64 + ci_in + warp + 8 * k;
        const float v = kx == 0 ? slot[192 * 128 + col * 64 + co_in + lane]
                                : slot[col * 128 + (kx == 2 ? 0 : 64) + co_in + lane];
        acc[tap][k] += v;
      }
    }
    if (bias_block) bsum += slot[192 * 128 + 192 * 64 + co_in + threadIdx.x];
  }
  if (bias_block) dbias[co0 + threadIdx.x] = accumulate ? dbias[co0 + threadIdx.x] + bsum : bsum;
#pragma unroll
  for (int tap = 0; tap < 9; ++tap)
#pragma unroll
    for (int k = 0; k < KR; ++k) tile[tap][warp + 8 * k][lane] = acc[tap][k];
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int co_l = warp + 8 * k;
    float* dst = dw + (static_cast<size_t>(co0 + co_l) * ctot + ci0) * 9;
#pragma unroll
    for (int e0 = 0; e0 < 9 * CI_T; e0 += 32) {
      const int e = e0 + lane;
      if (e < 9 * CI_T) {
        const int ci_l = e / 9, tap = e - 9 * ci_l;
        const float v = tile[tap][ci_l][co_l];
        dst[e] = accumulate ? dst[e] + v : v;
      }
    }
  }
}

// db[c] = sum over pixels of dZ[p][c]   (dZ: [npix][C] bf16); db must be zero-initialised.
// Thread t owns channel pair (t % C2) for pixels (t / C2) + k * rows; the rows of a block are reduced through shared
// memory so that a block issues one atomic per channel.
__global__ void __launch_bounds__(1024)
bias_grad_kernel(const __nv_bfloat162* __restrict__ dz, float* __restrict__ db, long long npix, int C2) {
  extern __shared__ float red[];  // [blockDim.x][2]
  const int cpair = threadIdx.x % C2;
  const int prow = threadIdx.x / C2;
  const int rows = blockDim.x / C2;
  float sx = 0.f, sy = 0.f;
  for (long long px = (long long)blockIdx.x * rows + prow; px < npix; px += (long long)gridDim.x * rows) {
    const float2 v = __bfloat1622float2(__ldg(dz + px * C2 + cpair));
    sx += v.x;
    sy += v.y;
  }
  red[2 * threadIdx.x] = sx;
  red[2 * threadIdx.x + 1] = sy;
  __syncthreads();
  if (prow == 0) {
    for (int r = 1; r < rows; ++r) {
      sx += red[2 * (r * C2 + cpair)];
      sy += red[2 * (r * C2 + cpair) + 1];
    }
    atomicAdd(db + 2 * cpair, sx);
    atomicAdd(db + 2 * cpair + 1, sy);
  }
}

}  // namespace pda

using namespace pda;

// scratch layout (fp32 words): [9 * cout * ctot] tap-major partial sums | [cout] bias sums
extern "C" long long pda_conv3x3_wgrad_scratch_floats(int ctot, int cout) {
  if (ctot <= 0 || cout <= 0) return 0;
  return 9LL * cout * ctot + cout;
}

static int wgrad_grid(int ctot, int cout, int B, int H, int W, long long* steps_out, int* tiles_out) {
  const long long nt = (long long)((W + 7) / 8) * ((H + 15) / 16) * B;
  const long long steps = (long long)(cout >> 6) * (ctot >> 6) * nt;
  if (steps_out) *steps_out = steps;
  if (tiles_out) *tiles_out = (int)nt;
  const int sms = sm_budget(0);
  return (int)(steps < sms ? steps : sms);
}
// slots a CTA may need: the items its contiguous step range can touch
static int wgrad_slots_per_cta(long long steps, int grid, int num_tiles) {
  const long long per = (steps + grid - 1) / grid;
  return (int)((per + num_tiles - 1) / num_tiles) + 1;
}

// deterministic variant: scratch = per-CTA partial accumulators (size depends on the shape AND the SM budget in force)
extern "C" long long pda_conv3x3_wgrad_det_scratch_floats(int ctot, int cout, int B, int H, int W) {
  if (ctot <= 0 || cout <= 0 || B <= 0 || H <= 0 || W <= 0) return 0;
  long long steps;
  int nt;
  const int grid = wgrad_grid(ctot, cout, B, H, W, &steps, &nt);
  return (long long)grid * wgrad_slots_per_cta(steps, grid, nt) * WG_SLOT_FLOATS;
}

static int wgrad_launch(const void* src0, int c0, const void* src1, int c1, const void* dz, float* scratch,
                        float* dw_oihw, float* dbias, int B, int H, int W, int cout, int accumulate,
                        int scratch_is_zero, int deterministic, void* stream_);

extern "C" int pda_conv3x3_wgrad_bf16_det(const void* src0, int c0, const void* src1, int c1, const void* dz,
                                          float* scratch, float* dw_oihw, float* dbias, int B, int H, int W, int cout,
                                          int accumulate, void* stream_) {
  return wgrad_launch(src0, c0, src1, c1, dz, scratch, dw_oihw, dbias, B, H, W, cout, accumulate, 0, 1, stream_);
}

extern "C" int pda_conv3x3_wgrad_bf16(const void* src0, int c0, const void* src1, int c1, const void* dz,
                                      float* scratch, float* dw_oihw, float* dbias, int B, int H, int W, int cout,
                                      int accumulate, int scratch_is_zero, void* stream_) {
  return wgrad_launch(src0, c0, src1, c1, dz, scratch, dw_oihw, dbias, B, H, W, cout, accumulate, scratch_is_zero, 0,
                      stream_);
}

static int wgrad_launch(const void* src0, int c0, const void* src1, int c1, const void* dz, float* scratch,
                        float* dw_oihw, float* dbias, int B, int H, int W, int cout, int accumulate,
                        int scratch_is_zero, int deterministic, void* stream_) {
  if (!src0 || !dz || !scratch || !dw_oihw || (c1 > 0 && !src1)) return PDA_ERR_ARG;
  if (c0 <= 0 || (c0 & 63) || (c1 & 63) || (cout & 63) || B <= 0 || H <= 0 || W <= 0) return PDA_ERR_SHAPE;
  cudaStream_t stream = (cudaStream_t)stream_;
  const int ctot = c0 + c1;
  WgradArgs a;
  a.B = B; a.H = H; a.W = W; a.c0 = c0; a.c1 = c1; a.cout = cout;
  a.tiles_x = (W + 7) / 8;
  a.tiles_y = (H + 15) / 16;
  const long long nt = (long long)a.tiles_x * a.tiles_y * B;
  if (nt > 0x7fffffffLL) return PDA_ERR_SHAPE;
  a.num_tiles = (int)nt;
  a.n64 = cout >> 6;
  a.items = a.n64 * (ctot >> 6);
  a.scratch = scratch;
  // the bias sums are accumulated in the cout floats BEHIND the weight scratch
  a.dbias = dbias ? scratch + 9ull * cout * ctot : nullptr;
  a.part = nullptr;
  a.slots_per_cta = 0;
  CUtensorMap tX0, tX1, tDZ;
  int r = make_act_tensor_map(&tX0, src0, B, H, W, c0, 8, 18, 64);
  if (r) return r;
  if (c1 > 0) {
    r = make_act_tensor_map(&tX1, src1, B, H, W, c1, 8, 18, 64);
    if (r) return r;
  } else {
    tX1 = tX0;
  }
  r = make_act_tensor_map(&tDZ, dz, B, H, W, cout, 16, 16, 64);
  if (r) return r;
  // scratch_is_zero: the caller keeps this scratch across calls; the scatter kernel leaves it all-zero again (no memset)
  if (!deterministic && !scratch_is_zero &&
      cudaMemsetAsync(scratch, 0, sizeof(float) * pda_conv3x3_wgrad_scratch_floats(ctot, cout), stream) != cudaSuccess)
    return PDA_ERR_CUDA;
  static int configured[64];
  if (dyn_smem_attr_needed(configured, WgradSmem::DYN_BYTES)) {
    if (cudaFuncSetAttribute(wgrad3x3_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             WgradSmem::DYN_BYTES) != cudaSuccess)
      return PDA_ERR_CUDA;
  }
  const long long steps = (long long)a.items * a.num_tiles;
  const int sms = sm_budget(0);
  const int grid = (int)(steps < sms ? steps : sms);
  if (deterministic) {
    a.part = scratch;
    a.slots_per_cta = wgrad_slots_per_cta(steps, grid, a.num_tiles);
    a.dbias = dbias;   // (only tested for null-ness by the kernel in this mode)
  }
  PDA_COUNT(1);
  wgrad3x3_tc_kernel<<<grid, 192, WgradSmem::DYN_BYTES, stream>>>(tX0, tX1, tDZ, a);
  if (cudaGetLastError() != cudaSuccess) return PDA_ERR_CUDA;
  PDA_COUNT(1);
  if (deterministic) {
    if ((ctot / 32) * (cout / 32) < 148)
      wgrad_reduce_kernel<8><<<dim3(ctot / 8, cout / 32), 256, 0, stream>>>(scratch, dw_oihw, dbias, cout, ctot, a.n64,
                                                                            a.num_tiles, grid, a.slots_per_cta, accumulate);
    else
      wgrad_reduce_kernel<32><<<dim3(ctot / 32, cout / 32), 256, 0, stream>>>(scratch, dw_oihw, dbias, cout, ctot, a.n64,
                                                                              a.num_tiles, grid, a.slots_per_cta, accumulate);
    return cudaGetLastError() == cudaSuccess ? PDA_OK : PDA_ERR_CUDA;
  }
  wgrad_scatter_kernel<<<dim3(ctot / 32, cout / 32), 256, 0, stream>>>(scratch, dw_oihw, cout, ctot, accumulate, dbias,
                                                                       scratch_is_zero);
  return cudaGetLastError() == cudaSuccess ? PDA_OK : PDA_ERR_CUDA;
}
