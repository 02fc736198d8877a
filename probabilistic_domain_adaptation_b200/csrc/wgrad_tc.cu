// Weight gradient of conv3x3 (pad 1) on tcgen05 tensor cores.
//
// Replaces the cuDNN wgrad that autograd runs for every nn.Conv2d(k=3) of
// /root/reference/prob_utils/my_models/unet_blocks.py:19-24 and probabilistic_unet.py:56-61 in backward.
//
//   dW[co][tap][ci] = sum over pixels p of  dZ[p][co] * X[p + tap][ci]
//
// GEMM view with the reduction over PIXELS as K: both operands are read straight from their NHWC tensors with the
// same TMA boxes the forward conv uses ([128 pixels][64 channels], SWIZZLE_128B) and fed to the MMA as MN-major
// operands (the channel dimension is contiguous), so no transpose is ever materialised:
//   A (M x K) = X^T shifted by the tap:  M = 128 = two (tap, 64-channel chunk) "row pairs", 64 rows each
//   B (N x K) = dZ^T:                    N = BN output channels (64 or 128)
//   D (M x N) fp32 in TMEM, accumulated over this CTA's slice of pixel tiles (split-K across blockIdx.z), then
//   reduced into a zero-initialised fp32 scratch [cout][9][ctot] with red.global.add (coalesced along ci).
#include "conv.cuh"
#include "ptx.cuh"

namespace pda {

struct WgradArgs {
  int B, H, W;
  int c0, c1, cout;
  int tile_w, tile_h, tiles_x, tiles_y;
  int num_tiles;   // pixel tiles = B * tiles_x * tiles_y
  int npairs;      // 9 * (c0 + c1) / 64
  float* scratch;  // [cout][9][ctot] fp32, zero-initialised by the caller
};

template <int BN, int STAGES>
struct WgradSmem {
  static constexpr int BOX = 128 * 128;             // one TMA box: 128 pixels x 64 channels bf16
  static constexpr int A_BYTES = 2 * BOX;           // two row pairs
  static constexpr int B_BYTES = (BN / 64) * BOX;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int SLOT_OFF = BAR_OFF + (2 * STAGES + 1) * 8;
  static constexpr int DYN_BYTES = SLOT_OFF + 16 + 1024;
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(192, 1)
wgrad3x3_tc_kernel(const __grid_constant__ CUtensorMap tmX0, const __grid_constant__ CUtensorMap tmX1,
                   const __grid_constant__ CUtensorMap tmDZ, const WgradArgs p) {
  using L = WgradSmem<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar_base = sbase + L::BAR_OFF;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  const uint32_t done_bar = bar_base + 8u * (2 * STAGES);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + L::SLOT_OFF);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const int ctot = p.c0 + p.c1;
  const int chunks = ctot >> 6;
  // the two (tap, channel-chunk) row pairs of this CTA; a dangling second pair repeats the first (discarded)
  int pair[2] = {2 * (int)blockIdx.x, 2 * (int)blockIdx.x + 1};
  const bool second_valid = pair[1] < p.npairs;
  if (!second_valid) pair[1] = pair[0];
  const int n0 = blockIdx.y * BN;
  // pixel tiles t = blockIdx.z, blockIdx.z + gridDim.z, ...
  const int my_tiles = (p.num_tiles - (int)blockIdx.z + (int)gridDim.z - 1) / (int)gridDim.z;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(done_bar, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmX0);
    tma_prefetch_desc(&tmX1);
    tma_prefetch_desc(&tmDZ);
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // TMA producer: warp-uniform control flow, one elected lane issues
    const bool leader = elect_one();
    for (int it = 0; it < my_tiles; ++it) {
      const int s = it % STAGES;
      const uint32_t ph = (it / STAGES) & 1;
      const int t = blockIdx.z + it * gridDim.z;
      const int tx = t % p.tiles_x;
      const int ty = (t / p.tiles_x) % p.tiles_y;
      const int img = t / (p.tiles_x * p.tiles_y);
      const int x0 = tx * p.tile_w, y0 = ty * p.tile_h;
      mbar_wait(empty_bar(s), ph ^ 1);
      if (leader) {
        mbar_expect_tx(full_bar(s), L::STAGE_BYTES);
        const uint32_t sa = sbase + s * L::STAGE_BYTES;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int tap = pair[h] / chunks;
          const int c = (pair[h] - tap * chunks) << 6;
          const int ky = tap / 3, kx = tap - 3 * ky;
          if (c < p.c0)
            tma_load_4d(sa + h * L::BOX, &tmX0, full_bar(s), c, x0 + kx - 1, y0 + ky - 1, img);
          else
            tma_load_4d(sa + h * L::BOX, &tmX1, full_bar(s), c - p.c0, x0 + kx - 1, y0 + ky - 1, img);
        }
#pragma unroll
        for (int j = 0; j < BN / 64; ++j)
          tma_load_4d(sa + L::A_BYTES + j * L::BOX, &tmDZ, full_bar(s), n0 + 64 * j, x0, y0, img);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    const bool leader = elect_one();
    constexpr uint32_t idesc = umma_idesc_bf16(128, BN, /*a_mn_major=*/1, /*b_mn_major=*/1);
    for (int it = 0; it < my_tiles; ++it) {
      const int s = it % STAGES;
      const uint32_t ph = (it / STAGES) & 1;
      mbar_wait(full_bar(s), ph);
      tc_fence_after();
      if (leader) {
        const uint32_t sa = sbase + s * L::STAGE_BYTES;
        // MN-major, 128-byte swizzle: 64-channel blocks L::BOX apart (LBO), 8-pixel K groups 1024 B apart (SBO)
        const uint64_t da = umma_desc_mn_sw128(sa, L::BOX, 1024);
        const uint64_t db = umma_desc_mn_sw128(sa + L::A_BYTES, L::BOX, 1024);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          // 16 pixels per MMA = 16 rows of 128 B = 2048 B: +128 in 16-byte units
          umma_bf16(tmem, da + 128 * k, db + 128 * k, idesc, (it | k) != 0 ? 1u : 0u);
        }
        umma_commit(empty_bar(s));
      }
      __syncwarp();
    }
    if (leader) umma_commit(done_bar);
    __syncwarp();
  } else {
    // epilogue: thread = accumulator row = (row pair, input channel); columns = output channels
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int h = row >> 6;
    const int tap = pair[h] / chunks;
    const int ci = ((pair[h] - tap * chunks) << 6) + (row & 63);
    const bool live = (my_tiles > 0) && (h == 0 || second_valid);
    if (my_tiles > 0) {
      mbar_wait(done_bar, 0);
      tc_fence_after();
#pragma unroll 1
      for (int cb = 0; cb < BN / 32; ++cb) {
        uint32_t v[32];
        tmem_ld32(tmem + (static_cast<uint32_t>(q * 32) << 16) + cb * 32, v);
        tmem_ld_wait();
        if (live) {
          float* dst = p.scratch + (static_cast<size_t>(n0 + cb * 32) * 9 + tap) * ctot + ci;
#pragma unroll
          for (int j = 0; j < 32; ++j) atomicAdd(dst + static_cast<size_t>(j) * 9 * ctot, __uint_as_float(v[j]));
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, BN);
}

// scratch [cout][9][ctot] fp32 -> dW OIHW fp32 [cout][ctot][3][3] (written, or accumulated when accumulate != 0)
__global__ void wgrad_scatter_kernel(const float* __restrict__ scratch, float* __restrict__ dw, int cout, int ctot,
                                     int accumulate) {
  const long long n = 9LL * cout * ctot;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int tap = i % 9;
    const int ci = (i / 9) % ctot;
    const int co = i / (9LL * ctot);
    const float v = scratch[((long long)co * 9 + tap) * ctot + ci];
    dw[i] = accumulate ? dw[i] + v : v;
  }
}

// db[c] = sum over pixels of dZ[p][c]   (dZ: [npix][C] bf16); db must be zero-initialised
__global__ void __launch_bounds__(256)
bias_grad_kernel(const __nv_bfloat162* __restrict__ dz, float* __restrict__ db, long long npix, int C2) {
  // block handles a slab of pixels; thread t handles channel pair (t % C2) for pixels (t / C2) + k * (256 / C2)
  const int cpair = threadIdx.x % C2;
  const int prow = threadIdx.x / C2;
  const int rows = blockDim.x / C2;
  float sx = 0.f, sy = 0.f;
  for (long long px = (long long)blockIdx.x * rows + prow; px < npix; px += (long long)gridDim.x * rows) {
    const float2 v = __bfloat1622float2(__ldg(dz + px * C2 + cpair));
    sx += v.x;
    sy += v.y;
  }
  atomicAdd(db + 2 * cpair, sx);
  atomicAdd(db + 2 * cpair + 1, sy);
}

template <int BN, int STAGES>
static int launch_wgrad(const CUtensorMap& x0, const CUtensorMap& x1, const CUtensorMap& dz, const WgradArgs& a,
                        int n_blocks, int ksplit, cudaStream_t stream) {
  using L = WgradSmem<BN, STAGES>;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(wgrad3x3_tc_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             L::DYN_BYTES) != cudaSuccess)
      return PDA_ERR_CUDA;
    configured = true;
  }
  dim3 grid((a.npairs + 1) / 2, n_blocks, ksplit);
  PDA_COUNT(1);
  wgrad3x3_tc_kernel<BN, STAGES><<<grid, 192, L::DYN_BYTES, stream>>>(x0, x1, dz, a);
  return cudaGetLastError() == cudaSuccess ? PDA_OK : PDA_ERR_CUDA;
}

}  // namespace pda

using namespace pda;

extern "C" int pda_conv3x3_wgrad_bf16(const void* src0, int c0, const void* src1, int c1, const void* dz,
                                      float* scratch, float* dw_oihw, float* dbias, int B, int H, int W, int cout,
                                      int accumulate, void* stream_) {
  if (!src0 || !dz || !scratch || !dw_oihw || (c1 > 0 && !src1)) return PDA_ERR_ARG;
  if (c0 <= 0 || (c0 & 63) || (c1 & 63) || (cout & 63) || B <= 0 || H <= 0 || W <= 0) return PDA_ERR_SHAPE;
  cudaStream_t stream = (cudaStream_t)stream_;
  const int ctot = c0 + c1;
  WgradArgs a;
  a.B = B; a.H = H; a.W = W; a.c0 = c0; a.c1 = c1; a.cout = cout;
  a.tile_w = (W > 8) ? 16 : 8;
  a.tile_h = 128 / a.tile_w;
  a.tiles_x = (W + a.tile_w - 1) / a.tile_w;
  a.tiles_y = (H + a.tile_h - 1) / a.tile_h;
  a.num_tiles = a.tiles_x * a.tiles_y * B;
  a.npairs = 9 * (ctot >> 6);
  a.scratch = scratch;
  const int bn = (cout % 128 == 0) ? 128 : 64;
  const int n_blocks = cout / bn;
  const int m_tiles = (a.npairs + 1) / 2;
  int ksplit = (148 * 2 + m_tiles * n_blocks - 1) / (m_tiles * n_blocks);
  if (ksplit > a.num_tiles) ksplit = a.num_tiles;
  if (ksplit < 1) ksplit = 1;
  if (ksplit > 65535) ksplit = 65535;
  CUtensorMap tX0, tX1, tDZ;
  int r = make_act_tensor_map(&tX0, src0, B, H, W, c0, a.tile_w, a.tile_h, 64);
  if (r) return r;
  if (c1 > 0) {
    r = make_act_tensor_map(&tX1, src1, B, H, W, c1, a.tile_w, a.tile_h, 64);
    if (r) return r;
  } else {
    tX1 = tX0;
  }
  r = make_act_tensor_map(&tDZ, dz, B, H, W, cout, a.tile_w, a.tile_h, 64);
  if (r) return r;
  if (cudaMemsetAsync(scratch, 0, sizeof(float) * 9ull * cout * ctot, stream) != cudaSuccess) return PDA_ERR_CUDA;
  r = (bn == 128) ? launch_wgrad<128, 3>(tX0, tX1, tDZ, a, n_blocks, ksplit, stream)
                  : launch_wgrad<64, 4>(tX0, tX1, tDZ, a, n_blocks, ksplit, stream);
  if (r) return r;
  const long long n = 9LL * cout * ctot;
  PDA_COUNT(1);
  wgrad_scatter_kernel<<<(int)((n + 255) / 256 > 148 * 8 ? 148 * 8 : (n + 255) / 256), 256, 0, stream>>>(
      scratch, dw_oihw, cout, ctot, accumulate);
  if (dbias) {
    if (!accumulate && cudaMemsetAsync(dbias, 0, sizeof(float) * cout, stream) != cudaSuccess) return PDA_ERR_CUDA;
    const int C2 = cout / 2;
    if (256 % C2 && C2 < 256) return PDA_ERR_SHAPE;
    const int threads = C2 >= 256 ? C2 : 256;
    const long long npix = (long long)B * H * W;
    int blocks = (int)((npix + 63) / 64);
    if (blocks > 148 * 4) blocks = 148 * 4;
    if (threads > 1024) return PDA_ERR_SHAPE;
    PDA_COUNT(1);
    bias_grad_kernel<<<blocks, threads, 0, stream>>>(static_cast<const __nv_bfloat162*>(dz), dbias, npix, C2);
  }
  return cudaGetLastError() == cudaSuccess ? PDA_OK : PDA_ERR_CUDA;
}
