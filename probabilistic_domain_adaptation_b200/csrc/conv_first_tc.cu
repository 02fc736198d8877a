// First conv layer (cin 1 or 2 fp32 planes -> 64 NHWC channels + bias + ReLU) on the tensor core.
//
// Reference: the first Conv2d + ReLU of every DownConvBlock stack (unet_blocks.py:17-24 through unet.py:22-31 and
// probabilistic_unet.py:53-61).
//
// The CUDA-core kernel (conv_first_kernel, misc_kernels.cu) needs 9 x 64 FMAs per pixel and is bound by the FMA pipe at
// 44 % of the HBM roofline of its 128 B/px output.  Here the layer is an M = 128 pixels x N = 64 x K GEMM of kind::tf32
// with fp32-equivalent accuracy by operand splitting: x = x_hi + x_lo, w = w_hi + w_lo with *_hi exactly representable
// in tf32, and  x*w ~= x_hi*w_hi + x_lo*w_hi + x_hi*w_lo  (the dropped x_lo*w_lo term and the tf32 rounding of the lo
// parts are ~2^-21 of |x||w| per product: measured <= 3e-5 absolute against the fp32 FMA kernel on outputs of magnitude 5,
// two orders below the 16-bit rounding of the stored result).  K layout for
// T = 9 cin taps:  [x_hi (T) | x_lo (T) | x_hi (T) | 1 | 1 | 0 ..]  against  [w_hi | w_hi | w_lo | b_hi | b_lo | 0 ..],
// padded to 32 (cin 1) / 64 (cin 2): 4 / 8 MMAs of K = 8 per 128 pixels -- far below the HBM time of the tile.
//
// Warp-specialised persistent CTA (2 per SM), the plumbing of fcomb_tc.cu: warps 0-3 build the A rows (one pixel per
// thread: 9 coalesced loads per input plane, split, tcgen05.st into a two-deep TMEM ring -- the operand never touches
// shared memory), warp 8 issues the MMAs (A from tensor memory, B = the split weights in shared memory), warps 4-7 drain
// the two-deep accumulator ring (tcgen05.ld -> ReLU -> 16-bit pack -> one 128-byte line per pixel).
#include "conv.cuh"
#include "ptx.cuh"

namespace pda {

constexpr int CF_TILE = 128;
constexpr int CF_THREADS = 288;
constexpr int CF_COUT = 64;

__host__ __device__ constexpr uint32_t umma_idesc_tf32(int m, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}
// D[tmem] (+)= A[tmem, one 32-bit column per k] * B[smem], kind::tf32, K = 8 per instruction
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// x rounded to tf32 (10-bit mantissa, low 13 bits zero): exactly representable as an MMA operand
__device__ __forceinline__ float tf32_hi(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

template <int CIN>
struct CfCfg {
  static constexpr int T = 9 * CIN;                // taps
  static constexpr int KP = CIN == 1 ? 32 : 64;    // padded K
  static constexpr int KSTEPS = KP / 8;
  static constexpr int D_COL = 0;                  // accumulator ring: 2 x 64 columns
  static constexpr int A_COL = 2 * CF_COUT;        // operand ring: 2 x KP columns
  static constexpr int TMEM_COLS = 256;
  static constexpr int B_BYTES = CF_COUT * KP * 4; // KP / 32 sub-tiles of [64 n][32 k] fp32, K-major SWIZZLE_128B
  static constexpr int STG_OFF = B_BYTES;          // 4 epilogue warps x 4 KB: transposes "one pixel per thread" into
                                                   // coalesced 512-byte store instructions
  static constexpr int BAR_OFF = STG_OFF + 4 * 4096;
  static constexpr int SLOT_OFF = BAR_OFF + 8 * 8;
  static constexpr int SMEM = SLOT_OFF + 16 + 1024;
  static_assert(3 * T + 2 <= KP && A_COL + 2 * KP <= TMEM_COLS, "K padding / tensor memory budget");
};

template <int CIN, bool F16>
__global__ void __launch_bounds__(CF_THREADS, 2)
conv_first_tc_kernel(const float* __restrict__ x0, const float* __restrict__ x1, const float* __restrict__ w,
                     const float* __restrict__ bias, uint4* __restrict__ out, int B, int H, int W, int relu,
                     int num_tiles) {
  using C = CfCfg<CIN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar0 = sbase + C::BAR_OFF;
  auto a_full = [&](int i) { return bar0 + 8u * i; };
  auto a_empty = [&](int i) { return bar0 + 8u * (2 + i); };
  auto d_full = [&](int i) { return bar0 + 8u * (4 + i); };
  auto d_empty = [&](int i) { return bar0 + 8u * (6 + i); };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + C::SLOT_OFF);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(a_full(i), 4);
      mbar_init(a_empty(i), 1);
      mbar_init(d_full(i), 1);
      mbar_init(d_empty(i), 4);
    }
    fence_mbar_init();
  }
  if (warp == 8) {
    tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), C::TMEM_COLS);
    tmem_relinquish();
  }
  // split weights + bias -> B operand [64 n][KP k] (see the header for the k layout)
  for (int i = tid; i < CF_COUT * C::KP; i += CF_THREADS) {
    const int n = i / C::KP, k = i - n * C::KP;
    float v = 0.f;
    if (k < 3 * C::T) {
      const float wv = w[n * C::T + (k % C::T)];
      const float hi = tf32_hi(wv);
      v = (k < 2 * C::T) ? hi : tf32_hi(wv - hi);
    } else if (k < 3 * C::T + 2) {
      const float bv = bias[n];
      const float hi = tf32_hi(bv);
      v = (k == 3 * C::T) ? hi : tf32_hi(bv - hi);
    }
    const int sub = k >> 5, kk = k & 31;
    *reinterpret_cast<float*>(smem + sub * (CF_COUT * 128) + n * 128 + (((kk >> 2) ^ (n & 7)) << 4) + (kk & 3) * 4) = v;
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const long long npix = (long long)B * H * W;

  if (warp < 4) {
    // ================================================================ producers: one pixel per thread
    const uint32_t lane_addr = tmem + (static_cast<uint32_t>(warp * 32) << 16);
    uint32_t it = 0;
    // the 9 (18) input values of this thread's pixel of one tile; zero outside the image / past the last pixel
    auto load_inputs = [&](int tile, float (&xv)[C::T]) {
      const long long p = (long long)tile * CF_TILE + tid;
#pragma unroll
      for (int t = 0; t < C::T; ++t) xv[t] = 0.f;
      if (tile < num_tiles && p < npix) {
        const long long row = p / W;               // = b * H + y   (one 64-bit division per pixel)
        const int xx = (int)(p - row * W);
        const int y = (int)((unsigned)row % (unsigned)H);   // B * H < 2^31 (checked by the host)
        const long long img_off = (row - y) * W;   // first pixel of image b
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci) {
          const float* plane = (ci == 0 ? x0 : x1) + img_off;
#pragma unroll
          for (int ky = 0; ky < 3; ++ky) {
            const int yy = y + ky - 1;
            const bool rowok = yy >= 0 && yy < H;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
              const int xs = xx + kx - 1;
              if (rowok && xs >= 0 && xs < W) xv[ci * 9 + ky * 3 + kx] = __ldg(plane + (long long)yy * W + xs);
            }
          }
        }
      }
    };
    // the loads of the NEXT tile are in flight while this tile's row goes to tensor memory: with one pixel per thread
    // and two CTAs per SM the load latency is otherwise on every tile's critical path
    float xv[C::T], xn[C::T];
    load_inputs(blockIdx.x, xv);
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      load_inputs(tile + gridDim.x, xn);
      const uint32_t slot = it & 1;
      mbar_wait(a_empty(slot), ((it >> 1) & 1) ^ 1);
      tc_fence_after();
#pragma unroll
      for (int half = 0; half < C::KP / 32; ++half) {
        uint32_t a[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int k = half * 32 + j;             // compile-time after unrolling
          float v = 0.f;
          if (k < C::T) v = tf32_hi(xv[k]);
          else if (k < 2 * C::T) v = tf32_hi(xv[k - C::T] - tf32_hi(xv[k - C::T]));   // (rounded, not truncated by the MMA)
          else if (k < 3 * C::T) v = tf32_hi(xv[k - 2 * C::T]);
          else if (k < 3 * C::T + 2) v = 1.f;
          a[j] = __float_as_uint(v);
        }
        tmem_st32(lane_addr + C::A_COL + slot * C::KP + half * 32, a);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_full(slot));
#pragma unroll
      for (int t = 0; t < C::T; ++t) xv[t] = xn[t];
    }
  } else if (warp < 8) {
    // ================================================================ epilogue: one pixel (128-byte output line) per thread
    const int q = warp & 3;
    const uint32_t lane_addr = tmem + (static_cast<uint32_t>(q * 32) << 16);
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const uint32_t buf = it & 1;
      mbar_wait(d_full(buf), (it >> 1) & 1);
      tc_fence_after();
      uint32_t v0[32], v1[32];
      tmem_ld32(lane_addr + C::D_COL + buf * CF_COUT, v0);
      tmem_ld32(lane_addr + C::D_COL + buf * CF_COUT + 32, v1);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(d_empty(buf));
      // ReLU + 16-bit pack into the warp's staging block (row = pixel, 16-byte chunks XOR-swizzled: conflict-free both
      // ways), then the 4 KB block -- contiguous in NHWC -- goes out as eight fully coalesced store instructions.
      // (Storing each thread's own 128-byte line directly makes every store instruction touch 32 lines: measured
      // 235 us for 4 x 1024^2, slower than the CUDA-core kernel.)
      uint8_t* stg = smem + C::STG_OFF + q * 4096;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t* v = j < 4 ? v0 + 8 * j : v1 + 8 * (j - 4);
        float f[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          f[i] = __uint_as_float(v[i]);
          if (relu) f[i] = fmaxf(f[i], 0.f);
        }
        uint4 pk;
        pk.x = pack_act2<F16>(f[0], f[1]);
        pk.y = pack_act2<F16>(f[2], f[3]);
        pk.z = pack_act2<F16>(f[4], f[5]);
        pk.w = pack_act2<F16>(f[6], f[7]);
        *reinterpret_cast<uint4*>(stg + lane * 128 + ((j ^ (lane & 7)) << 4)) = pk;
      }
      __syncwarp();
      const long long p0 = (long long)tile * CF_TILE + q * 32;   // first pixel of this warp's block
      uint4* o = out + p0 * 8;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int c = i * 32 + lane, px = c >> 3, j = c & 7;
        const uint4 pk = *reinterpret_cast<const uint4*>(stg + px * 128 + ((j ^ (px & 7)) << 4));
        if (p0 + px < npix) o[c] = pk;
      }
      __syncwarp();
    }
  } else {
    // ================================================================ control: every tcgen05.mma
    const bool leader = elect_one();
    constexpr uint32_t idesc = umma_idesc_tf32(128, CF_COUT);
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const uint32_t slot = it & 1;
      mbar_wait(a_full(slot), (it >> 1) & 1);
      mbar_wait(d_empty(slot), ((it >> 1) & 1) ^ 1);
      tc_fence_after();
      if (leader) {
        const uint32_t d = tmem + C::D_COL + slot * CF_COUT;
        const uint32_t ta = tmem + C::A_COL + slot * C::KP;
#pragma unroll
        for (int j = 0; j < C::KSTEPS; ++j) {
          // K-step j: 8 tf32 = 32 bytes inside the 128-byte row of sub-tile j / 4
          const uint64_t db = umma_desc_k_sw128(sbase + (j >> 2) * (CF_COUT * 128)) + 2ull * (j & 3);
          umma_tf32_ts(d, ta + 8 * j, db, idesc, j ? 1u : 0u);
        }
        umma_commit(d_full(slot));
        umma_commit(a_empty(slot));
      }
      __syncwarp();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem, C::TMEM_COLS);
}

// same contract as pda_conv3x3_first (cout = 64 only); selected by it
int conv3x3_first_tc(const float* x0, const float* x1, const float* w, const float* bias, void* out, int B, int H, int W,
                     int relu, int act_f16, cudaStream_t st) {
  const long long npix = (long long)B * H * W;
  const long long tiles = (npix + CF_TILE - 1) / CF_TILE;
  if (tiles <= 0 || tiles > 0x7fffffffLL || (long long)B * H > 0x7fffffffLL) return PDA_ERR_SHAPE;
  const int grid = (int)(tiles < 148 * 2 ? tiles : 148 * 2);
  uint4* o = static_cast<uint4*>(out);
  if (x1) {
    if (act_f16) conv_first_tc_kernel<2, true><<<grid, CF_THREADS, CfCfg<2>::SMEM, st>>>(x0, x1, w, bias, o, B, H, W, relu, (int)tiles);
    else conv_first_tc_kernel<2, false><<<grid, CF_THREADS, CfCfg<2>::SMEM, st>>>(x0, x1, w, bias, o, B, H, W, relu, (int)tiles);
  } else {
    if (act_f16) conv_first_tc_kernel<1, true><<<grid, CF_THREADS, CfCfg<1>::SMEM, st>>>(x0, x1, w, bias, o, B, H, W, relu, (int)tiles);
    else conv_first_tc_kernel<1, false><<<grid, CF_THREADS, CfCfg<1>::SMEM, st>>>(x0, x1, w, bias, o, B, H, W, relu, (int)tiles);
  }
  return cudaGetLastError() == cudaSuccess ? PDA_OK : PDA_ERR_CUDA;
}

}  // namespace pda
