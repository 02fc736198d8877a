// conv3x3 (pad 1, stride 1) + bias + ReLU (+ fused 2x2 average pool) as an implicit GEMM on the
// 5th-gen tensor cores: TMA (SWIZZLE_128B tiles, zero-filled out-of-bounds = the conv padding)
// -> shared memory ring -> tcgen05.mma (bf16 x bf16 -> fp32 in TMEM) -> tcgen05.ld epilogue.
//
// Replaces the cuDNN calls behind nn.Conv2d(k=3, padding=1)+ReLU(+AvgPool2d) of
//   /root/reference/prob_utils/my_models/unet_blocks.py:17-24   (DownConvBlock)
//   /root/reference/prob_utils/my_models/probabilistic_unet.py:53-61 (Encoder)
// and the torch.cat([up, bridge]) + conv of unet_blocks.py:56-57 (two K segments, no concat tensor).
//
// GEMM view:  D[pixel, co] = sum_{tap, ci} A[pixel shifted by tap, ci] * Wp[co, tap*Ctot + ci]
//   M tile = 128 output pixels = a tile_h x tile_w spatial patch of one image (tile_w in {8, 16})
//   N tile = BN output channels, K step = 64 input channels of one tap (one 128-byte swizzle row).
// Activations: NHWC bf16.  Weights: packed [Cout][9][Ctot] bf16 (K-major).
#include "conv.cuh"
#include "ptx.cuh"

namespace pda {

template <int BN, int STAGES>
struct ConvSmem {
  static constexpr int A_BYTES = 128 * 128;      // 128 pixels x 64 ch bf16
  static constexpr int B_BYTES = BN * 128;       // BN out-channels x 64 ch bf16
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int NBARS = 2 * STAGES + 1;
  static constexpr int TMEM_SLOT_OFF = BAR_OFF + NBARS * 8;
  static constexpr int BIAS_OFF = TMEM_SLOT_OFF + 16;
  static constexpr int TOTAL = BIAS_OFF + BN * 4;
  static constexpr int DYN_BYTES = TOTAL + 1024;  // slack for manual 1024-B alignment
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(192, 1)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                  const __grid_constant__ CUtensorMap tmB, const ConvArgs p) {
  using L = ConvSmem<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bar_base = smem_base + L::BAR_OFF;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  const uint32_t tmem_full_bar = bar_base + 8u * (2 * STAGES);
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + L::TMEM_SLOT_OFF);
  float* bias_s = reinterpret_cast<float*>(smem + L::BIAS_OFF);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // tile coordinates
  const int tile = blockIdx.x;
  const int tx = tile % p.tiles_x;
  const int ty = (tile / p.tiles_x) % p.tiles_y;
  const int img = tile / (p.tiles_x * p.tiles_y);
  const int x0 = tx * p.tile_w, y0 = ty * p.tile_h;
  const int n0 = blockIdx.y * BN;
  const int ctot = p.c0 + p.c1;
  const int chunks = ctot >> 6;
  const int num_k = 9 * chunks;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), BN < 32 ? 32 : BN);
    tmem_relinquish();
  }
  if (warp >= 2) {
    for (int i = threadIdx.x - 64; i < BN; i += 128) bias_s[i] = p.bias ? p.bias[n0 + i] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (one lane)
    if (lane == 0) {
      for (int it = 0; it < num_k; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        mbar_wait(empty_bar(s), ph ^ 1);
        mbar_expect_tx(full_bar(s), L::STAGE_BYTES);
        const int tap = it / chunks;
        const int c = (it - tap * chunks) << 6;
        const int ky = tap / 3, kx = tap - 3 * ky;
        const uint32_t sa = smem_base + s * L::STAGE_BYTES;
        const uint32_t sb = sa + L::A_BYTES;
        if (c < p.c0)
          tma_load_4d(sa, &tmA0, full_bar(s), c, x0 + kx - 1, y0 + ky - 1, img);
        else
          tma_load_4d(sa, &tmA1, full_bar(s), c - p.c0, x0 + kx - 1, y0 + ky - 1, img);
        tma_load_2d(sb, &tmB, full_bar(s), tap * ctot + c, n0);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (one lane)
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, BN);
      for (int it = 0; it < num_k; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        const uint32_t sa = smem_base + s * L::STAGE_BYTES;
        const uint64_t da = umma_desc_k_sw128(sa);
        const uint64_t db = umma_desc_k_sw128(sa + L::A_BYTES);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          // advance 16 bf16 = 32 bytes along K inside the 128-byte swizzle row: +2 in 16-byte units
          umma_bf16(tmem_base, da + 2 * k, db + 2 * k, idesc, (it | k) != 0 ? 1u : 0u);
        }
        umma_commit(empty_bar(s));  // frees the smem stage when these MMAs retire
      }
      umma_commit(tmem_full_bar);
    }
  } else {
    // ------------------------------------------------------------ epilogue: 4 warps, one output pixel per thread
    const int q = warp & 3;  // TMEM lane quarter this warp may read
    const int row = q * 32 + lane;
    const int lty = row / p.tile_w, ltx = row - lty * p.tile_w;
    const int y = y0 + lty, x = x0 + ltx;
    const bool valid = (y < p.H) && (x < p.W);
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    __nv_bfloat16* out_px = p.out ? p.out + ((static_cast<size_t>(img) * p.H + y) * p.W + x) * p.cout + n0 : nullptr;
    const int Hp = p.H >> 1, Wp = p.W >> 1;
    const bool pool_owner = valid && !(lty & 1) && !(ltx & 1);
    __nv_bfloat16* pool_px =
        p.out_pool ? p.out_pool + ((static_cast<size_t>(img) * Hp + (y >> 1)) * Wp + (x >> 1)) * p.cout + n0 : nullptr;
#pragma unroll 1
    for (int cb = 0; cb < BN / 32; ++cb) {
      uint32_t v[32];
      tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + cb * 32, v);
      tmem_ld_wait();
      float f[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float t = __uint_as_float(v[j]) + bias_s[cb * 32 + j];
        f[j] = p.relu ? fmaxf(t, 0.f) : t;
      }
      if (out_px && valid) {
        uint4* dst = reinterpret_cast<uint4*>(out_px + cb * 32);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 o;
          o.x = pack_bf16x2(f[8 * j + 0], f[8 * j + 1]);
          o.y = pack_bf16x2(f[8 * j + 2], f[8 * j + 3]);
          o.z = pack_bf16x2(f[8 * j + 4], f[8 * j + 5]);
          o.w = pack_bf16x2(f[8 * j + 6], f[8 * j + 7]);
          dst[j] = o;
        }
      }
      if (pool_px) {
        // 2x2 average of the fp32 post-ReLU values: partners are lane^1 (x) and lane^tile_w (y)
        // (H, W even => a valid even pixel always has valid partners).
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float t = valid ? f[j] : 0.f;
          t += __shfl_xor_sync(0xffffffffu, t, 1);
          t += __shfl_xor_sync(0xffffffffu, t, p.tile_w);
          f[j] = 0.25f * t;
        }
        if (pool_owner) {
          uint4* dst = reinterpret_cast<uint4*>(pool_px + cb * 32);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 o;
            o.x = pack_bf16x2(f[8 * j + 0], f[8 * j + 1]);
            o.y = pack_bf16x2(f[8 * j + 2], f[8 * j + 3]);
            o.z = pack_bf16x2(f[8 * j + 4], f[8 * j + 5]);
            o.w = pack_bf16x2(f[8 * j + 6], f[8 * j + 7]);
            dst[j] = o;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, BN < 32 ? 32 : BN);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

void* get_encode_tiled() { return reinterpret_cast<void*>(get_encode_fn()); }

int make_act_tensor_map(CUtensorMap* tm, const void* ptr, int B, int H, int W, int C, int box_w, int box_h,
                        int box_c) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return PDA_ERR_DRIVER;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? PDA_OK : PDA_ERR_TENSORMAP;
}

int make_mat_tensor_map(CUtensorMap* tm, const void* ptr, long long inner, long long outer, int box_inner,
                        int box_outer) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return PDA_ERR_DRIVER;
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)inner * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? PDA_OK : PDA_ERR_TENSORMAP;
}

template <int BN, int STAGES>
static int launch_conv(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b, const ConvArgs& args,
                       cudaStream_t stream) {
  using L = ConvSmem<BN, STAGES>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv3x3_tc_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         L::DYN_BYTES);
    if (e != cudaSuccess) return PDA_ERR_CUDA;
    configured = true;
  }
  dim3 grid(args.tiles_x * args.tiles_y * args.B, args.cout / BN);
  PDA_COUNT(1);
  conv3x3_tc_kernel<BN, STAGES><<<grid, 192, L::DYN_BYTES, stream>>>(a0, a1, b, args);
  return cudaGetLastError() == cudaSuccess ? PDA_OK : PDA_ERR_CUDA;
}

int conv3x3_tc(const void* src0, int c0, const void* src1, int c1, const void* wpacked, const float* bias, void* out,
               void* out_pool, int B, int H, int W, int cout, int relu, int bn_override, cudaStream_t stream) {
  if (c0 <= 0 || (c0 & 63) || (c1 & 63) || (cout & 63) || B <= 0 || H <= 0 || W <= 0) return PDA_ERR_SHAPE;
  if (out_pool && ((H & 1) || (W & 1))) return PDA_ERR_SHAPE;
  ConvArgs a;
  a.B = B; a.H = H; a.W = W; a.c0 = c0; a.c1 = c1; a.cout = cout; a.relu = relu;
  a.tile_w = (W > 8) ? 16 : 8;
  a.tile_h = 128 / a.tile_w;
  a.tiles_x = (W + a.tile_w - 1) / a.tile_w;
  a.tiles_y = (H + a.tile_h - 1) / a.tile_h;
  a.bias = bias;
  a.out = static_cast<__nv_bfloat16*>(out);
  a.out_pool = static_cast<__nv_bfloat16*>(out_pool);
  int bn = bn_override;
  if (bn == 0) bn = (cout % 128 == 0) ? 128 : 64;
  if (cout % bn) return PDA_ERR_SHAPE;
  CUtensorMap tA0, tA1, tB;
  int r = make_act_tensor_map(&tA0, src0, B, H, W, c0, a.tile_w, a.tile_h, 64);
  if (r) return r;
  if (c1 > 0) {
    r = make_act_tensor_map(&tA1, src1, B, H, W, c1, a.tile_w, a.tile_h, 64);
    if (r) return r;
  } else {
    tA1 = tA0;
  }
  r = make_mat_tensor_map(&tB, wpacked, 9LL * (c0 + c1), cout, 64, bn);
  if (r) return r;
  switch (bn) {
    case 64: return launch_conv<64, 4>(tA0, tA1, tB, a, stream);
    case 128: return launch_conv<128, 3>(tA0, tA1, tB, a, stream);
    case 256: return launch_conv<256, 4>(tA0, tA1, tB, a, stream);
    default: return PDA_ERR_SHAPE;
  }
}

}  // namespace pda
