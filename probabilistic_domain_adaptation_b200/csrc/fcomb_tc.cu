// Tensor-core version of the fused Fcomb + sigmoid + cross-sample mean + consensus kernel.
//
// Reference: S x Fcomb.forward (/root/reference/prob_utils/my_models/probabilistic_unet.py:200-214) + the consensus
// arithmetic of prob_utils/my_trainer/mean_teacher_trainer.py:74-86.
//
// The kernel is COMPUTE bound (~1000 FLOP/B at S=16), so the two 64x64 layers run on tcgen05:
//   per 128-pixel tile:  H1 = F . (W1f_hi + W1f_lo)^T         128x64x64 MMA x2 (bf16 hi/lo split of the fp32 weights
//                                                            -> ~16-bit mantissa), accumulator in TMEM, read ONCE
//   per sample s:        A1_s = relu(H1 + bz_s) -> fp16 (saturating) -> swizzled smem operand   (CUDA cores, f32x2)
//                        H2_s = A1_s . W2^T                  128x64x64 fp16 MMA into one of two TMEM buffers
//                        logit = w3 . relu(H2_s + b2) + b3 ; p = sigmoid ; mean / consensus in registers
// bz_s = b1 + W1z . z_s is the per-(sample, image) bias that replaces the tiled-z concat.  MMA s+1 is issued before
// the epilogue of sample s, so tensor pipe, TMEM loads and CUDA-core math overlap; 4 CTAs share an SM.
#include "conv.cuh"
#include "ptx.cuh"

namespace pda {

constexpr int FCT = 64;             // feature / hidden channels
constexpr int FC_TILE = 128;        // pixels per tile == TMEM lanes
constexpr int FC_TMEM_COLS = 128;   // [0,64): H1 then H2 buffer 0; [64,128): H2 buffer 1

struct FcombSmem {
  static constexpr int A_BYTES = FC_TILE * 128;                 // 128 rows x 64 bf16
  static constexpr int W_BYTES = FCT * 128;                     // 64 rows x 64 bf16
  static constexpr int A0_OFF = 0;                              // F tile, later A1 buffer 0
  static constexpr int A1_OFF = A_BYTES;                        // A1 buffer 1
  static constexpr int W1_OFF = 2 * A_BYTES;                    // bf16 hi part of W1[:, :64]
  static constexpr int W1L_OFF = W1_OFF + W_BYTES;              // bf16 lo part (w - hi)
  static constexpr int W2_OFF = W1L_OFF + W_BYTES;              // fp16 W2
  static constexpr int VEC_OFF = W2_OFF + W_BYTES;              // b2[64], w3[64] fp32
  static constexpr int BAR_OFF = VEC_OFF + 2 * FCT * 4;         // 4 mbarriers
  static constexpr int SLOT_OFF = BAR_OFF + 4 * 8;
  static constexpr int BZ_OFF = SLOT_OFF + 16;                  // bz[S][64] fp32
  static int bytes(int S) { return BZ_OFF + S * FCT * 4 + 1024; }
};

__device__ __forceinline__ uint64_t pack_f32x2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack_f32x2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t add_f32x2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
// {hi, lo} -> f16x2 with ReLU, saturating to the largest finite fp16 (lo in the low half)
__device__ __forceinline__ uint32_t relu_pack_f16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// fp32 [64][ld] weights (first 64 columns) -> K-major SWIZZLE_128B operand tiles in shared memory:
// bf16 hi + bf16 lo (w ~= hi + lo) when dst_lo != nullptr, else a single fp16 tile.
__device__ __forceinline__ void stage_weight_sw128(uint8_t* dst, uint8_t* dst_lo, const float* __restrict__ w, int ld) {
  for (int i = threadIdx.x; i < FCT * 8; i += blockDim.x) {
    const int n = i >> 3, c = i & 7;  // row n, 16-byte chunk c (8 k-values)
    const float* src = w + n * ld + c * 8;
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float a = src[2 * j], b = src[2 * j + 1];
      if (dst_lo) {
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

__global__ void __launch_bounds__(128, 4)
fcomb_tc_kernel(const __grid_constant__ CUtensorMap tmF, const float* __restrict__ z, const float* __restrict__ w1,
                const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ b2,
                const float* __restrict__ w3, const float* __restrict__ b3, int P, int S, int L, int B,
                int tiles_per_img, int num_tiles, float upper, float lower, float* __restrict__ mean_prob,
                float* __restrict__ cons_weight, int64_t* __restrict__ cons_mask, float* __restrict__ logits,
                float* __restrict__ probs) {
  using M = FcombSmem;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sbase = smem_u32(smem);
  const uint32_t barF = sbase + M::BAR_OFF, barM = barF + 8, barH0 = barF + 16, barH1 = barF + 24;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + M::SLOT_OFF);
  float* b2s = reinterpret_cast<float*>(smem + M::VEC_OFF);
  float* w3s = b2s + FCT;
  float* bzs = reinterpret_cast<float*>(smem + M::BZ_OFF);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int kin = FCT + L;

  if (tid == 0) {
    mbar_init(barF, 1);
    mbar_init(barM, 1);
    mbar_init(barH0, 1);
    mbar_init(barH1, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmF);
  }
  if (warp == 0) {
    tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), FC_TMEM_COLS);
    tmem_relinquish();
  }
  stage_weight_sw128(smem + M::W1_OFF, smem + M::W1L_OFF, w1, kin);
  stage_weight_sw128(smem + M::W2_OFF, nullptr, w2, FCT);
  if (tid < FCT) {
    b2s[tid] = b2[tid];
    w3s[tid] = w3[tid];
  }
  fence_proxy_async_smem();  // weight tiles were written by the generic proxy, read by the tensor core
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t lane_addr = tmem + (static_cast<uint32_t>(warp * 32) << 16);
  const float b3v = b3[0];
  constexpr uint32_t idesc1 = umma_idesc_bf16(128, FCT);  // F (bf16) x W1 hi/lo (bf16)
  constexpr uint32_t idesc2 = umma_idesc_f16(128, FCT);   // A1 (fp16) x W2 (fp16)
  const uint64_t dW1 = umma_desc_k_sw128(sbase + M::W1_OFF);
  const uint64_t dW1L = umma_desc_k_sw128(sbase + M::W1L_OFF);
  const uint64_t dW2 = umma_desc_k_sw128(sbase + M::W2_OFF);
  const uint64_t dA0 = umma_desc_k_sw128(sbase + M::A0_OFF), dA1 = umma_desc_k_sw128(sbase + M::A1_OFF);
  uint8_t* const arow0 = smem + M::A0_OFF + tid * 128;
  uint8_t* const arow1 = smem + M::A1_OFF + tid * 128;
  const int sw = tid & 7;
  uint32_t phF = 0, phM = 0, phH0 = 0, phH1 = 0;

  for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const int b = tile / tiles_per_img;
    const int p0 = (tile - b * tiles_per_img) * FC_TILE;
    // ---- F tile -> smem (TMA, rows beyond P are zero-filled), H1 = F . W1f^T
    if (tid == 0) {
      mbar_expect_tx(barF, M::A_BYTES);
      tma_load_3d(sbase + M::A0_OFF, &tmF, barF, 0, p0, b);
    }
    // per-(sample, image) bias of layer 1 while the tile is in flight
    for (int i = tid; i < S * FCT; i += blockDim.x) {
      const int s = i / FCT, j = i - s * FCT;
      float acc = b1[j];
      const float* zr = z + (static_cast<size_t>(s) * B + b) * L;
      for (int d = 0; d < L; ++d) acc = fmaf(w1[j * kin + FCT + d], zr[d], acc);
      bzs[i] = acc;
    }
    if (tid == 0) {
      mbar_wait(barF, phF);
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16(tmem, dA0 + 2 * k, dW1 + 2 * k, idesc1, k ? 1u : 0u);
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16(tmem, dA0 + 2 * k, dW1L + 2 * k, idesc1, 1u);
      umma_commit(barM);
    }
    phF ^= 1;
    __syncthreads();  // bzs visible
    mbar_wait(barM, phM);
    phM ^= 1;
    tc_fence_after();
    // ---- H1 -> registers (kept for all samples), as packed f32x2
    uint64_t h1[FCT / 2];
    {
      uint32_t v[32];
      tmem_ld32(lane_addr, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i) h1[i] = (static_cast<uint64_t>(v[2 * i + 1]) << 32) | v[2 * i];
      tmem_ld32(lane_addr + 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i) h1[16 + i] = (static_cast<uint64_t>(v[2 * i + 1]) << 32) | v[2 * i];
    }

    auto produce = [&](int s) {
      // A1_s = relu(H1 + bz_s) as fp16, this thread's 128-byte row, 16-byte chunks XOR-swizzled by (row & 7)
      const float4* bz4 = reinterpret_cast<const float4*>(bzs + s * FCT);
      uint8_t* row = (s & 1) ? arow1 : arow0;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float4 ba = bz4[2 * c], bb = bz4[2 * c + 1];
        float x0, x1, x2, x3, x4, x5, x6, x7;
        unpack_f32x2(add_f32x2(h1[4 * c + 0], pack_f32x2(ba.x, ba.y)), x0, x1);
        unpack_f32x2(add_f32x2(h1[4 * c + 1], pack_f32x2(ba.z, ba.w)), x2, x3);
        unpack_f32x2(add_f32x2(h1[4 * c + 2], pack_f32x2(bb.x, bb.y)), x4, x5);
        unpack_f32x2(add_f32x2(h1[4 * c + 3], pack_f32x2(bb.z, bb.w)), x6, x7);
        uint4 o;
        o.x = relu_pack_f16x2(x0, x1);
        o.y = relu_pack_f16x2(x2, x3);
        o.z = relu_pack_f16x2(x4, x5);
        o.w = relu_pack_f16x2(x6, x7);
        *reinterpret_cast<uint4*>(row + ((c ^ sw) << 4)) = o;
      }
    };
    auto issue = [&](int s) {
      // all 128 threads: publish smem writes to the async proxy, order prior TMEM reads, then one thread issues
      fence_proxy_async_smem();
      tc_fence_before();
      named_bar_sync(1, 128);
      if (tid == 0) {
        tc_fence_after();
        const uint32_t d = tmem + (s & 1) * FCT;
        const uint64_t da = (s & 1) ? dA1 : dA0;
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(d, da + 2 * k, dW2 + 2 * k, idesc2, k ? 1u : 0u);
        umma_commit((s & 1) ? barH1 : barH0);
      }
    };

    const int pix = p0 + tid;
    const bool valid = pix < P;
    const size_t gp = static_cast<size_t>(b) * P + pix;
    float psum = 0.f;
    int count = 0;
    produce(0);
    issue(0);
    for (int s = 0; s < S; ++s) {
      if (s + 1 < S) {
        produce(s + 1);
        issue(s + 1);
      }
      const int buf = s & 1;
      if (buf) {
        mbar_wait(barH1, phH1);
        phH1 ^= 1;
      } else {
        mbar_wait(barH0, phH0);
        phH0 ^= 1;
      }
      tc_fence_after();
      float logit = b3v;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t v[32];
        tmem_ld32(lane_addr + buf * FCT + half * 32, v);
        tmem_ld_wait();
        const float4* b24 = reinterpret_cast<const float4*>(b2s + half * 32);
        const float4* w34 = reinterpret_cast<const float4*>(w3s + half * 32);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 bb = b24[i], ww = w34[i];
          logit = fmaf(ww.x, fmaxf(__uint_as_float(v[4 * i + 0]) + bb.x, 0.f), logit);
          logit = fmaf(ww.y, fmaxf(__uint_as_float(v[4 * i + 1]) + bb.y, 0.f), logit);
          logit = fmaf(ww.z, fmaxf(__uint_as_float(v[4 * i + 2]) + bb.z, 0.f), logit);
          logit = fmaf(ww.w, fmaxf(__uint_as_float(v[4 * i + 3]) + bb.w, 0.f), logit);
        }
      }
      const float pr = 1.0f / (1.0f + expf(-logit));
      psum += pr;
      count += (pr >= upper || pr <= lower) ? 1 : 0;
      if (valid) {
        if (logits) logits[(static_cast<size_t>(s) * B + b) * P + pix] = logit;
        if (probs) probs[(static_cast<size_t>(s) * B + b) * P + pix] = pr;
      }
    }
    if (valid) {
      if (mean_prob) mean_prob[gp] = psum / static_cast<float>(S);
      if (cons_weight) cons_weight[gp] = static_cast<float>(count) / static_cast<float>(S);
      if (cons_mask) cons_mask[gp] = (count == S) ? 1 : 0;
    }
    // every thread has finished reading TMEM / bzs before the next tile overwrites them
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, FC_TMEM_COLS);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace pda

using namespace pda;

extern "C" int pda_fcomb_mc_consensus(const void* feat, const float* z, const float* w1, const float* b1,
                                      const float* w2, const float* b2, const float* w3, const float* b3, int B, int P,
                                      int S, int latent, float upper, float lower, float* mean_prob,
                                      float* cons_weight, int64_t* cons_mask, float* logits, float* probs,
                                      void* stream) {
  if (!feat || !z || !w1 || !b1 || !w2 || !b2 || !w3 || !b3) return PDA_ERR_ARG;
  if (B <= 0 || P <= 0 || S <= 0 || latent <= 0) return PDA_ERR_SHAPE;
  const int smem = FcombSmem::bytes(S);
  if (smem > 220 * 1024) return PDA_ERR_SHAPE;
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
  static int configured = 0;
  if (smem > configured) {
    if (cudaFuncSetAttribute(fcomb_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
      return PDA_ERR_CUDA;
    configured = smem;
  }
  const int tiles_per_img = (P + FC_TILE - 1) / FC_TILE;
  const long long num_tiles = (long long)tiles_per_img * B;
  if (num_tiles > 0x7fffffffLL) return PDA_ERR_SHAPE;
  const int grid = (int)(num_tiles < 148 * 4 ? num_tiles : 148 * 4);
  PDA_COUNT(1);
  fcomb_tc_kernel<<<grid, 128, smem, (cudaStream_t)stream>>>(tm, z, w1, b1, w2, b2, w3, b3, P, S, latent, B,
                                                             tiles_per_img, (int)num_tiles, upper, lower, mean_prob,
                                                             cons_weight, cons_mask, logits, probs);
  return cudaGetLastError() == cudaSuccess ? PDA_OK : PDA_ERR_CUDA;
}
