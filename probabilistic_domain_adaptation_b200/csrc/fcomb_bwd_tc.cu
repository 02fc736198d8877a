// Fcomb backward for ONE latent sample per image (the training form) on tcgen05 tensor cores.
//
// Reference: autograd through Fcomb.forward (/root/reference/prob_utils/my_models/probabilistic_unet.py:200-214)
// as called by reconstruct() in elbo() (:356-358).
//
// Per 128-pixel tile (thread = pixel row, accumulators in TMEM), five chained GEMMs:
//   G1  H1  = F . W1f^T              (bf16 hi + lo weights)        a1  = relu(H1 + bz)      bz = b1 + W1z . z[image]
//   G2  H2  = [A1 | 1 1] . [W2 | b2]^T   (fp16, exactly the operands of the forward kernel, so the ReLU masks are
//                                         those of the function that was actually evaluated)
//                                                                  dh2 = (H2 > 0) g w3      dW3 += g relu(H2)
//   G3  dA1 = dH2 . W2                                            dh1 = (a1 > 0) dA1
//   G4  dF  = dH1 . W1f                                           -> bf16, global
//   G5  [dW2 x; x dW1f] += [dH2 | dH1]^T . [A1 | F]   one M = 128, N = 128, K = 128 pixels MMA chain (MN-major
//       operands read from the very tiles the other GEMMs use; a bf16 copy of A1, because one MMA cannot mix fp16 and
//       bf16 operands), accumulated over ALL tiles of the CTA
//   G6  [db2 ; dbz] += [dH2 | dH1]^T . 1              N = 16 "ones" block (column sums), flushed per image
// A CTA runs two warpgroups, each on its own tile sequence, so that one group's CUDA-core phase overlaps the other
// group's MMAs; both accumulate into the same weight-gradient columns.
#include <cuda_fp16.h>

#include "conv.cuh"
#include "ptx.cuh"

namespace pda {

constexpr int FB_C = 64;
constexpr int FB_TILE_BYTES = 128 * 128;  // [128 px][64 ch] bf16
constexpr int FB_W_BYTES = 64 * 128;

struct FbSmem {
  // per warpgroup: A1 (bf16 copy), F (adjacent: the N operand of G5), DH2, DH1 (adjacent: the M operand of G5),
  // A1H (fp16: the A operand of G2 and the ReLU mask of layer 1)
  static constexpr int WG_BYTES = 5 * FB_TILE_BYTES;
  static constexpr int A1 = 0, F = FB_TILE_BYTES, DH2 = 2 * FB_TILE_BYTES, DH1 = 3 * FB_TILE_BYTES,
                       A1H = 4 * FB_TILE_BYTES;
  static constexpr int W1H = 2 * WG_BYTES, W1L = W1H + FB_W_BYTES, W2 = W1L + FB_W_BYTES, W2X = W2 + FB_W_BYTES,
                       W2T = W2X + FB_W_BYTES, W1T = W2T + FB_W_BYTES;
  static constexpr int AX = W1T + FB_W_BYTES;     // K-major "ones" A rows for the bias extension (1 KB)
  static constexpr int ONES = AX + 1024;          // MN-major all-ones B block (1 KB)
  static constexpr int BAR = ONES + 1024;         // 2 x (mbarF, mbarM)
  static constexpr int SLOT = BAR + 4 * 8;
  static constexpr int BZ = SLOT + 16;            // [2][64] fp32
  static constexpr int RED = BZ + 2 * 64 * 4;     // dW3 reduction [64] + db3 [1] (padded to 68 floats)
  static constexpr int W3 = RED + 68 * 4;         // last layer w3[64] fp32 (per-launch copy: no process-wide state)
  static constexpr int TOTAL = W3 + 64 * 4;
  static constexpr int DYN_BYTES = TOTAL + 1024;
};

__device__ __forceinline__ uint32_t fb_pack2(float lo, float hi) { return pack_bf16x2(lo, hi); }

__device__ __forceinline__ uint32_t fb_pack2h(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint32_t fb_relu_pack2h(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// fp32 [64 rows][ld] -> K-major SWIZZLE_128B tile(s): bf16 hi (+ bf16 lo), or a single fp16 tile (half = true).
// transpose: tile[n][k] = w[k][n].
__device__ __forceinline__ void fb_stage(uint8_t* dst, uint8_t* dst_lo, const float* __restrict__ w, int ld,
                                         bool transpose, bool half = false) {
  for (int i = threadIdx.x; i < FB_C * 8; i += blockDim.x) {
    const int n = i >> 3, c = i & 7;
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k0 = c * 8 + 2 * j;
      const float a = transpose ? w[k0 * ld + n] : w[n * ld + k0];
      const float b = transpose ? w[(k0 + 1) * ld + n] : w[n * ld + k0 + 1];
      hi[j] = half ? fb_pack2h(a, b) : fb_pack2(a, b);
      if (dst_lo) {
        const float ah = __bfloat162float(__float2bfloat16(a)), bh = __bfloat162float(__float2bfloat16(b));
        lo[j] = fb_pack2(a - ah, b - bh);
      }
    }
    const int off = n * 128 + ((c ^ (n & 7)) << 4);
    *reinterpret_cast<uint4*>(dst + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    if (dst_lo) *reinterpret_cast<uint4*>(dst_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

__device__ __forceinline__ void fb_named_bar(int id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }

__global__ void __launch_bounds__(256, 1)
fcomb_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmF, const float* __restrict__ bzg,
                    const float* __restrict__ w1, const float* __restrict__ w2, const float* __restrict__ b2,
                    const float* __restrict__ w3, const int* __restrict__ skip_flag,
                    const float* __restrict__ dlogit, int P, int L, int B, int tiles_per_img, int num_tiles,
                    __nv_bfloat16* __restrict__ dfeat, float* __restrict__ dw1f, float* __restrict__ dw2,
                    float* __restrict__ db2, float* __restrict__ dw3, float* __restrict__ db3,
                    float* __restrict__ dbz) {
  using M = FbSmem;
  // the forward took the fp32 path (fp16 range flag): the fp32 backward kernel produces the gradients instead
  if (skip_flag != nullptr && *skip_flag != 0) return;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wg = warp >> 2;              // warpgroup 0 / 1
  const int wt = tid & 127;              // thread inside the warpgroup = pixel row of its tile
  const int q = warp & 3;                // TMEM lane quarter
  const uint32_t mbarF = sbase + M::BAR + 16 * wg, mbarM = mbarF + 8;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + M::SLOT);
  float* bzs = reinterpret_cast<float*>(smem + M::BZ) + 64 * wg;
  float* red = reinterpret_cast<float*>(smem + M::RED);
  const float4* w3s = reinterpret_cast<const float4*>(smem + M::W3);
  uint8_t* tiles = smem + wg * M::WG_BYTES;
  const uint32_t tiles_u32 = sbase + wg * M::WG_BYTES;
  const int kin = FB_C + L;

  if (tid == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(sbase + M::BAR + 8 * i, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmF);
  }
  if (warp == 0) {
    tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), 512);
    tmem_relinquish();
  }
  fb_stage(smem + M::W1H, smem + M::W1L, w1, kin, false);
  fb_stage(smem + M::W2, nullptr, w2, FB_C, false, true);
  fb_stage(smem + M::W2T, nullptr, w2, FB_C, true);
  fb_stage(smem + M::W1T, nullptr, w1, kin, true);
  for (int i = tid; i < FB_C * 8; i += blockDim.x) {  // K-extension of W2: k = 0 -> b2_hi, k = 1 -> b2_lo
    const int n = i >> 3, c = i & 7;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (c == 0) {
      const float b = b2[n];
      const float bh = __half2float(__float2half_rn(b));
      v.x = fb_pack2h(bh, b - bh);
    }
    *reinterpret_cast<uint4*>(smem + M::W2X + n * 128 + ((c ^ (n & 7)) << 4)) = v;
  }
  if (tid < 64) {
    const int n = tid >> 3, c = tid & 7;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (c == 0) v.x = 0x3C003C00u;  // (1.0, 1.0) fp16
    *reinterpret_cast<uint4*>(smem + M::AX + n * 128 + ((c ^ (n & 7)) << 4)) = v;
    *reinterpret_cast<uint4*>(smem + M::ONES + tid * 16) = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
  }
  if (tid < 65) red[tid] = 0.f;
  if (tid >= 128 && tid < 128 + FB_C) reinterpret_cast<float*>(smem + M::W3)[tid - 128] = w3[tid - 128];
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t lane_base = tmem + (static_cast<uint32_t>(q * 32) << 16);
  const uint32_t colHA = wg * 128, colHB = wg * 128 + 64, colDW = 256, colON = 384 + 16 * wg;
  // zero the shared weight-gradient accumulator (both warpgroups accumulate into it from their first tile on)
  if (wg == 0) {
#pragma unroll
    for (int cb = 0; cb < 4; ++cb) tmem_st32_fill(lane_base + colDW + cb * 32, 0u);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  constexpr uint32_t id_bb = umma_idesc_mixed(128, FB_C, 1, 1);        // G1: bf16 x bf16, K-major
  constexpr uint32_t id_hh = umma_idesc_mixed(128, FB_C, 0, 0);        // G2: fp16 x fp16
  constexpr uint32_t id_bh = id_bb;                                    // G3 / G4: bf16 gradients x bf16 weights
  constexpr uint32_t id_w = umma_idesc_mixed(128, 128, 1, 1, 1, 1);    // G5: MN-major bf16 x bf16, N = 128
  constexpr uint32_t id_o = umma_idesc_mixed(128, 16, 1, 1, 1, 1);     // G6: x ones (bf16), N = 16
  const uint64_t dW1H = umma_desc_k_sw128(sbase + M::W1H), dW1L = umma_desc_k_sw128(sbase + M::W1L);
  const uint64_t dW2 = umma_desc_k_sw128(sbase + M::W2), dW2X = umma_desc_k_sw128(sbase + M::W2X);
  const uint64_t dW2T = umma_desc_k_sw128(sbase + M::W2T), dW1T = umma_desc_k_sw128(sbase + M::W1T);
  const uint64_t dAX = umma_desc_k_sw128(sbase + M::AX, 0);
  const uint64_t dF = umma_desc_k_sw128(tiles_u32 + M::F), dA1 = umma_desc_k_sw128(tiles_u32 + M::A1H);
  const uint64_t dDH2 = umma_desc_k_sw128(tiles_u32 + M::DH2), dDH1 = umma_desc_k_sw128(tiles_u32 + M::DH1);
  const uint64_t dMw = umma_desc_mn_sw128(tiles_u32 + M::DH2, FB_TILE_BYTES, 1024);  // [dH2 | dH1]^T
  const uint64_t dNw = umma_desc_mn_sw128(tiles_u32 + M::A1, FB_TILE_BYTES, 1024);   // [A1 (bf16) | F]
  const uint64_t dOnes = umma_desc_mn_sw128(sbase + M::ONES, 0, 0);
  const bool issuer = (wt == 0);
  uint8_t* const rowA1 = tiles + M::A1H + wt * 128;   // fp16 (G2 operand, layer-1 mask)
  uint8_t* const rowA1b = tiles + M::A1 + wt * 128;   // bf16 copy (G5 operand)
  uint8_t* const rowDH2 = tiles + M::DH2 + wt * 128;
  uint8_t* const rowDH1 = tiles + M::DH1 + wt * 128;
  const int sw = wt & 7;

  // this CTA's contiguous tile range; the warpgroups interleave inside it
  const int t0 = (int)((long long)num_tiles * blockIdx.x / gridDim.x);
  const int t1 = (int)((long long)num_tiles * (blockIdx.x + 1) / gridDim.x);
  float acc3[FB_C];
#pragma unroll
  for (int j = 0; j < FB_C; ++j) acc3[j] = 0.f;
  float accg = 0.f;
  uint32_t phF = 0, phM = 0;
  int cur_b = -1;
  bool ones_fresh = true;  // the next G6 overwrites the column-sum accumulator

  auto flush_ones = [&](int b) {
    // rows 0-63: column sums of dH2 (db2), rows 64-127: column sums of dH1 of image b (dbz)
    uint32_t v[16];
    tmem_ld16(lane_base + colON, v);
    tmem_ld_wait();
    const float s = __uint_as_float(v[0]);
    if (wt < 64) atomicAdd(db2 + wt, s);
    else atomicAdd(dbz + b * FB_C + (wt - 64), s);
    tc_fence_before();
  };
  auto wait_mma = [&]() {
    mbar_wait(mbarM, phM);
    phM ^= 1;
    tc_fence_after();
  };
  auto publish = [&]() {  // operand rows written -> visible to the tensor core, all rows of the group done
    fence_proxy_async_smem();
    tc_fence_before();
    fb_named_bar(1 + wg);
  };

  for (int tile = t0 + wg; tile < t1; tile += 2) {
    const int b = tile / tiles_per_img;
    const int p0 = (tile - b * tiles_per_img) * 128;
    if (b != cur_b) {
      if (cur_b >= 0) {
        flush_ones(cur_b);
        ones_fresh = true;
      }
      fb_named_bar(1 + wg);  // nobody still reads the previous image's bz
      if (wt < FB_C) bzs[wt] = bzg[b * FB_C + wt];
      cur_b = b;
    }
    if (issuer) {
      mbar_expect_tx(mbarF, FB_TILE_BYTES);
      tma_load_3d(tiles_u32 + M::F, &tmF, mbarF, 0, p0, b);
    }
    const int pix = p0 + wt;
    const bool valid = pix < P;
    const size_t gp = static_cast<size_t>(b) * P + pix;
    const float g = valid ? dlogit[gp] : 0.f;
    accg += g;
    fb_named_bar(1 + wg);  // bz visible
    // ---- G1
    if (issuer) {
      mbar_wait(mbarF, phF);
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16(tmem + colHA, dF + 2 * k, dW1H + 2 * k, id_bb, k ? 1u : 0u);
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16(tmem + colHA, dF + 2 * k, dW1L + 2 * k, id_bb, 1u);
      umma_commit(mbarM);
    }
    phF ^= 1;
    wait_mma();
    // ---- a1 = relu(H1 + bz) -> A1 row
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      uint32_t v[32];
      tmem_ld32(lane_base + colHA + half * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float x[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) x[e] = __uint_as_float(v[8 * c + e]) + bzs[half * 32 + 8 * c + e];
        // same rounding as the forward kernel: relu + saturating fp16
        *reinterpret_cast<uint4*>(rowA1 + (((half * 4 + c) ^ sw) << 4)) =
            make_uint4(fb_relu_pack2h(x[0], x[1]), fb_relu_pack2h(x[2], x[3]), fb_relu_pack2h(x[4], x[5]),
                       fb_relu_pack2h(x[6], x[7]));
#pragma unroll
        for (int e = 0; e < 8; ++e) x[e] = fmaxf(x[e], 0.f);
        *reinterpret_cast<uint4*>(rowA1b + (((half * 4 + c) ^ sw) << 4)) =
            make_uint4(fb_pack2(x[0], x[1]), fb_pack2(x[2], x[3]), fb_pack2(x[4], x[5]), fb_pack2(x[6], x[7]));
      }
    }
    publish();
    // ---- G2 (+ b2 through the K extension)
    if (issuer) {
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16(tmem + colHB, dA1 + 2 * k, dW2 + 2 * k, id_hh, k ? 1u : 0u);
      umma_bf16(tmem + colHB, dAX, dW2X, id_hh, 1u);
      umma_commit(mbarM);
    }
    wait_mma();
    // ---- dh2 = (H2 > 0) g w3 ; dW3 += g relu(H2)
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      uint32_t v[32];
      tmem_ld32(lane_base + colHB + half * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float x[8];
        const float4 wa = w3s[half * 8 + 2 * c], wb = w3s[half * 8 + 2 * c + 1];  // warp-wide broadcast reads
        const float w8[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int j = half * 32 + 8 * c + e;
          const float h2 = __uint_as_float(v[8 * c + e]);
          acc3[j] = fmaf(g, fmaxf(h2, 0.f), acc3[j]);
          x[e] = h2 > 0.f ? g * w8[e] : 0.f;
        }
        *reinterpret_cast<uint4*>(rowDH2 + (((half * 4 + c) ^ sw) << 4)) =
            make_uint4(fb_pack2(x[0], x[1]), fb_pack2(x[2], x[3]), fb_pack2(x[4], x[5]), fb_pack2(x[6], x[7]));
      }
    }
    publish();
    // ---- G3: dA1 = dH2 . W2   (B tile = W2^T, K-major)
    if (issuer) {
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16(tmem + colHA, dDH2 + 2 * k, dW2T + 2 * k, id_bh, k ? 1u : 0u);
      umma_commit(mbarM);
    }
    wait_mma();
    // ---- dh1 = (a1 > 0) dA1 -> DH1 row (a1 read back from this thread's A1 row)
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      uint32_t v[32];
      tmem_ld32(lane_base + colHA + half * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int ch = ((half * 4 + c) ^ sw) << 4;
        const uint4 a = *reinterpret_cast<const uint4*>(rowA1 + ch);
        const uint32_t aw[4] = {a.x, a.y, a.z, a.w};
        float x[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          // fp16 a1 >= 0 always: "> 0" <=> the 15 magnitude bits are non-zero
          const uint32_t bits = (aw[e >> 1] >> ((e & 1) * 16)) & 0x7fffu;
          x[e] = bits ? __uint_as_float(v[8 * c + e]) : 0.f;
        }
        *reinterpret_cast<uint4*>(rowDH1 + ch) =
            make_uint4(fb_pack2(x[0], x[1]), fb_pack2(x[2], x[3]), fb_pack2(x[4], x[5]), fb_pack2(x[6], x[7]));
      }
    }
    publish();
    // ---- G4: dF = dH1 . W1f ; G5: weight gradients ; G6: column sums
    if (issuer) {
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16(tmem + colHB, dDH1 + 2 * k, dW1T + 2 * k, id_bh, k ? 1u : 0u);
#pragma unroll
      for (int k = 0; k < 8; ++k) umma_bf16(tmem + colDW, dMw + 128 * k, dNw + 128 * k, id_w, 1u);
#pragma unroll
      for (int k = 0; k < 8; ++k) umma_bf16(tmem + colON, dMw + 128 * k, dOnes, id_o, (ones_fresh && k == 0) ? 0u : 1u);
      umma_commit(mbarM);
    }
    ones_fresh = false;
    wait_mma();
    // ---- dF -> bf16 -> global (one contiguous 128-byte row per thread)
    {
      uint4* dst = reinterpret_cast<uint4*>(dfeat + gp * FB_C);
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t v[32];
        tmem_ld32(lane_base + colHB + half * 32, v);
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int c = 0; c < 4; ++c)
            dst[half * 4 + c] = make_uint4(fb_pack2(__uint_as_float(v[8 * c + 0]), __uint_as_float(v[8 * c + 1])),
                                           fb_pack2(__uint_as_float(v[8 * c + 2]), __uint_as_float(v[8 * c + 3])),
                                           fb_pack2(__uint_as_float(v[8 * c + 4]), __uint_as_float(v[8 * c + 5])),
                                           fb_pack2(__uint_as_float(v[8 * c + 6]), __uint_as_float(v[8 * c + 7])));
        }
      }
    }
    tc_fence_before();
    fb_named_bar(1 + wg);  // every TMEM / smem read of this tile is done before the next tile overwrites
    tc_fence_after();
  }
  if (cur_b >= 0) flush_ones(cur_b);

  // ---- dW3 / db3: block reduction, then one atomic per element
#pragma unroll
  for (int j = 0; j < FB_C; ++j) atomicAdd(&red[j], acc3[j]);
  atomicAdd(&red[FB_C], accg);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid < FB_C) atomicAdd(dw3 + tid, red[tid]);
  if (tid == FB_C) atomicAdd(db3, red[FB_C]);
  // ---- shared weight-gradient accumulator: rows 0-63 x cols 0-63 = dW2[j][i]; rows 64-127 x cols 64-127 = dW1f[j][i]
  if (wg == 0) {
    const bool top = wt < 64;
    float* dst = top ? dw2 + wt * FB_C : dw1f + (wt - 64) * FB_C;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      uint32_t v[32];
      tmem_ld32(lane_base + colDW + (top ? 0 : 64) + half * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 32; ++e) atomicAdd(dst + half * 32 + e, __uint_as_float(v[e]));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// bz[b][j] = b1[j] + sum_d W1[j][64 + d] z[b][d]
__global__ void fb_bz_kernel(const float* __restrict__ z, const float* __restrict__ w1, const float* __restrict__ b1,
                             float* __restrict__ bz, int B, int L) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * FB_C) return;
  const int b = i / FB_C, j = i - b * FB_C;
  float acc = b1[j];
  for (int d = 0; d < L; ++d) acc = fmaf(w1[j * (FB_C + L) + FB_C + d], z[b * L + d], acc);
  bz[i] = acc;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// Launches the tensor-core backward.  bz: fp32 [B][64] scratch.  The accumulation targets must be zero-initialised.
int fcomb_bwd_tc(const void* feat, const float* z, const float* w1, const float* b1, const float* w2,
                 const float* b2, const float* w3, const float* dlogit, int B, int P, int L, void* dfeat, float* dw1f,
                 float* dw2, float* db2, float* dw3, float* db3, float* dbz, float* bz, const int* skip_flag,
                 cudaStream_t st) {
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(get_encode_tiled());
  if (!enc) return PDA_ERR_DRIVER;
  CUtensorMap tm;
  cuuint64_t dims[3] = {64, (cuuint64_t)P, (cuuint64_t)B};
  cuuint64_t strides[2] = {128, (cuuint64_t)P * 128};
  cuuint32_t box[3] = {64, 128, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(feat), dims, strides, box, estr,
          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return PDA_ERR_TENSORMAP;
  static int configured[64];
  if (dyn_smem_attr_needed(configured, FbSmem::DYN_BYTES)) {
    if (cudaFuncSetAttribute(fcomb_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FbSmem::DYN_BYTES) !=
        cudaSuccess)
      return PDA_ERR_CUDA;
  }
  const int tiles_per_img = (P + 127) / 128;
  const long long num_tiles = (long long)tiles_per_img * B;
  if (num_tiles > 0x7fffffffLL) return PDA_ERR_SHAPE;
  const int sms = sm_budget(0);
  const int grid = (int)(num_tiles < sms ? num_tiles : sms);
  PDA_COUNT(2);
  fb_bz_kernel<<<(B * FB_C + 255) / 256, 256, 0, st>>>(z, w1, b1, bz, B, L);
  fcomb_bwd_tc_kernel<<<grid, 256, FbSmem::DYN_BYTES, st>>>(tm, bz, w1, w2, b2, w3, skip_flag, dlogit, P, L, B, tiles_per_img,
                                                            (int)num_tiles, static_cast<__nv_bfloat16*>(dfeat), dw1f,
                                                            dw2, db2, dw3, db3, dbz);
  return cudaGetLastError() == cudaSuccess ? PDA_OK : PDA_ERR_CUDA;
}

}  // namespace pda
