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
#include <stdlib.h>

#ifndef PDA_CONV_PAIR_DEFAULT
#define PDA_CONV_PAIR_DEFAULT 1
#endif

#include "conv.cuh"
#include "ptx.cuh"

namespace pda {

// ---------------------------------------------------------------------------------------------
// Persistent kernel.  Work unit = (pixel tile of 16*MT rows x 8 columns of one image, block of BN output channels).
//
// Tap reuse: for one 64-channel chunk the producer loads three column-shifted input SLABS (kx = 0,1,2), each
// [16*MT + 2 rows][8 px][64 ch] with SWIZZLE_128B, so that an 8-pixel output row segment is one 1024-byte swizzle
// atom.  The A operand of tap (ky, kx) for M-tile mt is then simply slab kx advanced by (16*mt + ky) atoms: the
// same shared-memory bytes feed three taps and both M-tiles, i.e. 3.2-3.4 pixel-loads per output pixel instead of 9.
// Weights: one [BN][64] tile per (tap, chunk) through a second ring, or fully resident for a 64->64 layer.
// Accumulators: MT x BN fp32 columns, double buffered in TMEM, so the epilogue of unit i overlaps the MMAs of
// unit i+1.  Warp 0 = TMA producer, warp 1 = MMA issuer, warps 2-9 = epilogue: two warpgroups share the unit's
// [128 px x 64 ch] sub-tiles; a warp turns its 32 TMEM lanes (= 4 rows x 8 px) into bias + ReLU + bf16, stages them in
// a swizzled 4 KB shared-memory box and writes it with ONE TMA store (full 128-byte lines, clipped at the image edge).
// ---------------------------------------------------------------------------------------------
template <int BN, int MT, bool RES>
struct ConvCfg {
  static constexpr int SLAB_ROWS = 16 * MT + 2;
  static constexpr int A_BYTES = SLAB_ROWS * 1024;     // one slab: SLAB_ROWS x (8 px x 128 B)
  static constexpr int B_BYTES = BN * 128;             // one (tap, chunk) weight tile
  static constexpr int A_STAGES = 3;
  static constexpr int B_STAGES = RES ? 9 : (BN == 256 ? 4 : (MT == 2 ? (BN == 128 ? 5 : 8) : 6));
  static constexpr int A_OFF = 0;
  static constexpr int B_OFF = A_STAGES * A_BYTES;
  static constexpr int STG_OFF = B_OFF + B_STAGES * B_BYTES;   // 8 epilogue warps x 4 KB output staging
  static constexpr int BAR_OFF = STG_OFF + 8 * 4096;
  static constexpr int NBARS = 2 * A_STAGES + 2 * B_STAGES + 4;
  static constexpr int SLOT_OFF = BAR_OFF + NBARS * 8;
  static constexpr int BIAS_OFF = SLOT_OFF + 16;
  static constexpr int MAX_COUT = 512;
  static constexpr int TOTAL = BIAS_OFF + MAX_COUT * 4;
  static constexpr int DYN_BYTES = TOTAL + 1024;
  static constexpr int TMEM_COLS = 2 * MT * BN;
  static_assert(TMEM_COLS <= 512 && (TMEM_COLS & (TMEM_COLS - 1)) == 0, "TMEM columns");
  static_assert(DYN_BYTES <= 227 * 1024, "shared memory");
};

constexpr int CONV_THREADS = 320;

template <int BN, int MT, bool RES, bool F16>
__global__ void __launch_bounds__(CONV_THREADS, 1)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                  const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmOut,
                  const ConvArgs p) {
  using L = ConvCfg<BN, MT, RES>;
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
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + L::SLOT_OFF);
  float* bias_s = reinterpret_cast<float*>(smem + L::BIAS_OFF);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int ctot = p.c0 + p.c1;
  const int chunks = ctot >> 6;
  const int n_blocks = p.cout / BN;
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const int units = tiles_per_img * p.B * n_blocks;

  if (threadIdx.x == 0) {
    for (int s = 0; s < L::A_STAGES; ++s) {
      mbar_init(a_full(s), 1);
      mbar_init(a_empty(s), 1);
    }
    for (int s = 0; s < L::B_STAGES; ++s) {
      mbar_init(b_full(s), 1);
      mbar_init(b_empty(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(acc_full(s), 1);
      mbar_init(acc_empty(s), 8);  // one arrive per epilogue warp
    }
    fence_mbar_init();
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmOut);
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_slot)), L::TMEM_COLS);
    tmem_relinquish();
  }
  if (warp >= 2) {
    for (int i = threadIdx.x - 64; i < p.cout; i += CONV_THREADS - 64) bias_s[i] = p.bias ? p.bias[i] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    // The whole warp runs the (warp-uniform) control flow so that addresses and coordinates live in uniform
    // registers; one elected lane issues the TMA instructions.
    const bool leader = elect_one();
    int as = 0, bs = 0;
    uint32_t aph = 0, bph = 0;
    if (RES) {
      // whole weight matrix of this (single) n-block: 9 taps x 1 chunk, loaded once
      if (leader) {
        mbar_expect_tx(b_full(0), 9 * L::B_BYTES);
        for (int tap = 0; tap < 9; ++tap)
          tma_load_2d(sbase + L::B_OFF + tap * L::B_BYTES, &tmB, b_full(0), tap * ctot, 0);
      }
      __syncwarp();
    }
    for (int u = blockIdx.x; u < units; u += gridDim.x) {
      const int nb = u % n_blocks;
      const int mtile = u / n_blocks;
      const int img = mtile / tiles_per_img;
      const int t = mtile - img * tiles_per_img;
      const int ty = t / p.tiles_x, tx = t - ty * p.tiles_x;
      const int x0 = tx * 8, y0 = ty * (16 * MT);
      const int n0 = nb * BN;
      for (int ch = 0; ch < chunks; ++ch) {
        const int c = ch << 6;
        for (int kx = 0; kx < 3; ++kx) {
          mbar_wait(a_empty(as), aph ^ 1);
          if (leader) {
            mbar_expect_tx(a_full(as), L::A_BYTES);
            const uint32_t dst = sbase + L::A_OFF + as * L::A_BYTES;
            if (c < p.c0)
              tma_load_4d(dst, &tmA0, a_full(as), c, x0 + kx - 1, y0 - 1, img);
            else
              tma_load_4d(dst, &tmA1, a_full(as), c - p.c0, x0 + kx - 1, y0 - 1, img);
          }
          __syncwarp();
          if (++as == L::A_STAGES) { as = 0; aph ^= 1; }
          if (!RES) {
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
              mbar_wait(b_empty(bs), bph ^ 1);
              if (leader) {
                mbar_expect_tx(b_full(bs), L::B_BYTES);
                tma_load_2d(sbase + L::B_OFF + bs * L::B_BYTES, &tmB, b_full(bs), (ky * 3 + kx) * ctot + c, n0);
              }
              __syncwarp();
              if (++bs == L::B_STAGES) { bs = 0; bph ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (warp-uniform control, elected lane issues)
    const bool leader = elect_one();
    // operands: bf16 x bf16, or fp16 x fp16 on the fp16-activation (no-grad) path; fp32 accumulate either way
    constexpr uint32_t idesc = F16 ? umma_idesc_f16(128, BN) : umma_idesc_bf16(128, BN);
    int as = 0, bs = 0;
    uint32_t aph = 0, bph = 0;
    uint32_t it = 0;
    if (RES) {
      mbar_wait(b_full(0), 0);
      tc_fence_after();
    }
    for (int u = blockIdx.x; u < units; u += gridDim.x, ++it) {
      const uint32_t buf = it & 1;
      mbar_wait(acc_empty(buf), ((it >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t dcol = tmem_base + buf * (MT * BN);
      for (int ch = 0; ch < chunks; ++ch) {
        for (int kx = 0; kx < 3; ++kx) {
          mbar_wait(a_full(as), aph);
          tc_fence_after();
          const uint32_t sa = sbase + L::A_OFF + as * L::A_BYTES;
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
            if (leader) {
              const uint64_t db = umma_desc_k_sw128(sb);
              const uint32_t acc = (ch | kx | ky) != 0 ? 1u : 0u;
#pragma unroll
              for (int mt = 0; mt < MT; ++mt) {
                const uint64_t da = umma_desc_k_sw128(sa + (16 * mt + ky) * 1024);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma_bf16(dcol + mt * BN, da + 2 * k, db + 2 * k, idesc, (acc | k) != 0 ? 1u : 0u);
              }
              if (!RES) umma_commit(b_empty(bs));
            }
            __syncwarp();
            if (!RES) {
              if (++bs == L::B_STAGES) { bs = 0; bph ^= 1; }
            }
          }
          if (leader) umma_commit(a_empty(as));
          __syncwarp();
          if (++as == L::A_STAGES) { as = 0; aph ^= 1; }
        }
      }
      if (leader) umma_commit(acc_full(buf));
      __syncwarp();
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
    uint32_t it = 0;
    for (int u = blockIdx.x; u < units; u += gridDim.x, ++it) {
      const int nb = u % n_blocks;
      const int mtile = u / n_blocks;
      const int img = mtile / tiles_per_img;
      const int t = mtile - img * tiles_per_img;
      const int ty = t / p.tiles_x, tx = t - ty * p.tiles_x;
      const int n0 = nb * BN;
      const int x = tx * 8 + ltx;
      const uint32_t buf = it & 1;
      mbar_wait(acc_full(buf), (it >> 1) & 1);
      tc_fence_after();
      if (wg >= NSUB) {  // nothing to read for this warpgroup (MT = 1, BN = 64)
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty(buf));
      }
#pragma unroll 1
      for (int sub = wg; sub < NSUB; sub += 2) {
        const int mt = sub / NG, g = sub - mt * NG;
        const int y = ty * (16 * MT) + mt * 16 + lty;
        const bool valid = (y < p.H) && (x < p.W);
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
            // last TMEM read of this warp for this unit: hand the accumulator buffer back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty(buf));
          }
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float tv = __uint_as_float(v[j]) + bias_s[n0 + g * 64 + cb * 32 + j];
            f[j] = p.relu ? fmaxf(tv, 0.f) : tv;
          }
          if (F16) {
            // fp16 range guard: the stores below saturate at +-65504; a value beyond that raises the caller's flag
            float vmax = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) vmax = fmaxf(vmax, fabsf(f[j]));
            if (p.range_flag != nullptr && __any_sync(0xffffffffu, !(vmax <= F16_MAX)) && lane == 0)
              atomicOr(p.range_flag, 1);
          }
          if (p.mask) {
            // bf16 mask values are post-ReLU activations (>= 0): "> 0" <=> magnitude bits non-zero and sign clear
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
            // this pixel's 64 bytes of the 128-byte row, 16-byte chunks XOR-swizzled by (row & 7) as TMA expects
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
            // 2x2 average of the fp32 post-ReLU values: partners are lane^1 (x) and lane^8 (y)
            // (H, W even => a valid even pixel always has valid partners).
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
          if (lane == 0) {
            // this warp's 32 pixels = 4 rows x 8 px of the tile
            tma_store_4d(&tmOut, stage_u32, n0 + g * 64, tx * 8, ty * (16 * MT) + mt * 16 + q * 4, img);
            tma_store_commit();
          }
        }
      }
    }
    if (lane == 0) tma_store_wait<0>();  // shared memory must outlive the last store
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, L::TMEM_COLS);
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
                        int box_c, int swizzle128) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return PDA_ERR_DRIVER;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
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

// 1: conv3x3 launches use the CTA-pair (cta_group::2) kernel; initial value from PDA_CONV_PAIR, changed at run time
// through pda_set_conv_pair (A/B measurements, tests).  set < 0: query only.
int conv_pair_mode(int set) {
  static std::atomic<int> mode{[] {
    const char* e = getenv("PDA_CONV_PAIR");
    return e ? atoi(e) : PDA_CONV_PAIR_DEFAULT;
  }()};
  if (set >= 0) return mode.exchange(set);
  return mode.load();
}

// SMs the persistent tensor-core kernels (conv, conv pair, wgrad, Fcomb backward) may occupy.  Data-parallel training
// lowers it by a few SMs so that NCCL's all-reduce CTAs find room to run NEXT TO the backward kernels instead of
// between them (parallel.GradAllReducer); 148 = the whole B200.  set <= 0: query only.
int sm_budget(int set) {
  static std::atomic<int> budget{148};
  if (set > 0) return budget.exchange(set > 148 ? 148 : (set < 2 ? 2 : set));
  return budget.load();
}

template <int BN, int MT, bool RES, bool F16>
static int launch_conv_fmt(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b, const CUtensorMap& o,
                           const ConvArgs& args, cudaStream_t stream) {
  using L = ConvCfg<BN, MT, RES>;
  static int configured[64];
  if (dyn_smem_attr_needed(configured, L::DYN_BYTES)) {
    cudaError_t e = cudaFuncSetAttribute(conv3x3_tc_kernel<BN, MT, RES, F16>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, L::DYN_BYTES);
    if (e != cudaSuccess) return PDA_ERR_CUDA;
  }
  const long long units = (long long)args.tiles_x * args.tiles_y * args.B * (args.cout / BN);
  if (units > 0x7fffffffLL) return PDA_ERR_SHAPE;
  const int sms = sm_budget(0);
  const int grid = (int)(units < sms ? units : sms);
  PDA_COUNT(1);
  conv3x3_tc_kernel<BN, MT, RES, F16><<<grid, CONV_THREADS, L::DYN_BYTES, stream>>>(a0, a1, b, o, args);
  return cudaGetLastError() == cudaSuccess ? PDA_OK : PDA_ERR_CUDA;
}

template <int BN, int MT, bool RES>
static int launch_conv(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b, const CUtensorMap& o,
                       const ConvArgs& args, cudaStream_t stream) {
  return args.act_f16 ? launch_conv_fmt<BN, MT, RES, true>(a0, a1, b, o, args, stream)
                      : launch_conv_fmt<BN, MT, RES, false>(a0, a1, b, o, args, stream);
}

int conv3x3_tc(const void* src0, int c0, const void* src1, int c1, const void* wpacked, const float* bias, void* out,
               void* out_pool, const void* mask, int B, int H, int W, int cout, int relu, int bn_override,
               int act_f16, int* range_flag, cudaStream_t stream) {
  if (c0 <= 0 || (c0 & 63) || (c1 & 63) || (cout & 63) || cout > 512 || B <= 0 || H <= 0 || W <= 0)
    return PDA_ERR_SHAPE;
  if (out_pool && ((H & 1) || (W & 1))) return PDA_ERR_SHAPE;
  {
    // CTA-pair kernel (csrc/conv3x3_tc2.cu) for every launch with at least two pixel tiles
    const int pair_mode = conv_pair_mode(-1);
    const int th = ((H > 16 && !(bn_override == 256 || (bn_override == 0 && cout % 256 == 0))) ? 32 : 16);
    const long long mtiles = (long long)((W + 7) / 8) * ((H + th - 1) / th) * B;
    if (pair_mode && mtiles >= 2)
      return conv3x3_tc2(src0, c0, src1, c1, wpacked, bias, out, out_pool, mask, B, H, W, cout, relu, bn_override,
                         act_f16, range_flag, stream);
  }
  // N = 256 halves the shared-memory bytes per MAC of the B operand (N = 128 tiles sit exactly at the 128 B/clk
  // shared-memory port limit); its 2 x 256 accumulator columns leave room for one M-tile per unit only
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
  a.wide = a.wide_base_offset = 0;
  a.up_src = nullptr;
  a.range_flag = act_f16 ? range_flag : nullptr;
  CUtensorMap tA0, tA1, tB;
  int r = make_act_tensor_map(&tA0, src0, B, H, W, c0, 8, a.tile_h + 2, 64);
  if (r) return r;
  if (c1 > 0) {
    r = make_act_tensor_map(&tA1, src1, B, H, W, c1, 8, a.tile_h + 2, 64);
    if (r) return r;
  } else {
    tA1 = tA0;
  }
  r = make_mat_tensor_map(&tB, wpacked, 9LL * (c0 + c1), cout, 64, bn);
  if (r) return r;
  CUtensorMap tO = tA0;  // unused when out == nullptr
  if (out) {
    r = make_act_tensor_map(&tO, out, B, H, W, cout, 8, 4, 64);  // one epilogue warp = 4 rows x 8 px x 64 ch
    if (r) return r;
  }
  const bool resident = (bn == 64 && cout == 64 && c0 + c1 == 64 && mt == 2);
  if (mt == 2) {
    if (bn == 128) return launch_conv<128, 2, false>(tA0, tA1, tB, tO, a, stream);
    return resident ? launch_conv<64, 2, true>(tA0, tA1, tB, tO, a, stream)
                    : launch_conv<64, 2, false>(tA0, tA1, tB, tO, a, stream);
  }
  if (bn == 256) return launch_conv<256, 1, false>(tA0, tA1, tB, tO, a, stream);
  if (bn == 128) return launch_conv<128, 1, false>(tA0, tA1, tB, tO, a, stream);
  return launch_conv<64, 1, false>(tA0, tA1, tB, tO, a, stream);
}

}  // namespace pda
