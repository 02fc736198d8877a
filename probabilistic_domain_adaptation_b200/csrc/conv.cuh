// Internal declarations shared by the CUDA translation units (not part of the C ABI).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pda_b200.h"

#include <atomic>
extern std::atomic<long long> g_pda_launches;  // kernels launched through the C ABI (bench.py gpu_launches)
#define PDA_COUNT(n) g_pda_launches.fetch_add((n), std::memory_order_relaxed)

namespace pda {

struct ConvArgs {
  int B, H, W;        // spatial size of input == output (pad 1, stride 1)
  int c0, c1;         // channels of K-segment 0 / 1 (c1 == 0: single source)
  int cout;
  int tile_w, tile_h; // tile_w * tile_h == 128 output pixels per CTA
  int tiles_x, tiles_y;
  int relu;
  const float* bias;
  __nv_bfloat16* out;       // NHWC [B][H][W][cout] or nullptr
  __nv_bfloat16* out_pool;  // NHWC [B][H/2][W/2][cout] (2x2 average of the post-ReLU fp32 values) or nullptr
  const __nv_bfloat16* mask;  // NHWC [B][H][W][cout] or nullptr: outputs are zeroed where mask <= 0 (ReLU backward
                              // of the layer that produced this conv's input, fused into the dgrad epilogue)
  int act_f16;                // activations / weights / outputs are fp16 (no-grad path) instead of bf16
  int wide, wide_base_offset; // CTA-pair kernel: one 16-px slab per chunk (csrc/conv3x3_tc2.cu)
  const void* up_src;         // CTA-pair kernel: K segment 0 = bilinear x2 of this [B][H/2][W/2][c0] tensor, or nullptr
  int* range_flag;            // fp16 only, may be nullptr: set to 1 when an output exceeds the fp16 range (the store
                              // saturates at +-65504)
};

int conv3x3_first_tc(const float* x0, const float* x1, const float* w, const float* bias, void* out, int B, int H, int W,
                     int relu, int act_f16, cudaStream_t st);  // csrc/conv_first_tc.cu (cout = 64)
void* get_encode_tiled();  // cuTensorMapEncodeTiled driver entry point (or nullptr)
int make_act_tensor_map(CUtensorMap* tm, const void* ptr, int B, int H, int W, int C, int box_w, int box_h, int box_c,
                        int swizzle128 = 1);
int make_mat_tensor_map(CUtensorMap* tm, const void* ptr, long long inner, long long outer, int box_inner,
                        int box_outer);

int fcomb_bwd_tc(const void* feat, const float* z, const float* w1, const float* b1, const float* w2, const float* b2,
                 const float* w3, const float* dlogit, int B, int P, int L, void* dfeat, float* dw1f, float* dw2,
                 float* db2, float* dw3, float* db3, float* dbz, float* bz, const int* skip_flag, cudaStream_t st);

// exact fp32 Fcomb + consensus (csrc/fcomb.cu); run_flag != nullptr: only runs when *run_flag != 0 (device-side)
int fcomb_mc_fp32(const void* feat, const float* z, const float* w1, const float* b1, const float* w2, const float* b2,
                  const float* w3, const float* b3, int B, int P, int S, int latent, float upper, float lower,
                  float* mean_prob, float* cons_weight, int64_t* cons_mask, float* logits, float* probs,
                  const int* run_flag, int feat_f16, cudaStream_t stream);

int conv3x3_tc(const void* src0, int c0, const void* src1, int c1, const void* wpacked, const float* bias, void* out,
               void* out_pool, const void* mask, int B, int H, int W, int cout, int relu, int bn_override,
               int act_f16, int* range_flag, cudaStream_t stream);

int conv_pair_mode(int set);  // csrc/conv3x3_tc.cu
int sm_budget(int set);       // csrc/conv3x3_tc.cu

// CTA-pair (cta_group::2) variant, csrc/conv3x3_tc2.cu
int conv3x3_tc2(const void* src0, int c0, const void* src1, int c1, const void* wpacked, const float* bias, void* out,
                void* out_pool, const void* mask, int B, int H, int W, int cout, int relu, int bn_override,
                int act_f16, int* range_flag, cudaStream_t stream, const void* up_src = nullptr);

// log2(C / 8) for C = 8 * 2^k (the NHWC kernels address 16-byte channel chunks with shifts), else -1
static inline int c8_shift(int C) {
  if (C < 8 || (C & 7)) return -1;
  const int c8 = C >> 3;
  if (c8 & (c8 - 1)) return -1;
  int s = 0;
  while ((1 << s) < c8) ++s;
  return s;
}


// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device setting: remember, per device, the largest value already
// requested for one kernel (state: zero-initialised int[64]).  Returns true when the attribute must be (re)set.
static inline bool dyn_smem_attr_needed(int* state, int bytes) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
  if (state[dev] >= bytes) return false;
  state[dev] = bytes;
  return true;
}

// Launch shape of the row-decomposed NHWC elementwise kernels: x = 256-thread blocks across one row's (x, chunk)
// elements, y = blocks striding over the rows, sized to ~`target` blocks in total.
static inline dim3 row_grid(long long row_elems, long long rows, int target = 148 * 8) {
  const long long gx = (row_elems + 255) / 256;
  long long gy = (target + gx - 1) / gx;
  if (gy > rows) gy = rows;
  if (gy > 65535) gy = 65535;
  if (gy < 1) gy = 1;
  return dim3((unsigned)gx, (unsigned)gy, 1);
}

// Element index -> (channel chunk, x, y, image) with 32-bit arithmetic (C8 is a power of two: C in {64..512}).
struct Px {
  unsigned c, x, y, b;
};
__device__ __forceinline__ Px split_index(unsigned t, int c_shift, unsigned W, unsigned H) {
  Px p;
  p.c = t & ((1u << c_shift) - 1u);
  const unsigned pix = t >> c_shift;
  const unsigned row = pix / W;
  p.x = pix - row * W;
  p.b = row / H;
  p.y = row - p.b * H;
  return p;
}


}  // namespace pda
