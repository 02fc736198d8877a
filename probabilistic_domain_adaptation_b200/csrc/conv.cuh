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
};

void* get_encode_tiled();  // cuTensorMapEncodeTiled driver entry point (or nullptr)
int make_act_tensor_map(CUtensorMap* tm, const void* ptr, int B, int H, int W, int C, int box_w, int box_h, int box_c);
int make_mat_tensor_map(CUtensorMap* tm, const void* ptr, long long inner, long long outer, int box_inner,
                        int box_outer);

int conv3x3_tc(const void* src0, int c0, const void* src1, int c1, const void* wpacked, const float* bias, void* out,
               void* out_pool, int B, int H, int W, int cout, int relu, int bn_override, cudaStream_t stream);

}  // namespace pda
