// On-device weak / strong view augmentation.
//
// Reference: the target-domain loaders build two augmented views of every raw patch on the CPU, per sample, in the
// DataLoader workers (/root/reference/prob_utils/my_datasets/my_image_collection_dataset.py:349-357):
//     view = get_raw_transform(normalizer=my_standardize_torch, augmentation1=Compose([
//                my_standardize_torch,
//                RandomApply([GaussianBlur(kernel_size, sigma)], p),
//                RandomApply([AdditiveGaussianNoise(scale, clip_kwargs=False)], p),
//                RandomApply([RandomContrast(alpha, mean=0.0, clip_kwargs=False)], p)]))
// (/root/reference/MitoEM/common.py:50-68, LIVECell/livecell_fm.py:43-67, livecell_adamatch.py:16-38).
//   my_standardize_torch  prob_utils/my_utils/util.py:9-14:  x -= x.mean(); x /= (x.std() + 1e-7)   (unbiased std)
//   GaussianBlur          torchvision.transforms.GaussianBlur(k, sigma): reflect padding, kernel1d = pdf / sum(pdf),
//                         pdf = exp(-0.5 (x / sigma)^2), x = -(k-1)/2 .. (k-1)/2
//   AdditiveGaussianNoise x + N(0, scale)        RandomContrast  mean + alpha (x - mean)      (torch_em.transform.raw)
//
// Here: one statistics pass per raw batch (shared by all views) + ONE fused kernel per view.  The random decisions
// (apply flags, kernel size, sigma, scale, alpha) are drawn on the host in the reference's order and arrive as a small
// per-image parameter table; the unit-normal noise field is drawn by torch on the device.  HBM-bound: 4 B/px for the
// statistics pass, 8 B/px (+ 4 B/px of noise) per view.
#include <cuda_runtime.h>
#include <stdint.h>

#include "conv.cuh"

namespace pda {

constexpr int AUG_RMAX = 15;            // blur radius limit (kernel size <= 31; torch_em draws sizes <= 23)
constexpr int AUG_TILE = 32;            // output tile edge
constexpr int AUG_IN = AUG_TILE + 2 * AUG_RMAX;
constexpr int AUG_NPARAM = 8;           // per image: ksize, sigma, noise_scale, alpha, contrast_mean, n_standardize, -, -

// per-image sum and sum of squares (fp64), grid (gx, B); stats must be zeroed by the caller
__global__ void __launch_bounds__(256)
image_stats_kernel(const float* __restrict__ img, long long n, double* __restrict__ stats) {
  const int b = blockIdx.y;
  const float* p = img + (size_t)b * n;
  double s = 0.0, q = 0.0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if ((n & 3) == 0 && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
    const long long n4 = n >> 2;
    for (; i < n4; i += stride) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(p) + i);
      s += (double)v.x + (double)v.y + (double)v.z + (double)v.w;
      q += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
    }
  } else {
    for (; i < n; i += stride) {
      const double v = (double)p[i];
      s += v;
      q += v * v;
    }
  }
  __shared__ double sh[2][8];
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, d);
    q += __shfl_xor_sync(0xffffffffu, q, d);
  }
  if ((threadIdx.x & 31) == 0) {
    sh[0][threadIdx.x >> 5] = s;
    sh[1][threadIdx.x >> 5] = q;
  }
  __syncthreads();
  if (threadIdx.x < 2) {
    double a = 0.0;
    for (int k = 0; k < 8; ++k) a += sh[threadIdx.x][k];
    atomicAdd(stats + 2 * b + threadIdx.x, a);
  }
}

__device__ __forceinline__ int reflect_idx(int i, int n) {
  // torch 'reflect' padding (edge not repeated); valid for |overshoot| < n
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}

// One view: [standardise x n_std] -> [Gaussian blur] -> [+ scale * noise] -> [contrast].  grid (tiles_x, tiles_y, B).
__global__ void __launch_bounds__(256)
augment_view_kernel(const float* __restrict__ img, const float* __restrict__ noise, float* __restrict__ out, int H,
                    int W, const double* __restrict__ stats, const float* __restrict__ params, float eps) {
  __shared__ float tin[AUG_IN][AUG_IN + 1];
  __shared__ float tmp[AUG_IN][AUG_TILE + 1];
  __shared__ float wk[2 * AUG_RMAX + 1];
  const int b = blockIdx.z;
  const float* prm = params + (size_t)b * AUG_NPARAM;
  const int ksize = (int)prm[0];
  const float sigma = prm[1], nscale = prm[2], alpha = prm[3], cmean = prm[4];
  const int n_std = (int)prm[5];
  const int R = ksize > 1 ? (ksize - 1) / 2 : 0;
  const size_t n = (size_t)H * W;
  const float* src = img + (size_t)b * n;

  // standardisation as x -> (x - m) / d, applied n_std times.  First pass: m = mean, d = std + eps (unbiased std, as
  // torch.Tensor.std()).  A second pass sees mean 0 (up to rounding) and std s' = std / (std + eps): d2 = s' + eps.
  // The two divisions are folded into one reciprocal per block (<= 1 ulp from the reference's per-element divisions).
  float m = 0.f, inv = 1.f;
  if (n_std > 0) {
    const double mean = stats[2 * b] / (double)n;
    double var = (stats[2 * b + 1] - mean * stats[2 * b]) / (double)(n > 1 ? n - 1 : 1);
    if (var < 0.0) var = 0.0;
    const float sd = (float)sqrt(var);
    m = (float)mean;
    const float d = sd + eps;
    const float d2 = (n_std > 1) ? sd / d + eps : 1.f;
    inv = 1.f / (d * d2);
  }
  auto load = [&](int y, int x) { return (__ldg(src + (size_t)y * W + x) - m) * inv; };

  const int x0 = blockIdx.x * AUG_TILE, y0 = blockIdx.y * AUG_TILE;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  if (R > 0) {
    if (threadIdx.x < ksize) {
      const float xx = (float)threadIdx.x - 0.5f * (float)(ksize - 1);
      const float t = xx / sigma;
      wk[threadIdx.x] = expf(-0.5f * t * t);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      float s = 0.f;
      for (int k = 0; k < ksize; ++k) s += wk[k];
      for (int k = 0; k < ksize; ++k) wk[k] = wk[k] / s;
    }
    const int ext = AUG_TILE + 2 * R;
    for (int i = threadIdx.x; i < ext * ext; i += 256) {
      const int r = i / ext, c = i - r * ext;
      const int gy = reflect_idx(y0 + r - R, H), gx = reflect_idx(x0 + c - R, W);
      // rows / columns beyond the image on the far side of an edge tile are never used by a valid output
      tin[r][c] = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? load(gy, gx) : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < ext * AUG_TILE; i += 256) {  // horizontal pass
      const int r = i >> 5, c = i & 31;
      float acc = 0.f;
      for (int k = 0; k < ksize; ++k) acc = fmaf(wk[k], tin[r][c + k], acc);
      tmp[r][c] = acc;
    }
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < AUG_TILE / 8; ++j) {
    const int ly = ty + 8 * j;
    const int y = y0 + ly, x = x0 + tx;
    if (y >= H || x >= W) continue;
    float v;
    if (R > 0) {
      v = 0.f;
      for (int k = 0; k < ksize; ++k) v = fmaf(wk[k], tmp[ly + k][tx], v);  // vertical pass
    } else {
      v = load(y, x);
    }
    const size_t o = (size_t)b * n + (size_t)y * W + x;
    if (noise != nullptr && nscale != 0.f) v = fmaf(nscale, __ldg(noise + o), v);
    if (alpha != 1.f) v = cmean + alpha * (v - cmean);
    out[o] = v;
  }
}

}  // namespace pda

using namespace pda;

extern "C" {

int pda_image_stats(const float* img, int B, long long n, double* stats, void* stream) {
  if (!img || !stats) return PDA_ERR_ARG;
  if (B <= 0 || B > 65535 || n <= 0) return PDA_ERR_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemsetAsync(stats, 0, sizeof(double) * 2 * B, st) != cudaSuccess) return PDA_ERR_CUDA;
  long long gx = (n + 256 * 16 - 1) / (256 * 16);
  const long long cap = (148 * 8 + B - 1) / B;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  PDA_COUNT(1);
  image_stats_kernel<<<dim3((unsigned)gx, B), 256, 0, st>>>(img, n, stats);
  return cudaGetLastError() == cudaSuccess ? PDA_OK : PDA_ERR_CUDA;
}

int pda_augment_view(const float* img, const float* noise, float* out, int B, int H, int W, const double* stats,
                     const float* params, float eps, int max_ksize, void* stream) {
  if (!img || !out || !stats || !params) return PDA_ERR_ARG;
  if (B <= 0 || B > 65535 || H <= 0 || W <= 0) return PDA_ERR_SHAPE;
  // max_ksize: the largest kernel size in the (device) parameter table, known to the host that sampled it
  if (max_ksize > 2 * AUG_RMAX + 1 || (max_ksize > 1 && ((max_ksize & 1) == 0 || max_ksize / 2 >= H || max_ksize / 2 >= W)))
    return PDA_ERR_SHAPE;
  const dim3 grid((W + AUG_TILE - 1) / AUG_TILE, (H + AUG_TILE - 1) / AUG_TILE, B);
  if (grid.y > 65535) return PDA_ERR_SHAPE;
  PDA_COUNT(1);
  augment_view_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(img, noise, out, H, W, stats, params, eps);
  return cudaGetLastError() == cudaSuccess ? PDA_OK : PDA_ERR_CUDA;
}

}  // extern "C"
