// Backward / loss kernels of the PUNet training step that are not tensor-core GEMMs:
// ReLU+pool backward, bilinear-upsample backward, first-layer weight gradient, Gaussian-head backward, KL backward,
// reconstruction loss (Dice / BCE with consensus mask) forward+backward, multi-tensor L2 norm forward+backward,
// Fcomb backward.  References are cited per kernel (paths relative to /root/reference/prob_utils).
#include "conv.cuh"
#include "ptx.cuh"

#include <math.h>

namespace pda {

__device__ __forceinline__ void unpack8f(const uint4& v, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8f(const float (&f)[8]) {
  uint4 o;
  o.x = pack_bf16x2(f[0], f[1]);
  o.y = pack_bf16x2(f[2], f[3]);
  o.z = pack_bf16x2(f[4], f[5]);
  o.w = pack_bf16x2(f[6], f[7]);
  return o;
}
static inline int grid_cap(long long total, int block, int cap = 148 * 16) {
  long long g = (total + block - 1) / block;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

// ------------------------------------------------------------------------------------------------
// dZ = (dFull + 0.25 * dPool[y/2][x/2]) * (Y > 0): backward of ReLU (unet_blocks.py:20) and of the
// AvgPool2d that consumes the block output (unet_blocks.py:17).  NHWC bf16; dFull / dPool may be NULL.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
relu_pool_bwd_kernel(const uint4* __restrict__ dfull, const uint4* __restrict__ dpool, const uint4* __restrict__ y,
                     uint4* __restrict__ dz, float* __restrict__ dbias, int B, int H, int W, int c_shift) {
  __shared__ float red[256 * 8];
  const unsigned C8 = 1u << c_shift;
  const unsigned Hp = H >> 1, Wp = W >> 1;
  const unsigned xc = blockIdx.x * blockDim.x + threadIdx.x;  // x * C8 + c: a thread keeps its 8 channels
  const bool active = xc < (unsigned)W * C8;
  const unsigned x = xc >> c_shift, c = xc & (C8 - 1);
  float bsum[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) bsum[i] = 0.f;
  if (active) {
    for (unsigned r = blockIdx.y; r < (unsigned)B * H; r += gridDim.y) {
      const size_t t = ((size_t)r * W << c_shift) + xc;
      float g[8], a[8], yv[8];
      if (dfull) {
        unpack8f(__ldg(dfull + t), g);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) g[i] = 0.f;
      }
      if (dpool) {
        const unsigned b = r / (unsigned)H, yy = r - b * H;
        unpack8f(__ldg(dpool + ((((size_t)b * Hp + (yy >> 1)) * Wp + (x >> 1)) << c_shift) + c), a);
#pragma unroll
        for (int i = 0; i < 8; ++i) g[i] = fmaf(0.25f, a[i], g[i]);
      }
      if (y) {  // y == nullptr: plain average-pool backward, no ReLU mask
        unpack8f(__ldg(y + t), yv);
#pragma unroll
        for (int i = 0; i < 8; ++i) g[i] = yv[i] > 0.f ? g[i] : 0.f;
      }
      const uint4 packed = pack8f(g);
      dz[t] = packed;
      if (dbias) {  // sum the ROUNDED values: exactly what the weight-gradient GEMM sees
        float rr[8];
        unpack8f(packed, rr);
#pragma unroll
        for (int i = 0; i < 8; ++i) bsum[i] += rr[i];
      }
    }
  }
  if (dbias) {
    // threads of a block with the same channel chunk: tid, tid + C8, ... (256 % C8 == 0)
#pragma unroll
    for (int i = 0; i < 8; ++i) red[threadIdx.x * 8 + i] = bsum[i];
    __syncthreads();
    if (threadIdx.x < C8) {
      for (unsigned rr = threadIdx.x + C8; rr < blockDim.x; rr += C8) {
#pragma unroll
        for (int i = 0; i < 8; ++i) bsum[i] += red[rr * 8 + i];
      }
      const unsigned cc = (blockIdx.x * blockDim.x + threadIdx.x) & (C8 - 1);
#pragma unroll
      for (int i = 0; i < 8; ++i) atomicAdd(dbias + cc * 8 + i, bsum[i]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// backward of F.interpolate(bilinear, x2, align_corners=True) (unet_blocks.py:51), gather form:
// every input pixel sums the output-gradient pixels whose footprint contains it (deterministic).
// ------------------------------------------------------------------------------------------------
// weight with which output coordinate Q (of 2n) contributes to input coordinate q (of n): the forward's own arithmetic
__device__ __forceinline__ float up_weight(int Q, int q, int n, float r) {
  if (Q < 0 || Q > 2 * n - 1) return 0.f;
  const float s = r * (float)Q;
  const int q1 = (int)s;
  const int qp = (q1 < n - 1) ? 1 : 0;
  const float l1 = s - q1, l0 = 1.f - l1;
  float wgt = 0.f;
  if (q1 == q) wgt += l0;
  if (q1 + qp == q) wgt += l1;
  return wgt;
}

// The six candidate columns' weights depend only on the thread's x: computed once per thread; the six row weights once
// per row; the image index comes from blockIdx.z.  (The first version re-derived every weight for each of the 36
// candidates of every pixel and divided by h per row: instruction-bound at 25 % of the HBM roofline.)
__global__ void __launch_bounds__(256)
upsample2x_bwd_kernel(const uint4* __restrict__ dout, uint4* __restrict__ din, int h, int w, int c_shift) {
  const int Ho = 2 * h, Wo = 2 * w;
  const unsigned C8 = 1u << c_shift;
  const float rh = (Ho > 1) ? (float)(h - 1) / (float)(Ho - 1) : 0.f;
  const float rw = (Wo > 1) ? (float)(w - 1) / (float)(Wo - 1) : 0.f;
  const unsigned xc = blockIdx.x * blockDim.x + threadIdx.x;
  if (xc >= (unsigned)w * C8) return;
  const int x = xc >> c_shift;
  const unsigned c = xc & (C8 - 1);
  float wx[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) wx[k] = up_weight(2 * x - 2 + k, x, w, rw);
  const uint4* gimg = dout + (((size_t)blockIdx.z * Ho * Wo) << c_shift) + c;
  uint4* dimg = din + (((size_t)blockIdx.z * h * w) << c_shift);
  for (int y = blockIdx.y; y < h; y += gridDim.y) {
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll
    for (int a = 0; a < 6; ++a) {
      const int Y = 2 * y - 2 + a;
      const float wy = up_weight(Y, y, h, rh);
      if (wy == 0.f) continue;
      const uint4* grow = gimg + (((size_t)Y * Wo) << c_shift);
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        if (wx[k] == 0.f) continue;
        float g[8];
        unpack8f(__ldg(grow + ((size_t)(2 * x - 2 + k) << c_shift)), g);
        const float wgt = wy * wx[k];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fmaf(wgt, g[i], acc[i]);
      }
    }
    dimg[((size_t)y * w << c_shift) + xc] = pack8f(acc);
  }
}

// ------------------------------------------------------------------------------------------------
// first layer (cin 1 or 2) weight / bias gradient: dW[co][ci][tap] = sum_p dZ[p][co] * x_ci[p + tap],
// dZ = dOut * (out > 0).  Mirrors conv_first_kernel: thread owns 8 output channels, accumulates in registers.
// ------------------------------------------------------------------------------------------------
template <int CIN>
__global__ void __launch_bounds__(256)
conv_first_bwd_kernel(const float* __restrict__ x0, const float* __restrict__ x1, const __nv_bfloat16* __restrict__ out,
                      const __nv_bfloat16* __restrict__ dout, float* __restrict__ part, int B, int H, int W, int cout) {
  extern __shared__ float red[];  // [cout * CIN * 9 + cout]
  const int groups = cout >> 3;
  const int g = threadIdx.x % groups;
  const int lanes_px = blockDim.x / groups;
  const int lpx = threadIdx.x / groups;
  const int nred = cout * CIN * 9 + cout;
  for (int i = threadIdx.x; i < nred; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
  float acc[CIN][9][8], accb[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    accb[j] = 0.f;
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci)
#pragma unroll
      for (int t = 0; t < 9; ++t) acc[ci][t][j] = 0.f;
  }
  // A thread walks QUADS of QP consecutive pixels of one row (32-bit indices; the host checks B * H * W * cout < 2^31):
  // 3 x (QP + 2) input values feed QP x 72 FMAs per input plane, and the QP 16-byte dZ (and forward-output) loads of the
  // NEXT quad are issued before this quad's FMAs -- with ~200 registers per thread only 8 warps are resident per SM, so
  // the work per load and the loads in flight per thread decide the speed (one pixel per iteration: 9 dependent scalar
  // loads per 72 FMAs, measured 0.12 ms for 1 M pixels against 0.03 ms of issue time).
  // out == nullptr: dout is already masked by the layer's own ReLU (it comes out of the dgrad conv of the next layer,
  // whose epilogue applies exactly that mask): half the HBM traffic
  constexpr int QP = CIN == 1 ? 4 : 2;
  const bool masked = out != nullptr;
  const unsigned quads_per_row = ((unsigned)W + QP - 1) / QP;
  const unsigned nquad = (unsigned)B * H * quads_per_row;
  const unsigned qstride = gridDim.x * lanes_px;
  unsigned q = blockIdx.x * lanes_px + lpx;
  uint4 q_dz[QP], q_y[QP];
  auto load_quad = [&](unsigned qq) {
    const unsigned row = qq / quads_per_row;
    const int xq = (int)(qq - row * quads_per_row) * QP;
    const size_t base = ((size_t)row * W + xq) * cout + g * 8;
#pragma unroll
    for (int px = 0; px < QP; ++px) {
      q_dz[px] = make_uint4(0, 0, 0, 0);
      q_y[px] = q_dz[px];
      if (xq + px < W) {
        q_dz[px] = __ldg(reinterpret_cast<const uint4*>(dout + base + (size_t)px * cout));
        if (masked) q_y[px] = __ldg(reinterpret_cast<const uint4*>(out + base + (size_t)px * cout));
      }
    }
  };
  if (q < nquad) load_quad(q);
  for (; q < nquad; q += qstride) {
    const unsigned row = q / quads_per_row;          // = b * H + y
    const int xq = (int)(q - row * quads_per_row) * QP;
    const int y = (int)(row % (unsigned)H);
    float dz[QP][8];
#pragma unroll
    for (int px = 0; px < QP; ++px) {
      float yv[8];
      unpack8f(q_dz[px], dz[px]);
      unpack8f(q_y[px], yv);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (masked) dz[px][j] = yv[j] > 0.f ? dz[px][j] : 0.f;   // (pixels past the row end were loaded as zeros)
        accb[j] += dz[px][j];
      }
    }
    if (q + qstride < nquad) load_quad(q + qstride);
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci) {
      const float* plane = (ci == 0 ? x0 : x1) + (size_t)(row - y) * W;   // start of image b
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int yy = y + ky - 1;
        const bool rowok = (yy >= 0) && (yy < H);
        const float* rp = plane + (size_t)yy * W;
        float in[QP + 2];
#pragma unroll
        for (int k = 0; k < QP + 2; ++k) {
          const int xx = xq + k - 1;
          in[k] = (rowok && xx >= 0 && xx < W) ? __ldg(rp + xx) : 0.f;
        }
#pragma unroll
        for (int kx = 0; kx < 3; ++kx)
#pragma unroll
          for (int px = 0; px < QP; ++px)
#pragma unroll
            for (int j = 0; j < 8; ++j)
              acc[ci][ky * 3 + kx][j] = fmaf(in[px + kx], dz[px][j], acc[ci][ky * 3 + kx][j]);
      }
    }
  }
  // block reduction.  With 8 channel groups (cout = 64) a warp holds 4 pixel lanes x 8 groups: fold the pixel lanes with
  // two shuffles first, so that only 8 lanes per warp touch the shared accumulators (the 32-way conflicting shared
  // atomics of the direct form cost ten times the accumulation loop itself).
  const bool fold = (groups == 8);
  const bool writer = !fold || (threadIdx.x & 31) < 8;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int co = g * 8 + j;
    float v = accb[j];
    if (fold) {
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
    }
    if (writer) atomicAdd(&red[cout * CIN * 9 + co], v);
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci)
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        float a = acc[ci][t][j];
        if (fold) {
          a += __shfl_xor_sync(0xffffffffu, a, 8);
          a += __shfl_xor_sync(0xffffffffu, a, 16);
        }
        if (writer) atomicAdd(&red[(co * CIN + ci) * 9 + t], a);
      }
  }
  __syncthreads();
  // per-block totals go to part[block][nred]; a second kernel sums the blocks in a fixed order.  (Global atomics from
  // ~300 blocks onto the same 640 / 1216 addresses serialise in L2: ~45 us of a 160 us launch, and not deterministic.)
  for (int i = threadIdx.x; i < nred; i += blockDim.x) part[(size_t)blockIdx.x * nred + i] = red[i];
}

// dw[i] (i < nw) and db[i - nw] = sum over blocks of part[block][i]
__global__ void __launch_bounds__(256)
conv_first_bwd_reduce_kernel(const float* __restrict__ part, int nblocks, int nw, int nred, float* __restrict__ dw,
                             float* __restrict__ db) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nred) return;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int k = 0;
  for (; k + 3 < nblocks; k += 4) {
    s0 += part[(size_t)k * nred + i];
    s1 += part[(size_t)(k + 1) * nred + i];
    s2 += part[(size_t)(k + 2) * nred + i];
    s3 += part[(size_t)(k + 3) * nred + i];
  }
  for (; k < nblocks; ++k) s0 += part[(size_t)k * nred + i];
  const float v = (s0 + s1) + (s2 + s3);
  if (i < nw) dw[i] = v; else db[i - nw] = v;
}

// ------------------------------------------------------------------------------------------------
// Gaussian head backward (probabilistic_unet.py:126-130): d(mu|log_sigma)[B][2L] ->
//   dW[o][c] = sum_b d[b][o] * mean[b][c],  db[o] = sum_b d[b][o],  dmean[b][c] = sum_o d[b][o] * W[o][c]
//   dEnc[b][p][c] = dmean[b][c] / P * (enc > 0)   (ReLU of the last encoder conv folded in)
// ------------------------------------------------------------------------------------------------
__global__ void gauss_head_bwd_small_kernel(const float* __restrict__ dmls, const float* __restrict__ w,
                                            const float* __restrict__ mean, float* __restrict__ dw,
                                            float* __restrict__ db, float* __restrict__ dmean, int B, int C, int nout,
                                            float inv_p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nout * C) {
    const int o = i / C, c = i % C;
    float s = 0.f;
    for (int b = 0; b < B; ++b) s = fmaf(dmls[b * nout + o], mean[(long long)b * C + c], s);
    dw[i] = s;
  }
  if (i < nout) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += dmls[b * nout + i];
    db[i] = s;
  }
  if (i < B * C) {
    const int b = i / C, c = i % C;
    float s = 0.f;
    for (int o = 0; o < nout; ++o) s = fmaf(dmls[b * nout + o], w[(long long)o * C + c], s);
    dmean[i] = s * inv_p;
  }
}

__global__ void gauss_head_bwd_enc_kernel(const float* __restrict__ dmean, const uint4* __restrict__ enc,
                                          uint4* __restrict__ denc, int B, int P, int c_shift) {
  const unsigned total = ((unsigned)B * P) << c_shift;
  for (unsigned t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    const unsigned c = t & ((1u << c_shift) - 1u);
    const unsigned b = (t >> c_shift) / (unsigned)P;
    float e[8], g[8];
    unpack8f(__ldg(enc + t), e);
    const float* dm = dmean + (((size_t)b << c_shift) + c) * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) g[i] = e[i] > 0.f ? dm[i] : 0.f;
    denc[t] = pack8f(g);
  }
}

// mean over pixels [B][C] from the stage-1 partial sums of the forward (kept for backward)
__global__ void mean_from_partials_kernel(const float* __restrict__ partial, float* __restrict__ mean, int C, int nchunk,
                                          float inv_p, int total) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int b = i / C, c = i % C;
  float s = 0.f;
  for (int k = 0; k < nchunk; ++k) s += partial[((long long)b * nchunk + k) * C + c];
  mean[i] = s * inv_p;
}

// ------------------------------------------------------------------------------------------------
// KL(q||p) backward for diagonal Gaussians parametrised by (mu | log_sigma) (probabilistic_unet.py:332)
// ------------------------------------------------------------------------------------------------
__global__ void kl_bwd_kernel(const float* __restrict__ q, const float* __restrict__ p, const float* __restrict__ dkl,
                              float* __restrict__ dq, float* __restrict__ dp, int B, int L) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * L) return;
  const int b = i / L, d = i % L;
  const float g = dkl[b];
  const float mq = q[b * 2 * L + d], lq = q[b * 2 * L + L + d];
  const float mp = p[b * 2 * L + d], lp = p[b * 2 * L + L + d];
  const float sp = expf(lp);
  const float r = expf(lq) / sp, vr = r * r;
  const float dm = (mq - mp) / sp;
  // kl = 0.5 * (vr + dm^2 - 1 - log vr);  vr = exp(2 lq - 2 lp)
  dq[b * 2 * L + d] = g * dm / sp;
  dq[b * 2 * L + L + d] = g * (vr - 1.f);
  dp[b * 2 * L + d] = -g * dm / sp;
  dp[b * 2 * L + L + d] = g * (1.f - vr - dm * dm);
}

// ------------------------------------------------------------------------------------------------
// Reconstruction loss (probabilistic_unet.py:347-369).  x = logit * c, t = segm * c (c = consensus, optional).
//   BCE : l = (1 - t) x - log_sigmoid(x)  (ATen binary_cross_entropy_with_logits), outputs sum and mean
//   Dice: 1 - 2 sum(p t) / max(sum p^2 + sum t^2, 1e-7), p = sigmoid(x)   (torch_em DiceLossWithLogits)
// stats[0..2] = (S0, S1, S2): BCE -> (sum l, -, -);  Dice -> (sum p t, sum p^2, sum t^2)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float consm_at(const float* cf, const long long* ci, long long i) {
  return cf ? cf[i] : (ci ? (float)ci[i] : 1.f);
}

__global__ void __launch_bounds__(256)
recon_loss_partial_kernel(const float* __restrict__ logits, const float* __restrict__ segm, const float* __restrict__ cf,
                          const long long* __restrict__ ci, long long n, int dice, double* __restrict__ partial) {
  double s0 = 0.0, s1 = 0.0, s2 = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float c = consm_at(cf, ci, i);
    const float x = logits[i] * c, t = segm[i] * c;
    if (dice) {
      const float p = 1.0f / (1.0f + expf(-x));
      s0 += (double)(p * t);
      s1 += (double)(p * p);
      s2 += (double)(t * t);
    } else {
      const float ls = fminf(x, 0.f) - log1pf(expf(-fabsf(x)));
      s0 += (double)((1.f - t) * x - ls);
    }
  }
  __shared__ double sh[3][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    s0 += __shfl_xor_sync(0xffffffffu, s0, d);
    s1 += __shfl_xor_sync(0xffffffffu, s1, d);
    s2 += __shfl_xor_sync(0xffffffffu, s2, d);
  }
  if (lane == 0) {
    sh[0][warp] = s0;
    sh[1][warp] = s1;
    sh[2][warp] = s2;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    double s = 0.0;
    for (int k = 0; k < 8; ++k) s += sh[threadIdx.x][k];
    partial[blockIdx.x * 3 + threadIdx.x] = s;
  }
}

// Final reductions over per-block partial sums: ONE warp, lane l owns partials l, l + 32, ... (fixed order), then a
// fixed shuffle tree -- deterministic, and 1/32 of the dependent-load chain of a single-thread loop (these kernels sit
// on the critical path of every training step).
__device__ __forceinline__ void warp_sum3(const double* __restrict__ partial, int nblocks, double& s0, double& s1,
                                          double& s2) {
  s0 = s1 = s2 = 0.0;
  for (int k = threadIdx.x & 31; k < nblocks; k += 32) {
    s0 += partial[3 * k];
    s1 += partial[3 * k + 1];
    s2 += partial[3 * k + 2];
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    s0 += __shfl_xor_sync(0xffffffffu, s0, d);
    s1 += __shfl_xor_sync(0xffffffffu, s1, d);
    s2 += __shfl_xor_sync(0xffffffffu, s2, d);
  }
}

__global__ void recon_loss_final_kernel(const double* __restrict__ partial, int nblocks, long long n, int dice,
                                        float* __restrict__ out, float* __restrict__ stats) {
  if (threadIdx.x >= 32 || blockIdx.x != 0) return;
  double s0, s1, s2;
  warp_sum3(partial, nblocks, s0, s1, s2);
  if (threadIdx.x != 0) return;
  stats[0] = (float)s0;
  stats[1] = (float)s1;
  stats[2] = (float)s2;
  if (dice) {
    const float den = fmaxf((float)(s1 + s2), 1e-7f);
    const float loss = 1.f - 2.f * ((float)s0 / den);
    out[0] = loss;  // sum of a scalar
    out[1] = loss;  // mean of a scalar
  } else {
    out[0] = (float)s0;
    out[1] = (float)(s0 / (double)n);
  }
}

__global__ void recon_loss_bwd_kernel(const float* __restrict__ logits, const float* __restrict__ segm,
                                      const float* __restrict__ cf, const long long* __restrict__ ci, long long n,
                                      int dice, const float* __restrict__ stats, const float* __restrict__ gout,
                                      float* __restrict__ dlogits) {
  // gout[0] = dL/d(sum), gout[1] = dL/d(mean)
  const float gs = gout[0], gm = gout[1];
  float k_num = 0.f, k_den = 0.f;
  if (dice) {
    const float num = stats[0], den_raw = stats[1] + stats[2];
    const float den = fmaxf(den_raw, 1e-7f);
    // loss = 1 - 2 num / den ; dloss/dp_i = -2 (t_i / den - num * 2 p_i / den^2) (second term only if not clamped)
    k_num = -2.f / den;
    k_den = (den_raw > 1e-7f) ? 4.f * num / (den * den) : 0.f;
  }
  const float g_bce = gs + gm / (float)n;
  const float g_dice = gs + gm;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float c = consm_at(cf, ci, i);
    const float x = logits[i] * c, t = segm[i] * c;
    const float p = 1.0f / (1.0f + expf(-x));
    float d;
    if (dice)
      d = g_dice * (k_num * t + k_den * p) * p * (1.f - p);
    else
      d = g_bce * (p - t);
    dlogits[i] = d * c;
  }
}

// ------------------------------------------------------------------------------------------------
// multi-tensor L2 norm (utils.py:32-40): out = sum over tensors of ||W_t||_2
// table rows: (ptr, numel, tensor_index, unused), chunks of <= 65536 elements
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) l2_partial_kernel(const long long* __restrict__ table, double* __restrict__ partial) {
  const long long* e = table + 4LL * blockIdx.x;
  const float* w = reinterpret_cast<const float*>(e[0]);
  const int n = (int)e[1];
  double s = 0.0;
  if ((reinterpret_cast<uintptr_t>(w) & 15) == 0) {
    const float4* w4 = reinterpret_cast<const float4*>(w);
    const int n4 = n >> 2;
    for (int i = threadIdx.x; i < n4; i += blockDim.x) {
      const float4 v = __ldg(w4 + i);
      // fp32 products are exact in double; four per load keeps the loop short
      s += ((double)v.x * v.x + (double)v.y * v.y) + ((double)v.z * v.z + (double)v.w * v.w);
    }
    for (int i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) s += (double)w[i] * (double)w[i];
  } else {
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += (double)w[i] * (double)w[i];
  }
  __shared__ double sh[8];
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int k = 0; k < 8; ++k) t += sh[k];
    partial[blockIdx.x] = t;
  }
}

__global__ void l2_final_kernel(const long long* __restrict__ table, const double* __restrict__ partial, int n_chunks,
                                int n_tensors, float* __restrict__ norms, float* __restrict__ out) {
  // one warp per tensor (warps stride over the tensors); lanes stride over ALL chunks and keep those of their tensor:
  // fixed partition + fixed shuffle tree = deterministic
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int t = warp; t < n_tensors; t += nwarps) {
    double s = 0.0;
    for (int k = lane; k < n_chunks; k += 32)
      if ((int)table[4LL * k + 2] == t) s += partial[k];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    if (lane == 0) norms[t] = (float)sqrt(s);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float acc = 0.f;  // same left-to-right fp32 sum as the reference's python loop
    for (int t = 0; t < n_tensors; ++t) acc += norms[t];
    out[0] = acc;
  }
}

// grad table rows: (w_ptr, byte offset of the gradient inside grad_base, numel, tensor_index)
__global__ void __launch_bounds__(256)
l2_bwd_kernel(const long long* __restrict__ table, const float* __restrict__ norms, const float* __restrict__ gout,
              char* __restrict__ grad_base) {
  const long long* e = table + 4LL * blockIdx.x;
  const float* w = reinterpret_cast<const float*>(e[0]);
  float* g = reinterpret_cast<float*>(grad_base + e[1]);
  const int n = (int)e[2];
  const float nrm = norms[(int)e[3]];
  const float k = nrm > 0.f ? gout[0] / nrm : 0.f;
  if (((reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(g)) & 15) == 0) {
    const int n4 = n >> 2;
    for (int i = threadIdx.x; i < n4; i += blockDim.x) {
      float4 v = __ldg(reinterpret_cast<const float4*>(w) + i);
      v.x *= k; v.y *= k; v.z *= k; v.w *= k;
      reinterpret_cast<float4*>(g)[i] = v;
    }
    for (int i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) g[i] = k * w[i];
  } else {
    for (int i = threadIdx.x; i < n; i += blockDim.x) g[i] = k * w[i];
  }
}

// ------------------------------------------------------------------------------------------------
// Fcomb backward for ONE latent sample (the training form, probabilistic_unet.py:200-214 under autograd).
// fp32 CUDA cores.  Persistent CTAs of 128 threads; one pixel per thread per 128-pixel tile of one image.
//   recompute: a1 = relu(W1f F + bz), h2 = W2 a1 + b2, a2 = relu(h2)
//   backward : da2 = g w3 (h2>0); da1 = W2^T da2 (a1>0); dF = W1f^T da1
//   weight grads accumulate in registers across the CTA's tiles (each thread owns a 4x8 block of dW2 and of
//   dW1f), column sums through shared-memory tiles; one atomicAdd pass per CTA at the end.
// ------------------------------------------------------------------------------------------------
constexpr int FB = 64;
constexpr int FB_LD = 68;  // padded row length of the shared tiles (floats): 16-byte aligned rows, conflict-free

__global__ void __launch_bounds__(128, 1)
fcomb_bwd_kernel(const __nv_bfloat16* __restrict__ feat, const float* __restrict__ z, const float* __restrict__ w1,
                 const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ b2,
                 const float* __restrict__ w3, const float* __restrict__ dlogit, int P, int L, int B, int tiles_per_img,
                 int num_tiles, __nv_bfloat16* __restrict__ dfeat, float* __restrict__ dw1f, float* __restrict__ dw2,
                 float* __restrict__ db2, float* __restrict__ dw3, float* __restrict__ db3, float* __restrict__ dbz,
                 const int* __restrict__ run_flag) {
  // run_flag != nullptr: fallback of the tensor-core backward, runs only when the forward raised its fp16 range flag
  if (run_flag != nullptr && *run_flag == 0) return;
  extern __shared__ __align__(16) float sm[];
  float* w1s = sm;                       // [64][64]
  float* w2s = w1s + FB * FB;            // [64][64]
  float* tF = w2s + FB * FB;             // [128][FB_LD]
  float* tA1 = tF + 128 * FB_LD;
  float* tD2 = tA1 + 128 * FB_LD;
  float* tD1 = tD2 + 128 * FB_LD;
  float* vec = tD1 + 128 * FB_LD;        // bz[64] b2[64] w3[64] dw3acc[64] db3acc[1]
  float* bzs = vec, *b2s = vec + 64, *w3s = vec + 128, *dw3acc = vec + 192, *db3acc = vec + 256;
  const int tid = threadIdx.x;
  const int kin = FB + L;
  for (int i = tid; i < FB * FB; i += 128) {
    w1s[i] = w1[(i / FB) * kin + (i % FB)];
    w2s[i] = w2[i];
  }
  if (tid < 64) {
    b2s[tid] = b2[tid];
    w3s[tid] = w3[tid];
    dw3acc[tid] = 0.f;
  }
  if (tid == 0) db3acc[0] = 0.f;
  // this thread's 4x8 block of each 64x64 weight gradient
  const int jb = (tid >> 3) * 4, ib = (tid & 7) * 8;
  float accW2[4][8], accW1[4][8];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < 8; ++c) accW2[a][c] = accW1[a][c] = 0.f;
  float accb2 = 0.f;  // column tid (tid < 64) of db2
  __syncthreads();

  for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const int b = tile / tiles_per_img;
    const int p0 = (tile - b * tiles_per_img) * 128;
    if (tid < 64) {
      float acc = b1[tid];
      for (int d = 0; d < L; ++d) acc = fmaf(w1[tid * kin + FB + d], z[b * L + d], acc);
      bzs[tid] = acc;
    }
    __syncthreads();
    const int pix = p0 + tid;
    const bool valid = pix < P;
    const size_t gp = (size_t)b * P + pix;
    float a1[FB];
    {
      float f[FB];
      if (valid) {
        const uint4* src = reinterpret_cast<const uint4*>(feat + gp * FB);
#pragma unroll
        for (int k = 0; k < FB / 8; ++k) {
          float t8[8];
          unpack8f(__ldg(src + k), t8);
#pragma unroll
          for (int i = 0; i < 8; ++i) f[8 * k + i] = t8[i];
        }
      } else {
#pragma unroll
        for (int i = 0; i < FB; ++i) f[i] = 0.f;
      }
#pragma unroll
      for (int i = 0; i < FB; i += 4)
        *reinterpret_cast<float4*>(tF + tid * FB_LD + i) = make_float4(f[i], f[i + 1], f[i + 2], f[i + 3]);
#pragma unroll 2
      for (int j = 0; j < FB; ++j) {
        const float4* wr = reinterpret_cast<const float4*>(w1s + j * FB);
        float s0 = bzs[j], s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
        for (int i = 0; i < FB / 4; ++i) {
          const float4 w = wr[i];
          s0 = fmaf(w.x, f[4 * i], s0);
          s1 = fmaf(w.y, f[4 * i + 1], s1);
          s2 = fmaf(w.z, f[4 * i + 2], s2);
          s3 = fmaf(w.w, f[4 * i + 3], s3);
        }
        tA1[tid * FB_LD + j] = fmaxf((s0 + s1) + (s2 + s3), 0.f);  // dynamic index: through shared memory
      }
    }
#pragma unroll
    for (int i = 0; i < FB; i += 4) {
      const float4 v = *reinterpret_cast<const float4*>(tA1 + tid * FB_LD + i);
      a1[i] = v.x;
      a1[i + 1] = v.y;
      a1[i + 2] = v.z;
      a1[i + 3] = v.w;
    }
    const float g = valid ? dlogit[gp] : 0.f;
    float da1[FB];
#pragma unroll
    for (int i = 0; i < FB; ++i) da1[i] = 0.f;
    float dw3_part = 0.f;
#pragma unroll 1
    for (int j = 0; j < FB; ++j) {
      const float4* wr = reinterpret_cast<const float4*>(w2s + j * FB);
      float s0 = b2s[j], s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
      for (int i = 0; i < FB / 4; ++i) {
        const float4 w = wr[i];
        s0 = fmaf(w.x, a1[4 * i], s0);
        s1 = fmaf(w.y, a1[4 * i + 1], s1);
        s2 = fmaf(w.z, a1[4 * i + 2], s2);
        s3 = fmaf(w.w, a1[4 * i + 3], s3);
      }
      const float h2 = (s0 + s1) + (s2 + s3);
      const float d2 = h2 > 0.f ? g * w3s[j] : 0.f;
      tD2[tid * FB_LD + j] = d2;
      // dW3[j] = sum_p g * relu(h2): warp reduce, one shared atomic per warp
      dw3_part = g * fmaxf(h2, 0.f);
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) dw3_part += __shfl_xor_sync(0xffffffffu, dw3_part, d);
      if ((tid & 31) == 0) atomicAdd(&dw3acc[j], dw3_part);
#pragma unroll
      for (int i = 0; i < FB / 4; ++i) {
        const float4 w = wr[i];
        da1[4 * i] = fmaf(w.x, d2, da1[4 * i]);
        da1[4 * i + 1] = fmaf(w.y, d2, da1[4 * i + 1]);
        da1[4 * i + 2] = fmaf(w.z, d2, da1[4 * i + 2]);
        da1[4 * i + 3] = fmaf(w.w, d2, da1[4 * i + 3]);
      }
    }
    {
      float gsum = g;
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) gsum += __shfl_xor_sync(0xffffffffu, gsum, d);
      if ((tid & 31) == 0) atomicAdd(db3acc, gsum);
    }
#pragma unroll
    for (int i = 0; i < FB; ++i) da1[i] = a1[i] > 0.f ? da1[i] : 0.f;
#pragma unroll
    for (int i = 0; i < FB; i += 4)
      *reinterpret_cast<float4*>(tD1 + tid * FB_LD + i) = make_float4(da1[i], da1[i + 1], da1[i + 2], da1[i + 3]);
    // dF = W1f^T da1 (a1 registers are dead now: reuse as the accumulator)
#pragma unroll
    for (int i = 0; i < FB; ++i) a1[i] = 0.f;
#pragma unroll 1
    for (int j = 0; j < FB; ++j) {
      const float4* wr = reinterpret_cast<const float4*>(w1s + j * FB);
      const float d = tD1[tid * FB_LD + j];
#pragma unroll
      for (int i = 0; i < FB / 4; ++i) {
        const float4 w = wr[i];
        a1[4 * i] = fmaf(w.x, d, a1[4 * i]);
        a1[4 * i + 1] = fmaf(w.y, d, a1[4 * i + 1]);
        a1[4 * i + 2] = fmaf(w.z, d, a1[4 * i + 2]);
        a1[4 * i + 3] = fmaf(w.w, d, a1[4 * i + 3]);
      }
    }
    if (valid) {
      uint4* dst = reinterpret_cast<uint4*>(dfeat + gp * FB);
#pragma unroll
      for (int k = 0; k < FB / 8; ++k) {
        float t8[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) t8[i] = a1[8 * k + i];
        dst[k] = pack8f(t8);
      }
    }
    __syncthreads();
    // weight-gradient partial sums over the tile's 128 pixels
#pragma unroll 2
    for (int p = 0; p < 128; ++p) {
      const float4 d2 = *reinterpret_cast<const float4*>(tD2 + p * FB_LD + jb);
      const float4 d1 = *reinterpret_cast<const float4*>(tD1 + p * FB_LD + jb);
      const float4 aa = *reinterpret_cast<const float4*>(tA1 + p * FB_LD + ib);
      const float4 ab = *reinterpret_cast<const float4*>(tA1 + p * FB_LD + ib + 4);
      const float4 fa = *reinterpret_cast<const float4*>(tF + p * FB_LD + ib);
      const float4 fb = *reinterpret_cast<const float4*>(tF + p * FB_LD + ib + 4);
      const float d2v[4] = {d2.x, d2.y, d2.z, d2.w}, d1v[4] = {d1.x, d1.y, d1.z, d1.w};
      const float av[8] = {aa.x, aa.y, aa.z, aa.w, ab.x, ab.y, ab.z, ab.w};
      const float fv[8] = {fa.x, fa.y, fa.z, fa.w, fb.x, fb.y, fb.z, fb.w};
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          accW2[a][c] = fmaf(d2v[a], av[c], accW2[a][c]);
          accW1[a][c] = fmaf(d1v[a], fv[c], accW1[a][c]);
        }
    }
    if (tid < 64) {
      float s2 = 0.f, s1 = 0.f;
      for (int p = 0; p < 128; ++p) {
        s2 += tD2[p * FB_LD + tid];
        s1 += tD1[p * FB_LD + tid];
      }
      accb2 += s2;
      atomicAdd(dbz + b * FB + tid, s1);  // per-image column sum of da1
    }
    __syncthreads();
  }
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      atomicAdd(dw2 + (jb + a) * FB + ib + c, accW2[a][c]);
      atomicAdd(dw1f + (jb + a) * FB + ib + c, accW1[a][c]);
    }
  if (tid < 64) {
    atomicAdd(db2 + tid, accb2);
    atomicAdd(dw3 + tid, dw3acc[tid]);
  }
  if (tid == 0) atomicAdd(db3, db3acc[0]);
}

// dbz [B][64] -> db1[64], dW1z[64][L] (written into the full dW1 [64][64+L]), dz[B][L]; also copies dW1f into dW1
__global__ void fcomb_bwd_finish_kernel(const float* __restrict__ dbz, const float* __restrict__ dw1f,
                                        const float* __restrict__ w1, const float* __restrict__ z,
                                        float* __restrict__ dw1, float* __restrict__ db1, float* __restrict__ dz, int B,
                                        int L) {
  const int kin = FB + L;
  const int tid = threadIdx.x;
  for (int i = tid; i < FB * FB; i += blockDim.x) dw1[(i / FB) * kin + (i % FB)] = dw1f[i];
  for (int i = tid; i < FB; i += blockDim.x) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += dbz[b * FB + i];
    db1[i] = s;
  }
  for (int i = tid; i < FB * L; i += blockDim.x) {
    const int j = i / L, d = i % L;
    float s = 0.f;
    for (int b = 0; b < B; ++b) s = fmaf(dbz[b * FB + j], z[b * L + d], s);
    dw1[j * kin + FB + d] = s;
  }
  for (int i = tid; i < B * L; i += blockDim.x) {
    const int b = i / L, d = i % L;
    float s = 0.f;
    for (int j = 0; j < FB; ++j) s = fmaf(w1[j * kin + FB + d], dbz[b * FB + j], s);
    dz[i] = s;
  }
}


// ------------------------------------------------------------------------------------------------
// multi-tensor Adam (torch.optim.Adam semantics, the optimizer of every reference script, e.g.
// LIVECell/livecell_punet.py:58): one launch for all parameter tensors, 28 B/param of HBM traffic.
// table rows: (param_ptr, grad_ptr, exp_avg_ptr, exp_avg_sq_ptr, numel <= 65536)
// scalars[0] = 1 / grad_scale (GradScaler unscale, 1 when unused); found_inf (optional) != 0 skips the update.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
adam_kernel(const long long* __restrict__ table, float lr, float beta2, float omb1, float omb2, float eps,
            float weight_decay, float bc1, float bc2_sqrt, const float* __restrict__ inv_scale, const float* __restrict__ found_inf,
            const float* __restrict__ lr_dev, const long long* __restrict__ step_dev, double beta1_d, double beta2_d) {
  if (found_inf && found_inf[0] != 0.f) return;
  if (step_dev) {
    // graph-capturable form: step count and learning rate live on the device (a replayed launch must not bake them in)
    const double step = (double)(step_dev[0] + 1);
    bc1 = (float)(1.0 - pow(beta1_d, step));
    bc2_sqrt = (float)sqrt(1.0 - pow(beta2_d, step));
    lr = lr_dev[0];
  }
  const long long* e = table + 5LL * blockIdx.x;
  float* p = reinterpret_cast<float*>(e[0]);
  const float* g = reinterpret_cast<const float*>(e[1]);
  float* m = reinterpret_cast<float*>(e[2]);
  float* v = reinterpret_cast<float*>(e[3]);
  const int n = (int)e[4];
  const float gs = inv_scale ? inv_scale[0] : 1.f;
  const float step_size = lr / bc1;
  auto update = [&](float& pi, float gi, float& mi, float& vi) {
    gi *= gs;
    if (weight_decay != 0.f) gi = fmaf(weight_decay, pi, gi);
    // torch: exp_avg.lerp_(grad, 1 - beta1); exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
    mi = mi + (gi - mi) * omb1;
    vi = fmaf(gi * gi, omb2, vi * beta2);
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    pi = pi - step_size * (mi / denom);
  };
  if (((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
        reinterpret_cast<uintptr_t>(v)) & 15) == 0) {
    const int n4 = n >> 2;
    for (int i = threadIdx.x; i < n4; i += blockDim.x) {
      float4 p4 = reinterpret_cast<float4*>(p)[i], m4 = reinterpret_cast<float4*>(m)[i],
             v4 = reinterpret_cast<float4*>(v)[i];
      const float4 g4 = __ldg(reinterpret_cast<const float4*>(g) + i);
      update(p4.x, g4.x, m4.x, v4.x);
      update(p4.y, g4.y, m4.y, v4.y);
      update(p4.z, g4.z, m4.z, v4.z);
      update(p4.w, g4.w, m4.w, v4.w);
      reinterpret_cast<float4*>(p)[i] = p4;
      reinterpret_cast<float4*>(m)[i] = m4;
      reinterpret_cast<float4*>(v)[i] = v4;
    }
    for (int i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) update(p[i], g[i], m[i], v[i]);
  } else {
    for (int i = threadIdx.x; i < n; i += blockDim.x) update(p[i], g[i], m[i], v[i]);
  }
}


// ------------------------------------------------------------------------------------------------
// FixMatch distribution alignment (fixmatch_trainer.py:77-84): the foreground / background frequencies of the binarised
// pseudo-label against the source-domain frequencies.  torch.unique(return_counts=True) (a sort + a host sync in the
// reference) becomes one counting pass; a batch with a single class yields target = [1.0], exactly what `unique`
// returning one count does to the reference's broadcast.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
count_fg_kernel(const float* __restrict__ y, long long n, unsigned long long* __restrict__ count) {
  unsigned int c = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    c += (y[i] >= 0.5f) ? 1u : 0u;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(count, (unsigned long long)c);
}

__global__ void __launch_bounds__(256)
dist_align_apply_kernel(const float* __restrict__ y, float* __restrict__ out, long long n,
                        const unsigned long long* __restrict__ count, const float* __restrict__ source,
                        float* __restrict__ ratio_out) {
  const unsigned long long n1 = count[0], n0 = (unsigned long long)n - n1;
  float t0 = 1.f, t1 = 1.f;
  if (n0 != 0 && n1 != 0) {
    t0 = (float)(long long)n0 / (float)n;
    t1 = (float)(long long)n1 / (float)n;
  }
  // single class: unique() returns ONE count -> target_distribution = [1.0], broadcast against both source entries
  const float r0 = source[0] / t0, r1 = source[1] / t1;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    ratio_out[0] = r0;
    ratio_out[1] = r1;
  }
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = y[i];
    const float a = (v < 0.5f) ? v * r0 : v * r1;
    out[i] = fminf(fmaxf(a, 0.f), 1.f);
  }
}

__global__ void adam_step_inc_kernel(long long* step_dev, const float* __restrict__ found_inf) {
  if (found_inf && found_inf[0] != 0.f) return;
  step_dev[0] += 1;
}

// ------------------------------------------------------------------------------------------------
// Validation metric (my_utils/util.py:17-44, called at punet_trainer.py:78-81): dice = 2 sum(gt * seg) /
// (sum(gt) + sum(seg) + 1e-7), optional thresholds (NaN = none: soft dice of the MC mean).  fp64 partial sums.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
dice_partial_kernel(const float* __restrict__ seg, const float* __restrict__ gt, long long n, float thr_seg,
                    float thr_gt, double* __restrict__ partial) {
  const bool ts = thr_seg == thr_seg, tg = thr_gt == thr_gt;  // NaN -> no threshold
  double s0 = 0.0, s1 = 0.0, s2 = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float a = seg[i], b = gt[i];
    if (ts) a = a > thr_seg ? 1.f : 0.f;
    if (tg) b = b > thr_gt ? 1.f : 0.f;
    s0 += (double)(a * b);
    s1 += (double)b;
    s2 += (double)a;
  }
  __shared__ double sh[3][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    s0 += __shfl_xor_sync(0xffffffffu, s0, d);
    s1 += __shfl_xor_sync(0xffffffffu, s1, d);
    s2 += __shfl_xor_sync(0xffffffffu, s2, d);
  }
  if (lane == 0) {
    sh[0][warp] = s0;
    sh[1][warp] = s1;
    sh[2][warp] = s2;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    double s = 0.0;
    for (int k = 0; k < 8; ++k) s += sh[threadIdx.x][k];
    partial[blockIdx.x * 3 + threadIdx.x] = s;
  }
}

__global__ void dice_final_kernel(const double* __restrict__ partial, int nblocks, float* __restrict__ out) {
  if (threadIdx.x >= 32 || blockIdx.x != 0) return;
  double s0, s1, s2;
  warp_sum3(partial, nblocks, s0, s1, s2);
  if (threadIdx.x != 0) return;
  out[0] = (float)(2.0 * s0 / (s1 + s2 + 1e-7));
}

}  // namespace pda

using namespace pda;
#define ST(s) ((cudaStream_t)(s))
#define LAUNCH_OK() (cudaGetLastError() == cudaSuccess ? PDA_OK : PDA_ERR_CUDA)

extern "C" {

int pda_relu_pool_bwd_bf16(const void* dfull, const void* dpool, const void* y, void* dz, float* dbias, int B, int H,
                           int W, int C, void* stream) {
  if (!dz || (!dfull && !dpool)) return PDA_ERR_ARG;
  if (c8_shift(C) < 0 || B <= 0 || (dpool && ((H & 1) || (W & 1)))) return PDA_ERR_SHAPE;
  const long long total = (long long)B * H * W * (C / 8);
  if (total >= 0x7fffffffLL) return PDA_ERR_SHAPE;
  if (dbias && cudaMemsetAsync(dbias, 0, sizeof(float) * C, ST(stream)) != cudaSuccess) return PDA_ERR_CUDA;
  PDA_COUNT(1);
  // with the fused bias reduction every block ends with C atomics on the same C addresses: keep the grid at two
  // blocks per SM (each thread then strides over more pixels, which is what feeds its register partial sums)
  relu_pool_bwd_kernel<<<row_grid((long long)W * (C / 8), (long long)B * H, dbias ? 148 * 2 : 148 * 8), 256, 0,
                         ST(stream)>>>(
      static_cast<const uint4*>(dfull), static_cast<const uint4*>(dpool), static_cast<const uint4*>(y),
      static_cast<uint4*>(dz), dbias, B, H, W, c8_shift(C));
  return LAUNCH_OK();
}

int pda_upsample2x_bilinear_bwd_bf16(const void* dout, void* din, int B, int h, int w, int C, void* stream) {
  if (!dout || !din) return PDA_ERR_ARG;
  if (c8_shift(C) < 0 || B <= 0 || h <= 0 || w <= 0) return PDA_ERR_SHAPE;
  const long long total = (long long)B * h * w * (C / 8);
  if (total >= 0x7fffffffLL) return PDA_ERR_SHAPE;
  PDA_COUNT(1);
  if (B > 65535) return PDA_ERR_SHAPE;
  dim3 grid = row_grid((long long)w * (C / 8), h, (148 * 8 + B - 1) / B);
  grid.z = B;
  upsample2x_bwd_kernel<<<grid, 256, 0, ST(stream)>>>(static_cast<const uint4*>(dout), static_cast<uint4*>(din), h, w,
                                                     c8_shift(C));
  return LAUNCH_OK();
}

static int first_bwd_grid(int B, int H, int W, int cout) {
  // one resident block per SM (242 registers per thread): one wave, at least one quad per thread
  return grid_cap((long long)B * H * ((W + 1) / 2) * (cout >> 3), 256, 148);
}

long long pda_conv3x3_first_bwd_scratch_floats(int B, int H, int W, int cout, int cin) {
  if (B <= 0 || H <= 0 || W <= 0 || cout <= 0 || cin <= 0) return 0;
  return (long long)first_bwd_grid(B, H, W, cout) * (cout * cin * 9 + cout);
}

int pda_conv3x3_first_bwd(const float* x0, const float* x1, const void* out, const void* dout, float* dw, float* db,
                          int B, int H, int W, int cout, float* scratch, void* stream) {
  if (!x0 || !dout || !dw || !db || !scratch) return PDA_ERR_ARG;  // out may be NULL: dout already ReLU-masked
  const int groups = cout >> 3;
  if (cout <= 0 || (cout & 7) || 256 % groups) return PDA_ERR_SHAPE;
  if ((long long)B * H * W * cout >= 0x7fffffffLL) return PDA_ERR_SHAPE;
  const int cin = x1 ? 2 : 1;
  cudaStream_t st = ST(stream);
  const int nw = cout * cin * 9, nred = nw + cout;
  const size_t smem = sizeof(float) * nred;
  const int grid = first_bwd_grid(B, H, W, cout);
  // per-block partial sums live in the caller's per-call scratch (no state shared between launches or streams)
  float* part = scratch;
  PDA_COUNT(2);
  if (x1)
    conv_first_bwd_kernel<2><<<grid, 256, smem, st>>>(x0, x1, static_cast<const __nv_bfloat16*>(out),
                                                      static_cast<const __nv_bfloat16*>(dout), part, B, H, W, cout);
  else
    conv_first_bwd_kernel<1><<<grid, 256, smem, st>>>(x0, x1, static_cast<const __nv_bfloat16*>(out),
                                                      static_cast<const __nv_bfloat16*>(dout), part, B, H, W, cout);
  conv_first_bwd_reduce_kernel<<<(nred + 255) / 256, 256, 0, st>>>(part, grid, nw, nred, dw, db);
  return LAUNCH_OK();
}

int pda_gauss_head_mean(const float* scratch, float* mean, int B, int P, int C, void* stream) {
  if (!scratch || !mean) return PDA_ERR_ARG;
  const int nchunk = pda_gauss_head_scratch_rows(P);
  const int total = B * C;
  PDA_COUNT(1);
  mean_from_partials_kernel<<<(total + 255) / 256, 256, 0, ST(stream)>>>(scratch, mean, C, nchunk, 1.f / (float)P,
                                                                         total);
  return LAUNCH_OK();
}

int pda_gauss_head_bwd(const float* dmls, const float* w_head, const float* mean, const void* enc, float* dw,
                       float* db, float* dmean_scratch, void* denc, int B, int P, int C, int latent, void* stream) {
  if (!dmls || !w_head || !mean || !enc || !dw || !db || !dmean_scratch || !denc) return PDA_ERR_ARG;
  if (c8_shift(C) < 0 || B <= 0 || P <= 0 || (long long)B * P * (C / 8) >= 0x7fffffffLL) return PDA_ERR_SHAPE;
  const int nout = 2 * latent;
  int n = nout * C;
  if (B * C > n) n = B * C;
  PDA_COUNT(2);
  gauss_head_bwd_small_kernel<<<(n + 255) / 256, 256, 0, ST(stream)>>>(dmls, w_head, mean, dw, db, dmean_scratch, B, C,
                                                                       nout, 1.f / (float)P);
  const long long total = (long long)B * P * (C / 8);
  gauss_head_bwd_enc_kernel<<<grid_cap(total, 256), 256, 0, ST(stream)>>>(
      dmean_scratch, static_cast<const uint4*>(enc), static_cast<uint4*>(denc), B, P, c8_shift(C));
  return LAUNCH_OK();
}

int pda_kl_diag_gauss_bwd(const float* q, const float* p, const float* dkl, float* dq, float* dp, int B, int latent,
                          void* stream) {
  if (!q || !p || !dkl || !dq || !dp) return PDA_ERR_ARG;
  const int n = B * latent;
  if (n <= 0) return PDA_ERR_SHAPE;
  PDA_COUNT(1);
  kl_bwd_kernel<<<(n + 127) / 128, 128, 0, ST(stream)>>>(q, p, dkl, dq, dp, B, latent);
  return LAUNCH_OK();
}

int pda_recon_loss_blocks(long long n) { return grid_cap(n, 256, 148 * 4); }

int pda_recon_loss_fwd(const float* logits, const float* segm, const float* consm_f32, const int64_t* consm_i64,
                       long long n, int dice, double* partial, float* out2, float* stats3, void* stream) {
  if (!logits || !segm || !partial || !out2 || !stats3) return PDA_ERR_ARG;
  if (n <= 0) return PDA_ERR_SHAPE;
  const int blocks = pda_recon_loss_blocks(n);
  PDA_COUNT(2);
  recon_loss_partial_kernel<<<blocks, 256, 0, ST(stream)>>>(logits, segm, consm_f32,
                                                            reinterpret_cast<const long long*>(consm_i64), n, dice,
                                                            partial);
  recon_loss_final_kernel<<<1, 32, 0, ST(stream)>>>(partial, blocks, n, dice, out2, stats3);
  return LAUNCH_OK();
}

int pda_recon_loss_bwd(const float* logits, const float* segm, const float* consm_f32, const int64_t* consm_i64,
                       long long n, int dice, const float* stats3, const float* gout2, float* dlogits, void* stream) {
  if (!logits || !segm || !stats3 || !gout2 || !dlogits) return PDA_ERR_ARG;
  if (n <= 0) return PDA_ERR_SHAPE;
  PDA_COUNT(1);
  recon_loss_bwd_kernel<<<grid_cap(n, 256), 256, 0, ST(stream)>>>(
      logits, segm, consm_f32, reinterpret_cast<const long long*>(consm_i64), n, dice, stats3, gout2, dlogits);
  return LAUNCH_OK();
}

int pda_multi_tensor_l2norm_fwd(const int64_t* table, int n_chunks, int n_tensors, double* partial, float* norms,
                                float* out, void* stream) {
  if (!table || !partial || !norms || !out) return PDA_ERR_ARG;
  if (n_chunks <= 0 || n_tensors <= 0) return PDA_ERR_SHAPE;
  PDA_COUNT(2);
  l2_partial_kernel<<<n_chunks, 256, 0, ST(stream)>>>(reinterpret_cast<const long long*>(table), partial);
  l2_final_kernel<<<1, 1024, 0, ST(stream)>>>(reinterpret_cast<const long long*>(table), partial, n_chunks, n_tensors,
                                             norms, out);
  return LAUNCH_OK();
}

int pda_multi_tensor_l2norm_bwd(const int64_t* grad_table, int n_chunks, const float* norms, const float* gout,
                                void* grad_base, void* stream) {
  if (!grad_table || !norms || !gout || !grad_base) return PDA_ERR_ARG;
  if (n_chunks <= 0) return PDA_ERR_SHAPE;
  PDA_COUNT(1);
  l2_bwd_kernel<<<n_chunks, 256, 0, ST(stream)>>>(reinterpret_cast<const long long*>(grad_table), norms, gout,
                                                  static_cast<char*>(grad_base));
  return LAUNCH_OK();
}

// zero-fills up to five fp32 buffers in ONE launch (the accumulation targets of the Fcomb backward)
struct ZeroList {
  float* p[5];
  int n[5];
};
__global__ void __launch_bounds__(256) zero_list_kernel(ZeroList zl) {
  float* p = zl.p[blockIdx.y];
  const int n = zl.n[blockIdx.y];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = 0.f;
}

static int fcomb_bwd_fp32_launch(const void* feat, const float* z, const float* w1, const float* b1, const float* w2,
                                 const float* b2, const float* w3, const float* dlogit, int B, int P, int latent,
                                 void* dfeat, float* dw1f, float* dw2, float* db2, float* dw3, float* db3, float* dbz,
                                 const int* run_flag, cudaStream_t st) {
  const int smem = sizeof(float) * (2 * FB * FB + 4 * 128 * FB_LD + 320);
  static int configured[64];
  if (dyn_smem_attr_needed(configured, smem)) {
    if (cudaFuncSetAttribute(fcomb_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
      return PDA_ERR_CUDA;
  }
  const int tiles_per_img = (P + 127) / 128;
  const long long num_tiles = (long long)tiles_per_img * B;
  if (num_tiles > 0x7fffffffLL) return PDA_ERR_SHAPE;
  const int grid = (int)(num_tiles < 148 ? num_tiles : 148);
  PDA_COUNT(1);
  fcomb_bwd_kernel<<<grid, 128, smem, st>>>(static_cast<const __nv_bfloat16*>(feat), z, w1, b1, w2, b2, w3, dlogit, P,
                                            latent, B, tiles_per_img, (int)num_tiles,
                                            static_cast<__nv_bfloat16*>(dfeat), dw1f, dw2, db2, dw3, db3, dbz,
                                            run_flag);
  return LAUNCH_OK();
}

static int fcomb_bwd_zero(float* scratch, int B, float* dw2, float* db2, float* dw3, float* db3, cudaStream_t st) {
  ZeroList zl;
  zl.p[0] = scratch; zl.n[0] = FB * FB + B * FB;
  zl.p[1] = dw2;     zl.n[1] = FB * FB;
  zl.p[2] = db2;     zl.n[2] = FB;
  zl.p[3] = dw3;     zl.n[3] = FB;
  zl.p[4] = db3;     zl.n[4] = 1;
  PDA_COUNT(1);
  zero_list_kernel<<<dim3((FB * FB + B * FB + 255) / 256, 5), 256, 0, st>>>(zl);
  return LAUNCH_OK();
}

int pda_fcomb_bwd(const void* feat, const float* z, const float* w1, const float* b1, const float* w2, const float* b2,
                  const float* w3, const float* dlogit, int B, int P, int latent, void* dfeat, float* dw1, float* db1,
                  float* dw2, float* db2, float* dw3, float* db3, float* dz, float* scratch, const int* fwd_range_flag,
                  void* stream) {
  // scratch: fp32, at least 64*64 + 2*B*64 elements (dW1f accumulator, per-image column sums, per-image layer-1 bias)
  if (!feat || !z || !w1 || !b1 || !w2 || !b2 || !w3 || !dlogit || !dfeat || !dw1 || !db1 || !dw2 || !db2 || !dw3 ||
      !db3 || !dz || !scratch)
    return PDA_ERR_ARG;
  if (B <= 0 || P <= 0 || latent <= 0 || B > (1 << 20)) return PDA_ERR_SHAPE;
  cudaStream_t st = ST(stream);
  float* dw1f = scratch;
  float* dbz = scratch + FB * FB;
  float* bz = dbz + (size_t)B * FB;
  int r = fcomb_bwd_zero(scratch, B, dw2, db2, dw3, db3, st);
  if (r) return r;
  r = fcomb_bwd_tc(feat, z, w1, b1, w2, b2, w3, dlogit, B, P, latent, dfeat, dw1f, dw2, db2, dw3, db3, dbz, bz,
                   fwd_range_flag, st);
  if (r) return r;
  if (fwd_range_flag != nullptr) {
    r = fcomb_bwd_fp32_launch(feat, z, w1, b1, w2, b2, w3, dlogit, B, P, latent, dfeat, dw1f, dw2, db2, dw3, db3, dbz,
                              fwd_range_flag, st);
    if (r) return r;
  }
  PDA_COUNT(1);
  fcomb_bwd_finish_kernel<<<1, 256, 0, st>>>(dbz, dw1f, w1, z, dw1, db1, dz, B, latent);
  return LAUNCH_OK();
}

int pda_fcomb_bwd_fp32(const void* feat, const float* z, const float* w1, const float* b1, const float* w2, const float* b2,
                  const float* w3, const float* dlogit, int B, int P, int latent, void* dfeat, float* dw1, float* db1,
                  float* dw2, float* db2, float* dw3, float* db3, float* dz, float* scratch, void* stream) {
  // scratch: fp32, at least 64*64 + B*64 elements (dW1f accumulator, per-image column sums)
  if (!feat || !z || !w1 || !b1 || !w2 || !b2 || !w3 || !dlogit || !dfeat || !dw1 || !db1 || !dw2 || !db2 || !dw3 ||
      !db3 || !dz || !scratch)
    return PDA_ERR_ARG;
  if (B <= 0 || P <= 0 || latent <= 0 || B > (1 << 20)) return PDA_ERR_SHAPE;
  cudaStream_t st = ST(stream);
  float* dw1f = scratch;
  float* dbz = scratch + FB * FB;
  int r = fcomb_bwd_zero(scratch, B, dw2, db2, dw3, db3, st);
  if (r) return r;
  r = fcomb_bwd_fp32_launch(feat, z, w1, b1, w2, b2, w3, dlogit, B, P, latent, dfeat, dw1f, dw2, db2, dw3, db3, dbz,
                            nullptr, st);
  if (r) return r;
  PDA_COUNT(1);
  fcomb_bwd_finish_kernel<<<1, 256, 0, st>>>(dbz, dw1f, w1, z, dw1, db1, dz, B, latent);
  return LAUNCH_OK();
}

int pda_multi_tensor_adam(const int64_t* table, int n_chunks, double lr, double beta1, double beta2, double eps,
                          double weight_decay, long long step, const float* inv_scale, const float* found_inf,
                          void* stream) {
  if (!table) return PDA_ERR_ARG;
  if (n_chunks <= 0 || step <= 0) return PDA_ERR_SHAPE;
  const double bc1 = 1.0 - pow(beta1, (double)step);
  const double bc2 = 1.0 - pow(beta2, (double)step);
  PDA_COUNT(1);
  adam_kernel<<<n_chunks, 256, 0, ST(stream)>>>(reinterpret_cast<const long long*>(table), (float)lr, (float)beta2,
                                                (float)(1.0 - beta1), (float)(1.0 - beta2), (float)eps, (float)weight_decay, (float)bc1,
                                                (float)sqrt(bc2), inv_scale, found_inf, nullptr, nullptr, 0.0, 0.0);
  return LAUNCH_OK();
}

int pda_multi_tensor_adam_capturable(const int64_t* table, int n_chunks, const float* lr_dev, double beta1, double beta2,
                                     double eps, double weight_decay, int64_t* step_dev, const float* inv_scale,
                                     const float* found_inf, void* stream) {
  if (!table || !lr_dev || !step_dev) return PDA_ERR_ARG;
  if (n_chunks <= 0) return PDA_ERR_SHAPE;
  PDA_COUNT(2);
  adam_kernel<<<n_chunks, 256, 0, ST(stream)>>>(reinterpret_cast<const long long*>(table), 0.f, (float)beta2,
                                                (float)(1.0 - beta1), (float)(1.0 - beta2), (float)eps,
                                                (float)weight_decay, 1.f, 1.f, inv_scale, found_inf, lr_dev,
                                                reinterpret_cast<const long long*>(step_dev), beta1, beta2);
  adam_step_inc_kernel<<<1, 1, 0, ST(stream)>>>(reinterpret_cast<long long*>(step_dev), found_inf);
  return LAUNCH_OK();
}

int pda_distribution_alignment(const float* y, long long n, const float* source_dist, unsigned long long* scratch,
                               float* out, float* ratio, void* stream) {
  if (!y || !source_dist || !scratch || !out || !ratio) return PDA_ERR_ARG;
  if (n <= 0) return PDA_ERR_SHAPE;
  if (cudaMemsetAsync(scratch, 0, sizeof(unsigned long long), ST(stream)) != cudaSuccess) return PDA_ERR_CUDA;
  const int blocks = grid_cap(n, 256, 148 * 4);
  PDA_COUNT(2);
  count_fg_kernel<<<blocks, 256, 0, ST(stream)>>>(y, n, scratch);
  dist_align_apply_kernel<<<blocks, 256, 0, ST(stream)>>>(y, out, n, scratch, source_dist, ratio);
  return LAUNCH_OK();
}

int pda_dice_score(const float* seg, const float* gt, long long n, float thr_seg, float thr_gt, double* partial,
                   float* out, void* stream) {
  if (!seg || !gt || !partial || !out) return PDA_ERR_ARG;
  if (n <= 0) return PDA_ERR_SHAPE;
  const int blocks = pda_recon_loss_blocks(n);
  PDA_COUNT(2);
  dice_partial_kernel<<<blocks, 256, 0, ST(stream)>>>(seg, gt, n, thr_seg, thr_gt, partial);
  dice_final_kernel<<<1, 32, 0, ST(stream)>>>(partial, blocks, out);
  return LAUNCH_OK();
}

}  // extern "C"
