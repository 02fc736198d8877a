// HBM-bound helper kernels of the PUNet path: weight packing, first-layer direct conv, 2x2 average pool,
// bilinear x2 upsample, Gaussian head (global mean + 1x1), latent sampling, KL, multi-tensor EMA,
// and a plain CUDA-core conv3x3 used as the on-GPU cross-check of the tcgen05 kernel.
#include <stdlib.h>

#include <atomic>

#include "conv.cuh"
#include "ptx.cuh"

std::atomic<long long> g_pda_launches{0};

namespace pda {

// ------------------------------------------------------------------------------------------------
// OIHW fp32 -> [cout][tap][cin] bf16   (rot180: [cin][8-tap][cout], the dgrad operand)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint16_t w16(float v, int f16) {
  // 16-bit weight operand: bf16, or fp16 (saturating) for the fp16-activation inference path
  return f16 ? (uint16_t)(pack_f16x2_sat(v, 0.f) & 0xffffu) : (uint16_t)(pack_bf16x2(v, 0.f) & 0xffffu);
}

__global__ void pack_w_kernel(const float* __restrict__ w, uint16_t* __restrict__ o, int cout, int cin,
                              int rot180, int f16) {
  const long long n = 9LL * cout * cin;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    // i indexes the OUTPUT so writes are coalesced
    if (!rot180) {
      const int ci = i % cin;
      const int tap = (i / cin) % 9;
      const int co = i / (9LL * cin);
      o[i] = w16(w[((long long)co * cin + ci) * 9 + tap], f16);
    } else {
      const int co = i % cout;
      const int tap = (i / cout) % 9;
      const int ci = i / (9LL * cout);
      o[i] = w16(w[((long long)co * cin + ci) * 9 + (8 - tap)], f16);
    }
  }
}

// multi-tensor variant: one launch refreshes the 16-bit operands of many convs after an optimizer / EMA step.
// One block = one 32 (co) x 32 (ci) x 9 (tap) tile of one conv.  table rows: (w_ptr, bf16 packed_ptr or 0, bf16 rot_ptr
// or 0, fp16 packed_ptr or 0, cout, cin, co0, ci0).  The tile goes through shared memory: for one co the 32 ci x 9 taps
// are 288 CONTIGUOUS floats of the OIHW source (coalesced reads; the first version read with a stride of 9 floats and
// took 0.27 ms per training step for ~300 MB of traffic); the outputs are written as 64-byte runs along ci (forward
// operand) and along co (rot180 operand).
__global__ void __launch_bounds__(256) pack_w_multi_kernel(const long long* __restrict__ table) {
  __shared__ float tile[32][289];
  const long long* e = table + 8LL * blockIdx.x;
  const float* w = reinterpret_cast<const float*>(e[0]);
  uint16_t* o = reinterpret_cast<uint16_t*>(e[1]);
  uint16_t* orot = reinterpret_cast<uint16_t*>(e[2]);
  uint16_t* oh = reinterpret_cast<uint16_t*>(e[3]);
  const int cout = (int)e[4], cin = (int)e[5], co0 = (int)e[6], ci0 = (int)e[7];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int co = warp + 8 * r;
    const float* src = w + (static_cast<size_t>(co0 + co) * cin + ci0) * 9;
#pragma unroll
    for (int k = 0; k < 9; ++k) tile[co][k * 32 + lane] = src[k * 32 + lane];
  }
  __syncthreads();
  // forward operands [cout][tap][cin]: rows (co, tap), 32 consecutive ci each
  for (int row = warp; row < 32 * 9; row += 8) {
    const int co = row / 9, tap = row - co * 9;
    const float v = tile[co][lane * 9 + tap];
    const size_t idx = (static_cast<size_t>(co0 + co) * 9 + tap) * cin + ci0 + lane;
    if (o) o[idx] = w16(v, 0);
    if (oh) oh[idx] = w16(v, 1);
  }
  // rot180 operand [cin][8 - tap][cout]: rows (ci, tap), 32 consecutive co each
  if (orot) {
    for (int row = warp; row < 32 * 9; row += 8) {
      const int ci = row / 9, tap = row - ci * 9;
      orot[(static_cast<size_t>(ci0 + ci) * 9 + (8 - tap)) * cout + co0 + lane] = w16(tile[lane][ci * 9 + tap], 0);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// first layer: cin in {1,2} fp32 planes -> NHWC bf16, fp32 math (K = 9 / 18 is too thin for the tensor cores; the
// layer is bound by its 128 B/px output).  A thread owns 8 output channels of 4 consecutive pixels of one row:
// 18 (36) input loads and 18 (36) 128-bit shared-memory weight reads feed 288 (576) FMAs; the 8 threads of a pixel
// write one contiguous 128-byte line.  All index math is 32-bit, one division per 4 pixels.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t pack_f32x2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack_f32x2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t fma_f32x2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

template <int CIN, bool F16>
__global__ void __launch_bounds__(256)
conv_first_kernel(const float* __restrict__ x0, const float* __restrict__ x1, const float* __restrict__ w,
                  const float* __restrict__ bias, __nv_bfloat16* __restrict__ out, int B, int H, int W, int cout,
                  int relu) {
  extern __shared__ __align__(16) float wsm[];  // [CIN][9][cout] then bias[cout]
  float* bsm = wsm + CIN * 9 * cout;
  for (int i = threadIdx.x; i < CIN * 9 * cout; i += blockDim.x) {
    const int co = i % cout, t = (i / cout) % 9, ci = i / (9 * cout);
    wsm[i] = w[(co * CIN + ci) * 9 + t];  // OIHW -> [ci][tap][co]
  }
  for (int i = threadIdx.x; i < cout; i += blockDim.x) bsm[i] = bias[i];
  __syncthreads();
  const int groups = cout >> 3;                    // 8-channel groups per pixel
  const int g = threadIdx.x % groups;
  const int quads_per_blk = blockDim.x / groups;
  const int lq = threadIdx.x / groups;
  const unsigned quads_per_row = (unsigned)(W + 3) >> 2;
  const unsigned total_quads = quads_per_row * (unsigned)H * (unsigned)B;
  for (unsigned q = blockIdx.x * quads_per_blk + lq; q < total_quads; q += gridDim.x * quads_per_blk) {
    const unsigned row = q / quads_per_row;        // = b * H + y
    const int xq = (int)(q - row * quads_per_row) << 2;
    const int y = (int)(row % (unsigned)H);
    // accumulators as packed fp32 pairs: the kernel is bound by the FMA pipe (9 x 64 FMAs per pixel = 0.13 ms for
    // 4 x 1024^2 at 64 FMA/clk/SM), and one FFMA2 retires two of them (bit-identical to the scalar FMAs)
    uint64_t acc2[4][4];
#pragma unroll
    for (int px = 0; px < 4; ++px)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc2[px][j] = pack_f32x2(bsm[g * 8 + 2 * j], bsm[g * 8 + 2 * j + 1]);
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci) {
      const float* plane = (ci == 0 ? x0 : x1) + (size_t)(row - y) * W;  // start of image b
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int yy = y + ky - 1;
        float in[6];
        const bool rowok = (yy >= 0) && (yy < H);
        const float* rp = plane + (size_t)yy * W;
#pragma unroll
        for (int k = 0; k < 6; ++k) {
          const int xx = xq + k - 1;
          in[k] = (rowok && xx >= 0 && xx < W) ? __ldg(rp + xx) : 0.f;
        }
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const float4 wa = *reinterpret_cast<const float4*>(wsm + (ci * 9 + ky * 3 + kx) * cout + g * 8);
          const float4 wb = *reinterpret_cast<const float4*>(wsm + (ci * 9 + ky * 3 + kx) * cout + g * 8 + 4);
          const uint64_t w2[4] = {pack_f32x2(wa.x, wa.y), pack_f32x2(wa.z, wa.w), pack_f32x2(wb.x, wb.y),
                                  pack_f32x2(wb.z, wb.w)};
#pragma unroll
          for (int px = 0; px < 4; ++px) {
            const uint64_t v2 = pack_f32x2(in[px + kx], in[px + kx]);
#pragma unroll
            for (int j = 0; j < 4; ++j) acc2[px][j] = fma_f32x2(v2, w2[j], acc2[px][j]);
          }
        }
      }
    }
    __nv_bfloat16* orow = out + ((size_t)row * W + xq) * cout + g * 8;
#pragma unroll
    for (int px = 0; px < 4; ++px) {
      if (xq + px < W) {
        float a[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) unpack_f32x2(acc2[px][j], a[2 * j], a[2 * j + 1]);
        if (relu) {
#pragma unroll
          for (int j = 0; j < 8; ++j) a[j] = fmaxf(a[j], 0.f);
        }
        uint4 o;
        o.x = pack_act2<F16>(a[0], a[1]);
        o.y = pack_act2<F16>(a[2], a[3]);
        o.z = pack_act2<F16>(a[4], a[5]);
        o.w = pack_act2<F16>(a[6], a[7]);
        *reinterpret_cast<uint4*>(orow + (size_t)px * cout) = o;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// 2x2 average pool, NHWC bf16, 8 channels (16 B) per thread
// ------------------------------------------------------------------------------------------------
template <bool F16 = false>
__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = unpack_act2<F16>(w4[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
template <bool F16 = false>
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 o;
  o.x = pack_act2<F16>(f[0], f[1]);
  o.y = pack_act2<F16>(f[2], f[3]);
  o.z = pack_act2<F16>(f[4], f[5]);
  o.w = pack_act2<F16>(f[6], f[7]);
  return o;
}

// The NHWC elementwise kernels use a 2-D decomposition: blockIdx.x / threadIdx.x cover (x, 16-byte channel chunk) of
// one output row, blockIdx.y strides over the B * H output rows: one 32-bit division per row instead of three 64-bit
// ones per element.
template <bool F16>
__global__ void __launch_bounds__(256)
avgpool2_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int B, int H, int W, int c_shift) {
  const unsigned Ho = H >> 1, Wo = W >> 1;
  const unsigned C8 = 1u << c_shift;
  const unsigned xc = blockIdx.x * blockDim.x + threadIdx.x;  // x * C8 + c
  if (xc >= Wo * C8) return;
  const unsigned x = xc >> c_shift, c = xc & (C8 - 1);
  for (unsigned r = blockIdx.y; r < (unsigned)B * Ho; r += gridDim.y) {
    const unsigned b = r / Ho, y = r - b * Ho;
    const size_t base = ((((size_t)b * H + 2 * y) * W + 2 * x) << c_shift) + c;
    float a[8], s[8];
    unpack8<F16>(__ldg(in + base), s);
    unpack8<F16>(__ldg(in + base + C8), a);
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] += a[i];
    unpack8<F16>(__ldg(in + base + (size_t)W * C8), a);
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] += a[i];
    unpack8<F16>(__ldg(in + base + (size_t)W * C8 + C8), a);
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] = (s[i] + a[i]) * 0.25f;
    out[((size_t)r * Wo << c_shift) + xc] = pack8<F16>(s);
  }
}

// ------------------------------------------------------------------------------------------------
// bilinear x2, align_corners=True (ATen upsample_bilinear2d arithmetic: src = dst*(in-1)/(out-1))
//
// One thread = one 2 x 2 block of output pixels x 8 channels.  With scale (h-1)/(2h-1) < 1/2 the two output rows 2i, 2i+1
// read the input rows y1, y1+1 and y1+1, y1+2 (same for the columns): a 3 x 3 input neighbourhood serves four outputs --
// 9 loads / bf16 unpacks instead of 16, the six horizontal blends are shared between the two output rows, and the
// image index comes from blockIdx.z (no integer division in the loop).  The first version (one output per thread)
// was instruction-bound (ncu: ALU pipe 61 %, issue slots 73 %, 38 % of the HBM roofline).  The first row / column pair
// (both outputs start in the same input row / column) takes the plain four-neighbour path.
// ------------------------------------------------------------------------------------------------
template <bool F16>
__device__ __forceinline__ float2 bf2_to_f2(uint32_t w) { return unpack_act2<F16>(w); }
__device__ __forceinline__ uint32_t word_of(const uint4& v, int k) {
  return k == 0 ? v.x : k == 1 ? v.y : k == 2 ? v.z : v.w;
}
__device__ __forceinline__ void set_word(uint4& v, int k, uint32_t w) {
  if (k == 0) v.x = w; else if (k == 1) v.y = w; else if (k == 2) v.z = w; else v.w = w;
}

// one output pixel from its four neighbours (the arithmetic of the reference, channel pair by channel pair)
template <bool F16>
__device__ __forceinline__ uint4 bilerp4(const uint4& q00, const uint4& q01, const uint4& q10, const uint4& q11,
                                         float lx0, float lx1, float l0, float l1) {
  uint4 o;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 a = bf2_to_f2<F16>(word_of(q00, k)), b = bf2_to_f2<F16>(word_of(q01, k));
    const float2 c = bf2_to_f2<F16>(word_of(q10, k)), d = bf2_to_f2<F16>(word_of(q11, k));
    const float ox = blend2(l0, l1, blend2(lx0, lx1, a.x, b.x), blend2(lx0, lx1, c.x, d.x));
    const float oy = blend2(l0, l1, blend2(lx0, lx1, a.y, b.y), blend2(lx0, lx1, c.y, d.y));
    set_word(o, k, pack_act2<F16>(ox, oy));
  }
  return o;
}

template <bool F16>
__global__ void __launch_bounds__(256)
upsample2x_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int h, int w, int c_shift) {
  const unsigned Ho = 2 * h, Wo = 2 * w;
  const unsigned C8 = 1u << c_shift;
  const float rh = (Ho > 1) ? (float)(h - 1) / (float)(Ho - 1) : 0.f;
  const float rw = (Wo > 1) ? (float)(w - 1) / (float)(Wo - 1) : 0.f;
  const unsigned jc = blockIdx.x * blockDim.x + threadIdx.x;  // j * C8 + c  (j = input column = output column pair)
  if (jc >= (unsigned)w * C8) return;
  const unsigned j = jc >> c_shift, c = jc & (C8 - 1);
  const uint4* img = in + (((size_t)blockIdx.z * h * w) << c_shift) + c;
  uint4* oimg = out + (((size_t)blockIdx.z * Ho * Wo) << c_shift) + c;
  // columns 2j, 2j+1
  const float sxa = rw * (float)(2 * j), sxb = rw * (float)(2 * j + 1);
  const int xa = (int)sxa, xb = (int)sxb;
  const float lxa1 = sxa - xa, lxa0 = 1.f - lxa1, lxb1 = sxb - xb, lxb0 = 1.f - lxb1;
  const int xpa = (xa < w - 1) ? 1 : 0, xpb = (xb < w - 1) ? 1 : 0;
  const bool fast_x = (xb == xa + 1);
  for (unsigned i = blockIdx.y; i < (unsigned)h; i += gridDim.y) {
    const float sya = rh * (float)(2 * i), syb = rh * (float)(2 * i + 1);
    const int ya = (int)sya, yb = (int)syb;
    const float lya1 = sya - ya, lya0 = 1.f - lya1, lyb1 = syb - yb, lyb0 = 1.f - lyb1;
    const int ypa = (ya < h - 1) ? 1 : 0, ypb = (yb < h - 1) ? 1 : 0;
    uint4* o0 = oimg + (((size_t)(2 * i) * Wo + 2 * j) << c_shift);
    uint4* o1 = o0 + ((size_t)Wo << c_shift);
    if (fast_x && yb == ya + 1) {
      // rows ya, ya+1 (= yb), yb+ypb; columns xa, xa+1 (= xb), xb+xpb
      const uint4* r0 = img + (((size_t)ya * w + xa) << c_shift);
      const uint4* r1 = r0 + ((size_t)w << c_shift);
      const uint4* r2 = r1 + (((size_t)ypb * w) << c_shift);
      const size_t c1 = C8, c2 = (size_t)(1 + xpb) << c_shift;
      const uint4 q00 = __ldg(r0), q01 = __ldg(r0 + c1), q02 = __ldg(r0 + c2);
      const uint4 q10 = __ldg(r1), q11 = __ldg(r1 + c1), q12 = __ldg(r1 + c2);
      const uint4 q20 = __ldg(r2), q21 = __ldg(r2 + c1), q22 = __ldg(r2 + c2);
      uint4 oaa, oab, oba, obb;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 v00 = bf2_to_f2<F16>(word_of(q00, k)), v01 = bf2_to_f2<F16>(word_of(q01, k)), v02 = bf2_to_f2<F16>(word_of(q02, k));
        const float2 v10 = bf2_to_f2<F16>(word_of(q10, k)), v11 = bf2_to_f2<F16>(word_of(q11, k)), v12 = bf2_to_f2<F16>(word_of(q12, k));
        const float2 v20 = bf2_to_f2<F16>(word_of(q20, k)), v21 = bf2_to_f2<F16>(word_of(q21, k)), v22 = bf2_to_f2<F16>(word_of(q22, k));
        // horizontal blends: column pair a = (xa, xa+1), b = (xb, xb+xpb)
        const float a0x = blend2(lxa0, lxa1, v00.x, v01.x), a0y = blend2(lxa0, lxa1, v00.y, v01.y);
        const float a1x = blend2(lxa0, lxa1, v10.x, v11.x), a1y = blend2(lxa0, lxa1, v10.y, v11.y);
        const float a2x = blend2(lxa0, lxa1, v20.x, v21.x), a2y = blend2(lxa0, lxa1, v20.y, v21.y);
        const float b0x = blend2(lxb0, lxb1, v01.x, v02.x), b0y = blend2(lxb0, lxb1, v01.y, v02.y);
        const float b1x = blend2(lxb0, lxb1, v11.x, v12.x), b1y = blend2(lxb0, lxb1, v11.y, v12.y);
        const float b2x = blend2(lxb0, lxb1, v21.x, v22.x), b2y = blend2(lxb0, lxb1, v21.y, v22.y);
        set_word(oaa, k, pack_act2<F16>(blend2(lya0, lya1, a0x, a1x), blend2(lya0, lya1, a0y, a1y)));
        set_word(oab, k, pack_act2<F16>(blend2(lya0, lya1, b0x, b1x), blend2(lya0, lya1, b0y, b1y)));
        set_word(oba, k, pack_act2<F16>(blend2(lyb0, lyb1, a1x, a2x), blend2(lyb0, lyb1, a1y, a2y)));
        set_word(obb, k, pack_act2<F16>(blend2(lyb0, lyb1, b1x, b2x), blend2(lyb0, lyb1, b1y, b2y)));
      }
      o0[0] = oaa;
      o0[C8] = oab;
      o1[0] = oba;
      o1[C8] = obb;
    } else {
      // first row / column pair (or a degenerate size): every output from its own four neighbours
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int y1 = u ? yb : ya, yp = u ? ypb : ypa;
        const float l1 = u ? lyb1 : lya1, l0 = u ? lyb0 : lya0;
#pragma unroll
        for (int v = 0; v < 2; ++v) {
          const int x1 = v ? xb : xa, xp = v ? xpb : xpa;
          const float lx1 = v ? lxb1 : lxa1, lx0 = v ? lxb0 : lxa0;
          const uint4* p = img + (((size_t)y1 * w + x1) << c_shift);
          const uint4 q00 = __ldg(p), q01 = __ldg(p + ((size_t)xp << c_shift));
          const uint4 q10 = __ldg(p + (((size_t)yp * w) << c_shift)), q11 = __ldg(p + (((size_t)yp * w + xp) << c_shift));
          (u ? o1 : o0)[v ? C8 : 0] = bilerp4<F16>(q00, q01, q10, q11, lx0, lx1, l0, l1);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Gaussian head: deterministic two-stage spatial mean, then the 1x1 conv C -> 2L
// ------------------------------------------------------------------------------------------------
constexpr int HEAD_ROWS_PER_BLOCK = 64;  // pixels reduced per stage-1 block

template <bool F16>
__global__ void __launch_bounds__(256)
mean_partial_kernel(const uint32_t* __restrict__ enc, float* __restrict__ partial, int P, int C2, int nchunk) {
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int p0 = chunk * HEAD_ROWS_PER_BLOCK;
  const int p1 = min(P, p0 + HEAD_ROWS_PER_BLOCK);
  for (int c = threadIdx.x; c < C2; c += blockDim.x) {
    float sx = 0.f, sy = 0.f;
    const uint32_t* src = enc + ((long long)b * P + p0) * C2 + c;
    for (int p = p0; p < p1; ++p, src += C2) {
      const float2 v = unpack_act2<F16>(__ldg(src));
      sx += v.x;
      sy += v.y;
    }
    float* dst = partial + ((long long)b * nchunk + chunk) * (2 * C2) + 2 * c;
    dst[0] = sx;
    dst[1] = sy;
  }
}

__global__ void __launch_bounds__(256)
gauss_head_kernel(const float* __restrict__ partial, const float* __restrict__ w, const float* __restrict__ bias,
                  float* __restrict__ out, int P, int C, int nchunk, int nout) {
  extern __shared__ float mean_s[];  // [C]
  const int b = blockIdx.x;
  const float inv = 1.f / (float)P;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < nchunk; ++k) s += partial[((long long)b * nchunk + k) * C + c];
    mean_s[c] = s * inv;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  for (int o = warp; o < nout; o += nwarp) {
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s = fmaf(mean_s[c], __ldg(w + (long long)o * C + c), s);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    if (lane == 0) out[(long long)b * nout + o] = s + bias[o];
  }
}

__global__ void latent_samples_kernel(const float* __restrict__ mls, const float* __restrict__ eps,
                                      float* __restrict__ z, int S, int B, int L) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= S * B * L) return;
  const int d = i % L, b = (i / L) % B;
  z[i] = mls[b * 2 * L + d] + expf(mls[b * 2 * L + L + d]) * eps[i];
}

__global__ void kl_kernel(const float* __restrict__ q, const float* __restrict__ p, float* __restrict__ kl, int B,
                          int L) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float s = 0.f;
  for (int d = 0; d < L; ++d) {
    // torch.distributions.kl._kl_normal_normal
    const float sq = expf(q[b * 2 * L + L + d]), sp = expf(p[b * 2 * L + L + d]);
    const float r = sq / sp;
    const float var_ratio = r * r;
    const float dm = (q[b * 2 * L + d] - p[b * 2 * L + d]) / sp;
    s += 0.5f * (var_ratio + dm * dm - 1.f - logf(var_ratio));
  }
  kl[b] = s;
}

// ------------------------------------------------------------------------------------------------
// multi-tensor EMA: one 256-thread block per chunk of <= 65536 elements
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ema_kernel(const int64_t* __restrict__ table, float m, float om,
                                                  const long long* __restrict__ iteration, double cap) {
  if (iteration != nullptr) {
    // AdaMT warm-up (adamt_trainer.py:41): min(1 - 1 / (iteration + 1), momentum), evaluated in double like the
    // reference's python expression, then m and (1 - m) rounded to fp32 like ATen's scalar multiplications
    const double md = fmin(1.0 - 1.0 / ((double)iteration[0] + 1.0), cap);
    m = (float)md;
    om = (float)(1.0 - md);
  }
  const int64_t* e = table + 3LL * blockIdx.x;
  float* t = reinterpret_cast<float*>(e[0]);
  const float* s = reinterpret_cast<const float*>(e[1]);
  const int n = (int)e[2];
  if (((reinterpret_cast<uintptr_t>(t) | reinterpret_cast<uintptr_t>(s)) & 15) == 0) {
    const int n4 = n >> 2;
    float4* t4 = reinterpret_cast<float4*>(t);
    const float4* s4 = reinterpret_cast<const float4*>(s);
    for (int i = threadIdx.x; i < n4; i += blockDim.x) {
      float4 a = t4[i];
      const float4 b = __ldg(s4 + i);
      // same rounding as the reference's  t*m + p*(1-m)  (two products, one add; no fma contraction)
      a.x = __fadd_rn(__fmul_rn(a.x, m), __fmul_rn(b.x, om));
      a.y = __fadd_rn(__fmul_rn(a.y, m), __fmul_rn(b.y, om));
      a.z = __fadd_rn(__fmul_rn(a.z, m), __fmul_rn(b.z, om));
      a.w = __fadd_rn(__fmul_rn(a.w, m), __fmul_rn(b.w, om));
      t4[i] = a;
    }
    for (int i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x)
      t[i] = __fadd_rn(__fmul_rn(t[i], m), __fmul_rn(s[i], om));
  } else {
    for (int i = threadIdx.x; i < n; i += blockDim.x) t[i] = __fadd_rn(__fmul_rn(t[i], m), __fmul_rn(s[i], om));
  }
}

// ------------------------------------------------------------------------------------------------
// plain CUDA-core conv3x3 (cross-check): one thread per (pixel, cout), fp32 accumulate over bf16 inputs
// ------------------------------------------------------------------------------------------------
__global__ void conv3x3_simt_kernel(const __nv_bfloat16* __restrict__ s0, int c0, const __nv_bfloat16* __restrict__ s1,
                                    int c1, const __nv_bfloat16* __restrict__ wp, const float* __restrict__ bias,
                                    __nv_bfloat16* __restrict__ out, float* __restrict__ out_f32, int B, int H, int W,
                                    int cout, int relu) {
  const int ctot = c0 + c1;
  const long long total = (long long)B * H * W * cout;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int co = t % cout;
    const long long pix = t / cout;
    const int x = pix % W;
    const int y = (pix / W) % H;
    const int b = pix / ((long long)W * H);
    float acc = 0.f;
    for (int tap = 0; tap < 9; ++tap) {
      const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
      if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
      const long long ip = ((long long)b * H + yy) * W + xx;
      const __nv_bfloat16* wr = wp + ((long long)co * 9 + tap) * ctot;
      for (int c = 0; c < c0; ++c) acc = fmaf(__bfloat162float(s0[ip * c0 + c]), __bfloat162float(wr[c]), acc);
      for (int c = 0; c < c1; ++c) acc = fmaf(__bfloat162float(s1[ip * c1 + c]), __bfloat162float(wr[c0 + c]), acc);
    }
    acc += bias ? bias[co] : 0.f;
    if (relu) acc = fmaxf(acc, 0.f);
    if (out) out[t] = __float2bfloat16(acc);
    if (out_f32) out_f32[t] = acc;
  }
}

__global__ void avgpool2_f32_to_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, int B, int H,
                                            int W, int C) {
  const int Ho = H >> 1, Wo = W >> 1;
  const long long total = (long long)B * Ho * Wo * C;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int c = t % C;
    const long long pix = t / C;
    const int x = pix % Wo;
    const int y = (pix / Wo) % Ho;
    const int b = pix / ((long long)Wo * Ho);
    const long long base = (((long long)b * H + 2 * y) * W + 2 * x) * C + c;
    // same association as the tensor-core epilogue: (a + right) + (below + below-right)
    const float s = (in[base] + in[base + C]) + (in[base + (long long)W * C] + in[base + (long long)W * C + C]);
    out[t] = __float2bfloat16(0.25f * s);
  }
}


// ------------------------------------------------------------------------------------------------
// Tiled prediction (the reference feeds torch_em.util.prediction.predict_with_halo one block at a time,
// punet_predictions.py:41-49): gather a batch of equally sized outer blocks from the image, standardise each block
// with its own mean / population std ((x - mean) / (std + 1e-7), torch_em.transform.raw.standardize), and scatter the
// inner (halo-cropped) part of a prediction batch back into the output image.  rois: int32 [T][4] = (y0, x0, h, w).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
tile_stats_kernel(const float* __restrict__ img, int W, const int* __restrict__ rois, int th, int tw,
                  double* __restrict__ stats) {
  const int t = blockIdx.y;
  const int y0 = rois[4 * t], x0 = rois[4 * t + 1];
  const int n = th * tw;
  double s = 0.0, q = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int y = i / tw, x = i - y * tw;
    const double v = (double)img[(size_t)(y0 + y) * W + x0 + x];
    s += v;
    q += v * v;
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
    atomicAdd(stats + 2 * t + threadIdx.x, a);
  }
}

__global__ void __launch_bounds__(256)
tile_gather_standardize_kernel(const float* __restrict__ img, int W, const int* __restrict__ rois, int th, int tw,
                               const double* __restrict__ stats, float* __restrict__ out) {
  const int t = blockIdx.y;
  const int y0 = rois[4 * t], x0 = rois[4 * t + 1];
  const int n = th * tw;
  const double mean = stats[2 * t] / n;
  double var = stats[2 * t + 1] / n - mean * mean;
  if (var < 0.0) var = 0.0;
  const float m = (float)mean, inv = 1.0f / ((float)sqrt(var) + 1e-7f);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int y = i / tw, x = i - y * tw;
    out[(size_t)t * n + i] = (img[(size_t)(y0 + y) * W + x0 + x] - m) * inv;
  }
}

// inner: int32 [T][4] = (y0, x0, h, w) in IMAGE coordinates; outer origin from rois
__global__ void __launch_bounds__(256)
tile_scatter_kernel(const float* __restrict__ pred, int th, int tw, const int* __restrict__ rois,
                    const int* __restrict__ inner, float* __restrict__ out, int W) {
  const int t = blockIdx.y;
  const int oy = rois[4 * t], ox = rois[4 * t + 1];
  const int iy = inner[4 * t], ix = inner[4 * t + 1], ih = inner[4 * t + 2], iw = inner[4 * t + 3];
  const int n = ih * iw;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int y = i / iw, x = i - y * iw;
    out[(size_t)(iy + y) * W + ix + x] = pred[((size_t)t * th + (iy - oy + y)) * tw + (ix - ox + x)];
  }
}

static inline int grid_for(long long total, int block, int cap = 148 * 16) {
  long long g = (total + block - 1) / block;
  if (g < 1) g = 1;
  if (g > cap) g = cap;
  return (int)g;
}

}  // namespace pda

using namespace pda;

extern "C" {

int pda_pack_conv3x3_weights(const float* w, void* o, int cout, int cin, int rot180, int f16, void* stream) {
  if (!w || !o) return PDA_ERR_ARG;
  if (cout <= 0 || cin <= 0) return PDA_ERR_SHAPE;
  PDA_COUNT(1);
  pack_w_kernel<<<grid_for(9LL * cout * cin, 256), 256, 0, (cudaStream_t)stream>>>(
      w, static_cast<uint16_t*>(o), cout, cin, rot180, f16);
  return cudaGetLastError() == cudaSuccess ? PDA_OK : PDA_ERR_CUDA;
}

int pda_pack_conv3x3_weights_multi(const int64_t* table, int n_chunks, void* stream) {
  if (!table) return PDA_ERR_ARG;
  if (n_chunks <= 0) return PDA_ERR_SHAPE;
  PDA_COUNT(1);
  pack_w_multi_kernel<<<n_chunks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const long long*>(table));
  return cudaGetLastError() == cudaSuccess ? PDA_OK : PDA_ERR_CUDA;
}

// 1: the first conv layer (cout = 64) runs on the tensor core; initial value from PDA_FIRST_TC.  set < 0: query only.
static int first_conv_tc_mode(int set) {
  static std::atomic<int> mode{[] {
    const char* e = getenv("PDA_FIRST_TC");
    return e ? atoi(e) : 1;
  }()};
  if (set >= 0) return mode.exchange(set);
  return mode.load();
}

int pda_set_first_conv_tc(int mode) { return first_conv_tc_mode(mode); }

int pda_conv3x3_first(const float* x0, const float* x1, const float* w, const float* bias, void* out, int B, int H,
                      int W, int cout, int relu, int act_f16, void* stream) {
  if (!x0 || !w || !bias || !out) return PDA_ERR_ARG;
  if (cout <= 0 || (cout & 7) || B <= 0 || H <= 0 || W <= 0) return PDA_ERR_SHAPE;
  const int groups = cout >> 3;
  if (256 % groups || cout > 256) return PDA_ERR_SHAPE;
  const long long quads = (long long)B * H * ((W + 3) / 4);
  if (quads * groups >= 0x7fffffffLL) return PDA_ERR_SHAPE;
  const int cin = x1 ? 2 : 1;
  const size_t smem = sizeof(float) * (cin * 9 * cout + cout);
  PDA_COUNT(1);
  cudaStream_t st = (cudaStream_t)stream;
  // cout = 64 (every script configuration): the tensor-core kernel (tf32 with split operands, csrc/conv_first_tc.cu);
  // PDA_FIRST_TC=0 / pda_set_first_conv_tc(0) keep the CUDA-core kernel below
  if (first_conv_tc_mode(-1) && cout == 64) return conv3x3_first_tc(x0, x1, w, bias, out, B, H, W, relu, act_f16, st);
  const int grid = grid_for(quads * groups, 256, 148 * 6);
  __nv_bfloat16* o = static_cast<__nv_bfloat16*>(out);
  if (x1) {
    if (act_f16) conv_first_kernel<2, true><<<grid, 256, smem, st>>>(x0, x1, w, bias, o, B, H, W, cout, relu);
    else conv_first_kernel<2, false><<<grid, 256, smem, st>>>(x0, x1, w, bias, o, B, H, W, cout, relu);
  } else {
    if (act_f16) conv_first_kernel<1, true><<<grid, 256, smem, st>>>(x0, x1, w, bias, o, B, H, W, cout, relu);
    else conv_first_kernel<1, false><<<grid, 256, smem, st>>>(x0, x1, w, bias, o, B, H, W, cout, relu);
  }
  return cudaGetLastError() == cudaSuccess ? PDA_OK : PDA_ERR_CUDA;
}

int pda_conv3x3_tc(const void* src0, int c0, const void* src1, int c1, const void* w_packed, const float* bias,
                   void* out, void* out_pool, const void* relu_mask, int B, int H, int W, int cout, int relu,
                   int bn_tile, int act_f16, int* range_flag, void* stream) {
  if (!src0 || !w_packed || (!out && !out_pool) || (c1 > 0 && !src1)) return PDA_ERR_ARG;
  return conv3x3_tc(src0, c0, src1, c1, w_packed, bias, out, out_pool, relu_mask, B, H, W, cout, relu, bn_tile,
                    act_f16, range_flag, (cudaStream_t)stream);
}

int pda_conv3x3_up_tc(const void* up_src, int c0, const void* src1, int c1, const void* w_packed, const float* bias,
                      void* out, void* out_pool, int B, int H, int W, int cout, int relu, int act_f16, int* range_flag,
                      void* stream) {
  if (!up_src || !src1 || !w_packed || (!out && !out_pool)) return PDA_ERR_ARG;
  // the fused form exists in the CTA-pair kernel only; at least two pixel tiles
  const int th = (H > 16 && cout % 256 != 0) ? 32 : 16;
  if ((long long)((W + 7) / 8) * ((H + th - 1) / th) * B < 2) return PDA_ERR_SHAPE;
  return conv3x3_tc2(nullptr, c0, src1, c1, w_packed, bias, out, out_pool, nullptr, B, H, W, cout, relu, 0, act_f16,
                     range_flag, (cudaStream_t)stream, up_src);
}

int pda_set_conv_pair(int mode) { return conv_pair_mode(mode); }

int pda_set_sm_budget(int sms) { return sm_budget(sms); }

int pda_conv3x3_bf16_simt(const void* src0, int c0, const void* src1, int c1, const void* w_packed,
                          const float* bias, void* out, void* out_pool, int B, int H, int W, int cout, int relu,
                          void* stream) {
  if (!src0 || !w_packed || (!out && !out_pool) || (c1 > 0 && !src1)) return PDA_ERR_ARG;
  if (c0 <= 0 || c1 < 0 || cout <= 0) return PDA_ERR_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  const long long total = (long long)B * H * W * cout;
  float* tmp = nullptr;
  if (out_pool) {
    if ((H & 1) || (W & 1)) return PDA_ERR_SHAPE;
    if (cudaMallocAsync(&tmp, total * sizeof(float), st) != cudaSuccess) return PDA_ERR_CUDA;
  }
  PDA_COUNT(1);
  conv3x3_simt_kernel<<<grid_for(total, 256, 148 * 32), 256, 0, st>>>(
      static_cast<const __nv_bfloat16*>(src0), c0, static_cast<const __nv_bfloat16*>(src1), c1,
      static_cast<const __nv_bfloat16*>(w_packed), bias, static_cast<__nv_bfloat16*>(out), tmp, B, H, W, cout, relu);
  if (out_pool) {
    PDA_COUNT(1);
  avgpool2_f32_to_bf16_kernel<<<grid_for(total / 4, 256), 256, 0, st>>>(tmp, static_cast<__nv_bfloat16*>(out_pool),
                                                                          B, H, W, cout);
    cudaFreeAsync(tmp, st);
  }
  return cudaGetLastError() == cudaSuccess ? PDA_OK : PDA_ERR_CUDA;
}

int pda_avgpool2(const void* in, void* out, int B, int H, int W, int C, int act_f16, void* stream) {
  if (!in || !out) return PDA_ERR_ARG;
  if ((H & 1) || (W & 1) || B <= 0 || c8_shift(C) < 0) return PDA_ERR_SHAPE;
  const long long total = (long long)B * (H / 2) * (W / 2) * (C / 8);
  if (total >= 0x7fffffffLL) return PDA_ERR_SHAPE;
  PDA_COUNT(1);
  if (act_f16)
    avgpool2_kernel<true><<<row_grid((W / 2) * (C / 8), B * (H / 2)), 256, 0, (cudaStream_t)stream>>>(
        static_cast<const uint4*>(in), static_cast<uint4*>(out), B, H, W, c8_shift(C));
  else
    avgpool2_kernel<false><<<row_grid((W / 2) * (C / 8), B * (H / 2)), 256, 0, (cudaStream_t)stream>>>(
        static_cast<const uint4*>(in), static_cast<uint4*>(out), B, H, W, c8_shift(C));
  return cudaGetLastError() == cudaSuccess ? PDA_OK : PDA_ERR_CUDA;
}

int pda_upsample2x_bilinear(const void* in, void* out, int B, int h, int w, int C, int act_f16, void* stream) {
  if (!in || !out) return PDA_ERR_ARG;
  if (c8_shift(C) < 0 || B <= 0 || h <= 0 || w <= 0) return PDA_ERR_SHAPE;
  const long long total = (long long)B * (2 * h) * (2 * w) * (C / 8);
  if (total >= 0x7fffffffLL) return PDA_ERR_SHAPE;
  PDA_COUNT(1);
  if (B > 65535) return PDA_ERR_SHAPE;
  dim3 grid = row_grid((long long)w * (C / 8), h, (148 * 8 + B - 1) / B);
  grid.z = B;
  if (act_f16)
    upsample2x_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(static_cast<const uint4*>(in),
                                                                   static_cast<uint4*>(out), h, w, c8_shift(C));
  else
    upsample2x_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(static_cast<const uint4*>(in),
                                                                    static_cast<uint4*>(out), h, w, c8_shift(C));
  return cudaGetLastError() == cudaSuccess ? PDA_OK : PDA_ERR_CUDA;
}

int pda_gauss_head_scratch_rows(int P) { return (P + HEAD_ROWS_PER_BLOCK - 1) / HEAD_ROWS_PER_BLOCK; }

int pda_gauss_head(const void* enc, const float* w_head, const float* b_head, float* scratch, float* mu_logsigma,
                   int B, int P, int C, int latent, int act_f16, void* stream) {
  if (!enc || !w_head || !b_head || !scratch || !mu_logsigma) return PDA_ERR_ARG;
  if (B <= 0 || P <= 0 || C <= 0 || (C & 1) || latent <= 0 || C * sizeof(float) > 48 * 1024) return PDA_ERR_SHAPE;
  const int nchunk = pda_gauss_head_scratch_rows(P);
  cudaStream_t st = (cudaStream_t)stream;
  PDA_COUNT(2);
  if (act_f16)
    mean_partial_kernel<true><<<dim3(nchunk, B), 256, 0, st>>>(static_cast<const uint32_t*>(enc), scratch, P, C / 2,
                                                               nchunk);
  else
    mean_partial_kernel<false><<<dim3(nchunk, B), 256, 0, st>>>(static_cast<const uint32_t*>(enc), scratch, P, C / 2,
                                                                nchunk);
  gauss_head_kernel<<<B, 256, C * sizeof(float), st>>>(scratch, w_head, b_head, mu_logsigma, P, C, nchunk,
                                                       2 * latent);
  return cudaGetLastError() == cudaSuccess ? PDA_OK : PDA_ERR_CUDA;
}

int pda_latent_samples(const float* mls, const float* eps, float* z, int S, int B, int latent, void* stream) {
  if (!mls || !eps || !z) return PDA_ERR_ARG;
  const int n = S * B * latent;
  if (n <= 0) return PDA_ERR_SHAPE;
  PDA_COUNT(1);
  latent_samples_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(mls, eps, z, S, B, latent);
  return cudaGetLastError() == cudaSuccess ? PDA_OK : PDA_ERR_CUDA;
}

int pda_kl_diag_gauss(const float* q, const float* p, float* kl, int B, int latent, void* stream) {
  if (!q || !p || !kl) return PDA_ERR_ARG;
  if (B <= 0 || latent <= 0) return PDA_ERR_SHAPE;
  PDA_COUNT(1);
  kl_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(q, p, kl, B, latent);
  return cudaGetLastError() == cudaSuccess ? PDA_OK : PDA_ERR_CUDA;
}

int pda_multi_tensor_ema(const int64_t* table, int n_chunks, double momentum, void* stream) {
  if (!table) return PDA_ERR_ARG;
  if (n_chunks <= 0) return PDA_ERR_SHAPE;
  // the reference multiplies by the python doubles m and (1. - m), each rounded to fp32 by ATen
  PDA_COUNT(1);
  ema_kernel<<<n_chunks, 256, 0, (cudaStream_t)stream>>>(table, (float)momentum, (float)(1.0 - momentum), nullptr, 0.0);
  return cudaGetLastError() == cudaSuccess ? PDA_OK : PDA_ERR_CUDA;
}

__global__ void ema_iteration_inc_kernel(long long* it) { it[0] += 1; }

int pda_multi_tensor_ema_warmup(const int64_t* table, int n_chunks, double momentum, int64_t* iteration_dev,
                                void* stream) {
  if (!table || !iteration_dev) return PDA_ERR_ARG;
  if (n_chunks <= 0) return PDA_ERR_SHAPE;
  PDA_COUNT(2);
  ema_kernel<<<n_chunks, 256, 0, (cudaStream_t)stream>>>(table, 0.f, 0.f,
                                                         reinterpret_cast<const long long*>(iteration_dev), momentum);
  ema_iteration_inc_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(reinterpret_cast<long long*>(iteration_dev));
  return cudaGetLastError() == cudaSuccess ? PDA_OK : PDA_ERR_CUDA;
}

int pda_tile_gather_standardize(const float* image, int H, int W, const int32_t* rois, int T, int th, int tw,
                                double* stats, float* out, void* stream) {
  if (!image || !rois || !stats || !out) return PDA_ERR_ARG;
  if (T <= 0 || th <= 0 || tw <= 0 || th > H || tw > W || T > 65535) return PDA_ERR_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemsetAsync(stats, 0, sizeof(double) * 2 * T, st) != cudaSuccess) return PDA_ERR_CUDA;
  int gx = (th * tw + 256 * 8 - 1) / (256 * 8);
  if (gx > 64) gx = 64;
  PDA_COUNT(2);
  tile_stats_kernel<<<dim3(gx, T), 256, 0, st>>>(image, W, rois, th, tw, stats);
  tile_gather_standardize_kernel<<<dim3(gx, T), 256, 0, st>>>(image, W, rois, th, tw, stats, out);
  return cudaGetLastError() == cudaSuccess ? PDA_OK : PDA_ERR_CUDA;
}

int pda_tile_scatter(const float* pred, int T, int th, int tw, const int32_t* rois, const int32_t* inner, float* out,
                     int H, int W, void* stream) {
  if (!pred || !rois || !inner || !out) return PDA_ERR_ARG;
  if (T <= 0 || th <= 0 || tw <= 0 || T > 65535 || H <= 0 || W <= 0) return PDA_ERR_SHAPE;
  int gx = (th * tw + 256 * 8 - 1) / (256 * 8);
  if (gx > 64) gx = 64;
  PDA_COUNT(1);
  tile_scatter_kernel<<<dim3(gx, T), 256, 0, (cudaStream_t)stream>>>(pred, th, tw, rois, inner, out, W);
  return cudaGetLastError() == cudaSuccess ? PDA_OK : PDA_ERR_CUDA;
}

int pda_abi_version(void) { return PDA_ABI_VERSION; }

long long pda_launch_count(void) { return g_pda_launches.load(); }
void pda_reset_launch_count(void) { g_pda_launches.store(0); }

const char* pda_error_string(int code) {
  switch (code) {
    case PDA_OK: return "ok";
    case PDA_ERR_SHAPE: return "unsupported or inconsistent shape";
    case PDA_ERR_CUDA: return "CUDA runtime error";
    case PDA_ERR_DRIVER: return "cuTensorMapEncodeTiled driver entry point unavailable";
    case PDA_ERR_TENSORMAP: return "TMA tensor map encoding failed";
    case PDA_ERR_ARG: return "invalid argument";
    default: return "unknown error";
  }
}

}  // extern "C"
