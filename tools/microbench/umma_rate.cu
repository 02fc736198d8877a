// Micro-benchmark (evidence tool, not product code): issue rate of tcgen05.mma (kind::f16, bf16 operands, M = 128) as a
// function of the N extent and of the operand major-ness, with both operands resident in shared memory -- no TMA, no
// epilogue.  One CTA per SM; one thread issues `iters` groups of `K_STEPS x ACCS` MMAs (K = 16 each) that walk through a
// stage buffer the way the conv / wgrad kernels do, then commits and waits.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench/umma_rate.bin tools/microbench/umma_rate.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../probabilistic_domain_adaptation_b200/csrc/ptx.cuh"

using namespace pda;

// always-accumulate MMA without the predicate set-up of the production wrapper
__device__ __forceinline__ void umma_acc(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.eq.u32 p, 1, 1;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc)
      : "memory");
}

// same with the A operand in tensor memory (K = 16 bf16 = 8 columns per MMA), as the fused Fcomb kernel issues it
__device__ __forceinline__ void umma_acc_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.eq.u32 p, 1, 1;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc)
      : "memory");
}

// Small-N / TMEM-A rates: `iters` groups of 4 K-steps on ONE accumulator (the Fcomb pattern: 64-deep K, one D).
template <int N, bool A_TM>
__global__ void __launch_bounds__(128, 1) umma_small_kernel(int iters, unsigned long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const uint32_t sbase = smem_u32(smem);
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem)[i] = 0x3F803F80u ^ ((i * 2654435761u) & 0x007F007Fu);
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    fence_mbar_init();
  }
  if (threadIdx.x < 32) {
    tmem_alloc(smem_u32(&slot), 512);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  // some finite data in the TMEM A region (columns 256..287)
  tmem_st32_fill(tmem + 256 + (static_cast<uint32_t>((threadIdx.x >> 5) * 32) << 16), 0x3C003C00u);
  tmem_st_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x < 32) {
    const bool leader = elect_one();
    constexpr uint32_t idesc = umma_idesc_bf16(128, N, 0, 0);
    const uint64_t da = umma_desc_k_sw128(sbase);
    const uint64_t db = umma_desc_k_sw128(sbase + 64 * 1024);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (leader) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (A_TM) umma_acc_ts(tmem, tmem + 256 + 8 * k, db + 2ull * k, idesc);
          else umma_acc(tmem, da + 2ull * k, db + 2ull * k, idesc);
        }
      }
      __syncwarp();
    }
    if (leader) umma_commit(smem_u32(&bar));
    __syncwarp();
    mbar_wait(smem_u32(&bar), 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0 && leader) cycles[0] = (unsigned long long)(t1 - t0);
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

template <int ACCS, int N, int A_MN, int B_MN, int M = 128>
__global__ void __launch_bounds__(128, 1) umma_rate_kernel(int iters, unsigned long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const uint32_t sbase = smem_u32(smem);
  // non-trivial operand data (zeros would lower the power draw and flatter the clocks)
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem)[i] = 0x3F803F80u ^ ((i * 2654435761u) & 0x007F007Fu);
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    fence_mbar_init();
  }
  if (threadIdx.x < 32) {
    tmem_alloc(smem_u32(&slot), 512);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (threadIdx.x < 32) {
    // warp-uniform control flow, one elected lane issues (as in the production kernels): descriptors stay in uniform
    // registers, the K walk and the accumulator offsets are compile-time immediates
    const bool leader = elect_one();
    constexpr uint32_t idesc = umma_idesc_bf16(M, N, A_MN, B_MN);
    // A tile at offset 0, B tile at 64 KB
    const uint64_t da = A_MN ? umma_desc_mn_sw128(sbase, 1024, 1024) : umma_desc_k_sw128(sbase);
    const uint64_t db = B_MN ? umma_desc_mn_sw128(sbase + 64 * 1024, 1024, 1024) : umma_desc_k_sw128(sbase + 64 * 1024);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (leader) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          // K-major: +32 B per K = 16 step inside the 128-B row (2 units of 16 B); MN-major: +2048 B (16 K rows)
          const uint64_t ka = A_MN ? 128ull * k : 2ull * (k & 3) + 512ull * (k >> 2);
          const uint64_t kb = B_MN ? 128ull * k : 2ull * (k & 3) + 1024ull * (k >> 2);
#pragma unroll
          for (int a = 0; a < ACCS; ++a) umma_acc(tmem + a * N, da + ka + 64ull * a, db + kb, idesc);
        }
      }
      __syncwarp();
    }
    if (leader) umma_commit(smem_u32(&bar));
    __syncwarp();
    mbar_wait(smem_u32(&bar), 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0 && leader) cycles[0] = (unsigned long long)(t1 - t0);
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

struct Row {
  const char* name;
  int n, accs, iters;
  void (*fn)(int, unsigned long long*);
  int m = 128;
};

// Layout probe: which TMEM lanes receive the rows of an M = 64 accumulator (cta_group::1)?  A (K-major, 64 rows) holds
// row index + 1 in k = 0, B (K-major, 16 rows) holds 1 in k = 0, so D[r][n] = r + 1; every lane then reports column 0.
__global__ void __launch_bounds__(128, 1) m64_layout_kernel(float* out) {
  __shared__ __align__(1024) uint8_t sm[16 * 1024 + 2048];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  uint8_t* a = sm;
  uint8_t* b = sm + 16 * 1024;
  for (int i = threadIdx.x; i < (16 * 1024 + 2048) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = 0;
  __syncthreads();
  if (threadIdx.x < 64) {  // row r: element k = 0 lives in 16-byte chunk (0 ^ (r & 7)) of its 128-byte row
    const int r = threadIdx.x;
    __nv_bfloat16 v = __float2bfloat16((float)(r + 1));
    *reinterpret_cast<__nv_bfloat16*>(a + r * 128 + ((0 ^ (r & 7)) << 4)) = v;
  }
  if (threadIdx.x < 16) {
    const int r = threadIdx.x;
    *reinterpret_cast<__nv_bfloat16*>(b + r * 128 + ((0 ^ (r & 7)) << 4)) = __float2bfloat16(1.0f);
  }
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    fence_mbar_init();
  }
  if (threadIdx.x < 32) {
    tmem_alloc(smem_u32(&slot), 32);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  // poison all 128 lanes first
  tmem_st32_fill(tmem + (static_cast<uint32_t>((threadIdx.x >> 5) * 32) << 16), __float_as_uint(-777.f));
  tmem_st_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) {
    umma_bf16(tmem, umma_desc_k_sw128(smem_u32(a)), umma_desc_k_sw128(smem_u32(b)), umma_idesc_bf16(64, 16), 0u);
    umma_commit(smem_u32(&bar));
  }
  __syncthreads();
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after();
  uint32_t v[16];
  tmem_ld16(tmem + (static_cast<uint32_t>((threadIdx.x >> 5) * 32) << 16), v);
  tmem_ld_wait();
  out[threadIdx.x] = __uint_as_float(v[0]);
  out[128 + threadIdx.x] = __uint_as_float(v[5]);
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 32);
}

int main() {
  const int smem = 200 * 1024;
  unsigned long long* d_cyc;
  cudaMalloc(&d_cyc, 8);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
#define ROW(name, ACCS, N, AMN, BMN, iters) {name, N, ACCS, iters, umma_rate_kernel<ACCS, N, AMN, BMN>}
#define ROW64(name, ACCS, N, AMN, BMN, iters) {name, N, ACCS, iters, umma_rate_kernel<ACCS, N, AMN, BMN, 64>, 64}
  std::vector<Row> rows = {
      ROW("K-major A,B   N=64  x1 acc", 1, 64, 0, 0, 400),
      ROW("K-major A,B   N=64  x4 acc (conv, cout 64)", 4, 64, 0, 0, 400),
      ROW("K-major A,B   N=128 x1 acc", 1, 128, 0, 0, 400),
      ROW("K-major A,B   N=128 x2 acc", 2, 128, 0, 0, 400),
      ROW("K-major A,B   N=256 x1 acc", 1, 256, 0, 0, 200),
      ROW("K-major A,B   N=256 x2 acc (conv, cout>=256)", 2, 256, 0, 0, 200),
      ROW("MN-major A,B  N=64  x5 acc (wgrad today)", 5, 64, 1, 1, 400),
      ROW("MN-major A,B  N=64  x6 acc (wgrad today + bias)", 6, 64, 1, 1, 400),
      ROW("MN-major A,B  N=128 x1 acc", 1, 128, 1, 1, 400),
      ROW("MN-major A,B  N=128 x3 acc", 3, 128, 1, 1, 300),
      ROW("MN-major A,B  N=192 x1 acc", 1, 192, 1, 1, 300),
      ROW("MN-major A,B  N=192 x2 acc", 2, 192, 1, 1, 300),
      ROW("MN-major A,B  N=256 x2 acc", 2, 256, 1, 1, 200),
      ROW("K A, MN B     N=192 x2 acc", 2, 192, 0, 1, 300),
      ROW("MN A, K B     N=192 x2 acc", 2, 192, 1, 0, 300),
      ROW64("M = 64: MN-major A,B  N=192 x2 acc", 2, 192, 1, 1, 300),
      ROW64("M = 64: MN-major A,B  N=64  x4 acc", 4, 64, 1, 1, 400),
      ROW64("M = 64: K-major A,B   N=256 x2 acc", 2, 256, 0, 0, 200),
  };
  printf("| operands / tile | MMAs per CTA | ms | TFLOP/s (all %d SMs) | cycles per MMA | smem operand bytes per clk per SM |\n", sms);
  printf("|---|---|---|---|---|---|\n");
  for (auto& r : rows) {
    cudaFuncSetAttribute(r.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    void* args[] = {&r.iters, &d_cyc};
    for (int w = 0; w < 2; ++w) cudaLaunchKernel((const void*)r.fn, dim3(sms), dim3(128), args, smem, 0);
    cudaEventRecord(e0);
    const int reps = 5;
    for (int w = 0; w < reps; ++w) cudaLaunchKernel((const void*)r.fn, dim3(sms), dim3(128), args, smem, 0);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) {
      printf("| %s | launch failed: %s |\n", r.name, cudaGetErrorString(err));
      return 1;
    }
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= reps;
    unsigned long long cyc = 0;
    cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost);
    const double mmas = (double)r.iters * 8 * r.accs;
    const double flop = mmas * 2.0 * r.m * r.n * 16 * sms;
    const double bytes_per_mma = r.m * 16 * 2 + r.n * 16 * 2;
    printf("| %s | %.0f | %.3f | %.0f | %.1f | %.0f |\n", r.name, mmas, ms, flop / (ms * 1e-3) / 1e12, (double)cyc / mmas,
           bytes_per_mma * mmas / (double)cyc);
  }
  // small N and the A operand in tensor memory (the fused Fcomb kernel's MMAs)
  {
    struct SRow { const char* name; int n; void (*fn)(int, unsigned long long*); };
    std::vector<SRow> srows = {
        {"N=64  A in shared memory", 64, umma_small_kernel<64, false>}, {"N=64  A in tensor memory", 64, umma_small_kernel<64, true>},
        {"N=32  A in shared memory", 32, umma_small_kernel<32, false>}, {"N=32  A in tensor memory", 32, umma_small_kernel<32, true>},
        {"N=16  A in shared memory", 16, umma_small_kernel<16, false>}, {"N=16  A in tensor memory", 16, umma_small_kernel<16, true>},
        {"N=128 A in tensor memory", 128, umma_small_kernel<128, true>}, {"N=256 A in tensor memory", 256, umma_small_kernel<256, true>},
    };
    printf("\n| M = 128, K-major B, one accumulator, 4 K-steps per group | MMAs per CTA | cycles per MMA |\n|---|---|---|\n");
    for (auto& r : srows) {
      cudaFuncSetAttribute(r.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      int iters = 2000;
      void* args[] = {&iters, &d_cyc};
      for (int w = 0; w < 3; ++w) cudaLaunchKernel((const void*)r.fn, dim3(sms), dim3(128), args, smem, 0);
      cudaError_t err = cudaDeviceSynchronize();
      if (err != cudaSuccess) {
        printf("| %s | launch failed: %s |\n", r.name, cudaGetErrorString(err));
        return 1;
      }
      unsigned long long cyc = 0;
      cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost);
      printf("| %s | %d | %.1f |\n", r.name, iters * 4, (double)cyc / (iters * 4.0));
    }
  }
  // M = 64 accumulator layout
  float* d_out;
  cudaMalloc(&d_out, 256 * sizeof(float));
  m64_layout_kernel<<<1, 128>>>(d_out);
  float h[256];
  if (cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost) != cudaSuccess) {
    printf("layout probe failed: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 1;
  }
  printf("\nM = 64 accumulator (cta_group::1): value r + 1 of row r as seen by TMEM lane (column 0 | column 5); -777 = untouched\n");
  for (int l = 0; l < 128; l += 16) {
    printf("lanes %3d..%3d:", l, l + 15);
    for (int j = 0; j < 16; ++j) printf(" %4.0f", h[l + j]);
    printf("   |");
    for (int j = 0; j < 16; ++j) printf(" %4.0f", h[128 + l + j]);
    printf("\n");
  }
  return 0;
}
