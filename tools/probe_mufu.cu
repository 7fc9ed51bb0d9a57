// MUFU throughput per SM on sm_100a: which special-function ops are full rate (16 lanes / clk / SM) and which are not.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/probe_mufu tools/probe_mufu.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__device__ __forceinline__ float op(float x) {
  float y;
  if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  if (OP == 1) asm volatile("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  if (OP == 2) asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  if (OP == 3) asm volatile("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  if (OP == 4) asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  if (OP == 5) asm volatile("sin.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  if (OP == 6) asm volatile("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int OP>
__global__ void k(float* out, long long* cycles, int iters) {
  float v[8];
  for (int i = 0; i < 8; ++i) v[i] = 1.0f + threadIdx.x * 1e-3f + i;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = op<OP>(v[i]);
  }
  const long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 8; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name) {
  const int threads = 1024, blocks = 148, iters = 4096;
  float* out;
  long long* cyc;
  cudaMalloc(&out, blocks * threads * sizeof(float));
  cudaMalloc(&cyc, blocks * sizeof(long long));
  k<OP><<<blocks, threads>>>(out, cyc, 16);
  k<OP><<<blocks, threads>>>(out, cyc, iters);
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = 0;
  for (int i = 0; i < blocks; ++i) c += h[i];
  c /= blocks;
  const double ops = double(threads) * iters * 8;  // per SM (one block per SM)
  printf("%-22s %.2f lanes/clk/SM  (%.1f clk per warp instruction per SM sub-partition)\n", name, ops / c,
         32.0 / (ops / c / 4));
  cudaFree(out);
  cudaFree(cyc);
}

// The class-LSE epilogue's instruction mix per score on register data (no TMEM, no shared memory): FADD, FFMA,
// MUFU.SQRT |x|, FFMA, MUFU.EX2, FADD.  Two MUFU ops per score bound it at 8 scores / clk / SM; what does the mix reach?
template <bool SMEM_CADD>
__global__ void epi_mix(float* out, long long* cycles, int iters) {
  __shared__ __align__(16) float cadd_s[256];
  if (threadIdx.x < 256) cadd_s[threadIdx.x] = 100.0f + threadIdx.x;
  float a[32], c[32];
  for (int i = 0; i < 32; ++i) {
    a[i] = 1.0f + threadIdx.x * 1e-3f + i;
    c[i] = 100.0f + i;
  }
  const float qn = 50.0f + threadIdx.x, m = -3.0f;
  float part[4] = {0.f, 0.f, 0.f, 0.f};
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const float4* c4 = reinterpret_cast<const float4*>(cadd_s + (it & 7) * 32);
#pragma unroll
    for (int i4 = 0; i4 < 8; ++i4) {
      float cc[4];
      if (SMEM_CADD) {
        const float4 v = c4[i4];
        cc[0] = v.x; cc[1] = v.y; cc[2] = v.z; cc[3] = v.w;
      } else {
        cc[0] = c[i4 * 4]; cc[1] = c[i4 * 4 + 1]; cc[2] = c[i4 * 4 + 2]; cc[3] = c[i4 * 4 + 3];
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int i = i4 * 4 + k;
        float d2 = fmaf(-2.0f, a[i], qn + cc[k]), dist, e;
        asm volatile("sqrt.approx.ftz.f32 %0, %1;" : "=f"(dist) : "f"(fabsf(d2)));
        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fmaf(dist, -1.4426950f, -m)));
        part[k] += e;
        a[i] = dist;  // next iteration depends on this one's result only through a[i]
      }
    }
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = (part[0] + part[1]) + (part[2] + part[3]);
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <bool SMEM_CADD>
void run_mix(const char* name, int threads) {
  const int blocks = 148, iters = 2048;
  float* out;
  long long* cyc;
  cudaMalloc(&out, blocks * threads * sizeof(float));
  cudaMalloc(&cyc, blocks * sizeof(long long));
  epi_mix<SMEM_CADD><<<blocks, threads>>>(out, cyc, 16);
  epi_mix<SMEM_CADD><<<blocks, threads>>>(out, cyc, iters);
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = 0;
  for (int i = 0; i < blocks; ++i) c += h[i];
  c /= blocks;
  const double scores = double(threads) * iters * 32;
  printf("%-46s %2d warps/SM: %.2f scores/clk/SM = %.0f %% of the 2-MUFU bound (8)\n", name, threads / 32, scores / c,
         scores / c / 8 * 100);
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  run_mix<false>("epilogue mix, column terms in registers", 128);
  run_mix<false>("epilogue mix, column terms in registers", 256);
  run_mix<false>("epilogue mix, column terms in registers", 512);
  run_mix<true>("epilogue mix, column terms from shared (LDS.128)", 256);
  run_mix<true>("epilogue mix, column terms from shared (LDS.128)", 512);
  run<0>("ex2.approx.ftz.f32");
  run<1>("sqrt.approx.ftz.f32");
  run<2>("rsqrt.approx.ftz.f32");
  run<3>("lg2.approx.ftz.f32");
  run<4>("rcp.approx.ftz.f32");
  run<5>("sin.approx.ftz.f32");
  run<6>("tanh.approx.f32");
  return cudaDeviceSynchronize() == cudaSuccess ? 0 : 1;
}
