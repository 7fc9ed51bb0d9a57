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

int main() {
  run<0>("ex2.approx.ftz.f32");
  run<1>("sqrt.approx.ftz.f32");
  run<2>("rsqrt.approx.ftz.f32");
  run<3>("lg2.approx.ftz.f32");
  run<4>("rcp.approx.ftz.f32");
  run<5>("sin.approx.ftz.f32");
  run<6>("tanh.approx.f32");
  return cudaDeviceSynchronize() == cudaSuccess ? 0 : 1;
}
