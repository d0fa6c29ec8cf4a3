// Probe: MUFU throughput per SM (tanh.approx.f32 vs ex2.approx.f32 vs rcp.approx.f32), 512 threads per SM, 8 independent
// chains per thread.   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o mufu_probe mufu_probe.cu && ./mufu_probe
#include <cuda_runtime.h>
#include <cstdio>

template <int OP>
__global__ void k(float* out, int iters, long long* cyc) {
  float v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = 0.001f * (threadIdx.x + i);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(v[i]));
      if (OP == 1) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
      if (OP == 2) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
      if (OP == 3) asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(*reinterpret_cast<unsigned*>(&v[i])));
    }
  }
  __syncthreads();
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 4096;
  const char* names[4] = {"tanh.approx.f32", "ex2.approx.f32", "rcp.approx.f32", "tanh.approx.f16x2 (2 results per op)"};
  for (int op = 0; op < 4; ++op) {
    for (int rep = 0; rep < 2; ++rep) {
      if (op == 0) k<0><<<148, 512>>>(out, iters, cyc);
      if (op == 1) k<1><<<148, 512>>>(out, iters, cyc);
      if (op == 2) k<2><<<148, 512>>>(out, iters, cyc);
      if (op == 3) k<3><<<148, 512>>>(out, iters, cyc);
      cudaDeviceSynchronize();
    }
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double ops = 512.0 * 8 * iters;
    printf("%-40s %.2f thread-ops / clk / SM  (%lld cycles)\n", names[op], ops / (double)h[0], h[0]);
  }
  return 0;
}
