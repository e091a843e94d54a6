// Microbenchmark: DMMA.8x8x4 issue rate per SM sub-partition on sm_100a as a function of warps per
// sub-partition and accumulator-tile shape (register-resident, no memory traffic).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/dmma_rate tools/micro/dmma_rate.cu && build/dmma_rate
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int MI, int NJ>
__global__ void __launch_bounds__(1024, 1) rate_kernel(double* out, int iters, long long* cyc) {
  double acc[MI][NJ][2];
  double af[MI], bf[NJ];
#pragma unroll
  for (int i = 0; i < MI; ++i) { af[i] = 1.0 + threadIdx.x * 1e-3 + i; }
#pragma unroll
  for (int j = 0; j < NJ; ++j) { bf[j] = 0.5 + threadIdx.x * 1e-3 + j; }
#pragma unroll
  for (int i = 0; i < MI; ++i)
#pragma unroll
    for (int j = 0; j < NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
      for (int j = 0; j < NJ; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
  }
  __syncthreads();
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < MI; ++i)
#pragma unroll
    for (int j = 0; j < NJ; ++j) s += acc[i][j][0] + acc[i][j][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MI, int NJ>
void run(int warps_per_block, int iters) {
  double* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * sizeof(double));
  cudaMalloc(&cyc, 148 * sizeof(long long));
  rate_kernel<MI, NJ><<<148, warps_per_block * 32>>>(out, iters, cyc);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  rate_kernel<MI, NJ><<<148, warps_per_block * 32>>>(out, iters, cyc);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  long long mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
  double dmma_per_smsp = (double)iters * MI * NJ * warps_per_block / 4.0;
  double flops = 2.0 * 256 * (double)iters * MI * NJ * warps_per_block * 148;
  printf("tile %dx%d warps/SM %2d (per SMSP %d): %.2f cycles/DMMA/SMSP, %.2f TFLOP/s (%.3f ms)%s\n", MI, NJ,
         warps_per_block, warps_per_block / 4, mx / dmma_per_smsp, flops / ms / 1e9, ms,
         cudaGetLastError() == cudaSuccess ? "" : " ERROR");
  cudaFree(out); cudaFree(cyc);
}

int main() {
  const int iters = 4000;
  run<8, 4>(4, iters); run<8, 4>(8, iters);
  run<4, 4>(4, iters); run<4, 4>(8, iters); run<4, 4>(16, iters);
  run<4, 2>(8, iters); run<4, 2>(16, iters); run<4, 2>(32, iters);
  run<2, 2>(16, iters); run<2, 2>(32, iters);
  run<8, 2>(8, iters); run<8, 2>(16, iters);
  run<4, 8>(4, iters); run<4, 8>(8, iters);
  return 0;
}
