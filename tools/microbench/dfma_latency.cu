// Dependent-issue latency of DFMA (and of a few other instructions of the INT8 Legendre kernel's chain)
// on one warp, and how it changes when more warps of the SM issue DFMA at the same time.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench/dfma_latency tools/microbench/dfma_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS>
__global__ void chain(double* out, long long* cyc, int n, double a, double b) {
  double x[CHAINS], y[CHAINS];
  for (int c = 0; c < CHAINS; ++c) {
    x[c] = 1.0 + threadIdx.x * 1e-3 + c;
    y[c] = 0.5 + c;
  }
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) {
      const double t = fma(a, x[c], -y[c]);  // the recurrence's shape: p_new = r p - p_old
      y[c] = x[c];
      x[c] = t;
    }
  }
  const long long t1 = clock64();
  double s = 0;
  for (int c = 0; c < CHAINS; ++c) s += x[c] + y[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + b;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

int main() {
  double* out;
  long long* cyc;
  cudaMalloc(&out, 1 << 24);
  cudaMalloc(&cyc, 8);
  const int n = 4096;
  for (int blocks : {1, 148, 296}) {
    for (int threads : {32, 64, 128, 256, 512}) {
      long long h[3];
      chain<1><<<blocks, threads>>>(out, cyc, n, 0.999, 0.0);
      cudaMemcpy(&h[0], cyc, 8, cudaMemcpyDeviceToHost);
      chain<2><<<blocks, threads>>>(out, cyc, n, 0.999, 0.0);
      cudaMemcpy(&h[1], cyc, 8, cudaMemcpyDeviceToHost);
      chain<4><<<blocks, threads>>>(out, cyc, n, 0.999, 0.0);
      cudaMemcpy(&h[2], cyc, 8, cudaMemcpyDeviceToHost);
      printf("blocks %3d threads %3d: cycles per dependent DFMA step: 1 chain %.1f, 2 chains %.1f, 4 chains %.1f\n", blocks, threads,
             (double)h[0] / n, (double)h[1] / n, (double)h[2] / n);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
