// DFMA throughput vs resident warps per SM and ILP (development microbenchmark).
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double* out, double a, double b, int iters) {
  extern __shared__ double sm[];
  double v[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) v[i] = (threadIdx.x + i) * 1e-3;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) v[i] = fma(v[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += v[i];
  if (s == 123.456) out[0] = s + sm[0];
}
template <int ILP>
void run(int threads, int ctas_per_sm, int nsm) {
  double* d; cudaMalloc(&d, 8);
  // force occupancy with dynamic smem
  int smem = 200 * 1024 / ctas_per_sm;
  cudaFuncSetAttribute(k<ILP>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  int blocks = nsm * ctas_per_sm * 4; int iters = 1 << 14;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9;
  for (int r = 0; r < 4; ++r) {
    cudaEventRecord(e0); k<ILP><<<blocks, threads, smem>>>(d, 0.999999, 1e-9, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms;
  }
  double tf = 2.0 * blocks * threads * (double)ILP * iters / (best * 1e-3) / 1e12;
  printf("ILP=%2d threads=%4d ctas/SM=%d -> warps/SM=%3d : %.2f TFLOP/s\n", ILP, threads, ctas_per_sm, threads / 32 * ctas_per_sm, tf);
  cudaFree(d);
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0); int nsm = p.multiProcessorCount;
  for (int w : {128, 256, 512, 1024}) { run<8>(w, 1, nsm); run<32>(w, 1, nsm); }
  run<8>(512, 2, nsm); run<32>(512, 2, nsm); run<8>(256, 8, nsm); run<4>(512, 1, nsm); run<2>(512, 1, nsm); run<16>(512,1,nsm);
  return 0;
}
