// FP64 tensor-core (DMMA) throughput on B200 and its overlap with vector DFMA (development microbenchmark).
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma16816(double (&d)[4], const double (&a)[8], const double (&b)[4]) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
      : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
      : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

// MODE 0: DMMA m8n8k4 only, NM independent accumulator tiles; MODE 1: DFMA only (NF chains); MODE 2: both interleaved
template <int MODE, int NM, int NF>
__global__ void __launch_bounds__(256) k884(double* out, double a, double b, int iters) {
  double acc[NM][2];
  double f[NF];
#pragma unroll
  for (int i = 0; i < NM; ++i) acc[i][0] = acc[i][1] = threadIdx.x * 1e-3 + i;
#pragma unroll
  for (int i = 0; i < NF; ++i) f[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
    if (MODE != 1) {
#pragma unroll
      for (int i = 0; i < NM; ++i) dmma884(acc[i][0], acc[i][1], a, b);
    }
    if (MODE != 0) {
#pragma unroll
      for (int i = 0; i < NF; ++i) f[i] = fma(f[i], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NM; ++i) s += acc[i][0] + acc[i][1];
#pragma unroll
  for (int i = 0; i < NF; ++i) s += f[i];
  if (s == 123.456) out[0] = s;
}

template <int NM>
__global__ void __launch_bounds__(256) k16816(double* out, double a, double b, int iters) {
  double acc[NM][4];
  double av[8], bv[4];
#pragma unroll
  for (int i = 0; i < 8; ++i) av[i] = a + i;
#pragma unroll
  for (int i = 0; i < 4; ++i) bv[i] = b + i;
#pragma unroll
  for (int i = 0; i < NM; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NM; ++i) dmma16816(acc[i], av, bv);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NM; ++i) s += acc[i][0] + acc[i][1] + acc[i][2] + acc[i][3];
  if (s == 123.456) out[0] = s;
}

template <class F>
float timeit(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9;
  for (int r = 0; r < 4; ++r) { cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms; }
  return best;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0); int nsm = p.multiProcessorCount;
  double* d; cudaMalloc(&d, 8);
  const int iters = 1 << 13;
  for (int cps : {1, 2, 4}) {  // resident CTAs of 256 threads per SM (8, 16, 32 warps)
    int blocks = nsm * cps;
    // occupancy is not forced; with small kernels all blocks are resident anyway (<= 8 CTAs/SM)
    float t0 = timeit([&] { k884<0, 8, 1><<<blocks, 256>>>(d, 0.999, 1e-9, iters); });
    double mma_flop = 2.0 * 8 * 8 * 4 * 8 * (double)iters * blocks * 8;  // per warp-instr 256 FMA; 8 tiles; 8 warps
    printf("CTAs/SM=%d  DMMA m8n8k4 only      : %.2f TFLOP/s\n", cps, mma_flop / (t0 * 1e-3) / 1e12);
    float t1 = timeit([&] { k884<1, 1, 16><<<blocks, 256>>>(d, 0.999, 1e-9, iters); });
    double fma_flop = 2.0 * 16 * (double)iters * blocks * 256;
    printf("CTAs/SM=%d  DFMA only (16 chains) : %.2f TFLOP/s\n", cps, fma_flop / (t1 * 1e-3) / 1e12);
    float t2 = timeit([&] { k884<2, 8, 16><<<blocks, 256>>>(d, 0.999, 1e-9, iters); });
    printf("CTAs/SM=%d  DMMA(8)+DFMA(16) mixed: %.2f TFLOP/s total (%.2f ms vs %.2f + %.2f)\n", cps, (mma_flop + fma_flop) / (t2 * 1e-3) / 1e12, t2, t0, t1);
    float t3 = timeit([&] { k884<2, 8, 2><<<blocks, 256>>>(d, 0.999, 1e-9, iters); });
    double fma2 = 2.0 * 2 * (double)iters * blocks * 256;
    printf("CTAs/SM=%d  DMMA(8)+DFMA(2) mixed : %.2f TFLOP/s total (%.2f ms vs DMMA-only %.2f)\n", cps, (mma_flop + fma2) / (t3 * 1e-3) / 1e12, t3, t0);
    float t4 = timeit([&] { k16816<4><<<blocks, 256>>>(d, 0.999, 1e-9, iters); });
    double mma2 = 2.0 * 16 * 8 * 16 * 4 * (double)iters * blocks * 8;
    printf("CTAs/SM=%d  DMMA m16n8k16 only    : %.2f TFLOP/s\n", cps, mma2 / (t4 * 1e-3) / 1e12);
  }
  return 0;
}
