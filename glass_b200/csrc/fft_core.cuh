// fft_core.cuh -- power-of-two complex FP64 FFTs in shared memory for the ring-FFT stage
// (sht_ringfft.cu).  Written as per-thread pass functions without barriers so that the same
// code runs on the device (the drivers at the bottom add __syncthreads) and, thread by thread,
// in the host unit test tests/native/fft_core_host.cpp.
//
// Design points (B200: 32 banks x 4 B, 128-bit shared accesses served per quarter-warp):
//  * register-blocked passes: 3 butterfly levels per radix-8 pass, and the last four levels
//    (half-spans 8,4,2,1) as one radix-16 pass on 16 CONTIGUOUS elements per thread, so an
//    8192-point transform is 4 passes over shared memory instead of 13;
//  * XOR swizzle of the low three index bits with bits 4-6 and with the top three bits,
//        phys(i) = i ^ (((i >> 4) ^ (i >> (n-3))) & 7),
//    makes every pass conflict-free: the strided radix-8 passes (8 consecutive lanes touch 8
//    consecutive elements), the contiguous radix-16 pass (lane stride 16 elements) and the
//    bit-reversing variant of it (lane stride M/8);
//  * DIF: natural in -> bit-reversed out; DIT: bit-reversed in -> natural out, so the Bluestein
//    convolution DIF -> (x chirp spectrum) -> DIT needs no reordering, and its three innermost
//    steps (DIF tail, multiply, DIT head) act on the same 16 elements and are ONE pass;
//  * the "reordering" DIF tail lets thread t process block bitrev(t); its 16 results are then
//    the elements t, t + M/16, ... of the natural-order output: consecutive lanes hold
//    consecutive output elements and store them straight to global memory, coalesced.
// sign: inverse == false -> e^{-2 pi i jk/M}, inverse == true -> e^{+2 pi i jk/M}; unnormalised.
#pragma once

#if defined(__CUDACC__)
#include "common.cuh"
#define GLB_FFT_HD __host__ __device__ __forceinline__
#else
#include <cmath>
#include <cstdint>
struct double2 {
  double x, y;
};
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
#define GLB_FFT_HD inline
namespace glb {
inline double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
inline double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
inline double2 cmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
}  // namespace glb
#endif

namespace glb {
namespace fft {

GLB_FFT_HD int ilog2(int v) {  // v a power of two
#if defined(__CUDA_ARCH__)
  return 31 - __clz(v);
#else
  int n = 0;
  while ((1 << n) < v) ++n;
  return n;
#endif
}

// bit reversal of the low `bits` bits
GLB_FFT_HD unsigned bitrev(unsigned v, int bits) {
#if defined(__CUDA_ARCH__)
  return bits ? (__brev(v) >> (32 - bits)) : 0u;
#else
  unsigned r = 0;
  for (int b = 0; b < bits; ++b) r |= ((v >> b) & 1u) << (bits - 1 - b);
  return r;
#endif
}

// physical position of logical element i in a buffer holding a 2^n-point transform
GLB_FFT_HD int sw(int i, int n) { return i ^ ((((i >> 4) ^ (i >> (n >= 10 ? n - 3 : 31))) & 7)); }

// the 3-bit XOR mask of sw(): phys(i) = i ^ swmask(i, sh), sh = swshift(n)
GLB_FFT_HD int swshift(int n) { return n >= 10 ? n - 3 : 31; }
GLB_FFT_HD int swmask(int i, int sh) { return ((i >> 4) ^ (i >> sh)) & 7; }

GLB_FFT_HD double2 csq(double2 a) { return make_double2(a.x * a.x - a.y * a.y, 2.0 * a.x * a.y); }

// v * e^{-2 pi i a/16} (forward) or its conjugate twiddle (inverse), a = 0..7; `a` is a
// compile-time constant wherever this is called (fully unrolled butterflies)
GLB_FFT_HD double2 mul_w16(double2 v, int a, bool inverse) {
  const double h = 0.70710678118654752440;   // cos(pi/4)
  const double c1 = 0.92387953251128675613;  // cos(pi/8)
  const double s1 = 0.38268343236508977173;  // sin(pi/8)
  double wr, wi;  // forward twiddle = wr - i wi
  switch (a) {
    case 0: return v;
    case 4: return inverse ? make_double2(-v.y, v.x) : make_double2(v.y, -v.x);
    case 2: return inverse ? make_double2(h * (v.x - v.y), h * (v.x + v.y)) : make_double2(h * (v.x + v.y), h * (v.y - v.x));
    case 6: return inverse ? make_double2(-h * (v.x + v.y), h * (v.x - v.y)) : make_double2(h * (v.y - v.x), -h * (v.x + v.y));
    case 1: wr = c1; wi = s1; break;
    case 3: wr = s1; wi = c1; break;
    case 5: wr = -s1; wi = c1; break;
    default: wr = -c1; wi = s1; break;
  }
  if (inverse) wi = -wi;
  return make_double2(v.x * wr + v.y * wi, v.y * wr - v.x * wi);
}

// K DIF levels on R = 2^K register values; w = e^{-+2 pi i lo/(2 s)} is the position-dependent
// twiddle of the first level (squared from level to level), UNIT: w == 1
template <int K, bool UNIT>
GLB_FFT_HD void dif_butterflies(double2* v, double2 w, bool inverse) {
  constexpr int R = 1 << K;
#pragma unroll
  for (int j = 0; j < K; ++j) {
    const int hs = R >> (j + 1);
#pragma unroll
    for (int t = 0; t < R; ++t) {
      if ((t & hs) == 0) {
        const double2 u = v[t], z = v[t + hs];
        v[t] = cadd(u, z);
        double2 d = csub(u, z);
        if (!UNIT) d = cmul(d, w);
        v[t + hs] = mul_w16(d, (t & (hs - 1)) * (8 / hs), inverse);
      }
    }
    if (!UNIT) w = csq(w);
  }
}

// K DIT levels; wl[j] = position-dependent twiddle of level j (ignored if UNIT)
template <int K, bool UNIT>
GLB_FFT_HD void dit_butterflies(double2* v, const double2* wl, bool inverse) {
  constexpr int R = 1 << K;
#pragma unroll
  for (int j = 0; j < K; ++j) {
    const int hs = 1 << j;
#pragma unroll
    for (int t = 0; t < R; ++t) {
      if ((t & hs) == 0) {
        double2 z = v[t + hs];
        if (!UNIT) z = cmul(z, wl[j]);
        z = mul_w16(z, (t & (hs - 1)) * (8 / hs), inverse);
        const double2 u = v[t];
        v[t] = cadd(u, z);
        v[t + hs] = csub(u, z);
      }
    }
  }
}

GLB_FFT_HD double2 load_tw(const double2* tw, int idx, bool inverse) {
#if defined(__CUDA_ARCH__)
  double2 w = __ldg(&tw[idx]);
#else
  double2 w = tw[idx];
#endif
  if (inverse) w.y = -w.y;
  return w;
}

// One strided DIF pass: K levels with half-spans s, s/2, ..., s/2^(K-1) of an M = 2^n point
// transform.  Elements with logical index >= nvalid are taken as zero (never read).
template <int K>
GLB_FFT_HD void dif_pass(double2* x, int n, int s, const double2* tw, int tw_n, bool inverse, int nvalid, int tid,
                         int nthreads) {
  constexpr int R = 1 << K;
  const int M = 1 << n;
  const int q = s >> (K - 1);
  const int lq = ilog2(q);
  const int tshift = ilog2(tw_n) - ilog2(2 * s);  // twiddle table stride tw_n / (2 s), powers of two
  // element t of a butterfly group sits at base + t q; base and t q occupy disjoint bits and
  // q >= 16, so the swizzle mask splits into a per-thread and a per-t (warp-uniform) part:
  //   phys(base + t q) = ((base ^ mask(base)) ^ mask(t q)) + t q
  const int sh = swshift(n);
  int ft[R];
#pragma unroll
  for (int t = 0; t < R; ++t) ft[t] = swmask(t * q, sh);
  for (int g = tid; g < (M >> K); g += nthreads) {
    const int lo = g & (q - 1);
    const int base = ((g >> lq) << (lq + K)) + lo;
    const int b0 = base ^ swmask(base, sh);
    double2 v[R];
#pragma unroll
    for (int t = 0; t < R; ++t) v[t] = (base + t * q < nvalid) ? x[(b0 ^ ft[t]) + t * q] : make_double2(0.0, 0.0);
    dif_butterflies<K, false>(v, load_tw(tw, lo << tshift, inverse), inverse);
#pragma unroll
    for (int t = 0; t < R; ++t) x[(b0 ^ ft[t]) + t * q] = v[t];
  }
}

// One strided DIT pass: K levels with half-spans s, 2s, ..., s*2^(K-1).
template <int K>
GLB_FFT_HD void dit_pass(double2* x, int n, int s, const double2* tw, int tw_n, bool inverse, int tid, int nthreads) {
  constexpr int R = 1 << K;
  const int M = 1 << n;
  const int q = s;
  const int lq = ilog2(q);
  const int tshift = ilog2(tw_n) - lq - K;  // tw_n / (q R)
  const int sh = swshift(n);
  int ft[R];
#pragma unroll
  for (int t = 0; t < R; ++t) ft[t] = swmask(t * q, sh);
  for (int g = tid; g < (M >> K); g += nthreads) {
    const int lo = g & (q - 1);
    const int base = ((g >> lq) << (lq + K)) + lo;
    const int b0 = base ^ swmask(base, sh);  // see dif_pass
    double2 v[R];
#pragma unroll
    for (int t = 0; t < R; ++t) v[t] = x[(b0 ^ ft[t]) + t * q];
    double2 wl[K];
    wl[K - 1] = load_tw(tw, lo << tshift, inverse);
#pragma unroll
    for (int j = K - 2; j >= 0; --j) wl[j] = csq(wl[j + 1]);
    dit_butterflies<K, false>(v, wl, inverse);
#pragma unroll
    for (int t = 0; t < R; ++t) x[(b0 ^ ft[t]) + t * q] = v[t];
  }
}

// schedule: the T = min(n, 4) innermost levels form the contiguous tail (DIF) / head (DIT); the
// remaining n - T levels are radix-8 passes plus, if needed, one radix-4 or radix-2 pass at the
// LARGEST spans (first in DIF, last in DIT), where its stride is conflict-free.
GLB_FFT_HD int tail_levels(int n) { return n < 4 ? n : 4; }

// DIF passes down to (excluding) the tail.  step(K, s) runs one pass and synchronises.
template <class Step>
GLB_FFT_HD void dif_upper_schedule(int n, Step&& step) {
  int r = n - tail_levels(n);
  int s = 1 << (n - 1);
  if (r % 3 == 1) {
    step(1, s);
    s >>= 1;
    r -= 1;
  } else if (r % 3 == 2) {
    step(2, s);
    s >>= 2;
    r -= 2;
  }
  for (; r > 0; r -= 3) {
    step(3, s);
    s >>= 3;
  }
}
// DIT passes above the head
template <class Step>
GLB_FFT_HD void dit_upper_schedule(int n, Step&& step) {
  int r = n - tail_levels(n);
  int s = 1 << tail_levels(n);
  for (; r >= 3; r -= 3) {
    step(3, s);
    s <<= 3;
  }
  if (r == 2)
    step(2, s);
  else if (r == 1)
    step(1, s);
}

// ---- contiguous blocks of R = 2^T elements held in registers ----
// (T == 4: the mask is the same for the 16 elements of a block, phys = 16 g + (t ^ mask))
template <int T>
GLB_FFT_HD void block_load(const double2* x, int n, int g, double2* v, int nvalid) {
  const int fg = (T == 4) ? swmask(g << 4, swshift(n)) : 0;
#pragma unroll
  for (int t = 0; t < (1 << T); ++t) {
    const int i = (g << T) + t;
    const int ph = (T == 4) ? (g << 4) + (t ^ fg) : sw(i, n);
    v[t] = (i < nvalid) ? x[ph] : make_double2(0.0, 0.0);
  }
}
template <int T>
GLB_FFT_HD void block_store(double2* x, int n, int g, const double2* v) {
  const int fg = (T == 4) ? swmask(g << 4, swshift(n)) : 0;
#pragma unroll
  for (int t = 0; t < (1 << T); ++t) x[(T == 4) ? (g << 4) + (t ^ fg) : sw((g << T) + t, n)] = v[t];
}

// DIF tail in place (result in bit-reversed order)
template <int T>
GLB_FFT_HD void dif_tail_inplace(double2* x, int n, bool inverse, int nvalid, int tid, int nthreads) {
  for (int g = tid; g < (1 << (n - T)); g += nthreads) {
    double2 v[1 << T];
    block_load<T>(x, n, g, v, nvalid);
    dif_butterflies<T, true>(v, make_double2(1.0, 0.0), inverse);
    block_store<T>(x, n, g, v);
  }
}
// DIT head in place (input in bit-reversed order)
template <int T>
GLB_FFT_HD void dit_head_inplace(double2* x, int n, bool inverse, int tid, int nthreads) {
  for (int g = tid; g < (1 << (n - T)); g += nthreads) {
    double2 v[1 << T];
    block_load<T>(x, n, g, v, 1 << n);
    dit_butterflies<T, true>(v, nullptr, inverse);
    block_store<T>(x, n, g, v);
  }
}
// Bluestein middle: DIF tail (forward), multiply by the chirp spectrum, DIT head (inverse), one
// pass.  The spectrum is stored "block transposed": element t of block g at bf[t * (M >> T) + g]
// (coalesced across the lanes of this pass), see bf_index().
GLB_FFT_HD int bf_index(int p, int n) {  // p: position in the DIF (bit-reversed) output
  const int T = tail_levels(n);
  return (p & ((1 << T) - 1)) * (1 << (n - T)) + (p >> T);
}
template <int T>
GLB_FFT_HD void bluestein_middle(double2* x, int n, const double2* bf, int nvalid, int tid, int nthreads) {
  const int nblk = 1 << (n - T);
  for (int g = tid; g < nblk; g += nthreads) {
    double2 v[1 << T];
    block_load<T>(x, n, g, v, nvalid);
    dif_butterflies<T, true>(v, make_double2(1.0, 0.0), false);
#pragma unroll
    for (int t = 0; t < (1 << T); ++t) {
#if defined(__CUDA_ARCH__)
      const double2 b = __ldg(&bf[t * nblk + g]);
#else
      const double2 b = bf[t * nblk + g];
#endif
      v[t] = cmul(v[t], b);
    }
    dit_butterflies<T, true>(v, nullptr, true);
    block_store<T>(x, n, g, v);
  }
}
// Reordering DIF tail: thread processes block bitrev(g); emit(j, value) receives the transform
// in NATURAL order, j = bitrev_T(t) * (M >> T) + g -- consecutive g, consecutive j.
// The caller must not let emit() write into x before every thread has loaded (barrier).
template <int T, class Emit>
GLB_FFT_HD void dif_tail_reorder(const double2* x, int n, bool inverse, int tid, int nthreads, Emit&& emit) {
  const int nblk = 1 << (n - T);
  for (int g = tid; g < nblk; g += nthreads) {
    double2 v[1 << T];
    block_load<T>(x, n, (int)bitrev((unsigned)g, n - T), v, 1 << n);
    dif_butterflies<T, true>(v, make_double2(1.0, 0.0), inverse);
#pragma unroll
    for (int t = 0; t < (1 << T); ++t) emit((int)bitrev((unsigned)t, T) * nblk + g, v[t]);
  }
}

#if defined(__CUDACC__)
// ---- device drivers (all threads of the CTA call these; they end with a barrier) ----
template <int THREADS>
__device__ __forceinline__ void dev_dif_upper(double2* x, int n, const double2* __restrict__ tw, int tw_n, bool inverse,
                                              int nvalid) {
  dif_upper_schedule(n, [&](int K, int s) {
    if (K == 3)
      dif_pass<3>(x, n, s, tw, tw_n, inverse, nvalid, threadIdx.x, THREADS);
    else if (K == 2)
      dif_pass<2>(x, n, s, tw, tw_n, inverse, nvalid, threadIdx.x, THREADS);
    else
      dif_pass<1>(x, n, s, tw, tw_n, inverse, nvalid, threadIdx.x, THREADS);
    nvalid = 1 << n;  // the first pass has filled every element
    __syncthreads();
  });
}
template <int THREADS>
__device__ __forceinline__ void dev_dit_upper(double2* x, int n, const double2* __restrict__ tw, int tw_n, bool inverse) {
  dit_upper_schedule(n, [&](int K, int s) {
    if (K == 3)
      dit_pass<3>(x, n, s, tw, tw_n, inverse, threadIdx.x, THREADS);
    else if (K == 2)
      dit_pass<2>(x, n, s, tw, tw_n, inverse, threadIdx.x, THREADS);
    else
      dit_pass<1>(x, n, s, tw, tw_n, inverse, threadIdx.x, THREADS);
    __syncthreads();
  });
}
// full DIF, result left in x in bit-reversed order (logical positions, swizzled layout)
template <int THREADS>
__device__ __forceinline__ void dev_fft_dif(double2* x, int n, const double2* __restrict__ tw, int tw_n, bool inverse) {
  dev_dif_upper<THREADS>(x, n, tw, tw_n, inverse, 1 << n);
  switch (tail_levels(n)) {
    case 4: dif_tail_inplace<4>(x, n, inverse, 1 << n, threadIdx.x, THREADS); break;
    case 3: dif_tail_inplace<3>(x, n, inverse, 1 << n, threadIdx.x, THREADS); break;
    case 2: dif_tail_inplace<2>(x, n, inverse, 1 << n, threadIdx.x, THREADS); break;
    case 1: dif_tail_inplace<1>(x, n, inverse, 1 << n, threadIdx.x, THREADS); break;
    default: break;
  }
  __syncthreads();
}
// full DIF with the natural-order results handed to emit(j, value); x is only read by the tail
template <int THREADS, class Emit>
__device__ __forceinline__ void dev_fft_dif_emit(double2* x, int n, const double2* __restrict__ tw, int tw_n, bool inverse,
                                                 Emit&& emit) {
  dev_dif_upper<THREADS>(x, n, tw, tw_n, inverse, 1 << n);
  switch (tail_levels(n)) {
    case 4: dif_tail_reorder<4>(x, n, inverse, threadIdx.x, THREADS, emit); break;
    case 3: dif_tail_reorder<3>(x, n, inverse, threadIdx.x, THREADS, emit); break;
    case 2: dif_tail_reorder<2>(x, n, inverse, threadIdx.x, THREADS, emit); break;
    case 1: dif_tail_reorder<1>(x, n, inverse, threadIdx.x, THREADS, emit); break;
    default:
      if (threadIdx.x == 0) emit(0, x[0]);
      break;
  }
}
// circular convolution with the chirp whose (scaled, block-transposed) spectrum is bf:
// x (first nvalid elements, the rest taken as zero) -> IDFT(DFT(x) .* bf), natural order
template <int THREADS>
__device__ __forceinline__ void dev_bluestein_conv(double2* x, int n, const double2* __restrict__ tw, int tw_n,
                                                   const double2* __restrict__ bf, int nvalid) {
  const int upper = n - tail_levels(n);
  dev_dif_upper<THREADS>(x, n, tw, tw_n, false, nvalid);
  const int nv = upper > 0 ? (1 << n) : nvalid;
  switch (tail_levels(n)) {
    case 4: bluestein_middle<4>(x, n, bf, nv, threadIdx.x, THREADS); break;
    case 3: bluestein_middle<3>(x, n, bf, nv, threadIdx.x, THREADS); break;
    case 2: bluestein_middle<2>(x, n, bf, nv, threadIdx.x, THREADS); break;
    case 1: bluestein_middle<1>(x, n, bf, nv, threadIdx.x, THREADS); break;
    default:
      if (threadIdx.x == 0) x[0] = cmul(nv > 0 ? x[0] : make_double2(0.0, 0.0), bf[0]);
      break;
  }
  __syncthreads();
  dev_dit_upper<THREADS>(x, n, tw, tw_n, true);
}
#endif

}  // namespace fft
}  // namespace glb
