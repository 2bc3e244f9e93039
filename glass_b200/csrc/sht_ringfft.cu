// sht_ringfft.cu -- phase shift + alias fold + per-ring complex-to-real FFT (K4) with the
// pixel transform (K5: lognormal / squared-normal, glass/grf/_transformations.py:83-89,
// :170) fused into the store.
//
// Replaces the Fourier part of healpy.alm2map (glass/healpix.py:71):
//     f(theta_r, phi_j) = Re F_0 + 2 Re sum_{m>0} F_m e^{i m (phi0_r + 2 pi j / nphi_r)}
//
// One CTA per (ring, map).  Ring lengths are nphi = 4i (caps) and 4*nside (belt), so the
// FFT length is arbitrary.  Per ring, with n = nphi, h = n/2:
//   1. fold   G[k] = sum_{m = k (mod n)} t_m + sum_{m = -k (mod n)} conj(t_m),
//             t_m = F_m e^{i m phi0}, m <= mlim(ring)               (k = 0..h)
//   2. pack   Z[k] = (G[k] + conj G[h-k]) + i e^{2 pi i k/n} (G[k] - conj G[h-k])
//             so that the length-h complex inverse DFT z of Z holds the real ring as
//             x[2j] = Re z[j], x[2j+1] = Im z[j]
//   3. DFT_h  h a power of two: register-blocked DIF in shared memory (fft_core.cuh: radix-8
//             passes + one contiguous radix-16 pass, swizzled conflict-free layout) whose last
//             pass hands out the result in natural order.
//             otherwise h = 2L: two length-L DFTs (even/odd k) by Bluestein's chirp-z
//             with a power-of-two circular convolution of length M >= 2L-1
//             (forward DIF -> multiply by the precomputed chirp spectrum, stored in the
//             matching order -> inverse DIT: no bit reversal anywhere, and DIF tail, multiply
//             and DIT head are one pass), then one radix-2 butterfly.  Intermediate values
//             live in registers so shared memory only ever holds one M-length buffer.
//   4. store  coalesced 16-byte stores of (x[2j], x[2j+1]) with the transform applied --
//             for the direct path straight from the registers of the last FFT pass.
#include <algorithm>
#include <vector>

#include "expm1_fast.cuh"
#include "fft_core.cuh"
#include "plan.h"

namespace glb {

struct FftParams {
  const RingDesc* rings;
  const int* order;
  const double2* phase;
  int64_t phase_map_stride;  // double2 units
  const int* mlim;           // per ring pair
  const double2* tw;
  const double2* bf;
  double* maps[4];   // output map of each batch entry
  // phase addressing: F(m) = base + rowidx(ring)*W + (m % G)*blk + m / G
  // single GPU: G = 1, W = mmax+1, blk = 0, rowidx = null (identity)
  int G, W;
  int64_t blk;
  const int* rowidx;
  int mmax;
  int tw_n;
  int stash_off;     // offset (double2) of the Bluestein stash behind the FFT buffer
  double2* scratch;  // long rings only: [gridDim.x][3][LONG_LB + 8] work buffers in global memory
  int nitems;        // long rings only: rings x maps handed out to the resident CTAs
  int nrings;
  int kind[4];
  double p0[4];
  double p1[4];
};

__device__ __forceinline__ double apply_transform(double x, int kind, double p0, double p1) {
  if (kind == GLB_T_LOGNORMAL) {
    x = expm1_fast(x - p0);
    if (p1 != 1.0) x = p1 * x;
  } else if (kind == GLB_T_SQUARED_NORMAL) {
    const double d = x - p0;
    x = d * d - 1.0;
    if (p1 != 1.0) x = p1 * x;
  }
  return x;
}

// Roots of unity of one ring from two small shared-memory tables: T(j) = e^{2 pi i j/(2 nphi)},
// j < 2 nphi, as hi[j >> 6] * lo[j & 63].  Every trigonometric factor of the ring stage is one of
// them (phase shift e^{i pi m/nphi} = T(m mod 2nphi); real-FFT twiddle e^{2 pi i k/nphi} = T(2k);
// chirp e^{i pi q^2/L} = T(4 (q^2 mod 2L)); e^{2 pi i q/h} = T(4q)), so a CTA evaluates
// sincospi ~600 times instead of ~6 times per ring element.
struct RingTrig {
  const double2* hi;
  const double2* lo;
  __device__ __forceinline__ double2 T(int j) const { return cmul(hi[j >> 6], lo[j & 63]); }
};
constexpr int TRIG_HI = 512;  // 2 * 16384 / 64
template <int THREADS>
__device__ __forceinline__ RingTrig ring_trig_setup(double2* s_hi, double2* s_lo, int nphi) {
  const double inv = 1.0 / (double)nphi;  // j/(2 nphi) turns = j/nphi half-turns
  for (int j = threadIdx.x; j < 64; j += THREADS) s_lo[j] = cispi((double)j * inv);
  for (int a = threadIdx.x; a * 64 < 2 * nphi; a += THREADS) s_hi[a] = cispi((double)(a * 64) * inv);
  __syncthreads();
  RingTrig t;
  t.hi = s_hi;
  t.lo = s_lo;
  return t;
}

// chirp c[q] = e^{i pi q^2 / L}
__device__ __forceinline__ double2 chirp(int q, int L) {
  const long long q2 = ((long long)q * q) % (2LL * L);
  return cispi((double)q2 / (double)L);
}
// the same from the ring's tables (q < 2^15: q^2 fits 32 bits unsigned)
__device__ __forceinline__ double2 chirp(const RingTrig& tr, int q, int L) {
  return tr.T(4 * (int)(((unsigned)q * (unsigned)q) % (unsigned)(2 * L)));
}

// Z[k] from the folded bins
__device__ __forceinline__ double2 pack_z(const RingTrig& tr, const double2* G, int k, int h) {
  const double2 g = G[k];
  const double2 gr = cconj(G[h - k]);
  const double2 s = cadd(g, gr);
  const double2 d = csub(g, gr);
  const double2 w = tr.T(2 * k);  // e^{2 pi i k/nphi}
  const double2 wd = cmul(w, d);  // i*wd = (-wd.y, wd.x)
  return make_double2(s.x - wd.y, s.y + wd.x);
}

template <int THREADS, int NREG>
__global__ void __launch_bounds__(THREADS) sht_ringfft_synth_kernel(const FftParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double2* buf = reinterpret_cast<double2*>(smem_raw);
  const int tid = threadIdx.x;
  const int ring = p.order[blockIdx.x];
  const int b = blockIdx.y;
  const RingDesc d = p.rings[ring];
  const int n = d.nphi, h = n >> 1;
  const int mlim = min(p.mlim[d.pair], p.mmax);
  const int row = p.rowidx ? p.rowidx[ring] : ring;
  const double2* __restrict__ Fb = p.phase + b * p.phase_map_stride + (int64_t)row * p.W;
  const bool one_gpu = (p.G == 1);  // the m-split layout needs a division per element
  auto F = [&](int m) -> double2 { return one_gpu ? Fb[m] : Fb[(int64_t)(m % p.G) * p.blk + m / p.G]; };
  double* __restrict__ out = p.maps[b] + d.start;
  const int kind = p.kind[b];
  const double tp0 = p.p0[b], tp1 = p.p1[b];
  __shared__ double2 s_thi[TRIG_HI], s_tlo[64];
  const RingTrig tr = ring_trig_setup<THREADS>(s_thi, s_tlo, n);

  // ---- 1. fold with phase shift ----
  if (h > mlim && h <= NREG * THREADS) {
    // no aliasing (every belt ring, and cap rings wider than their m range): G[k] = t_k for
    // k <= mlim and 0 above.  All loads of a thread are issued before the first use, so the
    // CTA has the whole ring row in flight at once instead of one 16-byte load per thread and
    // loop trip (ncu: the dependent load was 12 % of the kernel's stall samples).
    double2 v[NREG];
#pragma unroll
    for (int t = 0; t < NREG; ++t) {
      const int k = tid + t * THREADS;
      v[t] = (k <= mlim) ? F(k) : make_double2(0.0, 0.0);
    }
    if (tid == 0) {
      v[0].y = 0.0;
      buf[h] = make_double2(0.0, 0.0);
    }
#pragma unroll
    for (int t = 0; t < NREG; ++t) {
      const int k = tid + t * THREADS;
      if (k < h) {
        if (d.shifted && k <= mlim) v[t] = cmul(v[t], tr.T(k));
        buf[k] = v[t];
      }
    }
  } else {
    for (int k = tid; k <= h; k += THREADS) {
      double2 g = make_double2(0.0, 0.0);
      // m = k, k + n, ...: m mod 2n alternates between j and j + n (mod 2n), no division needed
      int j = k;  // k <= n/2 < 2n
      for (int m = k; m <= mlim; m += n) {
        double2 t = F(m);
        if (m == 0) t.y = 0.0;
        if (d.shifted) t = cmul(t, tr.T(j));
        g = cadd(g, t);
        j += n;
        if (j >= 2 * n) j -= 2 * n;
      }
      j = n - k;  // n/2 <= n - k <= n
      for (int m = n - k; m <= mlim; m += n) {
        double2 t = F(m);
        if (d.shifted) t = cmul(t, tr.T(j));
        g = cadd(g, cconj(t));
        j += n;
        if (j >= 2 * n) j -= 2 * n;
      }
      buf[k] = g;
    }
  }
  __syncthreads();

  double2 reg[NREG];
  const int nfft = 31 - __clz(d.M);  // log2 of the power-of-two FFT length
  if (d.L == 0) {
    // ---- direct power-of-two path ----
#pragma unroll
    for (int t = 0; t < NREG; ++t) {
      const int k = tid + t * THREADS;
      if (k < h) reg[t] = pack_z(tr, buf, k, h);
    }
    __syncthreads();
#pragma unroll
    for (int t = 0; t < NREG; ++t) {
      const int k = tid + t * THREADS;
      if (k < h) buf[fft::sw(k, nfft)] = reg[t];
    }
    __syncthreads();
    // the last pass emits z[j] in natural order: consecutive lanes, consecutive 16-byte stores
    fft::dev_fft_dif_emit<THREADS>(buf, nfft, p.tw, p.tw_n, true, [&](int j, double2 zv) {
      double2 o;
      o.x = apply_transform(zv.x, kind, tp0, tp1);
      o.y = apply_transform(zv.y, kind, tp0, tp1);
      *reinterpret_cast<double2*>(out + 2 * j) = o;
    });
    return;
  }

  // ---- Bluestein path: h = 2L ----
  // even-k sequence in registers, odd-k sequence parked in the shared-memory stash while the
  // FFT buffer is busy with the other one
  const int L = d.L;
  constexpr int NQ = NREG / 2;
  double2* stash = buf + p.stash_off;
  double2* ev = reg;
#pragma unroll
  for (int t = 0; t < NQ; ++t) {
    const int q = tid + t * THREADS;
    if (q < L) {
      const double2 c = chirp(tr, q, L);
      ev[t] = cmul(pack_z(tr, buf, 2 * q, h), c);
      stash[q] = cmul(pack_z(tr, buf, 2 * q + 1, h), c);
    }
  }
  __syncthreads();
  const double2* __restrict__ bf = p.bf + d.bf_off;
#pragma unroll
  for (int t = 0; t < NQ; ++t) {
    const int q = tid + t * THREADS;
    if (q < L) buf[fft::sw(q, nfft)] = ev[t];
  }
  __syncthreads();
  fft::dev_bluestein_conv<THREADS>(buf, nfft, p.tw, p.tw_n, bf, L);  // elements >= L count as zero
#pragma unroll
  for (int t = 0; t < NQ; ++t) {
    const int q = tid + t * THREADS;
    if (q < L) ev[t] = buf[fft::sw(q, nfft)];
  }
  __syncthreads();
#pragma unroll
  for (int t = 0; t < NQ; ++t) {
    const int q = tid + t * THREADS;
    if (q < L) buf[fft::sw(q, nfft)] = stash[q];
  }
  __syncthreads();
  fft::dev_bluestein_conv<THREADS>(buf, nfft, p.tw, p.tw_n, bf, L);
#pragma unroll
  for (int t = 0; t < NQ; ++t) {
    const int q = tid + t * THREADS;
    if (q < L) {
      const double2 c = chirp(tr, q, L);
      const double2 E = cmul(ev[t], c);
      const double2 O = cmul(cmul(buf[fft::sw(q, nfft)], c), tr.T(4 * q));  // e^{2 pi i q/h}
      const double2 z0 = cadd(E, O), z1 = csub(E, O);
      double2 o;
      o.x = apply_transform(z0.x, kind, tp0, tp1);
      o.y = apply_transform(z0.y, kind, tp0, tp1);
      *reinterpret_cast<double2*>(out + 2 * q) = o;
      o.x = apply_transform(z1.x, kind, tp0, tp1);
      o.y = apply_transform(z1.y, kind, tp0, tp1);
      *reinterpret_cast<double2*>(out + 2 * (q + L)) = o;
    }
  }
}

// -------------------------------------------------------------------------------------
// analysis direction (first stage of map2alm, glass/healpix.py:270): one CTA per (ring, map)
//   real ring x[0..n) -> z[j] = x[2j] + i x[2j+1] -> forward DFT_h (same machinery, run as
//   conj(IDFT(conj .))) -> X[k] = sum_t x_t e^{-2 pi i tk/n} -> G_m = X[m mod n] e^{-i m phi0}
//   scaled by the ring's quadrature weight * 4 pi / npix, for m <= mlim(ring).
// -------------------------------------------------------------------------------------
struct FftAnaParams {
  const RingDesc* rings;
  const int* order;
  double2* phase;
  int64_t phase_map_stride;
  const int* mlim;
  const double2* tw;
  const double2* bf;
  const double* maps[4];
  const double* ring_w;   // [nring] or null
  double norm;            // 4 pi / npix
  int mmax;
  int tw_n;
  int stash_off;
  double2* scratch;       // long rings only, as in FftParams
  int nitems;
  int nrings;
};

template <int THREADS, int NREG>
__global__ void __launch_bounds__(THREADS) sht_ringfft_analysis_kernel(const FftAnaParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double2* buf = reinterpret_cast<double2*>(smem_raw);
  const int tid = threadIdx.x;
  const int ring = p.order[blockIdx.x];
  const int b = blockIdx.y;
  const RingDesc d = p.rings[ring];
  const int n = d.nphi, h = n >> 1;
  const int mlim = min(p.mlim[d.pair], p.mmax);
  const double* __restrict__ in = p.maps[b] + d.start;
  double2* __restrict__ F = p.phase + b * p.phase_map_stride + (int64_t)ring * (p.mmax + 1);
  const double wscale = p.norm * (p.ring_w ? p.ring_w[ring] : 1.0);
  const int nfft = 31 - __clz(d.M);
  bool bitrev_out;
  __shared__ double2 s_thi[TRIG_HI], s_tlo[64];
  const RingTrig tr = ring_trig_setup<THREADS>(s_thi, s_tlo, n);

  if (d.L == 0) {
    for (int j = tid; j < h; j += THREADS) buf[fft::sw(j, nfft)] = *reinterpret_cast<const double2*>(in + 2 * j);
    __syncthreads();
    fft::dev_fft_dif<THREADS>(buf, nfft, p.tw, p.tw_n, false);
    bitrev_out = true;
  } else {
    const int L = d.L;
    constexpr int NQ = NREG / 2;
    double2* stash = buf + p.stash_off;
    double2 ev[NQ], od[NQ];
#pragma unroll
    for (int t = 0; t < NQ; ++t) {
      const int q = tid + t * THREADS;
      if (q < L) {
        const double2 c = chirp(tr, q, L);
        const double2 ze = *reinterpret_cast<const double2*>(in + 4 * q);
        const double2 zo = *reinterpret_cast<const double2*>(in + 4 * q + 2);
        buf[fft::sw(q, nfft)] = cmul(cconj(ze), c);  // forward DFT as conj(IDFT(conj .))
        stash[q] = cmul(cconj(zo), c);
      }
    }
    __syncthreads();
    const double2* __restrict__ bf = p.bf + d.bf_off;
    fft::dev_bluestein_conv<THREADS>(buf, nfft, p.tw, p.tw_n, bf, L);
#pragma unroll
    for (int t = 0; t < NQ; ++t) {
      const int q = tid + t * THREADS;
      if (q < L) ev[t] = buf[fft::sw(q, nfft)];
    }
    __syncthreads();
#pragma unroll
    for (int t = 0; t < NQ; ++t) {
      const int q = tid + t * THREADS;
      if (q < L) buf[fft::sw(q, nfft)] = stash[q];
    }
    __syncthreads();
    fft::dev_bluestein_conv<THREADS>(buf, nfft, p.tw, p.tw_n, bf, L);
#pragma unroll
    for (int t = 0; t < NQ; ++t) {
      const int q = tid + t * THREADS;
      if (q < L) od[t] = buf[fft::sw(q, nfft)];
    }
    __syncthreads();
    // Zf[k] = E[k] + w^k O[k], Zf[k+L] = E[k] - w^k O[k], w = e^{-2 pi i / h}; natural order in buf
#pragma unroll
    for (int t = 0; t < NQ; ++t) {
      const int q = tid + t * THREADS;
      if (q < L) {
        const double2 c = chirp(tr, q, L);
        const double2 E = cconj(cmul(ev[t], c));
        const double2 O = cmul(cconj(cmul(od[t], c)), cconj(tr.T(4 * q)));  // e^{-2 pi i q/h}
        buf[q] = cadd(E, O);
        buf[q + L] = csub(E, O);
      }
    }
    __syncthreads();
    bitrev_out = false;
  }

  for (int m = tid; m <= mlim; m += THREADS) {
    const int k = m % n;
    const int kk = (k <= h) ? k : n - k;
    int i0 = kk % h, i1 = (h - kk) % h;
    if (bitrev_out) {  // logical bit-reversed position, swizzled layout
      i0 = fft::sw((int)fft::bitrev((unsigned)i0, nfft), nfft);
      i1 = fft::sw((int)fft::bitrev((unsigned)i1, nfft), nfft);
    }
    const double2 zk = buf[i0];
    const double2 zr = cconj(buf[i1]);
    const double2 sum = cadd(zk, zr), dif = csub(zk, zr);
    const double2 wd = cmul(cconj(tr.T(2 * kk)), dif);  // -i*wd = (wd.y, -wd.x)
    double2 X = make_double2(0.5 * (sum.x + wd.y), 0.5 * (sum.y - wd.x));
    if (k > h) X = cconj(X);
    if (d.shifted) X = cmul(X, cconj(tr.T(m % (2 * n))));
    F[m] = cscale(X, wscale);
  }
}

// -------------------------------------------------------------------------------------
// LONG rings: FFT length above 8192 complex points (ring length above 16384 pixels, i.e. nside
// above 4096).  A 16384-point complex FP64 buffer is 256 KB and no longer fits one SM's shared
// memory, so these rings run the SAME algorithm (fold / pack / power-of-two or two-sequence
// Bluestein with the fft_core passes) on three work buffers in GLOBAL memory -- one set per
// resident CTA, a few hundred KB each, so that they live in the 126 MB L2 -- with a persistent
// grid handing out (ring, map) items.  Every pass is then an L2 round trip instead of a
// shared-memory one: slower per ring than the shared-memory kernels, but the Fourier stage is a
// few per cent of a transform whose Legendre stage grows with nside^3.  Covers FFT lengths up to
// 16384 (nside <= 8192).
// -------------------------------------------------------------------------------------
constexpr int LONG_LB = 16384;
constexpr int LONG_THREADS = 512;
constexpr int LONG_TRIG_HI = 2 * 4 * 8192 / 64;  // 2 nphi / 64 for nphi = 4 * 8192

__global__ void __launch_bounds__(LONG_THREADS) sht_ringfft_synth_long_kernel(const FftParams p) {
  constexpr int THREADS = LONG_THREADS;
  __shared__ double2 s_thi[LONG_TRIG_HI], s_tlo[64];
  double2* bufG = p.scratch + (int64_t)blockIdx.x * 3 * (LONG_LB + 8);
  double2* bufA = bufG + (LONG_LB + 8);
  double2* bufB = bufA + (LONG_LB + 8);
  const int tid = threadIdx.x;
  for (int item = blockIdx.x; item < p.nitems; item += gridDim.x) {
    const int ring = p.order[item % p.nrings];
    const int b = item / p.nrings;
    const RingDesc d = p.rings[ring];
    const int n = d.nphi, h = n >> 1;
    const int mlim = min(p.mlim[d.pair], p.mmax);
    const int row = p.rowidx ? p.rowidx[ring] : ring;
    const double2* __restrict__ Fb = p.phase + b * p.phase_map_stride + (int64_t)row * p.W;
    const bool one_gpu = (p.G == 1);
    auto F = [&](int m) -> double2 { return one_gpu ? Fb[m] : Fb[(int64_t)(m % p.G) * p.blk + m / p.G]; };
    double* __restrict__ out = p.maps[b] + d.start;
    const int kind = p.kind[b];
    const double tp0 = p.p0[b], tp1 = p.p1[b];
    __syncthreads();  // the previous item's readers of the tables and buffers are done
    const RingTrig tr = ring_trig_setup<THREADS>(s_thi, s_tlo, n);
    // 1. fold with phase shift -> G[0..h]
    for (int k = tid; k <= h; k += THREADS) {
      double2 g = make_double2(0.0, 0.0);
      int j = k;
      for (int m = k; m <= mlim; m += n) {
        double2 t = F(m);
        if (m == 0) t.y = 0.0;
        if (d.shifted) t = cmul(t, tr.T(j));
        g = cadd(g, t);
        j += n;
        if (j >= 2 * n) j -= 2 * n;
      }
      j = n - k;
      for (int m = n - k; m <= mlim; m += n) {
        double2 t = F(m);
        if (d.shifted) t = cmul(t, tr.T(j));
        g = cadd(g, cconj(t));
        j += n;
        if (j >= 2 * n) j -= 2 * n;
      }
      bufG[k] = g;
    }
    __syncthreads();
    const int nfft = 31 - __clz(d.M);
    if (d.L == 0) {
      for (int k = tid; k < h; k += THREADS) bufA[fft::sw(k, nfft)] = pack_z(tr, bufG, k, h);
      __syncthreads();
      fft::dev_fft_dif_emit<THREADS>(bufA, nfft, p.tw, p.tw_n, true, [&](int j, double2 zv) {
        double2 o;
        o.x = apply_transform(zv.x, kind, tp0, tp1);
        o.y = apply_transform(zv.y, kind, tp0, tp1);
        *reinterpret_cast<double2*>(out + 2 * j) = o;
      });
      continue;
    }
    const int L = d.L;
    for (int q = tid; q < L; q += THREADS) {
      const double2 c = chirp(tr, q, L);
      bufA[fft::sw(q, nfft)] = cmul(pack_z(tr, bufG, 2 * q, h), c);
      bufB[fft::sw(q, nfft)] = cmul(pack_z(tr, bufG, 2 * q + 1, h), c);
    }
    __syncthreads();
    const double2* __restrict__ bf = p.bf + d.bf_off;
    fft::dev_bluestein_conv<THREADS>(bufA, nfft, p.tw, p.tw_n, bf, L);
    fft::dev_bluestein_conv<THREADS>(bufB, nfft, p.tw, p.tw_n, bf, L);
    for (int q = tid; q < L; q += THREADS) {
      const double2 c = chirp(tr, q, L);
      const double2 E = cmul(bufA[fft::sw(q, nfft)], c);
      const double2 O = cmul(cmul(bufB[fft::sw(q, nfft)], c), tr.T(4 * q));  // e^{2 pi i q/h}
      const double2 z0 = cadd(E, O), z1 = csub(E, O);
      double2 o;
      o.x = apply_transform(z0.x, kind, tp0, tp1);
      o.y = apply_transform(z0.y, kind, tp0, tp1);
      *reinterpret_cast<double2*>(out + 2 * q) = o;
      o.x = apply_transform(z1.x, kind, tp0, tp1);
      o.y = apply_transform(z1.y, kind, tp0, tp1);
      *reinterpret_cast<double2*>(out + 2 * (q + L)) = o;
    }
  }
}

__global__ void __launch_bounds__(LONG_THREADS) sht_ringfft_analysis_long_kernel(const FftAnaParams p) {
  constexpr int THREADS = LONG_THREADS;
  __shared__ double2 s_thi[LONG_TRIG_HI], s_tlo[64];
  double2* bufG = p.scratch + (int64_t)blockIdx.x * 3 * (LONG_LB + 8);
  double2* bufA = bufG + (LONG_LB + 8);
  double2* bufB = bufA + (LONG_LB + 8);
  const int tid = threadIdx.x;
  for (int item = blockIdx.x; item < p.nitems; item += gridDim.x) {
    const int ring = p.order[item % p.nrings];
    const int b = item / p.nrings;
    const RingDesc d = p.rings[ring];
    const int n = d.nphi, h = n >> 1;
    const int mlim = min(p.mlim[d.pair], p.mmax);
    const double* __restrict__ in = p.maps[b] + d.start;
    double2* __restrict__ F = p.phase + b * p.phase_map_stride + (int64_t)ring * (p.mmax + 1);
    const double wscale = p.norm * (p.ring_w ? p.ring_w[ring] : 1.0);
    const int nfft = 31 - __clz(d.M);
    __syncthreads();
    const RingTrig tr = ring_trig_setup<THREADS>(s_thi, s_tlo, n);
    const double2* Z;   // forward transform of z[j] = x[2j] + i x[2j+1], length h
    bool bitrev_out;
    if (d.L == 0) {
      for (int j = tid; j < h; j += THREADS) bufA[fft::sw(j, nfft)] = *reinterpret_cast<const double2*>(in + 2 * j);
      __syncthreads();
      fft::dev_fft_dif<THREADS>(bufA, nfft, p.tw, p.tw_n, false);
      Z = bufA;
      bitrev_out = true;
    } else {
      const int L = d.L;
      for (int q = tid; q < L; q += THREADS) {
        const double2 c = chirp(tr, q, L);
        const double2 ze = *reinterpret_cast<const double2*>(in + 4 * q);
        const double2 zo = *reinterpret_cast<const double2*>(in + 4 * q + 2);
        bufA[fft::sw(q, nfft)] = cmul(cconj(ze), c);  // forward DFT as conj(IDFT(conj .))
        bufB[fft::sw(q, nfft)] = cmul(cconj(zo), c);
      }
      __syncthreads();
      const double2* __restrict__ bf = p.bf + d.bf_off;
      fft::dev_bluestein_conv<THREADS>(bufA, nfft, p.tw, p.tw_n, bf, L);
      fft::dev_bluestein_conv<THREADS>(bufB, nfft, p.tw, p.tw_n, bf, L);
      for (int q = tid; q < L; q += THREADS) {
        const double2 c = chirp(tr, q, L);
        const double2 E = cconj(cmul(bufA[fft::sw(q, nfft)], c));
        const double2 O = cmul(cconj(cmul(bufB[fft::sw(q, nfft)], c)), cconj(tr.T(4 * q)));  // e^{-2 pi i q/h}
        bufG[q] = cadd(E, O);
        bufG[q + L] = csub(E, O);
      }
      __syncthreads();
      Z = bufG;
      bitrev_out = false;
    }
    for (int m = tid; m <= mlim; m += THREADS) {
      const int k = m % n;
      const int kk = (k <= h) ? k : n - k;
      int i0 = kk % h, i1 = (h - kk) % h;
      if (bitrev_out) {
        i0 = fft::sw((int)fft::bitrev((unsigned)i0, nfft), nfft);
        i1 = fft::sw((int)fft::bitrev((unsigned)i1, nfft), nfft);
      }
      const double2 zk = Z[i0];
      const double2 zr = cconj(Z[i1]);
      const double2 sum = cadd(zk, zr), dif = csub(zk, zr);
      const double2 wd = cmul(cconj(tr.T(2 * kk)), dif);
      double2 X = make_double2(0.5 * (sum.x + wd.y), 0.5 * (sum.y - wd.x));
      if (k > h) X = cconj(X);
      if (d.shifted) X = cmul(X, cconj(tr.T(m % (2 * n))));
      F[m] = cscale(X, wscale);
    }
  }
}

// chirp spectrum of a LONG ring's Bluestein length, work buffer in global memory (one CTA each)
__global__ void __launch_bounds__(LONG_THREADS) bluestein_spectrum_long_kernel(const int* Ls, const int* Ms, const int64_t* offs,
                                                                               const double2* tw, int tw_n, double2* bf,
                                                                               double2* scratch) {
  constexpr int THREADS = LONG_THREADS;
  double2* buf = scratch + (int64_t)blockIdx.x * (LONG_LB + 8);
  const int tid = threadIdx.x;
  const int L = Ls[blockIdx.x], M = Ms[blockIdx.x];
  const int nfft = 31 - __clz(M);
  for (int i = tid; i < M + 8; i += THREADS) buf[i] = make_double2(0.0, 0.0);
  __syncthreads();
  for (int q = tid; q < L; q += THREADS) {
    const double2 c = cconj(chirp(q, L));
    buf[fft::sw(q, nfft)] = c;
    if (q > 0) buf[fft::sw(M - q, nfft)] = c;
  }
  __syncthreads();
  fft::dev_fft_dif<THREADS>(buf, nfft, tw, tw_n, false);
  const double inv = 1.0 / (double)M;
  double2* o = bf + offs[blockIdx.x];
  for (int i = tid; i < M; i += THREADS) o[fft::bf_index(i, nfft)] = cscale(buf[fft::sw(i, nfft)], inv);
}

// chirp spectrum  Bf = DIF_M( b_wrapped ) / M,  b[d] = conj(c[d]) = e^{-i pi d^2 / L}, stored in the
// block-transposed order the fused Bluestein middle pass reads (fft::bf_index)
template <int THREADS>
__global__ void __launch_bounds__(THREADS) bluestein_spectrum_kernel(const int* Ls, const int* Ms, const int64_t* offs,
                                                                     const double2* tw, int tw_n, double2* bf) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double2* buf = reinterpret_cast<double2*>(smem_raw);
  const int tid = threadIdx.x;
  const int L = Ls[blockIdx.x], M = Ms[blockIdx.x];
  const int nfft = 31 - __clz(M);
  for (int i = tid; i < M + 8; i += THREADS) buf[i] = make_double2(0.0, 0.0);
  __syncthreads();
  for (int q = tid; q < L; q += THREADS) {
    const double2 c = cconj(chirp(q, L));
    buf[fft::sw(q, nfft)] = c;
    if (q > 0) buf[fft::sw(M - q, nfft)] = c;
  }
  __syncthreads();
  fft::dev_fft_dif<THREADS>(buf, nfft, tw, tw_n, false);
  const double inv = 1.0 / (double)M;
  double2* o = bf + offs[blockIdx.x];
  for (int i = tid; i < M; i += THREADS) o[fft::bf_index(i, nfft)] = cscale(buf[fft::sw(i, nfft)], inv);
}

// -------------------------------------------------------------------------------------
// host side
// -------------------------------------------------------------------------------------
// classes 0..2: shared-memory kernels; class 3: LONG rings (buffers in global memory)
static const int kClassLB[4] = {512, 2048, 8192, LONG_LB};

int ringfft_class_of(int lbuf) {
  for (int c = 0; c < 4; ++c)
    if (lbuf <= kClassLB[c]) return c;
  return -1;
}

// resident CTAs of the long-ring kernels and their work buffers (allocated with the plan when it
// has long rings): two CTAs per SM keep the L2 round trips of the passes overlapped
int ringfft_long_slots(int device) {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  return 2 * sms;
}
size_t ringfft_long_scratch_bytes(int device) { return (size_t)ringfft_long_slots(device) * 3 * (LONG_LB + 8) * sizeof(double2); }

int ringfft_build_spectra(glb_plan* pl, const std::vector<int>& Ls, const std::vector<int>& Ms,
                          const std::vector<int64_t>& offs, cudaStream_t st) {
  if (Ls.empty()) return GLB_OK;
  int *dL = nullptr, *dM = nullptr;
  int64_t* dO = nullptr;
  const size_t n = Ls.size();
  GLB_CUDA_CHECK(cudaMalloc(&dL, n * sizeof(int)));
  GLB_CUDA_CHECK(cudaMalloc(&dM, n * sizeof(int)));
  GLB_CUDA_CHECK(cudaMalloc(&dO, n * sizeof(int64_t)));
  GLB_CUDA_CHECK(cudaMemcpyAsync(dL, Ls.data(), n * sizeof(int), cudaMemcpyHostToDevice, st));
  GLB_CUDA_CHECK(cudaMemcpyAsync(dM, Ms.data(), n * sizeof(int), cudaMemcpyHostToDevice, st));
  GLB_CUDA_CHECK(cudaMemcpyAsync(dO, offs.data(), n * sizeof(int64_t), cudaMemcpyHostToDevice, st));
  // the lists are ordered by ring, i.e. not by length: split into the shared-memory kernel's
  // share (M <= 8192) and the long one's, each launched on its own sub-list
  std::vector<int> idx_s, idx_l;
  for (size_t i = 0; i < n; ++i) (Ms[i] <= kClassLB[2] ? idx_s : idx_l).push_back((int)i);
  auto gather = [&](const std::vector<int>& idx, std::vector<int>& l, std::vector<int>& m, std::vector<int64_t>& o) {
    for (int i : idx) {
      l.push_back(Ls[i]);
      m.push_back(Ms[i]);
      o.push_back(offs[i]);
    }
  };
  std::vector<int> l2, m2;
  std::vector<int64_t> o2;
  gather(idx_s, l2, m2, o2);
  const size_t ns = l2.size();
  gather(idx_l, l2, m2, o2);
  GLB_CUDA_CHECK(cudaMemcpyAsync(dL, l2.data(), n * sizeof(int), cudaMemcpyHostToDevice, st));
  GLB_CUDA_CHECK(cudaMemcpyAsync(dM, m2.data(), n * sizeof(int), cudaMemcpyHostToDevice, st));
  GLB_CUDA_CHECK(cudaMemcpyAsync(dO, o2.data(), n * sizeof(int64_t), cudaMemcpyHostToDevice, st));
  if (ns) {
    int maxM = 0;
    for (size_t i = 0; i < ns; ++i) maxM = std::max(maxM, m2[i]);
    const size_t smem = (size_t)(maxM + 8) * sizeof(double2);
    GLB_CUDA_CHECK(cudaFuncSetAttribute(bluestein_spectrum_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)smem));
    bluestein_spectrum_kernel<512><<<(unsigned)ns, 512, smem, st>>>(dL, dM, dO, pl->d_tw, pl->tw_n, pl->d_bf);
    GLB_CUDA_CHECK(cudaGetLastError());
  }
  if (n > ns) {  // long lengths: one global work buffer per CTA, in chunks of the scratch's slots
    const int slots = 3 * ringfft_long_slots(pl->device);
    for (size_t a = ns; a < n; a += slots) {
      const unsigned cnt = (unsigned)std::min<size_t>(slots, n - a);
      bluestein_spectrum_long_kernel<<<cnt, LONG_THREADS, 0, st>>>(dL + a, dM + a, dO + a, pl->d_tw, pl->tw_n, pl->d_bf,
                                                                   pl->d_long_scratch);
      GLB_CUDA_CHECK(cudaGetLastError());
    }
  }
  GLB_CUDA_CHECK(cudaStreamSynchronize(st));
  cudaFree(dL);
  cudaFree(dM);
  cudaFree(dO);
  return GLB_OK;
}

template <int THREADS, int NREG>
static int launch_class(FftParams p, int nrings, int nb, int lb, cudaStream_t st) {
  if (nrings == 0) return GLB_OK;
  // FFT buffer (+8: the swizzle permutes inside aligned groups of 8) and the Bluestein stash (L <= lb/2)
  p.stash_off = lb + 8;
  const size_t smem = (size_t)(lb + 8 + lb / 2) * sizeof(double2);
  static bool attr_set = false;
  if (!attr_set) {
    GLB_CUDA_CHECK(cudaFuncSetAttribute(sht_ringfft_synth_kernel<THREADS, NREG>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  dim3 grid((unsigned)nrings, (unsigned)nb);
  sht_ringfft_synth_kernel<THREADS, NREG><<<grid, THREADS, smem, st>>>(p);
  GLB_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return GLB_OK;
}

// phase [nb][nring][mmax+1] -> map [nb][npix]; nb <= 4
int sht_phase2map_group(glb_plan* pl, const double2* d_phase, int nb, double* const* d_maps, const int* kind,
                        const double* tparams, const int* d_mlim, cudaStream_t st, bool dist) {
  FftParams p;
  p.rings = pl->d_rings;
  p.phase = d_phase;
  if (dist) {
    p.G = pl->dist_world;
    p.W = pl->dist_W;
    p.blk = (int64_t)pl->dist_rows_local * pl->dist_W;
    p.rowidx = pl->d_dist_rowidx;
    p.phase_map_stride = (int64_t)pl->dist_world * p.blk;
  } else {
    p.G = 1;
    p.W = pl->mmax + 1;
    p.blk = 0;
    p.rowidx = nullptr;
    p.phase_map_stride = (int64_t)pl->nring * (pl->mmax + 1);
  }
  p.mlim = d_mlim ? d_mlim : pl->d_mlim;
  p.tw = pl->d_tw;
  p.bf = pl->d_bf;
  for (int b = 0; b < 4; ++b) p.maps[b] = (b < nb) ? d_maps[b] : nullptr;
  p.mmax = pl->mmax;
  p.tw_n = pl->tw_n;
  for (int b = 0; b < 4; ++b) {
    p.kind[b] = (kind && b < nb) ? kind[b] : GLB_T_NORMAL;
    p.p0[b] = (tparams && b < nb) ? tparams[2 * b] : 0.0;
    p.p1[b] = (tparams && b < nb) ? tparams[2 * b + 1] : 1.0;
  }
  int rc;
  int* const* order = dist ? pl->d_dist_ring_order : pl->d_ring_order;
  const int* count = dist ? pl->n_dist_ring_class : pl->n_ring_class;
  if (count[3] > 0) {  // long rings: persistent grid on global work buffers
    p.order = order[3];
    p.scratch = pl->d_long_scratch;
    p.nrings = count[3];
    p.nitems = count[3] * nb;
    sht_ringfft_synth_long_kernel<<<std::min(ringfft_long_slots(pl->device), p.nitems), LONG_THREADS, 0, st>>>(p);
    GLB_CUDA_CHECK(cudaGetLastError());
    count_launch();
  }
  // largest rings first (they take longest)
  p.order = order[2];
  if ((rc = launch_class<512, 16>(p, count[2], nb, kClassLB[2], st)) != GLB_OK) return rc;
  p.order = order[1];
  if ((rc = launch_class<256, 8>(p, count[1], nb, kClassLB[1], st)) != GLB_OK) return rc;
  p.order = order[0];
  if ((rc = launch_class<64, 8>(p, count[0], nb, kClassLB[0], st)) != GLB_OK) return rc;
  return GLB_OK;
}

template <int THREADS, int NREG>
static int launch_class_ana(FftAnaParams p, int nrings, int nb, int lb, cudaStream_t st) {
  if (nrings == 0) return GLB_OK;
  p.stash_off = lb + 8;
  const size_t smem = (size_t)(lb + 8 + lb / 2) * sizeof(double2);
  static bool attr_set = false;
  if (!attr_set) {
    GLB_CUDA_CHECK(cudaFuncSetAttribute(sht_ringfft_analysis_kernel<THREADS, NREG>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  dim3 grid((unsigned)nrings, (unsigned)nb);
  sht_ringfft_analysis_kernel<THREADS, NREG><<<grid, THREADS, smem, st>>>(p);
  GLB_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return GLB_OK;
}

// maps [nb] -> weighted phases [nb][nring][mmax+1]
int sht_map2phase_group(glb_plan* pl, const double* const* d_maps, int nb, const double* d_ring_w, double2* d_phase,
                        cudaStream_t st) {
  FftAnaParams p;
  p.rings = pl->d_rings;
  p.phase = d_phase;
  p.phase_map_stride = (int64_t)pl->nring * (pl->mmax + 1);
  p.mlim = pl->d_mlim;
  p.tw = pl->d_tw;
  p.bf = pl->d_bf;
  for (int b = 0; b < 4; ++b) p.maps[b] = (b < nb) ? d_maps[b] : nullptr;
  p.ring_w = d_ring_w;
  p.norm = 4.0 * 3.14159265358979323846 / (double)pl->npix;
  p.mmax = pl->mmax;
  p.tw_n = pl->tw_n;
  int rc;
  if (pl->n_ring_class[3] > 0) {
    p.order = pl->d_ring_order[3];
    p.scratch = pl->d_long_scratch;
    p.nrings = pl->n_ring_class[3];
    p.nitems = pl->n_ring_class[3] * nb;
    sht_ringfft_analysis_long_kernel<<<std::min(ringfft_long_slots(pl->device), p.nitems), LONG_THREADS, 0, st>>>(p);
    GLB_CUDA_CHECK(cudaGetLastError());
    count_launch();
  }
  p.order = pl->d_ring_order[2];
  if ((rc = launch_class_ana<512, 16>(p, pl->n_ring_class[2], nb, kClassLB[2], st)) != GLB_OK) return rc;
  p.order = pl->d_ring_order[1];
  if ((rc = launch_class_ana<256, 8>(p, pl->n_ring_class[1], nb, kClassLB[1], st)) != GLB_OK) return rc;
  p.order = pl->d_ring_order[0];
  if ((rc = launch_class_ana<64, 8>(p, pl->n_ring_class[0], nb, kClassLB[0], st)) != GLB_OK) return rc;
  return GLB_OK;
}

}  // namespace glb
