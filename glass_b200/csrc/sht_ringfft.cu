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
//   3. DFT_h  h a power of two: radix-2 DIF in shared memory, output read bit-reversed.
//             otherwise h = 2L: two length-L DFTs (even/odd k) by Bluestein's chirp-z
//             with a power-of-two circular convolution of length M >= 2L-1
//             (forward DIF -> multiply by the precomputed chirp spectrum, stored in the
//             same bit-reversed order -> inverse DIT: no bit reversal anywhere), then
//             one radix-2 butterfly.  Intermediate values live in registers so shared
//             memory only ever holds one M-length buffer.
//   4. store  coalesced 16-byte stores of (x[2j], x[2j+1]) with the transform applied.
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
  int kind[4];
  double p0[4];
  double p1[4];
};

// ---- shared-memory power-of-two FFTs, radix-8 passes (three butterfly levels per pass held in
// registers: 5 passes instead of 13 for 8192 points, one table twiddle per thread and pass --
// the others follow by squaring and by constant 8th roots of unity).
// DIF: natural in -> bit-reversed out.  DIT: bit-reversed in -> natural out.
// sign: inverse == false -> e^{-2 pi i jk/M}, inverse == true -> e^{+2 pi i jk/M}.
__device__ __forceinline__ double2 csq(double2 a) { return make_double2(a.x * a.x - a.y * a.y, 2.0 * a.x * a.y); }
// multiply by e^{-i pi a/4} (forward) or its conjugate (inverse), a = 0..3
template <int A>
__device__ __forceinline__ double2 mul_root8(double2 v, bool inverse) {
  const double h = 0.70710678118654752440;
  if (A == 0) return v;
  if (A == 2) return inverse ? make_double2(-v.y, v.x) : make_double2(v.y, -v.x);            // -+ i
  if (A == 1) return inverse ? make_double2(h * (v.x - v.y), h * (v.x + v.y))                  // (1+i)/sqrt2
                             : make_double2(h * (v.x + v.y), h * (v.y - v.x));                 // (1-i)/sqrt2
  return inverse ? make_double2(-h * (v.x + v.y), h * (v.x - v.y))                            // (-1+i)/sqrt2
                 : make_double2(h * (v.y - v.x), -h * (v.x + v.y));                            // (-1-i)/sqrt2
}

// one DIF pass fusing K levels with half-spans s, s/2, ..., s/2^(K-1)
template <int THREADS, int K>
__device__ __forceinline__ void dif_pass(double2* x, int M, int s, const double2* __restrict__ tw, int tw_n, bool inverse) {
  constexpr int R = 1 << K;
  const int q = s >> (K - 1);
  const int lq = 31 - __clz(q);
  const int tstep = tw_n / (2 * s);
  for (int g = threadIdx.x; g < (M >> K); g += THREADS) {
    const int lo = g & (q - 1);
    const int base = ((g >> lq) << (lq + K)) + lo;
    double2 v[R];
#pragma unroll
    for (int t = 0; t < R; ++t) v[t] = x[base + t * q];
    double2 w = __ldg(&tw[lo * tstep]);  // e^{-2 pi i lo/(2s)}
    if (inverse) w.y = -w.y;
#pragma unroll
    for (int j = 0; j < K; ++j) {
      constexpr int dummy = 0;
      (void)dummy;
      const int hs = R >> (j + 1);
#pragma unroll
      for (int t = 0; t < R; ++t) {
        if ((t & hs) == 0) {
          const double2 u = v[t], z = v[t + hs];
          v[t] = cadd(u, z);
          double2 d = cmul(csub(u, z), w);
          const int a = (t & (hs - 1)) * (4 / hs);  // (t mod hs)/(2hs) turns, in units of 1/8
          if (a == 1) d = mul_root8<1>(d, inverse);
          else if (a == 2) d = mul_root8<2>(d, inverse);
          else if (a == 3) d = mul_root8<3>(d, inverse);
          v[t + hs] = d;
        }
      }
      w = csq(w);
    }
#pragma unroll
    for (int t = 0; t < R; ++t) x[base + t * q] = v[t];
  }
  __syncthreads();
}

// one DIT pass fusing K levels with half-spans s, 2s, ..., s*2^(K-1)
template <int THREADS, int K>
__device__ __forceinline__ void dit_pass(double2* x, int M, int s, const double2* __restrict__ tw, int tw_n, bool inverse) {
  constexpr int R = 1 << K;
  const int q = s;
  const int lq = 31 - __clz(q);
  const int tstep = tw_n / (q * R);  // top level: span q*R
  for (int g = threadIdx.x; g < (M >> K); g += THREADS) {
    const int lo = g & (q - 1);
    const int base = ((g >> lq) << (lq + K)) + lo;
    double2 v[R];
#pragma unroll
    for (int t = 0; t < R; ++t) v[t] = x[base + t * q];
    // lo-dependent twiddle of each level: wl[K-1] = e^{-2 pi i lo/(q R)}, wl[j] = wl[j+1]^2
    double2 wl[K];
    wl[K - 1] = __ldg(&tw[lo * tstep]);
    if (inverse) wl[K - 1].y = -wl[K - 1].y;
#pragma unroll
    for (int j = K - 2; j >= 0; --j) wl[j] = csq(wl[j + 1]);
#pragma unroll
    for (int j = 0; j < K; ++j) {
      const int hs = 1 << j;
#pragma unroll
      for (int t = 0; t < R; ++t) {
        if ((t & hs) == 0) {
          double2 z = cmul(v[t + hs], wl[j]);
          const int a = (t & (hs - 1)) * (4 / hs);
          if (a == 1) z = mul_root8<1>(z, inverse);
          else if (a == 2) z = mul_root8<2>(z, inverse);
          else if (a == 3) z = mul_root8<3>(z, inverse);
          const double2 u = v[t];
          v[t] = cadd(u, z);
          v[t + hs] = csub(u, z);
        }
      }
    }
#pragma unroll
    for (int t = 0; t < R; ++t) x[base + t * q] = v[t];
  }
  __syncthreads();
}

template <int THREADS>
__device__ __forceinline__ void fft_dif(double2* x, int M, const double2* __restrict__ tw, int tw_n, bool inverse) {
  int s = M >> 1;
  while (s >= 4) {
    dif_pass<THREADS, 3>(x, M, s, tw, tw_n, inverse);
    s >>= 3;
  }
  if (s == 2)
    dif_pass<THREADS, 2>(x, M, s, tw, tw_n, inverse);
  else if (s == 1)
    dif_pass<THREADS, 1>(x, M, s, tw, tw_n, inverse);
}

template <int THREADS>
__device__ __forceinline__ void fft_dit(double2* x, int M, const double2* __restrict__ tw, int tw_n, bool inverse) {
  int s = 1;
  while (s * 8 <= M) {
    dit_pass<THREADS, 3>(x, M, s, tw, tw_n, inverse);
    s <<= 3;
  }
  if (s * 4 == M)
    dit_pass<THREADS, 2>(x, M, s, tw, tw_n, inverse);
  else if (s * 2 == M)
    dit_pass<THREADS, 1>(x, M, s, tw, tw_n, inverse);
}

__device__ __forceinline__ double apply_transform(double x, int kind, double p0, double p1) {
  if (kind == GLB_T_LOGNORMAL) {
    x = expm1(x - p0);
    if (p1 != 1.0) x = p1 * x;
  } else if (kind == GLB_T_SQUARED_NORMAL) {
    const double d = x - p0;
    x = d * d - 1.0;
    if (p1 != 1.0) x = p1 * x;
  }
  return x;
}

// chirp c[q] = e^{i pi q^2 / L}
__device__ __forceinline__ double2 chirp(int q, int L) {
  const long long q2 = ((long long)q * q) % (2LL * L);
  return cispi((double)q2 / (double)L);
}

// Z[k] from the folded bins
__device__ __forceinline__ double2 pack_z(const double2* G, int k, int h, int n) {
  const double2 g = G[k];
  const double2 gr = cconj(G[h - k]);
  const double2 s = cadd(g, gr);
  const double2 d = csub(g, gr);
  const double2 w = cispi(2.0 * (double)k / (double)n);
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
  auto F = [&](int m) -> double2 { return Fb[(int64_t)(m % p.G) * p.blk + m / p.G]; };
  double* __restrict__ out = p.maps[b] + d.start;
  const int kind = p.kind[b];
  const double tp0 = p.p0[b], tp1 = p.p1[b];
  const double inv_n = 1.0 / (double)n;

  // ---- 1. fold with phase shift ----
  for (int k = tid; k <= h; k += THREADS) {
    double2 g = make_double2(0.0, 0.0);
    for (int m = k; m <= mlim; m += n) {
      double2 t = F(m);
      if (m == 0) t.y = 0.0;
      if (d.shifted) t = cmul(t, cispi((double)(m % (2 * n)) * inv_n));
      g = cadd(g, t);
    }
    for (int m = n - k; m <= mlim; m += n) {
      double2 t = F(m);
      if (d.shifted) t = cmul(t, cispi((double)(m % (2 * n)) * inv_n));
      g = cadd(g, cconj(t));
    }
    buf[k] = g;
  }
  __syncthreads();

  double2 reg[NREG];
  if (d.L == 0) {
    // ---- direct power-of-two path ----
#pragma unroll
    for (int t = 0; t < NREG; ++t) {
      const int k = tid + t * THREADS;
      if (k < h) reg[t] = pack_z(buf, k, h, n);
    }
    __syncthreads();
#pragma unroll
    for (int t = 0; t < NREG; ++t) {
      const int k = tid + t * THREADS;
      if (k < h) buf[k] = reg[t];
    }
    __syncthreads();
    fft_dif<THREADS>(buf, h, p.tw, p.tw_n, true);
    const int lg = 31 - __clz(h);
    for (int j = tid; j < h; j += THREADS) {
      const int jr = (lg == 0) ? 0 : (int)(__brev((unsigned)j) >> (32 - lg));
      const double2 zv = buf[jr];
      double2 o;
      o.x = apply_transform(zv.x, kind, tp0, tp1);
      o.y = apply_transform(zv.y, kind, tp0, tp1);
      *reinterpret_cast<double2*>(out + 2 * j) = o;
    }
    return;
  }

  // ---- Bluestein path: h = 2L ----
  const int L = d.L, M = d.M;
  constexpr int NQ = NREG / 2;
  double2* ae = reg;
  double2* ao = reg + NQ;
#pragma unroll
  for (int t = 0; t < NQ; ++t) {
    const int q = tid + t * THREADS;
    if (q < L) {
      const double2 c = chirp(q, L);
      ae[t] = cmul(pack_z(buf, 2 * q, h, n), c);
      ao[t] = cmul(pack_z(buf, 2 * q + 1, h, n), c);
    }
  }
  __syncthreads();
  const double2* __restrict__ bf = p.bf + d.bf_off;
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    double2* a = pass ? ao : ae;
    for (int i = tid; i < M; i += THREADS) buf[i] = make_double2(0.0, 0.0);
    __syncthreads();
#pragma unroll
    for (int t = 0; t < NQ; ++t) {
      const int q = tid + t * THREADS;
      if (q < L) buf[q] = a[t];
    }
    __syncthreads();
    fft_dif<THREADS>(buf, M, p.tw, p.tw_n, false);
    for (int i = tid; i < M; i += THREADS) buf[i] = cmul(buf[i], bf[i]);
    __syncthreads();
    fft_dit<THREADS>(buf, M, p.tw, p.tw_n, true);
#pragma unroll
    for (int t = 0; t < NQ; ++t) {
      const int q = tid + t * THREADS;
      if (q < L) a[t] = buf[q];
    }
    __syncthreads();
  }
#pragma unroll
  for (int t = 0; t < NQ; ++t) {
    const int q = tid + t * THREADS;
    if (q < L) {
      const double2 c = chirp(q, L);
      const double2 E = cmul(ae[t], c);
      const double2 O = cmul(cmul(ao[t], c), cispi(2.0 * (double)q / (double)h));
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
  const int lg = 31 - __clz(h);
  bool bitrev_out;

  if (d.L == 0) {
    for (int j = tid; j < h; j += THREADS) buf[j] = *reinterpret_cast<const double2*>(in + 2 * j);
    __syncthreads();
    fft_dif<THREADS>(buf, h, p.tw, p.tw_n, false);
    bitrev_out = true;
  } else {
    const int L = d.L, M = d.M;
    constexpr int NQ = NREG / 2;
    double2 reg[NREG];
    double2* ae = reg;
    double2* ao = reg + NQ;
#pragma unroll
    for (int t = 0; t < NQ; ++t) {
      const int q = tid + t * THREADS;
      if (q < L) {
        const double2 c = chirp(q, L);
        const double2 ze = *reinterpret_cast<const double2*>(in + 4 * q);
        const double2 zo = *reinterpret_cast<const double2*>(in + 4 * q + 2);
        ae[t] = cmul(cconj(ze), c);  // forward DFT as conj(IDFT(conj .))
        ao[t] = cmul(cconj(zo), c);
      }
    }
    const double2* __restrict__ bf = p.bf + d.bf_off;
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
      double2* a = pass ? ao : ae;
      for (int i = tid; i < M; i += THREADS) buf[i] = make_double2(0.0, 0.0);
      __syncthreads();
#pragma unroll
      for (int t = 0; t < NQ; ++t) {
        const int q = tid + t * THREADS;
        if (q < L) buf[q] = a[t];
      }
      __syncthreads();
      fft_dif<THREADS>(buf, M, p.tw, p.tw_n, false);
      for (int i = tid; i < M; i += THREADS) buf[i] = cmul(buf[i], bf[i]);
      __syncthreads();
      fft_dit<THREADS>(buf, M, p.tw, p.tw_n, true);
#pragma unroll
      for (int t = 0; t < NQ; ++t) {
        const int q = tid + t * THREADS;
        if (q < L) a[t] = buf[q];
      }
      __syncthreads();
    }
    // Zf[k] = E[k] + w^k O[k], Zf[k+L] = E[k] - w^k O[k], w = e^{-2 pi i / h}; natural order in buf
#pragma unroll
    for (int t = 0; t < NQ; ++t) {
      const int q = tid + t * THREADS;
      if (q < L) {
        const double2 c = chirp(q, L);
        const double2 E = cconj(cmul(ae[t], c));
        const double2 O = cmul(cconj(cmul(ao[t], c)), cispi(-2.0 * (double)q / (double)h));
        buf[q] = cadd(E, O);
        buf[q + L] = csub(E, O);
      }
    }
    __syncthreads();
    bitrev_out = false;
  }

  const double inv_n = 1.0 / (double)n;
  for (int m = tid; m <= mlim; m += THREADS) {
    const int k = m % n;
    const int kk = (k <= h) ? k : n - k;
    int i0 = kk % h, i1 = (h - kk) % h;
    if (bitrev_out && lg > 0) {
      i0 = (int)(__brev((unsigned)i0) >> (32 - lg));
      i1 = (int)(__brev((unsigned)i1) >> (32 - lg));
    }
    const double2 zk = buf[i0];
    const double2 zr = cconj(buf[i1]);
    const double2 sum = cadd(zk, zr), dif = csub(zk, zr);
    const double2 wd = cmul(cispi(-2.0 * (double)kk * inv_n), dif);  // -i*wd = (wd.y, -wd.x)
    double2 X = make_double2(0.5 * (sum.x + wd.y), 0.5 * (sum.y - wd.x));
    if (k > h) X = cconj(X);
    if (d.shifted) X = cmul(X, cispi(-(double)(m % (2 * n)) * inv_n));
    F[m] = cscale(X, wscale);
  }
}

// chirp spectrum  Bf = DIF_M( b_wrapped ) / M,  b[d] = conj(c[d]) = e^{-i pi d^2 / L}
template <int THREADS>
__global__ void __launch_bounds__(THREADS) bluestein_spectrum_kernel(const int* Ls, const int* Ms, const int64_t* offs,
                                                                     const double2* tw, int tw_n, double2* bf) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double2* buf = reinterpret_cast<double2*>(smem_raw);
  const int tid = threadIdx.x;
  const int L = Ls[blockIdx.x], M = Ms[blockIdx.x];
  for (int i = tid; i < M; i += THREADS) buf[i] = make_double2(0.0, 0.0);
  __syncthreads();
  for (int q = tid; q < L; q += THREADS) {
    const double2 c = cconj(chirp(q, L));
    buf[q] = c;
    if (q > 0) buf[M - q] = c;
  }
  __syncthreads();
  fft_dif<THREADS>(buf, M, tw, tw_n, false);
  const double inv = 1.0 / (double)M;
  double2* o = bf + offs[blockIdx.x];
  for (int i = tid; i < M; i += THREADS) o[i] = cscale(buf[i], inv);
}

// -------------------------------------------------------------------------------------
// host side
// -------------------------------------------------------------------------------------
static const int kClassLB[3] = {512, 2048, 8192};

int ringfft_class_of(int lbuf) {
  for (int c = 0; c < 3; ++c)
    if (lbuf <= kClassLB[c]) return c;
  return -1;
}

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
  int maxM = 0;
  for (int m : Ms) maxM = std::max(maxM, m);
  const size_t smem = (size_t)maxM * sizeof(double2);
  GLB_CUDA_CHECK(cudaFuncSetAttribute(bluestein_spectrum_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
  bluestein_spectrum_kernel<512><<<(unsigned)n, 512, smem, st>>>(dL, dM, dO, pl->d_tw, pl->tw_n, pl->d_bf);
  GLB_CUDA_CHECK(cudaGetLastError());
  GLB_CUDA_CHECK(cudaStreamSynchronize(st));
  cudaFree(dL);
  cudaFree(dM);
  cudaFree(dO);
  return GLB_OK;
}

template <int THREADS, int NREG>
static int launch_class(const FftParams& p, int nrings, int nb, int lb, cudaStream_t st) {
  if (nrings == 0) return GLB_OK;
  const size_t smem = (size_t)(lb + 2) * sizeof(double2);
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
static int launch_class_ana(const FftAnaParams& p, int nrings, int nb, int lb, cudaStream_t st) {
  if (nrings == 0) return GLB_OK;
  const size_t smem = (size_t)(lb + 2) * sizeof(double2);
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
  p.order = pl->d_ring_order[2];
  if ((rc = launch_class_ana<512, 16>(p, pl->n_ring_class[2], nb, kClassLB[2], st)) != GLB_OK) return rc;
  p.order = pl->d_ring_order[1];
  if ((rc = launch_class_ana<256, 8>(p, pl->n_ring_class[1], nb, kClassLB[1], st)) != GLB_OK) return rc;
  p.order = pl->d_ring_order[0];
  if ((rc = launch_class_ana<64, 8>(p, pl->n_ring_class[0], nb, kClassLB[0], st)) != GLB_OK) return rc;
  return GLB_OK;
}

}  // namespace glb
