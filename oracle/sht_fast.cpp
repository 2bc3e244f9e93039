/*
 * oracle/sht_fast.cpp -- the TIMED CPU arm of bench.py: scalar HEALPix synthesis
 * (healpy.alm2map(pol=False), glass/healpix.py:71) written the way a production CPU library
 * is -- SIMD across rings, register-blocked Legendre recurrence, OpenMP over m, real FFT of
 * half length -- so that the CPU baseline is not flattered by a naive port.
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): bench.py's `cpu_baseline` and
 * `--impl reference` legs time it; tests/ pin it against oracle/sht_ref.c (the checker,
 * which stays the simple scalar restatement).  Nothing under glass_b200/ uses it.
 *
 * healpy / libsharp2 are absent from /root/reference and from this image; this follows the
 * published structure of libsharp (Reinecke & Seljebotn 2013): per m, blocks of rings held in
 * vector registers walk the standard three-term recurrence in l with a per-ring power-of-two
 * scale until the values become significant, then a branch-free FMA loop; north/south rings
 * share the recurrence by l-parity; rings whose mlim is below m are skipped; per ring one
 * complex FFT of length nphi/2 (power of two directly, Bluestein otherwise) gives the pixels.
 */
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef double v8 __attribute__((vector_size(64)));
typedef long long v8l __attribute__((vector_size(64)));
typedef std::complex<double> cplx;

namespace {

constexpr int VL = 8;        // doubles per vector
constexpr int K = 4;         // vectors per register block (measured best of 2..5 at nside 2048)
constexpr int NBLK = VL * K; // ring pairs per block
constexpr int SCALE_BITS = 400;

struct Ring {
  int64_t start;
  int nphi;
  bool shifted;
  double z, sth;
};

Ring ring_geom(int nside, int r) {
  const int64_t N = nside, i = r + 1;
  Ring g;
  if (i < N) {
    const double t = double(i * i) / double(3 * N * N);
    g = {2 * i * (i - 1), int(4 * i), true, 1 - t, std::sqrt(t * (2 - t))};
  } else if (i <= 3 * N) {
    const double z = double(2 * N - i) * 2 / double(3 * N);
    g = {2 * N * (N - 1) + (i - N) * 4 * N, int(4 * N), (i - N) % 2 == 0, z, std::sqrt((1 - z) * (1 + z))};
  } else {
    const int64_t ip = 4 * N - i;
    const double t = double(ip * ip) / double(3 * N * N);
    g = {12 * N * N - 2 * ip * (ip + 1), int(4 * ip), true, -(1 - t), std::sqrt(t * (2 - t))};
  }
  return g;
}

int mlim_of(int lmax, double sth) {
  const double ofs = std::max(100., lmax * 0.01);
  return int(std::min(double(lmax), lmax * sth + ofs) + 0.5);
}

inline v8 vabs(v8 x) { return (v8)((v8l)x & 0x7fffffffffffffffLL); }
inline bool any(v8l m) {
  long long r = 0;
  for (int i = 0; i < VL; ++i) r |= m[i];
  return r != 0;
}
inline bool all(v8l m) {
  long long r = -1;
  for (int i = 0; i < VL; ++i) r &= m[i];
  return r != 0;
}

/* Row m of F holds 4 nside slots: north ring r (incl. the equator) at r, its mirror at
 * 4 nside - 1 - r, so that the mirrors of an aligned group of four north rings are an aligned
 * group of four slots (one cache line each in the FFT stage). */

/* ---- Legendre stage: F[m][slot] = sum_l a_lm lambda_lm(z_ring) -------------------------------- */
void legendre(int nside, int lmax, const cplx* alm, cplx* F) {
  const int npair = 2 * nside;
  const int64_t stride = 4 * (int64_t)nside;
  std::vector<double> zs(npair + NBLK, 0.0), logsth(npair + NBLK, 0.0);
  std::vector<int> mlim(npair);
  for (int r = 0; r < npair; ++r) {
    const Ring g = ring_geom(nside, r);
    zs[r] = g.z;
    logsth[r] = std::log2(g.sth);
    mlim[r] = mlim_of(lmax, g.sth);
  }
  std::vector<double> log2c(lmax + 1);
  log2c[0] = -0.5 * std::log2(4 * M_PI);
  for (int m = 1; m <= lmax; ++m) log2c[m] = log2c[m - 1] + 0.5 * std::log2(double(2 * m + 1) / double(2 * m));
  const double BIG = std::ldexp(1.0, SCALE_BITS / 2), SMALL = std::ldexp(1.0, -SCALE_BITS);

#pragma omp parallel
  {
    std::vector<double> a(lmax + 4, 0.0), c(lmax + 4, 0.0), ar(lmax + 4, 0.0), ai(lmax + 4, 0.0);
#pragma omp for schedule(dynamic, 1)
    for (int mm = 0; mm <= lmax; ++mm) {
      /* interleave small and large m: the cost of an m falls with m */
      const int m = (mm & 1) ? lmax - mm / 2 : mm / 2;
      const cplx* a_m = alm + ((int64_t)m * (2 * lmax + 1 - m)) / 2; /* indexed by l */
      for (int l = m; l <= lmax; ++l) {
        ar[l] = a_m[l].real();
        ai[l] = m == 0 ? 0.0 : a_m[l].imag();
      }
      ar[lmax + 1] = ai[lmax + 1] = ar[lmax + 2] = ai[lmax + 2] = 0.0;
      /* lambda_l = a_l z lambda_{l-1} - c_l lambda_{l-2},  a_l = A_l,  c_l = A_l / A_{l-1} */
      double prev = 0;
      for (int l = m + 1; l <= lmax + 2; ++l) {
        const double l2 = double(l) * l, m2 = double(m) * m;
        const double A = std::sqrt((4 * l2 - 1) / (l2 - m2));
        a[l] = A;
        c[l] = l == m + 1 ? 0.0 : A / prev;
        prev = A;
      }
      int rmin = 0;
      while (rmin < npair && mlim[rmin] < m) ++rmin;
      cplx* Fm = F + (int64_t)m * stride;
      for (int r0 = rmin; r0 < npair; r0 += NBLK) {
        v8 z[K], l1[K], l2[K], sc[K], er[K], ei[K], orr[K], oi[K];
        for (int k = 0; k < K; ++k)
          for (int i = 0; i < VL; ++i) {
            /* lanes past the equator repeat it (same phase changes as their neighbours; not stored) */
            const int r = std::min(r0 + k * VL + i, npair - 1);
            z[k][i] = zs[r];
            double s = 0, lg2 = log2c[m] + m * logsth[r];
            if (lg2 < -SCALE_BITS / 2) {
              s = std::trunc(lg2 / SCALE_BITS);
              lg2 -= s * SCALE_BITS;
            }
            const double seed = std::exp2(lg2) * ((m & 1) ? -1.0 : 1.0);
            l2[k][i] = seed;
            sc[k][i] = s;
            l1[k][i] = 0;
            er[k][i] = ei[k][i] = orr[k][i] = oi[k][i] = 0;
          }
        int l = m;
        /* SKIP: every ring of the block still scaled down -- recurrence and rescaling only */
        for (;;) {
          bool live = false;
          for (int k = 0; k < K; ++k) live |= any(sc[k] == 0.0);
          if (live || l > lmax) break;
          for (int k = 0; k < K; ++k) {
            v8 t = a[l + 1] * (z[k] * l2[k]) - c[l + 1] * l1[k];
            l1[k] = l2[k];
            l2[k] = t;
            const v8l big = (vabs(t) > BIG) & (sc[k] < 0.0);
            l1[k] = big ? l1[k] * SMALL : l1[k];
            l2[k] = big ? l2[k] * SMALL : l2[k];
            sc[k] = big ? sc[k] + 1.0 : sc[k];
          }
          ++l;
        }
        /* MIXED: some rings significant -- masked accumulate; also aligns l to even l - m */
        for (;;) {
          bool done = true;
          for (int k = 0; k < K; ++k) done &= all(sc[k] == 0.0);
          if ((done && ((l - m) & 1) == 0) || l > lmax) break;
          const bool odd = (l - m) & 1;
          for (int k = 0; k < K; ++k) {
            const v8 zero = {0, 0, 0, 0, 0, 0, 0, 0};
            const v8 v = (sc[k] == 0.0) ? l2[k] : zero;
            if (odd) {
              orr[k] += ar[l] * v;
              oi[k] += ai[l] * v;
            } else {
              er[k] += ar[l] * v;
              ei[k] += ai[l] * v;
            }
            v8 t = a[l + 1] * (z[k] * l2[k]) - c[l + 1] * l1[k];
            l1[k] = l2[k];
            l2[k] = t;
            const v8l big = (vabs(t) > BIG) & (sc[k] < 0.0);
            l1[k] = big ? l1[k] * SMALL : l1[k];
            l2[k] = big ? l2[k] * SMALL : l2[k];
            sc[k] = big ? sc[k] + 1.0 : sc[k];
          }
          ++l;
        }
        /* FAST: two l per round, everything in registers; a, c, ar, ai are padded past lmax */
        for (; l <= lmax; l += 2) {
          const double a1 = a[l + 1], c1 = c[l + 1], a2 = a[l + 2], c2 = c[l + 2];
          const double re0 = ar[l], im0 = ai[l], re1 = ar[l + 1], im1 = ai[l + 1];
          for (int k = 0; k < K; ++k) {
            er[k] += re0 * l2[k];
            ei[k] += im0 * l2[k];
            const v8 t = a1 * (z[k] * l2[k]) - c1 * l1[k];
            orr[k] += re1 * t;
            oi[k] += im1 * t;
            l1[k] = t;
            l2[k] = a2 * (z[k] * t) - c2 * l2[k];
          }
        }
        for (int k = 0; k < K; ++k)
          for (int i = 0; i < VL; ++i) {
            const int r = r0 + k * VL + i;
            if (r >= npair) break;
            Fm[r] = cplx(er[k][i] + orr[k][i], ei[k][i] + oi[k][i]);
            if (r != npair - 1) Fm[stride - 1 - r] = cplx(er[k][i] - orr[k][i], ei[k][i] - oi[k][i]);
          }
      }
    }
  }
}

/* ---- FFT stage ------------------------------------------------------------------------------- */
struct Twiddles {  // per butterfly level len = 2, 4, ..., nmax: e^{2 pi i j / len}, j < len / 2, contiguous
  int nmax;
  std::vector<double> re, im;  // level `len` starts at offset len / 2 - 1
  explicit Twiddles(int n) : nmax(n), re(n), im(n) {
    for (int len = 2; len <= n; len <<= 1)
      for (int j = 0; j < len / 2; ++j) {
        re[len / 2 - 1 + j] = std::cos(2 * M_PI * j / len);
        im[len / 2 - 1 + j] = std::sin(2 * M_PI * j / len);
      }
  }
};

/* unnormalised power-of-two DFT, sign = +1 (e^{+i}) or -1, in place */
void fft_pow2(cplx* x, int n, int sign, const Twiddles& tw) {
  for (int i = 1, j = 0; i < n; ++i) {
    int bit = n >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) std::swap(x[i], x[j]);
  }
  const double sg = sign;
  for (int len = 2; len <= n; len <<= 1) {
    const int half = len >> 1;
    const double* wr = tw.re.data() + half - 1;
    const double* wi = tw.im.data() + half - 1;
    for (int i = 0; i < n; i += len) {
      double* p = reinterpret_cast<double*>(x + i);
      double* q = reinterpret_cast<double*>(x + i + half);
#pragma omp simd
      for (int j = 0; j < half; ++j) {
        const double cr = wr[j], ci = sg * wi[j];
        const double vr = q[2 * j] * cr - q[2 * j + 1] * ci, vi = q[2 * j] * ci + q[2 * j + 1] * cr;
        const double ur = p[2 * j], ui = p[2 * j + 1];
        p[2 * j] = ur + vr;
        p[2 * j + 1] = ui + vi;
        q[2 * j] = ur - vr;
        q[2 * j + 1] = ui - vi;
      }
    }
  }
}

struct Bluestein {  // y[j] = sum_k x[k] e^{+2 pi i jk/h}, any h; chirp and its spectrum are per length
  int h = 0, M = 0;
  std::vector<cplx> chirp, bspec, work;
  void prepare(int h_, const Twiddles& tw) {
    if (h_ == h) return;
    h = h_;
    M = 1;
    while (M < 2 * h - 1) M <<= 1;
    chirp.resize(h);
    bspec.assign(M, cplx(0, 0));
    work.resize(M);
    for (int t = 0; t < h; ++t) {
      const long long t2 = ((long long)t * t) % (2LL * h);
      chirp[t] = std::polar(1.0, M_PI * double(t2) / double(h));
    }
    bspec[0] = std::conj(chirp[0]);
    for (int t = 1; t < h; ++t) bspec[t] = bspec[M - t] = std::conj(chirp[t]);
    fft_pow2(bspec.data(), M, -1, tw);
    const double inv = 1.0 / M;
    for (int t = 0; t < M; ++t) bspec[t] *= inv;
  }
  void run(cplx* x, const Twiddles& tw) {
    for (int t = 0; t < h; ++t) work[t] = x[t] * chirp[t];
    std::fill(work.begin() + h, work.end(), cplx(0, 0));
    fft_pow2(work.data(), M, -1, tw);
    for (int t = 0; t < M; ++t) work[t] *= bspec[t];
    fft_pow2(work.data(), M, +1, tw);
    for (int t = 0; t < h; ++t) x[t] = work[t] * chirp[t];
  }
};

/* one ring: half spectrum X[0..h] (h = nphi/2) -> nphi real pixels through a complex DFT of length h */
void ring_c2r(cplx* X, int n, double* out, cplx* Z, Bluestein& bl, const Twiddles& tw) {
  const int h = n / 2;
  const cplx w1 = std::polar(1.0, 2 * M_PI / n);
  cplx w(1, 0);
  for (int k = 0; k < h; ++k) {
    if ((k & 31) == 0) w = std::polar(1.0, 2 * M_PI * k / n);  // bounds the drift of the rotation
    const cplx xk = X[k], xc = std::conj(X[h - k]);
    const cplx E = xk + xc, O = (xk - xc) * w;
    Z[k] = E + cplx(-O.imag(), O.real());
    w *= w1;
  }
  if ((h & (h - 1)) == 0)
    fft_pow2(Z, h, +1, tw);
  else {
    bl.prepare(h, tw);
    bl.run(Z, tw);
  }
  std::memcpy(out, Z, sizeof(double) * n);  // z[j] = x[2j] + i x[2j+1]
}

void phases_to_map(int nside, int lmax, const cplx* F, double* map) {
  const int npair = 2 * nside, nring = 4 * nside - 1;
  const int64_t stride = 4 * (int64_t)nside;
  int mfft = 1;
  while (mfft < 4 * nside) mfft <<= 1;  // covers Bluestein M >= 2 (2 nside) - 1 and direct n/2
  const Twiddles tw(mfft);
#pragma omp parallel
  {
    std::vector<cplx> X(8 * (2 * nside + 1)), Z(2 * nside);
    Bluestein bl;
#pragma omp for schedule(dynamic, 1)
    for (int g = 0; g < npair / 4; ++g) {
      /* four north rings r0..r0+3 and their mirrors: one cache line of F per m and hemisphere */
      const int r0 = 4 * g;
      Ring rg[4];
      int ml[4], hh[4], mmax = 0;
      cplx w1[4], w[4];
      for (int i = 0; i < 4; ++i) {
        rg[i] = ring_geom(nside, r0 + i);
        ml[i] = std::min(lmax, mlim_of(lmax, rg[i].sth));
        mmax = std::max(mmax, ml[i]);
        hh[i] = rg[i].nphi / 2;
        w1[i] = rg[i].shifted ? std::polar(1.0, M_PI / rg[i].nphi) : cplx(1, 0);
        w[i] = cplx(1, 0);
      }
      const int xs = 2 * nside + 1;
      std::fill(X.begin(), X.end(), cplx(0, 0));
      for (int m = 0; m <= mmax; ++m) {
        const cplx* fn = F + (int64_t)m * stride + r0;
        const cplx* fs = F + (int64_t)m * stride + (stride - 4 - r0);
        for (int i = 0; i < 4; ++i) {
          if (m > ml[i]) continue;
          if ((m & 63) == 0 && rg[i].shifted) w[i] = std::polar(1.0, M_PI * double(m % (2 * rg[i].nphi)) / rg[i].nphi);
          const int n = rg[i].nphi, h = hh[i], k = m % n;
          cplx tn = fn[i] * w[i], ts = fs[3 - i] * w[i];
          w[i] *= w1[i];
          cplx* xn = &X[(2 * i) * xs];
          cplx* xsou = &X[(2 * i + 1) * xs];
          if (m == 0) {
            xn[0] += tn.real();
            xsou[0] += ts.real();
          } else if (k == 0 || k == h) {
            xn[k] += 2 * tn.real();
            xsou[k] += 2 * ts.real();
          } else if (k < h) {
            xn[k] += tn;
            xsou[k] += ts;
          } else {
            xn[n - k] += std::conj(tn);
            xsou[n - k] += std::conj(ts);
          }
        }
      }
      for (int i = 0; i < 4; ++i) {
        const int r = r0 + i;
        ring_c2r(&X[(2 * i) * xs], rg[i].nphi, map + rg[i].start, Z.data(), bl, tw);
        if (r != npair - 1) {
          const Ring s = ring_geom(nside, nring - 1 - r);
          ring_c2r(&X[(2 * i + 1) * xs], s.nphi, map + s.start, Z.data(), bl, tw);
        }
      }
    }
  }
}

}  // namespace

extern "C" {

/* healpy.alm2map(alm, nside, pol=False, pixwin=False) with libsharp's mlim ring skipping */
int fast_alm2map(int nside, int lmax, const double* alm, double* map, int nthreads) {
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
  if (nside < 2 || (nside & (nside - 1))) return -2;
  const size_t nF = (size_t)(lmax + 1) * 4 * nside;
  cplx* F = static_cast<cplx*>(std::calloc(nF, sizeof(cplx)));
  if (!F) return -1;
  legendre(nside, lmax, reinterpret_cast<const cplx*>(alm), F);
  phases_to_map(nside, lmax, F, map);
  std::free(F);
  return 0;
}

/* seconds spent in the two stages of one call (for the bench's `sample` note) */
int fast_alm2map_timed(int nside, int lmax, const double* alm, double* map, int nthreads, double* t_legendre, double* t_fft) {
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
  if (nside < 2 || (nside & (nside - 1))) return -2;
  const size_t nF = (size_t)(lmax + 1) * 4 * nside;
  cplx* F = static_cast<cplx*>(std::calloc(nF, sizeof(cplx)));
  if (!F) return -1;
  const double t0 = omp_get_wtime();
  legendre(nside, lmax, reinterpret_cast<const cplx*>(alm), F);
  const double t1 = omp_get_wtime();
  phases_to_map(nside, lmax, F, map);
  const double t2 = omp_get_wtime();
  *t_legendre = t1 - t0;
  *t_fft = t2 - t1;
  std::free(F);
  return 0;
#else
  *t_legendre = *t_fft = 0;
  return fast_alm2map(nside, lmax, alm, map, nthreads);
#endif
}
}
