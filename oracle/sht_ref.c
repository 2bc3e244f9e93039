/*
 * oracle/sht_ref.c -- plain-C CPU restatement of the HEALPix ring spherical-harmonic
 * transforms that the reference reaches through healpy (glass/healpix.py:71,107,270).
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): used by tests/ as the checker at
 * sizes the NumPy oracle is too slow for, and by bench.py as the timed CPU arm.
 *
 * healpy / libsharp2 are absent from /root/reference and from this image, so this follows
 * the published algorithm (Gorski et al. 2005 ring geometry; Reinecke & Seljebotn 2013
 * libsharp: per-m Legendre recurrence with range scaling, per-ring FFT with alias
 * folding).  It deliberately uses the STANDARD three-term recurrence
 * in l -- not the x^2 recurrence of the CUDA kernels -- so the two implementations are
 * independent.  Pinned against scipy.special.sph_harm_y direct sums through
 * oracle/healpix_ref.py (tests/test_oracle_sht.py).  Parity with healpy itself: unpinned.
 *
 * Build (oracle/Makefile):  REAL=double   -> liboracle_sht.so     (OpenMP, baseline + checker)
 *                           REAL=long double -> liboracle_sht_ld.so (80-bit truth)
 */
#include <complex.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifdef USE_LONG_DOUBLE
typedef long double real;
#define R_SQRT sqrtl
#define R_LOG logl
#define R_EXP expl
#define R_FABS fabsl
#define R_LDEXP ldexpl
#else
typedef double real;
#define R_SQRT sqrt
#define R_LOG log
#define R_EXP exp
#define R_FABS fabs
#define R_LDEXP ldexp
#endif

#define SCALE_BITS 400

/* ---- ring geometry, rings 1..4N-1, index r = 0..4N-2 ---------------- */
typedef struct {
  int64_t start;
  int nphi;
  int shifted;
  real z, sth;
} ring_t;

static void ring_geom(int nside, int r, ring_t* g) {
  const int64_t N = nside;
  const int64_t i = r + 1;
  if (i < N) {
    const real t = (real)(i * i) / (real)(3 * N * N);
    g->nphi = (int)(4 * i);
    g->z = 1 - t;
    g->sth = R_SQRT(t * (2 - t));
    g->shifted = 1;
    g->start = 2 * i * (i - 1);
  } else if (i <= 3 * N) {
    g->nphi = (int)(4 * N);
    g->z = (real)(2 * N - i) * 2 / (real)(3 * N);
    g->sth = R_SQRT((1 - g->z) * (1 + g->z));
    g->shifted = ((i - N) % 2 == 0);
    g->start = 2 * N * (N - 1) + (i - N) * 4 * N;
  } else {
    const int64_t ip = 4 * N - i;
    const real t = (real)(ip * ip) / (real)(3 * N * N);
    g->nphi = (int)(4 * ip);
    g->z = -(1 - t);
    g->sth = R_SQRT(t * (2 - t));
    g->shifted = 1;
    g->start = 12 * N * N - 2 * ip * (ip + 1);
  }
}

static int mlim_of(int lmax, double sth) {
  double ofs = lmax * 0.01;
  if (ofs < 100.) ofs = 100.;
  double res = lmax * sth + ofs;
  if (res > lmax) res = lmax;
  return (int)(res + 0.5);
}

/* ---- FFT: radix-2 for powers of two, Bluestein otherwise ---------------------------- */
static void fft_pow2(double complex* x, int n, int sign) {
  for (int i = 1, j = 0; i < n; ++i) {
    int bit = n >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) {
      double complex t = x[i];
      x[i] = x[j];
      x[j] = t;
    }
  }
  for (int len = 2; len <= n; len <<= 1) {
    const int half = len >> 1;
    for (int j = 0; j < half; ++j) {
      const double ang = sign * 2.0 * M_PI * (double)j / (double)len;
      const double complex w = cos(ang) + I * sin(ang);
      for (int i = j; i < n; i += len) {
        const double complex u = x[i], v = x[i + half] * w;
        x[i] = u + v;
        x[i + half] = u - v;
      }
    }
  }
}

/* y[j] = sum_k x[k] exp(sign 2 pi i jk/n), any n, in place; work arrays allocated here */
static void dft_any(double complex* x, int n, int sign) {
  if ((n & (n - 1)) == 0) {
    fft_pow2(x, n, sign);
    return;
  }
  int M = 1;
  while (M < 2 * n - 1) M <<= 1;
  double complex* a = (double complex*)calloc((size_t)M, sizeof(double complex));
  double complex* b = (double complex*)calloc((size_t)M, sizeof(double complex));
  double complex* c = (double complex*)malloc((size_t)n * sizeof(double complex));
  for (int t = 0; t < n; ++t) {
    const long long t2 = ((long long)t * t) % (2LL * n);
    const double ang = sign * M_PI * (double)t2 / (double)n;
    c[t] = cos(ang) + I * sin(ang);
  }
  for (int t = 0; t < n; ++t) a[t] = x[t] * c[t];
  b[0] = conj(c[0]);
  for (int t = 1; t < n; ++t) b[t] = b[M - t] = conj(c[t]);
  fft_pow2(a, M, -1);
  fft_pow2(b, M, -1);
  for (int t = 0; t < M; ++t) a[t] *= b[t];
  fft_pow2(a, M, +1);
  for (int t = 0; t < n; ++t) x[t] = a[t] * c[t] / (double)M;
  free(a);
  free(b);
  free(c);
}

/* ---- Legendre: F[ring][m] = sum_l a_lm lambda_lm(z_ring), standard recurrence --------
 * lambda_l = A_l (z lambda_{l-1} - B_l lambda_{l-2}),  A_l = sqrt((4l^2-1)/(l^2-m^2)),
 * B_l = 1/A_{l-1};  lambda_mm = (-1)^m c_m sin^m(theta), seeded in log space and carried as
 * value * 2^(SCALE_BITS*scale) so that it never underflows.                                */
static void legendre_synth(int nside, int lmax, const double complex* alm, double complex* F, int use_mlim) {
  const int nring = 4 * nside - 1, npair = 2 * nside;
  ring_t* rg = (ring_t*)malloc((size_t)npair * sizeof(ring_t));
  for (int r = 0; r < npair; ++r) ring_geom(nside, r, &rg[r]);
  /* log c_m */
  real* logc = (real*)malloc((size_t)(lmax + 1) * sizeof(real));
  logc[0] = -0.5 * R_LOG((real)4 * (real)3.14159265358979323846264338327950288L);
  for (int m = 1; m <= lmax; ++m) logc[m] = logc[m - 1] + 0.5 * R_LOG((real)(2 * m + 1) / (real)(2 * m));
  const real ln2 = R_LOG((real)2);
  const real BIG = R_LDEXP((real)1, SCALE_BITS / 2), SMALL = R_LDEXP((real)1, -SCALE_BITS);

#pragma omp parallel
  {
    real* A = (real*)malloc((size_t)(lmax + 2) * sizeof(real));
    real* B = (real*)malloc((size_t)(lmax + 2) * sizeof(real));
#pragma omp for schedule(dynamic, 1)
    for (int m = 0; m <= lmax; ++m) {
      const double complex* a_m = alm + ((int64_t)m * (2 * lmax + 1 - m)) / 2; /* index by l */
      for (int l = m + 1; l <= lmax; ++l) {
        const real l2 = (real)l * l, m2 = (real)m * m;
        A[l] = R_SQRT((4 * l2 - 1) / (l2 - m2));
      }
      for (int l = m + 2; l <= lmax; ++l) B[l] = 1 / A[l - 1];
      for (int r = 0; r < npair; ++r) {
        if (use_mlim && mlim_of(lmax, (double)rg[r].sth) < m) continue;
        const real z = rg[r].z;
        /* seed */
        real lg2 = (logc[m] + m * R_LOG(rg[r].sth)) / ln2; /* log2 |lambda_mm| */
        int scale = 0;
        if (lg2 < -SCALE_BITS / 2) {
          scale = (int)(lg2 / SCALE_BITS); /* negative, truncation */
          lg2 -= (real)scale * SCALE_BITS;
        }
        real lam1 = 0, lam2 = R_EXP(lg2 * ln2) * ((m & 1) ? -1 : 1);
        real fer = 0, fei = 0, forr = 0, foi = 0;
        for (int l = m; l <= lmax; ++l) {
          if (l > m) {
            const real t = (l == m + 1) ? A[l] * z * lam2 : A[l] * (z * lam2 - B[l] * lam1);
            lam1 = lam2;
            lam2 = t;
            if (scale < 0 && R_FABS(lam2) > BIG) {
              lam1 *= SMALL;
              lam2 *= SMALL;
              ++scale;
            }
          }
          if (scale == 0) {
            const real ar = (real)creal(a_m[l]), ai = (m == 0) ? 0 : (real)cimag(a_m[l]);
            if ((l - m) & 1) {
              forr += ar * lam2;
              foi += ai * lam2;
            } else {
              fer += ar * lam2;
              fei += ai * lam2;
            }
          }
        }
        F[(int64_t)r * (lmax + 1) + m] = (double)(fer + forr) + I * (double)(fei + foi);
        if (r != npair - 1) F[(int64_t)(nring - 1 - r) * (lmax + 1) + m] = (double)(fer - forr) + I * (double)(fei - foi);
      }
    }
    free(A);
    free(B);
  }
  free(rg);
  free(logc);
}

static void phases_to_map(int nside, int lmax, const double complex* F, double* map, int use_mlim) {
  const int nring = 4 * nside - 1;
#pragma omp parallel
  {
    double complex* x = (double complex*)malloc((size_t)(4 * nside) * sizeof(double complex));
#pragma omp for schedule(dynamic, 4)
    for (int r = 0; r < nring; ++r) {
      ring_t g;
      ring_geom(nside, r, &g);
      const int n = g.nphi;
      const double phi0 = g.shifted ? M_PI / n : 0.0;
      int mmax = lmax;
      if (use_mlim) {
        const int ml = mlim_of(lmax, (double)g.sth);
        if (ml < mmax) mmax = ml;
      }
      memset(x, 0, (size_t)n * sizeof(double complex));
      const double complex* Fr = F + (int64_t)r * (lmax + 1);
      x[0] += creal(Fr[0]);
      for (int m = 1; m <= mmax; ++m) {
        const double ang = phi0 * (double)(m % (2 * n));
        const double complex t = Fr[m] * (cos(ang) + I * sin(ang));
        const int k = m % n;
        /* full Hermitian spectrum: bin k gets t, bin n-k gets conj(t) */
        x[k] += t;
        x[(n - k) % n] += conj(t);
      }
      dft_any(x, n, +1);
      for (int j = 0; j < n; ++j) map[g.start + j] = creal(x[j]);
    }
    free(x);
  }
}

/* healpy.alm2map(alm, nside, pol=False, pixwin=False)  (glass/healpix.py:71) */
int ref_alm2map(int nside, int lmax, const double* alm, double* map, int use_mlim, int nthreads) {
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
  const int nring = 4 * nside - 1;
  double complex* F = (double complex*)calloc((size_t)nring * (lmax + 1), sizeof(double complex));
  if (!F) return -1;
  legendre_synth(nside, lmax, (const double complex*)alm, F, use_mlim);
  phases_to_map(nside, lmax, F, map, use_mlim);
  free(F);
  return 0;
}

/* Legendre stage only: F[nring][lmax+1] complex */
int ref_alm2phase(int nside, int lmax, const double* alm, double* F, int use_mlim, int nthreads) {
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
  memset(F, 0, (size_t)(4 * nside - 1) * (lmax + 1) * 2 * sizeof(double));
  legendre_synth(nside, lmax, (const double complex*)alm, (double complex*)F, use_mlim);
  return 0;
}

/* ---- scalar analysis, one pass (the adjoint of the synthesis above) --------------------
 * G_m(r) = w_r 4pi/npix sum_j f(r,j) e^{-i m phi_j};  a_lm = sum_r lambda_lm(z_r) G_m(r).
 * healpy.map2alm(pol=False) is this pass followed by `iter` Jacobi refinements, which the
 * caller (oracle/sht_c.py) builds from this function and ref_alm2map.                     */
static void map_to_phases(int nside, int lmax, const double* map, const double* ring_w, double complex* G) {
  const int nring = 4 * nside - 1;
  const double norm = 4.0 * M_PI / (12.0 * (double)nside * (double)nside);
#pragma omp parallel
  {
    double complex* x = (double complex*)malloc((size_t)(4 * nside) * sizeof(double complex));
#pragma omp for schedule(dynamic, 4)
    for (int r = 0; r < nring; ++r) {
      ring_t g;
      ring_geom(nside, r, &g);
      const int n = g.nphi;
      const double phi0 = g.shifted ? M_PI / n : 0.0;
      for (int j = 0; j < n; ++j) x[j] = map[g.start + j];
      dft_any(x, n, -1);
      const double w = (ring_w ? ring_w[r] : 1.0) * norm;
      double complex* Gr = G + (int64_t)r * (lmax + 1);
      for (int m = 0; m <= lmax; ++m) {
        const double ang = -phi0 * (double)(m % (2 * n));
        Gr[m] = x[m % n] * (cos(ang) + I * sin(ang)) * w;
      }
    }
    free(x);
  }
}

static void legendre_analysis(int nside, int lmax, const double complex* G, double complex* alm) {
  const int nring = 4 * nside - 1, npair = 2 * nside;
  ring_t* rg = (ring_t*)malloc((size_t)npair * sizeof(ring_t));
  for (int r = 0; r < npair; ++r) ring_geom(nside, r, &rg[r]);
  real* logc = (real*)malloc((size_t)(lmax + 1) * sizeof(real));
  logc[0] = -0.5 * R_LOG((real)4 * (real)3.14159265358979323846264338327950288L);
  for (int m = 1; m <= lmax; ++m) logc[m] = logc[m - 1] + 0.5 * R_LOG((real)(2 * m + 1) / (real)(2 * m));
  const real ln2 = R_LOG((real)2);
  const real BIG = R_LDEXP((real)1, SCALE_BITS / 2), SMALL = R_LDEXP((real)1, -SCALE_BITS);
#pragma omp parallel
  {
    real* A = (real*)malloc((size_t)(lmax + 2) * sizeof(real));
    real* B = (real*)malloc((size_t)(lmax + 2) * sizeof(real));
    real* accr = (real*)malloc((size_t)(lmax + 1) * sizeof(real));
    real* acci = (real*)malloc((size_t)(lmax + 1) * sizeof(real));
#pragma omp for schedule(dynamic, 1)
    for (int m = 0; m <= lmax; ++m) {
      double complex* a_m = alm + ((int64_t)m * (2 * lmax + 1 - m)) / 2; /* index by l */
      for (int l = m + 1; l <= lmax; ++l) {
        const real l2 = (real)l * l, m2 = (real)m * m;
        A[l] = R_SQRT((4 * l2 - 1) / (l2 - m2));
      }
      for (int l = m + 2; l <= lmax; ++l) B[l] = 1 / A[l - 1];
      for (int l = m; l <= lmax; ++l) accr[l] = acci[l] = 0;
      for (int r = 0; r < npair; ++r) {
        const real z = rg[r].z;
        const double complex gn = G[(int64_t)r * (lmax + 1) + m];
        const double complex gs = (r != npair - 1) ? G[(int64_t)(nring - 1 - r) * (lmax + 1) + m] : 0;
        const real ser = (real)creal(gn) + (real)creal(gs), sei = (real)cimag(gn) + (real)cimag(gs);
        const real sor = (real)creal(gn) - (real)creal(gs), soi = (real)cimag(gn) - (real)cimag(gs);
        real lg2 = (logc[m] + m * R_LOG(rg[r].sth)) / ln2;
        int scale = 0;
        if (lg2 < -SCALE_BITS / 2) {
          scale = (int)(lg2 / SCALE_BITS);
          lg2 -= (real)scale * SCALE_BITS;
        }
        real lam1 = 0, lam2 = R_EXP(lg2 * ln2) * ((m & 1) ? -1 : 1);
        for (int l = m; l <= lmax; ++l) {
          if (l > m) {
            const real t = (l == m + 1) ? A[l] * z * lam2 : A[l] * (z * lam2 - B[l] * lam1);
            lam1 = lam2;
            lam2 = t;
            if (scale < 0 && R_FABS(lam2) > BIG) {
              lam1 *= SMALL;
              lam2 *= SMALL;
              ++scale;
            }
          }
          if (scale == 0) {
            if ((l - m) & 1) {
              accr[l] += lam2 * sor;
              acci[l] += lam2 * soi;
            } else {
              accr[l] += lam2 * ser;
              acci[l] += lam2 * sei;
            }
          }
        }
      }
      for (int l = m; l <= lmax; ++l) a_m[l] = (double)accr[l] + I * (double)((m == 0) ? 0 : acci[l]);
    }
    free(A);
    free(B);
    free(accr);
    free(acci);
  }
  free(rg);
  free(logc);
}

/* one analysis pass: alm[nalm] = A(map); ring_w [4 nside - 1] or NULL (uniform) */
int ref_map2alm_pass(int nside, int lmax, const double* map, const double* ring_w, double* alm, int nthreads) {
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
  const int nring = 4 * nside - 1;
  double complex* G = (double complex*)calloc((size_t)nring * (lmax + 1), sizeof(double complex));
  if (!G) return -1;
  map_to_phases(nside, lmax, map, ring_w, G);
  legendre_analysis(nside, lmax, G, (double complex*)alm);
  free(G);
  return 0;
}

/* ---- spin-weighted synthesis (healpy.alm2map_spin, glass/healpix.py:107) ----------------
 * sY_lm(theta, 0) = (-1)^s sqrt((2l+1)/4pi) d^l_{m,-s}(theta), Wigner d by its three-term
 * recurrence in l from l0 = max(|m|, |s|), seeded in log space and carried with the same
 * power-of-two scale as the scalar functions.  Restates oracle/healpix_ref.py::wigner_d_l /
 * slam_lm / alm2map_spin (validated there against Goldberg's closed form) for sizes NumPy is
 * too slow for.                                                                           */
static real log_fact(int n) {
#ifdef USE_LONG_DOUBLE
  return lgammal((real)n + 1);
#else
  return lgamma((double)n + 1);
#endif
}

/* acc += sum_l c_l sY_{l,mm}(theta) for l = l0..lmax, c_l complex (cr[l], ci[l]); mm may be negative */
static void spin_sum(int lmax, int mm, int s, real cth, real ch, real sh, const real* cr, const real* ci, real* outr, real* outi) {
  const int mp = -s;
  const int l0 = (abs(mm) > abs(mp)) ? abs(mm) : abs(mp);
  *outr = *outi = 0;
  if (l0 > lmax) return;
  /* seed d^{l0}_{mm,mp}: reduce to a = l0 >= |b| by d_{m',m} = (-1)^{m-m'} d_{m,m'} = d_{-m,-m'} */
  int a = mm, b = mp;
  real sign = 1;
  if (abs(a) < abs(b)) {
    const int t = a;
    a = b;
    b = t;
    if ((a - b) & 1) sign = -sign;
  }
  if (a < 0) {
    if ((a - b) & 1) sign = -sign;
    a = -a;
    b = -b;
  }
  const int j = l0;
  const real ln2 = R_LOG((real)2);
  real lg2 = (0.5 * (log_fact(2 * j) - log_fact(j + b) - log_fact(j - b)) + (j + b) * R_LOG(ch) + (j - b) * R_LOG(sh)) / ln2;
  int scale = 0;
  if (lg2 < -SCALE_BITS / 2) {
    scale = (int)(lg2 / SCALE_BITS);
    lg2 -= (real)scale * SCALE_BITS;
  }
  if ((j - b) & 1) sign = -sign;
  const real BIG = R_LDEXP((real)1, SCALE_BITS / 2), SMALL = R_LDEXP((real)1, -SCALE_BITS);
  const real fourpi = (real)4 * (real)3.14159265358979323846264338327950288L;
  const real ssign = (s & 1) ? -1 : 1;
  real d1 = 0, d2 = sign * R_EXP(lg2 * ln2); /* d^{l-1}, d^l */
  real sr = 0, si = 0;
  for (int l = l0; l <= lmax; ++l) {
    if (scale == 0) {
      const real y = ssign * R_SQRT((real)(2 * l + 1) / fourpi) * d2;
      sr += cr[l] * y;
      si += ci[l] * y;
    }
    if (l == lmax) break;
    real t;
    if (l == 0) {
      t = cth * d2; /* d^1_{00} = cos(theta) */
    } else {
      const real lp = (real)l + 1;
      const real den = (real)l * R_SQRT((lp * lp - (real)mm * mm) * (lp * lp - (real)mp * mp));
      const real t1 = (real)(2 * l + 1) * ((real)l * lp * cth - (real)mm * mp);
      const real t2 = lp * R_SQRT(((real)l * l - (real)mm * mm) * ((real)l * l - (real)mp * mp));
      t = (t1 * d2 - t2 * d1) / den;
    }
    d1 = d2;
    d2 = t;
    if (scale < 0 && R_FABS(d2) > BIG) {
      d1 *= SMALL;
      d2 *= SMALL;
      ++scale;
    }
  }
  *outr = sr;
  *outi = si;
}

/* map1 + i map2 = sum_{lm} -(alm1 + i alm2)_lm sY_lm; alm2 may be NULL (E-only, what GLASS passes) */
int ref_alm2map_spin(int nside, int lmax, int spin, const double* alm1, const double* alm2, double* map1, double* map2,
                     int nthreads) {
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
  const int nring = 4 * nside - 1;
  const double complex* E = (const double complex*)alm1;
  const double complex* Bm = (const double complex*)alm2;
  double complex* FA = (double complex*)calloc((size_t)nring * (lmax + 1), sizeof(double complex));
  double complex* FB = (double complex*)calloc((size_t)nring * (lmax + 1), sizeof(double complex));
  if (!FA || !FB) return -1;
#pragma omp parallel
  {
    real* pr = (real*)malloc((size_t)(lmax + 1) * sizeof(real));
    real* pi_ = (real*)malloc((size_t)(lmax + 1) * sizeof(real));
    real* nr = (real*)malloc((size_t)(lmax + 1) * sizeof(real));
    real* ni = (real*)malloc((size_t)(lmax + 1) * sizeof(real));
#pragma omp for schedule(dynamic, 1)
    for (int m = 0; m <= lmax; ++m) {
      const int64_t base = ((int64_t)m * (2 * lmax + 1 - m)) / 2;
      const real sg = (m & 1) ? -1 : 1;
      for (int l = m; l <= lmax; ++l) {
        const real er = (real)creal(E[base + l]), ei = (m == 0) ? 0 : (real)cimag(E[base + l]);
        const real br = Bm ? (real)creal(Bm[base + l]) : 0, bi = (Bm && m != 0) ? (real)cimag(Bm[base + l]) : 0;
        /* +m: -(e + i b);   -m: -((-1)^m conj(e) + i (-1)^m conj(b)) */
        pr[l] = -(er - bi);
        pi_[l] = -(ei + br);
        nr[l] = -sg * (er + bi);
        ni[l] = -sg * (-ei + br);
      }
      for (int r = 0; r < nring; ++r) {
        ring_t g;
        ring_geom(nside, r, &g);
        /* half-angle functions from the small quantity in the caps (1 -+ z = i^2 / 3N^2) */
        const int64_t N = nside, i = r + 1;
        real omz, opz; /* 1 - z, 1 + z */
        if (i < N) {
          omz = (real)(i * i) / (real)(3 * N * N);
          opz = 2 - omz;
        } else if (i > 3 * N) {
          opz = (real)((4 * N - i) * (4 * N - i)) / (real)(3 * N * N);
          omz = 2 - opz;
        } else {
          omz = 1 - g.z;
          opz = 1 + g.z;
        }
        const real sh = R_SQRT(omz / 2), ch = R_SQRT(opz / 2);
        real fpr, fpi, fnr = 0, fni = 0;
        spin_sum(lmax, m, spin, g.z, ch, sh, pr, pi_, &fpr, &fpi);
        if (m > 0) spin_sum(lmax, -m, spin, g.z, ch, sh, nr, ni, &fnr, &fni);
        double complex A, Bc;
        if (m == 0) {
          A = (double)fpr;
          Bc = (double)fpi;
        } else {
          /* A = (Fp + conj(Fn)) / 2,  B = -i (Fp - conj(Fn)) / 2 */
          A = (double)(0.5 * (fpr + fnr)) + I * (double)(0.5 * (fpi - fni));
          Bc = (double)(0.5 * (fpi + fni)) + I * (double)(-0.5 * (fpr - fnr));
        }
        FA[(int64_t)r * (lmax + 1) + m] = A;
        FB[(int64_t)r * (lmax + 1) + m] = Bc;
      }
    }
    free(pr);
    free(pi_);
    free(nr);
    free(ni);
  }
  phases_to_map(nside, lmax, FA, map1, 0);
  phases_to_map(nside, lmax, FB, map2, 0);
  free(FA);
  free(FB);
  return 0;
}

int ref_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
