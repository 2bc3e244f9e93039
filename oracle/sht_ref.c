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

int ref_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
