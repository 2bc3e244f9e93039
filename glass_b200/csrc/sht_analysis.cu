// sht_analysis.cu -- Legendre stage of the scalar analysis (K10) and the map2alm driver.
//
// Replaces healpy.map2alm(map, lmax, pol=False, use_pixel_weights=True) -> libsharp2
// (glass/healpix.py:270, called from glass/lensing.py:306,408):
//     a_lm = sum_rings w_r lambda_lm(theta_r) G_m(r),   G_m(r) = sum_j f(r,j) e^{-i m phi_j} 4pi/npix
// followed by `niter` Jacobi refinements  alm += A(map - S(alm))  (healpy's default iter=3).
//
// It is the exact adjoint of the synthesis in sht_legendre.cu and shares its x^2 recurrence:
// with Ge = G_N + G_S, Go = x (G_N - G_S) per ring pair,
//     a_{m+2k}   = alpha_k sum_r p_k(r) Ge(r)
//     a_{m+2k+1} = v_k,   v_k = s1_k y_k - c_{k-1} v_{k-1},   y_k = sum_r p_k(r) Go(r)
// (the forward recursion is the transpose of the backward one in sht_prep_kernel).
// A CTA owns (m, tile of ring pairs): threads walk l for their R ring pairs, partial sums for 8
// consecutive k are held in registers, reduced across the warp by a shuffle transpose
// (31 adds for 32 values), across warps through shared memory, and written per tile;
// a final per-m pass adds the tiles in fixed order (deterministic) and applies alpha / the
// odd recursion.
#include <algorithm>

#include "plan.h"

namespace glb {

constexpr int AN_KT = 256;     // l-pairs per smem chunk (multiple of AN_KB)
constexpr int AN_KB = 8;       // l-pairs per reduction round
constexpr int AN_STAGES = 3;
constexpr int AN_BEXP_BIG = 1023 + 256;

__device__ __forceinline__ int an_bexp(double v) { return (__double2hiint(v) >> 20) & 0x7ff; }

__host__ __device__ __forceinline__ double an_eps(int l, int m) {
  if (l <= m) return 0.0;
  const double dl = (double)l, dm = (double)m;
  return sqrt(((dl - dm) * (dl + dm)) / (4.0 * dl * dl - 1.0));
}

// recurrence coefficients only: rec[roff[m] + k] = {a_k, b_k}
__global__ void __launch_bounds__(128) analysis_coef_kernel(int lmax, int mmax, const int64_t* __restrict__ roff,
                                                            double2* __restrict__ rec) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m > mmax) return;
  const int K = (lmax - m) / 2 + 1;
  double2* r = rec + roff[m];
  double alpha_km1 = 0.0, alpha_k = 1.0;
  double e_lm1 = 0.0, e_l = 0.0, e_lp1 = an_eps(m + 1, m), e_lp2 = an_eps(m + 2, m);
  for (int k = 0; k < K; ++k) {
    const int l = m + 2 * k;
    const double e_lp3 = an_eps(l + 3, m), e_lp4 = an_eps(l + 4, m);
    const double alpha_kp1 = (k == 0) ? 1.0 : alpha_km1 * ((e_l * e_lm1) / (e_lp1 * e_lp2));
    const double a = alpha_k / (e_lp1 * e_lp2 * alpha_kp1);
    r[k] = make_double2(a, -(e_lp1 * e_lp1 + e_l * e_l) * a);
    alpha_km1 = alpha_k;
    alpha_k = alpha_kp1;
    e_lm1 = e_lp1;
    e_l = e_lp2;
    e_lp1 = e_lp3;
    e_lp2 = e_lp4;
  }
}

__device__ __forceinline__ void an_lam_mm_scaled(int m, double sth, double cm_mant, int cm_exp, double& val, int& scale) {
  int e;
  double bv = frexp(sth, &e);
  int be = e;
  double rv = 1.0;
  int re = 0, mm = m;
  while (mm) {
    if (mm & 1) {
      rv *= bv;
      re += be;
      if (rv < 0.5) {
        rv *= 2.0;
        re -= 1;
      }
    }
    bv *= bv;
    be *= 2;
    if (bv < 0.5) {
      bv *= 2.0;
      be -= 1;
    }
    mm >>= 1;
  }
  double mant = rv * cm_mant;
  const int E = re + cm_exp;
  if (m & 1) mant = -mant;
  if (E >= 0) {
    scale = 0;
    val = scalbn(mant, E);
  } else {
    const int s = (-E) / 512;
    scale = -s;
    val = scalbn(mant, E + s * 512);
  }
}

struct AnaParams {
  const LegItem* items;
  const double2* rec;         // {a_k, b_k}
  const int64_t* roff;
  const double* z;
  const double* sth;
  const int* mlim;
  const double* cm_mant;
  const int* cm_exp;
  const double2* phase;       // [nring][mmax+1] weighted G_m(ring)
  double* partial;            // [ntile][nrec][4]
  int64_t nrec;
  int lmax, mmax, npair, nring;
};

template <int R, int THREADS>
__global__ void __launch_bounds__(THREADS) legendre_analysis_kernel(const AnaParams p) {
  constexpr int NWARPS = THREADS / 32;
  __shared__ __align__(128) double2 s_rec[AN_STAGES][AN_KT];
  __shared__ __align__(8) uint64_t s_full[AN_STAGES];
  __shared__ __align__(8) uint64_t s_empty[AN_STAGES];
  __shared__ double s_wsum[2][NWARPS][32];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const LegItem item = p.items[blockIdx.x];
  const int m = item.m;
  const int K = (p.lmax - m) / 2 + 1;
  const int nchunks = (K + AN_KT - 1) / AN_KT;
  const double2* rec_m = p.rec + p.roff[m];
  double* out_m = p.partial + ((int64_t)item.tile * p.nrec + p.roff[m]) * 4;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < AN_STAGES; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&s_empty[s], NWARPS);
    }
    mbar_fence_init();
  }
  __syncthreads();
  auto issue = [&](int c) {
    const int s = c % AN_STAGES;
    const int kc = min(AN_KT, K - c * AN_KT);
    const uint32_t bytes = (uint32_t)(kc * sizeof(double2));
    mbar_arrive_expect_tx(&s_full[s], bytes);
    bulk_g2s(&s_rec[s][0], rec_m + (int64_t)c * AN_KT, bytes, &s_full[s]);
  };
  if (tid == 0)
    for (int c = 0; c < AN_STAGES - 1 && c < nchunks; ++c) issue(c);

  double p1[R], p2[R], x2[R], ge_r[R], ge_i[R], go_r[R], go_i[R];
  int sc[R];
  const int pair0 = item.tile * (THREADS * R) + tid * R;
  const double cm_mant = p.cm_mant[m];
  const int cm_exp = p.cm_exp[m];
#pragma unroll
  for (int j = 0; j < R; ++j) {
    const int r = pair0 + j;
    const bool live = (r < p.npair) && (p.mlim[min(r, p.npair - 1)] >= m);
    p1[j] = p2[j] = x2[j] = ge_r[j] = ge_i[j] = go_r[j] = go_i[j] = 0.0;
    sc[j] = 0;
    if (live) {
      const double zz = p.z[r];
      x2[j] = zz * zz;
      an_lam_mm_scaled(m, p.sth[r], cm_mant, cm_exp, p2[j], sc[j]);
      const double2 gn = p.phase[(int64_t)r * (p.mmax + 1) + m];
      double2 gs = make_double2(0.0, 0.0);
      if (r != p.npair - 1) gs = p.phase[(int64_t)(p.nring - 1 - r) * (p.mmax + 1) + m];
      ge_r[j] = gn.x + gs.x;
      ge_i[j] = gn.y + gs.y;
      go_r[j] = (gn.x - gs.x) * zz;
      go_i[j] = (gn.y - gs.y) * zz;
    }
  }
  const double SMALL = 7.458340731200207e-155;  // 2^-512
  int wbuf = 0;

  for (int c = 0; c < nchunks; ++c) {
    const int s = c % AN_STAGES;
    if (tid == 0) {
      const int cn = c + AN_STAGES - 1;
      if (cn < nchunks) {
        if (cn >= AN_STAGES) mbar_wait(&s_empty[cn % AN_STAGES], ((cn / AN_STAGES) - 1) & 1);
        issue(cn);
      }
    }
    mbar_wait(&s_full[s], (c / AN_STAGES) & 1);
    const double2* ck = &s_rec[s][0];
    const int kc = min(AN_KT, K - c * AN_KT);

    for (int k0 = 0; k0 < kc; k0 += AN_KB) {
      double part[AN_KB * 4];
#pragma unroll
      for (int i = 0; i < AN_KB * 4; ++i) part[i] = 0.0;
      bool contrib = false;
#pragma unroll
      for (int j = 0; j < R; ++j) contrib |= (sc[j] == 0) && (p2[j] != 0.0);
#pragma unroll
      for (int kk = 0; kk < AN_KB; ++kk) {
        if (k0 + kk < kc) {
          const double2 ab = ck[k0 + kk];
#pragma unroll
          for (int j = 0; j < R; ++j) {
            const double pa = (sc[j] == 0) ? p2[j] : 0.0;
            part[kk * 4 + 0] = fma(pa, ge_r[j], part[kk * 4 + 0]);
            part[kk * 4 + 1] = fma(pa, ge_i[j], part[kk * 4 + 1]);
            part[kk * 4 + 2] = fma(pa, go_r[j], part[kk * 4 + 2]);
            part[kk * 4 + 3] = fma(pa, go_i[j], part[kk * 4 + 3]);
            const double rr = fma(ab.x, x2[j], ab.y);
            const double t = fma(rr, p2[j], -p1[j]);
            p1[j] = p2[j];
            p2[j] = t;
            if (an_bexp(p2[j]) >= AN_BEXP_BIG) {
              p1[j] *= SMALL;
              p2[j] *= SMALL;
              sc[j] += 1;
            }
          }
        }
      }
#pragma unroll
      for (int j = 0; j < R; ++j) contrib |= (sc[j] == 0) && (p2[j] != 0.0);
      // warp transpose-reduce: 32 values per lane -> lane i holds the warp total of value i
#define GLB_TR_STEP(O, NH)                                                   \
  {                                                                          \
    const bool up = (lane & (O)) != 0;                                       \
    _Pragma("unroll") for (int i = 0; i < (NH); ++i) {                       \
      const double send = up ? part[i] : part[i + (NH)];                     \
      const double keep = up ? part[i + (NH)] : part[i];                     \
      part[i] = keep + __shfl_xor_sync(0xffffffffu, send, (O));              \
    }                                                                        \
  }
      GLB_TR_STEP(16, 16)
      GLB_TR_STEP(8, 8)
      GLB_TR_STEP(4, 4)
      GLB_TR_STEP(2, 2)
      GLB_TR_STEP(1, 1)
#undef GLB_TR_STEP
      s_wsum[wbuf][warp][lane] = part[0];
      const int any = __syncthreads_or(contrib ? 1 : 0);
      if (any && warp == 0) {
        double tot = 0.0;
#pragma unroll
        for (int w = 0; w < NWARPS; ++w) tot += s_wsum[wbuf][w][lane];
        const int kidx = c * AN_KT + k0 + (lane >> 2);
        if (kidx < K) out_m[(int64_t)kidx * 4 + (lane & 3)] = tot;
      }
      wbuf ^= 1;
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&s_empty[s]);
  }
}

// per m: add the tiles in order, a_{m+2k} = alpha_k ce_k, odd coefficients by the forward
// recursion; accumulate != 0 adds to alm (Jacobi refinement)
__global__ void __launch_bounds__(128) analysis_finalize_kernel(int lmax, int mmax, const int64_t* __restrict__ roff,
                                                                const double* __restrict__ partial, int ntile,
                                                                int64_t nrec, int accumulate, double2* __restrict__ alm) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m > mmax) return;
  const int K = (lmax - m) / 2 + 1;
  const int64_t base = (int64_t)m * (2 * lmax + 1 - m) / 2;
  double alpha_km1 = 0.0, alpha_k = 1.0;
  double e_lm1 = 0.0, e_l = 0.0, e_lp1 = an_eps(m + 1, m), e_lp2 = an_eps(m + 2, m);
  double vr = 0.0, vi = 0.0, c_prev = 0.0;
  for (int k = 0; k < K; ++k) {
    const int l = m + 2 * k;
    const double e_lp3 = an_eps(l + 3, m), e_lp4 = an_eps(l + 4, m);
    const double alpha_kp1 = (k == 0) ? 1.0 : alpha_km1 * ((e_l * e_lm1) / (e_lp1 * e_lp2));
    double s0 = 0.0, s1v = 0.0, s2 = 0.0, s3 = 0.0;
    for (int t = 0; t < ntile; ++t) {
      const double* q = partial + ((int64_t)t * nrec + roff[m] + k) * 4;
      s0 += q[0];
      s1v += q[1];
      s2 += q[2];
      s3 += q[3];
    }
    double2 ev = make_double2(alpha_k * s0, alpha_k * s1v);
    const double s1 = alpha_k / e_lp1;
    vr = s1 * s2 - c_prev * vr;
    vi = s1 * s3 - c_prev * vi;
    c_prev = e_lp2 / e_lp3;
    if (m == 0) ev.y = 0.0;
    if (accumulate) {
      const double2 o = alm[base + l];
      ev.x += o.x;
      ev.y += o.y;
    }
    alm[base + l] = ev;
    if (l + 1 <= lmax) {
      double2 ov = make_double2(vr, (m == 0) ? 0.0 : vi);
      if (accumulate) {
        const double2 o = alm[base + l + 1];
        ov.x += o.x;
        ov.y += o.y;
      }
      alm[base + l + 1] = ov;
    }
    alpha_km1 = alpha_k;
    alpha_k = alpha_kp1;
    e_lm1 = e_lp1;
    e_l = e_lp2;
    e_lp1 = e_lp3;
    e_lp2 = e_lp4;
  }
}

__global__ void __launch_bounds__(256) residual_kernel(const double* __restrict__ a, const double* __restrict__ b,
                                                       int64_t n, double* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = a[i] - b[i];
}

int sht_map2phase_group(glb_plan* pl, const double* const* d_maps, int nb, const double* d_ring_w, double2* d_phase,
                        cudaStream_t st);

int plan_ensure_analysis(glb_plan* pl) {
  if (pl->d_partial) return GLB_OK;
  const int T = pl->leg_threads * pl->leg_R;
  pl->ana_ntile = (pl->npair + T - 1) / T;
  const size_t bytes = (size_t)pl->ana_ntile * pl->nrec * 4 * sizeof(double);
  GLB_CUDA_CHECK(cudaMalloc((void**)&pl->d_partial, bytes));
  GLB_CUDA_CHECK(cudaMalloc((void**)&pl->d_tmpmap, (size_t)pl->npix * sizeof(double) * 2));
  pl->workspace_bytes += (int64_t)bytes + pl->npix * 16;
  return GLB_OK;
}

// one analysis pass: map -> alm (overwrite or accumulate)
int sht_analysis_pass(glb_plan* pl, const double* d_map, const double* d_ring_w, int accumulate, double2* d_alm,
                      cudaStream_t st) {
  int rc;
  const double* maps[1] = {d_map};
  if ((rc = sht_map2phase_group(pl, maps, 1, d_ring_w, pl->d_phase, st)) != GLB_OK) return rc;
  GLB_CUDA_CHECK(cudaMemsetAsync(pl->d_partial, 0, (size_t)pl->ana_ntile * pl->nrec * 4 * sizeof(double), st));
  const int threads = 128, blocks = (pl->mmax + threads) / threads;
  analysis_coef_kernel<<<blocks, threads, 0, st>>>(pl->lmax, pl->mmax, pl->d_roff, reinterpret_cast<double2*>(pl->d_rec));
  AnaParams p;
  p.items = pl->d_items;
  p.rec = reinterpret_cast<const double2*>(pl->d_rec);
  p.roff = pl->d_roff;
  p.z = pl->d_z;
  p.sth = pl->d_sth;
  p.mlim = pl->d_mlim;
  p.cm_mant = pl->d_cm_mant;
  p.cm_exp = pl->d_cm_exp;
  p.phase = pl->d_phase;
  p.partial = pl->d_partial;
  p.nrec = pl->nrec;
  p.lmax = pl->lmax;
  p.mmax = pl->mmax;
  p.npair = pl->npair;
  p.nring = pl->nring;
  constexpr int R = 4;
  if (pl->leg_threads == 64)
    legendre_analysis_kernel<R, 64><<<pl->nitems, 64, 0, st>>>(p);
  else if (pl->leg_threads == 128)
    legendre_analysis_kernel<R, 128><<<pl->nitems, 128, 0, st>>>(p);
  else
    legendre_analysis_kernel<R, 256><<<pl->nitems, 256, 0, st>>>(p);
  analysis_finalize_kernel<<<blocks, threads, 0, st>>>(pl->lmax, pl->mmax, pl->d_roff, pl->d_partial, pl->ana_ntile,
                                                       pl->nrec, accumulate, d_alm);
  GLB_CUDA_CHECK(cudaGetLastError());
  count_launch(3);
  return GLB_OK;
}

int sht_residual(const double* a, const double* b, int64_t n, double* out, cudaStream_t st) {
  residual_kernel<<<(unsigned)std::min<int64_t>((n + 255) / 256, 148 * 16), 256, 0, st>>>(a, b, n, out);
  GLB_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return GLB_OK;
}

}  // namespace glb
