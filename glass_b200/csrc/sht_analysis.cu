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
// (31 adds for 32 values) and written per WARP tile -- there is no CTA-wide reduction and no
// barrier in the l loop, so warps in cheap phases (rings near the poles) run ahead instead of
// waiting for the slowest warp every 64 l-pairs; a final per-m pass adds the warp tiles in
// fixed order (deterministic) and applies alpha / the odd recursion.
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "plan.h"
#include "sht_tables.cuh"

namespace glb {

constexpr int AN_KT = 256;     // l-pairs per smem chunk (multiple of AN_SK)
constexpr int AN_KB = 8;       // l-pairs per warp-level reduction round (AN_KB * 4 = 32 values = one per lane)
constexpr int AN_TRS = 34;     // row stride (doubles) of the per-warp transpose buffer
constexpr int AN_STAGES = 3;
constexpr int AN_BEXP_BIG = 1023 + 256;
constexpr int AN_BEXP_SIG = 1023 - 70;   // as BEXP_SIG of the synthesis kernel

__device__ __forceinline__ int an_bexp(double v) { return (__double2hiint(v) >> 20) & 0x7ff; }

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
  const double2* rec;         // two pairs per l-pair: {a_k, b_k}, {-a_k, a_k + b_k}
  const int64_t* roff;
  const double* z;
  const double* sth;
  const int* mlim;
  const double* cm_mant;
  const int* cm_exp;
  const double2* phase;       // [nring][mmax+1] weighted G_m(ring)
  double* partial;            // [ntile * warps per CTA][nrec][4]: one slab per warp tile
  int64_t nrec;
  int lmax, mmax, npair, nring;
};

template <int R, int THREADS>
__global__ void __launch_bounds__(THREADS, (R > 4 ? 256 : 512) / THREADS) legendre_analysis_kernel(const AnaParams p) {
  constexpr int NWARPS = THREADS / 32;
  __shared__ __align__(128) double2 s_rec[AN_STAGES][2 * AN_KT];
  __shared__ __align__(8) uint64_t s_full[AN_STAGES];
  __shared__ __align__(8) uint64_t s_empty[AN_STAGES];
  extern __shared__ __align__(16) double s_tr_dyn[];  // [NWARPS][32 * AN_TRS] per-warp transpose buffers of the reduction

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const LegItem item = p.items[blockIdx.x];
  const int m = item.m;
  const int K = (p.lmax - m) / 2 + 1;
  const int nchunks = (K + AN_KT - 1) / AN_KT;
  const double2* rec_m = p.rec + 2 * p.roff[m];
  double* out_m = p.partial + (((int64_t)item.tile * NWARPS + (threadIdx.x >> 5)) * p.nrec + p.roff[m]) * 4;

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
    const uint32_t bytes = (uint32_t)(kc * 2 * sizeof(double2));
    mbar_arrive_expect_tx(&s_full[s], bytes);
    bulk_g2s(&s_rec[s][0], rec_m + (int64_t)c * (2 * AN_KT), bytes, &s_full[s]);
  };
  if (tid == 0)
    for (int c = 0; c < AN_STAGES - 1 && c < nchunks; ++c) issue(c);

  double p1[R], p2[R], x2[R], ge_r[R], ge_i[R], go_r[R], go_i[R];
  int sc[R];
  const int pair0 = item.tile * (THREADS * R) + tid * R;
  const double cm_mant = p.cm_mant[m];
  const int cm_exp = p.cm_exp[m];
  // warp-uniform recurrence variable, as in the synthesis kernel: u = sin^2 where z^2 >= 1/2
  const double zw = p.z[min(item.tile * (THREADS * R) + (tid & ~31) * R, p.npair - 1)];
  const bool use_u = zw * zw >= 0.5;
#pragma unroll
  for (int j = 0; j < R; ++j) {
    const int r = pair0 + j;
    const bool live = (r < p.npair) && (p.mlim[min(r, p.npair - 1)] >= m);
    p1[j] = p2[j] = x2[j] = ge_r[j] = ge_i[j] = go_r[j] = go_i[j] = 0.0;
    sc[j] = 0;
    if (live) {
      const double zz = p.z[r];
      const double sth = p.sth[r];
      x2[j] = use_u ? sth * sth : zz * zz;  // the warp's recurrence variable, see sht_prep_tables_kernel
      an_lam_mm_scaled(m, sth, cm_mant, cm_exp, p2[j], sc[j]);
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
  bool skipping = true;  // warp-uniform: no ring has come near significance yet

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
    const double2* ck = &s_rec[s][use_u ? 1 : 0];  // pair of l-pair k at ck[2 k]
    const int kc = min(AN_KT, K - c * AN_KT);

    for (int k0 = 0; k0 < kc; k0 += AN_KB) {
      double* dst = out_m + ((int64_t)c * AN_KT + k0) * 4;
      if (skipping && k0 + AN_KB <= kc) {
        // SKIP round: while no ring of the warp is both at scale 0 and within 2^-120 of the
        // significance threshold 2^-70 (the synthesis kernel's rule, sht_legendre.cu), nothing
        // can contribute during the next AN_KB steps (per step |p| grows by at most ~2^14):
        // run the recurrence only -- 2 DFMA per ring and step instead of 6 plus selects and
        // exponent tests -- and write the round's zeros.  One rescale test per round is enough
        // (|p| < 2^256 before, < 2^368 after eight steps).
        bool near = false;
#pragma unroll
        for (int j = 0; j < R; ++j) near |= (sc[j] == 0) && (an_bexp(p2[j]) >= AN_BEXP_SIG - 120);
        if (!__any_sync(0xffffffffu, near)) {
#pragma unroll
          for (int kk = 0; kk < AN_KB; ++kk) {
            const double2 ab = ck[2 * (k0 + kk)];
#pragma unroll
            for (int j = 0; j < R; ++j) {
              const double rr = fma(ab.x, x2[j], ab.y);
              const double t = fma(rr, p2[j], -p1[j]);
              p1[j] = p2[j];
              p2[j] = t;
            }
          }
#pragma unroll
          for (int j = 0; j < R; ++j) {
            if (an_bexp(p2[j]) >= AN_BEXP_BIG) {
              p1[j] *= SMALL;
              p2[j] *= SMALL;
              sc[j] += 1;
            }
          }
          dst[lane] = 0.0;
          continue;
        }
        skipping = false;
      }
      double part[AN_KB * 4];
#pragma unroll
      for (int i = 0; i < AN_KB * 4; ++i) part[i] = 0.0;
      // FAST when every ring of the warp is at scale 0 (no select, no rescale test) and the
      // round is complete; otherwise the CHECKED form
      bool allz = (k0 + AN_KB <= kc);
#pragma unroll
      for (int j = 0; j < R; ++j) allz &= (sc[j] == 0);
      if (__all_sync(0xffffffffu, allz)) {
#pragma unroll
        for (int kk = 0; kk < AN_KB; ++kk) {
          const double2 ab = ck[2 * (k0 + kk)];
#pragma unroll
          for (int j = 0; j < R; ++j) {
            part[kk * 4 + 0] = fma(p2[j], ge_r[j], part[kk * 4 + 0]);
            part[kk * 4 + 1] = fma(p2[j], ge_i[j], part[kk * 4 + 1]);
            part[kk * 4 + 2] = fma(p2[j], go_r[j], part[kk * 4 + 2]);
            part[kk * 4 + 3] = fma(p2[j], go_i[j], part[kk * 4 + 3]);
            const double rr = fma(ab.x, x2[j], ab.y);
            const double t = fma(rr, p2[j], -p1[j]);
            p1[j] = p2[j];
            p2[j] = t;
          }
        }
      } else {
#pragma unroll
        for (int kk = 0; kk < AN_KB; ++kk) {
          if (k0 + kk < kc) {
            const double2 ab = ck[2 * (k0 + kk)];
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
      }
      // warp reduction of the AN_KB*4 = 32 values through shared memory: lane L stores value i
      // at [i][L] (consecutive lanes, conflict-free), then reads row L with 16-byte loads (row
      // stride 34 doubles: 8 consecutive lanes start in 8 different 16-byte bank groups) and adds
      // the 32 entries in a fixed order.  79 instructions per round against 217 for the
      // select-and-shuffle transpose (4 FSEL + 2 SHFL + 1 DADD per output).
      // value v = kk*4 + q of this round belongs to l-pair k0 + kk: consecutive lanes write
      // consecutive doubles of this warp tile's slab (always written: no memset, no flags)
      double* tr = s_tr_dyn + warp * (32 * AN_TRS);
#pragma unroll
      for (int i = 0; i < AN_KB * 4; ++i) tr[i * AN_TRS + lane] = part[i];
      __syncwarp();
      // eight independent chains of four adds, then a three-level tree (11 dependent DADDs
      // became 7): the adds sit between two rounds of FMAs with nothing else to overlap
      double t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0, t4 = 0.0, t5 = 0.0, t6 = 0.0, t7 = 0.0;
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        const double2 a = *reinterpret_cast<const double2*>(tr + lane * AN_TRS + i);
        const double2 b = *reinterpret_cast<const double2*>(tr + lane * AN_TRS + i + 2);
        const double2 cc = *reinterpret_cast<const double2*>(tr + lane * AN_TRS + i + 4);
        const double2 dd = *reinterpret_cast<const double2*>(tr + lane * AN_TRS + i + 6);
        t0 += a.x;
        t1 += a.y;
        t2 += b.x;
        t3 += b.y;
        t4 += cc.x;
        t5 += cc.y;
        t6 += dd.x;
        t7 += dd.y;
      }
      __syncwarp();
      if (k0 + (lane >> 2) < kc) dst[lane] = ((t0 + t1) + (t2 + t3)) + ((t4 + t5) + (t6 + t7));
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&s_empty[s]);
  }
}

// rec[2i], rec[2i+1] = {a_k, b_k}, {-a_k, a_k + b_k} gathered from the static prep table (once per plan)
__global__ void __launch_bounds__(256) analysis_ab_gather_kernel(const double* __restrict__ tab, int64_t nrec,
                                                                 double2* __restrict__ ab) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nrec) {
    const double* t = tab + i * PREP_TAB;
    ab[2 * i] = make_double2(t[TAB_A], t[TAB_B]);
    ab[2 * i + 1] = make_double2(-t[TAB_A], t[TAB_AB]);
  }
}

// One CTA per m: add the tiles in fixed order (coalesced, deterministic), a_{m+2k} = alpha_k ce_k,
// odd coefficients by the forward recursion v_k = s1_k y_k - c_{k-1} v_{k-1} run by two threads
// (re, im) on shared memory; accumulate != 0 adds to alm (Jacobi refinement).
constexpr int FIN_CH = 512;
constexpr int FIN_THREADS = 128;

__global__ void __launch_bounds__(FIN_THREADS) analysis_finalize_kernel(int lmax, int mmax, const int64_t* __restrict__ roff,
                                                                        const double* __restrict__ tab,
                                                                        const double* __restrict__ partial, int ntile,
                                                                        const int* __restrict__ first_tile,
                                                                        int64_t nrec, int accumulate,
                                                                        double2* __restrict__ alm) {
  __shared__ double s_y[2][FIN_CH + 1];   // s1_k * y_k (re, im), then v_k in place
  __shared__ double s_c[FIN_CH];          // c_{k-1}
  __shared__ double s_carry[2];
  const int m = blockIdx.x, tid = threadIdx.x;
  const int K = (lmax - m) / 2 + 1;
  const int64_t base = (int64_t)m * (2 * lmax + 1 - m) / 2;
  const double* t = tab + roff[m] * PREP_TAB;
  const int tl0 = first_tile[m];
  if (tid < 2) s_carry[tid] = 0.0;
  __syncthreads();
  for (int klo = 0; klo < K; klo += FIN_CH) {
    const int n = min(FIN_CH, K - klo);
    for (int i = tid; i < n; i += FIN_THREADS) {
      const int k = klo + i;
      const int l = m + 2 * k;
      double s0 = 0.0, s1v = 0.0, s2 = 0.0, s3 = 0.0;
      for (int tl = tl0; tl < ntile; ++tl) {  // warp tiles of CTA tiles without a live ring were never written
        const double4 q = *reinterpret_cast<const double4*>(partial + ((int64_t)tl * nrec + roff[m] + k) * 4);
        s0 += q.x;
        s1v += q.y;
        s2 += q.z;
        s3 += q.w;
      }
      const double* tk = t + (int64_t)k * PREP_TAB;
      const double alpha = tk[TAB_ALPHA], s1 = tk[TAB_S1];
      double2 ev = make_double2(alpha * s0, (m == 0) ? 0.0 : alpha * s1v);
      if (accumulate) {
        const double2 o = alm[base + l];
        ev.x += o.x;
        ev.y += o.y;
      }
      alm[base + l] = ev;
      s_y[0][i] = s1 * s2;
      s_y[1][i] = s1 * s3;
      s_c[i] = (k > 0) ? t[(int64_t)(k - 1) * PREP_TAB + TAB_C] : 0.0;
    }
    __syncthreads();
    if (tid < 2) {
      double v = s_carry[tid];
      for (int i = 0; i < n; ++i) {
        v = s_y[tid][i] - s_c[i] * v;
        s_y[tid][i] = v;
      }
      s_carry[tid] = v;
    }
    __syncthreads();
    for (int i = tid; i < n; i += FIN_THREADS) {
      const int l = m + 2 * (klo + i);
      if (l + 1 <= lmax) {
        double2 ov = make_double2(s_y[0][i], (m == 0) ? 0.0 : s_y[1][i]);
        if (accumulate) {
          const double2 o = alm[base + l + 1];
          ov.x += o.x;
          ov.y += o.y;
        }
        alm[base + l + 1] = ov;
      }
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) residual_kernel(const double* __restrict__ a, const double* __restrict__ b,
                                                       int64_t n, double* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = a[i] - b[i];
}

int sht_map2phase_group(glb_plan* pl, const double* const* d_maps, int nb, const double* d_ring_w, double2* d_phase,
                        cudaStream_t st);

int plan_items(glb_plan* pl, int tile, int G, int rank, LegItem** d_items, int* nitems);

// ring pairs per thread of the analysis kernel: the cross-lane reduction of a round costs the
// same for any R, so more rings per thread amortise it better (8: one CTA of 256 threads per SM)
static int analysis_R(const glb_plan* pl) {
  int R = (pl->npair >= 2048) ? 8 : 4;
  if (const char* env = getenv("GLB_AN_R")) {
    const int r = atoi(env);
    if (r == 4 || r == 8) R = r;
  }
  return R;
}

int plan_ensure_analysis(glb_plan* pl) {
  if (pl->d_partial) return GLB_OK;
  const int T = pl->leg_threads * analysis_R(pl);
  const int nwarps = pl->leg_threads / 32;
  pl->ana_ntile = (pl->npair + T - 1) / T * nwarps;  // warp tiles
  {
    // first warp tile written for each m: CTA tiles below rmin[m] / T have no work item
    std::vector<int> ft(pl->mmax + 1);
    for (int m = 0; m <= pl->mmax; ++m) ft[m] = std::min(pl->h_rmin[m] / T, (pl->npair + T - 1) / T) * nwarps;
    GLB_CUDA_CHECK(cudaMalloc((void**)&pl->d_ana_first_tile, ft.size() * sizeof(int)));
    GLB_CUDA_CHECK(cudaMemcpy(pl->d_ana_first_tile, ft.data(), ft.size() * sizeof(int), cudaMemcpyHostToDevice));
  }
  const size_t bytes = (size_t)pl->ana_ntile * pl->nrec * 4 * sizeof(double);
  GLB_CUDA_CHECK(cudaMalloc((void**)&pl->d_partial, bytes));
  GLB_CUDA_CHECK(cudaMalloc((void**)&pl->d_tmpmap, (size_t)pl->npix * sizeof(double) * (pl->max_batch + 1)));
  GLB_CUDA_CHECK(cudaMalloc((void**)&pl->d_ab_tab, (size_t)pl->nrec * 2 * sizeof(double2)));
  analysis_ab_gather_kernel<<<(unsigned)((pl->nrec + 255) / 256), 256>>>(pl->d_prep_tab, pl->nrec,
                                                                        reinterpret_cast<double2*>(pl->d_ab_tab));
  GLB_CUDA_CHECK(cudaGetLastError());
  GLB_CUDA_CHECK(cudaDeviceSynchronize());
  count_launch();
  pl->workspace_bytes += (int64_t)bytes + pl->npix * 8 * (pl->max_batch + 1);
  return GLB_OK;
}

// one analysis pass: map -> alm (overwrite or accumulate)
int sht_analysis_pass(glb_plan* pl, const double* d_map, const double* d_ring_w, int accumulate, double2* d_alm,
                      cudaStream_t st) {
  int rc;
  const double* maps[1] = {d_map};
  if ((rc = sht_map2phase_group(pl, maps, 1, d_ring_w, pl->d_phase, st)) != GLB_OK) return rc;
  AnaParams p;
  const int R = analysis_R(pl);
  LegItem* items = nullptr;
  int nitems = 0;
  if ((rc = plan_items(pl, R * pl->leg_threads, 1, 0, &items, &nitems)) != GLB_OK) return rc;
  p.items = items;
  p.rec = reinterpret_cast<const double2*>(pl->d_ab_tab);
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
#define GLB_AN_LAUNCH(RR, TT)                                                                                  \
  {                                                                                                            \
    const size_t smem = (size_t)(TT / 32) * 32 * AN_TRS * sizeof(double);                                      \
    static bool attr_set = false;                                                                              \
    if (!attr_set) {                                                                                           \
      GLB_CUDA_CHECK(cudaFuncSetAttribute(legendre_analysis_kernel<RR, TT>,                                    \
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));            \
      attr_set = true;                                                                                         \
    }                                                                                                          \
    legendre_analysis_kernel<RR, TT><<<nitems, TT, smem, st>>>(p);                                             \
  }
  if (R == 8) {
    if (pl->leg_threads == 64)
      GLB_AN_LAUNCH(8, 64)
    else if (pl->leg_threads == 128)
      GLB_AN_LAUNCH(8, 128)
    else
      GLB_AN_LAUNCH(8, 256)
  } else {
    if (pl->leg_threads == 64)
      GLB_AN_LAUNCH(4, 64)
    else if (pl->leg_threads == 128)
      GLB_AN_LAUNCH(4, 128)
    else
      GLB_AN_LAUNCH(4, 256)
  }
#undef GLB_AN_LAUNCH
  analysis_finalize_kernel<<<pl->mmax + 1, FIN_THREADS, 0, st>>>(pl->lmax, pl->mmax, pl->d_roff, pl->d_prep_tab,
                                                                 pl->d_partial, pl->ana_ntile, pl->d_ana_first_tile,
                                                                 pl->nrec, accumulate, d_alm);
  GLB_CUDA_CHECK(cudaGetLastError());
  count_launch(2);
  return GLB_OK;
}

int sht_residual(const double* a, const double* b, int64_t n, double* out, cudaStream_t st) {
  residual_kernel<<<(unsigned)std::min<int64_t>((n + 255) / 256, 148 * 16), 256, 0, st>>>(a, b, n, out);
  GLB_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return GLB_OK;
}

}  // namespace glb
