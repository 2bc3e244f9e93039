// sht_legendre.cu -- Legendre stage of the scalar spherical-harmonic synthesis (K3).
//
// Replaces the Legendre part of healpy.alm2map -> libsharp2 (glass/healpix.py:71, called
// from glass/fields.py:429 and glass/lensing.py:326):
//     F_m(theta_r) = sum_{l=m..lmax} a_lm lambda_lm(cos theta_r)
// for every ring r and every m <= mlim(r).
//
// Algorithm (B200-first; not a translation of libsharp):
//  * x^2 recurrence.  From  x lam_l = e_{l+1} lam_{l+1} + e_l lam_{l-1},
//    e_l = sqrt((l^2-m^2)/(4l^2-1)), the even-offset functions lam_{m+2k} obey a
//    three-term recurrence in x^2.  Rescaling p_k = lam_{m+2k}/alpha_k makes it
//        p_{k+1} = (a_k x^2 + b_k) p_k - p_{k-1}        (2 DFMA per TWO l values)
//    and the odd-offset functions are x times a combination of the even ones, so
//        F_m(+-x) = sum_k p_k(x^2) (Ae_k +- x Ao_k)
//    with coefficients (Ae, Ao) obtained from a_lm by an O(lmax) pass per m
//    (sht_prep_kernel).  North and south ring of a pair share everything but the sign.
//  * one thread owns R adjacent ring pairs and walks l; a CTA owns (m, tile of ring
//    pairs); the per-m record stream {a_k, b_k, -a_k, a_k+b_k, Ae_k, Ao_k} is staged through shared
//    memory by the TMA engine (cp.async.bulk + mbarrier, multi-stage), every thread
//    reads it with broadcast LDS.128.  The inner loop is pure DFMA: (2 + 4B) per ring
//    pair per l-pair for B maps batched on one recurrence.
//  * dynamic range: lam_mm ~ sin^m(theta) underflows FP64; values carry an integer
//    scale (true = v * 2^(512*scale)).  Per warp three phases: SKIP (recurrence only,
//    nothing is significant yet), CHECK (accumulate with per-ring select + rescale
//    test), FAST (all rings at scale 0: no tests).
//  * rings with mlim(ring) < m are skipped entirely (mlim as in libsharp's
//    sharp_get_mlim heuristic, shared with the FFT stage through the plan).
#include <cstdio>
#include <cstdlib>

#include "plan.h"
#include "sht_seed.cuh"
#include "sht_tables.cuh"

namespace glb {

#ifndef GLB_LEG_UNROLL
#define GLB_LEG_UNROLL 2
#endif
constexpr int LEG_UNROLL = GLB_LEG_UNROLL;  // unroll factor of the FAST loop (development knob)
constexpr int LEG_KT = 64;      // l-pairs per smem chunk
constexpr int LEG_STAGES = 4;   // chunks in flight


// -------------------------------------------------------------------------------------
// static per-plan tables (depend on l, m only), one thread per m, run once at plan creation:
//   tab[roff[m]+k] = {a_k, b_k, a_k + b_k, alpha_k, s1_k = alpha_k/e_{l+1}, c_k = e_{l+2}/e_{l+3}},  l = m+2k
// computed in double-double and rounded once (sht_tables.cuh explains why that matters).
//
// Conditioning.  p_k is a polynomial of degree k in y = x^2 and, like any such polynomial, is
// sensitive to its argument at BOTH ends of [0, 1]: |dp_k/dy| ~ k / sqrt(y (1 - y)).  The factor
// of the recurrence is therefore evaluated per warp in the variable that is known to RELATIVE
// precision at the warp's rings:
//     y <  1/2 (towards the equator):  a_k y + b_k            with y = z^2
//     y >= 1/2 (towards the poles):    (a_k + b_k) - a_k u    with u = sin^2(theta) = t (2 - t),
//                                                             t = i^2 / (3 nside^2)
// Both are ONE FMA on a coefficient pair of the record, {a_k, b_k} or {-a_k, a_k + b_k}, chosen by
// a warp-uniform offset.  With y = fl(z z) everywhere the ABSOLUTE error 1e-16 of y near the
// poles, the same in every step, cost 2e-9 of the map at l = 8191 (d lambda_l0 / dz = l^2 / 2
// there); with u everywhere the rings next to the equator lose as much (measured:
// tests/test_gpu_fullsize.py, tests/test_gpu_sht.py::test_alm2map_vs_long_double_oracle).
// -------------------------------------------------------------------------------------
__global__ void __launch_bounds__(64) sht_prep_tables_kernel(int lmax, int mmax, const int64_t* __restrict__ roff,
                                                             double* __restrict__ tab) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m > mmax) return;
  prep_tables_for_m(lmax, m, tab + roff[m] * PREP_TAB);
}

// -------------------------------------------------------------------------------------
// prep: a_lm (m-major) -> per-m record stream {a_k, b_k, -a_k, a_k + b_k, (Ae_re, Ae_im, Ao_re, Ao_im) x B}
//   Ae_k = alpha_k a_{m+2k};  Ao_k = s1_k t_k,  t_k = a_{m+2k+1} - c_k t_{k+1}  (backward in k)
// One CTA per m.  Chunks of PREP_CH l-pairs, from the top: coalesced loads into shared memory,
// the short sequential recursion by 2B threads (one per map and re/im), coalesced record
// writes by all threads.
// -------------------------------------------------------------------------------------
constexpr int PREP_CH = 512;
constexpr int PREP_THREADS = 128;

template <int B>
__global__ void __launch_bounds__(PREP_THREADS) sht_prep_kernel(const double2* __restrict__ alm, int64_t alm_stride,
                                                                 int lmax, int mmax, const int64_t* __restrict__ roff,
                                                                 const double* __restrict__ tab, double* __restrict__ rec) {
  constexpr int REC = 4 + 4 * B;
  __shared__ double s_t[2 * B][PREP_CH + 1];  // odd-coefficient stream, then t_k in place (+1: no bank conflicts)
  __shared__ double s_c[PREP_CH];
  __shared__ double s_carry[2 * B];
  const int m = blockIdx.x;
  const int tid = threadIdx.x;
  const int K = (lmax - m) / 2 + 1;
  const double* t = tab + roff[m] * PREP_TAB;
  double* r = rec + roff[m] * REC;
  const int64_t base = (int64_t)m * (2 * lmax + 1 - m) / 2;  // index of (l=0, m) (virtual)
  if (tid < 2 * B) s_carry[tid] = 0.0;
  __syncthreads();
  for (int khi = K; khi > 0; khi -= PREP_CH) {
    const int klo = max(khi - PREP_CH, 0);
    const int n = khi - klo;
    // load o_k = a_{m+2k+1} (zero beyond lmax; m = 0: imaginary part ignored) and c_k
    for (int i = tid; i < n; i += PREP_THREADS) {
      const int k = klo + i;
      const int l = m + 2 * k;
      s_c[i] = t[(int64_t)k * PREP_TAB + TAB_C];
#pragma unroll
      for (int bb = 0; bb < B; ++bb) {
        double2 o = make_double2(0.0, 0.0);
        if (l + 1 <= lmax) {
          o = alm[bb * alm_stride + base + l + 1];
          if (m == 0) o.y = 0.0;
        }
        s_t[2 * bb][i] = o.x;
        s_t[2 * bb + 1][i] = o.y;
      }
    }
    __syncthreads();
    if (tid < 2 * B) {
      double tv = s_carry[tid];
      for (int i = n - 1; i >= 0; --i) {
        tv = s_t[tid][i] - tv * s_c[i];
        s_t[tid][i] = tv;
      }
      s_carry[tid] = tv;
    }
    __syncthreads();
    for (int i = tid; i < n; i += PREP_THREADS) {
      const int k = klo + i;
      const int l = m + 2 * k;
      const double* tk = t + (int64_t)k * PREP_TAB;
      const double alpha = tk[TAB_ALPHA], s1 = tk[TAB_S1];
      double* rk = r + (int64_t)k * REC;
      rk[0] = tk[TAB_A];
      rk[1] = tk[TAB_B];
      rk[2] = -tk[TAB_A];
      rk[3] = tk[TAB_AB];
#pragma unroll
      for (int bb = 0; bb < B; ++bb) {
        double2 v = alm[bb * alm_stride + base + l];
        if (m == 0) v.y = 0.0;  // a_l0 is real (healpy ignores the imaginary part)
        rk[4 + 4 * bb + 0] = v.x * alpha;
        rk[4 + 4 * bb + 1] = v.y * alpha;
        rk[4 + 4 * bb + 2] = s_t[2 * bb][i] * s1;
        rk[4 + 4 * bb + 3] = s_t[2 * bb + 1][i] * s1;
      }
    }
    __syncthreads();
  }
}

int plan_items(glb_plan* pl, int tile, int G, int rank, LegItem** d_items, int* nitems);

struct LegParams {
  const LegItem* items;
  const double* rec;
  const int64_t* roff;
  const double* z;
  const double* sth;
  const int* mlim;
  const double* cm_mant;
  const int* cm_exp;
  double2* phase;
  int64_t phase_map_stride;  // in double2
  int lmax, mmax, npair, nring;
  // phase layout: row(ring) * W + m / G.  Single GPU: G = 1, W = mmax+1, rowmap = null (identity).
  int G, W;
  const int* rowmap;
  // P2P instantiation only (appended, so the layout seen by the other instantiations is unchanged)
  const LegP2P* p2p;
};

// P2P = true: m-split over GPUs with the transpose fused into the store -- every F_m(ring) goes
// straight to the receive buffer of the rank that owns the ring (peer-mapped over NVLink) instead
// of a local send buffer that an all-to-all moves afterwards.
template <int R, int B, int THREADS, bool P2P = false>
__global__ void __launch_bounds__(THREADS) sht_legendre_synth_kernel(const LegParams p) {
  constexpr int REC = 4 + 4 * B;
  constexpr int CHUNK_DOUBLES = LEG_KT * REC;
  constexpr int NWARPS = THREADS / 32;
  __shared__ __align__(128) double s_rec[LEG_STAGES][CHUNK_DOUBLES];
  __shared__ __align__(8) uint64_t s_full[LEG_STAGES];
  __shared__ __align__(8) uint64_t s_empty[LEG_STAGES];

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const LegItem item = p.items[blockIdx.x];
  const int m = item.m;
  const int K = (p.lmax - m) / 2 + 1;
  const int nchunks = (K + LEG_KT - 1) / LEG_KT;
  const double* rec_m = p.rec + p.roff[m] * REC;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < LEG_STAGES; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&s_empty[s], NWARPS);
    }
    mbar_fence_init();
  }
  __syncthreads();

  auto issue = [&](int c) {
    const int s = c % LEG_STAGES;
    const int kc = min(LEG_KT, K - c * LEG_KT);
    const uint32_t bytes = (uint32_t)(kc * REC * sizeof(double));
    mbar_arrive_expect_tx(&s_full[s], bytes);
    bulk_g2s(&s_rec[s][0], rec_m + (int64_t)c * CHUNK_DOUBLES, bytes, &s_full[s]);
  };
  if (tid == 0) {
    for (int c = 0; c < LEG_STAGES - 1 && c < nchunks; ++c) issue(c);
  }

  // ---- per-thread ring state ----
  double p1[R], p2[R], x2[R], zz[R];
  int sc[R];
  bool live[R];
  double acc[R][B][4];
  const int pair0 = item.tile * (THREADS * R) + tid * R;
  const double cm_mant = p.cm_mant[m];
  const int cm_exp = p.cm_exp[m];
  // warp-uniform choice of the recurrence variable (see sht_prep_tables_kernel): u = sin^2 for
  // warps whose first ring has z^2 >= 1/2, y = z^2 otherwise; `ab_off` selects the matching
  // coefficient pair of the record
  const double zw = p.z[min(item.tile * (THREADS * R) + (tid & ~31) * R, p.npair - 1)];
  const bool use_u = zw * zw >= 0.5;
  const int ab_off = use_u ? 2 : 0;
#pragma unroll
  for (int j = 0; j < R; ++j) {
    const int r = pair0 + j;
    live[j] = (r < p.npair) && (p.mlim[min(r, p.npair - 1)] >= m);
    p1[j] = 0.0;
    p2[j] = 0.0;
    sc[j] = 0;
    x2[j] = 0.0;
    zz[j] = 0.0;
    if (live[j]) {
      zz[j] = p.z[r];
      const double sth = p.sth[r];
      x2[j] = use_u ? sth * sth : zz[j] * zz[j];  // the warp's recurrence variable: u or y
      lam_mm_scaled(m, sth, cm_mant, cm_exp, p2[j], sc[j]);
    }
#pragma unroll
    for (int b = 0; b < B; ++b) acc[j][b][0] = acc[j][b][1] = acc[j][b][2] = acc[j][b][3] = 0.0;
  }

  const double SMALL = 7.458340731200207e-155;  // 2^-512
  int phase = 0;  // 0 skip, 1 check, 2 fast (warp-uniform)

  for (int c = 0; c < nchunks; ++c) {
    const int s = c % LEG_STAGES;
    if (tid == 0) {
      const int cn = c + LEG_STAGES - 1;
      if (cn < nchunks) {
        if (cn >= LEG_STAGES) mbar_wait(&s_empty[cn % LEG_STAGES], ((cn / LEG_STAGES) - 1) & 1);
        issue(cn);
      }
    }
    mbar_wait(&s_full[s], (c / LEG_STAGES) & 1);
    const double* ck = &s_rec[s][0];
    const int kc = min(LEG_KT, K - c * LEG_KT);
    int k = 0;

    if (phase == 0) {
      // blocks of 4 l-pairs with ONE exponent test per block while every ring is still far
      // (2^-80) below the significance threshold: per step |p| grows by at most ~2^14 (first
      // steps of a large m), so four unchecked steps can neither overflow nor cross the
      // threshold.  The exponent tests are integer instructions that would otherwise
      // outnumber the 2 DFMA per ring and step.
      while (k + 4 <= kc) {
        bool near = false;
#pragma unroll
        for (int j = 0; j < R; ++j) near |= (sc[j] == 0) && (bexp(p2[j]) >= BEXP_SIG - 80);
        if (__any_sync(0xffffffffu, near)) break;  // finish with the exact per-step loop below
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const double2 ab = *reinterpret_cast<const double2*>(ck + (k + u) * REC + ab_off);
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
          if (bexp(p2[j]) >= BEXP_BIG) {
            p1[j] *= SMALL;
            p2[j] *= SMALL;
            sc[j] += 1;
          }
        }
        k += 4;
      }
      while (phase == 0 && k < kc) {
        bool sig = false;
#pragma unroll
        for (int j = 0; j < R; ++j) sig |= (sc[j] == 0) && (bexp(p2[j]) >= BEXP_SIG);
        if (__any_sync(0xffffffffu, sig)) {
          phase = 1;
          break;
        }
        const double2 ab = *reinterpret_cast<const double2*>(ck + k * REC + ab_off);
#pragma unroll
        for (int j = 0; j < R; ++j) {
          const double rr = fma(ab.x, x2[j], ab.y);
          const double t = fma(rr, p2[j], -p1[j]);
          p1[j] = p2[j];
          p2[j] = t;
          if (bexp(p2[j]) >= BEXP_BIG) {
            p1[j] *= SMALL;
            p2[j] *= SMALL;
            sc[j] += 1;
          }
        }
        ++k;
      }
    }
    if (phase == 1) {
      while (k < kc) {
        bool allz = true;
#pragma unroll
        for (int j = 0; j < R; ++j) allz &= (sc[j] == 0);
        if (__all_sync(0xffffffffu, allz)) {
          phase = 2;
          break;
        }
        const double* rk = ck + k * REC;
        const double2 ab = *reinterpret_cast<const double2*>(rk + ab_off);
        double2 ce[B], co[B];
#pragma unroll
        for (int b = 0; b < B; ++b) {
          ce[b] = *reinterpret_cast<const double2*>(rk + 4 + 4 * b);
          co[b] = *reinterpret_cast<const double2*>(rk + 6 + 4 * b);
        }
#pragma unroll
        for (int j = 0; j < R; ++j) {
          const double pa = (sc[j] == 0) ? p2[j] : 0.0;
#pragma unroll
          for (int b = 0; b < B; ++b) {
            acc[j][b][0] = fma(pa, ce[b].x, acc[j][b][0]);
            acc[j][b][1] = fma(pa, ce[b].y, acc[j][b][1]);
            acc[j][b][2] = fma(pa, co[b].x, acc[j][b][2]);
            acc[j][b][3] = fma(pa, co[b].y, acc[j][b][3]);
          }
          const double rr = fma(ab.x, x2[j], ab.y);
          const double t = fma(rr, p2[j], -p1[j]);
          p1[j] = p2[j];
          p2[j] = t;
          if (bexp(p2[j]) >= BEXP_BIG) {
            p1[j] *= SMALL;
            p2[j] *= SMALL;
            sc[j] += 1;
          }
        }
        ++k;
      }
    }
    if (phase == 2) {
#pragma unroll LEG_UNROLL
      for (; k < kc; ++k) {
        const double* rk = ck + k * REC;
        const double2 ab = *reinterpret_cast<const double2*>(rk + ab_off);
        double2 ce[B], co[B];
#pragma unroll
        for (int b = 0; b < B; ++b) {
          ce[b] = *reinterpret_cast<const double2*>(rk + 4 + 4 * b);
          co[b] = *reinterpret_cast<const double2*>(rk + 6 + 4 * b);
        }
#pragma unroll
        for (int j = 0; j < R; ++j) {
#pragma unroll
          for (int b = 0; b < B; ++b) {
            acc[j][b][0] = fma(p2[j], ce[b].x, acc[j][b][0]);
            acc[j][b][1] = fma(p2[j], ce[b].y, acc[j][b][1]);
            acc[j][b][2] = fma(p2[j], co[b].x, acc[j][b][2]);
            acc[j][b][3] = fma(p2[j], co[b].y, acc[j][b][3]);
          }
          const double rr = fma(ab.x, x2[j], ab.y);
          const double t = fma(rr, p2[j], -p1[j]);
          p1[j] = p2[j];
          p2[j] = t;
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&s_empty[s]);
  }

  // ---- write F_m for the north and south ring of every live pair ----
  if constexpr (P2P) {
    const LegP2P* __restrict__ t = p.p2p;
    const int world = __ldg(&t->world), me = __ldg(&t->rank);
    const int slot = m / p.G;
#pragma unroll
    for (int j = 0; j < R; ++j) {
      if (!live[j]) continue;
      const int r = pair0 + j;
      const int rs = p.nring - 1 - r;
      // a ring and its mirror have the same owner (msplit_layout), consecutive pairs mostly too
      const int row_n = p.rowmap[r];
      int d = 0;
      while (row_n >= __ldg(&t->rowstart[d + 1])) ++d;
      const int first = __ldg(&t->rowstart[d]);
      const int64_t rows_d = __ldg(&t->rowstart[d + 1]) - first;
      double2* base = reinterpret_cast<double2*>(__ldg(reinterpret_cast<const unsigned long long*>(&t->base[d])));
      const int row_s = r != p.npair - 1 ? p.rowmap[rs] : -1;
#pragma unroll
      for (int b = 0; b < B; ++b) {
        const double er = acc[j][b][0], ei = acc[j][b][1];
        const double orr = acc[j][b][2] * zz[j], oi = acc[j][b][3] * zz[j];
        double2* dst = base + (int64_t)(b * world + me) * rows_d * p.W + slot;
        dst[(int64_t)(row_n - first) * p.W] = make_double2(er + orr, ei + oi);
        if (row_s >= 0) dst[(int64_t)(row_s - first) * p.W] = make_double2(er - orr, ei - oi);
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < R; ++j) {
      if (!live[j]) continue;
      const int r = pair0 + j;
#pragma unroll
      for (int b = 0; b < B; ++b) {
        const double er = acc[j][b][0], ei = acc[j][b][1];
        const double orr = acc[j][b][2] * zz[j], oi = acc[j][b][3] * zz[j];
        double2* ph = p.phase + b * p.phase_map_stride;
        const int slot = m / p.G;
        const int rs = p.nring - 1 - r;
        const int row_n = p.rowmap ? p.rowmap[r] : r;
        ph[(int64_t)row_n * p.W + slot] = make_double2(er + orr, ei + oi);
        if (r != p.npair - 1) {
          const int row_s = p.rowmap ? p.rowmap[rs] : rs;
          ph[(int64_t)row_s * p.W + slot] = make_double2(er - orr, ei - oi);
        }
      }
    }
  }
}

// -------------------------------------------------------------------------------------
// host launchers
// -------------------------------------------------------------------------------------
template <int B>
static int launch_prep(glb_plan* pl, const double2* d_alm, cudaStream_t st) {
  sht_prep_kernel<B><<<pl->mmax + 1, PREP_THREADS, 0, st>>>(d_alm, pl->nalm, pl->lmax, pl->mmax, pl->d_roff,
                                                             pl->d_prep_tab, pl->d_rec);
  GLB_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return GLB_OK;
}

template <int B>
static int launch_legendre(glb_plan* pl, double2* d_phase, cudaStream_t st, bool dist = false, int p2p_buffer = -1) {
  LegParams p;
  p.p2p = p2p_buffer >= 0 ? pl->d_p2p_tab + p2p_buffer : nullptr;
  p.items = pl->d_items;
  p.rec = pl->d_rec;
  p.roff = pl->d_roff;
  p.z = pl->d_z;
  p.sth = pl->d_sth;
  p.mlim = pl->d_mlim;
  p.cm_mant = pl->d_cm_mant;
  p.cm_exp = pl->d_cm_exp;
  p.phase = d_phase;
  p.lmax = pl->lmax;
  p.mmax = pl->mmax;
  p.npair = pl->npair;
  p.nring = pl->nring;
  p.G = dist ? pl->dist_world : 1;
  p.W = dist ? pl->dist_W : pl->mmax + 1;
  p.rowmap = dist ? pl->d_dist_rowmap : nullptr;
  p.phase_map_stride = (int64_t)pl->nring * p.W;
  // (R ring pairs per thread, THREADS) per batch size.  Shared-memory wavefronts per DFMA scale
  // as 2/R (every broadcast LDS.128 costs two), registers as 8*R*B for the accumulators.
  int R = 4, threads = pl->leg_threads;
  if (B == 4 && pl->npair >= 1024) {
    R = 4;
    threads = 256;
    if (const char* env = getenv("GLB_LEG_CFG4")) {  // tuning knob "R,THREADS"
      int r = 0, t = 0;
      if (sscanf(env, "%d,%d", &r, &t) == 2) {
        R = r;
        threads = t;
      }
    }
  }
  LegItem* items = nullptr;
  int nitems = 0;
  int rc = plan_items(pl, R * threads, p.G, dist ? pl->dist_rank : 0, &items, &nitems);
  if (rc != GLB_OK) return rc;
  p.items = items;
#define GLB_LEG_LAUNCH(RR, TT)                                                   \
  if (R == RR && threads == TT) {                                                \
    sht_legendre_synth_kernel<RR, B, TT><<<nitems, TT, 0, st>>>(p);              \
    launched = true;                                                             \
  }
  bool launched = false;
  if (p.p2p) {  // fused transpose: the default configurations only
#define GLB_LEG_LAUNCH_P2P(RR, TT)                                               \
  if (R == RR && threads == TT) {                                                \
    sht_legendre_synth_kernel<RR, B, TT, true><<<nitems, TT, 0, st>>>(p);        \
    launched = true;                                                             \
  }
    GLB_LEG_LAUNCH_P2P(4, 64)
    GLB_LEG_LAUNCH_P2P(4, 128)
    GLB_LEG_LAUNCH_P2P(4, 256)
#undef GLB_LEG_LAUNCH_P2P
  } else {
  GLB_LEG_LAUNCH(4, 64)
  GLB_LEG_LAUNCH(4, 128)
  GLB_LEG_LAUNCH(4, 256)
  }
  if (B == 4 && !p.p2p) {
    GLB_LEG_LAUNCH(2, 512)
    GLB_LEG_LAUNCH(2, 256)
    GLB_LEG_LAUNCH(3, 384)
    GLB_LEG_LAUNCH(3, 256)
    GLB_LEG_LAUNCH(3, 320)
    GLB_LEG_LAUNCH(4, 192)
    GLB_LEG_LAUNCH(4, 320)
  }
#undef GLB_LEG_LAUNCH
  if (nitems == 0) launched = true;  // nothing to do for this rank
  if (!launched) {
    set_last_error("internal: legendre (R, threads) configuration not instantiated");
    return GLB_ERR_INVALID_ARG;
  }
  GLB_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return GLB_OK;
}

// static tables of the prep stage (once per plan)
int sht_build_prep_tables(glb_plan* pl, cudaStream_t st) {
  GLB_CUDA_CHECK(cudaMalloc((void**)&pl->d_prep_tab, (size_t)pl->nrec * PREP_TAB * sizeof(double)));
  const int threads = 64;  // one thread per m, sequential in l: small blocks spread the long m over the SMs
  sht_prep_tables_kernel<<<(pl->mmax + threads) / threads, threads, 0, st>>>(pl->lmax, pl->mmax, pl->d_roff,
                                                                             pl->d_prep_tab);
  GLB_CUDA_CHECK(cudaGetLastError());
  GLB_CUDA_CHECK(cudaStreamSynchronize(st));
  count_launch();
  return GLB_OK;
}

// alm [nb][nalm] -> Legendre records, nb in {1,2,4}
int sht_prep_group(glb_plan* pl, const double2* d_alm, int nb, cudaStream_t st) {
  switch (nb) {
    case 1: return launch_prep<1>(pl, d_alm, st);
    case 2: return launch_prep<2>(pl, d_alm, st);
    case 4: return launch_prep<4>(pl, d_alm, st);
    default:
      set_last_error("internal: batch group must be 1, 2 or 4");
      return GLB_ERR_INVALID_ARG;
  }
}

// records -> phase [nb][nring][mmax+1]  (dist: [nb][nring (permuted rows)][W], this rank's m only)
int sht_legendre_group(glb_plan* pl, int nb, double2* d_phase, cudaStream_t st, bool dist, int p2p_buffer) {
  switch (nb) {
    case 1: return launch_legendre<1>(pl, d_phase, st, dist, p2p_buffer);
    case 2: return launch_legendre<2>(pl, d_phase, st, dist, p2p_buffer);
    case 4: return launch_legendre<4>(pl, d_phase, st, dist, p2p_buffer);
    default:
      set_last_error("internal: batch group must be 1, 2 or 4");
      return GLB_ERR_INVALID_ARG;
  }
}

int sht_alm2phase_group(glb_plan* pl, const double2* d_alm, int nb, double2* d_phase, cudaStream_t st) {
  const int rc = sht_prep_group(pl, d_alm, nb, st);
  if (rc != GLB_OK) return rc;
  return sht_legendre_group(pl, nb, d_phase, st, false, -1);
}

}  // namespace glb
