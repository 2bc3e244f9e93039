// sht_spin.cu -- Legendre stage of the spin-weighted synthesis (K11).
//
// Replaces healpy.alm2map_spin -> libsharp2 (glass/healpix.py:107, called from
// glass/lensing.py:343 (spin 1), :366 and :428 (spin 2)):
//     map1 + i map2 = sum_{lm} -(alm1 + i alm2)_lm  sY_lm(theta, phi)
//
// With lam+_l = sY_{l,m}(theta,0), lam-_l = sY_{l,-m}(theta,0) (real, Wigner-d based) the
// Fourier coefficients of the two REAL maps on a ring are, for E-only input,
//     A_m = -sum_l E_l W_l,   B_m = i sum_l E_l X_l,   W,X = (lam+ +- (-1)^m lam-)/2
// (a B-mode input is the same transform with the outputs rotated: map1(0,B) = -map2(B,0),
// map2(0,B) = map1(B,0)).  Both functions follow the three-term recurrence in l of
// d^l_{m,-s}; rescaled by sigma_l (sigma_{l+1} = C_l sigma_{l-1}) it becomes
//     p+-_{l+1} = (x A'_l +- B'_l) p+-_l - p+-_{l-1}            (2 DFMA per function per l)
// The factor is evaluated per warp in the variable that is known to relative precision at its
// rings: x A' +- B' with x = cos(theta) where x < 1/2, (A' +- B') - A' t with t = 1 - x =
// 2 sin^2(theta/2) where x >= 1/2 (the absolute rounding error of x next to the poles, the same
// in every step, is amplified by l^2/2 -- 4e-10 of the functions at l = 8191, 2e-11 with t;
// same reasoning as in sht_legendre.cu).  Both are one FMA on a coefficient triple of the
// record, {A', B', -B'} or {-A', A'+B', A'-B'}, chosen by a warp-uniform offset.
// and the southern ring of a pair needs no second recurrence:
//     lam+(pi-theta) = (-1)^{l+s} lam-(theta)
// so sums are kept separately for even and odd (l+m+s).  Everything else -- one thread per
// R ring pairs, the (m, ring tile) work list, TMA-staged record stream, scaled arithmetic
// with SKIP/CHECK/FAST phases -- mirrors sht_legendre.cu.
#include <algorithm>
#include <cmath>

#include "plan.h"

namespace glb {

constexpr int SP_KT = 64;      // l values per smem chunk (even)
constexpr int SP_STAGES = 4;
constexpr int SP_SCALE_BITS = 512;
constexpr int SP_BEXP_BIG = 1023 + 256;
constexpr int SP_BEXP_SIG = 1023 - 70;

__device__ __forceinline__ int sp_bexp(double v) { return (__double2hiint(v) >> 20) & 0x7ff; }

// ---- static tables (once per plan and spin): tab[soff[m] + (l-l0)] = {A'_l, B'_l, sigma_l} ----
__global__ void __launch_bounds__(128) spin_tables_kernel(int lmax, int mmax, int spin, const int64_t* __restrict__ soff,
                                                          double* __restrict__ tab) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m > mmax) return;
  const int l0 = max(m, spin);
  if (l0 > lmax) return;
  double* r = tab + soff[m] * 3;
  const double dm = (double)m, ds = (double)spin;
  auto Rfun = [&](int l) {
    const double dl = (double)l;
    const double v = ((dl - dm) * (dl + dm)) * ((dl - ds) * (dl + ds));
    return sqrt(fmax(v, 0.0));
  };
  double sig_lm1 = 1.0, sig_l = 1.0;  // sigma_{l-1}, sigma_l
  double R_l = Rfun(l0);
  for (int l = l0; l <= lmax; ++l) {
    const double dl = (double)l;
    const double R_lp1 = Rfun(l + 1);
    const double nr = sqrt((2.0 * dl + 3.0) / (2.0 * dl + 1.0));
    const double Al = (2.0 * dl + 1.0) * (dl + 1.0) * nr / R_lp1;
    const double Bl = (2.0 * dl + 1.0) * dm * ds * nr / (dl * R_lp1);
    double sig_lp1 = 1.0;
    if (l > l0) {
      const double Cl = (dl + 1.0) * R_l * sqrt((2.0 * dl + 3.0) / (2.0 * dl - 1.0)) / (dl * R_lp1);
      sig_lp1 = Cl * sig_lm1;
    }
    const double ratio = sig_l / sig_lp1;
    double* rk = r + (int64_t)(l - l0) * 3;
    rk[0] = Al * ratio;
    rk[1] = Bl * ratio;
    rk[2] = sig_l;
    sig_lm1 = sig_l;
    sig_l = sig_lp1;
    R_l = R_lp1;
  }
}

// ---- prep (per call, fully parallel): records {A', B', -B', 0, -A', A'+B', A'-B', 0, (a_l sigma_l) x NB} ----
// NB coefficient sets share the recurrence: E and B of one map, or the E modes of NB maps.
// grid: (ceil((lmax+1)/256), mmax+1): blockIdx.y = m, threads over l
struct SpinAlms {
  const double2* a[4];
};
template <int NB>
__global__ void __launch_bounds__(256) spin_prep_kernel(const SpinAlms alms, int lmax, int spin,
                                                         const int64_t* __restrict__ soff, const double* __restrict__ tab,
                                                         double* __restrict__ rec) {
  constexpr int REC = 8 + 2 * NB;
  const int m = blockIdx.y;
  const int l0 = max(m, spin);
  const int l = l0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (l > lmax) return;
  const int64_t i = soff[m] + (l - l0);
  const int64_t base = (int64_t)m * (2 * lmax + 1 - m) / 2;
  const double* tk = tab + i * 3;
  const double sig = tk[2];
  double* rk = rec + i * REC;
  const double A = tk[0], B = tk[1];
  rk[0] = A;
  rk[1] = B;
  rk[2] = -B;
  rk[3] = 0.0;
  rk[4] = -A;
  rk[5] = A + B;
  rk[6] = A - B;
  rk[7] = 0.0;
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    double2 e = alms.a[b][base + l];
    if (m == 0) e.y = 0.0;
    rk[8 + 2 * b] = e.x * sig;
    rk[9 + 2 * b] = e.y * sig;
  }
}

// base^n as mantissa in [0.5,1) times 2^e
__device__ __forceinline__ void pow_scaled(double base, int n, double& mant, int& ex) {
  int e;
  double bv = frexp(base, &e);
  int be = e;
  double rv = 1.0;
  int re = 0;
  while (n) {
    if (n & 1) {
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
    n >>= 1;
  }
  mant = rv;
  ex = re;
}

struct SpinParams {
  const LegItem* items;
  const double* rec;
  const int64_t* soff;
  const double* z;
  const double* ch;
  const double* sh;
  const int* mlim;
  const double* sn_mant;
  const int* sn_exp;
  double2* phase;            // [2][nring][mmax+1]
  int64_t phase_map_stride;
  int lmax, mmax, npair, nring, spin;
};

// EB: the NB = 2 coefficient sets are the E and B modes of ONE map pair; otherwise they are the E
// modes of NB independent map pairs (several convergence planes sheared at once), written to the
// phase maps 2b, 2b + 1.  Per l and ring pair the loop executes 6 DFMA for the two recurrences and
// 4 NB for the sums: 10 for one E-only map, 7 per map for two, 5.5 per map for four.
template <int R, int NB, bool EB, int THREADS>
__global__ void __launch_bounds__(THREADS, (R * (NB + 1) <= 8 && THREADS <= 256) ? 512 / THREADS : 1) spin_legendre_synth_kernel(const SpinParams p) {
  constexpr int REC = 8 + 2 * NB;
  constexpr int CHUNK_DOUBLES = SP_KT * REC;
  constexpr int NWARPS = THREADS / 32;
  __shared__ __align__(128) double s_rec[SP_STAGES][CHUNK_DOUBLES];
  __shared__ __align__(8) uint64_t s_full[SP_STAGES];
  __shared__ __align__(8) uint64_t s_empty[SP_STAGES];

  const int tid = threadIdx.x, lane = tid & 31;
  const LegItem item = p.items[blockIdx.x];
  const int m = item.m, s = p.spin;
  const int l0 = max(m, s);
  const int K = p.lmax - l0 + 1;  // number of l values (>= 1: items exist only if l0 <= lmax)
  const int nchunks = (K + SP_KT - 1) / SP_KT;
  const double* rec_m = p.rec + p.soff[m] * REC;

  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < SP_STAGES; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], NWARPS);
    }
    mbar_fence_init();
  }
  __syncthreads();
  auto issue = [&](int c) {
    const int st = c % SP_STAGES;
    const int kc = min(SP_KT, K - c * SP_KT);
    const uint32_t bytes = (uint32_t)(kc * REC * sizeof(double));
    mbar_arrive_expect_tx(&s_full[st], bytes);
    bulk_g2s(&s_rec[st][0], rec_m + (int64_t)c * CHUNK_DOUBLES, bytes, &s_full[st]);
  };
  if (tid == 0)
    for (int c = 0; c < SP_STAGES - 1 && c < nchunks; ++c) issue(c);

  // ---- per-ring state ----
  double pp1[R], pp2[R], pm1[R], pm2[R], xx[R];
  int sc[R];
  bool live[R];
  // accumulators: [ring][map][parity A/B][T+ re, T+ im, T- re, T- im]
  double acc[R][NB][2][4];
  const int pair0 = item.tile * (THREADS * R) + tid * R;
  const double sn_mant = p.sn_mant[m];
  const int sn_exp = p.sn_exp[m];
  const int ea = abs(m - s), eb = m + s;
  // warp-uniform recurrence variable: t = 1 - x for warps whose first ring has x >= 1/2
  const bool use_t = p.z[min(item.tile * (THREADS * R) + (tid & ~31) * R, p.npair - 1)] >= 0.5;
  const int c_off = use_t ? 4 : 0;
#pragma unroll
  for (int j = 0; j < R; ++j) {
    const int r = pair0 + j;
    live[j] = (r < p.npair) && (p.mlim[min(r, p.npair - 1)] >= m);
    pp1[j] = pp2[j] = pm1[j] = pm2[j] = xx[j] = 0.0;
    sc[j] = 0;
    if (live[j]) {
      const double ch = p.ch[r], sh = p.sh[r];
      xx[j] = use_t ? 2.0 * sh * sh : p.z[r];  // t = 1 - x = 2 sin^2(theta/2), or x
      double mca, msb, mcb, msa;
      int eca, esb, ecb, esa;
      pow_scaled(ch, ea, mca, eca);  // ch^|m-s|
      pow_scaled(sh, eb, msb, esb);  // sh^(m+s)
      pow_scaled(ch, eb, mcb, ecb);  // ch^(m+s)
      pow_scaled(sh, ea, msa, esa);  // sh^|m-s|
      double mp = sn_mant * mca * msb;  // lam+ magnitude
      double mm = sn_mant * mcb * msa;  // lam- magnitude
      const int Ep = sn_exp + eca + esb, Em = sn_exp + ecb + esa;
      if (m & 1) mp = -mp;                                     // (-1)^m
      const bool negm = (m >= s) ? (s & 1) : (m & 1);          // (-1)^s or (-1)^m
      if (negm) mm = -mm;
      const int Emax = max(Ep, Em);
      int shift = 0;
      if (Emax < 0) {
        const int q = (-Emax) / SP_SCALE_BITS;
        sc[j] = -q;
        shift = q * SP_SCALE_BITS;
      }
      pp2[j] = scalbn(mp, max(Ep + shift, -2000));
      pm2[j] = scalbn(mm, max(Em + shift, -2000));
    }
#pragma unroll
    for (int b = 0; b < NB; ++b)
#pragma unroll
      for (int q = 0; q < 2; ++q) acc[j][b][q][0] = acc[j][b][q][1] = acc[j][b][q][2] = acc[j][b][q][3] = 0.0;
  }

  const double SMALL = 7.458340731200207e-155;  // 2^-512
  int phase = 0;

  // one l step with accumulation into parity slot Q (compile time) using multiplier-selected p
  // P, Qp, Qm = {A', B', -B'} (variable x) or {-A', A'+B', A'-B'} (variable t)
  auto recur = [&](int j, double P, double Qp, double Qm) {
    const double rp = fma(xx[j], P, Qp);
    const double rm = fma(xx[j], P, Qm);
    const double tp = fma(rp, pp2[j], -pp1[j]);
    const double tm = fma(rm, pm2[j], -pm1[j]);
    pp1[j] = pp2[j];
    pp2[j] = tp;
    pm1[j] = pm2[j];
    pm2[j] = tm;
  };
  auto rescale = [&](int j) {
    if (max(sp_bexp(pp2[j]), sp_bexp(pm2[j])) >= SP_BEXP_BIG) {
      pp1[j] *= SMALL;
      pp2[j] *= SMALL;
      pm1[j] *= SMALL;
      pm2[j] *= SMALL;
      sc[j] += 1;
    }
  };

  for (int c = 0; c < nchunks; ++c) {
    const int st = c % SP_STAGES;
    if (tid == 0) {
      const int cn = c + SP_STAGES - 1;
      if (cn < nchunks) {
        if (cn >= SP_STAGES) mbar_wait(&s_empty[cn % SP_STAGES], ((cn / SP_STAGES) - 1) & 1);
        issue(cn);
      }
    }
    mbar_wait(&s_full[st], (c / SP_STAGES) & 1);
    const double* ck = &s_rec[st][0];
    const int kc = min(SP_KT, K - c * SP_KT);
    int k = 0;  // chunk-local l index; parity slot of k is (k & 1) because SP_KT is even

    if (phase == 0) {
      // blocks of 4 steps with one exponent test while every ring is still 2^-80 below the
      // significance threshold (see sht_legendre.cu): the tests are integer instructions that
      // would otherwise outnumber the 4 DFMA per ring and step
      while (k + 4 <= kc) {
        bool near = false;
#pragma unroll
        for (int j = 0; j < R; ++j)
          near |= (sc[j] == 0) && (max(sp_bexp(pp2[j]), sp_bexp(pm2[j])) >= SP_BEXP_SIG - 80);
        if (__any_sync(0xffffffffu, near)) break;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const double2 ab = *reinterpret_cast<const double2*>(ck + (k + u) * REC + c_off);
          const double qm = ck[(k + u) * REC + c_off + 2];
#pragma unroll
          for (int j = 0; j < R; ++j) recur(j, ab.x, ab.y, qm);
        }
#pragma unroll
        for (int j = 0; j < R; ++j) rescale(j);
        k += 4;
      }
      while (k < kc) {
        bool sig = false;
#pragma unroll
        for (int j = 0; j < R; ++j)
          sig |= (sc[j] == 0) && (max(sp_bexp(pp2[j]), sp_bexp(pm2[j])) >= SP_BEXP_SIG);
        if (__any_sync(0xffffffffu, sig)) {
          phase = 1;
          break;
        }
        const double2 ab = *reinterpret_cast<const double2*>(ck + k * REC + c_off);
        const double qm = ck[k * REC + c_off + 2];
#pragma unroll
        for (int j = 0; j < R; ++j) {
          recur(j, ab.x, ab.y, qm);
          rescale(j);
        }
        ++k;
      }
    }
    // generic (runtime parity) accumulate step used by the CHECK phase and odd leftovers
    auto check_step = [&](int kk) {
      const double* rk = ck + kk * REC;
      const double2 ab = *reinterpret_cast<const double2*>(rk + c_off);
      const double qm = rk[c_off + 2];
      double2 e[NB];
#pragma unroll
      for (int b = 0; b < NB; ++b) e[b] = *reinterpret_cast<const double2*>(rk + 8 + 2 * b);
      const bool odd = kk & 1;
#pragma unroll
      for (int j = 0; j < R; ++j) {
        const bool on = (sc[j] == 0);
        const double a0p = (on && !odd) ? pp2[j] : 0.0, a1p = (on && odd) ? pp2[j] : 0.0;
        const double a0m = (on && !odd) ? pm2[j] : 0.0, a1m = (on && odd) ? pm2[j] : 0.0;
#pragma unroll
        for (int b = 0; b < NB; ++b) {
          acc[j][b][0][0] = fma(a0p, e[b].x, acc[j][b][0][0]);
          acc[j][b][0][1] = fma(a0p, e[b].y, acc[j][b][0][1]);
          acc[j][b][0][2] = fma(a0m, e[b].x, acc[j][b][0][2]);
          acc[j][b][0][3] = fma(a0m, e[b].y, acc[j][b][0][3]);
          acc[j][b][1][0] = fma(a1p, e[b].x, acc[j][b][1][0]);
          acc[j][b][1][1] = fma(a1p, e[b].y, acc[j][b][1][1]);
          acc[j][b][1][2] = fma(a1m, e[b].x, acc[j][b][1][2]);
          acc[j][b][1][3] = fma(a1m, e[b].y, acc[j][b][1][3]);
        }
        recur(j, ab.x, ab.y, qm);
        rescale(j);
      }
    };
    if (phase == 1) {
      while (k < kc) {
        if ((k & 1) == 0) {  // switch to the FAST loop only at an even offset (compile-time parity there)
          bool allz = true;
#pragma unroll
          for (int j = 0; j < R; ++j) allz &= (sc[j] == 0);
          if (__all_sync(0xffffffffu, allz)) {
            phase = 2;
            break;
          }
        }
        check_step(k);
        ++k;
      }
    }
    if (phase == 2) {
      const int kend = k + ((kc - k) & ~1);
      for (; k < kend; k += 2) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const double* rk = ck + (k + q) * REC;
          const double2 ab = *reinterpret_cast<const double2*>(rk + c_off);
          const double qm = rk[c_off + 2];
          double2 e[NB];
#pragma unroll
          for (int b = 0; b < NB; ++b) e[b] = *reinterpret_cast<const double2*>(rk + 8 + 2 * b);
#pragma unroll
          for (int j = 0; j < R; ++j) {
#pragma unroll
            for (int b = 0; b < NB; ++b) {
              acc[j][b][q][0] = fma(pp2[j], e[b].x, acc[j][b][q][0]);
              acc[j][b][q][1] = fma(pp2[j], e[b].y, acc[j][b][q][1]);
              acc[j][b][q][2] = fma(pm2[j], e[b].x, acc[j][b][q][2]);
              acc[j][b][q][3] = fma(pm2[j], e[b].y, acc[j][b][q][3]);
            }
            recur(j, ab.x, ab.y, qm);
          }
        }
      }
      if (k < kc) {  // odd leftover (last chunk only)
        check_step(k);
        ++k;
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&s_empty[st]);
  }

  // ---- combine and write A_m (map 1) and B_m (map 2) for north and south ----
  const int par0 = (l0 + m + s) & 1;          // parity of (l+m+s) at the first l: slot 0 holds it
  const double sg = (m & 1) ? -1.0 : 1.0;     // (-1)^m
#pragma unroll
  for (int j = 0; j < R; ++j) {
    if (!live[j]) continue;
    const int r = pair0 + j;
    double2 An = make_double2(0, 0), Bn = An, As = An, Bs = An;
    const int64_t in = (int64_t)r * (p.mmax + 1) + m;
    const int64_t is = (int64_t)(p.nring - 1 - r) * (p.mmax + 1) + m;
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      // even / odd (l+m+s) sums of W and X
      double2 W[2], X[2];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int par = q ^ par0;  // slot q holds parity (q + par0) & 1
        W[par] = make_double2(0.5 * (acc[j][b][q][0] + sg * acc[j][b][q][2]), 0.5 * (acc[j][b][q][1] + sg * acc[j][b][q][3]));
        X[par] = make_double2(0.5 * (acc[j][b][q][0] - sg * acc[j][b][q][2]), 0.5 * (acc[j][b][q][1] - sg * acc[j][b][q][3]));
      }
      // E-type input: A = -(We +- Wo), B = +-i (Xe +- Xo)
      const double2 a_n = make_double2(-(W[0].x + W[1].x), -(W[0].y + W[1].y));
      const double2 a_s = make_double2(-(W[0].x - W[1].x), -(W[0].y - W[1].y));
      const double2 xn = make_double2(X[0].x + X[1].x, X[0].y + X[1].y);
      const double2 xs = make_double2(X[0].x - X[1].x, X[0].y - X[1].y);
      const double2 b_n = make_double2(-xn.y, xn.x);   // i * xn
      const double2 b_s = make_double2(xs.y, -xs.x);   // -i * xs
      if (!EB && NB > 1) {  // an E-only map pair of its own
        double2* ph1 = p.phase + (int64_t)(2 * b) * p.phase_map_stride;
        double2* ph2 = ph1 + p.phase_map_stride;
        ph1[in] = a_n;
        ph2[in] = b_n;
        if (r != p.npair - 1) {
          ph1[is] = a_s;
          ph2[is] = b_s;
        }
      } else if (b == 0) {
        An = a_n; Bn = b_n; As = a_s; Bs = b_s;
      } else {
        // B-mode input: map1 -= map2(B as E), map2 += map1(B as E)
        An.x -= b_n.x; An.y -= b_n.y; Bn.x += a_n.x; Bn.y += a_n.y;
        As.x -= b_s.x; As.y -= b_s.y; Bs.x += a_s.x; Bs.y += a_s.y;
      }
    }
    if (EB || NB == 1) {
      double2* ph1 = p.phase;
      double2* ph2 = p.phase + p.phase_map_stride;
      ph1[in] = An;
      ph2[in] = Bn;
      if (r != p.npair - 1) {
        ph1[is] = As;
        ph2[is] = Bs;
      }
    }
  }
}

// -------------------------------------------------------------------------------------
// host side
// -------------------------------------------------------------------------------------
static int spin_mlim(int lmax, int spin, double sth, double cth) {
  double ofs = lmax * 0.01;
  if (ofs < 100.) ofs = 100.;
  const double b = -2.0 * spin * std::fabs(cth);
  const double t1 = lmax * sth + ofs;
  const double c = (double)spin * spin - t1 * t1;
  const double discr = b * b - 4.0 * c;
  if (discr <= 0) return lmax;
  double res = (-b + std::sqrt(discr)) / 2.0;
  if (res > lmax) res = lmax;
  return (int)(res + 0.5);
}

template <typename T>
static int upload_vec(T** dptr, const std::vector<T>& h) {
  if (*dptr) cudaFree(*dptr);
  *dptr = nullptr;
  GLB_CUDA_CHECK(cudaMalloc((void**)dptr, std::max<size_t>(h.size(), 1) * sizeof(T)));
  if (!h.empty()) GLB_CUDA_CHECK(cudaMemcpy(*dptr, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  return GLB_OK;
}

int plan_ensure_spin(glb_plan* pl, int spin) {
  if (pl->spin_ready == spin) return GLB_OK;
  const int N = pl->nside, lmax = pl->lmax;
  int rc;
  // half-angle functions per ring pair, from the small quantity 1 - z in the caps
  std::vector<double> ch(pl->npair), sh(pl->npair);
  pl->h_mlim_spin.resize(pl->npair);
  for (int r = 0; r < pl->npair; ++r) {
    const int i = r + 1;
    double omz;  // 1 - z
    if (i < N)
      omz = (double)i * i / (3.0 * N * N);
    else
      omz = 1.0 - pl->h_z[r];
    sh[r] = std::sqrt(0.5 * omz);
    ch[r] = std::sqrt(1.0 - 0.5 * omz);
    pl->h_mlim_spin[r] = std::min(spin_mlim(lmax, spin, pl->h_sth[r], pl->h_z[r]), pl->mmax);
  }
  for (int r = 1; r < pl->npair; ++r) pl->h_mlim_spin[r] = std::max(pl->h_mlim_spin[r], pl->h_mlim_spin[r - 1]);
  if ((rc = upload_vec(&pl->d_ch, ch)) != GLB_OK) return rc;
  if ((rc = upload_vec(&pl->d_sh, sh)) != GLB_OK) return rc;
  if ((rc = upload_vec(&pl->d_mlim_spin, pl->h_mlim_spin)) != GLB_OK) return rc;
  // seed norm N_j sqrt((2j)!/((j+q)!(j-q)!)), j = max(m,s), q = min(m,s), as mantissa/exponent
  {
    std::vector<double> mant(pl->mmax + 1);
    std::vector<int> ex(pl->mmax + 1);
    const long double fourpi = 4.0L * 3.14159265358979323846264338327950288L;
    for (int m = 0; m <= pl->mmax; ++m) {
      const int j = std::max(m, spin), q = std::min(m, spin);
      // (2j)!/((j+q)!(j-q)!) = C(2j, j+q): product_{t=1..j-q} (j+q+t)/t, at most... j-q can be
      // large only when q = spin (m >= s), i.e. j - q = m - s terms; accumulate with frexp
      long double c = 1.0L;
      int e = 0;
      for (int t = 1; t <= j - q; ++t) {
        c *= (long double)(j + q + t) / (long double)t;
        int de;
        c = frexpl(c, &de);
        e += de;
      }
      // sqrt of (c * 2^e) * (2j+1)/(4 pi)
      long double v = c * (long double)(2 * j + 1) / fourpi;
      if (e & 1) {
        v *= 2.0L;
        e -= 1;
      }
      v = sqrtl(v);
      int de;
      v = frexpl(v, &de);
      mant[m] = (double)v;
      ex[m] = e / 2 + de;
    }
    if ((rc = upload_vec(&pl->d_sn_mant, mant)) != GLB_OK) return rc;
    if ((rc = upload_vec(&pl->d_sn_exp, ex)) != GLB_OK) return rc;
  }
  // record offsets and work list
  {
    std::vector<int64_t> soff(pl->mmax + 2);
    int64_t acc = 0;
    for (int m = 0; m <= pl->mmax; ++m) {
      soff[m] = acc;
      const int l0 = std::max(m, spin);
      if (l0 <= lmax) acc += lmax - l0 + 1;
    }
    soff[pl->mmax + 1] = acc;
    pl->nrec_spin = acc;
    if ((rc = upload_vec(&pl->d_soff, soff)) != GLB_OK) return rc;
    if (acc * 12 > pl->rec_capacity) {
      set_last_error("internal: record workspace too small for the spin transform");
      return GLB_ERR_NOMEM;
    }
    // work lists are built per tile size on demand (spin_items); drop those of another spin
    for (auto& kv : pl->spin_item_lists) cudaFree(kv.second.first);
    pl->spin_item_lists.clear();
  }
  // static recurrence tables for this spin
  cudaFree(pl->d_spin_tab);
  pl->d_spin_tab = nullptr;
  GLB_CUDA_CHECK(cudaMalloc((void**)&pl->d_spin_tab, std::max<int64_t>(pl->nrec_spin, 1) * 3 * sizeof(double)));
  spin_tables_kernel<<<(pl->mmax + 128) / 128, 128>>>(lmax, pl->mmax, spin, pl->d_soff, pl->d_spin_tab);
  GLB_CUDA_CHECK(cudaGetLastError());
  GLB_CUDA_CHECK(cudaDeviceSynchronize());
  count_launch();
  pl->spin_ready = spin;
  return GLB_OK;
}

// (m, ring tile) work list of the spin transform for tiles of T ring pairs, most expensive first
static int spin_items(glb_plan* pl, int spin, int T, LegItem** d_items, int* nitems) {
  auto it = pl->spin_item_lists.find(T);
  if (it == pl->spin_item_lists.end()) {
    const int lmax = pl->lmax;
    std::vector<int> rmin(pl->mmax + 1, pl->npair);
    int r = 0;
    for (int m = 0; m <= pl->mmax; ++m) {
      while (r < pl->npair && pl->h_mlim_spin[r] < m) ++r;
      rmin[m] = r;
    }
    const int ntile = (pl->npair + T - 1) / T;
    struct Tmp {
      LegItem it;
      double cost;
    };
    std::vector<Tmp> tmp;
    for (int m = 0; m <= pl->mmax; ++m) {
      const int l0 = std::max(m, spin);
      if (l0 > lmax) continue;
      for (int t = 0; t < ntile; ++t) {
        const int lo = std::max(t * T, rmin[m]), hi = std::min((t + 1) * T, pl->npair);
        if (hi <= lo) continue;
        tmp.push_back({{m, t}, (double)(lmax - l0 + 1) * (hi - lo)});
      }
    }
    std::stable_sort(tmp.begin(), tmp.end(), [](const Tmp& a, const Tmp& b) { return a.cost > b.cost; });
    std::vector<LegItem> items(tmp.size());
    for (size_t i = 0; i < tmp.size(); ++i) items[i] = tmp[i].it;
    LegItem* d = nullptr;
    int rc;
    if ((rc = upload_vec(&d, items)) != GLB_OK) return rc;
    it = pl->spin_item_lists.emplace(T, std::make_pair(d, (int)items.size())).first;
  }
  *d_items = it->second.first;
  *nitems = it->second.second;
  return GLB_OK;
}

template <int R, int NB, bool EB, int THREADS>
static int launch_spin_cfg(glb_plan* pl, const SpinAlms& alms, int spin, double2* d_phase, cudaStream_t st) {
  dim3 pgrid((unsigned)((pl->lmax + 1 + 255) / 256), (unsigned)(pl->mmax + 1));
  if ((int64_t)pl->nrec_spin * (8 + 2 * NB) > pl->rec_capacity) {
    set_last_error("internal: record workspace too small for the batched spin transform");
    return GLB_ERR_NOMEM;
  }
  spin_prep_kernel<NB><<<pgrid, 256, 0, st>>>(alms, pl->lmax, spin, pl->d_soff, pl->d_spin_tab, pl->d_rec);
  GLB_CUDA_CHECK(cudaGetLastError());
  count_launch();
  // phases of rings beyond mlim are never written nor read; zero the rest for m < spin rows etc.
  const int nphase = EB ? 2 : 2 * NB;
  GLB_CUDA_CHECK(cudaMemsetAsync(d_phase, 0, (size_t)nphase * pl->nring * (pl->mmax + 1) * sizeof(double2), st));
  LegItem* items = nullptr;
  int nitems = 0;
  int rc = spin_items(pl, spin, R * THREADS, &items, &nitems);
  if (rc != GLB_OK) return rc;
  if (nitems == 0) return GLB_OK;
  SpinParams p;
  p.items = items;
  p.rec = pl->d_rec;
  p.soff = pl->d_soff;
  p.z = pl->d_z;
  p.ch = pl->d_ch;
  p.sh = pl->d_sh;
  p.mlim = pl->d_mlim_spin;
  p.sn_mant = pl->d_sn_mant;
  p.sn_exp = pl->d_sn_exp;
  p.phase = d_phase;
  p.phase_map_stride = (int64_t)pl->nring * (pl->mmax + 1);
  p.lmax = pl->lmax;
  p.mmax = pl->mmax;
  p.npair = pl->npair;
  p.nring = pl->nring;
  p.spin = spin;
  spin_legendre_synth_kernel<R, NB, EB, THREADS><<<nitems, THREADS, 0, st>>>(p);
  GLB_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return GLB_OK;
}

// alm1 (E-like), alm2 (B-like, may be null) -> phases [2][nring][mmax+1]
// E-only: 4 ring pairs per thread (fewer shared-memory wavefronts per DFMA, 2 CTAs per SM);
// E+B: 2 ring pairs per thread (twice the accumulators); tiles of 1024 ring pairs either way.
int sht_spin_alm2phase(glb_plan* pl, const double2* d_alm1, const double2* d_alm2, int spin, double2* d_phase,
                       cudaStream_t st) {
  SpinAlms a = {{d_alm1, d_alm2, nullptr, nullptr}};
  if (d_alm2) return launch_spin_cfg<2, 2, true, 512>(pl, a, spin, d_phase, st);
  return launch_spin_cfg<4, 1, false, 256>(pl, a, spin, d_phase, st);
}

// E modes of nb = 2 or 4 map pairs on ONE pair of recurrences -> phases [2 nb][nring][mmax+1]
// (map pair b in phase maps 2b, 2b + 1).  Variant (tuning knob GLB_SPIN4_R, nb = 4 only): ring pairs
// per thread, 2 (256 threads, 8 warps per SM) or 1 (512 threads, 16 warps per SM); tiles of 512.
int sht_spin_alm2phase_multi(glb_plan* pl, const double2* const* d_alms, int nb, int spin, double2* d_phase,
                             cudaStream_t st) {
  SpinAlms a = {{nullptr, nullptr, nullptr, nullptr}};
  for (int b = 0; b < nb; ++b) a.a[b] = d_alms[b];
  if (nb == 2) return launch_spin_cfg<2, 2, false, 512>(pl, a, spin, d_phase, st);
  if (nb == 4) {
    static const int r4 = [] {
      const char* e = getenv("GLB_SPIN4_R");
      return e ? atoi(e) : 2;
    }();
    if (r4 == 1) return launch_spin_cfg<1, 4, false, 512>(pl, a, spin, d_phase, st);
    return launch_spin_cfg<2, 4, false, 256>(pl, a, spin, d_phase, st);
  }
  set_last_error("internal: batched spin transform takes 2 or 4 maps");
  return GLB_ERR_INVALID_ARG;
}

}  // namespace glb
