// sht_ozaki.cu -- Legendre stage of the scalar synthesis with the contraction on the INT8 tensor cores.
//
// Same mathematics as sht_legendre.cu (replaces the Legendre part of healpy.alm2map, glass/healpix.py:71):
//     F_m(+-x) = sum_k p_k(x^2) (Ae_k +- x Ao_k),   p_{k+1} = (a_k x^2 + b_k) p_k - p_{k-1}
// but only the 2 DFMA of the recurrence stay on the FP64 pipe.  The contraction over k
//     D[ring, c] = sum_k p_k(ring) A_k[c]        c = (Ae_re, Ae_im, Ao_re, Ao_im) x maps
// is an Ozaki-type exact integer product on tcgen05.mma kind::i8 with int32 accumulators in TMEM:
//
//   * per (ring, tile of 64 l-pairs) the values p_k are cut into six base-256 digits relative to the
//     largest |p_k| of the tile:  V = rint(p s) + BIAS comes out of ONE FMA with the magic constant
//     2^52 + BIAS (the low 48 mantissa bits are V), and with BIAS = sum_j 128 256^j the BYTES u_j of V
//     are the balanced digits d_j = u_j - 128 in [-128, 127] -- as an int8 that is u_j ^ 0x80: no
//     shifts, masks or carries, only byte permutes (a 4 x 4 byte transpose per four l-pairs) and one
//     XOR per word;
//   * the coefficients A_k[c] are cut the same way per (column, tile) by a preparation kernel;
//   * digit products with i + j >= 5 (21 of 36) are accumulated by significance s = i + j into six
//     column groups of D:  p digit a against the coefficient digits 5-a..5 is ONE MMA of N = NC (a+1)
//     columns; 12 MMAs (M = 128 rings, K = 32) per tile; |D| <= 6 64 2^14 < 2^23;
//   * epilogue per tile: D from TMEM (thread = TMEM lane = ring), pairs of groups combined in int32
//     (< 2^31), three FP64 operations per column, accumulated in FP64 registers with the two scales.
//
// Accuracy: exact integer arithmetic on 48-bit operands; the dropped digit products are below 2^-46 of
// max |p| max |A| per term (emulated step by step in tests/studies/ozaki_device_scheme.py).  One thread owns one ring pair (128 per CTA), walks l in
// two passes per tile (pass 1: recurrence only, finds the tile's largest |p|; pass 2: the same
// recurrence again, digits to shared memory in the tensor core's K-major core-matrix layout).
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "plan.h"
#include "sht_seed.cuh"
#include "sht_tables.cuh"

namespace glb {

constexpr int OZ_KT = 64;      // l-pairs per tile
constexpr int OZ_ND = 6;       // base-256 digits per operand
constexpr int OZ_ROWS = 128;   // ring pairs per CTA = threads = TMEM lanes
constexpr int OZ_STAGES = 3;   // tile blocks in flight (TMA)
constexpr int OZ_SUP = 8;      // tiles per super-tile: one set of column scales, D may accumulate across them
constexpr int OZ_A_SLICE = (OZ_KT / 16) * OZ_ROWS * 16;  // bytes of one digit plane of the p operand
constexpr int OZ_A_BYTES = OZ_ND * OZ_A_SLICE;

// tile block in global memory (one bulk copy): {a, b, -a, a+b}[64] | coefficient digits | inverse
// column scales | the tile's power of two
template <int NC>
struct OzTile {
  static constexpr int AB_BYTES = OZ_KT * 4 * 8;
  static constexpr int BOP_LBO = OZ_ND * NC * 16;           // bytes between the 16-byte k chunks
  static constexpr int BOP_BYTES = BOP_LBO * (OZ_KT / 16);
  static constexpr int INVA_OFF = AB_BYTES + BOP_BYTES;     // double[NC]
  static constexpr int PSC_OFF = INVA_OFF + NC * 8;         // double: the tile's power of two 2^sig (see oz_prep_kernel)
  static constexpr int BYTES = PSC_OFF + 32;
  static_assert(BYTES % 32 == 0, "bulk copies move multiples of 16 bytes, the coefficient rows are stored as double4");
};

constexpr double OZ_HEADROOM = 0.99;  // |V| <= 0.99 * 2^47 keeps V + BIAS inside 48 bits
__host__ __device__ constexpr int64_t oz_bias() {
  int64_t b = 0;
  for (int j = 0; j < OZ_ND; ++j) b += (int64_t)128 << (8 * j);
  return b;
}
constexpr double OZ_MAGIC = 4503599627370496.0 + (double)oz_bias();  // 2^52 + BIAS

// scale s = 0.99 * 2^(47 - e) and its inverse for |x| < 2^e, from the biased exponent field eb of the
// largest |x| (e = eb - 1022); tiny maxima (eb < 64) count as zero
__device__ __forceinline__ void oz_scales(int eb, double& s, double& inv) {
  if (eb < 64) {
    s = 0.0;
    inv = 0.0;
  } else {
    s = __hiloint2double((2092 - eb) << 20, 0) * OZ_HEADROOM;
    inv = __hiloint2double((eb - 46) << 20, 0) * (1.0 / OZ_HEADROOM);
  }
}

// ---- tcgen05 / TMEM PTX ------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst) {  // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "n"(COLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {  // the allocating warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&v)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor, K-major, no swizzle: 8 rows x 16 bytes core matrices, `sbo` bytes
// between 8-row groups, `lbo` bytes between the two 16-byte k chunks of one K = 32 instruction
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3ffff) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) |
         ((uint64_t)1 << 46);
}
// instruction descriptor of kind::i8: D = s32, A = B = s8, both K-major, M = 128
__host__ __device__ constexpr uint32_t umma_idesc_i8(int n) {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(OZ_ROWS >> 4) << 24);
}
__device__ __forceinline__ void umma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// -------------------------------------------------------------------------------------
// preparation: Legendre records (sht_prep_kernel<4>, 16 columns) -> tile blocks.  One CTA of 64
// threads per (super-tile of OZ_SUP tiles, m); thread = l-pair of a tile.
//
// Scales.  The rescaling alpha_k of the recurrence (lambda = alpha_k p_k, sht_tables.cuh) drifts
// with l, so p_k and the coefficients A_k = alpha_k a_lm drift in opposite directions.  Every tile
// therefore carries ONE power of two 2^sig (the exponent of alpha at its first l-pair): the
// coefficients are cut as A 2^-sig, the kernel cuts p 2^sig -- both then vary like the physical
// lambda_lm and a_lm -- and the product is unchanged.  On top of that the COLUMN scales are common
// to the super-tile (largest |A 2^-sig| of its 512 l-pairs), so that the integer sums of its tiles
// can be added up in TMEM.  col0 = first column of this pass (B = 8: two passes).
// -------------------------------------------------------------------------------------
template <int NC>
__global__ void __launch_bounds__(OZ_KT) oz_prep_kernel(const double* __restrict__ rec, const double* __restrict__ tab,
                                                        const int64_t* __restrict__ roff, const int64_t* __restrict__ toff,
                                                        int lmax, int col0, uint8_t* __restrict__ oz) {
  using T = OzTile<NC>;
  constexpr int REC = 4 + 16;
  __shared__ int s_max[16];
  __shared__ __align__(16) int8_t s_dig[OZ_ND * 16][OZ_KT];  // [digit * 16 + column][k]
  const int m = blockIdx.y, k = threadIdx.x;
  const int K = (lmax - m) / 2 + 1;
  const int t0 = blockIdx.x * OZ_SUP;
  if (t0 * OZ_KT >= K) return;
  const int t1 = min(t0 + OZ_SUP, (K + OZ_KT - 1) / OZ_KT);
  const double* rm = rec + roff[m] * REC;
  const double* am = tab + roff[m] * PREP_TAB + TAB_ALPHA;
  auto tile_field = [&](int t) { return (__double2hiint(am[(int64_t)t * OZ_KT * PREP_TAB]) >> 20) & 0x7ff; };  // exponent field of alpha
  if (k < 16) s_max[k] = 0;
  __syncthreads();
  // largest |A 2^-sig| per column over the super-tile
  {
    int mx[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) mx[c] = 0;
    for (int t = t0; t < t1; ++t) {
      const int kk = t * OZ_KT + k;
      if (kk >= K) break;
      const double down = __hiloint2double((2046 - tile_field(t)) << 20, 0);  // 2^-sig
      const double* r = rm + (int64_t)kk * REC;
#pragma unroll
      for (int c = 0; c < 16; ++c) mx[c] = max(mx[c], __double2hiint(r[4 + c] * down) & 0x7fffffff);
    }
#pragma unroll
    for (int c = 0; c < 16; ++c) atomicMax(&s_max[c], mx[c]);
  }
  __syncthreads();
  for (int t = t0; t < t1; ++t) {
    const int kk = t * OZ_KT + k;
    const bool valid = kk < K;
    const double* r = rm + (int64_t)kk * REC;
    uint8_t* blk = oz + (toff[m] + t) * (int64_t)T::BYTES;
    const int field = tile_field(t);
    const double down = __hiloint2double((2046 - field) << 20, 0);
    if (col0 == 0) {
      double4 ab = valid ? *reinterpret_cast<const double4*>(r) : make_double4(0.0, 0.0, 0.0, 0.0);
      reinterpret_cast<double4*>(blk)[k] = ab;
      if (k == 0) *reinterpret_cast<double*>(blk + T::PSC_OFF) = __hiloint2double(field << 20, 0);  // 2^sig
    }
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      double sc, inv;
      oz_scales(s_max[c] >> 20, sc, inv);
      const double v = valid ? r[4 + c] * down : 0.0;
      const double tt = fma(v, sc, OZ_MAGIC);
      const uint32_t lo = (uint32_t)__double2loint(tt), hi = (uint32_t)__double2hiint(tt);
#pragma unroll
      for (int j = 0; j < OZ_ND; ++j) {
        const uint32_t u = j < 4 ? (lo >> (8 * j)) & 255u : (hi >> (8 * (j - 4))) & 255u;
        s_dig[j * 16 + c][k] = (int8_t)(u ^ 0x80u);
      }
      if (k == 0) reinterpret_cast<double*>(blk + T::INVA_OFF)[col0 + c] = inv * 1099511627776.0;  // 256^5: digit products with i + j >= 5
    }
    __syncthreads();
    // digit planes to the block in the tensor core's layout: [k chunk][digit * NC + column][16 bytes]
    for (int i = k; i < OZ_ND * 16 * (OZ_KT / 16); i += OZ_KT) {
      const int row = i >> 2, ch = i & 3;
      const int j = row >> 4, c = row & 15;
      const uint4 q = *reinterpret_cast<const uint4*>(&s_dig[row][ch * 16]);
      *reinterpret_cast<uint4*>(blk + T::AB_BYTES + ch * T::BOP_LBO + (j * NC + col0 + c) * 16) = q;
    }
    __syncthreads();
  }
}

struct OzParams {
  const LegItem* items;
  const uint8_t* oz;
  const int64_t* toff;
  const double* z;
  const double* sth;
  const int* mlim;
  const double* cm_mant;
  const int* cm_exp;
  double2* phase;
  int64_t phase_map_stride;  // in double2
  int lmax, mmax, npair, nring;
  double bound_pole, bound_c;  // |lambda_lm(theta)| <= min(bound_pole, bound_c / sqrt(sin theta)) (with margin)
  unsigned long long* dbg;  // event counters (development), or null
};

// the six digit planes of four consecutive values: a 4 x 4 byte transpose of the low words, a 4 x 2
// one of the high words (__byte_perm), and the sign flip of the balanced digits
__device__ __forceinline__ void oz_planes(const double (&tt)[4], uint32_t (&w)[OZ_ND]) {
  const uint32_t l0 = (uint32_t)__double2loint(tt[0]), l1 = (uint32_t)__double2loint(tt[1]);
  const uint32_t l2 = (uint32_t)__double2loint(tt[2]), l3 = (uint32_t)__double2loint(tt[3]);
  const uint32_t h0 = (uint32_t)__double2hiint(tt[0]), h1 = (uint32_t)__double2hiint(tt[1]);
  const uint32_t h2 = (uint32_t)__double2hiint(tt[2]), h3 = (uint32_t)__double2hiint(tt[3]);
  const uint32_t t0 = __byte_perm(l0, l1, 0x5140), t1 = __byte_perm(l0, l1, 0x7362);
  const uint32_t t2 = __byte_perm(l2, l3, 0x5140), t3 = __byte_perm(l2, l3, 0x7362);
  const uint32_t s0 = __byte_perm(h0, h1, 0x5140), s1 = __byte_perm(h2, h3, 0x5140);
  w[0] = __byte_perm(t0, t2, 0x5410) ^ 0x80808080u;
  w[1] = __byte_perm(t0, t2, 0x7632) ^ 0x80808080u;
  w[2] = __byte_perm(t1, t3, 0x5410) ^ 0x80808080u;
  w[3] = __byte_perm(t1, t3, 0x7632) ^ 0x80808080u;
  w[4] = __byte_perm(s0, s1, 0x5410) ^ 0x80808080u;
  w[5] = __byte_perm(s0, s1, 0x7632) ^ 0x80808080u;
}

// int32 -> double without a conversion instruction: the mantissa of 2^52 + 2^31 + x
__device__ __forceinline__ double oz_i2d(int x) {
  return __hiloint2double(0x43300000, (int)((uint32_t)x ^ 0x80000000u)) - 4503601774854144.0;
}

// pass 2 of one half tile (32 l-pairs from k0): the recurrence, six digit planes of every value to
// the operand buffer (row `dst`), and the largest |p| that was cut (high word).  FAST: every ring of
// the warp is at scale 0 (no range tests); FULL: all 32 l-pairs exist.
template <bool FAST, bool FULL>
__device__ __forceinline__ void oz_pass2_half(const double* __restrict__ ab, int k0, int kc, double x2, double scale,
                                              double& p1, double& p2, int& sc, int& maxhi, uint8_t* __restrict__ dst) {
  const double SMALL = 7.458340731200207e-155;  // 2^-512
#pragma unroll 1
  for (int ch = 0; ch < 2; ++ch) {
    uint32_t w[4][OZ_ND];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      double tt[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int k = k0 + ch * 16 + q * 4 + i;
        const bool in = FULL || k < kc;
        const double v = (in && (FAST || sc == 0)) ? p2 : 0.0;
        maxhi = max(maxhi, __double2hiint(v) & 0x7fffffff);
        tt[i] = fma(v, scale, OZ_MAGIC);
        if (in) {
          const double2 c2 = *reinterpret_cast<const double2*>(ab + 4 * k);
          const double rr = fma(c2.x, x2, c2.y);
          const double tn = fma(rr, p2, -p1);
          p1 = p2;
          p2 = tn;
          if (!FAST && bexp(p2) >= BEXP_BIG) {
            p1 *= SMALL;
            p2 *= SMALL;
            sc += 1;
          }
        }
      }
      oz_planes(tt, w[q]);
    }
#pragma unroll
    for (int j = 0; j < OZ_ND; ++j)
      *reinterpret_cast<uint4*>(dst + j * OZ_A_SLICE + ch * (OZ_ROWS * 16)) = make_uint4(w[0][j], w[1][j], w[2][j], w[3][j]);
  }
}

// recurrence alone over l-pairs [k0, k1) of a tile, every ring at scale 0: largest |p| that enters the sum
__device__ __forceinline__ int oz_scan_fast(const double* __restrict__ ab, int k0, int k1, double x2, double& p1, double& p2) {
  int maxhi = 0;
#pragma unroll 4
  for (int k = k0; k < k1; ++k) {
    const double2 c2 = *reinterpret_cast<const double2*>(ab + 4 * k);
    maxhi = max(maxhi, __double2hiint(p2) & 0x7fffffff);
    const double rr = fma(c2.x, x2, c2.y);
    const double tn = fma(rr, p2, -p1);
    p1 = p2;
    p2 = tn;
  }
  return maxhi;
}

// Ring scales.  |lambda_lm(theta)| <= min(sqrt((2l+1)/4pi), c / sqrt(sin theta)) and p 2^sig is lambda up to a
// factor below two, so ONE absolute bound per ring serves every tile (2^3 on most rings: 48-bit
// fixed point with an absolute error of 2^-45 per value, which is what matters for a map whose error
// is measured against its largest pixel).  Fixed scales mean that the integer sums of consecutive tiles can stay
// in TMEM (a "run", at most one super-tile = one set of column scales) and D is read and converted
// once per run ("flush") instead of once per tile.  Every cut checks the bound half tile by half
// tile, BEFORE the half's MMAs are issued; a ring that exceeds it (never observed) closes the run --
// everything summed so far is valid -- and the half is cut again with its exact range.
// A warp whose rings are all still negligible stays silent: it only runs the recurrence (pass 1),
// and cuts the tile in a second pass once a ring becomes significant; from the first tile cut with
// every ring at scale 0 on it goes straight to the cut.  The tile's two halves (32 l-pairs, one
// K = 32 instruction per digit each) are pipelined: the MMAs of a half run while the threads cut
// the next one.
template <int NC>
__global__ void __launch_bounds__(OZ_ROWS) sht_legendre_ozaki_kernel(const OzParams p) {
  using T = OzTile<NC>;
  constexpr int B = NC / 4;
  constexpr int TMEM_COLS = (OZ_ND * NC <= 128) ? 128 : 256;
  constexpr int HALF = 2 * OZ_ROWS * 16;  // bytes of one half tile inside a digit plane
  extern __shared__ __align__(128) uint8_t oz_smem[];
  uint8_t* sA = oz_smem;
  uint8_t* sT = oz_smem + OZ_A_BYTES;
  uint64_t* s_full = reinterpret_cast<uint64_t*>(sT + OZ_STAGES * T::BYTES);
  uint64_t* s_mma = s_full + OZ_STAGES;  // [h]: the MMAs of half h issued so far are done
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_mma + 2);
  int* s_flag = reinterpret_cast<int*>(s_tmem + 1);  // [3] votes of the CTA-wide decisions, rotating

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const LegItem item = p.items[blockIdx.x];
  const int m = item.m;
  const int K = (p.lmax - m) / 2 + 1;
  const int ntiles = (K + OZ_KT - 1) / OZ_KT;
  const uint8_t* blk_m = p.oz + p.toff[m] * (int64_t)T::BYTES;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < OZ_STAGES; ++s) mbar_init(&s_full[s], 1);
    mbar_init(&s_mma[0], 1);
    mbar_init(&s_mma[1], 1);
    mbar_fence_init();
    s_flag[0] = s_flag[1] = s_flag[2] = 0;
  }
  if (warp == 0) tmem_alloc<TMEM_COLS>(s_tmem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;

  auto issue = [&](int t) {
    const int s = t % OZ_STAGES;
    mbar_arrive_expect_tx(&s_full[s], (uint32_t)T::BYTES);
    bulk_g2s(sT + s * T::BYTES, blk_m + (int64_t)t * T::BYTES, (uint32_t)T::BYTES, &s_full[s]);
  };
  if (tid == 0) {
    for (int t = 0; t < OZ_STAGES - 1 && t < ntiles; ++t) issue(t);
  }

  // ---- per-thread ring state ----
  const int r = item.tile * OZ_ROWS + tid;
  const bool live = (r < p.npair) && (p.mlim[min(r, p.npair - 1)] >= m);
  const double zw = p.z[min(item.tile * OZ_ROWS + (tid & ~31), p.npair - 1)];
  const bool use_u = zw * zw >= 0.5;  // the warp's recurrence variable (sht_legendre.cu)
  const int ab_off = use_u ? 2 : 0;
  double p1 = 0.0, p2 = 0.0, x2 = 0.0, zz = 0.0;
  int sc = 0, eg_ring = 1023;
  if (live) {
    zz = p.z[r];
    const double sth = p.sth[r];
    x2 = use_u ? sth * sth : zz * zz;
    lam_mm_scaled(m, sth, p.cm_mant[m], p.cm_exp[m], p2, sc);
    // exponent field bounding |p 2^sig| on this ring: 2^(eg_ring - 1022) > min(bound_pole, bound_c / sqrt(sin theta))
    eg_ring = ((__double2hiint(fmin(p.bound_pole, p.bound_c * rsqrt(sth))) >> 20) & 0x7ff) + 1;
  }
  double acc[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) acc[c] = 0.0;

  const double SMALL = 7.458340731200207e-155;  // 2^-512
  int nc0 = 0, nc1 = 0;            // commits so far on s_mma[0], s_mma[1] (CTA-uniform)
  int nsync = 0;                   // votes so far (CTA-uniform)
  bool run_open = false;           // CTA-uniform: D holds sums that have not been read
  bool run_rows = false;           // warp-uniform: this warp has rows in them
  int run_stage = 0;               // a resident tile block of the run's super-tile (column scales)
  double run_inv = 0.0;            // inverse ring scale of the run
  int eg = 0;                      // exponent field bounding this ring's |p 2^sig| in the run
  double scale = 0.0;
  bool awake = false;              // warp-uniform: the warp has cut a tile with every ring at scale 0 (no pass 1 any more)

  // CTA-wide OR of a few bits (one barrier)
  auto vote = [&](int bits) -> int {
    int* f = &s_flag[nsync % 3];
    bits = __reduce_or_sync(0xffffffffu, bits);
    if (lane == 0 && bits) atomicOr(f, bits);
    __syncthreads();
    const int res = *f;
    if (tid == 0) s_flag[(nsync + 2) % 3] = 0;
    ++nsync;
    return res;
  };
  // every MMA issued so far is complete
  auto mma_wait_all = [&]() {
    if (nc0) mbar_wait(&s_mma[0], (uint32_t)((nc0 - 1) & 1));
    if (nc1) mbar_wait(&s_mma[1], (uint32_t)((nc1 - 1) & 1));
  };
  // close the run: D -> FP64 accumulators (a barrier follows before the next MMA is issued)
  auto count = [&](int i) {
    if (p.dbg && lane == 0) atomicAdd(p.dbg + i, 1ull);
  };
  auto flush = [&]() {
    count(0 + (run_rows ? 1 : 0));  // warp-flushes without / with rows
    mma_wait_all();
    tc_fence_after();
    if (run_rows) {
      const uint8_t* tp = sT + run_stage * T::BYTES;
      const double* invA = reinterpret_cast<const double*>(tp + T::INVA_OFF);
      const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
#pragma unroll
      for (int c0 = 0; c0 < NC; c0 += 8) {
        uint32_t d[OZ_ND][8];
#pragma unroll
        for (int g = 0; g < OZ_ND; ++g) tmem_ld8(lane_base + g * NC + c0, d[g]);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          double val = oz_i2d((int)d[OZ_ND - 1][i]);
#pragma unroll
          for (int g = OZ_ND - 2; g >= 0; --g) val = fma(val, 256.0, oz_i2d((int)d[g][i]));
          acc[c0 + i] = fma(val, run_inv * invA[c0 + i], acc[c0 + i]);
        }
      }
    }
    tc_fence_before();
    run_open = false;
    run_rows = false;
  };
  // the six MMAs of one half tile (one thread)
  auto mma_half = [&](const uint8_t* tp, int h, bool first) {
    const uint32_t a0 = smem_u32(sA) + h * HALF, b0 = smem_u32(tp + T::AB_BYTES) + h * 2 * T::BOP_LBO;
#pragma unroll
    for (int a = OZ_ND - 1; a >= 0; --a) {
      const uint64_t ad = umma_desc(a0 + a * OZ_A_SLICE, OZ_ROWS * 16, 128);
      const uint64_t bd = umma_desc(b0 + NC * (OZ_ND - 1 - a) * 16, T::BOP_LBO, 128);
      umma_i8(tmem, ad, bd, umma_idesc_i8(NC * (a + 1)), (first && a == OZ_ND - 1) ? 0u : 1u);
    }
  };

  for (int t = 0; t < ntiles; ++t) {
    const int stage = t % OZ_STAGES;
    mbar_wait(&s_full[stage], (uint32_t)((t / OZ_STAGES) & 1));
    const uint8_t* tp = sT + stage * T::BYTES;
    const double* ab = reinterpret_cast<const double*>(tp) + ab_off;
    const int kc = min(OZ_KT, K - t * OZ_KT);
    const bool full = kc == OZ_KT;
    uint8_t* dst = sA + tid * 16;
    const double psc = *reinterpret_cast<const double*>(tp + T::PSC_OFF);  // the tile's 2^sig: p is cut as p 2^sig
    const int sig = ((__double2hiint(psc) >> 20) & 0x7ff) - 1023;
    const bool newsup = (t % OZ_SUP) == 0;

    const double s_p1 = p1, s_p2 = p2;
    const int s_sc = sc;
    const bool fast = __all_sync(0xffffffffu, sc == 0);
    const bool has = run_open && run_rows;  // this warp has valid sums in D
    // ---- this warp's part in the tile: cut (emit) or silent; the ring scale is the fixed absolute one
    //      (p.eg_fix bounds every |lambda_lm|) unless a value should ever exceed it ----
    bool one, emit;
    int need = 0;
    if (awake && fast && full) {
      one = true;  // straight to the cut; the range is checked half by half
      emit = true;
      count(2);
    } else {
      // ---- pass 1: the recurrence alone; largest |p| that enters the sum (scale 0) ----
      one = false;
      int maxhi = 0;
      if (fast) {
        maxhi = oz_scan_fast(ab, 0, kc, x2, p1, p2);
      } else {
        int k = 0;
        // far below significance: blocks of four steps with one range test (sht_legendre.cu, SKIP phase)
        while (k + 4 <= kc) {
          const bool near = (sc == 0) && (bexp(p2) >= BEXP_SIG - 80);
          if (__any_sync(0xffffffffu, near)) break;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const double2 c2 = *reinterpret_cast<const double2*>(ab + 4 * (k + u));
            const double rr = fma(c2.x, x2, c2.y);
            const double tn = fma(rr, p2, -p1);
            p1 = p2;
            p2 = tn;
          }
          if (bexp(p2) >= BEXP_BIG) {
            p1 *= SMALL;
            p2 *= SMALL;
            sc += 1;
          }
          k += 4;
        }
#pragma unroll 2
        for (; k < kc; ++k) {
          const double2 c2 = *reinterpret_cast<const double2*>(ab + 4 * k);
          if (sc == 0) maxhi = max(maxhi, __double2hiint(p2) & 0x7fffffff);
          const double rr = fma(c2.x, x2, c2.y);
          const double tn = fma(rr, p2, -p1);
          p1 = p2;
          p2 = tn;
          if (bexp(p2) >= BEXP_BIG) {
            p1 *= SMALL;
            p2 *= SMALL;
            sc += 1;
          }
        }
      }
      emit = __any_sync(0xffffffffu, (maxhi >> 20) >= BEXP_SIG);
      count(emit ? (fast ? 4 : 5) : 6);
      need = maxhi ? (maxhi >> 20) + sig : 0;
      if (emit) {
        p1 = s_p1;
        p2 = s_p2;
        sc = s_sc;
        awake = awake || fast;
      }
    }
    int eg_new = max(need, eg_ring);
    double scale_new = scale, inv_new = run_inv;
    if (eg_new != eg) oz_scales(eg_new, scale_new, inv_new);
    // D must be read first if this warp's rows change scale, appear or disappear, or the column scales change
    const bool brk = emit ? (run_open && !(has && eg_new == eg && !newsup)) : has;
    double sc_t = scale_new * psc;

    // ---- first half (operand half 0 is free once the MMAs issued on it are done) ----
    int seen = 0;  // largest |p| actually cut
    if (nc0) mbar_wait(&s_mma[0], (uint32_t)((nc0 - 1) & 1));
    if (emit) {
      if (fast && full)
        oz_pass2_half<true, true>(ab, 0, kc, x2, sc_t, p1, p2, sc, seen, dst);
      else
        oz_pass2_half<false, false>(ab, 0, kc, x2, sc_t, p1, p2, sc, seen, dst);
      fence_async_smem();
    }
    bool bad = one && seen != 0 && (seen >> 20) + sig > eg_new;  // a ring outgrew the guessed range
    const int v = vote((emit ? 1 : 0) | (bad ? 2 : 0) | (brk ? 8 : 0));
    const bool any = (v & 1) != 0;
    if (__any_sync(0xffffffffu, bad)) count(7);
    if (brk) count(8);
    if (v & 10) {
      // nothing of this tile has been issued: close the run with the old scales
      if (run_open) flush();
      if ((v & 2) && __any_sync(0xffffffffu, bad)) {
        // exact range of the whole tile for this warp
        p1 = s_p1;
        p2 = s_p2;
        const int mx = oz_scan_fast(ab, 0, kc, x2, p1, p2);
        eg_new = mx ? (mx >> 20) + sig : 0;
        oz_scales(eg_new, scale_new, inv_new);
        sc_t = scale_new * psc;
        p1 = s_p1;
        p2 = s_p2;
        seen = 0;
        oz_pass2_half<true, true>(ab, 0, kc, x2, sc_t, p1, p2, sc, seen, dst);
        fence_async_smem();
        one = false;
      }
      __syncthreads();
    }
    eg = eg_new;
    scale = scale_new;
    run_inv = inv_new;
    if (any) {
      if (tid == 0) {
        tc_fence_after();
        mma_half(tp, 0, !run_open);
        umma_commit(&s_mma[0]);
      }
      ++nc0;
      run_open = true;
      run_stage = stage;
      run_rows = run_rows || emit;
    }

    // ---- second half ----
    const double h_p1 = p1, h_p2 = p2;
    if (nc1) mbar_wait(&s_mma[1], (uint32_t)((nc1 - 1) & 1));
    if (emit && kc > 32) {
      if (fast && full)
        oz_pass2_half<true, true>(ab, 32, kc, x2, sc_t, p1, p2, sc, seen, dst + HALF);
      else
        oz_pass2_half<false, false>(ab, 32, kc, x2, sc_t, p1, p2, sc, seen, dst + HALF);
      fence_async_smem();
    }
    bad = one && seen != 0 && (seen >> 20) + sig > eg;
    if (__any_sync(0xffffffffu, bad)) count(9);
    const int redo = __syncthreads_or(bad ? 1 : 0);
    if (any) {
      if (redo) {
        // the first half is in D with the old scales: close the run, then the second half with its exact range
        flush();
        if (__any_sync(0xffffffffu, bad)) {
          p1 = h_p1;
          p2 = h_p2;
          const int mx = oz_scan_fast(ab, 32, kc, x2, p1, p2);
          eg = mx ? (mx >> 20) + sig : 0;
          oz_scales(eg, scale, run_inv);
          sc_t = scale * psc;
          p1 = h_p1;
          p2 = h_p2;
          int dummy = 0;
          oz_pass2_half<true, true>(ab, 32, kc, x2, sc_t, p1, p2, sc, dummy, dst + HALF);
          fence_async_smem();
        }
        __syncthreads();
      }
      if (kc > 32) {
        if (tid == 0) {
          tc_fence_after();
          mma_half(tp, 1, !run_open);
          umma_commit(&s_mma[1]);
        }
        ++nc1;
        run_open = true;
        run_rows = run_rows || emit;
      }
    }
    if (tid == 0 && t + OZ_STAGES - 1 < ntiles) issue(t + OZ_STAGES - 1);
  }
  if (run_open) flush();
  __syncthreads();
  if (warp == 0) tmem_dealloc<TMEM_COLS>(tmem);

  // ---- F_m of the north and south ring of the pair ----
  if (live) {
    const int slot = m;
    const int W = p.mmax + 1;
    const int rs = p.nring - 1 - r;
#pragma unroll
    for (int b = 0; b < B; ++b) {
      const double er = acc[4 * b + 0], ei = acc[4 * b + 1];
      const double orr = acc[4 * b + 2] * zz, oi = acc[4 * b + 3] * zz;
      double2* ph = p.phase + b * p.phase_map_stride;
      ph[(int64_t)r * W + slot] = make_double2(er + orr, ei + oi);
      if (r != p.npair - 1) ph[(int64_t)rs * W + slot] = make_double2(er - orr, ei - oi);
    }
  }
}

// -------------------------------------------------------------------------------------
// host side
// -------------------------------------------------------------------------------------
int plan_items(glb_plan* pl, int tile, int G, int rank, LegItem** d_items, int* nitems);
int sht_prep_group(glb_plan* pl, const double2* d_alm, int nb, cudaStream_t st);

template <int NC>
static int ozaki_prep(glb_plan* pl, const double2* d_alm, cudaStream_t st) {
  using T = OzTile<NC>;
  constexpr int B = NC / 4;
  // tile offsets per m
  if (!pl->d_oz_toff) {
    std::vector<int64_t> toff(pl->mmax + 2, 0);
    for (int m = 0; m <= pl->mmax; ++m) toff[m + 1] = toff[m] + ((pl->lmax - m) / 2 + 1 + OZ_KT - 1) / OZ_KT;
    pl->oz_tiles = toff[pl->mmax + 1];
    GLB_CUDA_CHECK(cudaMalloc((void**)&pl->d_oz_toff, toff.size() * sizeof(int64_t)));
    GLB_CUDA_CHECK(cudaMemcpy(pl->d_oz_toff, toff.data(), toff.size() * sizeof(int64_t), cudaMemcpyHostToDevice));
  }
  const int64_t need = pl->oz_tiles * (int64_t)T::BYTES;
  if (pl->oz_bytes < need) {
    cudaFree(pl->d_oz);
    pl->d_oz = nullptr;
    pl->oz_bytes = 0;
    if (cudaMalloc((void**)&pl->d_oz, (size_t)need) != cudaSuccess) {
      cudaGetLastError();
      set_last_error("out of device memory for the INT8 Legendre tile blocks");
      return GLB_ERR_NOMEM;
    }
    pl->oz_bytes = need;
  }
  if ((int64_t)pl->nrec * (4 + 16) > pl->rec_capacity) {
    set_last_error("the INT8 Legendre path needs a plan with max_batch >= 4");
    return GLB_ERR_INVALID_ARG;
  }
  const int ntile_max = (pl->lmax / 2 + 1 + OZ_KT - 1) / OZ_KT;
  for (int g = 0; g < B / 4; ++g) {
    int rc = sht_prep_group(pl, d_alm + (int64_t)g * 4 * pl->nalm, 4, st);
    if (rc != GLB_OK) return rc;
    oz_prep_kernel<NC><<<dim3((ntile_max + OZ_SUP - 1) / OZ_SUP, pl->mmax + 1), OZ_KT, 0, st>>>(pl->d_rec, pl->d_prep_tab, pl->d_roff, pl->d_oz_toff, pl->lmax,
                                                                       16 * g, pl->d_oz);
    GLB_CUDA_CHECK(cudaGetLastError());
    count_launch();
  }
  return GLB_OK;
}

template <int NC>
static int ozaki_legendre(glb_plan* pl, double2* d_phase, cudaStream_t st) {
  using T = OzTile<NC>;
  OzParams p;
  int nitems = 0;
  LegItem* items = nullptr;
  int rc = plan_items(pl, OZ_ROWS, 1, 0, &items, &nitems);
  if (rc != GLB_OK) return rc;
  p.items = items;
  p.oz = pl->d_oz;
  p.toff = pl->d_oz_toff;
  p.z = pl->d_z;
  p.sth = pl->d_sth;
  p.mlim = pl->d_mlim;
  p.cm_mant = pl->d_cm_mant;
  p.cm_exp = pl->d_cm_exp;
  p.phase = d_phase;
  p.phase_map_stride = (int64_t)pl->nring * (pl->mmax + 1);
  p.lmax = pl->lmax;
  p.mmax = pl->mmax;
  p.npair = pl->npair;
  p.nring = pl->nring;
  // |lambda_lm| <= sqrt((2l+1)/4pi) everywhere; away from the poles the Airy peak at the turning point is the
  // largest value: |lambda_lm| sqrt(sin theta) < 0.54 lmax^(1/6) (measured with the oracle up to lmax 2047: 1.69 at
  // lmax 1023); a quarter on top for the drift of alpha inside a tile.  The kernel checks the bound on every value.
  p.bound_pole = 1.25 * std::sqrt((2.0 * pl->lmax + 1.0) / (4.0 * 3.14159265358979323846));
  p.bound_c = 1.25 * 0.6 * std::pow((double)std::max(pl->lmax, 1), 1.0 / 6.0);
  static unsigned long long* d_dbg = nullptr;
  const bool debug = getenv("GLB_OZ_DEBUG") != nullptr;
  if (debug && !d_dbg) GLB_CUDA_CHECK(cudaMalloc((void**)&d_dbg, 16 * sizeof(unsigned long long)));
  if (debug) GLB_CUDA_CHECK(cudaMemsetAsync(d_dbg, 0, 16 * sizeof(unsigned long long), st));
  p.dbg = debug ? d_dbg : nullptr;
  const size_t smem = OZ_A_BYTES + OZ_STAGES * T::BYTES + (OZ_STAGES + 2) * sizeof(uint64_t) + 32;
  static bool attr_set = false;
  if (!attr_set) {
    GLB_CUDA_CHECK(cudaFuncSetAttribute(sht_legendre_ozaki_kernel<NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  if (nitems > 0) {
    sht_legendre_ozaki_kernel<NC><<<nitems, OZ_ROWS, smem, st>>>(p);
    GLB_CUDA_CHECK(cudaGetLastError());
    count_launch();
  }
  if (debug) {
    unsigned long long h[16];
    GLB_CUDA_CHECK(cudaMemcpyAsync(h, d_dbg, sizeof(h), cudaMemcpyDeviceToHost, st));
    GLB_CUDA_CHECK(cudaStreamSynchronize(st));
    fprintf(stderr, "[oz NC=%d] warp-tiles: one-pass %llu, two-pass fast %llu / checked %llu, silent %llu | warp flushes %llu (+%llu without rows), bad half0 %llu half1 %llu, brk %llu\n", NC, h[2], h[4], h[5], h[6], h[1], h[0], h[7], h[9], h[8]);
  }
  return GLB_OK;
}

// the two halves of the INT8 Legendre stage, nb in {4, 8}: alm [nb][nalm] -> tile blocks (records, digit planes), and
// tile blocks -> phase [nb][nring][mmax+1]
int sht_ozaki_prep(glb_plan* pl, const double2* d_alm, int nb, cudaStream_t st) {
  if (nb == 4) return ozaki_prep<16>(pl, d_alm, st);
  if (nb == 8) return ozaki_prep<32>(pl, d_alm, st);
  set_last_error("the INT8 Legendre path takes 4 or 8 maps");
  return GLB_ERR_INVALID_ARG;
}
int sht_ozaki_legendre(glb_plan* pl, int nb, double2* d_phase, cudaStream_t st) {
  if (nb == 4) return ozaki_legendre<16>(pl, d_phase, st);
  if (nb == 8) return ozaki_legendre<32>(pl, d_phase, st);
  set_last_error("the INT8 Legendre path takes 4 or 8 maps");
  return GLB_ERR_INVALID_ARG;
}
int sht_alm2phase_ozaki(glb_plan* pl, const double2* d_alm, int nb, double2* d_phase, cudaStream_t st) {
  const int rc = sht_ozaki_prep(pl, d_alm, nb, st);
  if (rc != GLB_OK) return rc;
  return sht_ozaki_legendre(pl, nb, d_phase, st);
}

}  // namespace glb
