// sht_ozaki.cu -- Legendre stage of the scalar synthesis with the contraction on the INT8 tensor cores.
//
// Same mathematics as sht_legendre.cu (replaces the Legendre part of healpy.alm2map, glass/healpix.py:71):
//     F_m(+-x) = sum_k p_k(x^2) (Ae_k +- x Ao_k),   p_{k+1} = (a_k x^2 + b_k) p_k - p_{k-1}
// but only the 2 DFMA of the recurrence stay on the FP64 pipe.  The contraction over k
//     D[ring, c] = sum_k p_k(ring) A_k[c]        c = (Ae_re, Ae_im, Ao_re, Ao_im) x maps
// is an Ozaki-type exact integer product on tcgen05.mma kind::i8 with int32 accumulators in TMEM:
//
//   * every value p_k is turned into 48-bit fixed point by ONE FMA with a magic constant:
//     t = fma(p, s, 2^52 + BIAS) has V = rint(p s) + BIAS in its low mantissa bits, and with
//     BIAS = sum_j 128 256^j the BYTES u_j of V are the balanced base-256 digits d_j = u_j - 128 in
//     [-128, 127] -- as an int8 that is u_j ^ 0x80: no shifts, masks or carries, only byte permutes (a
//     4 x 4 byte transpose per four l-pairs) and one XOR per word put the six digit planes into shared
//     memory in the tensor core's K-major core-matrix layout;
//   * the coefficients A_k[c] are cut the same way by a preparation kernel (oz_prep_kernel) into tile
//     blocks of 64 l-pairs that one TMA bulk copy brings in;
//   * digit products with i + j >= 5 (21 of 36) are accumulated by significance s = i + j into six
//     column groups of D:  p digit a against the coefficient digits 5-a..5 is ONE MMA of N = NC (a+1)
//     columns; 12 MMAs (M = 128 rings, K = 32) per tile; |D| <= 6 64 2^14 < 2^23 per tile;
//   * the scales are FIXED -- one absolute bound per ring (|lambda_lm| <= min(sqrt((2l+1)/4pi),
//     c / sqrt(sin theta))), one power of two per tile that follows the rescaling alpha_k of the
//     recurrence, column scales per super-tile of 16 tiles -- so the integer sums of consecutive tiles
//     stay in TMEM and D is read, converted and scaled once per super-tile.
//
// Accuracy: exact integer arithmetic on 48-bit operands; the dropped digit products are below 2^-46 of
// max |p| max |A| per term (emulated step by step in tests/studies/ozaki_device_scheme.py); measured
// 4-8e-12 of the largest phase at nside 4096 against the FP64 kernel.  The comment above the kernel
// describes the scales, the roles of the warps and the synchronisation; DESIGN.md 3.1b the measurements.
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "plan.h"
#include "oz_digits.cuh"
#include "sht_seed.cuh"
#include "sht_tables.cuh"

namespace glb {

constexpr int OZ_KT = 64;      // l-pairs per tile
constexpr int OZ_ROWS = 128;   // ring pairs per CTA = threads = TMEM lanes
constexpr int OZ_STAGES = 3;   // tile blocks in flight (TMA)
constexpr int OZ_SUP = 16;     // tiles per super-tile: one set of column scales, D may accumulate across them
constexpr int OZ_PERK = 2;     // the first tiles of every m carry one power of two per l-pair (alpha falls fast there)
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
  static constexpr int PSC_OFF = INVA_OFF + NC * 8;         // double: the tile's power of two 2^sig (see oz_prep_kernel), 0: per l-pair
  static constexpr int SK_OFF = PSC_OFF + 32;               // double[64]: powers of two per l-pair (tiles with PSC = 0)
  static constexpr int BYTES = SK_OFF + OZ_KT * 8;
  static_assert(BYTES % 32 == 0, "bulk copies move multiples of 16 bytes, the coefficient rows are stored as double4");
};

// mbarrier wait with a watchdog: a protocol error becomes a launch failure instead of a hung GPU
template <bool BACKOFF = false>
__device__ __forceinline__ void oz_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  long long t0 = 0;
  for (uint32_t spin = 0;; ++spin) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (ok) return;
    if (BACKOFF) __nanosleep(200);  // the issuing thread: do not take issue slots from the producers of its sub-partition
    if ((spin & 1023u) == 1023u) {
      if (t0 == 0) t0 = clock64();
      else if (clock64() - t0 > 4000000000ll) __trap();  // two seconds in one wait: never in a correct run
    }
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
__device__ __forceinline__ void tmem_st8_zero(uint32_t taddr) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr), "r"(0u) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
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
// therefore carries ONE power of two 2^sig (the exponent of alpha at its first l-pair; the first two
// tiles of every m, where alpha falls by a factor of ten, carry one per l-pair): the
// coefficients are cut as A 2^-sig, the kernel cuts p 2^sig -- both then vary like the physical
// lambda_lm and a_lm -- and the product is unchanged.  On top of that the COLUMN scales are common
// to the super-tile (largest |A 2^-sig| of its 1024 l-pairs), so that the integer sums of its tiles
// can be added up in TMEM.  col0 = first column of this pass (B = 8: two passes).
// -------------------------------------------------------------------------------------
template <int NC>
__global__ void __launch_bounds__(OZ_KT) oz_prep_kernel(const double* __restrict__ rec, const double* __restrict__ tab,
                                                        const int64_t* __restrict__ roff, const int64_t* __restrict__ toff,
                                                        int lmax, int col0, uint8_t* __restrict__ oz) {
  using T = OzTile<NC>;
  constexpr int REC = 4 + 16;
  __shared__ int s_max[16];
  __shared__ __align__(16) double s_rec[OZ_KT][REC + 1];     // one tile of records (+1: the rows are read column-wise)
  __shared__ __align__(16) int8_t s_dig[OZ_ND * 16][OZ_KT];  // [digit * 16 + column][k]
  const int m = blockIdx.y, k = threadIdx.x;
  const int K = (lmax - m) / 2 + 1;
  const int t0 = blockIdx.x * OZ_SUP;
  if (t0 * OZ_KT >= K) return;
  const int t1 = min(t0 + OZ_SUP, (K + OZ_KT - 1) / OZ_KT);
  const double* rm = rec + roff[m] * REC;
  const double* am = tab + roff[m] * PREP_TAB + TAB_ALPHA;
  // exponent field of alpha: per l-pair in the first OZ_PERK tiles (alpha falls by a factor of ten there), of the
  // tile's first l-pair afterwards (a few per cent of drift per tile)
  auto field_of = [&](int t, int kk) {
    const int kr = t < OZ_PERK ? min(kk, K - 1) : t * OZ_KT;
    return (__double2hiint(am[(int64_t)kr * PREP_TAB]) >> 20) & 0x7ff;
  };
  // coalesced copy of tile t's records (contiguous in global memory) into s_rec, zero rows beyond K
  auto stage_tile = [&](int t) {
    const int n = min(OZ_KT, K - t * OZ_KT) * REC;  // doubles
    const double* src = rm + (int64_t)t * OZ_KT * REC;
    for (int i = k; i < OZ_KT * REC; i += OZ_KT) s_rec[i / REC][i % REC] = i < n ? src[i] : 0.0;
  };
  if (k < 16) s_max[k] = 0;
  // largest |A 2^-sig| per column over the super-tile
  int mx[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) mx[c] = 0;
  for (int t = t0; t < t1; ++t) {
    __syncthreads();
    stage_tile(t);
    __syncthreads();
    const double down = __hiloint2double((2046 - field_of(t, t * OZ_KT + k)) << 20, 0);  // 2^-sig
#pragma unroll
    for (int c = 0; c < 16; ++c) mx[c] = max(mx[c], __double2hiint(s_rec[k][4 + c] * down) & 0x7fffffff);
  }
#pragma unroll
  for (int c = 0; c < 16; ++c) atomicMax(&s_max[c], mx[c]);
  for (int t = t0; t < t1; ++t) {
    __syncthreads();
    stage_tile(t);
    __syncthreads();
    uint8_t* blk = oz + (toff[m] + t) * (int64_t)T::BYTES;
    const int field = field_of(t, t * OZ_KT + k);
    const double down = __hiloint2double((2046 - field) << 20, 0);
    if (col0 == 0) {
      reinterpret_cast<double4*>(blk)[k] = make_double4(s_rec[k][0], s_rec[k][1], s_rec[k][2], s_rec[k][3]);
      reinterpret_cast<double*>(blk + T::SK_OFF)[k] = __hiloint2double(field << 20, 0);  // 2^sig of this l-pair
      if (k == 0) *reinterpret_cast<double*>(blk + T::PSC_OFF) = t < OZ_PERK ? 0.0 : __hiloint2double(field << 20, 0);
    }
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      double sc, inv;
      oz_scales(s_max[c] >> 20, sc, inv);
      const double tt = fma(s_rec[k][4 + c] * down, sc, OZ_MAGIC);
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

// cut of one half tile (32 l-pairs from k0): the recurrence, six digit planes of every value to the
// operand buffer (row `dst`).  Range check: every FMA result must lie in [2^52, 2^52 + 2^48), i.e.
// its high word is 0x4330xxxx -- `hor` / `hand` collect the OR and the AND of the high words.
// FAST: every ring of the warp is at scale 0 (no range tests); FULL: all 32 l-pairs exist.
template <bool FAST, bool FULL, bool PERK = false>
__device__ __forceinline__ void oz_cut_half(const double* __restrict__ ab, int k0, int kc, double x2, double scale,
                                            double& p1, double& p2, int& sc, uint32_t& hor, uint32_t& hand,
                                            uint8_t* __restrict__ dst, const double* __restrict__ sk = nullptr) {
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
        tt[i] = fma(v, PERK ? scale * sk[k] : scale, OZ_MAGIC);
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
      const uint32_t h0 = (uint32_t)__double2hiint(tt[0]), h1 = (uint32_t)__double2hiint(tt[1]);
      const uint32_t h2 = (uint32_t)__double2hiint(tt[2]), h3 = (uint32_t)__double2hiint(tt[3]);
      hor = (hor | h0 | h1) | (h2 | h3);
      hand = (hand & h0 & h1) & (h2 & h3);
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
// is measured against its largest pixel).  Fixed scales mean that the integer sums of consecutive
// tiles stay in TMEM and D is read and converted once per super-tile (= one set of column scales,
// 1024 l-pairs) instead of once per tile.
//
// Roles.  Four PRODUCER warps own 32 rings each (thread = ring = TMEM lane): they run the recurrence
// and cut the digits of their rows, half tile by half tile (32 l-pairs = one K = 32 instruction per
// digit), into the operand buffer.  A fifth warp ISSUES: one thread waits until the four producers
// have written a half (mbarrier), issues its six MMAs, commits them to the mbarrier that frees the
// half for the next tile, and keeps the TMA pipeline of tile blocks filled.  There is no CTA-wide
// barrier in the loop: a warp is only ever held up by the operand half it wants to overwrite.
//   * A warp whose rings are all still negligible stays silent: its operand rows are zero (its rows
//     of D therefore stay zero), it only runs the recurrence (pass 1) and cuts the tile in a second
//     pass once a ring becomes significant; after that it always goes straight to the cut.
//   * Reading D ("flush") is PRIVATE to a warp: it waits for the MMAs issued so far, loads its 32
//     lanes, adds them to its FP64 accumulators with its ring scales and the super-tile's column
//     scales, and clears the lanes (tcgen05.st) -- the MMAs always accumulate.  Every warp flushes
//     at a super-tile boundary before it cuts the first half of the new super-tile; the issuer
//     cannot run ahead of that because it needs all four warps' halves.
//   * Every cut checks the bound (oz_in_range).  A ring that exceeds it (never observed) makes its
//     warp flush, cut the half again with the exact range of its values, and flush once more when it
//     returns to the fixed scale.
constexpr int OZ_THREADS = OZ_ROWS + 32;
#ifdef GLB_OZ_TIMERS
constexpr bool OZ_TIMERS = true;  // development: where the producer warps and the issuer spend their cycles (GLB_OZ_DEBUG=1)
#else
constexpr bool OZ_TIMERS = false;
#endif

template <int NC>
__global__ void __launch_bounds__(OZ_THREADS) sht_legendre_ozaki_kernel(const OzParams p) {
  using T = OzTile<NC>;
  constexpr int B = NC / 4;
  constexpr int TMEM_COLS = (OZ_ND * NC <= 128) ? 128 : 256;
  constexpr int HALF = 2 * OZ_ROWS * 16;  // bytes of one half tile inside a digit plane
  extern __shared__ __align__(128) uint8_t oz_smem[];
  uint8_t* sA = oz_smem;
  uint8_t* sT = oz_smem + OZ_A_BYTES;
  uint64_t* s_full = reinterpret_cast<uint64_t*>(sT + OZ_STAGES * T::BYTES);
  uint64_t* s_mma = s_full + OZ_STAGES;   // [h]: the MMAs of half h of a tile are done (or there were none)
  uint64_t* s_ready = s_mma + 2;          // [h]: the four producer warps have written half h of a tile
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_ready + 2);
  int* s_emit = reinterpret_cast<int*>(s_tmem + 1);  // [2][2] (tile parity, half): a producer cut this half

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
    mbar_init(&s_ready[0], OZ_ROWS / 32);
    mbar_init(&s_ready[1], OZ_ROWS / 32);
    mbar_fence_init();
    s_emit[0] = s_emit[1] = s_emit[2] = s_emit[3] = 0;
  }
  if (warp == 0) tmem_alloc<TMEM_COLS>(s_tmem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;

  if (warp == OZ_ROWS / 32) {
    // =========================== issuer ===========================
    if (lane == 0) {
      auto issue = [&](int t) {
        const int s = t % OZ_STAGES;
        mbar_arrive_expect_tx(&s_full[s], (uint32_t)T::BYTES);
        bulk_g2s(sT + s * T::BYTES, blk_m + (int64_t)t * T::BYTES, (uint32_t)T::BYTES, &s_full[s]);
      };
      for (int t = 0; t < OZ_STAGES - 1 && t < ntiles; ++t) issue(t);
      bool first = true;  // D has not been written yet
      long long iw = 0, ii = 0, iq = (OZ_TIMERS && p.dbg) ? clock64() : 0;
      for (int t = 0; t < ntiles; ++t) {
        const int stage = t % OZ_STAGES;
        const uint8_t* tp = sT + stage * T::BYTES;
        const uint32_t par = (uint32_t)(t & 1);
        const int kc = min(OZ_KT, K - t * OZ_KT);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          oz_wait<true>(&s_ready[h], par);
          if (OZ_TIMERS && p.dbg) {
            const long long now = clock64();
            iw += now - iq;
            iq = now;
          }
          if (h == 0 && t + OZ_STAGES - 1 < ntiles) {
            // stage (t - 1) % OZ_STAGES: its MMAs are done, the producers have moved on to tile t
            if (t > 0) oz_wait(&s_mma[1], par ^ 1u);
            issue(t + OZ_STAGES - 1);
          }
          int* flag = &s_emit[(t & 1) * 2 + h];
          const bool any = *reinterpret_cast<volatile int*>(flag) != 0 && (h == 0 || kc > 32);
          *flag = 0;
          if (any) {
            oz_wait(&s_full[stage], (uint32_t)((t / OZ_STAGES) & 1));  // (the producers have seen it already)
            tc_fence_after();
            const uint32_t a0 = smem_u32(sA) + h * HALF, b0 = smem_u32(tp + T::AB_BYTES) + h * 2 * T::BOP_LBO;
#pragma unroll
            for (int a = OZ_ND - 1; a >= 0; --a) {
              const uint64_t ad = umma_desc(a0 + a * OZ_A_SLICE, OZ_ROWS * 16, 128);
              const uint64_t bd = umma_desc(b0 + NC * (OZ_ND - 1 - a) * 16, T::BOP_LBO, 128);
              umma_i8(tmem, ad, bd, umma_idesc_i8(NC * (a + 1)), (first && a == OZ_ND - 1) ? 0u : 1u);
            }
            first = false;
            umma_commit(&s_mma[h]);
          } else {
            mbar_arrive(&s_mma[h]);
          }
          if (OZ_TIMERS && p.dbg) {
            const long long now = clock64();
            ii += now - iq;
            iq = now;
          }
        }
      }
      if (OZ_TIMERS && p.dbg) {
        atomicAdd(p.dbg + 8, (unsigned long long)iw);
        atomicAdd(p.dbg + 9, (unsigned long long)ii);
      }
    }
  } else {
    // =========================== producers ===========================
    const int r = item.tile * OZ_ROWS + tid;
    const bool live = (r < p.npair) && (p.mlim[min(r, p.npair - 1)] >= m);
    const double zw = p.z[min(item.tile * OZ_ROWS + (tid & ~31), p.npair - 1)];
    const bool use_u = zw * zw >= 0.5;  // the warp's recurrence variable (sht_legendre.cu)
    const int ab_off = use_u ? 2 : 0;
    double p1 = 0.0, p2 = 0.0, x2 = 0.0, zz = 0.0;
    int sc = 0, eg_ring = 1023;
    double bound = 1.0;
    if (live) {
      zz = p.z[r];
      const double sth = p.sth[r];
      x2 = use_u ? sth * sth : zz * zz;
      lam_mm_scaled(m, sth, p.cm_mant[m], p.cm_exp[m], p2, sc);
      // bound on |p 2^sig| on this ring and its exponent field: 2^(eg_ring - 1022) > bound
      bound = fmin(p.bound_pole, p.bound_c * rsqrt(sth));
      eg_ring = (__double2hiint(bound) >> 20) & 0x7ff;
    }
    double acc[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) acc[c] = 0.0;
    uint8_t* dst = sA + tid * 16;
    // silent rows are zero digits
#pragma unroll
    for (int j = 0; j < OZ_ND; ++j)
#pragma unroll
      for (int ch = 0; ch < OZ_KT / 16; ++ch)
        *reinterpret_cast<uint4*>(dst + j * OZ_A_SLICE + ch * (OZ_ROWS * 16)) = make_uint4(0u, 0u, 0u, 0u);
    fence_async_smem();

    const double SMALL = 7.458340731200207e-155;  // 2^-512
    auto count = [&](int i) {
      if (p.dbg && lane == 0) atomicAdd(p.dbg + i, 1ull);
    };
    long long tm[6] = {0, 0, 0, 0, 0, 0}, tq = 0;  // development: cycles in wait-TMA, pass 1, wait-operand, cut, fence + arrive, flush
    auto tick = [&](int i) {
      if (OZ_TIMERS && p.dbg) {
        const long long now = clock64();
        tm[i] += now - tq;
        tq = now;
      }
    };
    if (OZ_TIMERS && p.dbg) tq = clock64();
    bool awake = false;   // warp-uniform: the warp cuts (its rows of D are in use)
    bool dirty = false;   // warp-uniform: its lanes of D hold sums that have not been read
    int eg = 0;           // exponent field bounding this ring's |p 2^sig| at the scale in use
    double scale = 0.0, run_inv = 0.0;
    int run_stage = 0;    // a resident tile block of the current super-tile (column scales)
    const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);

    // this warp's lanes of D -> FP64 accumulators, lanes cleared.  `t`, `h`: the last half whose MMAs were issued.
    auto flush = [&](int t, int h) {
      count(1);
      oz_wait(&s_mma[h], (uint32_t)(t & 1));  // in order: everything issued before it is complete as well
      tc_fence_after();
      const double* invA = reinterpret_cast<const double*>(sT + run_stage * T::BYTES + T::INVA_OFF);
#pragma unroll
      for (int c0 = 0; c0 < NC; c0 += 8) {
        uint32_t d[OZ_ND][8];
#pragma unroll
        for (int g = 0; g < OZ_ND; ++g) tmem_ld8(lane_base + g * NC + c0, d[g]);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < OZ_ND; ++g) tmem_st8_zero(lane_base + g * NC + c0);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          double val = oz_i2d((int)d[OZ_ND - 1][i]);
#pragma unroll
          for (int g = OZ_ND - 2; g >= 0; --g) val = fma(val, 256.0, oz_i2d((int)d[g][i]));
          acc[c0 + i] = fma(val, run_inv * invA[c0 + i], acc[c0 + i]);
        }
      }
      tmem_st_wait();
      tc_fence_before();
      dirty = false;
    };

    for (int t = 0; t < ntiles; ++t) {
      const int stage = t % OZ_STAGES;
      oz_wait(&s_full[stage], (uint32_t)((t / OZ_STAGES) & 1));
      tick(0);
      const uint8_t* tp = sT + stage * T::BYTES;
      const double* ab = reinterpret_cast<const double*>(tp) + ab_off;
      const int kc = min(OZ_KT, K - t * OZ_KT);
      const bool full = kc == OZ_KT;
      const double psc = *reinterpret_cast<const double*>(tp + T::PSC_OFF);  // the tile's 2^sig: p is cut as p 2^sig
      const bool perk = psc == 0.0;                                          // ... or one power of two per l-pair
      const double* sk = reinterpret_cast<const double*>(tp + T::SK_OFF);
      const int sig = perk ? 0 : ((__double2hiint(psc) >> 20) & 0x7ff) - 1023;
      const uint32_t prev = (uint32_t)((t & 1) ^ 1);

      // ---- new column scales: read what the old ones produced ----
      if (dirty && (t % OZ_SUP) == 0) flush(t - 1, 1);  // (column scales of tile t - 1, whose block is still resident)
      run_stage = stage;
      tick(5);

      const double s_p1 = p1, s_p2 = p2;
      const int s_sc = sc;
      const bool fast = __all_sync(0xffffffffu, sc == 0);
      bool emit = awake;
      if (!awake) {
        // ---- pass 1: the recurrence alone; does any ring become significant in this tile? ----
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
        count(emit ? 4 : 6);
        if (emit) {
          p1 = s_p1;
          p2 = s_p2;
          sc = s_sc;
          awake = true;
        }
      } else {
        count(2);
      }
      // the fixed scale of the ring (after an excursion to an exact range: read D first)
      if (__any_sync(0xffffffffu, emit && eg != eg_ring)) {  // (warp-uniform: the flush is a warp-wide operation)
        if (dirty) flush(t - 1, 1);
        eg = eg_ring;
        scale = (OZ_HEADROOM * 140737488355328.0) / bound;  // 0.99 * 2^47 / bound: the ring's fixed scale
        run_inv = bound * (1.0 / (OZ_HEADROOM * 140737488355328.0));
      }
      double sc_t = perk ? scale : scale * psc;
      tick(1);

#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        const int k0 = 32 * h;
        if (t > 0) oz_wait(&s_mma[h], prev);  // the operand half is free: the MMAs of the previous tile have read it
        tick(2);
        const bool work = emit && k0 < kc;
        if (work) {
          const double h_p1 = p1, h_p2 = p2;
          const int h_sc = sc;
          uint32_t hor = 0u, hand = 0xffffffffu;
          if (perk && fast && full)
            oz_cut_half<true, true, true>(ab, k0, kc, x2, sc_t, p1, p2, sc, hor, hand, dst + h * HALF, sk);
          else if (perk)
            oz_cut_half<false, false, true>(ab, k0, kc, x2, sc_t, p1, p2, sc, hor, hand, dst + h * HALF, sk);
          else if (fast && full)
            oz_cut_half<true, true>(ab, k0, kc, x2, sc_t, p1, p2, sc, hor, hand, dst + h * HALF);
          else
            oz_cut_half<false, false>(ab, k0, kc, x2, sc_t, p1, p2, sc, hor, hand, dst + h * HALF);
          if (__any_sync(0xffffffffu, !oz_in_range(hor, hand))) {
            // a value beyond the bound: read D at the old scale, then this half with the exact range of its values
            count(7);
            if (dirty) flush(h == 0 ? t - 1 : t, h == 0 ? 1 : 0);
            p1 = h_p1;
            p2 = h_p2;
            sc = h_sc;
            int mx = 0;
            for (int k = k0; k < min(k0 + 32, kc); ++k) {
              const double2 c2 = *reinterpret_cast<const double2*>(ab + 4 * k);
              if (sc == 0) mx = max(mx, __double2hiint(p2) & 0x7fffffff);
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
            eg = max(mx ? (mx >> 20) + sig : 0, eg);
            oz_scales(eg, scale, run_inv);
            sc_t = perk ? scale : scale * psc;
            p1 = h_p1;
            p2 = h_p2;
            sc = h_sc;
            if (perk)
              oz_cut_half<false, false, true>(ab, k0, kc, x2, sc_t, p1, p2, sc, hor, hand, dst + h * HALF, sk);
            else
              oz_cut_half<false, false>(ab, k0, kc, x2, sc_t, p1, p2, sc, hor, hand, dst + h * HALF);
          }
          tick(3);
          fence_async_smem();
          dirty = true;
        }
        __syncwarp();
        if (lane == 0) {
          if (work) atomicOr(&s_emit[(t & 1) * 2 + h], 1);
          __threadfence_block();
          mbar_arrive(&s_ready[h]);
        }
        tick(4);
      }
    }
    if (dirty) flush(ntiles - 1, (min(OZ_KT, K - (ntiles - 1) * OZ_KT) > 32) ? 1 : 0);
    tick(5);
    if (OZ_TIMERS && p.dbg && lane == 0)
      for (int i = 0; i < 6; ++i) atomicAdd(p.dbg + 10 + i, (unsigned long long)tm[i]);

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
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<TMEM_COLS>(tmem);
}

// -------------------------------------------------------------------------------------
// host side
// -------------------------------------------------------------------------------------
int plan_items(glb_plan* pl, int tile, int G, int rank, LegItem** d_items, int* nitems);
int sht_prep_group(glb_plan* pl, const double2* d_alm, int nb, cudaStream_t st);

template <int NC>
static int ozaki_prep(glb_plan* pl, const double2* d_alm, cudaStream_t st, int slot) {
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
  uint8_t*& d_oz = slot == 0 ? pl->d_oz : (slot == 1 ? pl->d_oz1 : pl->d_oz2);
  int64_t& oz_bytes = slot == 0 ? pl->oz_bytes : (slot == 1 ? pl->oz_bytes1 : pl->oz_bytes2);
  if (oz_bytes < need) {
    cudaFree(d_oz);
    d_oz = nullptr;
    oz_bytes = 0;
    if (cudaMalloc((void**)&d_oz, (size_t)need) != cudaSuccess) {
      cudaGetLastError();
      set_last_error("out of device memory for the INT8 Legendre tile blocks");
      return GLB_ERR_NOMEM;
    }
    oz_bytes = need;
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
                                                                       16 * g, d_oz);
    GLB_CUDA_CHECK(cudaGetLastError());
    count_launch();
  }
  return GLB_OK;
}

template <int NC>
static int ozaki_legendre(glb_plan* pl, double2* d_phase, cudaStream_t st, int slot) {
  using T = OzTile<NC>;
  OzParams p;
  int nitems = 0;
  LegItem* items = nullptr;
  int rc = plan_items(pl, OZ_ROWS, 1, 0, &items, &nitems);
  if (rc != GLB_OK) return rc;
  p.items = items;
  p.oz = slot == 0 ? pl->d_oz : (slot == 1 ? pl->d_oz1 : pl->d_oz2);
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
  // lmax 1023); a fifth on top for the drift of alpha inside a tile.  The kernel checks the bound on every value.
  p.bound_pole = 1.2 * std::sqrt((2.0 * pl->lmax + 1.0) / (4.0 * 3.14159265358979323846));
  p.bound_c = 1.2 * 0.6 * std::pow((double)std::max(pl->lmax, 1), 1.0 / 6.0);
  static unsigned long long* d_dbg = nullptr;
  const bool debug = getenv("GLB_OZ_DEBUG") != nullptr;
  if (debug && !d_dbg) GLB_CUDA_CHECK(cudaMalloc((void**)&d_dbg, 16 * sizeof(unsigned long long)));
  if (debug) GLB_CUDA_CHECK(cudaMemsetAsync(d_dbg, 0, 16 * sizeof(unsigned long long), st));
  p.dbg = debug ? d_dbg : nullptr;
  const size_t smem = OZ_A_BYTES + OZ_STAGES * T::BYTES + (OZ_STAGES + 4) * sizeof(uint64_t) + 32;
  static bool attr_set = false;
  if (!attr_set) {
    GLB_CUDA_CHECK(cudaFuncSetAttribute(sht_legendre_ozaki_kernel<NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  if (nitems > 0) {
    sht_legendre_ozaki_kernel<NC><<<nitems, OZ_THREADS, smem, st>>>(p);
    GLB_CUDA_CHECK(cudaGetLastError());
    count_launch();
  }
  if (debug) {
    unsigned long long h[16];
    GLB_CUDA_CHECK(cudaMemcpyAsync(h, d_dbg, sizeof(h), cudaMemcpyDeviceToHost, st));
    GLB_CUDA_CHECK(cudaStreamSynchronize(st));
    fprintf(stderr, "[oz NC=%d] warp-tiles: cut directly %llu, woken up %llu, silent %llu | warp flushes %llu, halves beyond the bound %llu\n", NC, h[2], h[4], h[6], h[1], h[7]);
    if (OZ_TIMERS) {
    const double tot = (double)(h[10] + h[11] + h[12] + h[13] + h[14] + h[15]);
    fprintf(stderr, "[oz NC=%d] producer warp cycles: wait TMA %.1f%%, mode / pass 1 %.1f%%, wait operand half %.1f%%, cut %.1f%%, fence + arrive %.1f%%, flush %.1f%% of %.3g; issuer: waiting %.3g, issuing %.3g cycles\n", NC, 100 * h[10] / tot, 100 * h[11] / tot, 100 * h[12] / tot, 100 * h[13] / tot, 100 * h[14] / tot, 100 * h[15] / tot, tot, (double)h[8], (double)h[9]);
    }
  }
  return GLB_OK;
}

// the two halves of the INT8 Legendre stage, nb in {4, 8}: alm [nb][nalm] -> tile blocks (records, digit planes), and
// tile blocks -> phase [nb][nring][mmax+1]
int sht_ozaki_prep(glb_plan* pl, const double2* d_alm, int nb, cudaStream_t st, int slot) {
  if (nb == 4) return ozaki_prep<16>(pl, d_alm, st, slot);
  if (nb == 8) return ozaki_prep<32>(pl, d_alm, st, slot);
  set_last_error("the INT8 Legendre path takes 4 or 8 maps");
  return GLB_ERR_INVALID_ARG;
}
int sht_ozaki_legendre(glb_plan* pl, int nb, double2* d_phase, cudaStream_t st, int slot) {
  if (nb == 4) return ozaki_legendre<16>(pl, d_phase, st, slot);
  if (nb == 8) return ozaki_legendre<32>(pl, d_phase, st, slot);
  set_last_error("the INT8 Legendre path takes 4 or 8 maps");
  return GLB_ERR_INVALID_ARG;
}
int sht_alm2phase_ozaki(glb_plan* pl, const double2* d_alm, int nb, double2* d_phase, cudaStream_t st) {
  const int rc = sht_ozaki_prep(pl, d_alm, nb, st, 2);
  if (rc != GLB_OK) return rc;
  return sht_ozaki_legendre(pl, nb, d_phase, st, 2);
}

}  // namespace glb
