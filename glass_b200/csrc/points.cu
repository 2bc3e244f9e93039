// points.cu -- galaxy counts and in-pixel positions (K6, K7, K8).
//
// Replaces, per population (glass/points.py:520-540):
//   n  = bias_model(delta, b)                                  points.py:243-249 (linear :134, loglinear :157-160)
//   n  = (n [- mean(n)] + 1) * ARCMIN2_SPHERE/npix * ngal      points.py:279-288
//   n *= vis                                                   points.py:314-316
//   n  = poisson(clip(n, 0))                                   points.py:340-348
//   ipix = repeat(arange(start, stop), n[start:stop])          points.py:426
//   lon, lat = randang(nside, ipix, lonlat=True)               points.py:427 -> healpix.py:426-431
//
// K6 fuses bias, normalisation, visibility, clip and the Poisson draw, and K7 (exclusive pixel
// offsets; the batch cuts of points.py:409-424 are then a searchsorted on that array) rides in
// the same pass as a chained scan with decoupled look-back; K8 walks pixels and
// writes every galaxy's (lon, lat) at its offset, galaxies ordered by ring pixel index as
// in the reference.  All arithmetic that the reference does in NumPy is rounded the same
// way (separate multiplies/adds, no FMA contraction) so that the expected-count map is
// bit-identical for the linear / no-bias models.
#include <algorithm>

#include "common.cuh"
#include "healpix_geom.cuh"
#include "points_cuts.cuh"
#include "rng.cuh"

namespace glb {

constexpr int PT_THREADS = 256;
#ifndef GLB_PT_ITEMS
#define GLB_PT_ITEMS 8
#endif
#ifndef GLB_PT_MINBLOCKS
#define GLB_PT_MINBLOCKS 4
#endif
constexpr int PT_ITEMS = GLB_PT_ITEMS;  // pixels per thread and tile (4 or 8): tile = 1024 or 2048 pixels
constexpr int PT_TILE = PT_THREADS * PT_ITEMS;

enum { BIAS_NONE = 0, BIAS_LINEAR = 1, BIAS_LOGLINEAR = 2 };

__device__ __forceinline__ double biased(double d, int model, double b) {
  if (model == BIAS_LINEAR) return __dmul_rn(b, d);
  if (model == BIAS_LOGLINEAR) return expm1(__dmul_rn(log1p(d), b));
  return d;
}

// ---- deterministic two-stage sum (for remove_monopole's mean) -------------------------
__global__ void __launch_bounds__(PT_THREADS) points_partial_sum_kernel(const double* __restrict__ delta, int64_t npix,
                                                                        int model, double b, double* __restrict__ partial) {
  __shared__ double sh[PT_THREADS / 32];
  const int64_t base = (int64_t)blockIdx.x * PT_TILE;
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < PT_ITEMS; ++i) {
    const int64_t p = base + threadIdx.x + (int64_t)i * PT_THREADS;
    if (p < npix) s += biased(delta[p], model, b);
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < PT_THREADS / 32; ++w) t += sh[w];
    partial[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(1024) final_sum_kernel(const double* __restrict__ partial, int n, double scale,
                                                         double* __restrict__ out) {
  __shared__ double sh[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += partial[i];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w];
    out[0] = t * scale;
  }
}

// ---- Poisson from Philox --------------------------------------------------------------
// PTRS, Hoermann 1993 (the algorithm NumPy uses for lam >= 10); kept out of line so that its
// registers (log, lgamma) do not limit the occupancy of the common small-lambda path
__device__ __noinline__ int64_t poisson_ptrs(double lam, uint32_t k0, uint32_t k1, uint64_t idx, uint32_t stream) {
  const double slam = sqrt(lam), loglam = log(lam);
  const double b = 0.931 + 2.53 * slam;
  const double a = -0.059 + 0.02483 * b;
  const double invalpha = 1.1239 + 1.1328 / (b - 3.4);
  const double vr = 0.9277 - 3.6224 / (b - 2.0);
  for (uint32_t attempt = 0; attempt < 1024; ++attempt) {
    const Philox4 r = philox4x32_10((uint32_t)idx, (uint32_t)(idx >> 32), stream, RNG_TAG_POISSON + 1u + attempt, k0, k1);
    const double U = u01_closed_open(r.v[0], r.v[1]) - 0.5;
    const double V = u01_open_closed(r.v[2], r.v[3]);
    const double us = 0.5 - fabs(U);
    const double kf = floor((2.0 * a / us + b) * U + lam + 0.43);
    if (us >= 0.07 && V <= vr) return (int64_t)kf;
    if (kf < 0.0 || (us < 0.013 && V > us)) continue;
    if ((log(V) + log(invalpha) - log(a / (us * us) + b)) <= (-lam + kf * loglam - lgamma(kf + 1.0))) return (int64_t)kf;
  }
  return (int64_t)floor(lam);
}

// slow path of the small-lambda inversion / PTRS: only pixels that did not resolve to zero
// galaxies by the cheap test get here, compacted to dense warps by the caller
__device__ __forceinline__ int64_t poisson_slow(double lam, double u, uint32_t k0, uint32_t k1, uint64_t pix,
                                                uint32_t stream) {
  if (lam < 10.0) {
    // inversion by sequential search
    double p = exp(-lam), F = p;
    int64_t x = 0;
    while (u > F && x < 256) {
      ++x;
      p *= lam / (double)x;
      F += p;
    }
    return x;
  }
  return poisson_ptrs(lam, k0, k1, pix, stream);
}

struct CountParams {
  const double* delta;
  const double* vis;        // may be null
  const double* mean;       // may be null (device scalar to subtract: remove_monopole)
  const int64_t* counts_in; // parity mode: supplied counts (may be null)
  double* nbar_out;         // may be null
  int64_t* counts;          // out [npix]; may be null (list mode)
  int64_t* off;             // out [npix+1]; may be null (list mode)
  int64_t* gpix;            // out: ring pixel of every galaxy, galaxies in pixel order (may be null)
  int64_t gpix_cap;         // entries of gpix that may be written (the total is exact either way)
  int64_t* total;           // out [1]: number of galaxies (may be null)
  unsigned long long* status;  // [ntiles] decoupled look-back words, zeroed before the launch
  unsigned int* ticket;        // tile ticket counter, zeroed before the launch
  int64_t npix;
  int ntiles;
  double scale;             // ARCMIN2_SPHERE / npix * ngal   (computed on the host like points.py:287)
  double bias;
  int model;
  int sample;               // 1: draw Poisson, 0: copy counts_in
  uint32_t k0, k1, stream;
};

constexpr int PT_WARPS = PT_THREADS / 32;     // worker warps; one more warp does the look-back
constexpr int PT_WSEG = PT_TILE / PT_WARPS;   // pixels per worker warp (contiguous)
constexpr int PT_QUADS = PT_WSEG / 128;       // 4-pixel groups per lane
constexpr int PT_QCAP = PT_WSEG;              // slow-pixel queue entries per warp (worst case: every pixel)
constexpr unsigned long long ST_AGG = 1ull << 62, ST_INCL = 2ull << 62, ST_VAL = (1ull << 62) - 1;

__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ int64_t warp_sum_i64(int64_t v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// named barriers: bar.sync blocks, bar.arrive only signals (producer side)
__device__ __forceinline__ void named_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void named_arrive(int id, int nthreads) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// K6 + K7 in ONE pass over the map: expected count -> Poisson count -> exclusive offsets.
//  * a CTA takes a tile of 2048 pixels by atomic ticket; a worker warp owns 256 contiguous
//    pixels, a lane two groups of 4 (coalesced 16-byte loads);
//  * one Philox call gives the four pixels of a group 32 random bits each.  A pixel whose bits
//    already prove u < 1 - lam <= e^{-lam} has zero galaxies (at 0.08 galaxies per pixel: 11 of
//    12).  The others are queued in shared memory, get the remaining 21 bits of their uniform
//    from a second Philox call and are drawn by DENSE warps, so the exp / division / PTRS code
//    runs once per 32 queued pixels instead of once per group of 32 mostly idle lanes;
//  * the tile's counts are scanned with shuffles and chained to the preceding tiles by a
//    decoupled look-back on one 64-bit status word per tile (flag | value).  A ninth warp does
//    the look-back WHILE the workers draw, and the counts are stored before the workers wait
//    for it, so the offsets cost no second pass: 8 B read + 16 B written per pixel.
//  (Measured alternatives, nside 4096, 0.083 galaxies/pixel: three kernels count / scan /
//  offsets 2.57 ms; look-back by warp 0 after the draw 2.04 ms; this version 1.73 ms; a
//  persistent variant that defers the offsets of a tile by one tile 1.99 ms; 128-wide
//  look-back windows 2.51 ms.)
template <bool VIS, bool SAMPLE, bool LOGLIN>
__global__ void __launch_bounds__(PT_THREADS + 32, GLB_PT_MINBLOCKS) points_count_scan_kernel(const CountParams p) {
  // SAMPLE only: per-pixel counts of the tile and the per-warp queue of pixels that need the slow draw
  __shared__ __align__(16) int s_cnt[SAMPLE ? PT_TILE : 4];
  __shared__ double s_qlam[SAMPLE ? PT_WARPS : 1][SAMPLE ? PT_QCAP : 1];
  __shared__ unsigned s_qbits[SAMPLE ? PT_WARPS : 1][SAMPLE ? PT_QCAP : 1];
  __shared__ unsigned short s_qidx[SAMPLE ? PT_WARPS : 1][SAMPLE ? PT_QCAP : 1];
  __shared__ int64_t s_wtot[PT_WARPS];
  __shared__ int64_t s_excl;
  __shared__ unsigned s_tile;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned full = 0xffffffffu;
  constexpr int NALL = PT_THREADS + 32;
  // named barriers: 1 warp totals ready (workers arrive, look-back warp syncs), 2 tile prefix
  // ready (look-back warp arrives, workers sync), 3 workers only
  if (tid == 0) s_tile = atomicAdd(p.ticket, 1u);
  __syncthreads();
  const unsigned tile = s_tile;

  if (warp == PT_WARPS) {
    // ---- look-back warp: exclusive prefix of this tile from its predecessors' status words ----
    int64_t excl = 0;
    if (tile > 0) {
      int64_t look = (int64_t)tile - 1;
      for (;;) {
        const int64_t t = look - lane;
        unsigned long long w = ST_INCL;  // before the first tile: inclusive prefix 0
        do {
          if (t >= 0) w = ld_relaxed_u64(p.status + t);
        } while (__any_sync(full, (w >> 62) == 0));
        const unsigned inc = __ballot_sync(full, (w >> 62) == 2);
        const int64_t v = (int64_t)(w & ST_VAL);
        if (inc) {
          const int first = __ffs(inc) - 1;  // nearest predecessor holding an inclusive prefix
          excl += warp_sum_i64(lane <= first ? v : 0);
          break;
        }
        excl += warp_sum_i64(v);
        look -= 32;
      }
    }
    if (lane == 0) s_excl = excl;
    named_sync(1, NALL);
    if (lane == 0) {
      int64_t agg = 0;
#pragma unroll
      for (int w = 0; w < PT_WARPS; ++w) agg += s_wtot[w];
      // max: never replace an inclusive word by the aggregate the workers publish
      atomicMax(p.status + tile, ST_INCL | (unsigned long long)(excl + agg));
      if ((int)tile == p.ntiles - 1) {
        if (p.off) p.off[p.npix] = excl + agg;
        if (p.total) p.total[0] = excl + agg;
      }
    }
    named_arrive(2, NALL);
    return;
  }

  const int64_t wbase = (int64_t)tile * PT_TILE + warp * PT_WSEG;
  const double mean = p.mean ? p.mean[0] : 0.0;

  // issue all loads of the tile first (memory-level parallelism), then do the arithmetic
  double2 dv[PT_QUADS][2], vv[PT_QUADS][2];
  longlong2 c[PT_QUADS][2];
#pragma unroll
  for (int i = 0; i < PT_QUADS; ++i) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int64_t pix = wbase + i * 128 + 4 * lane + 2 * h;
      const bool ok = pix < p.npix;
      dv[i][h] = ok ? *reinterpret_cast<const double2*>(p.delta + pix) : make_double2(0.0, 0.0);
      vv[i][h] = (ok && VIS) ? *reinterpret_cast<const double2*>(p.vis + pix) : make_double2(1.0, 1.0);
      c[i][h] = (ok && !SAMPLE) ? *reinterpret_cast<const longlong2*>(p.counts_in + pix) : make_longlong2(0, 0);
    }
  }

  int qn = 0;  // warp-uniform queue length
#pragma unroll
  for (int i = 0; i < PT_QUADS; ++i) {
    const int loc = i * 128 + 4 * lane;
    const int64_t pix = wbase + loc;
    double n[4] = {dv[i][0].x, dv[i][0].y, dv[i][1].x, dv[i][1].y};
    const double vs[4] = {vv[i][0].x, vv[i][0].y, vv[i][1].x, vv[i][1].y};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      // log-linear bias (expm1, log1p) lives in its own instantiation: the linear / no-bias
      // streaming path stays free of out-of-line math calls
      n[e] = LOGLIN ? biased(n[e], BIAS_LOGLINEAR, p.bias) : (p.model == BIAS_LINEAR ? __dmul_rn(p.bias, n[e]) : n[e]);
      if (p.mean) n[e] = __dsub_rn(n[e], mean);
      n[e] = __dmul_rn(__dadd_rn(n[e], 1.0), p.scale);
      if (VIS) n[e] = __dmul_rn(n[e], vs[e]);
    }
    if (p.nbar_out) {
      if (pix < p.npix) *reinterpret_cast<double2*>(p.nbar_out + pix) = make_double2(n[0], n[1]);
      if (pix + 2 < p.npix) *reinterpret_cast<double2*>(p.nbar_out + pix + 2) = make_double2(n[2], n[3]);
    }
    if (SAMPLE) {
      *reinterpret_cast<int4*>(s_cnt + warp * PT_WSEG + loc) = make_int4(0, 0, 0, 0);
      // one Philox call per 4 pixels: word e -> the leading 32 bits of pixel e's uniform
      const uint64_t qd = (uint64_t)pix >> 2;
      const Philox4 r = philox4x32_10((uint32_t)qd, (uint32_t)(qd >> 32), p.stream, RNG_TAG_POISSON, p.k0, p.k1);
      const unsigned lt = (1u << lane) - 1u;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        // u < (bits+1) 2^-32; zero galaxies for sure if that bound is <= 1 - lam <= e^{-lam}
        const double ub = __hiloint2double(0x43300000, (int)r.v[e]) - 4503599627370495.0;  // bits + 1, exact
        const bool slow =
            (pix + e < p.npix) && (n[e] > 0.0) && !(n[e] < 10.0 && ub <= fma(-n[e], 4294967296.0, 4294967296.0));
        const unsigned m = __ballot_sync(full, slow);
        if (slow) {
          const int k = qn + __popc(m & lt);
          s_qlam[warp][k] = n[e];
          s_qbits[warp][k] = r.v[e];
          s_qidx[warp][k] = (unsigned short)(loc + e);
        }
        qn += __popc(m);
      }
    }
  }
  if (SAMPLE) {
    __syncwarp();
    // the queued pixels, 32 at a time on dense lanes
    for (int k = lane; k < qn; k += 32) {
      const int loc = s_qidx[warp][k];
      const uint64_t pix = (uint64_t)(wbase + loc);
      // the other 21 bits of the 53-bit uniform
      const Philox4 r = philox4x32_10((uint32_t)pix, (uint32_t)(pix >> 32), p.stream, RNG_TAG_POISSON + 0x80u, p.k0, p.k1);
      const double u = ((double)s_qbits[warp][k] * 2097152.0 + (double)(r.v[0] >> 11)) * (1.0 / 9007199254740992.0);
      const int64_t v = poisson_slow(s_qlam[warp][k], u, p.k0, p.k1, pix, p.stream);
      s_cnt[warp * PT_WSEG + loc] = (int)min(v, (int64_t)2147483647);
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < PT_QUADS; ++i) {
      const int4 v = *reinterpret_cast<const int4*>(s_cnt + warp * PT_WSEG + i * 128 + 4 * lane);
      c[i][0] = make_longlong2(v.x, v.y);
      c[i][1] = make_longlong2(v.z, v.w);
    }
  }

  // ---- scan: lane-level, then warp totals; the tile prefix comes from the look-back warp ----
  int64_t ex[PT_QUADS];
  int64_t run = 0;
#pragma unroll
  for (int i = 0; i < PT_QUADS; ++i) {
    const int64_t s = (c[i][0].x + c[i][0].y) + (c[i][1].x + c[i][1].y);
    int64_t incl = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int64_t t = __shfl_up_sync(full, incl, o);
      if (lane >= o) incl += t;
    }
    ex[i] = run + incl - s;
    run += __shfl_sync(full, incl, 31);
  }
  if (lane == 0) s_wtot[warp] = run;
  named_arrive(1, NALL);
  named_sync(3, PT_THREADS);
  int64_t wpre = 0, agg = 0;
#pragma unroll
  for (int w = 0; w < PT_WARPS; ++w) {
    const int64_t t = s_wtot[w];
    agg += t;
    if (w < warp) wpre += t;
  }
  if (tid == 0 && tile > 0) atomicMax(p.status + tile, ST_AGG | (unsigned long long)agg);
  // the counts do not depend on the prefix: store them while the look-back finishes
  if (p.counts) {
#pragma unroll
    for (int i = 0; i < PT_QUADS; ++i) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int64_t pix = wbase + i * 128 + 4 * lane + 2 * h;
        if (pix < p.npix) *reinterpret_cast<longlong2*>(p.counts + pix) = c[i][h];
      }
    }
  }
  named_sync(2, NALL);
  wpre += s_excl;
  if (p.off) {
#pragma unroll
    for (int i = 0; i < PT_QUADS; ++i) {
      const int64_t pix = wbase + i * 128 + 4 * lane;
      const int64_t o0 = wpre + ex[i];
      const int64_t o1 = o0 + c[i][0].x, o2 = o1 + c[i][0].y, o3 = o2 + c[i][1].x;
      if (pix < p.npix) *reinterpret_cast<longlong2*>(p.off + pix) = make_longlong2(o0, o1);
      if (pix + 2 < p.npix) *reinterpret_cast<longlong2*>(p.off + pix + 2) = make_longlong2(o2, o3);
    }
  }
  if (p.gpix) {
    // LIST mode: the galaxy -> pixel map itself (ipix = repeat(arange, n), points.py:426), written at
    // the galaxies' global positions: a sparse map (0.08 galaxies per pixel at the north-star density)
    // costs 8 B per GALAXY here instead of 16 B per PIXEL of counts + offsets, the cut rule and the
    // position kernel then never touch a per-pixel array again.  A tile's galaxies are consecutive,
    // so its stores fall into a few adjacent lines.
#pragma unroll
    for (int i = 0; i < PT_QUADS; ++i) {
      const int64_t pix = wbase + i * 128 + 4 * lane;
      int64_t o = wpre + ex[i];
      const int64_t ce[4] = {c[i][0].x, c[i][0].y, c[i][1].x, c[i][1].y};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        for (int64_t k = 0; k < ce[e]; ++k)
          if (o + k < p.gpix_cap) p.gpix[o + k] = pix + e;
        o += ce[e];
      }
    }
  }
}

struct FillParams {
  const int64_t* counts;
  const int64_t* off;
  const double* u;   // supplied in-pixel offsets for THIS pixel range (may be null)
  const double* v;
  double* lon;       // outputs for this pixel range, index 0 = first galaxy of pixel p0
  double* lat;
  int64_t* ipix;     // may be null
  int64_t p0, p1;
  int64_t nside;
  uint32_t k0, k1, stream;
};

// One CTA per tile of FILL_TILE pixels: the tile's counts are expanded in shared memory
// (block scan -> local offsets), then threads walk GALAXIES, not pixels, so the expensive
// pixel->angle arithmetic runs on dense warps and the (lon, lat) stores are coalesced.
constexpr int FILL_TILE = 3072;  // ~255 galaxies per tile at 0.083 galaxies/pixel: dense lanes
constexpr int FILL_THREADS = 256;
constexpr int FILL_ITEMS = FILL_TILE / FILL_THREADS;

__global__ void __launch_bounds__(FILL_THREADS) points_fill_kernel(const FillParams p) {
  __shared__ int s_off[FILL_TILE + 1];
  const int tid = threadIdx.x;
  const int64_t tile0 = p.p0 + (int64_t)blockIdx.x * FILL_TILE;
  const int npx = (int)min((int64_t)FILL_TILE, p.p1 - tile0);
  const int64_t g0 = p.off[tile0];          // global index of the tile's first galaxy
  // local offsets straight from the global exclusive scan (coalesced load, no block scan)
#pragma unroll
  for (int i = 0; i <= FILL_ITEMS; ++i) {
    const int k = tid + i * FILL_THREADS;
    if (k <= FILL_TILE) s_off[k] = (int)(p.off[tile0 + min(k, npx)] - g0);
  }
  __syncthreads();
  const int total = s_off[FILL_TILE];
  if (total == 0) return;
  const int64_t o0 = g0 - p.off[p.p0];      // its position in this call's output
  const double rad2deg = 57.295779513082320877;  // 180/pi, as np.degrees
  for (int g = tid; g < total; g += FILL_THREADS) {
    // largest i with s_off[i] <= g
    int lo = 0, hi = FILL_TILE;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (s_off[mid] <= g)
        lo = mid;
      else
        hi = mid;
    }
    const int64_t pix = tile0 + lo;
    double u, v;
    if (p.u) {
      u = p.u[o0 + g];
      v = p.v[o0 + g];
    } else {
      const uint64_t gi = (uint64_t)(g0 + g);
      const Philox4 r = philox4x32_10((uint32_t)gi, (uint32_t)(gi >> 32), p.stream, RNG_TAG_POS, p.k0, p.k1);
      u = u01_closed_open(r.v[0], r.v[1]);
      v = u01_closed_open(r.v[2], r.v[3]);
    }
    int x, y, f;
    ring2xyf(p.nside, pix, x, y, f);
    double z, sth, phi;
    hpc2loc((double)p.nside, x, y, f, u, v, z, sth, phi);
    p.lon[o0 + g] = phi * rad2deg;
    p.lat[o0 + g] = 90.0 - atan2(sth, z) * rad2deg;
    if (p.ipix) p.ipix[o0 + g] = pix;
  }
}

__global__ void __launch_bounds__(256) ring2ang_uv_kernel(int64_t nside, const int64_t* __restrict__ ipix,
                                                          const double* __restrict__ u, const double* __restrict__ v,
                                                          int64_t n, int lonlat, double* __restrict__ o1,
                                                          double* __restrict__ o2) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int x, y, f;
  ring2xyf(nside, ipix[i], x, y, f);
  double z, sth, phi;
  hpc2loc((double)nside, x, y, f, u[i], v[i], z, sth, phi);
  const double theta = atan2(sth, z);
  if (lonlat) {
    const double rad2deg = 57.295779513082320877;
    o1[i] = phi * rad2deg;
    o2[i] = 90.0 - theta * rad2deg;
  } else {
    o1[i] = theta;
    o2[i] = phi;
  }
}

// healpix.randang: (u, v) from Philox keyed by (seed, stream, element index)
__global__ void __launch_bounds__(256) randang_kernel(int64_t nside, const int64_t* __restrict__ ipix, int64_t n,
                                                      uint32_t k0, uint32_t k1, uint32_t stream, int lonlat,
                                                      double* __restrict__ o1, double* __restrict__ o2) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Philox4 r = philox4x32_10((uint32_t)i, (uint32_t)((uint64_t)i >> 32), stream, RNG_TAG_POS, k0, k1);
  const double u = u01_closed_open(r.v[0], r.v[1]);
  const double v = u01_closed_open(r.v[2], r.v[3]);
  int x, y, f;
  ring2xyf(nside, ipix[i], x, y, f);
  double z, sth, phi;
  hpc2loc((double)nside, x, y, f, u, v, z, sth, phi);
  const double theta = atan2(sth, z);
  if (lonlat) {
    const double rad2deg = 57.295779513082320877;
    o1[i] = phi * rad2deg;
    o2[i] = 90.0 - theta * rad2deg;
  } else {
    o1[i] = theta;
    o2[i] = phi;
  }
}

// glass.uniform_positions (glass/points.py:543-607): lon = uniform(-180, 180), lat = degrees(asin(uniform(-1, 1)))
// per point; (u1, u2) from Philox keyed by (seed, stream, point index) or supplied arrays.  NumPy's
// uniform(low, high) is low + (high - low) * random(), reproduced with separately rounded operations.
__global__ void __launch_bounds__(256) uniform_positions_kernel(int64_t n, const double* __restrict__ u1,
                                                                const double* __restrict__ u2, uint32_t k0, uint32_t k1,
                                                                uint32_t stream, double* __restrict__ lon,
                                                                double* __restrict__ lat) {
  const double rad2deg = 57.295779513082320877;
  // two points per thread: 16-byte stores
  const int64_t i = 2 * ((int64_t)blockIdx.x * blockDim.x + threadIdx.x);
  if (i >= n) return;
  double a[2], b[2];
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const int64_t j = min(i + e, n - 1);
    if (u1) {
      a[e] = u1[j];
      b[e] = u2[j];
    } else {
      const Philox4 r = philox4x32_10((uint32_t)j, (uint32_t)((uint64_t)j >> 32), stream, RNG_TAG_POS + 0x40u, k0, k1);
      a[e] = u01_closed_open(r.v[0], r.v[1]);
      b[e] = u01_closed_open(r.v[2], r.v[3]);
    }
    a[e] = __dadd_rn(-180.0, __dmul_rn(360.0, a[e]));
    b[e] = __dmul_rn(asin(__dadd_rn(-1.0, __dmul_rn(2.0, b[e]))), rad2deg);
  }
  if (i + 1 < n) {
    *reinterpret_cast<double2*>(lon + i) = make_double2(a[0], a[1]);
    *reinterpret_cast<double2*>(lat + i) = make_double2(b[0], b[1]);
  } else {
    lon[i] = a[0];
    lat[i] = b[0];
  }
}

__global__ void __launch_bounds__(256) ang2pix_kernel(int64_t nside, const double* __restrict__ a, const double* __restrict__ b,
                                                      int64_t n, int lonlat, int64_t* __restrict__ ipix) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double theta, phi;
  if (lonlat) {
    const double deg2rad = 0.017453292519943295769;
    theta = (90.0 - b[i]) * deg2rad;
    phi = a[i] * deg2rad;
  } else {
    theta = a[i];
    phi = b[i];
  }
  double s, c;
  sincos(theta, &s, &c);
  ipix[i] = zphi2pix_ring(nside, c, s, phi);
}

// K8, list mode: one thread per galaxy, its pixel read from the list K6 wrote -- no per-pixel
// array, no search, no shared memory; consecutive lanes store consecutive galaxies.
__global__ void __launch_bounds__(256) points_fill_list_kernel(const int64_t* __restrict__ gpix, int64_t g0, int64_t n,
                                                               const double* __restrict__ us, const double* __restrict__ vs,
                                                               int64_t nside, uint32_t k0, uint32_t k1, uint32_t stream,
                                                               double* __restrict__ lon, double* __restrict__ lat) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n) return;
  const uint64_t gi = (uint64_t)(g0 + g);
  const int64_t pix = gpix[gi];
  double u, v;
  if (us) {
    u = us[g];
    v = vs[g];
  } else {
    const Philox4 r = philox4x32_10((uint32_t)gi, (uint32_t)(gi >> 32), stream, RNG_TAG_POS, k0, k1);
    u = u01_closed_open(r.v[0], r.v[1]);
    v = u01_closed_open(r.v[2], r.v[3]);
  }
  int x, y, f;
  ring2xyf(nside, pix, x, y, f);
  double z, sth, phi;
  hpc2loc((double)nside, x, y, f, u, v, z, sth, phi);
  const double rad2deg = 57.295779513082320877;  // 180/pi, as np.degrees
  lon[g] = phi * rad2deg;
  lat[g] = 90.0 - atan2(sth, z) * rad2deg;
}

__global__ void points_cuts_list_kernel(const int64_t* __restrict__ gpix, int64_t total, int64_t npix, int64_t batch,
                                        int64_t start, int64_t remaining, int max_cuts, int64_t* __restrict__ cuts,
                                        int64_t* __restrict__ state) {
  if (blockIdx.x == 0 && threadIdx.x == 0) cuts_list_chain(gpix, total, npix, batch, start, remaining, max_cuts, cuts, state);
}

// The chain of batch cuts is sequential (every cut starts where the previous one stopped) but tiny:
// two binary searches over the offsets per cut.  One thread walks it, so the host fetches the cuts
// of a population with one copy instead of synchronising several times per batch.
__global__ void points_cuts_kernel(const int64_t* __restrict__ off, int64_t npix, int64_t batch, int64_t start,
                                   int64_t remaining, int max_cuts, int64_t* __restrict__ cuts,
                                   int64_t* __restrict__ state) {
  if (blockIdx.x == 0 && threadIdx.x == 0) cuts_chain(off, npix, batch, start, remaining, max_cuts, cuts, state);
}

}  // namespace glb

using namespace glb;

extern "C" {

size_t glb_points_workspace_bytes(int64_t npix) {
  const int64_t nblocks = (npix + PT_TILE - 1) / PT_TILE;
  return (size_t)(2 * nblocks + 2) * sizeof(int64_t) + (size_t)(nblocks + 1) * sizeof(double);
}

int glb_points_counts(int64_t npix, const double* d_delta, const double* d_vis, int bias_model, double bias,
                      double scale, int remove_monopole, const int64_t* d_counts_in, uint64_t seed, uint32_t stream_id,
                      double* d_nbar_out, int64_t* d_counts, int64_t* d_off, int64_t* d_gpix, int64_t gpix_capacity,
                      int64_t* d_total, void* d_workspace, void* stream) {
  GLB_REQUIRE(npix > 0 && d_delta && d_workspace, "null pointer or empty map");
  GLB_REQUIRE(d_off || d_total, "nowhere to write the total (d_off and d_total both NULL)");
  GLB_REQUIRE(d_gpix == nullptr || gpix_capacity >= 0, "negative list capacity");
  GLB_REQUIRE(bias_model >= BIAS_NONE && bias_model <= BIAS_LOGLINEAR, "unknown bias model");
  cudaStream_t st = (cudaStream_t)stream;
  const int nblocks = (int)((npix + PT_TILE - 1) / PT_TILE);
  int64_t* block_sums = reinterpret_cast<int64_t*>(d_workspace);
  int64_t* block_off = block_sums + nblocks;
  double* partial = reinterpret_cast<double*>(block_off + nblocks + 2);
  double* mean = partial + nblocks;
  if (remove_monopole) {
    points_partial_sum_kernel<<<nblocks, PT_THREADS, 0, st>>>(d_delta, npix, bias_model, bias, partial);
    final_sum_kernel<<<1, 1024, 0, st>>>(partial, nblocks, 1.0 / (double)npix, mean);
    count_launch(2);
  }
  GLB_REQUIRE(npix % 2 == 0, "npix must be even (12 nside^2)");
  GLB_REQUIRE(((uintptr_t)d_delta | (uintptr_t)d_counts | (uintptr_t)d_off | (uintptr_t)d_vis | (uintptr_t)d_counts_in |
               (uintptr_t)d_nbar_out) % 16 == 0,
              "map pointers must be 16-byte aligned");
  CountParams p;
  p.delta = d_delta;
  p.vis = d_vis;
  p.mean = remove_monopole ? mean : nullptr;
  p.counts_in = d_counts_in;
  p.nbar_out = d_nbar_out;
  p.counts = d_counts;
  p.off = d_off;
  p.gpix = d_gpix;
  p.gpix_cap = gpix_capacity;
  p.total = d_total;
  p.status = reinterpret_cast<unsigned long long*>(block_sums);
  p.ticket = reinterpret_cast<unsigned int*>(block_sums + nblocks);
  p.npix = npix;
  p.ntiles = nblocks;
  p.scale = scale;
  p.bias = bias;
  p.model = bias_model;
  p.sample = d_counts_in ? 0 : 1;
  p.k0 = (uint32_t)seed;
  p.k1 = (uint32_t)(seed >> 32);
  p.stream = stream_id;
  GLB_CUDA_CHECK(cudaMemsetAsync(block_sums, 0, (size_t)(nblocks + 1) * sizeof(int64_t), st));
  const int grid = nblocks;
  const int variant = (d_vis ? 4 : 0) | (p.sample ? 2 : 0) | (bias_model == BIAS_LOGLINEAR ? 1 : 0);
#define GLB_COUNT_LAUNCH(V, VIS, SAMPLE, LOGLIN)                                                           \
  case V:                                                                                                  \
    points_count_scan_kernel<VIS, SAMPLE, LOGLIN><<<grid, PT_THREADS + 32, 0, st>>>(p);                     \
    break;
  switch (variant) {
    GLB_COUNT_LAUNCH(0, false, false, false)
    GLB_COUNT_LAUNCH(1, false, false, true)
    GLB_COUNT_LAUNCH(2, false, true, false)
    GLB_COUNT_LAUNCH(3, false, true, true)
    GLB_COUNT_LAUNCH(4, true, false, false)
    GLB_COUNT_LAUNCH(5, true, false, true)
    GLB_COUNT_LAUNCH(6, true, true, false)
    GLB_COUNT_LAUNCH(7, true, true, true)
  }
#undef GLB_COUNT_LAUNCH
  GLB_CUDA_CHECK(cudaGetLastError());
  count_launch(1);
  return GLB_OK;
}

int glb_points_cuts(const int64_t* d_off, int64_t npix, int64_t batch, int64_t start, int64_t remaining, int max_cuts,
                    int64_t* d_cuts, int64_t* d_state, void* stream) {
  GLB_REQUIRE(d_off && d_cuts && d_state, "null pointer");
  GLB_REQUIRE(npix >= 1 && batch >= 1 && start >= 0 && start <= npix && remaining >= 0 && max_cuts >= 1, "bad size");
  points_cuts_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(d_off, npix, batch, start, remaining, max_cuts, d_cuts, d_state);
  GLB_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return GLB_OK;
}

int glb_points_cuts_list(const int64_t* d_gpix, int64_t total, int64_t npix, int64_t batch, int64_t start,
                         int64_t remaining, int max_cuts, int64_t* d_cuts, int64_t* d_state, void* stream) {
  GLB_REQUIRE(d_gpix && d_cuts && d_state, "null pointer");
  GLB_REQUIRE(npix >= 1 && batch >= 1 && start >= 0 && start <= npix && remaining >= 0 && remaining <= total && max_cuts >= 1,
              "bad size");
  points_cuts_list_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(d_gpix, total, npix, batch, start, remaining, max_cuts, d_cuts,
                                                             d_state);
  GLB_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return GLB_OK;
}

int glb_points_fill_list(int64_t nside, const int64_t* d_gpix, int64_t g0, int64_t g1, const double* d_u,
                         const double* d_v, uint64_t seed, uint32_t stream_id, double* d_lon, double* d_lat,
                         void* stream) {
  GLB_REQUIRE(nside >= 1 && d_gpix && d_lon && d_lat, "null pointer");
  GLB_REQUIRE(g0 >= 0 && g1 >= g0, "bad galaxy range");
  GLB_REQUIRE((d_u == nullptr) == (d_v == nullptr), "u and v must be given together");
  const int64_t n = g1 - g0;
  if (n == 0) return GLB_OK;
  points_fill_list_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      d_gpix, g0, n, d_u, d_v, nside, (uint32_t)seed, (uint32_t)(seed >> 32), stream_id, d_lon, d_lat);
  GLB_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return GLB_OK;
}

int glb_points_fill(int64_t nside, const int64_t* d_counts, const int64_t* d_off, int64_t pix0, int64_t pix1,
                    const double* d_u, const double* d_v, uint64_t seed, uint32_t stream_id, double* d_lon,
                    double* d_lat, int64_t* d_ipix, void* stream) {
  GLB_REQUIRE(nside >= 1 && d_counts && d_off && d_lon && d_lat, "null pointer");
  GLB_REQUIRE(pix0 >= 0 && pix1 >= pix0 && pix1 <= 12 * nside * nside, "bad pixel range");
  GLB_REQUIRE((d_u == nullptr) == (d_v == nullptr), "u and v must be given together");
  if (pix1 == pix0) return GLB_OK;
  FillParams p;
  p.counts = d_counts;
  p.off = d_off;
  p.u = d_u;
  p.v = d_v;
  p.lon = d_lon;
  p.lat = d_lat;
  p.ipix = d_ipix;
  p.p0 = pix0;
  p.p1 = pix1;
  p.nside = nside;
  p.k0 = (uint32_t)seed;
  p.k1 = (uint32_t)(seed >> 32);
  p.stream = stream_id;
  const int64_t n = pix1 - pix0;
  points_fill_kernel<<<(unsigned)((n + FILL_TILE - 1) / FILL_TILE), FILL_THREADS, 0, (cudaStream_t)stream>>>(p);
  GLB_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return GLB_OK;
}

int glb_ring2ang_uv(int64_t nside, const int64_t* d_ipix, const double* d_u, const double* d_v, int64_t n, int lonlat,
                    double* d_out1, double* d_out2, void* stream) {
  GLB_REQUIRE(nside >= 1 && n >= 0, "bad size");
  if (n == 0) return GLB_OK;
  GLB_REQUIRE(d_ipix && d_u && d_v && d_out1 && d_out2, "null pointer");
  ring2ang_uv_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(nside, d_ipix, d_u, d_v, n, lonlat,
                                                                                   d_out1, d_out2);
  GLB_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return GLB_OK;
}

int glb_randang(int64_t nside, const int64_t* d_ipix, int64_t n, uint64_t seed, uint32_t stream_id, int lonlat,
                double* d_out1, double* d_out2, void* stream) {
  GLB_REQUIRE(nside >= 1 && n >= 0, "bad size");
  if (n == 0) return GLB_OK;
  GLB_REQUIRE(d_ipix && d_out1 && d_out2, "null pointer");
  randang_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      nside, d_ipix, n, (uint32_t)seed, (uint32_t)(seed >> 32), stream_id, lonlat, d_out1, d_out2);
  GLB_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return GLB_OK;
}

int glb_uniform_positions(int64_t n, const double* d_u1, const double* d_u2, uint64_t seed, uint32_t stream_id,
                          double* d_lon, double* d_lat, void* stream) {
  GLB_REQUIRE(n >= 0, "bad size");
  if (n == 0) return GLB_OK;
  GLB_REQUIRE(d_lon && d_lat, "null pointer");
  GLB_REQUIRE((d_u1 == nullptr) == (d_u2 == nullptr), "the two uniform arrays must be given together");
  GLB_REQUIRE(((uintptr_t)d_lon | (uintptr_t)d_lat) % 16 == 0, "output pointers must be 16-byte aligned");
  const int64_t nthreads = (n + 1) / 2;
  uniform_positions_kernel<<<(unsigned)((nthreads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      n, d_u1, d_u2, (uint32_t)seed, (uint32_t)(seed >> 32), stream_id, d_lon, d_lat);
  GLB_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return GLB_OK;
}

int glb_ang2pix(int64_t nside, const double* d_a, const double* d_b, int64_t n, int lonlat, int64_t* d_ipix,
                void* stream) {
  GLB_REQUIRE(nside >= 1 && n >= 0, "bad size");
  if (n == 0) return GLB_OK;
  GLB_REQUIRE(d_a && d_b && d_ipix, "null pointer");
  ang2pix_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(nside, d_a, d_b, n, lonlat, d_ipix);
  GLB_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return GLB_OK;
}

}  // extern "C"
