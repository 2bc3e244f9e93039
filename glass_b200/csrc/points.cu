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
// K6 fuses bias, normalisation, visibility, clip and the Poisson draw into one pass that
// also emits per-block count sums; K7 turns them into exclusive pixel offsets (the batch
// cuts of points.py:409-424 are then a searchsorted on that array); K8 walks pixels and
// writes every galaxy's (lon, lat) at its offset, galaxies ordered by ring pixel index as
// in the reference.  All arithmetic that the reference does in NumPy is rounded the same
// way (separate multiplies/adds, no FMA contraction) so that the expected-count map is
// bit-identical for the linear / no-bias models.
#include "common.cuh"
#include "healpix_geom.cuh"
#include "rng.cuh"

namespace glb {

constexpr int PT_THREADS = 256;
constexpr int PT_ITEMS = 8;
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

// u: the pixel's own uniform (one Philox call serves the two pixels of a pair)
__device__ __forceinline__ int64_t poisson_from_uniform(double lam, double u, uint32_t k0, uint32_t k1, uint64_t pix,
                                                        uint32_t stream) {
  if (!(lam > 0.0)) return 0;
  if (lam < 10.0) {
    // inversion by sequential search
    if (u < 1.0 - lam) return 0;  // u < 1 - lam <= e^{-lam}: zero galaxies without evaluating exp
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
  int64_t* counts;          // out
  int64_t* block_sums;      // out [nblocks]
  int64_t npix;
  double scale;             // ARCMIN2_SPHERE / npix * ngal   (computed on the host like points.py:287)
  double bias;
  int model;
  int sample;               // 1: draw Poisson, 0: copy counts_in
  uint32_t k0, k1, stream;
};

__global__ void __launch_bounds__(PT_THREADS, 4) points_count_kernel(const CountParams p) {
  __shared__ int64_t sh[PT_THREADS / 32];
  const int64_t base = (int64_t)blockIdx.x * PT_TILE;
  const double mean = p.mean ? p.mean[0] : 0.0;
  int64_t s = 0;
  constexpr int NP = PT_ITEMS / 2;  // pixel pairs per thread (npix = 12 nside^2 is even)
  // issue all loads of the tile first (memory-level parallelism), then do the arithmetic
  double2 dv[NP], vv[NP];
  longlong2 cin[NP];
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    const int64_t pix = base + 2 * (threadIdx.x + (int64_t)i * PT_THREADS);
    const bool ok = pix < p.npix;
    dv[i] = ok ? *reinterpret_cast<const double2*>(p.delta + pix) : make_double2(0.0, 0.0);
    vv[i] = (ok && p.vis) ? *reinterpret_cast<const double2*>(p.vis + pix) : make_double2(1.0, 1.0);
    cin[i] = (ok && !p.sample) ? *reinterpret_cast<const longlong2*>(p.counts_in + pix) : make_longlong2(0, 0);
  }
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    const int64_t pix = base + 2 * (threadIdx.x + (int64_t)i * PT_THREADS);
    if (pix >= p.npix) continue;
    double n0 = biased(dv[i].x, p.model, p.bias), n1 = biased(dv[i].y, p.model, p.bias);
    if (p.mean) {
      n0 = __dsub_rn(n0, mean);
      n1 = __dsub_rn(n1, mean);
    }
    n0 = __dmul_rn(__dadd_rn(n0, 1.0), p.scale);
    n1 = __dmul_rn(__dadd_rn(n1, 1.0), p.scale);
    if (p.vis) {
      n0 = __dmul_rn(n0, vv[i].x);
      n1 = __dmul_rn(n1, vv[i].y);
    }
    if (p.nbar_out) *reinterpret_cast<double2*>(p.nbar_out + pix) = make_double2(n0, n1);
    longlong2 c;
    if (p.sample) {
      // one Philox call per pixel pair: words (0,1) -> even pixel, (2,3) -> odd pixel
      const uint64_t pr = (uint64_t)pix >> 1;
      const Philox4 r = philox4x32_10((uint32_t)pr, (uint32_t)(pr >> 32), p.stream, RNG_TAG_POISSON, p.k0, p.k1);
      c.x = poisson_from_uniform(fmax(n0, 0.0), u01_closed_open(r.v[0], r.v[1]), p.k0, p.k1, (uint64_t)pix, p.stream);
      c.y = poisson_from_uniform(fmax(n1, 0.0), u01_closed_open(r.v[2], r.v[3]), p.k0, p.k1, (uint64_t)pix + 1, p.stream);
    } else {
      c = cin[i];
    }
    *reinterpret_cast<longlong2*>(p.counts + pix) = c;
    s += c.x + c.y;
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    int64_t t = 0;
    for (int w = 0; w < PT_THREADS / 32; ++w) t += sh[w];
    p.block_sums[blockIdx.x] = t;
  }
}

// exclusive scan of the block sums (single CTA), total written to block_off[n]
__global__ void __launch_bounds__(1024) scan_block_sums_kernel(const int64_t* __restrict__ sums, int n,
                                                               int64_t* __restrict__ block_off) {
  __shared__ int64_t sh[1024];
  __shared__ int64_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 1024) {
    const int i = base + threadIdx.x;
    const int64_t v = (i < n) ? sums[i] : 0;
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      const int64_t t = (threadIdx.x >= o) ? sh[threadIdx.x - o] : 0;
      __syncthreads();
      sh[threadIdx.x] += t;
      __syncthreads();
    }
    if (i < n) block_off[i] = carry + sh[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry += sh[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) block_off[n] = carry;
}

// per-pixel exclusive offsets: off[p] = sum_{q<p} counts[q]; off[npix] = total
__global__ void __launch_bounds__(PT_THREADS) points_offsets_kernel(const int64_t* __restrict__ counts, int64_t npix,
                                                                    const int64_t* __restrict__ block_off, int nblocks,
                                                                    int64_t* __restrict__ off) {
  __shared__ int64_t s_c[PT_TILE];      // counts, then offsets, of the tile
  __shared__ int64_t sh[PT_THREADS / 32];
  const int64_t base = (int64_t)blockIdx.x * PT_TILE;
  const int tid = threadIdx.x;
  // coalesced load into shared memory
#pragma unroll
  for (int i = 0; i < PT_ITEMS; ++i) {
    const int k = tid + i * PT_THREADS;
    s_c[k] = (base + k < npix) ? counts[base + k] : 0;
  }
  __syncthreads();
  // thread t owns PT_ITEMS consecutive pixels
  int64_t c[PT_ITEMS];
  int64_t s = 0;
#pragma unroll
  for (int i = 0; i < PT_ITEMS; ++i) {
    c[i] = s_c[tid * PT_ITEMS + i];
    s += c[i];
  }
  // block scan of the per-thread sums: warp shuffles, then the warp totals
  int64_t incl = s;
  const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int64_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) sh[warp] = incl;
  __syncthreads();
  int64_t woff = 0;
#pragma unroll
  for (int w = 0; w < PT_THREADS / 32; ++w)
    if (w < warp) woff += sh[w];
  int64_t run = block_off[blockIdx.x] + woff + incl - s;
#pragma unroll
  for (int i = 0; i < PT_ITEMS; ++i) {
    s_c[tid * PT_ITEMS + i] = run;
    run += c[i];
  }
  __syncthreads();
  // coalesced store
#pragma unroll
  for (int i = 0; i < PT_ITEMS; ++i) {
    const int k = tid + i * PT_THREADS;
    if (base + k < npix) off[base + k] = s_c[k];
  }
  if (blockIdx.x == nblocks - 1 && tid == 0) off[npix] = block_off[nblocks];
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
constexpr int FILL_TILE = 2048;
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

}  // namespace glb

using namespace glb;

extern "C" {

size_t glb_points_workspace_bytes(int64_t npix) {
  const int64_t nblocks = (npix + PT_TILE - 1) / PT_TILE;
  return (size_t)(2 * nblocks + 2) * sizeof(int64_t) + (size_t)(nblocks + 1) * sizeof(double);
}

int glb_points_counts(int64_t npix, const double* d_delta, const double* d_vis, int bias_model, double bias,
                      double scale, int remove_monopole, const int64_t* d_counts_in, uint64_t seed, uint32_t stream_id,
                      double* d_nbar_out, int64_t* d_counts, int64_t* d_off, void* d_workspace, void* stream) {
  GLB_REQUIRE(npix > 0 && d_delta && d_counts && d_off && d_workspace, "null pointer or empty map");
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
  CountParams p;
  p.delta = d_delta;
  p.vis = d_vis;
  p.mean = remove_monopole ? mean : nullptr;
  p.counts_in = d_counts_in;
  p.nbar_out = d_nbar_out;
  p.counts = d_counts;
  p.block_sums = block_sums;
  p.npix = npix;
  p.scale = scale;
  p.bias = bias;
  p.model = bias_model;
  p.sample = d_counts_in ? 0 : 1;
  p.k0 = (uint32_t)seed;
  p.k1 = (uint32_t)(seed >> 32);
  p.stream = stream_id;
  points_count_kernel<<<nblocks, PT_THREADS, 0, st>>>(p);
  scan_block_sums_kernel<<<1, 1024, 0, st>>>(block_sums, nblocks, block_off);
  points_offsets_kernel<<<nblocks, PT_THREADS, 0, st>>>(d_counts, npix, block_off, nblocks, d_off);
  GLB_CUDA_CHECK(cudaGetLastError());
  count_launch(3);
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
