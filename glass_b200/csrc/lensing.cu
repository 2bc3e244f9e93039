// lensing.cu -- pixel- and galaxy-space lensing kernels (K9, K12) and the per-galaxy
// samplers that feed them (ellipticities, redshifts).
//
//   K9  MultiPlaneConvergence.add_plane update       glass/lensing.py:580-586
//         kappa3 <- (1-t) kappa1 + t kappa2 + f delta_prev     one fused pass, rounded like
//         NumPy's three in-place passes (separate multiplies and adds)
//   K12 galaxy_shear                                   glass/galaxies.py:311-347
//         ipix = ang2pix(lon, lat); gather kappa, gamma1, gamma2; reduced-shear formula
//   ellipticity_intnorm / ellipticity_gaussian         glass/shapes.py:323-362, 255-285
//   redshifts_from_nz inverse-CDF draw                 glass/galaxies.py:77-89
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "healpix_geom.cuh"
#include "rng.cuh"

namespace glb {

__global__ void __launch_bounds__(256) multiplane_update_kernel(double* __restrict__ k3, const double* __restrict__ k2,
                                                                const double* __restrict__ delta2, int64_t npix, double t,
                                                                double f, double delta2_scalar) {
  const double omt = 1.0 - t;  // host computes the same (1 - t) in double
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (int64_t)gridDim.x * blockDim.x) {
    double k = __dmul_rn(k3[i], omt);
    k = __dadd_rn(k, __dmul_rn(t, k2[i]));
    const double d = delta2 ? delta2[i] : delta2_scalar;
    k = __dadd_rn(k, __dmul_rn(f, d));
    k3[i] = k;
  }
}

__device__ __forceinline__ double2 cdiv(double2 a, double2 b) {
  const double den = b.x * b.x + b.y * b.y;
  return make_double2((a.x * b.x + a.y * b.y) / den, (a.y * b.x - a.x * b.y) / den);
}

__global__ void __launch_bounds__(256) galaxy_shear_kernel(int64_t nside, const double* __restrict__ lon,
                                                           const double* __restrict__ lat, const int64_t* __restrict__ ipix_in,
                                                           const double2* __restrict__ eps, int64_t n,
                                                           const double* __restrict__ kappa, const double* __restrict__ g1,
                                                           const double* __restrict__ g2, int reduced,
                                                           double2* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int64_t ip;
  if (ipix_in) {
    ip = ipix_in[i];
  } else {
    const double deg2rad = 0.017453292519943295769;
    const double theta = (90.0 - lat[i]) * deg2rad;
    double s, c;
    sincos(theta, &s, &c);
    ip = zphi2pix_ring(nside, c, s, lon[i] * deg2rad);
  }
  const double k = kappa[ip];
  double2 g = make_double2(g1[ip], g2[ip]);
  const double2 e = eps[i];
  if (reduced) {
    const double d = 1.0 - k;
    g.x /= d;
    g.y /= d;
    // (eps + g) / (1 + conj(g) eps)
    const double2 num = make_double2(e.x + g.x, e.y + g.y);
    const double2 ge = make_double2(g.x * e.x + g.y * e.y, g.x * e.y - g.y * e.x);
    g = cdiv(num, make_double2(1.0 + ge.x, ge.y));
  } else {
    g.x += e.x;
    g.y += e.y;
  }
  out[i] = g;
}

// glass.displace (glass/points.py:654-716, sign = +1) and glass.deflect (glass/lensing.py:687-778,
// sign = -1): the exponential map on the sphere.  The point (lon, lat) moves the angular distance
// |alpha| along the geodesic with bearing arg(alpha); same formulas and operation order as the
// reference (great-circle navigation, lat instead of co-lat).
__global__ void __launch_bounds__(256) displace_kernel(const double* __restrict__ lon, const double* __restrict__ lat,
                                                       const double* __restrict__ a1, const double* __restrict__ a2,
                                                       int64_t a_stride, double sign, int64_t n,
                                                       double* __restrict__ out_lon, double* __restrict__ out_lat) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double pi = 3.141592653589793;
  const double t = lat[i] / 180.0 * pi;
  double ct, st;
  sincos(t, &ct, &st);  // sin and cos flipped: lat, not co-lat
  const double x = a1[i * a_stride], y = a2[i * a_stride];
  const double a = hypot(x, y);
  const double g = atan2(y, x);
  double sa, ca, sg, cg;
  sincos(a, &sa, &ca);
  sincos(g, &sg, &cg);
  const double tp = atan2(ct * ca + st * sa * cg, hypot(ct * sa - st * ca * cg, st * sg));
  const double d = atan2(sa * sg, st * ca - ct * sa * cg);
  out_lon[i] = lon[i] + sign * (d / pi * 180.0);
  out_lat[i] = tp / pi * 180.0;
}

// glass.displacement (glass/points.py:719-772): the complex displacement that takes
// (from_lon, from_lat) to (to_lon, to_lat): r e^{i x}.
__global__ void __launch_bounds__(256) displacement_kernel(const double* __restrict__ from_lon,
                                                           const double* __restrict__ from_lat,
                                                           const double* __restrict__ to_lon,
                                                           const double* __restrict__ to_lat, int64_t n,
                                                           double2* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double rad = 0.017453292519943295;
  double sa, ca, sb, cb, sg, cg;
  sincos(from_lat[i] * rad, &sa, &ca);
  sincos(to_lat[i] * rad, &sb, &cb);
  sincos((to_lon[i] - from_lon[i]) * rad, &sg, &cg);
  const double u = cb * sg, v = ca * sb - sa * cb * cg;
  const double r = atan2(hypot(u, v), sa * sb + ca * cb * cg);
  const double x = atan2(u, v);
  double sx, cx;
  sincos(x, &sx, &cx);
  out[i] = make_double2(r * cx, r * sx);
}

__device__ __forceinline__ double2 philox_normal_pair(uint32_t k0, uint32_t k1, uint64_t idx, uint32_t stream, uint32_t tag) {
  const Philox4 r = philox4x32_10((uint32_t)idx, (uint32_t)(idx >> 32), stream, tag, k0, k1);
  const double u1 = u01_open_closed(r.v[0], r.v[1]);
  const double u2 = u01_closed_open(r.v[2], r.v[3]);
  const double rad = sqrt(-2.0 * log(u1));
  double s, c;
  sincospi(2.0 * u2, &s, &c);
  return make_double2(rad * c, rad * s);
}

// glass.gaussian_phz (glass/galaxies.py:350-455): zphot = normal(z, (1 + z) sigma_0), re-drawn while
// outside [lower, upper] (plain rejection, like the reference's while loop).  sigma_0 / lower / upper
// are arrays or scalars.  Philox mode: the whole rejection loop runs in the thread (attempt number in
// the counter).  Supplied-deviate mode replays the reference's rounds: one launch per round with that
// round's full-size normal array; redraw_only = 0 draws every element, 1 only those still out of
// bounds; nbad counts the elements out of bounds after the round.
__global__ void __launch_bounds__(256) gaussian_phz_kernel(const double* __restrict__ z, const double* __restrict__ sigma0_arr,
                                                           double sigma0, const double* __restrict__ lower_arr, double lower,
                                                           const double* __restrict__ upper_arr, double upper,
                                                           const double* __restrict__ normals, int redraw_only, int64_t n,
                                                           uint32_t k0, uint32_t k1, uint32_t stream,
                                                           double* __restrict__ zphot, unsigned long long* __restrict__ nbad) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double zi = z[i];
  const double s0 = sigma0_arr ? sigma0_arr[i] : sigma0;
  const double lo = lower_arr ? lower_arr[i] : lower;
  const double hi = upper_arr ? upper_arr[i] : upper;
  const double sigma = __dmul_rn(__dadd_rn(1.0, zi), s0);  // xp.add(1, z) * sigma_0
  if (normals) {
    double v = zphot[i];
    if (!redraw_only || v < lo || v > hi) v = __dadd_rn(zi, __dmul_rn(sigma, normals[i]));  // loc + scale * N(0,1)
    zphot[i] = v;
    if (v < lo || v > hi) atomicAdd(nbad, 1ull);
    return;
  }
  double v = zi;
  for (uint32_t attempt = 0; attempt < 4096; ++attempt) {
    // two normals per Philox call: even attempts use .x, odd attempts .y
    const double2 nn = philox_normal_pair(k0, k1, (uint64_t)i, stream, RNG_TAG_REDSHIFT + 0x10u + (attempt >> 1));
    v = __dadd_rn(zi, __dmul_rn(sigma, (attempt & 1) ? nn.y : nn.x));
    if (!(v < lo || v > hi)) break;
  }
  zphot[i] = v;
}

// mode 0: intrinsic-normal (shapes.py:323-362), sigma = sigma_eta; mode 1: clipped Gaussian
// (shapes.py:255-285), re-drawn until |e| <= 1.  normals (optional): supplied complex deviates.
__global__ void __launch_bounds__(256) ellipticity_kernel(int mode, double sigma, const double2* __restrict__ normals,
                                                          int64_t n, uint32_t k0, uint32_t k1, uint32_t stream,
                                                          uint64_t index0, double2* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (mode == 0) {
    double2 e = normals ? normals[i] : philox_normal_pair(k0, k1, index0 + (uint64_t)i, stream, RNG_TAG_EPS);
    e.x *= sigma;
    e.y *= sigma;
    const double r = hypot(e.x, e.y);
    const double s = (r > 0.0) ? tanh(r / 2) / r : 1.0;
    out[i] = make_double2(e.x * s, e.y * s);
  } else {
    double2 e;
    for (uint32_t attempt = 0; attempt < 4096; ++attempt) {
      e = philox_normal_pair(k0, k1, index0 + (uint64_t)i, stream, RNG_TAG_EPS + 1u + attempt);
      e.x *= sigma;
      e.y *= sigma;
      if (hypot(e.x, e.y) <= 1.0) break;
    }
    out[i] = e;
  }
}

// z = interp(u, cdf, zgrid)  (np.interp semantics: clamp outside, linear inside)
__global__ void __launch_bounds__(256) redshift_kernel(const double* __restrict__ cdf, const double* __restrict__ zgrid,
                                                       int nz, const double* __restrict__ u_in, int64_t n, uint32_t k0,
                                                       uint32_t k1, uint32_t stream, uint64_t index0,
                                                       double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double u;
  if (u_in) {
    u = u_in[i];
  } else {
    const uint64_t g = index0 + (uint64_t)i;
    const Philox4 r = philox4x32_10((uint32_t)g, (uint32_t)(g >> 32), stream, RNG_TAG_REDSHIFT, k0, k1);
    u = u01_closed_open(r.v[0], r.v[1]);
  }
  if (u <= cdf[0]) {
    out[i] = zgrid[0];
    return;
  }
  if (u >= cdf[nz - 1]) {
    out[i] = zgrid[nz - 1];
    return;
  }
  // largest j with cdf[j] <= u
  int lo = 0, hi = nz - 1;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (cdf[mid] <= u)
      lo = mid;
    else
      hi = mid;
  }
  const double slope = (zgrid[lo + 1] - zgrid[lo]) / (cdf[lo + 1] - cdf[lo]);
  out[i] = slope * (u - cdf[lo]) + zgrid[lo];
}

}  // namespace glb

using namespace glb;

extern "C" {

int glb_multiplane_update(double* d_kappa3, const double* d_kappa2, const double* d_delta2, double delta2_scalar,
                          int64_t npix, double t, double f, void* stream) {
  GLB_REQUIRE(d_kappa3 && d_kappa2 && npix > 0, "null pointer");
  const int blocks = (int)std::min<int64_t>((npix + 255) / 256, 148 * 16);
  multiplane_update_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(d_kappa3, d_kappa2, d_delta2, npix, t, f, delta2_scalar);
  GLB_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return GLB_OK;
}

int glb_galaxy_shear(int64_t nside, const double* d_lon, const double* d_lat, const int64_t* d_ipix, const double* d_eps,
                     int64_t n, const double* d_kappa, const double* d_gamma1, const double* d_gamma2, int reduced_shear,
                     double* d_out, void* stream) {
  GLB_REQUIRE(nside >= 1 && n >= 0, "bad size");
  if (n == 0) return GLB_OK;
  GLB_REQUIRE((d_ipix || (d_lon && d_lat)) && d_eps && d_kappa && d_gamma1 && d_gamma2 && d_out, "null pointer");
  // Experiment knob (off unless set): GLB_L2_FETCH_BYTES=32 asks the L2 for 32-byte instead of 64-byte
  // fetches.  The three map gathers of a galaxy use 8 bytes of what they pull (ncu: 243 B of DRAM reads
  // per galaxy at 87 % of peak DRAM throughput).  Device-wide hint, set once and left in place.
  static const int fetch_bytes = [] {
    const char* env = getenv("GLB_L2_FETCH_BYTES");
    return env ? atoi(env) : 0;
  }();
  static bool fetch_set = false;
  if (fetch_bytes > 0 && !fetch_set) {
    GLB_CUDA_CHECK(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)fetch_bytes));
    fetch_set = true;
  }
  galaxy_shear_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      nside, d_lon, d_lat, d_ipix, reinterpret_cast<const double2*>(d_eps), n, d_kappa, d_gamma1, d_gamma2,
      reduced_shear, reinterpret_cast<double2*>(d_out));
  GLB_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return GLB_OK;
}

int glb_displace(const double* d_lon, const double* d_lat, const double* d_alpha1, const double* d_alpha2,
                 int64_t alpha_stride, int deflect, int64_t n, double* d_out_lon, double* d_out_lat, void* stream) {
  GLB_REQUIRE(n >= 0 && alpha_stride >= 1, "bad size");
  if (n == 0) return GLB_OK;
  GLB_REQUIRE(d_lon && d_lat && d_alpha1 && d_alpha2 && d_out_lon && d_out_lat, "null pointer");
  displace_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      d_lon, d_lat, d_alpha1, d_alpha2, alpha_stride, deflect ? -1.0 : 1.0, n, d_out_lon, d_out_lat);
  GLB_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return GLB_OK;
}

int glb_displacement(const double* d_from_lon, const double* d_from_lat, const double* d_to_lon,
                     const double* d_to_lat, int64_t n, double* d_out, void* stream) {
  GLB_REQUIRE(n >= 0, "bad size");
  if (n == 0) return GLB_OK;
  GLB_REQUIRE(d_from_lon && d_from_lat && d_to_lon && d_to_lat && d_out, "null pointer");
  displacement_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      d_from_lon, d_from_lat, d_to_lon, d_to_lat, n, reinterpret_cast<double2*>(d_out));
  GLB_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return GLB_OK;
}

int glb_gaussian_phz(const double* d_z, const double* d_sigma0, double sigma0, const double* d_lower, double lower,
                     const double* d_upper, double upper, const double* d_normals, int redraw_only, int64_t n,
                     uint64_t seed, uint32_t stream_id, double* d_zphot, int64_t* d_nbad, void* stream) {
  GLB_REQUIRE(n >= 0, "bad size");
  if (n == 0) return GLB_OK;
  GLB_REQUIRE(d_z && d_zphot, "null pointer");
  GLB_REQUIRE(!d_normals || d_nbad, "a supplied-deviate round needs the out-of-bounds counter");
  gaussian_phz_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      d_z, d_sigma0, sigma0, d_lower, lower, d_upper, upper, d_normals, redraw_only, n, (uint32_t)seed,
      (uint32_t)(seed >> 32), stream_id, d_zphot, reinterpret_cast<unsigned long long*>(d_nbad));
  GLB_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return GLB_OK;
}

int glb_ellipticity(int mode, double sigma, const double* d_normals, int64_t n, uint64_t seed, uint32_t stream_id,
                    uint64_t index0, double* d_out, void* stream) {
  GLB_REQUIRE(mode == 0 || mode == 1, "mode must be 0 (intnorm) or 1 (gaussian)");
  GLB_REQUIRE(n >= 0, "bad size");
  if (n == 0) return GLB_OK;
  GLB_REQUIRE(d_out != nullptr, "null pointer");
  ellipticity_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      mode, sigma, reinterpret_cast<const double2*>(d_normals), n, (uint32_t)seed, (uint32_t)(seed >> 32), stream_id,
      index0, reinterpret_cast<double2*>(d_out));
  GLB_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return GLB_OK;
}

int glb_redshifts_from_cdf(const double* d_cdf, const double* d_z, int nz, const double* d_u, int64_t n, uint64_t seed,
                           uint32_t stream_id, uint64_t index0, double* d_out, void* stream) {
  GLB_REQUIRE(nz >= 2 && n >= 0, "bad size");
  if (n == 0) return GLB_OK;
  GLB_REQUIRE(d_cdf && d_z && d_out, "null pointer");
  redshift_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      d_cdf, d_z, nz, d_u, n, (uint32_t)seed, (uint32_t)(seed >> 32), stream_id, index0, d_out);
  GLB_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return GLB_OK;
}

}  // extern "C"
