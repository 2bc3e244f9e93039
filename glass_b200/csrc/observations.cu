// observations.cu -- visibility-mask construction and the small spectra kernels of the "next" rows
// (SURVEY 8f ranks 2 and 4).
//
//   glb_query_strip        healpy.query_strip behind hp.query_strip           glass/healpix.py:359-396
//   glb_rotate_map_pixel   healpy.Rotator(coord=).rotate_map_pixel            glass/healpix.py:457-471
//                          (both used by glass.vmap_galactic_ecliptic,        glass/observations.py:96-100)
//   glb_cls_window         cl[:n] * pw[:n]**2 of glass.discretized_cls        glass/fields.py:290-299
//   glb_effective_cls      the weighted sum of glass.effective_cls            glass/fields.py:682-691
//
// healpy / healpix_cxx are absent from the reference tree: the strip and the bilinear interpolation
// follow the published HEALPix C++ algorithms (ring_above, query_strip_internal, get_ring_info2,
// get_interpol), healpy's Rotator conventions are applied on the host (glass_b200/healpix.py).
// All of it is HBM/gather-bound element-wise work: one thread per output element, coalesced stores.
#include "common.cuh"
#include "healpix_geom.cuh"

namespace glb {

__host__ __device__ __forceinline__ int64_t ring_above(int64_t nside, double z) {
  const double az = fabs(z);
  if (az <= 2.0 / 3.0) return (int64_t)((double)nside * (2.0 - 1.5 * z));
  const int64_t ir = (int64_t)((double)nside * sqrt(3.0 * (1.0 - az)));
  return z > 0.0 ? ir : 4 * nside - ir - 1;
}

// first pixel and length of ring 0..4 nside-1 (ring 0: the empty "ring" above the pole)
static void ring_small(int64_t n, int64_t ring, int64_t& sp, int64_t& nr) {
  if (ring < n) {
    sp = 2 * ring * (ring - 1);
    nr = 4 * ring;
  } else if (ring < 3 * n) {
    sp = 2 * n * (n - 1) + (ring - n) * 4 * n;
    nr = 4 * n;
  } else {
    const int64_t r = 4 * n - ring;
    sp = 12 * n * n - 2 * r * (r + 1);
    nr = 4 * r;
  }
}

// pixel range [p0, p1) of one query_strip_internal call (RING, inclusive = false)
static void strip_range(int64_t n, double theta1, double theta2, int64_t& p0, int64_t& p1) {
  int64_t ring1 = 1 + ring_above(n, cos(theta1));
  if (ring1 < 1) ring1 = 1;
  int64_t ring2 = ring_above(n, cos(theta2));
  if (ring2 > 4 * n - 1) ring2 = 4 * n - 1;
  int64_t sp1, rp1, sp2, rp2;
  ring_small(n, ring1, sp1, rp1);
  ring_small(n, ring2, sp2, rp2);
  p0 = sp1;
  p1 = sp2 + rp2;
  if (p1 < p0) p1 = p0;
}

__global__ void __launch_bounds__(256) query_strip_kernel(int64_t npix, int64_t a0, int64_t a1, int64_t b0,
                                                          int64_t b1, double* __restrict__ out) {
  const int64_t p = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (p >= npix) return;
  out[p] = ((p >= a0 && p < a1) || (p >= b0 && p < b1)) ? 1.0 : 0.0;
}

// ring 1..4 nside-1 -> first pixel, length, colatitude, shift (get_ring_info2)
__device__ __forceinline__ void ring_info2(int64_t n, int64_t ring, int64_t& sp, int64_t& nr, double& theta,
                                           bool& shifted) {
  const int64_t north = ring > 2 * n ? 4 * n - ring : ring;
  if (north < n) {
    const double tmp = (double)(north * north) / (3.0 * (double)n * (double)n);
    theta = atan2(sqrt(tmp * (2.0 - tmp)), 1.0 - tmp);
    nr = 4 * north;
    shifted = true;
    sp = 2 * north * (north - 1);
  } else {
    theta = acos((double)(2 * n - north) * (2.0 / (3.0 * (double)n)));
    nr = 4 * n;
    shifted = ((north - n) & 1) == 0;
    sp = 2 * n * (n - 1) + (north - n) * 4 * n;
  }
  if (north != ring) {
    theta = 3.141592653589793238462643383279 - theta;
    sp = 12 * n * n - sp - nr;
  }
}

// the two pixels of one ring next to azimuth phi and the weight of the second
__device__ __forceinline__ void ring_pair(int64_t n, int64_t ring, double phi, int64_t& pa, int64_t& pb, double& w1,
                                          double& theta) {
  int64_t sp, nr;
  bool shifted;
  ring_info2(n, ring, sp, nr, theta, shifted);
  const double dphi = 6.283185307179586476925286766559 / (double)nr;
  const double sh = shifted ? 0.5 : 0.0;
  const double tmp = phi / dphi - sh;
  int64_t i1 = tmp < 0.0 ? (int64_t)tmp - 1 : (int64_t)tmp;
  w1 = (phi - ((double)i1 + sh) * dphi) / dphi;
  int64_t i2 = i1 + 1;
  if (i1 < 0) i1 += nr;
  if (i2 >= nr) i2 -= nr;
  pa = sp + i1;
  pb = sp + i2;
}

// out[p] = bilinear interpolation (HEALPix get_interpol) of `in` at R * (centre of pixel p);
// R row-major = the matrix of healpy's Rotator.I (back-rotation into the frame of the input map)
struct Rot3 {
  double m[9];
};

__global__ void __launch_bounds__(256) rotate_map_pixel_kernel(int64_t nside, Rot3 R, const double* __restrict__ in,
                                                               double* __restrict__ out) {
  const int64_t npix = 12 * nside * nside;
  const int64_t p = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (p >= npix) return;
  int x, y, f;
  ring2xyf(nside, p, x, y, f);
  double z, sth, phi;
  hpc2loc((double)nside, x, y, f, 0.5, 0.5, z, sth, phi);
  double sp_, cp_;
  sincos(phi, &sp_, &cp_);
  const double vx = sth * cp_, vy = sth * sp_, vz = z;
  const double rx = R.m[0] * vx + R.m[1] * vy + R.m[2] * vz;
  const double ry = R.m[3] * vx + R.m[4] * vy + R.m[5] * vz;
  const double rz = R.m[6] * vx + R.m[7] * vy + R.m[8] * vz;
  const double theta = atan2(sqrt(rx * rx + ry * ry), rz);
  const double twopi = 6.283185307179586476925286766559;
  double ph = atan2(ry, rx);
  if (ph < 0.0) ph += twopi;
  if (ph >= twopi) ph -= twopi;

  const int64_t ir1 = ring_above(nside, cos(theta)), ir2 = ir1 + 1;
  int64_t p0 = 0, p1 = 0, p2 = 0, p3 = 0;
  double w0 = 0.0, w1 = 0.0, w2 = 0.0, w3 = 0.0, theta1 = 0.0, theta2 = 0.0, w;
  if (ir1 > 0) {
    ring_pair(nside, ir1, ph, p0, p1, w, theta1);
    w0 = 1.0 - w;
    w1 = w;
  }
  if (ir2 < 4 * nside) {
    ring_pair(nside, ir2, ph, p2, p3, w, theta2);
    w2 = 1.0 - w;
    w3 = w;
  }
  if (ir1 == 0) {
    const double wt = theta / theta2, fac = (1.0 - wt) * 0.25;
    w2 = w2 * wt + fac;
    w3 = w3 * wt + fac;
    w0 = fac;
    w1 = fac;
    p0 = (p2 + 2) & 3;
    p1 = (p3 + 2) & 3;
  } else if (ir2 == 4 * nside) {
    const double wt = (theta - theta1) / (3.141592653589793238462643383279 - theta1), fac = wt * 0.25;
    w0 = w0 * (1.0 - wt) + fac;
    w1 = w1 * (1.0 - wt) + fac;
    w2 = fac;
    w3 = fac;
    p2 = ((p0 + 2) & 3) + npix - 4;
    p3 = ((p1 + 2) & 3) + npix - 4;
  } else {
    const double wt = (theta - theta1) / (theta2 - theta1);
    w0 *= 1.0 - wt;
    w1 *= 1.0 - wt;
    w2 *= wt;
    w3 *= wt;
  }
  // np.sum(m[p] * w, 0): rows added in order, every product and sum rounded
  double acc = __dmul_rn(__ldg(in + p0), w0);
  acc = __dadd_rn(acc, __dmul_rn(__ldg(in + p1), w1));
  acc = __dadd_rn(acc, __dmul_rn(__ldg(in + p2), w2));
  acc = __dadd_rn(acc, __dmul_rn(__ldg(in + p3), w3));
  out[p] = acc;
}

// spectra packed [nspec][ld]; out[s][l] = cl[s][l] * (pw[l] * pw[l])  for l < n, separately rounded
__global__ void __launch_bounds__(256) cls_window_kernel(int nspec, int n, int64_t ld_in, int64_t ld_out,
                                                         const double* __restrict__ cl, const double* __restrict__ pw,
                                                         double* __restrict__ out) {
  const int l = blockIdx.x * 256 + threadIdx.x;
  const int s = blockIdx.y;
  if (l >= n || s >= nspec) return;
  const double w = pw[l];
  out[(int64_t)s * ld_out + l] = __dmul_rn(cl[(int64_t)s * ld_in + l], __dmul_rn(w, w));
}

// out[j1][j2][l] = sum_{i1} sum_{i2} (w1[i1][j1] * w2[i2][j2]) * C_l^{i1 i2}, accumulated from 0.0 in
// the reference's order (i1 outer, i2 inner); C^{ij} = row i(i+1)/2 + i - j (i >= j) of cls[nspec][ld]
__global__ void __launch_bounds__(128) effective_cls_kernel(int nf, int J1, int J2, int L, int64_t ld, int symmetric,
                                                            const double* __restrict__ cls,
                                                            const double* __restrict__ w1,
                                                            const double* __restrict__ w2, double* __restrict__ out) {
  const int l = blockIdx.x * 128 + threadIdx.x;
  const int jo2 = blockIdx.y, jo1 = blockIdx.z;
  if (l >= L) return;
  // weights2 is weights1: the reference computes j1 <= j2 only and copies the transpose
  const bool swap = symmetric && jo1 > jo2;
  const int j1 = swap ? jo2 : jo1, j2 = swap ? jo1 : jo2;
  double acc = 0.0;
  for (int i1 = 0; i1 < nf; ++i1) {
    const double a = w1[(int64_t)i1 * J1 + j1];
    for (int i2 = 0; i2 < nf; ++i2) {
      const int hi = i1 > i2 ? i1 : i2, lo = i1 > i2 ? i2 : i1;
      const double c = cls[(int64_t)(hi * (hi + 1) / 2 + hi - lo) * ld + l];
      acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(a, w2[(int64_t)i2 * J2 + j2]), c));
    }
  }
  out[((int64_t)jo1 * J2 + jo2) * L + l] = acc;
}

}  // namespace glb

using namespace glb;

extern "C" {

int glb_query_strip(int64_t nside, double theta1, double theta2, double* d_mask, void* stream) {
  GLB_REQUIRE(nside >= 1 && nside <= (1 << 24), "bad nside");
  GLB_REQUIRE(d_mask != nullptr, "null pointer");
  int64_t a0 = 0, a1 = 0, b0 = 0, b1 = 0;
  if (theta1 < theta2) {
    strip_range(nside, theta1, theta2, a0, a1);
  } else {  // the complement: [0, theta2] and [theta1, pi]
    strip_range(nside, 0.0, theta2, a0, a1);
    strip_range(nside, theta1, 3.141592653589793238462643383279, b0, b1);
  }
  const int64_t npix = 12 * nside * nside;
  query_strip_kernel<<<(unsigned)((npix + 255) / 256), 256, 0, (cudaStream_t)stream>>>(npix, a0, a1, b0, b1, d_mask);
  GLB_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return GLB_OK;
}

int glb_rotate_map_pixel(int64_t nside, const double* rot9, const double* d_in, double* d_out, void* stream) {
  GLB_REQUIRE(nside >= 1 && nside <= (1 << 24), "bad nside");
  GLB_REQUIRE(rot9 && d_in && d_out, "null pointer");
  GLB_REQUIRE(d_in != d_out, "rotate_map_pixel cannot work in place");
  Rot3 R;
  for (int i = 0; i < 9; ++i) R.m[i] = rot9[i];
  const int64_t npix = 12 * nside * nside;
  rotate_map_pixel_kernel<<<(unsigned)((npix + 255) / 256), 256, 0, (cudaStream_t)stream>>>(nside, R, d_in, d_out);
  GLB_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return GLB_OK;
}

int glb_cls_window(int nspec, int n, int64_t ld_in, int64_t ld_out, const double* d_cl, const double* d_pw,
                   double* d_out, void* stream) {
  GLB_REQUIRE(nspec >= 0 && n >= 0 && ld_in >= n && ld_out >= n, "bad size");
  if (nspec == 0 || n == 0) return GLB_OK;
  GLB_REQUIRE(nspec <= 65535, "too many spectra for one launch");
  GLB_REQUIRE(d_cl && d_pw && d_out, "null pointer");
  cls_window_kernel<<<dim3((unsigned)((n + 255) / 256), (unsigned)nspec), 256, 0, (cudaStream_t)stream>>>(
      nspec, n, ld_in, ld_out, d_cl, d_pw, d_out);
  GLB_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return GLB_OK;
}

int glb_effective_cls(int nf, int J1, int J2, int L, int64_t ld, int symmetric, const double* d_cls,
                      const double* d_w1, const double* d_w2, double* d_out, void* stream) {
  GLB_REQUIRE(nf >= 1 && J1 >= 1 && J2 >= 1 && L >= 0 && ld >= L, "bad size");
  if (L == 0) return GLB_OK;
  GLB_REQUIRE(J1 <= 65535 && J2 <= 65535, "too many weight columns for one launch");
  GLB_REQUIRE(d_cls && d_w1 && d_w2 && d_out, "null pointer");
  GLB_REQUIRE(!symmetric || (d_w1 == d_w2 && J1 == J2), "symmetric needs one weight array");
  effective_cls_kernel<<<dim3((unsigned)((L + 127) / 128), (unsigned)J2, (unsigned)J1), 128, 0,
                         (cudaStream_t)stream>>>(nf, J1, J2, L, ld, symmetric, d_cls, d_w1, d_w2, d_out);
  GLB_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return GLB_OK;
}

}  // extern "C"
