// healpix_geom.cuh -- HEALPix RING-scheme pixel <-> angle device functions.
// Replaces healpix `_chp.ring2ang_uv` (healpix.randang, glass/healpix.py:426-431) and
// `healpix.ang2pix` (glass/healpix.py:172).  Formulas: Gorski et al. 2005 / healpix_bare,
// SURVEY.md Appendix A.3-A.5.
#pragma once
#include <stdint.h>

namespace glb {

__device__ __constant__ const int8_t c_jrll[12] = {2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4};
__device__ __constant__ const int8_t c_jpll[12] = {1, 3, 5, 7, 0, 2, 4, 6, 1, 3, 5, 7};

// 12 nside^2 <= 2^30: every intermediate of ring2xyf / zphi2pix_ring fits 32 bits, and 32-bit
// integer division costs a quarter of the 64-bit one (these kernels are instruction bound)
__host__ __device__ __forceinline__ bool healpix_fits_i32(int64_t nside) { return 12 * nside * nside <= (1LL << 30); }

template <typename I>
__device__ __forceinline__ I isqrt_t(I a) {
  I r = (I)sqrt((double)a);
  if (r * r > a) --r;
  if ((r + 1) * (r + 1) <= a) ++r;
  return r;
}

// ring pixel -> face and integer in-face coordinates (I = int32_t or int64_t)
template <typename I>
__device__ __forceinline__ void ring2xyf_t(I nside, I pix, int& x, int& y, int& f) {
  const I ncap = 2 * nside * (nside - 1);
  const I npix = 12 * nside * nside;
  I iring, iphi, kshift, nr;
  if (pix < ncap) {
    iring = (1 + isqrt_t<I>(1 + 2 * pix)) >> 1;
    iphi = pix + 1 - 2 * iring * (iring - 1);
    kshift = 0;
    nr = iring;
    f = (int)((iphi - 1) / nr);
  } else if (pix < npix - ncap) {
    const I ip = pix - ncap;
    const I tmp = ip / (4 * nside);
    iring = tmp + nside;
    iphi = ip - tmp * 4 * nside + 1;
    kshift = (iring + nside) & 1;
    nr = nside;
    const I ire = tmp + 1, irm = 2 * nside + 1 - tmp;
    const I ifm = (iphi - ire / 2 + nside - 1) / nside;
    const I ifp = (iphi - irm / 2 + nside - 1) / nside;
    f = (int)((ifp == ifm) ? (ifp | 4) : ((ifp < ifm) ? ifp : (ifm + 8)));
  } else {
    const I ip = npix - pix;
    const I ir = (1 + isqrt_t<I>(2 * ip - 1)) >> 1;
    iphi = 4 * ir + 1 - (ip - 2 * ir * (ir - 1));
    kshift = 0;
    nr = ir;
    iring = 4 * nside - ir;
    f = 8 + (int)((iphi - 1) / nr);
  }
  const I irt = iring - (I)c_jrll[f] * nside + 1;
  I ipt = 2 * iphi - (I)c_jpll[f] * nr - kshift - 1;
  if (ipt >= 2 * nside) ipt -= 8 * nside;
  x = (int)((ipt - irt) >> 1);
  y = (int)((-ipt - irt) >> 1);
}
__device__ __forceinline__ void ring2xyf(int64_t nside, int64_t pix, int& x, int& y, int& f) {
  if (healpix_fits_i32(nside))
    ring2xyf_t<int32_t>((int32_t)nside, (int32_t)pix, x, y, f);
  else
    ring2xyf_t<int64_t>(nside, pix, x, y, f);
}

// continuous face coordinates -> (z, sin theta, phi)
// (inv_nside = 1/nside: exact for the usual power-of-two nside, so X, Y are then the same
// doubles as (x+u)/nside; otherwise within one ulp)
__device__ __forceinline__ void hpc2loc(double nside, int x, int y, int f, double u, double v, double& z, double& sth,
                                        double& phi) {
  const double inv_nside = 1.0 / nside;  // uniform: hoisted by the compiler
  const double X = ((double)x + u) * inv_nside;
  const double Y = ((double)y + v) * inv_nside;
  const double jr = (double)c_jrll[f] - X - Y;
  double tmpphi;
  if (jr < 1.0) {
    const double tmp = jr * jr / 3.0;
    z = 1.0 - tmp;
    sth = sqrt(tmp * (2.0 - tmp));
    tmpphi = (X - Y) / jr;
  } else if (jr > 3.0) {
    const double nr = 4.0 - jr;
    const double tmp = nr * nr / 3.0;
    z = -(1.0 - tmp);
    sth = sqrt(tmp * (2.0 - tmp));
    tmpphi = (X - Y) / nr;
  } else {
    z = (2.0 - jr) * 2.0 / 3.0;
    sth = sqrt(fmax((1.0 - z) * (1.0 + z), 0.0));
    tmpphi = X - Y;
  }
  tmpphi += (double)c_jpll[f];
  if (tmpphi < 0.0) tmpphi += 8.0;
  if (tmpphi >= 8.0) tmpphi -= 8.0;
  phi = 0.78539816339744830962 * tmpphi;
}

// (z, sin theta, phi) -> ring pixel
template <typename I>
__device__ __forceinline__ int64_t zphi2pix_ring_t(I nside, double z, double sth, double phi) {
  const double za = fabs(z);
  const double twopi = 6.283185307179586476925286766559;
  double pm = fmod(phi, twopi);
  if (pm < 0.0) pm += twopi;
  double tt = pm * 0.63661977236758134308;  // 2/pi
  if (tt >= 4.0) tt -= 4.0;
  const I ncap = 2 * nside * (nside - 1);
  if (za <= 2.0 / 3.0) {
    const double t1 = (double)nside * (0.5 + tt);
    const double t2 = (double)nside * z * 0.75;
    const I jp = (I)floor(t1 - t2);
    const I jm = (I)floor(t1 + t2);
    const I ir = nside + 1 + jp - jm;
    const I kshift = 1 - (ir & 1);
    const I t = jp + jm - nside + kshift + 1 + 8 * nside;
    I ip = t >> 1;  // in [4 nside, 8 nside]: modulo 4 nside without a division
    while (ip >= 4 * nside) ip -= 4 * nside;
    return (int64_t)(ncap + (ir - 1) * 4 * nside + ip);
  }
  const double tp = tt - floor(tt);
  const double tmp = (za > 0.99) ? (double)nside * sth / sqrt((1.0 + za) / 3.0) : (double)nside * sqrt(3.0 * (1.0 - za));
  const I jp = (I)(tp * tmp);
  const I jm = (I)((1.0 - tp) * tmp);
  const I ir = jp + jm + 1;
  I ip = (I)(tt * (double)ir);
  if (ip >= 4 * ir) ip -= 4 * ir;
  if (ip < 0) ip += 4 * ir;
  return (int64_t)((z > 0.0) ? 2 * ir * (ir - 1) + ip : 12 * nside * nside - 2 * ir * (ir + 1) + ip);
}
__device__ __forceinline__ int64_t zphi2pix_ring(int64_t nside, double z, double sth, double phi) {
  return healpix_fits_i32(nside) ? zphi2pix_ring_t<int32_t>((int32_t)nside, z, sth, phi)
                                 : zphi2pix_ring_t<int64_t>(nside, z, sth, phi);
}

}  // namespace glb
