/*
 * glass_b200.h -- C ABI of libglassb200.so: the B200-native (sm_100a) kernels behind
 * the GLASS per-shell field-generation hot path.
 *
 * This is the drop-in boundary.  The reference (glass-dev/glass) is pure Python and
 * reaches native code only through the seam glass/healpix.py (healpy / healpix C
 * extensions) and NumPy.  Every entry point below names the reference interface it
 * replaces (file:line relative to the reference checkout).  INTEGRATION.md shows the
 * ctypes stub a GLASS maintainer would add in glass/healpix.py to bind them.
 *
 * Conventions
 *  - plain pointers and sizes only; no torch / CUDA types in signatures
 *    (`stream` is a cudaStream_t passed as void*; NULL = default stream).
 *  - `d_` pointers are DEVICE pointers on the plan's device, `h_` are HOST pointers.
 *  - complex128 arrays are interleaved (re, im) doubles.
 *  - maps are HEALPix RING order float64 of length 12*nside^2.
 *  - alm are m-major ("HEALPix order", glass/fields.py:959-962):
 *        index(l, m) = m*(2*lmax+1-m)/2 + l,  0 <= m <= l <= lmax  (mmax == lmax).
 *  - all calls are asynchronous on `stream`, return 0 (GLB_OK) or a negative glb_status;
 *    no call allocates caller-visible memory; a plan is not thread-safe (one per
 *    host thread), distinct plans are independent.
 */
#ifndef GLASS_B200_H
#define GLASS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum glb_status {
  GLB_OK = 0,
  GLB_ERR_INVALID_ARG = -1,   /* -> ValueError                                              */
  GLB_ERR_UNSUPPORTED = -2,   /* size outside what the kernels handle (e.g. nside > 4096)   */
  GLB_ERR_CUDA = -3,          /* a CUDA runtime call failed; see glb_last_error()           */
  GLB_ERR_NOMEM = -4,
  GLB_ERR_NOT_POSDEF = -10,   /* "covariance matrix is not positive definite" fields.py:181 */
  GLB_ERR_NEGATIVE_CL = -11   /* "negative values in cl"                      fields.py:229 */
} glb_status;

/* pixel transform fused into the ring-FFT epilogue (glass/grf/_transformations.py) */
typedef enum glb_transform {
  GLB_T_NORMAL = 0,         /* Normal.__call__       :26   t(x) = x                            */
  GLB_T_LOGNORMAL = 1,      /* Lognormal.__call__    :83   t(x) = lamda*expm1(x - var/2)       */
  GLB_T_SQUARED_NORMAL = 2  /* SquaredNormal.__call__:170  t(x) = lamda*((x-a)^2 - 1)          */
} glb_transform;

typedef struct glb_plan glb_plan; /* opaque: ring tables, twiddles, chirp spectra, workspace */

const char* glb_version(void);
const char* glb_last_error(void);      /* thread-local description of the last failure */
const char* glb_status_string(int status);

/* ---- plan -------------------------------------------------------------------------- */
/* One plan per (nside, lmax, device).  Replaces the implicit geometry set-up inside
 * healpy.alm2map / map2alm (glass/healpix.py:71,270).  max_batch = how many maps one
 * call may transform together (workspace is sized for it). */
int glb_plan_create(glb_plan** plan, int nside, int lmax, int max_batch, int device);
int glb_plan_destroy(glb_plan* plan);
int glb_plan_info(const glb_plan* plan, int* nside, int* lmax, int64_t* npix, int64_t* nalm,
                  int* max_batch, int64_t* workspace_bytes);

/* ---- spherical-harmonic transforms (the seam glass/healpix.py) ---------------------- */
/* healpy.alm2map(alm, nside, pol=False, pixwin=False)   glass/healpix.py:71
 * (called from glass/fields.py:429, glass/lensing.py:326).
 * nmaps alm sets [nmaps][nalm] complex128 -> nmaps maps [nmaps][npix] float64.
 * transform/tparams (may be NULL): per-map glb_transform and 2 doubles (p0, p1):
 *   LOGNORMAL: p0 = var/2, p1 = lamda;  SQUARED_NORMAL: p0 = a, p1 = lamda. */
int glb_alm2map(glb_plan* plan, const double* d_alm, int nmaps, double* d_map,
                const int* h_transform, const double* h_tparams, void* stream);

/* healpy.alm2map_spin([alm1, alm2], nside, spin, lmax)   glass/healpix.py:107
 * (called from glass/lensing.py:343,366,428).  d_alm2 may be NULL (B-modes zero, which
 * is what GLASS always passes). */
int glb_alm2map_spin(glb_plan* plan, const double* d_alm1, const double* d_alm2, int spin,
                     double* d_map1, double* d_map2, void* stream);
/* The same for the E modes of nb = 1..4 map pairs at once (several convergence planes sheared
 * together, glass/lensing.py:428 once per plane in the reference): d_alms [nb][nalm] complex128,
 * d_maps1 / d_maps2 [nb][npix].  The planes share the Wigner-d recurrences. */
int glb_alm2map_spin_batch(glb_plan* plan, const double* d_alms, int nb, int spin, double* d_maps1, double* d_maps2,
                           void* stream);

/* healpy.map2alm(map, lmax, pol=False, use_pixel_weights=True)   glass/healpix.py:270
 * (called from glass/lensing.py:306,408).  d_ring_weights: per-ring quadrature weights
 * [4*nside-1] or NULL (uniform); niter Jacobi refinements (healpy default 3). */
int glb_map2alm(glb_plan* plan, const double* d_map, const double* d_ring_weights, int niter,
                double* d_alm, void* stream);
/* The same for nb = 1, 2 or 4 maps at once (d_maps [nb][npix], d_alms [nb][nalm], nb <= the
 * plan's max_batch): healpy.map2alm accepts a sequence of maps (pol=False).  The synthesis of
 * every refinement runs as one batched transform. */
int glb_map2alm_batch(glb_plan* plan, const double* d_maps, int nb, const double* d_ring_weights, int niter,
                      double* d_alms, void* stream);

/* healpy.almxfl(alm, fl)   glass/healpix.py:136 (called from glass/lensing.py:322,339,363,425)
 * in place; fl has nfl entries, treated as zero beyond. */
int glb_almxfl(int lmax, double* d_alm, const double* d_fl, int nfl, void* stream);

/* ---- correlated a_lm sampling across shells (glass/fields.py:404-425) ------------------ */
/* z = rng.standard_normal((N_lm, 2)) @ [1, 1j]   fields.py:407.  Counter-based
 * Philox4x32-10 + Box-Muller keyed by (seed, shell, GLASS index l(l+1)/2+m); the result is
 * written in m-major order (so _glass_to_healpix_alm, fields.py:943-962, is folded in). */
int glb_alm_draw(int lmax, uint64_t seed, uint32_t shell, double* d_z, void* stream);
/* _glass_to_healpix_alm   fields.py:943-962: l-major -> m-major gather (for supplied z). */
int glb_alm_glass_to_healpix(int lmax, const double* d_in, double* d_out, void* stream);
/* iternorm   fields.py:101-188, one step (one shell) for all n = lmax+1 multipoles at once.
 * State (caller-allocated, zero-initialised, kept between steps): d_m [k][k][n], d_a [k][n],
 * d_s [n]; d_tmp [k][n] scratch.  d_row: [n][k+1] row of cls2cov (fields.py:191-236), d_w:
 * [n][k+1] result [a, s].  first != 0 on the first step.  *d_flag is OR-ed with 1 where the
 * reference raises "covariance matrix is not positive definite" (fields.py:184-186). */
int glb_iternorm_step(int n, int k, int first, const double* d_row, double* d_m, double* d_a, double* d_s,
                      double* d_tmp, double* d_w, int* d_flag, void* stream);
/* alm = sum_i multalm(z_i, w[:, i])   fields.py:420 + harmonics.py:46-47, then the m = 0
 * fix alm = Re + Im (fields.py:425).  h_zptrs: HOST array of nterms DEVICE pointers to
 * m-major z arrays, oldest first; d_w: [lmax+1][w_stride] float64, column i scales z_i.
 * Operation order and rounding match NumPy (no FMA contraction): bit-exact for given z.
 * Any nterms >= 1 (the reference has no limit, ncorr=None correlates all shells): above 64 terms
 * the sum continues over further launches in the same left-to-right order. */
int glb_alm_combine(int lmax, int nterms, const double* const* h_zptrs, const double* d_w, int w_stride,
                    double* d_alm, void* stream);

/* ---- galaxy counts and positions (glass/points.py:520-540) --------------------------- */
/* bytes of scratch glb_points_counts needs for a map of npix pixels */
size_t glb_points_workspace_bytes(int64_t npix);
/* One population: biased density -> expected count -> Poisson count per pixel, plus the
 * exclusive prefix sum of the counts.  Replaces _compute_density_contrast (points.py:243-249;
 * bias_model 0 = none/copy, 1 = linear_bias :134, 2 = loglinear_bias :157-160),
 * _compute_expected_count (:279-288; scale = ARCMIN2_SPHERE/npix*ngal computed by the caller,
 * remove_monopole subtracts the map mean), _apply_visibility (:314-316; d_vis may be NULL) and
 * _sample_number_galaxies (:340-348; Philox Poisson keyed by (seed, stream_id, pixel), or, in
 * parity mode, d_counts_in supplies the deviates).  Outputs, each optional (NULL = not wanted):
 * d_counts [npix] int64; d_off [npix+1] int64 (d_off[p] = galaxies before pixel p, d_off[npix] =
 * total); d_nbar_out [npix] = expected counts before clipping; d_total [1] = number of galaxies;
 * and the GALAXY LIST d_gpix [gpix_capacity] int64: ring pixel of every galaxy, galaxies in pixel
 * order = np.repeat(np.arange(npix), counts) (points.py:426).  Only the first gpix_capacity galaxies
 * are listed -- the total is exact regardless, so a caller whose guess was too small calls again
 * with a larger list (the counts are a function of (seed, stream_id, pixel) and come out the same).
 * For sparse maps the list replaces counts and offsets altogether (8 B per galaxy instead of 16 B
 * per pixel): glb_points_cuts_list and glb_points_fill_list work from it alone. */
int glb_points_counts(int64_t npix, const double* d_delta, const double* d_vis, int bias_model, double bias,
                      double scale, int remove_monopole, const int64_t* d_counts_in, uint64_t seed,
                      uint32_t stream_id, double* d_nbar_out, int64_t* d_counts, int64_t* d_off, int64_t* d_gpix,
                      int64_t gpix_capacity, int64_t* d_total, void* d_workspace, void* stream);
/* The pixel ranges of the batches of _sample_galaxies_per_pixel (points.py:409-437: 1000-pixel
 * stepping, searchsorted(side="right"), "first pixel alone" rule) from the exclusive scan d_off,
 * walked on the device: up to max_cuts cuts from pixel `start` with `remaining` galaxies to hand out.
 * d_cuts [max_cuts][3] = {start, stop, galaxies}; d_state [3] = {cuts written, next start, galaxies
 * remaining} -- call again from there while galaxies remain. */
int glb_points_cuts(const int64_t* d_off, int64_t npix, int64_t batch, int64_t start, int64_t remaining,
                    int max_cuts, int64_t* d_cuts, int64_t* d_state, void* stream);
/* The same cuts from the galaxy list of glb_points_counts (d_gpix [total]); `remaining` of the
 * `total` galaxies are still to hand out, i.e. the next batch starts at galaxy total - remaining. */
int glb_points_cuts_list(const int64_t* d_gpix, int64_t total, int64_t npix, int64_t batch, int64_t start,
                         int64_t remaining, int max_cuts, int64_t* d_cuts, int64_t* d_state, void* stream);
/* healpix.randang(nside, ipix, lonlat=True) (points.py:427 -> glass/healpix.py:426-431) for the
 * galaxies [g0, g1) of the list: d_lon/d_lat [g1 - g0]; (u, v) from Philox keyed by (seed,
 * stream_id, global galaxy index) -- the same draws as glb_points_fill -- or supplied arrays
 * [g1 - g0] (parity mode).  The galaxies' pixel indices are d_gpix[g0:g1] itself. */
int glb_points_fill_list(int64_t nside, const int64_t* d_gpix, int64_t g0, int64_t g1, const double* d_u,
                         const double* d_v, uint64_t seed, uint32_t stream_id, double* d_lon, double* d_lat,
                         void* stream);
/* Positions of every galaxy in ring pixels [pix0, pix1): ipix = repeat(arange, n) (points.py:426)
 * and healpix.randang(nside, ipix, lonlat=True) (points.py:427 -> glass/healpix.py:426-431),
 * written at index d_off[p] - d_off[pix0] + i.  (u, v) in-pixel offsets: Philox keyed by
 * (seed, stream_id, global galaxy index) or supplied arrays (parity mode).  d_ipix optional. */
int glb_points_fill(int64_t nside, const int64_t* d_counts, const int64_t* d_off, int64_t pix0, int64_t pix1,
                    const double* d_u, const double* d_v, uint64_t seed, uint32_t stream_id, double* d_lon,
                    double* d_lat, int64_t* d_ipix, void* stream);
/* healpix `_chp.ring2ang_uv` behind healpix.randang (glass/healpix.py:426-431): lonlat=1 ->
 * (lon, lat) degrees, else (theta, phi) radians. */
int glb_ring2ang_uv(int64_t nside, const int64_t* d_ipix, const double* d_u, const double* d_v, int64_t n,
                    int lonlat, double* d_out1, double* d_out2, void* stream);
/* healpix.randang(nside, ipix, lonlat) as called at glass/healpix.py:426-431: in-pixel
 * offsets from Philox keyed by (seed, stream_id, element index) -- the same sequence on every
 * call with the same seed, like the reference's fresh default_rng(42) per call. */
int glb_randang(int64_t nside, const int64_t* d_ipix, int64_t n, uint64_t seed, uint32_t stream_id, int lonlat,
                double* d_out1, double* d_out2, void* stream);
/* glass.uniform_positions (glass/points.py:543-607), one population of n points: lon = uniform(-180, 180),
 * lat = degrees(asin(uniform(-1, 1))) (points.py:594-595).  The two uniforms per point come from Philox
 * keyed by (seed, stream_id, point index) or from d_u1/d_u2 (parity mode, both or neither). */
int glb_uniform_positions(int64_t n, const double* d_u1, const double* d_u2, uint64_t seed, uint32_t stream_id,
                          double* d_lon, double* d_lat, void* stream);
/* glass.gaussian_phz (glass/galaxies.py:350-455): zphot = normal(z, (1+z) sigma_0) with rejection outside
 * [lower, upper].  sigma_0 / lower / upper: device array (per galaxy) or, if the pointer is NULL, the scalar.
 * d_normals == NULL: Philox deviates keyed by (seed, stream_id, galaxy, attempt), whole rejection loop in
 * one launch.  d_normals != NULL (parity mode): ONE round of the reference's loop with that round's
 * full-size normal array; redraw_only = 0 draws every element, 1 only those out of bounds; *d_nbad (int64,
 * caller-zeroed) receives the number of elements still out of bounds. */
int glb_gaussian_phz(const double* d_z, const double* d_sigma0, double sigma0, const double* d_lower, double lower,
                     const double* d_upper, double upper, const double* d_normals, int redraw_only, int64_t n,
                     uint64_t seed, uint32_t stream_id, double* d_zphot, int64_t* d_nbad, void* stream);
/* healpix.ang2pix(nside, theta|lon, phi|lat, lonlat)   glass/healpix.py:172 (RING scheme). */
int glb_ang2pix(int64_t nside, const double* d_a, const double* d_b, int64_t n, int lonlat, int64_t* d_ipix,
                void* stream);

/* ---- lensing in pixel / galaxy space -------------------------------------------------- */
/* MultiPlaneConvergence.add_plane update (glass/lensing.py:584-586), one fused pass:
 * kappa3 = ((kappa3*(1-t)) + t*kappa2) + f*delta2, rounded like NumPy's three passes.
 * d_delta2 NULL -> the scalar delta2_scalar is used (first plane: delta is 0-d, :486). */
int glb_multiplane_update(double* d_kappa3, const double* d_kappa2, const double* d_delta2, double delta2_scalar,
                          int64_t npix, double t, double f, void* stream);
/* galaxy_shear (glass/galaxies.py:311-347): pixel lookup via ang2pix(lon, lat, lonlat=True)
 * -- or d_ipix if the caller already knows it -- gather kappa, gamma1, gamma2, then
 * g = gamma/(1-kappa); (eps+g)/(1+conj(g) eps)  [reduced_shear]  or  gamma + eps.
 * d_eps, d_out complex128 [n]. */
int glb_galaxy_shear(int64_t nside, const double* d_lon, const double* d_lat, const int64_t* d_ipix,
                     const double* d_eps, int64_t n, const double* d_kappa, const double* d_gamma1,
                     const double* d_gamma2, int reduced_shear, double* d_out, void* stream);
/* glass.displace (glass/points.py:654-716; deflect = 0) and glass.deflect (glass/lensing.py:687-778;
 * deflect = 1, the displaced longitude is lon - d instead of lon + d): exponential map on the
 * sphere, degrees in and out.  alpha = d_alpha1 + i d_alpha2, element i at [i * alpha_stride]
 * (stride 2 with d_alpha2 = d_alpha1 + 1 for a complex128 array, stride 1 for two real arrays). */
int glb_displace(const double* d_lon, const double* d_lat, const double* d_alpha1, const double* d_alpha2,
                 int64_t alpha_stride, int deflect, int64_t n, double* d_out_lon, double* d_out_lat, void* stream);
/* glass.displacement (glass/points.py:719-772): complex displacement from -> to, d_out complex128 [n]. */
int glb_displacement(const double* d_from_lon, const double* d_from_lat, const double* d_to_lon,
                     const double* d_to_lat, int64_t n, double* d_out, void* stream);
/* ellipticity_intnorm (glass/shapes.py:323-362; mode 0, sigma = sigma_eta computed by the
 * caller) / ellipticity_gaussian (shapes.py:255-285; mode 1, redraw while |e| > 1).
 * d_normals (mode 0 only, may be NULL): supplied complex standard normals. */
int glb_ellipticity(int mode, double sigma, const double* d_normals, int64_t n, uint64_t seed, uint32_t stream_id,
                    uint64_t index0, double* d_out, void* stream);
/* _draw_nz (glass/galaxies.py:77-89): z = interp(U[0,1), cdf, zgrid); d_u may be NULL. */
int glb_redshifts_from_cdf(const double* d_cdf, const double* d_z, int nz, const double* d_u, int64_t n,
                           uint64_t seed, uint32_t stream_id, uint64_t index0, double* d_out, void* stream);

/* ---- visibility masks and spectra helpers --------------------- */
/* hp.query_strip(nside, thetas, dtype=float64) (glass/healpix.py:359-396 -> healpy.query_strip,
 * RING, inclusive = False): d_mask[p] = 1.0 for pixels whose centre ring lies in the colatitude
 * strip, 0.0 elsewhere; theta1 >= theta2 selects the complement [0, theta2] + [theta1, pi]. */
int glb_query_strip(int64_t nside, double theta1, double theta2, double* d_mask, void* stream);
/* hp.Rotator(coord=).rotate_map_pixel(m) (glass/healpix.py:457-471 -> healpy): d_out[p] = HEALPix
 * bilinear interpolation (get_interp_val, four pixels on the two neighbouring rings) of d_in at
 * rot9 * (centre of pixel p); rot9 = row-major 3x3 matrix of the BACK rotation (healpy's
 * Rotator.I), host memory.  Not in place. */
int glb_rotate_map_pixel(int64_t nside, const double* rot9, const double* d_in, double* d_out, void* stream);
/* glass.discretized_cls (glass/fields.py:290-299), the window step on spectra packed as rows:
 * d_out[s * ld_out + l] = d_cl[s * ld_in + l] * (d_pw[l] * d_pw[l]) for s < nspec, l < n. */
int glb_cls_window(int nspec, int n, int64_t ld_in, int64_t ld_out, const double* d_cl, const double* d_pw,
                   double* d_out, void* stream);
/* glass.effective_cls (glass/fields.py:682-691): d_out[(j1 * J2 + j2) * L + l] =
 * sum_{i1, i2 < nf} (d_w1[i1 * J1 + j1] * d_w2[i2 * J2 + j2]) * C_l^{i1 i2}, accumulated in the
 * reference's order; C^{ij} (i >= j) is row i (i + 1) / 2 + i - j of d_cls[nspec][ld] (zero padded).
 * symmetric = 1 (weights2 is weights1, d_w2 == d_w1): elements with j1 > j2 are the transposed element's
 * sum, as the reference copies them (fields.py:689-690). */
int glb_effective_cls(int nf, int J1, int J2, int L, int64_t ld, int symmetric, const double* d_cls,
                      const double* d_w1, const double* d_w2, double* d_out, void* stream);

/* host-buffer forms of the two transforms GLASS calls per shell: H2D, kernels, D2H on
 * `stream`, synchronous on return.  These are what a ctypes binding inside
 * glass/healpix.py would call with NumPy buffers. */
int glb_alm2map_host(glb_plan* plan, const double* h_alm, int nmaps, double* h_map,
                     const int* h_transform, const double* h_tparams, void* stream);

/* ---- m-split of ONE transform over the GPUs of a box (axis 2) ----------------
 * Legendre stage m-sharded (rank r owns m = r mod world), Fourier stage ring-sharded; between
 * them the caller does ONE all-to-all of the phase array over NVLink (torch.distributed /
 * ncclAllToAll).  Layouts are chosen so that both sides of the all-to-all are contiguous:
 *   send  [map][row][W]           row = h_rowmap[ring]: rings grouped by owning rank
 *   recv  [map][src rank][local row][W]      W = ceil((lmax+1)/world) m slots, slot = m / world
 * h_my_rings lists the rings this rank owns, in local-row order (see glass_b200/sharding.py). */
int glb_dist_setup(glb_plan* plan, int world, int rank, const int* h_rowmap, const int* h_my_rings,
                   int n_my_rings);
int glb_dist_alm2phase(glb_plan* plan, const double* d_alm, int nmaps, double* d_send, void* stream);
int glb_dist_phase2map(glb_plan* plan, const double* d_recv, int nmaps, double* d_map,
                       const int* h_transform, const double* h_tparams, void* stream);

/* Fused form of the same split: the Legendre kernel stores every F_m(ring) straight into the
 * receive buffer of the rank that owns the ring, through peer mappings of those buffers
 * (NVLink / NVSwitch, CUDA IPC between the per-GPU processes), so the m -> ring transpose
 * overlaps the FP64 work tile by tile and no all-to-all runs afterwards.
 *   1. glb_dist_setup, then glb_dist_p2p_alloc (same nmaps_max on every rank): two receive
 *      buffers [nmaps_max][world][rows][W] are allocated as one block; h_handle receives its
 *      64-byte IPC handle;
 *   2. the caller gathers the handles of all ranks (any host exchange) and calls
 *      glb_dist_p2p_open(all handles [world][64], rows per rank [world]);
 *   3. per transform: glb_dist_alm2phase_p2p(buffer = 0/1 alternating), a barrier over the
 *      ranks ordered on `stream` (every rank's stores are complete when its kernel has ended),
 *      then glb_dist_phase2map on the pointer glb_dist_p2p_recv returns.  Alternating the two
 *      buffers makes that one barrier per transform also protect the buffer a slower rank is
 *      still reading.
 * Same kernels and the same bits as the all-to-all form (and as one GPU). */
int glb_dist_p2p_alloc(glb_plan* plan, int nmaps_max, void* h_handle);
int glb_dist_p2p_open(glb_plan* plan, const void* h_all_handles, const int* h_rows);
int glb_dist_alm2phase_p2p(glb_plan* plan, const double* d_alm, int nmaps, int buffer, void* stream);
int glb_dist_p2p_recv(glb_plan* plan, int buffer, double** d_recv);

/* ---- measurement hooks (bench.py) --------------------------------------------------- */
/* Per-stage device time of glb_alm2map, from CUDA events recorded on the launch stream:
 * ms3 / launches3 = {prep, Legendre, ring FFT}; nmaps = maps transformed since enable. */
int glb_plan_timing_enable(glb_plan* plan, int enable);
int glb_plan_timing_read(glb_plan* plan, double* ms3, int64_t* launches3, int64_t* nmaps);
/* number of kernels this library has launched in this process */
uint64_t glb_kernel_launch_count(void);
/* FP64 roofline denominator: register-resident DFMA chains on every SM, timed with CUDA
 * events (MEASURED_PEAKS.json has no FP64 entry). */
int glb_measure_fp64_peak(int device, double* tflops, double* ms, void* stream);

/* Where the Legendre stage of glb_alm2map runs its contraction over l (the recurrence itself is always FP64):
 * 0 = auto: groups of EIGHT maps on the INT8 tensor cores (tcgen05.mma kind::i8, exact integer digit products,
 * csrc/sht_ozaki.cu) at nside >= 1024, everything else on the FP64 pipe; 1 = FP64 pipe only; 2 = INT8 tensor cores
 * for groups of four and eight maps at any nside.  Both agree to ~1e-11 of the largest phase. */
int glb_plan_set_legendre_mode(glb_plan* plan, int mode);
/* glb_alm2map for EIGHT maps in two halves (the INT8 path, same results): `prepare` turns the a_lm into Legendre records and
 * digit planes (set `slot` = 0 or 1 of the plan's tile blocks), `finish` runs the Legendre kernel on that set and the ring
 * FFTs with the fused transformations.  A caller that gives them different streams -- prepare of batch i+1 ordered after
 * finish of batch i-1 (same slot), finish of batch i after its prepare -- overlaps the memory-bound preparation with the
 * latency-bound Legendre kernel (glass_b200/fields.py does).  GLB_ERR_UNSUPPORTED: not a group the INT8 path takes
 * (nmaps != 8, FP64-only mode, nside < 1024 in auto mode, no memory): use glb_alm2map. */
int glb_alm2map_prepare(glb_plan* plan, const double* d_alm, int nmaps, int slot, void* stream);
int glb_alm2map_finish(glb_plan* plan, int nmaps, int slot, double* d_map, const int* h_transform, const double* h_tparams,
                       void* stream);
/* Give back the buffers a plan allocates on first use and can rebuild on the next (the tile blocks of the INT8 Legendre
 * path, 4 GB at nside 4096): for callers that are done generating and need the memory for maps.  Synchronises the device. */
int glb_plan_release_scratch(glb_plan* plan);

/* ---- debug / test taps (stable, used by tests/ only) --------------------------------- */
/* Legendre stage only: alm -> phase array F_m(ring), [nmaps][nring][lmax+1] complex128 */
int glb_debug_alm2phase(glb_plan* plan, const double* d_alm, int nmaps, double* d_phase,
                        void* stream);
/* the same stage with the contraction over l on the INT8 tensor cores (tcgen05.mma kind::i8, exact
 * integer digit products; csrc/sht_ozaki.cu), nmaps = 4 or 8 maps on one recurrence */
int glb_debug_alm2phase_int8(glb_plan* plan, const double* d_alm, int nmaps, double* d_phase, void* stream);
/* ring-FFT stage only: phases -> map */
int glb_debug_phase2map(glb_plan* plan, const double* d_phase, int nmaps, double* d_map,
                        void* stream);
/* mlim[ring pair] table the two stages share (host copy), npair = 2*nside entries */
int glb_debug_mlim(const glb_plan* plan, int* h_mlim);

#ifdef __cplusplus
}
#endif
#endif /* GLASS_B200_H */
