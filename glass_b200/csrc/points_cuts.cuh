// points_cuts.cuh -- the batch-cut rule of glass/points.py:409-437 in closed form, on the exclusive
// scan of the galaxy counts.  Host/device: the kernel (points.cu) walks the chain of cuts on the
// device so that the host reads ONE small array per population instead of synchronising four or
// five times per batch; tests/native/points_cuts_host.cpp runs the same functions on the CPU
// against the reference's 1000-pixel stepping loop.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define GLB_CUTS_HD __host__ __device__ __forceinline__
#else
#define GLB_CUTS_HD inline
#endif

namespace glb {

// first index i in [0, n] with off[i] > target (strict = true) or off[i] >= target (strict = false);
// n + 1 if there is none.  off[0..n] is non-decreasing.
GLB_CUTS_HD int64_t cuts_search(const int64_t* off, int64_t n, int64_t target, bool strict) {
  int64_t lo = 0, hi = n + 1;
  while (lo < hi) {
    const int64_t mid = lo + ((hi - lo) >> 1);
    const int64_t v = off[mid];
    if (strict ? (v <= target) : (v < target))
      lo = mid + 1;
    else
      hi = mid;
  }
  return lo;
}

// The batch that starts at pixel `start` with `remaining` galaxies still to hand out.  The reference
// advances in groups of 1000 pixels (counted from `start`) until the group in which the running
// total reaches min(batch, remaining) -- at pixel q* -- and cuts inside that group with
// searchsorted(side="right"): after the last pixel whose running total is still <= batch, but not
// beyond the end of the group.  Hence stop = min(p, end of the group of q*), p the largest index with
// off[p] <= off[start] + batch.  A batch may be EMPTY (zero-count pixels before a pixel that alone
// exceeds `batch`); a first pixel that alone exceeds `batch` is taken by itself; on an exact fit, and
// for the last batch, trailing empty pixels are included up to the end of the group.
// Returns stop; *n = galaxies in [start, stop).
GLB_CUTS_HD int64_t cuts_next(const int64_t* off, int64_t npix, int64_t batch, int64_t start, int64_t remaining,
                              int64_t* n) {
  const int64_t base = off[start];
  int64_t p = cuts_search(off, npix, base + batch, true) - 1;  // largest p with off[p] <= base + batch
  if (p > npix) p = npix;
  int64_t stop;
  if (p <= start) {
    stop = start + 1;
  } else {
    const int64_t need = batch < remaining ? batch : remaining;
    const int64_t qstar = cuts_search(off, npix, base + need, false) - 1;  // pixel completing `need`
    const int64_t group_end = start + 1000 * ((qstar - start) / 1000 + 1);
    stop = group_end < p ? group_end : p;
  }
  *n = off[stop] - base;
  return stop;
}

// Up to max_cuts cuts from `start` with `remaining` galaxies still to hand out: cuts[i] = {start,
// stop, n}; state = {number of cuts written, next start, galaxies remaining}.
GLB_CUTS_HD void cuts_chain(const int64_t* off, int64_t npix, int64_t batch, int64_t start, int64_t remaining,
                            int max_cuts, int64_t* cuts, int64_t* state) {
  int k = 0;
  while (remaining > 0 && k < max_cuts) {
    int64_t n;
    const int64_t stop = cuts_next(off, npix, batch, start, remaining, &n);
    cuts[3 * k + 0] = start;
    cuts[3 * k + 1] = stop;
    cuts[3 * k + 2] = n;
    ++k;
    start = stop;
    remaining -= n;
  }
  state[0] = k;
  state[1] = start;
  state[2] = remaining;
}

// ---- the same rule on the GALAXY LIST ---------------------------------------------------------
// gpix[g] = ring pixel of galaxy g, non-decreasing, g in [0, total) -- what K6 emits instead of
// counts and offsets when the map is sparsely populated.  With off[p] = #{g : gpix[g] < p}:
//   largest p with off[p] <= T     = gpix[T]       (T < total; npix otherwise)
//   pixel completing V galaxies    = gpix[V - 1]
//   off[stop]                      = lower bound of `stop` in gpix
// so the cuts need neither the counts nor the offsets.  `base` = index of the first galaxy not yet
// handed out (= total - remaining).

// first g in [lo, hi) with gpix[g] >= pix; hi if none
GLB_CUTS_HD int64_t cuts_list_lower_bound(const int64_t* gpix, int64_t lo, int64_t hi, int64_t pix) {
  while (lo < hi) {
    const int64_t mid = lo + ((hi - lo) >> 1);
    if (gpix[mid] < pix)
      lo = mid + 1;
    else
      hi = mid;
  }
  return lo;
}

GLB_CUTS_HD int64_t cuts_list_next(const int64_t* gpix, int64_t total, int64_t npix, int64_t batch, int64_t start,
                                   int64_t remaining, int64_t* n) {
  const int64_t base = total - remaining;
  // base + batch may overflow for absurd batch sizes: compare through the difference
  const int64_t p = (batch < total - base) ? gpix[base + batch] : npix;
  int64_t stop;
  if (p <= start) {
    stop = start + 1;
  } else {
    const int64_t need = batch < remaining ? batch : remaining;
    const int64_t qstar = gpix[base + need - 1];
    const int64_t group_end = start + 1000 * ((qstar - start) / 1000 + 1);
    stop = group_end < p ? group_end : p;
  }
  *n = cuts_list_lower_bound(gpix, base, total, stop) - base;
  return stop;
}

GLB_CUTS_HD void cuts_list_chain(const int64_t* gpix, int64_t total, int64_t npix, int64_t batch, int64_t start,
                                 int64_t remaining, int max_cuts, int64_t* cuts, int64_t* state) {
  int k = 0;
  while (remaining > 0 && k < max_cuts) {
    int64_t n;
    const int64_t stop = cuts_list_next(gpix, total, npix, batch, start, remaining, &n);
    cuts[3 * k + 0] = start;
    cuts[3 * k + 1] = stop;
    cuts[3 * k + 2] = n;
    ++k;
    start = stop;
    remaining -= n;
  }
  state[0] = k;
  state[1] = start;
  state[2] = remaining;
}

}  // namespace glb
