// Host check of glass_b200/csrc/points_cuts.cuh: the closed-form batch cuts against the reference's
// loop (glass/points.py:409-437: 1000-pixel stepping, searchsorted(side="right"), "first pixel
// alone" rule), restated here step by step.  Built and run by tests/test_cpu_host.py.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <vector>

#include "../../glass_b200/csrc/points_cuts.cuh"

struct Cut { int64_t start, stop, n; };

static std::vector<Cut> reference_loop(const std::vector<int64_t>& cnt, int64_t batch) {
  const int64_t npix = (int64_t)cnt.size();
  int64_t count = 0;
  for (int64_t c : cnt) count += c;
  std::vector<Cut> cuts;
  const int64_t step = 1000;
  int64_t start = 0, stop = 0, size = 0;
  while (count) {
    const int64_t hi = std::min(npix, stop + step);
    std::vector<int64_t> q;
    int64_t acc = 0;
    for (int64_t i = stop; i < hi; ++i) q.push_back(acc += cnt[i]);
    if (size + q.back() < std::min(batch, count)) {
      stop += step;
      size += q.back();
    } else {
      stop += std::upper_bound(q.begin(), q.end(), batch - size) - q.begin();  // searchsorted(side="right")
      if (stop == start) stop += 1;
      int64_t tot = 0;
      for (int64_t i = start; i < stop; ++i) tot += cnt[i];
      cuts.push_back({start, stop, tot});
      start = stop;
      size = 0;
      count -= tot;
    }
  }
  return cuts;
}

static uint64_t s = 88172645463325252ULL;
static double rnd() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (double)(s >> 11) / 9007199254740992.0; }

static int check(const std::vector<int64_t>& cnt, int64_t batch, int chunk) {
  const int64_t npix = (int64_t)cnt.size();
  std::vector<int64_t> off(npix + 1, 0);
  for (int64_t i = 0; i < npix; ++i) off[i + 1] = off[i] + cnt[i];
  const std::vector<Cut> ref = reference_loop(cnt, batch);
  std::vector<Cut> got;
  int64_t start = 0, remaining = off[npix];
  std::vector<int64_t> buf(3 * chunk), state(3);
  while (remaining > 0) {  // the way the host drives the kernel: chunks of at most `chunk` cuts
    glb::cuts_chain(off.data(), npix, batch, start, remaining, chunk, buf.data(), state.data());
    for (int k = 0; k < state[0]; ++k) got.push_back({buf[3 * k], buf[3 * k + 1], buf[3 * k + 2]});
    start = state[1];
    remaining = state[2];
    if (state[0] == 0) return 2;
  }
  if (got.size() != ref.size()) return 1;
  for (size_t i = 0; i < ref.size(); ++i)
    if (got[i].start != ref[i].start || got[i].stop != ref[i].stop || got[i].n != ref[i].n) return 1;
  // the same rule walked on the galaxy list (gpix[g] = pixel of galaxy g), what K6 emits for sparse maps
  std::vector<int64_t> gpix;
  for (int64_t i = 0; i < npix; ++i)
    for (int64_t k = 0; k < cnt[i]; ++k) gpix.push_back(i);
  const int64_t total = (int64_t)gpix.size();
  got.clear();
  start = 0;
  remaining = total;
  while (remaining > 0) {
    glb::cuts_list_chain(gpix.data(), total, npix, batch, start, remaining, chunk, buf.data(), state.data());
    for (int k = 0; k < state[0]; ++k) got.push_back({buf[3 * k], buf[3 * k + 1], buf[3 * k + 2]});
    start = state[1];
    remaining = state[2];
    if (state[0] == 0) return 4;
  }
  if (got.size() != ref.size()) return 3;
  for (size_t i = 0; i < ref.size(); ++i)
    if (got[i].start != ref[i].start || got[i].stop != ref[i].stop || got[i].n != ref[i].n) return 3;
  return 0;
}

int main() {
  int ncases = 0;
  // hand-made corner cases: exact fit ending at a 1000-pixel group boundary, empty batches before an
  // oversize pixel, trailing zeros, a single pixel, everything in the last pixel
  {
    std::vector<int64_t> c(3072, 0);
    c[0] = 5; c[1500] = 9; c[1501] = 1; c[2999] = 2;
    if (check(c, 5, 4)) { std::printf("case A\n"); return 1; }
    std::vector<int64_t> d(3072, 0);
    d[999] = 3; d[1000] = 2; d[2500] = 4;
    if (check(d, 3, 1)) { std::printf("case B\n"); return 1; }
    std::vector<int64_t> e(1, 7);
    if (check(e, 3, 8) || check(e, 7, 8) || check(e, 100, 8)) { std::printf("case C\n"); return 1; }
    std::vector<int64_t> f(5000, 0);
    f[4999] = 11;
    if (check(f, 4, 8) || check(f, 11, 8) || check(f, 12, 8)) { std::printf("case D\n"); return 1; }
    ncases += 8;
  }
  // random maps: sparse and dense Poisson-like counts, batches from 1 to beyond the total, exact fits
  for (int it = 0; it < 3000; ++it) {
    const int64_t npix = 1 + (int64_t)(rnd() * 6000);
    const double density = std::pow(10.0, -3.0 + 4.0 * rnd());
    std::vector<int64_t> c(npix);
    int64_t total = 0;
    for (auto& x : c) {
      const double u = rnd();
      x = u < std::exp(-density) ? 0 : (int64_t)(1 + density * 2 * rnd() + (rnd() < 0.01 ? 50 * rnd() : 0));
      total += x;
    }
    if (total == 0) continue;
    int64_t batch = 1 + (int64_t)(rnd() * rnd() * 2.0 * (double)total);
    if (it % 7 == 0) {  // force exact fits: a batch equal to a partial sum
      int64_t acc = 0;
      const int64_t upto = (int64_t)(rnd() * npix);
      for (int64_t i = 0; i <= upto; ++i) acc += c[i];
      if (acc > 0) batch = acc;
    }
    const int rc = check(c, batch, 1 + (int)(rnd() * 64));
    if (rc) { std::printf("random case %d (npix %lld batch %lld) rc %d\n", it, (long long)npix, (long long)batch, rc); return 1; }
    ++ncases;
  }
  std::printf("points_cuts ok (%d cases)\n", ncases);
  return 0;
}
