// Host check of glass_b200/csrc/fft_core.cuh: every pass function is executed "thread by
// thread" and compared with a naive long-double DFT.  Built and run by tests/test_cpu_host.py.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../glass_b200/csrc/fft_core.cuh"

using namespace glb;
using namespace glb::fft;
typedef long double ld;

// textbook recursive radix-2 FFT in long double (reference for the long transforms)
static void ref_fft_rec(std::vector<ld>& re, std::vector<ld>& im, bool inverse) {
  const size_t M = re.size();
  if (M == 1) return;
  std::vector<ld> er(M / 2), ei(M / 2), orr(M / 2), oi(M / 2);
  for (size_t i = 0; i < M / 2; ++i) { er[i] = re[2 * i]; ei[i] = im[2 * i]; orr[i] = re[2 * i + 1]; oi[i] = im[2 * i + 1]; }
  ref_fft_rec(er, ei, inverse);
  ref_fft_rec(orr, oi, inverse);
  const ld twopi = 6.283185307179586476925286766559005768L;
  for (size_t k = 0; k < M / 2; ++k) {
    const ld a = twopi * (ld)k / (ld)M * (inverse ? 1 : -1);
    const ld c = cosl(a), s = sinl(a);
    const ld tr = orr[k] * c - oi[k] * s, ti = orr[k] * s + oi[k] * c;
    re[k] = er[k] + tr; im[k] = ei[k] + ti;
    re[k + M / 2] = er[k] - tr; im[k + M / 2] = ei[k] - ti;
  }
}
static std::vector<double2> naive_dft(const std::vector<double2>& x, bool inverse) {
  const int M = (int)x.size();
  if (M > 256) {
    std::vector<ld> re(M), im(M);
    for (int i = 0; i < M; ++i) { re[i] = x[i].x; im[i] = x[i].y; }
    ref_fft_rec(re, im, inverse);
    std::vector<double2> y(M);
    for (int i = 0; i < M; ++i) y[i] = make_double2((double)re[i], (double)im[i]);
    return y;
  }
  std::vector<double2> y(M);
  const ld twopi = 6.283185307179586476925286766559005768L;
  for (int k = 0; k < M; ++k) {
    ld sr = 0, si = 0;
    for (int j = 0; j < M; ++j) {
      const ld a = twopi * (ld)(((long long)j * k) % M) / (ld)M * (inverse ? 1 : -1);
      const ld c = cosl(a), s = sinl(a);
      sr += x[j].x * c - x[j].y * s;
      si += x[j].x * s + x[j].y * c;
    }
    y[k] = make_double2((double)sr, (double)si);
  }
  return y;
}

static double maxdiff(const std::vector<double2>& a, const std::vector<double2>& b) {
  double d = 0, s = 0;
  for (size_t i = 0; i < a.size(); ++i) {
    d = std::fmax(d, std::fmax(std::fabs(a[i].x - b[i].x), std::fabs(a[i].y - b[i].y)));
    s = std::fmax(s, std::fmax(std::fabs(b[i].x), std::fabs(b[i].y)));
  }
  return d / (s > 0 ? s : 1);
}

struct Ctx {
  int n, M, T, nthreads, tw_n;
  std::vector<double2> tw;
};

static void run_dif_upper(std::vector<double2>& buf, const Ctx& c, bool inverse, int nvalid) {
  int nv = nvalid;
  dif_upper_schedule(c.n, [&](int K, int s) {
    for (int tid = 0; tid < c.nthreads; ++tid) {
      if (K == 3) dif_pass<3>(buf.data(), c.n, s, c.tw.data(), c.tw_n, inverse, nv, tid, c.nthreads);
      if (K == 2) dif_pass<2>(buf.data(), c.n, s, c.tw.data(), c.tw_n, inverse, nv, tid, c.nthreads);
      if (K == 1) dif_pass<1>(buf.data(), c.n, s, c.tw.data(), c.tw_n, inverse, nv, tid, c.nthreads);
    }
    nv = c.M;
  });
}
static void run_dit_upper(std::vector<double2>& buf, const Ctx& c, bool inverse) {
  dit_upper_schedule(c.n, [&](int K, int s) {
    for (int tid = 0; tid < c.nthreads; ++tid) {
      if (K == 3) dit_pass<3>(buf.data(), c.n, s, c.tw.data(), c.tw_n, inverse, tid, c.nthreads);
      if (K == 2) dit_pass<2>(buf.data(), c.n, s, c.tw.data(), c.tw_n, inverse, tid, c.nthreads);
      if (K == 1) dit_pass<1>(buf.data(), c.n, s, c.tw.data(), c.tw_n, inverse, tid, c.nthreads);
    }
  });
}
#define BY_T(T_, CALL4, CALL3, CALL2, CALL1) \
  switch (T_) { case 4: CALL4; break; case 3: CALL3; break; case 2: CALL2; break; case 1: CALL1; break; default: break; }

int main() {
  int fails = 0;
  srand(1234);
  for (int n = 0; n <= 13; ++n) {
    Ctx c;
    c.n = n;
    c.M = 1 << n;
    c.T = tail_levels(n);
    c.nthreads = n >= 9 ? 64 : 8;
    c.tw_n = 8192;
    c.tw.resize(c.tw_n / 2);
    const ld twopi = 6.283185307179586476925286766559005768L;
    for (int t = 0; t < c.tw_n / 2; ++t) c.tw[t] = make_double2((double)cosl(twopi * t / c.tw_n), (double)-sinl(twopi * t / c.tw_n));
    const int M = c.M;
    std::vector<double2> x(M);
    for (auto& v : x) v = make_double2(rand() / (double)RAND_MAX - 0.5, rand() / (double)RAND_MAX - 0.5);
    const int cap = M + 8;  // swizzle stays inside aligned blocks of 8
    for (int inv = 0; inv < 2; ++inv) {
      const bool inverse = inv;
      const std::vector<double2> ref = naive_dft(x, inverse);
      // (1) DIF in place -> bit-reversed
      std::vector<double2> buf(cap, make_double2(1e300, 1e300));
      for (int i = 0; i < M; ++i) buf[sw(i, n)] = x[i];
      run_dif_upper(buf, c, inverse, M);
      for (int tid = 0; tid < c.nthreads; ++tid)
        BY_T(c.T, dif_tail_inplace<4>(buf.data(), n, inverse, M, tid, c.nthreads), dif_tail_inplace<3>(buf.data(), n, inverse, M, tid, c.nthreads),
             dif_tail_inplace<2>(buf.data(), n, inverse, M, tid, c.nthreads), dif_tail_inplace<1>(buf.data(), n, inverse, M, tid, c.nthreads))
      std::vector<double2> got(M);
      for (int p = 0; p < M; ++p) got[bitrev((unsigned)p, n)] = buf[sw(p, n)];
      std::vector<double2> dif_nat = got;
      if (!ref.empty()) {
        const double e = maxdiff(got, ref);
        if (!(e < 1e-13)) { printf("FAIL dif n=%d inv=%d err=%g\n", n, inv, e); ++fails; }
      }
      // (2) DIF with reordering tail -> natural order through emit
      std::vector<double2> buf2(cap, make_double2(1e300, 1e300));
      for (int i = 0; i < M; ++i) buf2[sw(i, n)] = x[i];
      run_dif_upper(buf2, c, inverse, M);
      std::vector<double2> got2(M, make_double2(1e300, 0));
      auto emit = [&](int j, double2 v) { got2[j] = v; };
      if (c.T == 0) emit(0, buf2[0]);
      for (int tid = 0; tid < c.nthreads; ++tid)
        BY_T(c.T, dif_tail_reorder<4>(buf2.data(), n, inverse, tid, c.nthreads, emit), dif_tail_reorder<3>(buf2.data(), n, inverse, tid, c.nthreads, emit),
             dif_tail_reorder<2>(buf2.data(), n, inverse, tid, c.nthreads, emit), dif_tail_reorder<1>(buf2.data(), n, inverse, tid, c.nthreads, emit))
      {
        const double e = maxdiff(got2, dif_nat);
        if (!(e == 0.0)) { printf("FAIL dif-reorder n=%d inv=%d err=%g\n", n, inv, e); ++fails; }
      }
      // (3) DIT: bit-reversed in -> natural out
      std::vector<double2> buf3(cap, make_double2(1e300, 1e300));
      for (int p = 0; p < M; ++p) buf3[sw(p, n)] = x[bitrev((unsigned)p, n)];
      for (int tid = 0; tid < c.nthreads; ++tid)
        BY_T(c.T, dit_head_inplace<4>(buf3.data(), n, inverse, tid, c.nthreads), dit_head_inplace<3>(buf3.data(), n, inverse, tid, c.nthreads),
             dit_head_inplace<2>(buf3.data(), n, inverse, tid, c.nthreads), dit_head_inplace<1>(buf3.data(), n, inverse, tid, c.nthreads))
      run_dit_upper(buf3, c, inverse);
      std::vector<double2> got3(M);
      for (int i = 0; i < M; ++i) got3[i] = buf3[sw(i, n)];
      {
        const double e = maxdiff(got3, dif_nat);
        if (!(e < 1e-13)) { printf("FAIL dit n=%d inv=%d err=%g\n", n, inv, e); ++fails; }
      }
    }
    // (4) Bluestein convolution: y = IDFT(DFT(x zero-padded beyond nvalid) .* B) with B = DFT(b)/M
    if (n >= 1) {
      const int nvalid = M / 2 + 1 > M ? M : (M / 2 + 1);
      std::vector<double2> b(M);
      for (auto& v : b) v = make_double2(rand() / (double)RAND_MAX - 0.5, rand() / (double)RAND_MAX - 0.5);
      // spectrum of b by the DIF path (bit-reversed), scaled, block-transposed
      std::vector<double2> bb(cap);
      for (int i = 0; i < M; ++i) bb[sw(i, n)] = b[i];
      run_dif_upper(bb, c, false, M);
      for (int tid = 0; tid < c.nthreads; ++tid)
        BY_T(c.T, dif_tail_inplace<4>(bb.data(), n, false, M, tid, c.nthreads), dif_tail_inplace<3>(bb.data(), n, false, M, tid, c.nthreads),
             dif_tail_inplace<2>(bb.data(), n, false, M, tid, c.nthreads), dif_tail_inplace<1>(bb.data(), n, false, M, tid, c.nthreads))
      std::vector<double2> bf(M);
      for (int p = 0; p < M; ++p) {
        const double2 v = bb[sw(p, n)];
        bf[bf_index(p, n)] = make_double2(v.x / M, v.y / M);
      }
      std::vector<double2> buf(cap, make_double2(1e300, 1e300));  // garbage beyond nvalid must be ignored
      for (int i = 0; i < nvalid; ++i) buf[sw(i, n)] = x[i];
      run_dif_upper(buf, c, false, nvalid);
      const int nv = (n - c.T) > 0 ? M : nvalid;
      for (int tid = 0; tid < c.nthreads; ++tid)
        BY_T(c.T, bluestein_middle<4>(buf.data(), n, bf.data(), nv, tid, c.nthreads), bluestein_middle<3>(buf.data(), n, bf.data(), nv, tid, c.nthreads),
             bluestein_middle<2>(buf.data(), n, bf.data(), nv, tid, c.nthreads), bluestein_middle<1>(buf.data(), n, bf.data(), nv, tid, c.nthreads))
      run_dit_upper(buf, c, true);
      if (n <= 11) {
        std::vector<double2> ref(M);
        for (int k = 0; k < M; ++k) {
          ld sr = 0, si = 0;
          for (int j = 0; j < nvalid; ++j) {
            const double2 bv = b[(k - j + M) % M];
            sr += (ld)x[j].x * bv.x - (ld)x[j].y * bv.y;
            si += (ld)x[j].x * bv.y + (ld)x[j].y * bv.x;
          }
          ref[k] = make_double2((double)sr, (double)si);
        }
        std::vector<double2> got(M);
        for (int i = 0; i < M; ++i) got[i] = buf[sw(i, n)];
        const double e = maxdiff(got, ref);
        if (!(e < 1e-12)) { printf("FAIL bluestein n=%d err=%g\n", n, e); ++fails; }
      }
    }
  }
  // swizzle is a bijection within aligned blocks of 8
  for (int n = 0; n <= 13; ++n)
    for (int i = 0; i < (1 << n); ++i)
      if ((sw(i, n) >> 3) != (i >> 3) || sw(sw(i, n), n) != i) { printf("FAIL swizzle n=%d i=%d\n", n, i); ++fails; n = 99; break; }
  printf(fails ? "FAILED %d\n" : "fft_core ok\n", fails);
  return fails ? 1 : 0;
}
