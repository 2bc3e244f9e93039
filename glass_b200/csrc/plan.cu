// plan.cu -- plan construction: HEALPix ring geometry (RING scheme), the mlim table shared
// by the Legendre and FFT stages, the Legendre work list, FFT twiddles and Bluestein chirp
// spectra, and the workspace.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <map>

#include "plan.h"

namespace glb {

static thread_local std::string g_last_error;
void set_last_error(const std::string& s) { g_last_error = s; }
const std::string& last_error() { return g_last_error; }
static unsigned long long g_launches = 0;
void count_launch(int n) { __atomic_fetch_add(&g_launches, (unsigned long long)n, __ATOMIC_RELAXED); }
unsigned long long launch_count() { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

int ringfft_class_of(int lbuf);
size_t ringfft_long_scratch_bytes(int device);
int sht_build_prep_tables(glb_plan* pl, cudaStream_t st);
int ringfft_build_spectra(glb_plan* pl, const std::vector<int>& Ls, const std::vector<int>& Ms,
                          const std::vector<int64_t>& offs, cudaStream_t st);

static inline bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }
static inline int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

// libsharp-style cut-off: beyond mlim the lambda_lm(theta) are negligible for all l <= lmax
static int compute_mlim(int lmax, int spin, double sth, double cth) {
  double ofs = lmax * 0.01;
  if (ofs < 100.) ofs = 100.;
  const double b = -2.0 * spin * std::fabs(cth);
  const double t1 = lmax * sth + ofs;
  const double c = (double)spin * spin - t1 * t1;
  const double discr = b * b - 4.0 * c;
  if (discr <= 0) return lmax;
  double res = (-b + std::sqrt(discr)) / 2.0;
  if (res > lmax) res = lmax;
  return (int)(res + 0.5);
}

template <typename T>
static int upload(T** dptr, const std::vector<T>& h) {
  GLB_CUDA_CHECK(cudaMalloc((void**)dptr, std::max<size_t>(h.size(), 1) * sizeof(T)));
  if (!h.empty()) GLB_CUDA_CHECK(cudaMemcpy(*dptr, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  return GLB_OK;
}

int plan_build(glb_plan* pl) {
  const int N = pl->nside, lmax = pl->lmax;
  pl->mmax = lmax;
  pl->nring = 4 * N - 1;
  pl->npair = 2 * N;
  pl->npix = 12LL * N * N;
  pl->nalm = (int64_t)(lmax + 1) * (lmax + 2) / 2;
  GLB_CUDA_CHECK(cudaSetDevice(pl->device));

  // ---- ring geometry ----
  pl->h_z.resize(pl->npair);
  pl->h_sth.resize(pl->npair);
  pl->h_mlim.resize(pl->npair);
  for (int r = 0; r < pl->npair; ++r) {
    const int i = r + 1;  // ring number 1..2N
    double z, sth;
    if (i < N) {
      const double t = (double)i * i / (3.0 * N * N);
      z = 1.0 - t;
      sth = std::sqrt(t * (2.0 - t));
    } else {
      z = (2.0 * N - i) * 2.0 / (3.0 * N);
      sth = std::sqrt((1.0 - z) * (1.0 + z));
    }
    pl->h_z[r] = z;
    pl->h_sth[r] = sth;
    pl->h_mlim[r] = std::min(compute_mlim(lmax, 0, sth, z), pl->mmax);
  }
  // mlim must be monotone towards the equator for the tile skip logic
  for (int r = 1; r < pl->npair; ++r) pl->h_mlim[r] = std::max(pl->h_mlim[r], pl->h_mlim[r - 1]);
  pl->h_rmin.assign(pl->mmax + 1, pl->npair);
  {
    int r = 0;
    for (int m = 0; m <= pl->mmax; ++m) {
      while (r < pl->npair && pl->h_mlim[r] < m) ++r;
      pl->h_rmin[m] = r;
    }
  }
  int rc;
  if ((rc = upload(&pl->d_z, pl->h_z)) != GLB_OK) return rc;
  if ((rc = upload(&pl->d_sth, pl->h_sth)) != GLB_OK) return rc;
  if ((rc = upload(&pl->d_mlim, pl->h_mlim)) != GLB_OK) return rc;

  // ---- lambda_mm prefactor c_m = sqrt((2m+1)!!/(4 pi (2m)!!)) as mantissa * 2^exp ----
  {
    std::vector<double> mant(pl->mmax + 1);
    std::vector<int> ex(pl->mmax + 1);
    long double c = sqrtl(1.0L / (4.0L * 3.14159265358979323846264338327950288L));
    int e = 0;
    for (int m = 0; m <= pl->mmax; ++m) {
      if (m > 0) c *= sqrtl((2.0L * m + 1.0L) / (2.0L * m));
      int de;
      c = frexpl(c, &de);
      e += de;
      mant[m] = (double)c;
      ex[m] = e;
    }
    if ((rc = upload(&pl->d_cm_mant, mant)) != GLB_OK) return rc;
    if ((rc = upload(&pl->d_cm_exp, ex)) != GLB_OK) return rc;
  }

  // ---- record offsets ----
  {
    std::vector<int64_t> roff(pl->mmax + 2);
    int64_t acc = 0;
    for (int m = 0; m <= pl->mmax; ++m) {
      roff[m] = acc;
      acc += (lmax - m) / 2 + 1;
    }
    roff[pl->mmax + 1] = acc;
    pl->nrec = acc;
    if ((rc = upload(&pl->d_roff, roff)) != GLB_OK) return rc;
  }

  if ((rc = sht_build_prep_tables(pl, 0)) != GLB_OK) return rc;

  // ---- Legendre work list, most expensive first ----
  {
    pl->leg_R = 4;
    pl->leg_threads = (pl->npair >= 1024) ? 256 : (pl->npair >= 512 ? 128 : 64);
    if (const char* env = getenv("GLB_LEG_TILE")) {  // tuning knob: ring pairs per CTA tile (256/512/1024)
      const int t = atoi(env);
      if (t == 256 || t == 512 || t == 1024) pl->leg_threads = t / 4;
    }
    const int T = pl->leg_threads * pl->leg_R;
    const int ntile = (pl->npair + T - 1) / T;
    struct Tmp {
      LegItem it;
      double cost;
    };
    std::vector<Tmp> tmp;
    for (int m = 0; m <= pl->mmax; ++m) {
      const int K = (lmax - m) / 2 + 1;
      for (int t = 0; t < ntile; ++t) {
        const int lo = std::max(t * T, pl->h_rmin[m]);
        const int hi = std::min((t + 1) * T, pl->npair);
        if (hi <= lo) continue;
        tmp.push_back({{m, t}, (double)K * (hi - lo)});
      }
    }
    std::stable_sort(tmp.begin(), tmp.end(), [](const Tmp& a, const Tmp& b) { return a.cost > b.cost; });
    std::vector<LegItem> items(tmp.size());
    for (size_t i = 0; i < tmp.size(); ++i) items[i] = tmp[i].it;
    pl->nitems = (int)items.size();
    if ((rc = upload(&pl->d_items, items)) != GLB_OK) return rc;
  }

  // ---- ring descriptors ----
  pl->h_rings.resize(pl->nring);
  std::map<int, int64_t> bf_off_of_L;  // distinct Bluestein lengths
  std::vector<int> Ls, Ms;
  std::vector<int64_t> offs;
  int64_t bf_total = 0;
  int max_fft = 2;
  for (int r = 0; r < pl->nring; ++r) {
    RingDesc d;
    const int i = r + 1;
    const bool south = i > 3 * N;
    const bool north = i < N;
    const int ip = south ? 4 * N - i : i;
    if (north || south) {
      d.nphi = 4 * ip;
      d.shifted = 1;
      d.start = north ? 2LL * ip * (ip - 1) : pl->npix - 2LL * ip * (ip + 1);
    } else {
      d.nphi = 4 * N;
      d.shifted = ((i - N) % 2 == 0) ? 1 : 0;
      d.start = 2LL * N * (N - 1) + (int64_t)(i - N) * 4 * N;
    }
    d.pair = (r < pl->npair) ? r : pl->nring - 1 - r;
    const int h = d.nphi / 2;
    if (is_pow2(h)) {
      d.L = 0;
      d.M = h;
      d.bf_off = -1;
    } else {
      d.L = h / 2;
      d.M = next_pow2(2 * d.L - 1);
      auto it = bf_off_of_L.find(d.L);
      if (it == bf_off_of_L.end()) {
        bf_off_of_L[d.L] = bf_total;
        Ls.push_back(d.L);
        Ms.push_back(d.M);
        offs.push_back(bf_total);
        d.bf_off = bf_total;
        bf_total += d.M;
      } else {
        d.bf_off = it->second;
      }
    }
    max_fft = std::max(max_fft, d.M);
    if (ringfft_class_of(d.M) < 0) {
      set_last_error("nside too large for the ring FFT (FFT length " + std::to_string(d.M) + " > 16384, i.e. nside > 8192)");
      return GLB_ERR_UNSUPPORTED;
    }
    pl->h_rings[r] = d;
  }
  if ((rc = upload(&pl->d_rings, pl->h_rings)) != GLB_OK) return rc;
  {
    std::vector<int> order[4];
    // pairs adjacent (north ring, its southern mirror), largest first
    std::vector<int> pairs(pl->npair);
    for (int r = 0; r < pl->npair; ++r) pairs[r] = r;
    std::stable_sort(pairs.begin(), pairs.end(),
                     [&](int a, int b) { return pl->h_rings[a].M > pl->h_rings[b].M; });
    for (int r : pairs) {
      const int c = ringfft_class_of(pl->h_rings[r].M);
      order[c].push_back(r);
      if (r != pl->npair - 1) order[c].push_back(pl->nring - 1 - r);
    }
    for (int c = 0; c < 4; ++c) {
      pl->n_ring_class[c] = (int)order[c].size();
      if ((rc = upload(&pl->d_ring_order[c], order[c])) != GLB_OK) return rc;
    }
    if (pl->n_ring_class[3] > 0) GLB_CUDA_CHECK(cudaMalloc((void**)&pl->d_long_scratch, ringfft_long_scratch_bytes(pl->device)));
  }
  // ---- twiddles e^{-2 pi i t / tw_n}, t < tw_n/2 (long double on host) ----
  {
    pl->tw_n = max_fft;
    std::vector<double2> tw(std::max(pl->tw_n / 2, 1));
    const long double twopi = 6.283185307179586476925286766559005768L;
    for (int t = 0; t < pl->tw_n / 2; ++t) {
      const long double a = twopi * (long double)t / (long double)pl->tw_n;
      tw[t] = make_double2((double)cosl(a), (double)-sinl(a));
    }
    if ((rc = upload(&pl->d_tw, tw)) != GLB_OK) return rc;
  }
  // ---- Bluestein chirp spectra ----
  pl->bf_total = bf_total;
  GLB_CUDA_CHECK(cudaMalloc((void**)&pl->d_bf, std::max<int64_t>(bf_total, 1) * sizeof(double2)));
  if ((rc = ringfft_build_spectra(pl, Ls, Ms, offs, 0)) != GLB_OK) return rc;

  // ---- workspace ----
  const int gb = std::min(pl->max_batch, 4);
  const int gmax = gb >= 4 ? 4 : (gb >= 2 ? 2 : 1);
  // records: scalar synthesis needs nrec*(4+4B) doubles, the spin transform 8 + 2 per coefficient
  // set and (l,m): 12 for E and B of one map, 16 for the E modes of four maps
  pl->rec_capacity = std::max<int64_t>(pl->nrec * (4 + 4 * gmax), pl->nalm * (gmax >= 4 ? 16 : 12));
  const size_t rec_bytes = (size_t)pl->rec_capacity * sizeof(double);
  const size_t phase_bytes = (size_t)std::max(gmax, 2) * pl->nring * (pl->mmax + 1) * sizeof(double2);
  GLB_CUDA_CHECK(cudaMalloc((void**)&pl->d_rec, rec_bytes));
  GLB_CUDA_CHECK(cudaMalloc((void**)&pl->d_phase, phase_bytes));
  pl->workspace_bytes = (int64_t)(rec_bytes + phase_bytes + bf_total * sizeof(double2) + (size_t)pl->nrec * 6 * sizeof(double));
  return GLB_OK;
}

// scalar (m, ring tile) work list for an arbitrary tile size, most expensive first
int plan_items(glb_plan* pl, int tile, int G, int rank, LegItem** d_items, int* nitems) {
  const int64_t key = (int64_t)tile | ((int64_t)G << 24) | ((int64_t)rank << 36);
  auto it = pl->item_lists.find(key);
  if (it == pl->item_lists.end()) {
    const int ntile = (pl->npair + tile - 1) / tile;
    struct Tmp {
      LegItem it;
      double cost;
    };
    std::vector<Tmp> tmp;
    for (int m = 0; m <= pl->mmax; ++m) {
      if (m % G != rank) continue;  // m-split: this rank's share of the m values
      const int K = (pl->lmax - m) / 2 + 1;
      for (int t = 0; t < ntile; ++t) {
        const int lo = std::max(t * tile, pl->h_rmin[m]);
        const int hi = std::min((t + 1) * tile, pl->npair);
        if (hi <= lo) continue;
        tmp.push_back({{m, t}, (double)K * (hi - lo)});
      }
    }
    std::stable_sort(tmp.begin(), tmp.end(), [](const Tmp& a, const Tmp& b) { return a.cost > b.cost; });
    std::vector<LegItem> items(tmp.size());
    for (size_t i = 0; i < tmp.size(); ++i) items[i] = tmp[i].it;
    LegItem* d = nullptr;
    const int rc = upload(&d, items);
    if (rc != GLB_OK) return rc;
    it = pl->item_lists.emplace(key, std::make_pair(d, (int)items.size())).first;
  }
  *d_items = it->second.first;
  *nitems = it->second.second;
  return GLB_OK;
}

// m-split set-up: rowmap[ring] = row in the permuted send layout, my_rings = rings this rank owns
int plan_dist_setup(glb_plan* pl, int world, int rank, const int* h_rowmap, const int* h_my_rings, int n_my) {
  GLB_CUDA_CHECK(cudaSetDevice(pl->device));
  pl->dist_world = world;
  pl->dist_rank = rank;
  pl->dist_W = (pl->mmax + 1 + world - 1) / world;
  pl->dist_rows_local = n_my;
  std::vector<int> rowmap(h_rowmap, h_rowmap + pl->nring), rowidx(pl->nring, -1);
  for (int i = 0; i < n_my; ++i) {
    if (h_my_rings[i] < 0 || h_my_rings[i] >= pl->nring) {
      set_last_error("glb_dist_setup: ring index out of range");
      return GLB_ERR_INVALID_ARG;
    }
    rowidx[h_my_rings[i]] = i;
  }
  cudaFree(pl->d_dist_rowmap);
  cudaFree(pl->d_dist_rowidx);
  pl->d_dist_rowmap = pl->d_dist_rowidx = nullptr;
  int rc;
  if ((rc = upload(&pl->d_dist_rowmap, rowmap)) != GLB_OK) return rc;
  if ((rc = upload(&pl->d_dist_rowidx, rowidx)) != GLB_OK) return rc;
  // ring lists per FFT size class, largest first (same order as the single-GPU lists)
  std::vector<int> mine(h_my_rings, h_my_rings + n_my);
  std::stable_sort(mine.begin(), mine.end(), [&](int a, int b) { return pl->h_rings[a].M > pl->h_rings[b].M; });
  std::vector<int> order[4];
  for (int r : mine) order[ringfft_class_of(pl->h_rings[r].M)].push_back(r);
  for (int c = 0; c < 4; ++c) {
    cudaFree(pl->d_dist_ring_order[c]);
    pl->d_dist_ring_order[c] = nullptr;
    pl->n_dist_ring_class[c] = (int)order[c].size();
    if ((rc = upload(&pl->d_dist_ring_order[c], order[c])) != GLB_OK) return rc;
  }
  return GLB_OK;
}

void plan_free(glb_plan* pl) {
  for (auto& kv : pl->item_lists) cudaFree(kv.second.first);
  pl->item_lists.clear();
  cudaSetDevice(pl->device);
  cudaFree(pl->d_z);
  cudaFree(pl->d_sth);
  cudaFree(pl->d_mlim);
  cudaFree(pl->d_cm_mant);
  cudaFree(pl->d_cm_exp);
  cudaFree(pl->d_roff);
  cudaFree(pl->d_prep_tab);
  cudaFree(pl->d_items);
  cudaFree(pl->d_rings);
  for (int c = 0; c < 4; ++c) cudaFree(pl->d_ring_order[c]);
  cudaFree(pl->d_long_scratch);
  cudaFree(pl->d_tw);
  cudaFree(pl->d_bf);
  cudaFree(pl->d_rec);
  cudaFree(pl->d_oz);
  cudaFree(pl->d_oz1);
  cudaFree(pl->d_oz2);
  cudaFree(pl->d_oz_toff);
  cudaFree(pl->d_phase);
  cudaFree(pl->d_ch);
  cudaFree(pl->d_sh);
  cudaFree(pl->d_mlim_spin);
  cudaFree(pl->d_sn_mant);
  cudaFree(pl->d_sn_exp);
  cudaFree(pl->d_soff);
  for (auto& kv : pl->spin_item_lists) cudaFree(kv.second.first);
  cudaFree(pl->d_phase_spin);
  cudaFree(pl->d_spin_tab);
  cudaFree(pl->d_dist_rowmap);
  cudaFree(pl->d_dist_rowidx);
  for (int c = 0; c < 4; ++c) cudaFree(pl->d_dist_ring_order[c]);
  // peer mappings first (the owners free the memory itself), then this rank's receive buffers
  for (int d = 0; d < P2P_MAX_WORLD; ++d)
    if (pl->p2p_peer[d]) cudaIpcCloseMemHandle(pl->p2p_peer[d]);
  cudaFree(pl->d_p2p_recv[0]);  // one block holds both buffers
  cudaFree(pl->d_p2p_tab);
  cudaFree(pl->d_partial);
  cudaFree(pl->d_ana_first_tile);
  cudaFree(pl->d_tmpmap);
  cudaFree(pl->d_ab_tab);
  if (pl->h_pin_in) cudaFreeHost(pl->h_pin_in);
  if (pl->h_pin_out) cudaFreeHost(pl->h_pin_out);
  cudaFree(pl->d_stage_alm);
  cudaFree(pl->d_stage_map);
  for (cudaEvent_t e : pl->ev_pool) cudaEventDestroy(e);
  pl->ev_pool.clear();
}

}  // namespace glb
