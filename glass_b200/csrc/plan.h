// plan.h -- the opaque glb_plan: HEALPix ring geometry, Legendre work list, FFT tables
// and workspace for one (nside, lmax, device).
#pragma once
#include <map>
#include <utility>
#include <vector>

#include "common.cuh"

namespace glb {

// One Legendre work item: a fixed m and a tile of ring pairs.
struct LegItem {
  int m;
  int tile;
};

// Per-ring descriptor for the ring-FFT stage.
struct RingDesc {
  int64_t start;   // first pixel
  int nphi;        // pixels on the ring (multiple of 4)
  int shifted;     // phi0 = pi/nphi if 1, else 0
  int pair;        // ring-pair index (0 = polar ... 2*nside-1 = equator)
  int L;           // Bluestein length (nphi/4) or 0 if nphi/2 is a power of two
  int M;           // power-of-two convolution length (Bluestein) or nphi/2 (direct)
  int64_t bf_off;  // offset (in double2) of the chirp spectrum in d_bf, -1 if direct
};

// m-split with peer stores: where the Legendre kernel of rank `rank` puts F_m(ring) -- straight
// into the receive buffer [map][src rank][local row][W] of the rank that owns the ring, through
// peer mappings of those buffers (NVLink).  One table per receive buffer, in device memory.
constexpr int P2P_MAX_WORLD = 8;
constexpr int P2P_NBUF = 2;
struct LegP2P {
  double2* base[P2P_MAX_WORLD];       // receive buffer of every rank as mapped into THIS process
  int rowstart[P2P_MAX_WORLD + 1];    // first row (permuted send order) owned by each rank
  int rank, world;
};

}  // namespace glb

struct glb_plan {
  int nside = 0, lmax = 0, mmax = 0, device = 0, max_batch = 1;
  int nring = 0, npair = 0;
  int64_t npix = 0, nalm = 0;

  // ---- host copies ----
  std::vector<double> h_z, h_sth;   // per ring pair (north ring / equator)
  std::vector<int> h_mlim;          // per ring pair
  std::vector<int> h_rmin;          // per m: first ring pair with mlim >= m
  std::vector<glb::RingDesc> h_rings;

  // ---- device tables ----
  double* d_z = nullptr;            // [npair]
  double* d_sth = nullptr;          // [npair]
  int* d_mlim = nullptr;            // [npair]
  double* d_cm_mant = nullptr;      // [mmax+1] mantissa of lambda_mm prefactor
  int* d_cm_exp = nullptr;          // [mmax+1] binary exponent
  int64_t* d_roff = nullptr;        // [mmax+2] record offsets (in l-pairs) per m
  int64_t nrec = 0;                 // total l-pairs over all m
  double* d_prep_tab = nullptr;     // [nrec][5] static recurrence / rescaling tables

  // Legendre work lists per (threads, R) configuration actually used
  glb::LegItem* d_items = nullptr;
  int nitems = 0;
  int leg_threads = 256, leg_R = 4;  // tile = leg_threads * leg_R ring pairs
  // extra scalar work lists keyed by tile size (ring pairs per CTA), built on demand
  std::map<int64_t, std::pair<glb::LegItem*, int>> item_lists;

  // spin-weighted synthesis (built lazily by plan_ensure_spin for one spin at a time)
  int spin_ready = 0;                // spin the tables below were built for (0 = none)
  double* d_ch = nullptr;            // [npair] cos(theta/2)
  double* d_sh = nullptr;            // [npair] sin(theta/2)
  int* d_mlim_spin = nullptr;        // [npair]
  std::vector<int> h_mlim_spin;
  double* d_sn_mant = nullptr;       // [mmax+1] seed norm N_j*sqrt((2j)!/((j+q)!(j-q)!)), mantissa
  int* d_sn_exp = nullptr;           // [mmax+1] ... binary exponent
  int64_t* d_soff = nullptr;         // [mmax+2] spin record offsets (one record per l >= max(m,s))
  int64_t nrec_spin = 0;
  double* d_spin_tab = nullptr;      // [nrec_spin][3] {A'_l, B'_l, sigma_l}
  std::map<int, std::pair<glb::LegItem*, int>> spin_item_lists;  // work lists keyed by tile size, built on demand
  double2* d_phase_spin = nullptr;   // [8][nring][mmax+1] phases of a 4-plane batched spin synthesis, allocated on first use
  int64_t rec_capacity = 0;          // doubles available in d_rec

  // analysis (map2alm): per-tile partial sums and a scratch map pair, built lazily
  double* d_partial = nullptr;       // [ana_ntile][nrec][4], one slab per warp tile
  double* d_tmpmap = nullptr;        // [max_batch + 1][npix]: synthesised maps of a Jacobi step, residual
  double* d_ab_tab = nullptr;        // [nrec][2] {a_k, b_k}, {-a_k, a_k+b_k} recurrence coefficients (TMA-streamed by the analysis kernel)
  int ana_ntile = 0;                 // warp tiles (ring-pair tiles x warps per CTA)
  int* d_ana_first_tile = nullptr;   // [mmax+1] first warp tile the analysis kernel writes for m

  // m-split distribution over GPUs (glb_dist_setup): this rank computes the Legendre stage for
  // m = rank (mod world) and the Fourier stage for its own ring bands
  int dist_world = 1, dist_rank = 0;
  int dist_W = 0;                    // m slots per rank = ceil((mmax+1)/world)
  int dist_rows_local = 0;           // rings owned by this rank
  int* d_dist_rowmap = nullptr;      // [nring] row of each ring in the permuted send layout
  int* d_dist_rowidx = nullptr;      // [nring] local row of each owned ring (-1 otherwise)
  int* d_dist_ring_order[4] = {nullptr, nullptr, nullptr, nullptr};
  int n_dist_ring_class[4] = {0, 0, 0, 0};
  // fused Legendre + transpose (glb_dist_p2p_*): receive buffers allocated here with cudaMalloc
  // (CUDA IPC needs that), the peers' buffers opened through their IPC handles
  int p2p_nb = 0;                                                // maps a receive buffer holds
  double2* d_p2p_recv[glb::P2P_NBUF] = {nullptr, nullptr};       // [p2p_nb][world][rows_local][W], halves of ONE cudaMalloc block
  void* p2p_peer[glb::P2P_MAX_WORLD] = {};                       // the peers' blocks as opened here (null: own rank)
  glb::LegP2P* d_p2p_tab = nullptr;                              // [P2P_NBUF]

  // ring FFT
  glb::RingDesc* d_rings = nullptr;  // [nring]
  int* d_ring_order[4] = {nullptr, nullptr, nullptr, nullptr};  // ring index lists per size class (3: long rings)
  int n_ring_class[4] = {0, 0, 0, 0};
  double2* d_long_scratch = nullptr; // work buffers of the long-ring FFT kernels (plans with FFT lengths above 8192)
  double2* d_tw = nullptr;           // twiddles e^{-2 pi i t / TW_N}, t < TW_N/2
  int tw_n = 0;
  double2* d_bf = nullptr;           // concatenated chirp spectra (bit-reversed order, scaled 1/M)
  int64_t bf_total = 0;

  // INT8 tensor-core Legendre path (sht_ozaki.cu), built on first use
  uint8_t* d_oz = nullptr;           // tile blocks: recurrence coefficients + digit planes of the a_lm coefficients
  int64_t* d_oz_toff = nullptr;      // [mmax+2] first tile of every m
  int64_t oz_bytes = 0, oz_tiles = 0;
  uint8_t* d_oz1 = nullptr;          // sets 0 and 1: glb_alm2map_prepare / _finish (the next batch is prepared on a side stream
  int64_t oz_bytes1 = 0;             // while the Legendre kernel of the current one reads the other set; a prepared set must
  uint8_t* d_oz2 = nullptr;          // survive whatever the caller runs in between); set 2: every other call (glb_alm2map,
  int64_t oz_bytes2 = 0;             // the refinement syntheses of glb_map2alm_batch)
  std::vector<cudaEvent_t> ev_prep, ev_fin;  // stage timing of the split form: pairs (prep) and triples (Legendre, ring FFT)
  int legendre_mode = 0;             // 0 auto (INT8 for groups of eight maps at nside >= 1024), 1 FP64 only, 2 INT8 for groups of four and eight

  // workspace
  double* d_rec = nullptr;           // Legendre records  [nrec * (4 + 4*B)]
  double2* d_phase = nullptr;        // [max_batch][nring][mmax+1]
  int64_t workspace_bytes = 0;

  // optional per-stage timing (CUDA events on the launch stream): prep, legendre, ringfft
  bool timing = false;
  std::vector<cudaEvent_t> ev_pool;     // recorded quadruples e0..e3 per group
  double stage_ms[3] = {0.0, 0.0, 0.0};
  int64_t stage_launches[3] = {0, 0, 0};
  int64_t stage_maps = 0;               // maps transformed while timing was on

  // pinned host staging for the host-buffer API
  double* h_pin_in = nullptr;
  double* h_pin_out = nullptr;
  double* d_stage_alm = nullptr;
  double* d_stage_map = nullptr;
};
