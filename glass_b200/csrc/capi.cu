// capi.cu -- extern "C" entry points declared in include/glass_b200.h.
#include <algorithm>
#include <cstring>
#include <new>

#include "plan.h"

namespace glb {
const std::string& last_error();
int plan_build(glb_plan* pl);
void plan_free(glb_plan* pl);
int sht_alm2phase_group(glb_plan* pl, const double2* d_alm, int nb, double2* d_phase, cudaStream_t st);
int sht_prep_group(glb_plan* pl, const double2* d_alm, int nb, cudaStream_t st);
int sht_alm2phase_ozaki(glb_plan* pl, const double2* d_alm, int nb, double2* d_phase, cudaStream_t st);
int sht_ozaki_prep(glb_plan* pl, const double2* d_alm, int nb, cudaStream_t st, int slot = 2);
int sht_ozaki_legendre(glb_plan* pl, int nb, double2* d_phase, cudaStream_t st, int slot = 2);
int sht_legendre_group(glb_plan* pl, int nb, double2* d_phase, cudaStream_t st, bool dist = false, int p2p_buffer = -1);
unsigned long long launch_count();
int measure_fp64_peak(int device, double* tflops, double* ms, cudaStream_t st);

// fold finished event quadruples into the per-stage totals
static int timing_collect(glb_plan* pl) {
  for (size_t i = 0; i + 3 < pl->ev_pool.size(); i += 4) {
    GLB_CUDA_CHECK(cudaEventSynchronize(pl->ev_pool[i + 3]));
    for (int s = 0; s < 3; ++s) {
      float ms = 0.f;
      GLB_CUDA_CHECK(cudaEventElapsedTime(&ms, pl->ev_pool[i + s], pl->ev_pool[i + s + 1]));
      pl->stage_ms[s] += ms;
      pl->stage_launches[s] += 1;
    }
  }
  for (cudaEvent_t e : pl->ev_pool) cudaEventDestroy(e);
  pl->ev_pool.clear();
  // the split form (glb_alm2map_prepare on one stream, glb_alm2map_finish on another)
  for (size_t i = 0; i + 1 < pl->ev_prep.size(); i += 2) {
    float ms = 0.f;
    GLB_CUDA_CHECK(cudaEventSynchronize(pl->ev_prep[i + 1]));
    GLB_CUDA_CHECK(cudaEventElapsedTime(&ms, pl->ev_prep[i], pl->ev_prep[i + 1]));
    pl->stage_ms[0] += ms;
    pl->stage_launches[0] += 1;
  }
  for (size_t i = 0; i + 2 < pl->ev_fin.size(); i += 3) {
    GLB_CUDA_CHECK(cudaEventSynchronize(pl->ev_fin[i + 2]));
    for (int s = 0; s < 2; ++s) {
      float ms = 0.f;
      GLB_CUDA_CHECK(cudaEventElapsedTime(&ms, pl->ev_fin[i + s], pl->ev_fin[i + s + 1]));
      pl->stage_ms[1 + s] += ms;
      pl->stage_launches[1 + s] += 1;
    }
  }
  for (cudaEvent_t e : pl->ev_prep) cudaEventDestroy(e);
  for (cudaEvent_t e : pl->ev_fin) cudaEventDestroy(e);
  pl->ev_prep.clear();
  pl->ev_fin.clear();
  return GLB_OK;
}
int sht_phase2map_group(glb_plan* pl, const double2* d_phase, int nb, double* const* d_maps, const int* kind,
                        const double* tparams, const int* d_mlim, cudaStream_t st, bool dist = false);
int plan_dist_setup(glb_plan* pl, int world, int rank, const int* h_rowmap, const int* h_my_rings, int n_my);
int sht_spin_alm2phase_multi(glb_plan* pl, const double2* const* d_alms, int nb, int spin, double2* d_phase,
                             cudaStream_t st);
int sht_spin_alm2phase(glb_plan* pl, const double2* d_alm1, const double2* d_alm2, int spin, double2* d_phase,
                       cudaStream_t st);
int plan_ensure_spin(glb_plan* pl, int spin);
int plan_ensure_analysis(glb_plan* pl);
int sht_analysis_pass(glb_plan* pl, const double* d_map, const double* d_ring_w, int accumulate, double2* d_alm,
                      cudaStream_t st);
int sht_residual(const double* a, const double* b, int64_t n, double* out, cudaStream_t st);

static inline int group_size(int remaining, int max_batch) {
  const int cap = max_batch >= 4 ? 4 : (max_batch >= 2 ? 2 : 1);
  int g = remaining >= 4 ? 4 : (remaining >= 2 ? 2 : 1);
  return g > cap ? cap : g;
}
}  // namespace glb

using namespace glb;

extern "C" {

const char* glb_version(void) { return "glass_b200 0.1 (sm_100a)"; }
const char* glb_last_error(void) { return glb::last_error().c_str(); }

const char* glb_status_string(int status) {
  switch (status) {
    case GLB_OK: return "ok";
    case GLB_ERR_INVALID_ARG: return "invalid argument";
    case GLB_ERR_UNSUPPORTED: return "unsupported size";
    case GLB_ERR_CUDA: return "CUDA error";
    case GLB_ERR_NOMEM: return "out of memory";
    case GLB_ERR_NOT_POSDEF: return "covariance matrix is not positive definite";
    case GLB_ERR_NEGATIVE_CL: return "negative values in cl";
    default: return "unknown status";
  }
}

int glb_plan_create(glb_plan** plan, int nside, int lmax, int max_batch, int device) {
  GLB_REQUIRE(plan != nullptr, "plan pointer is null");
  GLB_REQUIRE(nside >= 1, "nside must be >= 1");
  GLB_REQUIRE(lmax >= 0, "lmax must be >= 0");
  GLB_REQUIRE(max_batch >= 1, "max_batch must be >= 1");
  glb_plan* pl = new (std::nothrow) glb_plan();
  if (!pl) return GLB_ERR_NOMEM;
  pl->nside = nside;
  pl->lmax = lmax;
  pl->max_batch = max_batch;
  pl->device = device;
  const int rc = plan_build(pl);
  if (rc != GLB_OK) {
    plan_free(pl);
    delete pl;
    return rc;
  }
  *plan = pl;
  return GLB_OK;
}

int glb_plan_destroy(glb_plan* plan) {
  if (!plan) return GLB_OK;
  plan_free(plan);
  delete plan;
  return GLB_OK;
}

int glb_plan_info(const glb_plan* plan, int* nside, int* lmax, int64_t* npix, int64_t* nalm, int* max_batch,
                  int64_t* workspace_bytes) {
  GLB_REQUIRE(plan != nullptr, "plan is null");
  if (nside) *nside = plan->nside;
  if (lmax) *lmax = plan->lmax;
  if (npix) *npix = plan->npix;
  if (nalm) *nalm = plan->nalm;
  if (max_batch) *max_batch = plan->max_batch;
  if (workspace_bytes) *workspace_bytes = plan->workspace_bytes;
  return GLB_OK;
}

int glb_alm2map(glb_plan* plan, const double* d_alm, int nmaps, double* d_map, const int* h_transform,
                const double* h_tparams, void* stream) {
  GLB_REQUIRE(plan && d_alm && d_map, "null pointer");
  GLB_REQUIRE(nmaps >= 1, "nmaps must be >= 1");
  cudaStream_t st = (cudaStream_t)stream;
  GLB_CUDA_CHECK(cudaSetDevice(plan->device));
  int done = 0;
  const int64_t phase_map = (int64_t)plan->nring * (plan->mmax + 1);
  while (done < nmaps) {
    int g = group_size(nmaps - done, plan->max_batch);
    // Legendre stage on the INT8 tensor cores (csrc/sht_ozaki.cu): eight maps on one recurrence (auto, nside >= 1024),
    // or groups of four and eight when the plan says so.  Eight maps need eight phase maps: allocated on first use.
    bool int8 = false;
    double2* phase = plan->d_phase;
    if (plan->legendre_mode != 1 && plan->max_batch >= 4 && g == 4) {
      if (nmaps - done >= 8 && (plan->legendre_mode == 2 || plan->nside >= 1024)) {
        if (!plan->d_phase_spin &&
            cudaMalloc((void**)&plan->d_phase_spin, (size_t)8 * phase_map * sizeof(double2)) != cudaSuccess) {
          cudaGetLastError();  // not enough memory: stay with groups of four
          plan->d_phase_spin = nullptr;
        }
        if (plan->d_phase_spin) {
          g = 8;
          int8 = true;
          phase = plan->d_phase_spin;
        }
      }
      if (!int8 && plan->legendre_mode == 2) int8 = true;
    }
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    if (plan->timing) {
      for (int i = 0; i < 4; ++i) GLB_CUDA_CHECK(cudaEventCreate(&ev[i]));
      GLB_CUDA_CHECK(cudaEventRecord(ev[0], st));
    }
    const double2* alm = reinterpret_cast<const double2*>(d_alm) + (int64_t)done * plan->nalm;
    int rc = int8 ? sht_ozaki_prep(plan, alm, g, st) : sht_prep_group(plan, alm, g, st);
    if (rc != GLB_OK) return rc;
    if (plan->timing) GLB_CUDA_CHECK(cudaEventRecord(ev[1], st));
    rc = int8 ? sht_ozaki_legendre(plan, g, phase, st) : sht_legendre_group(plan, g, phase, st);
    if (rc != GLB_OK) return rc;
    if (plan->timing) GLB_CUDA_CHECK(cudaEventRecord(ev[2], st));
    for (int h = 0; h < g; h += 4) {  // ring FFTs in groups of at most four maps
      const int gh = g - h < 4 ? g - h : 4;
      double* outs[4];
      for (int b = 0; b < gh; ++b) outs[b] = d_map + (int64_t)(done + h + b) * plan->npix;
      rc = sht_phase2map_group(plan, phase + (int64_t)h * phase_map, gh, outs, h_transform ? h_transform + done + h : nullptr,
                               h_tparams ? h_tparams + 2 * (done + h) : nullptr, nullptr, st);
      if (rc != GLB_OK) return rc;
    }
    if (plan->timing) {
      GLB_CUDA_CHECK(cudaEventRecord(ev[3], st));
      for (int i = 0; i < 4; ++i) plan->ev_pool.push_back(ev[i]);
      plan->stage_maps += g;
      if (plan->ev_pool.size() > 4096) timing_collect(plan);
    }
    done += g;
  }
  return GLB_OK;
}

// glb_alm2map in two halves for EIGHT maps on the INT8 path, so that a caller can prepare the next batch (records, digit
// planes; `slot` = which of the two sets of tile blocks) on a side stream while the Legendre kernel of the current batch runs
static bool int8_group_of_eight(const glb_plan* plan) {
  return plan->legendre_mode != 1 && plan->max_batch >= 4 && (plan->legendre_mode == 2 || plan->nside >= 1024);
}

int glb_alm2map_prepare(glb_plan* plan, const double* d_alm, int nmaps, int slot, void* stream) {
  GLB_REQUIRE(plan && d_alm, "null pointer");
  GLB_REQUIRE(slot == 0 || slot == 1, "slot must be 0 or 1");
  if (nmaps != 8 || !int8_group_of_eight(plan)) {
    set_last_error("glb_alm2map_prepare: only groups of eight maps on the INT8 Legendre path are split");
    return GLB_ERR_UNSUPPORTED;
  }
  cudaStream_t st = (cudaStream_t)stream;
  GLB_CUDA_CHECK(cudaSetDevice(plan->device));
  const int64_t phase_map = (int64_t)plan->nring * (plan->mmax + 1);
  if (!plan->d_phase_spin && cudaMalloc((void**)&plan->d_phase_spin, (size_t)8 * phase_map * sizeof(double2)) != cudaSuccess) {
    cudaGetLastError();
    plan->d_phase_spin = nullptr;
    set_last_error("glb_alm2map_prepare: not enough memory for eight phase maps");
    return GLB_ERR_UNSUPPORTED;
  }
  cudaEvent_t ev[2] = {nullptr, nullptr};
  if (plan->timing) {
    for (int i = 0; i < 2; ++i) GLB_CUDA_CHECK(cudaEventCreate(&ev[i]));
    GLB_CUDA_CHECK(cudaEventRecord(ev[0], st));
  }
  const int rc = sht_ozaki_prep(plan, reinterpret_cast<const double2*>(d_alm), 8, st, slot);
  if (rc != GLB_OK) return rc == GLB_ERR_NOMEM ? GLB_ERR_UNSUPPORTED : rc;
  if (plan->timing) {
    GLB_CUDA_CHECK(cudaEventRecord(ev[1], st));
    plan->ev_prep.push_back(ev[0]);
    plan->ev_prep.push_back(ev[1]);
  }
  return GLB_OK;
}

int glb_alm2map_finish(glb_plan* plan, int nmaps, int slot, double* d_map, const int* h_transform, const double* h_tparams,
                       void* stream) {
  GLB_REQUIRE(plan && d_map, "null pointer");
  GLB_REQUIRE(nmaps == 8 && (slot == 0 || slot == 1) && plan->d_phase_spin, "glb_alm2map_finish follows glb_alm2map_prepare");
  cudaStream_t st = (cudaStream_t)stream;
  GLB_CUDA_CHECK(cudaSetDevice(plan->device));
  const int64_t phase_map = (int64_t)plan->nring * (plan->mmax + 1);
  cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
  if (plan->timing) {
    for (int i = 0; i < 3; ++i) GLB_CUDA_CHECK(cudaEventCreate(&ev[i]));
    GLB_CUDA_CHECK(cudaEventRecord(ev[0], st));
  }
  int rc = sht_ozaki_legendre(plan, 8, plan->d_phase_spin, st, slot);
  if (rc != GLB_OK) return rc;
  if (plan->timing) GLB_CUDA_CHECK(cudaEventRecord(ev[1], st));
  for (int h = 0; h < 8; h += 4) {
    double* outs[4];
    for (int b = 0; b < 4; ++b) outs[b] = d_map + (int64_t)(h + b) * plan->npix;
    rc = sht_phase2map_group(plan, plan->d_phase_spin + (int64_t)h * phase_map, 4, outs, h_transform ? h_transform + h : nullptr,
                             h_tparams ? h_tparams + 2 * h : nullptr, nullptr, st);
    if (rc != GLB_OK) return rc;
  }
  if (plan->timing) {
    GLB_CUDA_CHECK(cudaEventRecord(ev[2], st));
    for (int i = 0; i < 3; ++i) plan->ev_fin.push_back(ev[i]);
    plan->stage_maps += 8;
    if (plan->ev_fin.size() > 3072) timing_collect(plan);
  }
  return GLB_OK;
}

int glb_plan_release_scratch(glb_plan* plan) {
  GLB_REQUIRE(plan != nullptr, "plan is null");
  GLB_CUDA_CHECK(cudaSetDevice(plan->device));
  GLB_CUDA_CHECK(cudaDeviceSynchronize());
  cudaFree(plan->d_oz);
  cudaFree(plan->d_oz1);
  cudaFree(plan->d_oz2);
  plan->d_oz = plan->d_oz1 = plan->d_oz2 = nullptr;
  plan->oz_bytes = plan->oz_bytes1 = plan->oz_bytes2 = 0;
  return GLB_OK;
}

int glb_plan_set_legendre_mode(glb_plan* plan, int mode) {
  GLB_REQUIRE(plan != nullptr, "plan is null");
  GLB_REQUIRE(mode >= 0 && mode <= 2, "mode must be 0 (auto), 1 (FP64) or 2 (INT8)");
  plan->legendre_mode = mode;
  return GLB_OK;
}

int glb_alm2map_spin(glb_plan* plan, const double* d_alm1, const double* d_alm2, int spin, double* d_map1,
                     double* d_map2, void* stream) {
  GLB_REQUIRE(plan && d_alm1 && d_map1 && d_map2, "null pointer");
  GLB_REQUIRE(spin >= 1 && spin <= 3, "spin must be 1, 2 or 3");
  GLB_REQUIRE(spin <= plan->lmax, "spin exceeds lmax");
  cudaStream_t st = (cudaStream_t)stream;
  GLB_CUDA_CHECK(cudaSetDevice(plan->device));
  int rc = plan_ensure_spin(plan, spin);
  if (rc != GLB_OK) return rc;
  rc = sht_spin_alm2phase(plan, reinterpret_cast<const double2*>(d_alm1), reinterpret_cast<const double2*>(d_alm2), spin,
                          plan->d_phase, st);
  if (rc != GLB_OK) return rc;
  double* outs[4] = {d_map1, d_map2, nullptr, nullptr};
  return sht_phase2map_group(plan, plan->d_phase, 2, outs, nullptr, nullptr, plan->d_mlim_spin, st);
}

// E-only spin synthesis of nb map pairs at once: the two Wigner-d recurrences are shared by the
// nb coefficient sets (10 DFMA per (l, ring pair) for one map, 7 per map for two, 5.5 for four).
// Four maps need eight phase maps: allocated on first use (17 GB at nside 4096); if that fails the
// batch runs as two pairs in the plan's own phase buffer.
int glb_alm2map_spin_batch(glb_plan* plan, const double* d_alms, int nb, int spin, double* d_maps1, double* d_maps2,
                           void* stream) {
  GLB_REQUIRE(plan && d_alms && d_maps1 && d_maps2, "null pointer");
  GLB_REQUIRE(nb >= 1 && nb <= 4, "nb must be in [1, 4]");
  GLB_REQUIRE(spin >= 1 && spin <= 3, "spin must be 1, 2 or 3");
  GLB_REQUIRE(spin <= plan->lmax, "spin exceeds lmax");
  cudaStream_t st = (cudaStream_t)stream;
  GLB_CUDA_CHECK(cudaSetDevice(plan->device));
  int rc = plan_ensure_spin(plan, spin);
  if (rc != GLB_OK) return rc;
  const double2* alm = reinterpret_cast<const double2*>(d_alms);
  const int64_t phase_map = (int64_t)plan->nring * (plan->mmax + 1);
  int done = 0;
  while (done < nb) {
    int g = nb - done >= 4 ? 4 : (nb - done >= 2 ? 2 : 1);
    if (g > 1 && plan->max_batch < 4) g = 1;  // the plan's phase buffer holds two maps only
    double2* phase = plan->d_phase;
    if (g == 4) {
      if (!plan->d_phase_spin && cudaMalloc((void**)&plan->d_phase_spin, (size_t)8 * phase_map * sizeof(double2)) != cudaSuccess) {
        cudaGetLastError();  // not enough memory for eight phase maps: pairs instead
        plan->d_phase_spin = nullptr;
        g = 2;
      } else {
        phase = plan->d_phase_spin;
      }
    }
    if (g == 1) {
      rc = sht_spin_alm2phase(plan, alm + (int64_t)done * plan->nalm, nullptr, spin, phase, st);
    } else {
      const double2* ptrs[4] = {nullptr, nullptr, nullptr, nullptr};
      for (int b = 0; b < g; ++b) ptrs[b] = alm + (int64_t)(done + b) * plan->nalm;
      rc = sht_spin_alm2phase_multi(plan, ptrs, g, spin, phase, st);
    }
    if (rc != GLB_OK) return rc;
    for (int h = 0; h < 2 * g; h += 4) {  // ring FFTs, four maps per launch group
      double* outs[4] = {nullptr, nullptr, nullptr, nullptr};
      const int n = std::min(4, 2 * g - h);
      for (int q = 0; q < n; ++q) {
        const int b = done + (h + q) / 2;
        outs[q] = (((h + q) & 1) ? d_maps2 : d_maps1) + (int64_t)b * plan->npix;
      }
      rc = sht_phase2map_group(plan, phase + (int64_t)h * phase_map, n, outs, nullptr, nullptr, plan->d_mlim_spin, st);
      if (rc != GLB_OK) return rc;
    }
    done += g;
  }
  return GLB_OK;
}

// nb maps at once (nb = 1, 2 or 4 <= max_batch): the analyses run map by map, the synthesis of
// every Jacobi refinement runs as ONE batched transform -- nb maps share the Legendre recurrence,
// 34 ms per map at nside 4096 for nb = 4 against 50 ms alone.
int glb_map2alm_batch(glb_plan* plan, const double* d_maps, int nb, const double* d_ring_weights, int niter,
                      double* d_alms, void* stream) {
  GLB_REQUIRE(plan && d_maps && d_alms, "null pointer");
  GLB_REQUIRE(nb == 1 || nb == 2 || nb == 4, "nb must be 1, 2 or 4");
  GLB_REQUIRE(nb <= plan->max_batch, "nb exceeds max_batch");
  GLB_REQUIRE(niter >= 0 && niter <= 100, "niter must be in [0, 100]");
  cudaStream_t st = (cudaStream_t)stream;
  GLB_CUDA_CHECK(cudaSetDevice(plan->device));
  int rc = plan_ensure_analysis(plan);
  if (rc != GLB_OK) return rc;
  double2* alm = reinterpret_cast<double2*>(d_alms);
  for (int b = 0; b < nb; ++b)
    if ((rc = sht_analysis_pass(plan, d_maps + (int64_t)b * plan->npix, d_ring_weights, 0, alm + (int64_t)b * plan->nalm, st)) != GLB_OK)
      return rc;
  for (int it = 0; it < niter; ++it) {
    // alm += A(map - S(alm)); d_tmpmap: [max_batch] synthesised maps, then one residual
    double* resid = plan->d_tmpmap + (int64_t)plan->max_batch * plan->npix;
    // the refinement syntheses of four planes: on the INT8 tensor cores where that is faster (nside >= 4096: 115 against
    // 141 ms), their error (1e-11 of the map) is far below what an iteration still corrects
    const bool int8 = nb == 4 && plan->legendre_mode != 1 && (plan->legendre_mode == 2 || plan->nside >= 4096);
    if ((rc = int8 ? sht_alm2phase_ozaki(plan, alm, nb, plan->d_phase, st) : sht_alm2phase_group(plan, alm, nb, plan->d_phase, st)) != GLB_OK)
      return rc;
    double* outs[4] = {nullptr, nullptr, nullptr, nullptr};
    for (int b = 0; b < nb; ++b) outs[b] = plan->d_tmpmap + (int64_t)b * plan->npix;
    if ((rc = sht_phase2map_group(plan, plan->d_phase, nb, outs, nullptr, nullptr, nullptr, st)) != GLB_OK) return rc;
    for (int b = 0; b < nb; ++b) {
      if ((rc = sht_residual(d_maps + (int64_t)b * plan->npix, outs[b], plan->npix, resid, st)) != GLB_OK) return rc;
      if ((rc = sht_analysis_pass(plan, resid, d_ring_weights, 1, alm + (int64_t)b * plan->nalm, st)) != GLB_OK) return rc;
    }
  }
  return GLB_OK;
}

int glb_map2alm(glb_plan* plan, const double* d_map, const double* d_ring_weights, int niter, double* d_alm,
                void* stream) {
  return glb_map2alm_batch(plan, d_map, 1, d_ring_weights, niter, d_alm, stream);
}

// Host-buffer entry (the one INTEGRATION.md's stub binds): pageable NumPy memory in and out.
// A cudaMemcpy from pageable memory is staged by the driver through a small bounce buffer and
// blocks the stream; here both directions go through two page-locked chunks of the plan
// (allocated once), so the host memcpy of chunk i overlaps the DMA of chunk i+1 and the
// transfers run at the link rate.  Everything is on `stream`; the call returns when h_map is filled.
static constexpr size_t HOST_CHUNK = (size_t)64 << 20;

static int staged_h2d(glb_plan* plan, void* d_dst, const void* h_src, size_t bytes, cudaEvent_t* ev, cudaStream_t st) {
  char* pin[2] = {reinterpret_cast<char*>(plan->h_pin_in), reinterpret_cast<char*>(plan->h_pin_in) + HOST_CHUNK};
  size_t off = 0;
  for (int c = 0; off < bytes; ++c, off += HOST_CHUNK) {
    const size_t n = std::min(HOST_CHUNK, bytes - off);
    if (c >= 2) GLB_CUDA_CHECK(cudaEventSynchronize(ev[c & 1]));  // the DMA that last read this chunk has finished
    memcpy(pin[c & 1], static_cast<const char*>(h_src) + off, n);
    GLB_CUDA_CHECK(cudaMemcpyAsync(static_cast<char*>(d_dst) + off, pin[c & 1], n, cudaMemcpyHostToDevice, st));
    GLB_CUDA_CHECK(cudaEventRecord(ev[c & 1], st));
  }
  return GLB_OK;
}

static int staged_d2h(glb_plan* plan, void* h_dst, const void* d_src, size_t bytes, cudaEvent_t* ev, cudaStream_t st) {
  char* pin[2] = {reinterpret_cast<char*>(plan->h_pin_out), reinterpret_cast<char*>(plan->h_pin_out) + HOST_CHUNK};
  size_t off = 0, prev_off = 0, prev_n = 0;
  int c = 0;
  for (; off < bytes; ++c, off += HOST_CHUNK) {
    const size_t n = std::min(HOST_CHUNK, bytes - off);
    GLB_CUDA_CHECK(cudaMemcpyAsync(pin[c & 1], static_cast<const char*>(d_src) + off, n, cudaMemcpyDeviceToHost, st));
    GLB_CUDA_CHECK(cudaEventRecord(ev[c & 1], st));
    if (c >= 1) {  // hand the previous chunk to the caller while this one is in flight
      GLB_CUDA_CHECK(cudaEventSynchronize(ev[(c - 1) & 1]));
      memcpy(static_cast<char*>(h_dst) + prev_off, pin[(c - 1) & 1], prev_n);
    }
    prev_off = off;
    prev_n = n;
  }
  if (c >= 1) {
    GLB_CUDA_CHECK(cudaEventSynchronize(ev[(c - 1) & 1]));
    memcpy(static_cast<char*>(h_dst) + prev_off, pin[(c - 1) & 1], prev_n);
  }
  return GLB_OK;
}

int glb_alm2map_host(glb_plan* plan, const double* h_alm, int nmaps, double* h_map, const int* h_transform,
                     const double* h_tparams, void* stream) {
  GLB_REQUIRE(plan && h_alm && h_map, "null pointer");
  GLB_REQUIRE(nmaps >= 1 && nmaps <= plan->max_batch, "nmaps must be in [1, max_batch]");
  cudaStream_t st = (cudaStream_t)stream;
  GLB_CUDA_CHECK(cudaSetDevice(plan->device));
  const size_t alm_bytes = (size_t)plan->nalm * 2 * sizeof(double) * plan->max_batch;
  const size_t map_bytes = (size_t)plan->npix * sizeof(double) * plan->max_batch;
  if (!plan->d_stage_alm) GLB_CUDA_CHECK(cudaMalloc((void**)&plan->d_stage_alm, alm_bytes));
  if (!plan->d_stage_map) GLB_CUDA_CHECK(cudaMalloc((void**)&plan->d_stage_map, map_bytes));
  if (!plan->h_pin_in) GLB_CUDA_CHECK(cudaHostAlloc((void**)&plan->h_pin_in, 2 * HOST_CHUNK, cudaHostAllocDefault));
  if (!plan->h_pin_out) GLB_CUDA_CHECK(cudaHostAlloc((void**)&plan->h_pin_out, 2 * HOST_CHUNK, cudaHostAllocDefault));
  cudaEvent_t ev[2];
  for (auto& e : ev) GLB_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  const size_t a_bytes = (size_t)plan->nalm * 2 * sizeof(double) * nmaps;
  const size_t m_bytes = (size_t)plan->npix * sizeof(double) * nmaps;
  int rc = staged_h2d(plan, plan->d_stage_alm, h_alm, a_bytes, ev, st);
  if (rc == GLB_OK) rc = glb_alm2map(plan, plan->d_stage_alm, nmaps, plan->d_stage_map, h_transform, h_tparams, stream);
  if (rc == GLB_OK) rc = staged_d2h(plan, h_map, plan->d_stage_map, m_bytes, ev, st);
  const cudaError_t sync = cudaStreamSynchronize(st);
  for (auto& e : ev) cudaEventDestroy(e);
  if (rc != GLB_OK) return rc;
  GLB_CUDA_CHECK(sync);
  return GLB_OK;
}

int glb_debug_alm2phase(glb_plan* plan, const double* d_alm, int nmaps, double* d_phase, void* stream) {
  GLB_REQUIRE(plan && d_alm && d_phase, "null pointer");
  GLB_REQUIRE(nmaps == 1 || nmaps == 2 || nmaps == 4, "nmaps must be 1, 2 or 4");
  GLB_REQUIRE(nmaps <= plan->max_batch, "nmaps exceeds max_batch");
  GLB_CUDA_CHECK(cudaSetDevice(plan->device));
  GLB_CUDA_CHECK(cudaMemsetAsync(d_phase, 0, (size_t)nmaps * plan->nring * (plan->mmax + 1) * sizeof(double2),
                                 (cudaStream_t)stream));
  return sht_alm2phase_group(plan, reinterpret_cast<const double2*>(d_alm), nmaps,
                             reinterpret_cast<double2*>(d_phase), (cudaStream_t)stream);
}

int glb_debug_alm2phase_int8(glb_plan* plan, const double* d_alm, int nmaps, double* d_phase, void* stream) {
  GLB_REQUIRE(plan && d_alm && d_phase, "null pointer");
  GLB_REQUIRE(nmaps == 4 || nmaps == 8, "nmaps must be 4 or 8");
  GLB_CUDA_CHECK(cudaSetDevice(plan->device));
  GLB_CUDA_CHECK(cudaMemsetAsync(d_phase, 0, (size_t)nmaps * plan->nring * (plan->mmax + 1) * sizeof(double2),
                                 (cudaStream_t)stream));
  return sht_alm2phase_ozaki(plan, reinterpret_cast<const double2*>(d_alm), nmaps, reinterpret_cast<double2*>(d_phase),
                             (cudaStream_t)stream);
}

int glb_debug_phase2map(glb_plan* plan, const double* d_phase, int nmaps, double* d_map, void* stream) {
  GLB_REQUIRE(plan && d_phase && d_map, "null pointer");
  GLB_REQUIRE(nmaps >= 1 && nmaps <= 4, "nmaps must be in [1, 4]");
  GLB_CUDA_CHECK(cudaSetDevice(plan->device));
  double* outs[4];
  for (int b = 0; b < nmaps; ++b) outs[b] = d_map + (int64_t)b * plan->npix;
  return sht_phase2map_group(plan, reinterpret_cast<const double2*>(d_phase), nmaps, outs, nullptr, nullptr, nullptr,
                             (cudaStream_t)stream);
}

int glb_dist_setup(glb_plan* plan, int world, int rank, const int* h_rowmap, const int* h_my_rings, int n_my_rings) {
  GLB_REQUIRE(plan && h_rowmap && h_my_rings, "null pointer");
  GLB_REQUIRE(world >= 1 && rank >= 0 && rank < world && n_my_rings >= 0, "bad world/rank");
  return plan_dist_setup(plan, world, rank, h_rowmap, h_my_rings, n_my_rings);
}

int glb_dist_alm2phase(glb_plan* plan, const double* d_alm, int nmaps, double* d_send, void* stream) {
  GLB_REQUIRE(plan && d_alm && d_send, "null pointer");
  GLB_REQUIRE(nmaps == 1 || nmaps == 2 || nmaps == 4, "nmaps must be 1, 2 or 4");
  GLB_REQUIRE(nmaps <= plan->max_batch, "nmaps exceeds max_batch");
  GLB_REQUIRE(plan->dist_W > 0, "glb_dist_setup has not been called");
  cudaStream_t st = (cudaStream_t)stream;
  GLB_CUDA_CHECK(cudaSetDevice(plan->device));
  GLB_CUDA_CHECK(cudaMemsetAsync(d_send, 0, (size_t)nmaps * plan->nring * plan->dist_W * sizeof(double2), st));
  int rc = sht_prep_group(plan, reinterpret_cast<const double2*>(d_alm), nmaps, st);
  if (rc != GLB_OK) return rc;
  return sht_legendre_group(plan, nmaps, reinterpret_cast<double2*>(d_send), st, true);
}

int glb_dist_phase2map(glb_plan* plan, const double* d_recv, int nmaps, double* d_map, const int* h_transform,
                       const double* h_tparams, void* stream) {
  GLB_REQUIRE(plan && d_recv && d_map, "null pointer");
  GLB_REQUIRE(nmaps >= 1 && nmaps <= 4, "nmaps must be in [1, 4]");
  GLB_REQUIRE(plan->dist_W > 0, "glb_dist_setup has not been called");
  GLB_CUDA_CHECK(cudaSetDevice(plan->device));
  double* outs[4];
  for (int b = 0; b < nmaps; ++b) outs[b] = d_map + (int64_t)b * plan->npix;
  return sht_phase2map_group(plan, reinterpret_cast<const double2*>(d_recv), nmaps, outs, h_transform, h_tparams,
                             nullptr, (cudaStream_t)stream, true);
}

// ---- fused Legendre + transpose over peer memory ---------------------------------------------
// bytes of ONE receive buffer of a rank owning `rows` rings, rounded so that the second buffer
// of the block stays 256-byte aligned
static size_t p2p_buffer_bytes(const glb_plan* plan, int nmaps_max, int rows) {
  const size_t b = (size_t)nmaps_max * plan->dist_world * std::max(rows, 1) * plan->dist_W * sizeof(double2);
  return (b + 255) / 256 * 256;
}

int glb_dist_p2p_alloc(glb_plan* plan, int nmaps_max, void* h_handle) {
  GLB_REQUIRE(plan && h_handle, "null pointer");
  GLB_REQUIRE(plan->dist_W > 0, "glb_dist_setup has not been called");
  GLB_REQUIRE(nmaps_max >= 1 && nmaps_max <= 4, "nmaps_max must be in [1, 4]");
  GLB_REQUIRE(plan->dist_world <= glb::P2P_MAX_WORLD, "peer stores support at most 8 ranks");
  GLB_REQUIRE(plan->p2p_nb == 0, "receive buffers already allocated");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  GLB_CUDA_CHECK(cudaSetDevice(plan->device));
  const size_t bytes = p2p_buffer_bytes(plan, nmaps_max, plan->dist_rows_local);
  // ONE plain cudaMalloc block for both buffers (CUDA IPC cannot export memory of a stream-ordered
  // pool; one allocation = one handle for the peers to open).  Zeroed once: entries no rank ever
  // writes (rings beyond mlim) stay zero, the others are rewritten by every transform.
  char* block = nullptr;
  GLB_CUDA_CHECK(cudaMalloc((void**)&block, glb::P2P_NBUF * bytes));
  GLB_CUDA_CHECK(cudaMemset(block, 0, glb::P2P_NBUF * bytes));
  for (int i = 0; i < glb::P2P_NBUF; ++i) plan->d_p2p_recv[i] = reinterpret_cast<double2*>(block + i * bytes);
  cudaIpcMemHandle_t h;
  GLB_CUDA_CHECK(cudaIpcGetMemHandle(&h, block));
  memcpy(h_handle, &h, 64);
  GLB_CUDA_CHECK(cudaDeviceSynchronize());
  plan->p2p_nb = nmaps_max;
  return GLB_OK;
}

int glb_dist_p2p_open(glb_plan* plan, const void* h_all_handles, const int* h_rows) {
  GLB_REQUIRE(plan && h_all_handles && h_rows, "null pointer");
  GLB_REQUIRE(plan->p2p_nb > 0, "glb_dist_p2p_alloc has not been called");
  GLB_REQUIRE(plan->d_p2p_tab == nullptr, "peer buffers already opened");
  GLB_REQUIRE(h_rows[plan->dist_rank] == plan->dist_rows_local, "h_rows disagrees with glb_dist_setup");
  GLB_CUDA_CHECK(cudaSetDevice(plan->device));
  const int world = plan->dist_world;
  glb::LegP2P tab[glb::P2P_NBUF];
  memset(tab, 0, sizeof(tab));
  for (int i = 0; i < glb::P2P_NBUF; ++i) {
    tab[i].rank = plan->dist_rank;
    tab[i].world = world;
    for (int d = 0; d < world; ++d) tab[i].rowstart[d + 1] = tab[i].rowstart[d] + h_rows[d];
    for (int d = world; d < glb::P2P_MAX_WORLD; ++d) tab[i].rowstart[d + 1] = tab[i].rowstart[world];
    GLB_REQUIRE(tab[i].rowstart[world] == plan->nring, "h_rows does not add up to the number of rings");
  }
  for (int d = 0; d < world; ++d) {
    char* block = reinterpret_cast<char*>(plan->d_p2p_recv[0]);
    if (d != plan->dist_rank) {
      cudaIpcMemHandle_t h;
      memcpy(&h, static_cast<const char*>(h_all_handles) + 64 * (size_t)d, 64);
      void* ptr = nullptr;
      GLB_CUDA_CHECK(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
      plan->p2p_peer[d] = ptr;
      block = static_cast<char*>(ptr);
    }
    // every rank allocated with the same nmaps_max, so the peer's buffer size follows from its rows
    const size_t bytes_d = p2p_buffer_bytes(plan, plan->p2p_nb, h_rows[d]);
    for (int i = 0; i < glb::P2P_NBUF; ++i) tab[i].base[d] = reinterpret_cast<double2*>(block + i * bytes_d);
  }
  GLB_CUDA_CHECK(cudaMalloc((void**)&plan->d_p2p_tab, sizeof(tab)));
  GLB_CUDA_CHECK(cudaMemcpy(plan->d_p2p_tab, tab, sizeof(tab), cudaMemcpyHostToDevice));
  return GLB_OK;
}

int glb_dist_alm2phase_p2p(glb_plan* plan, const double* d_alm, int nmaps, int buffer, void* stream) {
  GLB_REQUIRE(plan && d_alm, "null pointer");
  GLB_REQUIRE(nmaps == 1 || nmaps == 2 || nmaps == 4, "nmaps must be 1, 2 or 4");
  GLB_REQUIRE(nmaps <= plan->max_batch && nmaps <= plan->p2p_nb, "nmaps exceeds max_batch or the receive buffers");
  GLB_REQUIRE(plan->d_p2p_tab != nullptr, "glb_dist_p2p_open has not been called");
  GLB_REQUIRE(buffer >= 0 && buffer < glb::P2P_NBUF, "buffer must be 0 or 1");
  cudaStream_t st = (cudaStream_t)stream;
  GLB_CUDA_CHECK(cudaSetDevice(plan->device));
  int rc = sht_prep_group(plan, reinterpret_cast<const double2*>(d_alm), nmaps, st);
  if (rc != GLB_OK) return rc;
  return sht_legendre_group(plan, nmaps, nullptr, st, true, buffer);
}

int glb_dist_p2p_recv(glb_plan* plan, int buffer, double** d_recv) {
  GLB_REQUIRE(plan && d_recv, "null pointer");
  GLB_REQUIRE(plan->p2p_nb > 0, "glb_dist_p2p_alloc has not been called");
  GLB_REQUIRE(buffer >= 0 && buffer < glb::P2P_NBUF, "buffer must be 0 or 1");
  *d_recv = reinterpret_cast<double*>(plan->d_p2p_recv[buffer]);
  return GLB_OK;
}

int glb_plan_timing_enable(glb_plan* plan, int enable) {
  GLB_REQUIRE(plan != nullptr, "plan is null");
  GLB_CUDA_CHECK(cudaSetDevice(plan->device));
  timing_collect(plan);
  plan->timing = enable != 0;
  for (int s = 0; s < 3; ++s) {
    plan->stage_ms[s] = 0.0;
    plan->stage_launches[s] = 0;
  }
  plan->stage_maps = 0;
  return GLB_OK;
}

int glb_plan_timing_read(glb_plan* plan, double* ms3, int64_t* launches3, int64_t* nmaps) {
  GLB_REQUIRE(plan && ms3 && launches3 && nmaps, "null pointer");
  GLB_CUDA_CHECK(cudaSetDevice(plan->device));
  const int rc = timing_collect(plan);
  if (rc != GLB_OK) return rc;
  for (int s = 0; s < 3; ++s) {
    ms3[s] = plan->stage_ms[s];
    launches3[s] = plan->stage_launches[s];
  }
  *nmaps = plan->stage_maps;
  return GLB_OK;
}

uint64_t glb_kernel_launch_count(void) { return (uint64_t)glb::launch_count(); }

int glb_measure_fp64_peak(int device, double* tflops, double* ms, void* stream) {
  GLB_REQUIRE(tflops && ms, "null pointer");
  return glb::measure_fp64_peak(device, tflops, ms, (cudaStream_t)stream);
}

int glb_debug_mlim(const glb_plan* plan, int* h_mlim) {
  GLB_REQUIRE(plan && h_mlim, "null pointer");
  for (int r = 0; r < plan->npair; ++r) h_mlim[r] = plan->h_mlim[r];
  return GLB_OK;
}

}  // extern "C"
