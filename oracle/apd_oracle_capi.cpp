// TEST INFRASTRUCTURE — C ABI of the CPU oracle, mirroring include/apdgicp.h
// one-to-one (prefix apdo_ instead of apd_) so that the parity tests drive the
// oracle and the CUDA library with the same call sequences.
#include <chrono>
#include <cstring>

#include "apd_oracle.hpp"

using namespace apdo;

namespace {
struct Handle {
  FastAPDGICP g;
};
inline Handle* H(void* h) { return reinterpret_cast<Handle*>(h); }

std::vector<PointXYZL> gather(const void* pts, int n, int stride, int xyz_off, int label_off) {
  std::vector<PointXYZL> out((size_t)(n > 0 ? n : 0));
  const char* base = reinterpret_cast<const char*>(pts);
  for (int i = 0; i < n; i++) {
    const char* p = base + (size_t)i * stride;
    float xyz[3], l = 0.f;
    std::memcpy(xyz, p + xyz_off, 12);
    if (label_off >= 0) std::memcpy(&l, p + label_off, 4);
    out[i] = PointXYZL{xyz[0], xyz[1], xyz[2], l};
  }
  return out;
}
void cov_to_4x4(const M3& c, double* o) {  // column-major 4x4, 3x3 block only
  for (int e = 0; e < 16; e++) o[e] = 0.0;
  for (int r = 0; r < 3; r++)
    for (int cc = 0; cc < 3; cc++) o[cc * 4 + r] = c(r, cc);
}
M3 cov_from_4x4(const double* o) {
  M3 c;
  for (int r = 0; r < 3; r++)
    for (int cc = 0; cc < 3; cc++) c(r, cc) = o[cc * 4 + r];
  return c;
}
M4 m4_from_colmajor(const double* T) {
  M4 t;
  for (int r = 0; r < 4; r++)
    for (int c = 0; c < 4; c++) t(r, c) = T[c * 4 + r];
  return t;
}
}  // namespace

extern "C" {

int apdo_has_nanoflann(void) {
#ifdef APDO_USE_NANOFLANN
  return 1;
#else
  return 0;
#endif
}
int apdo_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

int apdo_create(void** out) {
  *out = new Handle();
  return APD_OK;
}
int apdo_destroy(void* h) {
  delete H(h);
  return APD_OK;
}
const char* apdo_last_error(void* h) { return H(h)->g.error.c_str(); }
int apdo_default_params(apd_params* p) {
  FastAPDGICP g;
  *p = g.params;
  return APD_OK;
}
int apdo_set_params(void* h, const apd_params* p) {
  H(h)->g.params = *p;
  return APD_OK;
}
int apdo_get_params(void* h, apd_params* p) {
  *p = H(h)->g.params;
  return APD_OK;
}
// reference setNumThreads (fast_apdgicp_impl.hpp:34-42): 0 = omp_get_max_threads()
int apdo_set_num_threads(void* h, int n) {
  if (n == 0) n = apdo_max_threads();
  H(h)->g.num_threads = n;
  return APD_OK;
}
// 0 brute force, 1 kd-tree (default), 2 nanoflann (oracle/_ref build only)
int apdo_set_search(void* h, int kind) {
#ifndef APDO_USE_NANOFLANN
  if (kind == 2) return APD_ERR_UNSUPPORTED;
#endif
  H(h)->g.search_kind = kind;
  return APD_OK;
}

int apdo_set_source(void* h, const void* pts, int32_t n, int32_t stride, int32_t xyz_off, int32_t label_off, uint64_t key) {
  FastAPDGICP& g = H(h)->g;
  if (g.has_source && key != 0 && key == g.source_key) return APD_OK;
  g.setInputSource(gather(pts, n, stride, xyz_off, label_off), key);
  return APD_OK;
}
int apdo_set_target(void* h, const void* pts, int32_t n, int32_t stride, int32_t xyz_off, int32_t label_off, uint64_t key) {
  FastAPDGICP& g = H(h)->g;
  if (g.has_target && key != 0 && key == g.target_key) return APD_OK;
  g.setInputTarget(gather(pts, n, stride, xyz_off, label_off), key);
  return APD_OK;
}
int apdo_swap_source_and_target(void* h) {
  H(h)->g.swapSourceAndTarget();
  return APD_OK;
}
int apdo_clear_source(void* h) {
  H(h)->g.clearSource();
  return APD_OK;
}
int apdo_clear_target(void* h) {
  H(h)->g.clearTarget();
  return APD_OK;
}

static int set_covs(std::vector<M3>& dst, const double* covs, int n) {
  dst.resize(n);
  for (int i = 0; i < n; i++) dst[i] = cov_from_4x4(covs + (size_t)i * 16);
  return APD_OK;
}
int apdo_set_source_covariances(void* h, const double* covs, int32_t n) {
  H(h)->g.source_neighbors.clear();
  return set_covs(H(h)->g.source_covs, covs, n);
}
int apdo_set_target_covariances(void* h, const double* covs, int32_t n) {
  H(h)->g.target_neighbors.clear();
  return set_covs(H(h)->g.target_covs, covs, n);
}
static int get_covs(FastAPDGICP& g, bool src, double* covs, int n) {
  std::vector<PointXYZL>& cloud = src ? g.source : g.target;
  std::vector<M3>& cv = src ? g.source_covs : g.target_covs;
  if (!(src ? g.has_source : g.has_target)) return APD_ERR_INVALID;
  if (cv.size() != cloud.size()) {
    if (!g.calculate_covariances(cloud, src ? *g.source_search : *g.target_search, cv, src ? &g.source_neighbors : &g.target_neighbors))
      return APD_ERR_TOO_FEW;
  }
  if (n != (int)cv.size()) return APD_ERR_INVALID;
  for (int i = 0; i < n; i++) cov_to_4x4(cv[i], covs + (size_t)i * 16);
  return APD_OK;
}
int apdo_get_source_covariances(void* h, double* covs, int32_t n) { return get_covs(H(h)->g, true, covs, n); }
int apdo_get_target_covariances(void* h, double* covs, int32_t n) { return get_covs(H(h)->g, false, covs, n); }
int apdo_get_neighbors(void* h, int32_t which, int32_t* out, int32_t n, int32_t k) {
  FastAPDGICP& g = H(h)->g;
  const std::vector<int>& nb = which == 0 ? g.source_neighbors : g.target_neighbors;
  if ((size_t)n * k != nb.size()) return APD_ERR_INVALID;
  std::memcpy(out, nb.data(), nb.size() * sizeof(int));
  return APD_OK;
}

int apdo_align(void* h, const float* guess, float* T_out, double* T_out_f64, double* H_out, int32_t* converged,
               int32_t* iterations, float* aligned_xyz) {
  FastAPDGICP& g = H(h)->g;
  if (!g.align(guess)) return g.error.find("fewer") != std::string::npos ? APD_ERR_TOO_FEW : APD_ERR_INVALID;
  if (T_out) std::memcpy(T_out, g.final_transformation, sizeof(float) * 16);
  if (T_out_f64)
    for (int r = 0; r < 4; r++)
      for (int c = 0; c < 4; c++) T_out_f64[c * 4 + r] = g.final_pose_f64(r, c);
  if (H_out) std::memcpy(H_out, g.final_hessian, sizeof(double) * 36);
  if (converged) *converged = g.converged ? 1 : 0;
  if (iterations) *iterations = g.nr_iterations;
  if (aligned_xyz) {
    // pcl::transformPointCloud(*input_, output, final_transformation_) [ext]
    const float* T = g.final_transformation;
    for (size_t i = 0; i < g.source.size(); i++) {
      const PointXYZL& p = g.source[i];
      for (int r = 0; r < 3; r++) aligned_xyz[3 * i + r] = ((T[0 * 4 + r] * p.x + T[1 * 4 + r] * p.y) + T[2 * 4 + r] * p.z) + T[3 * 4 + r];
    }
  }
  return APD_OK;
}
int apdo_linearize(void* h, const double* T, double* Hm, double* b, double* err) {
  FastAPDGICP& g = H(h)->g;
  if (!g.ensure_covariances()) return g.error.find("fewer") != std::string::npos ? APD_ERR_TOO_FEW : APD_ERR_INVALID;
  const double e = g.linearize(m4_from_colmajor(T), Hm, b);
  if (err) *err = e;
  return APD_OK;
}
int apdo_compute_error(void* h, const double* T, double* err) {
  FastAPDGICP& g = H(h)->g;
  if (g.params.variant == APD_VARIANT_VGICP) {
    if (g.voxel_correspondences.size() != g.source.size() * (size_t)g.n_offsets()) return APD_ERR_INVALID;
  } else if (g.correspondences.size() != g.source.size()) {
    return APD_ERR_INVALID;
  }
  *err = g.compute_error(m4_from_colmajor(T));
  return APD_OK;
}
int apdo_update_correspondences(void* h, const double* T) {
  FastAPDGICP& g = H(h)->g;
  if (!g.ensure_covariances()) return g.error.find("fewer") != std::string::npos ? APD_ERR_TOO_FEW : APD_ERR_INVALID;
  if (g.params.variant == APD_VARIANT_VGICP) {
    if (!g.voxelmap_valid) g.create_voxelmap();
    g.vgicp_update_correspondences(m4_from_colmajor(T));
    return APD_OK;
  }
  g.update_correspondences(m4_from_colmajor(T));
  return APD_OK;
}
int apdo_get_correspondences(void* h, int32_t* idx, float* sq, int32_t n) {
  FastAPDGICP& g = H(h)->g;
  if ((size_t)n != g.correspondences.size()) return APD_ERR_INVALID;
  if (idx) std::memcpy(idx, g.correspondences.data(), sizeof(int) * n);
  if (sq) std::memcpy(sq, g.sq_distances.data(), sizeof(float) * n);
  return APD_OK;
}
int apdo_get_mahalanobis(void* h, double* maha, int32_t n) {
  FastAPDGICP& g = H(h)->g;
  if ((size_t)n != g.mahalanobis.size()) return APD_ERR_INVALID;
  for (int i = 0; i < n; i++) {
    if (g.correspondences[i] < 0) {
      for (int e = 0; e < 16; e++) maha[(size_t)i * 16 + e] = 0.0;
    } else {
      cov_to_4x4(g.mahalanobis[i], maha + (size_t)i * 16);
    }
  }
  return APD_OK;
}
int apdo_vgicp_get_voxels(void* h, int32_t* n_voxels, int32_t* coords, int32_t* counts, double* means, double* covs, int32_t capacity) {
  FastAPDGICP& g = H(h)->g;
  const int n = g.voxelmap_valid ? (int)g.voxels.size() : 0;
  if (n_voxels) *n_voxels = n;
  for (int i = 0; i < n && i < capacity; i++) {
    const FastAPDGICP::Voxel& v = g.voxels[i];
    if (coords) for (int a = 0; a < 3; a++) coords[3 * i + a] = v.coord[a];
    if (counts) counts[i] = v.num_points;
    if (means) for (int a = 0; a < 3; a++) means[3 * i + a] = v.mean[a];
    if (covs) for (int e = 0; e < 9; e++) covs[9 * (size_t)i + e] = v.cov.m[e];
  }
  return APD_OK;
}
int apdo_vgicp_get_correspondences(void* h, int32_t* voxel, double* maha, int32_t n_source, int32_t n_offsets) {
  FastAPDGICP& g = H(h)->g;
  const size_t n = (size_t)n_source * n_offsets;
  if (n != g.voxel_correspondences.size() || n_offsets != g.n_offsets()) return APD_ERR_INVALID;
  for (size_t i = 0; i < n; i++) {
    if (voxel) voxel[i] = g.voxel_correspondences[i];
    if (maha) for (int e = 0; e < 9; e++) maha[9 * i + e] = g.voxel_correspondences[i] >= 0 ? g.voxel_mahalanobis[i].m[e] : 0.0;
  }
  return APD_OK;
}
int apdo_fitness(void* h, const float* T, double max_range, double* score, int32_t* n_in_range, double inlier_sq_thr,
                 int32_t* n_inliers) {
  FastAPDGICP& g = H(h)->g;
  if (!g.has_source || !g.has_target) return APD_ERR_INVALID;
  int a = 0, b = 0;
  const double s = g.fitness(T, max_range, &a, inlier_sq_thr, &b);
  if (score) *score = s;
  if (n_in_range) *n_in_range = a;
  if (n_inliers) *n_inliers = b;
  return APD_OK;
}
int apdo_get_lm_trace(void* h, double* rows, int32_t max_rows, int32_t* n_rows) {
  FastAPDGICP& g = H(h)->g;
  const int n = std::min<int>(max_rows, (int)g.lm_trace.size());
  for (int i = 0; i < n; i++) std::memcpy(rows + (size_t)i * 8, &g.lm_trace[i], sizeof(double) * 8);
  *n_rows = n;
  return APD_OK;
}
// work counters (how many linearize / compute_error passes the last runs made)
int apdo_get_counters(void* h, int64_t* n_linearize, int64_t* n_compute_error) {
  *n_linearize = H(h)->g.n_linearize;
  *n_compute_error = H(h)->g.n_compute_error;
  return APD_OK;
}

// Timed CPU baseline, protocol of the reference's benchmark app
// (fast_apdgicp/src/align.cpp:57-83): `reps` times {clearTarget; clearSource;
// setInputTarget; setInputSource; align}. Returns per-rep wall milliseconds.
int apdo_bench_align(void* h, const void* src, int32_t n_src, const void* tgt, int32_t n_tgt, int32_t stride,
                     int32_t xyz_off, int32_t label_off, const float* guess, int32_t reps, double* ms_out) {
  FastAPDGICP& g = H(h)->g;
  for (int r = 0; r < reps; r++) {
    const auto t0 = std::chrono::steady_clock::now();
    g.clearTarget();
    g.clearSource();
    g.setInputTarget(gather(tgt, n_tgt, stride, xyz_off, label_off), 0);
    g.setInputSource(gather(src, n_src, stride, xyz_off, label_off), 0);
    if (!g.align(guess)) return APD_ERR_INVALID;
    const auto t1 = std::chrono::steady_clock::now();
    ms_out[r] = std::chrono::duration<double, std::milli>(t1 - t0).count();
  }
  return APD_OK;
}

}  // extern "C"
