// TEST INFRASTRUCTURE — CPU oracle for the FastAPDGICP hot path (see apd_oracle.hpp).
#include "apd_oracle.hpp"

#include <cassert>
#include <cmath>

#ifdef APDO_USE_NANOFLANN
// oracle/_ref build only: the kd-tree the reference tree vendors (nanoflann
// 1.3.2, 4DRadarSLAM/include/scan_context/nanoflann.hpp), included from where it
// lies under /root/reference at build time (never copied), standing in for
// PCL's FLANN KDTreeSingleIndex in the timed CPU baseline.
#include <nanoflann.hpp>
#endif

namespace apdo {

#ifdef APDO_USE_NANOFLANN
namespace {
struct CloudAdaptor {
  const std::vector<PointXYZL>& pts;
  explicit CloudAdaptor(const std::vector<PointXYZL>& p) : pts(p) {}
  inline size_t kdtree_get_point_count() const { return pts.size(); }
  inline float kdtree_get_pt(const size_t idx, const size_t dim) const {
    return dim == 0 ? pts[idx].x : (dim == 1 ? pts[idx].y : pts[idx].z);
  }
  template <class BBOX>
  bool kdtree_get_bbox(BBOX&) const { return false; }
};
class NanoSearch : public Search {
 public:
  using Tree = nanoflann::KDTreeSingleIndexAdaptor<nanoflann::L2_Simple_Adaptor<float, CloudAdaptor>, CloudAdaptor, 3, int>;
  explicit NanoSearch(const std::vector<PointXYZL>& pts)
      : adaptor_(pts), tree_(3, adaptor_, nanoflann::KDTreeSingleIndexAdaptorParams(15)) {
    tree_.buildIndex();
  }
  void knn(float qx, float qy, float qz, int k, std::vector<Neighbor>& out) const override {
    const int n = (int)adaptor_.pts.size();
    if (k > n) k = n;
    int idx[64];
    float d2[64];
    std::vector<int> vi;
    std::vector<float> vd;
    int* pi = idx;
    float* pd = d2;
    if (k > 64) {
      vi.resize(k);
      vd.resize(k);
      pi = vi.data();
      pd = vd.data();
    }
    const float q[3] = {qx, qy, qz};
    const size_t found = tree_.knnSearch(q, (size_t)k, pi, pd);
    out.resize(found);
    for (size_t i = 0; i < found; i++) out[i] = Neighbor{pd[i], pi[i]};
  }

 private:
  CloudAdaptor adaptor_;
  Tree tree_;
};
}  // namespace
#endif

std::unique_ptr<Search> make_search(const std::vector<PointXYZL>& pts, int kind) {
  if (kind == 0) return std::unique_ptr<Search>(new BruteSearch(pts));
#ifdef APDO_USE_NANOFLANN
  if (kind == 2) return std::unique_ptr<Search>(new NanoSearch(pts));
#endif
  return std::unique_ptr<Search>(new KdSearch(pts));
}

// reference FastAPDGICP::FastAPDGICP (fast_apdgicp_impl.hpp:14-28) and
// LsqRegistration::LsqRegistration (lsq_registration_impl.hpp:11-24)
FastAPDGICP::FastAPDGICP() {
  params.k_correspondences = 20;
  params.regularization = APD_REG_PLANE;
  params.max_correspondence_distance = (double)std::numeric_limits<float>::max();
  params.dist_var = 0.86;
  params.azimuth_var = 0.5;
  params.elevation_var = 1.0;
  params.max_iterations = 64;
  params.optimizer = APD_OPT_LEVENBERG_MARQUARDT;
  params.rotation_epsilon = 2e-3;
  params.transformation_epsilon = 5e-4;
  params.lm_max_iterations = 10;
  params.lm_debug_print = 0;
  params.lm_init_lambda_factor = 1e-9;
  params.maha_fp64 = 1;
  params.host_loop = 0;
  params.variant = APD_VARIANT_APDGICP;
  params.voxel_search = APD_VOXEL_DIRECT1;   // fast_vgicp_impl.hpp:22-24
  params.voxel_resolution = 1.0;
  params.voxel_mode = APD_VOXEL_ADDITIVE;
  params.reserved_ = 0;
  final_pose_f64 = M4::identity();
  for (int i = 0; i < 16; i++) final_transformation[i] = (i % 5 == 0) ? 1.f : 0.f;
  for (int i = 0; i < 36; i++) final_hessian[i] = (i % 7 == 0) ? 1.0 : 0.0;  // final_hessian_.setIdentity() (:23)
}

// fast_apdgicp_impl.hpp:115-124
void FastAPDGICP::setInputSource(const std::vector<PointXYZL>& cloud, uint64_t key) {
  if (has_source && key != 0 && key == source_key) return;
  source = cloud;
  source_key = key;
  has_source = true;
  source_search = make_search(source, search_kind);
  source_covs.clear();
  source_neighbors.clear();
}
// fast_apdgicp_impl.hpp:127-135
void FastAPDGICP::setInputTarget(const std::vector<PointXYZL>& cloud, uint64_t key) {
  if (has_target && key != 0 && key == target_key) return;
  target = cloud;
  target_key = key;
  has_target = true;
  target_search = make_search(target, search_kind);
  target_covs.clear();
  target_neighbors.clear();
  voxelmap_valid = false;  // FastVGICP::setInputTarget (fast_vgicp_impl.hpp:57-64)
}
// fast_apdgicp_impl.hpp:89-98
void FastAPDGICP::swapSourceAndTarget() {
  source.swap(target);
  std::swap(source_key, target_key);
  std::swap(has_source, has_target);
  // the Search objects hold references to the vectors, so rebuild them
  source_search = has_source ? make_search(source, search_kind) : nullptr;
  target_search = has_target ? make_search(target, search_kind) : nullptr;
  source_covs.swap(target_covs);
  source_neighbors.swap(target_neighbors);
  correspondences.clear();
  sq_distances.clear();
  voxelmap_valid = false;  // FastVGICP::swapSourceAndTarget (fast_vgicp_impl.hpp:46-54)
  voxel_correspondences.clear();
  voxel_mahalanobis.clear();
}
// fast_apdgicp_impl.hpp:101-112
void FastAPDGICP::clearSource() {
  source.clear();
  has_source = false;
  source_key = 0;
  source_search.reset();
  source_covs.clear();
  source_neighbors.clear();
}
void FastAPDGICP::clearTarget() {
  voxelmap_valid = false;
  target.clear();
  has_target = false;
  target_key = 0;
  target_search.reset();
  target_covs.clear();
  target_neighbors.clear();
}

// fast_apdgicp_impl.hpp:149-154
bool FastAPDGICP::ensure_covariances() {
  if (!has_source || !has_target) {
    error = "source or target cloud not set";
    return false;
  }
  if (source_covs.size() != source.size()) {
    if (!calculate_covariances(source, *source_search, source_covs, &source_neighbors)) return false;
  }
  if (target_covs.size() != target.size()) {
    if (!calculate_covariances(target, *target_search, target_covs, &target_neighbors)) return false;
  }
  return true;
}

// fast_apdgicp_impl.hpp:351-411
bool FastAPDGICP::calculate_covariances(const std::vector<PointXYZL>& cloud, const Search& search, std::vector<M3>& covs,
                                        std::vector<int>* neighbors) {
  const int n = (int)cloud.size();
  const int k = params.k_correspondences;
  if (n < k || k < 1) {
    // reference: fewer than k neighbours leaves columns of the 4xk matrix
    // uninitialised (:366-369) — undefined behaviour; the oracle defines an error.
    error = "cloud has fewer points than k_correspondences";
    return false;
  }
  covs.assign(n, zero3());
  if (neighbors) neighbors->assign((size_t)n * k, -1);
  const int reg = params.regularization;

#pragma omp parallel for num_threads(num_threads) schedule(guided, 8)
  for (int i = 0; i < n; i++) {
    std::vector<Neighbor> nb;
    search.knn(cloud[i].x, cloud[i].y, cloud[i].z, k, nb);  // :364

    // :366-369 neighbours as double columns (row 3 is all ones and drops out)
    double nx[64], ny[64], nz[64];
    std::vector<double> hx, hy, hz;
    double *px = nx, *py = ny, *pz = nz;
    if (k > 64) {
      hx.resize(k); hy.resize(k); hz.resize(k);
      px = hx.data(); py = hy.data(); pz = hz.data();
    }
    for (int j = 0; j < k; j++) {
      const PointXYZL& p = cloud[nb[j].idx];
      px[j] = (double)p.x;
      py[j] = (double)p.y;
      pz[j] = (double)p.z;
      if (neighbors) (*neighbors)[(size_t)i * k + j] = nb[j].idx;
    }
    // :371 subtract the row mean
    double mx = 0, my = 0, mz = 0;
    for (int j = 0; j < k; j++) {
      mx += px[j];
      my += py[j];
      mz += pz[j];
    }
    mx /= k;
    my /= k;
    mz /= k;
    // :372 cov = N N^T / k
    M3 cov = zero3();
    for (int j = 0; j < k; j++) {
      const double d[3] = {px[j] - mx, py[j] - my, pz[j] - mz};
      for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) cov(r, c) += d[r] * d[c];
    }
    for (int e = 0; e < 9; e++) cov.m[e] /= k;

    if (reg == APD_REG_NONE) {  // :374-376
      covs[i] = cov;
    } else if (reg == APD_REG_FROBENIUS) {  // :377-383
      const double lambda = 1e-3;
      M3 C = cov;
      C(0, 0) += lambda;
      C(1, 1) += lambda;
      C(2, 2) += lambda;
      M3 C_inv = inverse3(C);
      const double nrm = frobenius3(C_inv);
      M3 scaled;
      for (int e = 0; e < 9; e++) scaled.m[e] = C_inv.m[e] / nrm;
      covs[i] = inverse3(scaled);
    } else {  // :384-407
      double sv[3];
      M3 U, V;
      svd3_sym(cov, sv, U, V);
      double values[3];
      switch (reg) {
        default:
        case APD_REG_PLANE:  // :392-394
          values[0] = 1.0; values[1] = 1.0; values[2] = 1e-3;
          break;
        case APD_REG_MIN_EIG:  // :395-397
          for (int e = 0; e < 3; e++) values[e] = std::max(sv[e], 1e-3);
          break;
        case APD_REG_NORMALIZED_MIN_EIG:  // :398-401
          for (int e = 0; e < 3; e++) values[e] = std::max(sv[e] / sv[0], 1e-3);
          break;
      }
      covs[i] = udvt(U, values, V);  // :405
    }
  }
  return true;
}

namespace {
inline void m4_to_f32(const M4& t, float f[16]) {
  for (int i = 0; i < 16; i++) f[i] = (float)t.m[i];
}
// Eigen Isometry3f * Vector4f with w = 1 [ext]: rows of the 3x4 affine part,
// sum over the inner index in order, separate mul and add (SSE, no FMA).
inline void transform_f32(const float t[16] /*row-major*/, float x, float y, float z, float& ox, float& oy, float& oz) {
  ox = ((t[0] * x + t[1] * y) + t[2] * z) + t[3];
  oy = ((t[4] * x + t[5] * y) + t[6] * z) + t[7];
  oz = ((t[8] * x + t[9] * y) + t[10] * z) + t[11];
}
}  // namespace

// fast_apdgicp_impl.hpp:160-220
void FastAPDGICP::update_correspondences(const M4& trans) {
  assert(source_covs.size() == source.size());
  assert(target_covs.size() == target.size());
  float tf[16];
  m4_to_f32(trans, tf);  // :164 Isometry3f trans_f = trans.cast<float>()

  const int n = (int)source.size();
  correspondences.resize(n);
  sq_distances.resize(n);
  mahalanobis.resize(n);

  const double thr = params.max_correspondence_distance;
  const double thr_sq = thr * thr;  // :183 double product (corr_dist_threshold_ is double in PCL [ext])
  const double dv = params.dist_var, av = params.azimuth_var, ev = params.elevation_var;
  M3 R3;  // rotation block of trans
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) R3(r, c) = trans(r, c);
  const M3 R3t = transpose3(R3);

#pragma omp parallel for num_threads(num_threads) schedule(guided, 8)
  for (int i = 0; i < n; i++) {
    float px, py, pz;
    transform_f32(tf, source[i].x, source[i].y, source[i].z, px, py, pz);  // :176

    std::vector<Neighbor> nb;
    target_search->knn(px, py, pz, 1, nb);  // :178
    sq_distances[i] = nb[0].d2;             // :180
    correspondences[i] = ((double)nb[0].d2 < thr_sq) ? nb[0].idx : -1;  // :183
    if (correspondences[i] < 0) {
      continue;
    }
    const int target_index = correspondences[i];
    const M3& cov_A = source_covs[i];
    const M3& cov_B = target_covs[target_index];
    if (params.variant == APD_VARIANT_GICP) {
      // FastGICP (fast_gicp_impl.hpp:157-161): RCR = cov_B + T cov_A T^T, no noise term
      mahalanobis[i] = inverse3(add3(cov_B, mul3(mul3(R3, cov_A), R3t)));
      continue;
    }

    // :194-199 radar noise at the transformed source point
    const double dist = std::sqrt((double)px * (double)px + (double)py * (double)py + (double)pz * (double)pz);
    const double s_x = dist * dv / 400;
    const double s_y = dist * std::sin(av / 180 * M_PI);
    const double s_z = dist * std::sin(ev / 180 * M_PI);
    // `using namespace std` + float arguments select the float overloads of
    // sqrt / atan2 in the reference (:198-199); see the deviation note on atan2.
    const float rho_xy = std::sqrt(px * px + py * py);
    const double elevation = (double)(float)std::atan2((double)rho_xy, (double)pz);
    const double azimuth = (double)(float)std::atan2((double)py, (double)px);
    // :200-203 R = AngleAxis(azimuth, Z) * AngleAxis(elevation, Y): Eigen forms
    // the two quaternions, multiplies them and converts to a matrix [ext].
    const double cz = std::cos(azimuth * 0.5), sz = std::sin(azimuth * 0.5);
    const double cy = std::cos(elevation * 0.5), sy = std::sin(elevation * 0.5);
    const M3 R = quat_to_rot(cz * cy, -(sz * sy), cz * sy, sz * cy);
    // :204-210 A = R*S, cov_r = A*A^T
    const double s[3] = {s_x, s_y, s_z};
    M3 A;
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) A(r, c) = R(r, c) * s[c];
    const M3 cov_r = mul3(A, transpose3(A));

    // :213-215 RCR = (cov_B + cov_dist) + T (cov_A + cov_dist) T^T (3x3 block;
    // the 4th row/column of every term is zero and RCR(3,3) is set to 1)
    const M3 RCR = add3(add3(cov_B, cov_r), mul3(mul3(R3, add3(cov_A, cov_r)), R3t));
    // :217-218 4x4 inverse of blockdiag(RCR3, 1) = blockdiag(RCR3^-1, 1); (3,3) := 0
    mahalanobis[i] = inverse3(RCR);
  }
}

namespace {
// per-point geometric weight: sigma3/sigma1 of JacobiSVD(cov_A) (:266-269, :330-333)
inline double geo_weight_of(const M3& cov_A) {
  double sv[3];
  M3 U, V;
  svd3_sym(cov_A, sv, U, V);
  return sv[2] / sv[0];  // values = sv / sv.maxCoeff(); values(2)
}

struct Contribution {
  double err;
  double H[36];
  double b[6];
};

// one iteration body of the loops at :248-295 / :314-343
inline bool point_terms(const FastAPDGICP& g, const M4& trans, int i, bool want_hb, Contribution& out) {
  const int target_index = g.correspondences[i];
  if (target_index < 0) return false;
  const PointXYZL& a = g.source[i];
  const PointXYZL& bpt = g.target[target_index];
  const double mean_A[3] = {(double)a.x, (double)a.y, (double)a.z};
  const double mean_B[3] = {(double)bpt.x, (double)bpt.y, (double)bpt.z};
  // :262-263 transed_mean_A = trans * mean_A ; error = mean_B - transed_mean_A
  double tA[3], e[3];
  for (int r = 0; r < 3; r++) {
    tA[r] = ((trans(r, 0) * mean_A[0] + trans(r, 1) * mean_A[1]) + trans(r, 2) * mean_A[2]) + trans(r, 3);
    e[r] = mean_B[r] - tA[r];
  }
  const bool gicp = g.params.variant == APD_VARIANT_GICP;  // FastGICP: sum_errors += e^T M e (fast_gicp_impl.hpp:205,:255)
  const double geo_weight = gicp ? 0.0 : geo_weight_of(g.source_covs[i]);  // :266-269 (recomputed per call, as the reference does)
  double cl_weight = 0.0;
  if (!gicp && bpt.label == a.label) cl_weight = 1.0 / (double)g.correspondences.size();  // :271-273
  const M3& M = g.mahalanobis[i];
  double Me[3];
  for (int r = 0; r < 3; r++) Me[r] = (M(r, 0) * e[0] + M(r, 1) * e[1]) + M(r, 2) * e[2];
  const double q = (e[0] * Me[0] + e[1] * Me[1]) + e[2] * Me[2];
  out.err = gicp ? q : (1.0 + geo_weight + cl_weight) * q;  // :276
  if (!want_hb) return true;
  // :284-287 J = [skew(T a), -I] (3x6; the 4th row is zero)
  double J[3][6] = {{0}};
  J[0][1] = -tA[2]; J[0][2] = tA[1];
  J[1][0] = tA[2];  J[1][2] = -tA[0];
  J[2][0] = -tA[1]; J[2][1] = tA[0];
  J[0][3] = J[1][4] = J[2][5] = -1.0;
  // :289-290 H = J^T M J, b = J^T M e
  double MJ[3][6];
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 6; c++) MJ[r][c] = (M(r, 0) * J[0][c] + M(r, 1) * J[1][c]) + M(r, 2) * J[2][c];
  for (int r = 0; r < 6; r++) {
    for (int c = 0; c < 6; c++) out.H[r * 6 + c] = (J[0][r] * MJ[0][c] + J[1][r] * MJ[1][c]) + J[2][r] * MJ[2][c];
    out.b[r] = (J[0][r] * Me[0] + J[1][r] * Me[1]) + J[2][r] * Me[2];
  }
  return true;
}
}  // namespace

// fast_apdgicp_impl.hpp:224-307
double FastAPDGICP::linearize(const M4& trans, double* H36, double* b6) {
  if (params.variant == APD_VARIANT_VGICP) {  // FastVGICP::linearize (fast_vgicp_impl.hpp:121-181)
    if (!voxelmap_valid) create_voxelmap();   // :126-129
    vgicp_update_correspondences(trans);      // :131
    n_linearize++;
    return vgicp_sums(trans, H36, b6);
  }
  update_correspondences(trans);  // :226
  n_linearize++;
  const int n = (int)source.size();
  const bool want_hb = (H36 != nullptr && b6 != nullptr);
  double sum_errors = 0.0;
  const int nt = std::max(1, num_threads);
  std::vector<double> Hs((size_t)nt * 36, 0.0), bs((size_t)nt * 6, 0.0);  // :229-234

#pragma omp parallel for num_threads(num_threads) reduction(+ : sum_errors) schedule(guided, 8)
  for (int i = 0; i < n; i++) {
    Contribution c;
    if (!point_terms(*this, trans, i, want_hb, c)) continue;
    sum_errors += c.err;
    if (!want_hb) continue;
#ifdef _OPENMP
    const int t = omp_get_thread_num();
#else
    const int t = 0;
#endif
    for (int e = 0; e < 36; e++) Hs[(size_t)t * 36 + e] += c.H[e];  // :292-293
    for (int e = 0; e < 6; e++) bs[(size_t)t * 6 + e] += c.b[e];
  }
  if (want_hb) {  // :297-304
    for (int e = 0; e < 36; e++) H36[e] = 0.0;
    for (int e = 0; e < 6; e++) b6[e] = 0.0;
    for (int t = 0; t < nt; t++) {
      for (int e = 0; e < 36; e++) H36[e] += Hs[(size_t)t * 36 + e];
      for (int e = 0; e < 6; e++) b6[e] += bs[(size_t)t * 6 + e];
    }
  }
  return sum_errors;
}

// fast_apdgicp_impl.hpp:310-346 — stale correspondences_ / mahalanobis_, trial transform in e only
double FastAPDGICP::compute_error(const M4& trans) {
  if (params.variant == APD_VARIANT_VGICP) {  // FastVGICP::compute_error (fast_vgicp_impl.hpp:184-205)
    n_compute_error++;
    return vgicp_sums(trans, nullptr, nullptr);
  }
  n_compute_error++;
  const int n = (int)source.size();
  double sum_errors = 0.0;
#pragma omp parallel for num_threads(num_threads) reduction(+ : sum_errors) schedule(guided, 8)
  for (int i = 0; i < n; i++) {
    Contribution c;
    if (!point_terms(*this, trans, i, false, c)) continue;
    sum_errors += c.err;
  }
  return sum_errors;
}

// lsq_registration_impl.hpp:83-92
bool FastAPDGICP::is_converged(const M4& delta) const {
  double rmax = 0.0, tmax = 0.0;
  for (int r = 0; r < 3; r++) {
    for (int c = 0; c < 3; c++) {
      const double v = 1.0 / params.rotation_epsilon * std::fabs(delta(r, c) - (r == c ? 1.0 : 0.0));
      rmax = std::max(rmax, v);
    }
    tmax = std::max(tmax, 1.0 / params.transformation_epsilon * std::fabs(delta(r, 3)));
  }
  return std::max(rmax, tmax) < 1;
}

namespace {
inline M4 delta_from(const double d[6]) {
  M4 delta = M4::identity();
  const M3 R = so3_exp_rot(d);  // so3_exp(d.head<3>()).toRotationMatrix()
  for (int r = 0; r < 3; r++) {
    for (int c = 0; c < 3; c++) delta(r, c) = R(r, c);
    delta(r, 3) = d[3 + r];
  }
  return delta;
}
}  // namespace

// lsq_registration_impl.hpp:107-123
bool FastAPDGICP::step_gn(M4& x0, M4& delta) {
  double H[36], b[6];
  const double y0 = linearize(x0, H, b);
  double nb[6], d[6];
  for (int i = 0; i < 6; i++) nb[i] = -b[i];
  ldlt_solve6(H, nb, d);
  delta = delta_from(d);
  x0 = mul_isometry(delta, x0);
  std::memcpy(final_hessian, H, sizeof(H));
  double dn = 0;
  for (int i = 0; i < 6; i++) dn += d[i] * d[i];
  lm_trace.push_back(LmTraceRow{(double)trace_outer, 0.0, y0, y0, 0.0, 0.0, std::sqrt(dn), 1.0});
  return true;
}

// lsq_registration_impl.hpp:127-173
bool FastAPDGICP::step_lm(M4& x0, M4& delta) {
  double H[36], b[6];
  const double y0 = linearize(x0, H, b);  // :130
  if (lm_lambda < 0.0) {                  // :131-133
    double mx = 0.0;
    for (int i = 0; i < 6; i++) mx = std::max(mx, std::fabs(H[i * 6 + i]));
    lm_lambda = params.lm_init_lambda_factor * mx;
  }
  double nu = 2.0;
  for (int i = 0; i < params.lm_max_iterations; i++) {  // :136
    double Hl[36], nb[6], d[6];
    std::memcpy(Hl, H, sizeof(H));
    for (int j = 0; j < 6; j++) {
      Hl[j * 6 + j] += lm_lambda;
      nb[j] = -b[j];
    }
    ldlt_solve6(Hl, nb, d);  // :137-138
    delta = delta_from(d);   // :140-142
    const M4 xi = mul_isometry(delta, x0);  // :144
    const double yi = compute_error(xi);    // :145
    double denom = 0.0, dn = 0.0;
    for (int j = 0; j < 6; j++) {
      denom += d[j] * (lm_lambda * d[j] - b[j]);
      dn += d[j] * d[j];
    }
    const double rho = (y0 - yi) / denom;  // :146
    if (params.lm_debug_print) {
      if (i == 0) std::printf("--- LM optimization ---\n%5s %15s %15s %15s %15s %15s %5s\n", "i", "y0", "yi", "rho", "lambda", "|delta|", "dec");
      std::printf("%5d %15g %15g %15g %15g %15g %5c\n", i, y0, yi, rho, lm_lambda, std::sqrt(dn), rho > 0.0 ? 'x' : ' ');
    }
    lm_trace.push_back(LmTraceRow{(double)trace_outer, (double)i, y0, yi, rho, lm_lambda, std::sqrt(dn), rho < 0 ? 0.0 : 1.0});
    if (rho < 0) {  // :156-164 (NaN rho falls through to "accept", as in the reference)
      if (is_converged(delta)) {
        return true;
      }
      lm_lambda = nu * lm_lambda;
      nu = 2 * nu;
      continue;
    }
    x0 = xi;  // :166-169
    lm_lambda = lm_lambda * std::max(1.0 / 3.0, 1 - std::pow(2 * rho - 1, 3));
    std::memcpy(final_hessian, H, sizeof(H));
    return true;
  }
  return false;  // :172
}

// pcl::Registration::align(output, guess) [ext] -> FastAPDGICP::computeTransformation
// (fast_apdgicp_impl.hpp:148-157) -> LsqRegistration::computeTransformation
// (lsq_registration_impl.hpp:55-80)
bool FastAPDGICP::align(const float* guess_colmajor) {
  voxelmap_valid = false;  // FastVGICP::computeTransformation (fast_vgicp_impl.hpp:66-71)
  if (!ensure_covariances()) return false;
  M4 x0 = M4::identity();  // :56 Isometry3d(guess.cast<double>())
  if (guess_colmajor) {
    for (int r = 0; r < 4; r++)
      for (int c = 0; c < 4; c++) x0(r, c) = (double)guess_colmajor[c * 4 + r];
  }
  lm_lambda = -1.0;   // :58
  converged = false;  // :59
  lm_trace.clear();
  nr_iterations = 0;
  for (int i = 0; i < params.max_iterations && !converged; i++) {  // :67
    nr_iterations = i;                                              // :68
    trace_outer = i;
    M4 delta;
    const bool ok = (params.optimizer == APD_OPT_GAUSS_NEWTON) ? step_gn(x0, delta) : step_lm(x0, delta);
    if (!ok) {
      std::fprintf(stderr, "lm not converged!!\n");  // :72
      break;
    }
    converged = is_converged(delta);  // :75
  }
  final_pose_f64 = x0;
  for (int r = 0; r < 4; r++)
    for (int c = 0; c < 4; c++) final_transformation[c * 4 + r] = (float)x0(r, c);  // :78
  return true;
}

// pcl::Registration::getFitnessScore(max_range) [ext, PCL 1.10 registration.hpp]:
// transform the input by final_transformation_ (float), 1-NN in the target,
// sum d2 (double) over the points with d2 <= max_range, divide by their count;
// DBL_MAX if there are none.
double FastAPDGICP::fitness(const float* T_colmajor_or_null, double max_range, int* n_in_range, double inlier_sq_thr,
                            int* n_inliers) {
  const float* T = T_colmajor_or_null ? T_colmajor_or_null : final_transformation;
  float tf[16];
  for (int r = 0; r < 4; r++)
    for (int c = 0; c < 4; c++) tf[r * 4 + c] = T[c * 4 + r];
  double sum = 0.0;
  int nr = 0, inl = 0;
  std::vector<Neighbor> nb;
  for (size_t i = 0; i < source.size(); i++) {
    float px, py, pz;
    transform_f32(tf, source[i].x, source[i].y, source[i].z, px, py, pz);
    target_search->knn(px, py, pz, 1, nb);
    if ((double)nb[0].d2 <= max_range) {
      sum += (double)nb[0].d2;
      nr++;
    }
    if ((double)nb[0].d2 < inlier_sq_thr) inl++;
  }
  if (n_in_range) *n_in_range = nr;
  if (n_inliers) *n_inliers = inl;
  return nr > 0 ? sum / nr : std::numeric_limits<double>::max();
}

}  // namespace apdo
