// CPU oracle, FastVGICP part — TEST INFRASTRUCTURE ONLY (see apd_oracle.hpp). A restatement of
//   /root/reference/fast_apdgicp/include/fast_gicp/gicp/impl/fast_vgicp_impl.hpp   (update_correspondences :74-118,
//       linearize :121-181, compute_error :184-205, the voxel map's life cycle :46-71)
//   /root/reference/fast_apdgicp/include/fast_gicp/gicp/fast_vgicp_voxel.hpp       (neighbor_offsets :10-44, the
//       additive / multiplicative Gaussian voxels :79-124, create_voxelmap :131-158, voxel_coord :160-162)
// FastVGICP derives from FastGICP: covariances (fast_gicp_impl.hpp:215-262, the same code as FastAPDGICP's), optimizer
// and convergence test are the ones of apd_oracle.cpp. The reference holds no golden output for this class
// (gicp_test.cpp:148-166 states a 5 cm / 1 degree bar against a file that is absent): pinned against what the
// reference's own headers compute (ref_gicp.cpp -> tests/golden/gicp_reference.npz, tests/test_reference_code.py) and
// against tests/numpy_restatement.py, like the rest of the oracle.
// [ext] Eigen 3.3.7: `Vector4d / scalar` and `Matrix4d / scalar` divide element by element; Isometry3d * Vector4d is
// ((r0 x + r1 y) + r2 z) + t w with w = 1; the 4x4 inverse of blockdiag(A, 1) is blockdiag(A^-1, 1) (closed-form adjugate
// here, as everywhere in this oracle). The reference's correspondence LIST is in OpenMP thread order; here it is a table
// in (source index, offset) order — the sums differ from any run of the reference by their rounding only.
#include <algorithm>
#include <cmath>

#include "apd_oracle.hpp"

namespace apdo {

namespace {
// neighbor_offsets (fast_vgicp_voxel.hpp:10-44), in the reference's order
const int kOff1[1][3] = {{0, 0, 0}};
const int kOff7[7][3] = {{0, 0, 0}, {1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}};
inline void offset_of(int method, int o, int out[3]) {
  if (method == APD_VOXEL_DIRECT1) { out[0] = kOff1[o][0]; out[1] = kOff1[o][1]; out[2] = kOff1[o][2]; return; }
  if (method == APD_VOXEL_DIRECT7) { out[0] = kOff7[o][0]; out[1] = kOff7[o][1]; out[2] = kOff7[o][2]; return; }
  out[0] = o / 9 - 1; out[1] = (o / 3) % 3 - 1; out[2] = o % 3 - 1;  // :36-42 loops i, j, k: Vector3i(i - 1, j - 1, k - 1)
}
// voxel_coord (:160-162): floor(x / resolution - 0.5), double arithmetic
inline int coord_of(double x, double res) { return (int)std::floor(x / res - 0.5); }
inline bool coord_less(const int a[3], const int b[3]) {
  if (a[2] != b[2]) return a[2] < b[2];
  if (a[1] != b[1]) return a[1] < b[1];
  return a[0] < b[0];
}
}  // namespace

int FastAPDGICP::n_offsets() const {
  return params.voxel_search == APD_VOXEL_DIRECT1 ? 1 : (params.voxel_search == APD_VOXEL_DIRECT7 ? 7 : 27);
}

// GaussianVoxelMap::create_voxelmap (:131-158): every target point, in index order, is appended to the voxel of its
// coordinate; then every voxel is finalised
void FastAPDGICP::create_voxelmap() {
  const double res = params.voxel_resolution;
  const bool mult = params.voxel_mode == APD_VOXEL_MULTIPLICATIVE;
  const int n = (int)target.size();
  // group the points by voxel, keeping index order within a voxel (the unordered_map of the reference only decides where
  // a voxel lives, not what it holds)
  std::vector<int> order(n);
  std::vector<int> cx(n), cy(n), cz(n);
  for (int i = 0; i < n; i++) {
    order[i] = i;
    cx[i] = coord_of((double)target[i].x, res);
    cy[i] = coord_of((double)target[i].y, res);
    cz[i] = coord_of((double)target[i].z, res);
  }
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
    const int ca[3] = {cx[a], cy[a], cz[a]}, cb[3] = {cx[b], cy[b], cz[b]};
    return coord_less(ca, cb);
  });
  voxels.clear();
  for (int s = 0; s < n;) {
    const int first = order[s];
    Voxel v;
    v.coord[0] = cx[first]; v.coord[1] = cy[first]; v.coord[2] = cz[first];
    v.num_points = 0;
    v.mean[0] = v.mean[1] = v.mean[2] = 0.0;
    for (int e = 0; e < 9; e++) v.cov.m[e] = 0.0;
    int e2 = s;
    for (; e2 < n && cx[order[e2]] == v.coord[0] && cy[order[e2]] == v.coord[1] && cz[order[e2]] == v.coord[2]; e2++) {
      const int i = order[e2];
      const double m[3] = {(double)target[i].x, (double)target[i].y, (double)target[i].z};
      v.num_points++;
      if (!mult) {  // AdditiveGaussianVoxel::append (:108-112)
        for (int a = 0; a < 3; a++) v.mean[a] += m[a];
        v.cov = add3(v.cov, target_covs[i]);
      } else {  // MultiplicativeGaussianVoxel::append (:86-93): cov += cov_^-1, mean += cov_^-1 mean_
        const M3 ci = inverse3(target_covs[i]);
        v.cov = add3(v.cov, ci);
        for (int a = 0; a < 3; a++) v.mean[a] += (ci(a, 0) * m[0] + ci(a, 1) * m[1]) + ci(a, 2) * m[2];
      }
    }
    if (!mult) {  // finalize (:114-117)
      for (int a = 0; a < 3; a++) v.mean[a] /= (double)v.num_points;
      for (int e = 0; e < 9; e++) v.cov.m[e] /= (double)v.num_points;
    } else {  // finalize (:95-101): cov = cov^-1, mean = cov mean
      v.cov = inverse3(v.cov);
      const double m[3] = {v.mean[0], v.mean[1], v.mean[2]};
      for (int a = 0; a < 3; a++) v.mean[a] = (v.cov(a, 0) * m[0] + v.cov(a, 1) * m[1]) + v.cov(a, 2) * m[2];
    }
    voxels.push_back(v);
    s = e2;
  }
  voxelmap_valid = true;
}

// lookup_voxel (:169-176)
int FastAPDGICP::lookup_voxel(int x, int y, int z) const {
  const int key[3] = {x, y, z};
  int lo = 0, hi = (int)voxels.size();
  while (lo < hi) {
    const int mid = (lo + hi) / 2;
    if (coord_less(voxels[mid].coord, key)) lo = mid + 1;
    else hi = mid;
  }
  if (lo < (int)voxels.size() && voxels[lo].coord[0] == x && voxels[lo].coord[1] == y && voxels[lo].coord[2] == z) return lo;
  return -1;
}

// FastVGICP::update_correspondences (fast_vgicp_impl.hpp:74-118)
void FastAPDGICP::vgicp_update_correspondences(const M4& trans) {
  const int n = (int)source.size(), no = n_offsets();
  const double res = params.voxel_resolution;
  voxel_correspondences.assign((size_t)n * no, -1);
  voxel_mahalanobis.assign((size_t)n * no, M3());
  M3 R3;
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) R3(r, c) = trans(r, c);
  const M3 R3t = transpose3(R3);
  for (int i = 0; i < n; i++) {
    const double a[3] = {(double)source[i].x, (double)source[i].y, (double)source[i].z};
    double tA[3];
    for (int r = 0; r < 3; r++) tA[r] = ((trans(r, 0) * a[0] + trans(r, 1) * a[1]) + trans(r, 2) * a[2]) + trans(r, 3);  // :87
    const int c0[3] = {coord_of(tA[0], res), coord_of(tA[1], res), coord_of(tA[2], res)};  // :88
    for (int o = 0; o < no; o++) {
      int off[3];
      offset_of(params.voxel_search, o, off);
      const int v = lookup_voxel(c0[0] + off[0], c0[1] + off[1], c0[2] + off[2]);  // :91
      if (v < 0) continue;
      voxel_correspondences[(size_t)i * no + o] = v;
      // :111-116 RCR = cov_B + T cov_A T^T (RCR(3,3) = 1), inverse, (3,3) = 0
      voxel_mahalanobis[(size_t)i * no + o] = inverse3(add3(voxels[v].cov, mul3(mul3(R3, source_covs[i]), R3t)));
    }
  }
}

// the loops of linearize (:141-169) / compute_error (:186-202) over the correspondence table
double FastAPDGICP::vgicp_sums(const M4& trans, double* H36, double* b6) {
  const bool want_hb = (H36 != nullptr && b6 != nullptr);
  if (want_hb) {
    for (int e = 0; e < 36; e++) H36[e] = 0.0;
    for (int e = 0; e < 6; e++) b6[e] = 0.0;
  }
  const int n = (int)source.size(), no = n_offsets();
  double sum_errors = 0.0;
  if (voxel_correspondences.size() != (size_t)n * no) return 0.0;
  for (int i = 0; i < n; i++) {
    const double a[3] = {(double)source[i].x, (double)source[i].y, (double)source[i].z};
    double tA[3];
    for (int r = 0; r < 3; r++) tA[r] = ((trans(r, 0) * a[0] + trans(r, 1) * a[1]) + trans(r, 2) * a[2]) + trans(r, 3);  // :151
    for (int o = 0; o < no; o++) {
      const int v = voxel_correspondences[(size_t)i * no + o];
      if (v < 0) continue;
      const Voxel& vx = voxels[v];
      const M3& M = voxel_mahalanobis[(size_t)i * no + o];
      double e[3];
      for (int r = 0; r < 3; r++) e[r] = vx.mean[r] - tA[r];  // :152
      const double w = std::sqrt((double)vx.num_points);       // :154
      double Me[3];
      for (int r = 0; r < 3; r++) Me[r] = (M(r, 0) * e[0] + M(r, 1) * e[1]) + M(r, 2) * e[2];
      sum_errors += w * ((e[0] * Me[0] + e[1] * Me[1]) + e[2] * Me[2]);  // :155
      if (!want_hb) continue;
      // :161-167 J = [skew(T a), -I]; Hi = w J^T M J, bi = w J^T M e
      double J[3][6] = {{0}};
      J[0][1] = -tA[2]; J[0][2] = tA[1];
      J[1][0] = tA[2];  J[1][2] = -tA[0];
      J[2][0] = -tA[1]; J[2][1] = tA[0];
      J[0][3] = J[1][4] = J[2][5] = -1.0;
      double MJ[3][6];
      for (int r = 0; r < 3; r++)
        for (int c = 0; c < 6; c++) MJ[r][c] = (M(r, 0) * J[0][c] + M(r, 1) * J[1][c]) + M(r, 2) * J[2][c];
      for (int r = 0; r < 6; r++) {
        for (int c = 0; c < 6; c++) H36[r * 6 + c] += w * ((J[0][r] * MJ[0][c] + J[1][r] * MJ[1][c]) + J[2][r] * MJ[2][c]);
        b6[r] += w * ((J[0][r] * Me[0] + J[1][r] * Me[1]) + J[2][r] * Me[2]);
      }
    }
  }
  return sum_errors;
}

}  // namespace apdo
