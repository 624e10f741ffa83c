// TEST INFRASTRUCTURE — CPU oracle for the FastAPDGICP hot path.
//
// A plain C++17 restatement of the reference algorithm, each function citing
// the reference file:line it follows. It is NOT the product: only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may build or call it. The product (go-rio_b200/csrc) never includes this.
//
// HOW IT IS PINNED: the reference holds no test, golden vector or fixture for
// FastAPDGICP and its build needs Eigen / PCL / FLANN, which
// the image lacks. Its own HEADERS, however, compile where they
// lie under /root/reference against the small stand-ins of oracle/ref_stubs
// (ref_apdgicp.cpp, ref_gicp.cpp, ref_dbscan.cpp -> oracle/_ref); what that code
// computes is committed as tests/golden/{apdgicp,gicp,dbscan}_reference.npz, and
// tests/test_reference_code.py holds this restatement (and the CUDA library) to
// it: bit-exact correspondences and distances, same iteration counts, H / b / err
// within the atan2f deviation below. Also pinned against an independent
// NumPy/SciPy restatement (tests/test_oracle_numpy.py) and the property tests.
//
// Reference files restated (all under /root/reference/fast_apdgicp/include/fast_gicp):
//   gicp/impl/fast_apdgicp_impl.hpp:148-411   covariances, correspondences, linearize, compute_error
//   gicp/impl/lsq_registration_impl.hpp:55-173 optimizer control flow (defaults :11-24)
//   so3/so3.hpp:21-78                          skewd, so3_exp
//   gicp/fast_apdgicp.hpp:96-121               noise defaults
// Third-party behaviour restated from published semantics, marked [ext]:
//   PCL 1.10 pcl::Registration::align / getFitnessScore, pcl::search::KdTree
//   (xyz-only, FLANN 1.9.1 KDTreeSingleIndex + L2_Simple<float>), Eigen 3.3.7
//   JacobiSVD / inverse / LDLT / Transform products.
// Documented deviations:
//   - kNN ties are broken by (fp32 d2, lower index); FLANN's order at exact ties
//     depends on its tree layout.
//   - atan2 of the noise model is evaluated in double and rounded to float
//     (the reference calls glibc atan2f, fast_apdgicp_impl.hpp:198-199; the two
//     differ by at most one float ulp of the angle).
//   - 3x3 SVD = symmetric Jacobi eigen-solve (U = V), inverses are closed-form
//     adjugates, LDLT is a restated pivoted LDL^T: equal to Eigen up to rounding.
//   - sums run in index order (the reference's OpenMP order is not deterministic).
#pragma once
#include <cfloat>
#include <cstdint>
#include <cstdio>
#include <limits>
#include <memory>
#include <string>
#include <vector>

#include "apd_math.hpp"
#include "../include/apdgicp.h"

#ifdef _OPENMP
#include <omp.h>
#endif

namespace apdo {

struct PointXYZL {
  float x, y, z, label;
};

// fp32 squared distance exactly as FLANN L2_Simple<float> accumulates it [ext]
// (the same form is in the vendored nanoflann, 4DRadarSLAM/include/scan_context/
// nanoflann.hpp:432-440): result += diff*diff over x, y, z, no FMA contraction.
// (Built with -ffp-contract=off and without -mfma, enforced below, so every
// fp32 operation here rounds once, exactly like the reference's SSE build,
// fast_apdgicp/CMakeLists.txt:11-16.)
#if defined(__FMA__) || defined(__FAST_MATH__)
#error "the oracle must be built without FMA contraction and without fast-math"
#endif
static inline float sqdist_f32(float ax, float ay, float az, float bx, float by, float bz) {
  const float dx = ax - bx;
  const float dy = ay - by;
  const float dz = az - bz;
  return (dx * dx + dy * dy) + dz * dz;
}

struct Neighbor {
  float d2;
  int idx;
};
static inline bool nb_less(const Neighbor& a, const Neighbor& b) { return a.d2 < b.d2 || (a.d2 == b.d2 && a.idx < b.idx); }

// Exact kNN search structure over xyz [ext: pcl::search::KdTree -> FLANN].
// Result order: ascending (d2, idx). Two implementations that must agree:
//   brute force (definitional) and a kd-tree whose pruning is exact in fp32.
class Search {
 public:
  virtual ~Search() {}
  virtual void knn(float qx, float qy, float qz, int k, std::vector<Neighbor>& out) const = 0;
};

class BruteSearch : public Search {
 public:
  explicit BruteSearch(const std::vector<PointXYZL>& pts) : pts_(pts) {}
  void knn(float qx, float qy, float qz, int k, std::vector<Neighbor>& out) const override {
    out.clear();
    const int n = (int)pts_.size();
    if (k > n) k = n;  // [ext] FLANN clamps k to the cloud size
    for (int i = 0; i < n; i++) {
      Neighbor c{sqdist_f32(qx, qy, qz, pts_[i].x, pts_[i].y, pts_[i].z), i};
      if ((int)out.size() < k) {
        out.push_back(c);
        for (int j = (int)out.size() - 1; j > 0 && nb_less(out[j], out[j - 1]); j--) std::swap(out[j], out[j - 1]);
      } else if (nb_less(c, out.back())) {
        out.back() = c;
        for (int j = k - 1; j > 0 && nb_less(out[j], out[j - 1]); j--) std::swap(out[j], out[j - 1]);
      }
    }
  }

 private:
  const std::vector<PointXYZL>& pts_;
};

// kd-tree, median split on the widest axis, leaves of <= 15 points (FLANN's
// KDTreeSingleIndexParams(15) as PCL builds it [ext]). A far subtree is skipped
// only if fl(fl(q[axis]-split)^2) > worst.d2; because fp32 subtraction, squaring
// and the non-negative sum are monotone, every point in that subtree then has
// d2 >= that bound > worst, so skipping can never change the (d2, idx) result.
class KdSearch : public Search {
 public:
  explicit KdSearch(const std::vector<PointXYZL>& pts) : pts_(pts) {
    const int n = (int)pts.size();
    order_.resize(n);
    for (int i = 0; i < n; i++) order_[i] = i;
    if (n > 0) {
      nodes_.reserve(2 * (n / 8 + 1));
      build(0, n);
    }
  }
  void knn(float qx, float qy, float qz, int k, std::vector<Neighbor>& out) const override {
    out.clear();
    const int n = (int)pts_.size();
    if (k > n) k = n;
    if (k <= 0) return;
    const float q[3] = {qx, qy, qz};
    descend(0, q, k, out);
  }

 private:
  struct Node {
    int lo, hi;        // leaf: range in order_
    int left, right;   // children (-1 for leaf)
    int axis;
    float split_lo;    // max coordinate of the left subtree on axis
    float split_hi;    // min coordinate of the right subtree on axis
  };
  int build(int lo, int hi) {
    const int id = (int)nodes_.size();
    nodes_.push_back(Node{lo, hi, -1, -1, 0, 0.f, 0.f});
    if (hi - lo <= 15) return id;
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int i = lo; i < hi; i++) {
      const PointXYZL& p = pts_[order_[i]];
      const float c[3] = {p.x, p.y, p.z};
      for (int a = 0; a < 3; a++) {
        mn[a] = std::min(mn[a], c[a]);
        mx[a] = std::max(mx[a], c[a]);
      }
    }
    int axis = 0;
    for (int a = 1; a < 3; a++)
      if (mx[a] - mn[a] > mx[axis] - mn[axis]) axis = a;
    const int mid = (lo + hi) / 2;
    auto coord = [&](int i) {
      const PointXYZL& p = pts_[i];
      return axis == 0 ? p.x : (axis == 1 ? p.y : p.z);
    };
    std::nth_element(order_.begin() + lo, order_.begin() + mid, order_.begin() + hi,
                     [&](int a, int b) { return coord(a) < coord(b); });
    float slo = -FLT_MAX, shi = FLT_MAX;
    for (int i = lo; i < mid; i++) slo = std::max(slo, coord(order_[i]));
    for (int i = mid; i < hi; i++) shi = std::min(shi, coord(order_[i]));
    const int l = build(lo, mid);
    const int r = build(mid, hi);
    nodes_[id].left = l;
    nodes_[id].right = r;
    nodes_[id].axis = axis;
    nodes_[id].split_lo = slo;
    nodes_[id].split_hi = shi;
    return id;
  }
  void offer(const Neighbor& c, int k, std::vector<Neighbor>& out) const {
    if ((int)out.size() < k) {
      out.push_back(c);
      for (int j = (int)out.size() - 1; j > 0 && nb_less(out[j], out[j - 1]); j--) std::swap(out[j], out[j - 1]);
    } else if (nb_less(c, out.back())) {
      out.back() = c;
      for (int j = k - 1; j > 0 && nb_less(out[j], out[j - 1]); j--) std::swap(out[j], out[j - 1]);
    }
  }
  void descend(int id, const float q[3], int k, std::vector<Neighbor>& out) const {
    const Node& nd = nodes_[id];
    if (nd.left < 0) {
      for (int i = nd.lo; i < nd.hi; i++) {
        const int j = order_[i];
        offer(Neighbor{sqdist_f32(q[0], q[1], q[2], pts_[j].x, pts_[j].y, pts_[j].z), j}, k, out);
      }
      return;
    }
    const float qa = q[nd.axis];
    // distance from q to each child's slab along the split axis (0 if inside)
    auto slab = [&](bool left) -> float {
      if (left) {
        if (qa <= nd.split_lo) return 0.f;
        const float d = qa - nd.split_lo;
        return d * d;
      } else {
        if (qa >= nd.split_hi) return 0.f;
        const float d = qa - nd.split_hi;
        return d * d;
      }
    };
    const float bl = slab(true), br = slab(false);
    const bool left_first = bl <= br;
    const int first = left_first ? nd.left : nd.right;
    const int second = left_first ? nd.right : nd.left;
    const float b1 = left_first ? bl : br, b2 = left_first ? br : bl;
    if ((int)out.size() < k || !(b1 > out.back().d2)) descend(first, q, k, out);
    if ((int)out.size() < k || !(b2 > out.back().d2)) descend(second, q, k, out);
  }
  const std::vector<PointXYZL>& pts_;
  std::vector<int> order_;
  std::vector<Node> nodes_;
};

std::unique_ptr<Search> make_search(const std::vector<PointXYZL>& pts, int kind);  // 0 brute, 1 kd, 2 nanoflann (_ref only)

struct LmTraceRow {
  double outer, inner, y0, yi, rho, lambda, dnorm, accepted;
};

class FastAPDGICP {
 public:
  FastAPDGICP();

  apd_params params;
  int num_threads = 1;   // reference setNumThreads (fast_apdgicp_impl.hpp:34-42)
  int search_kind = 1;   // 0 brute force, 1 kd-tree, 2 nanoflann (oracle/_ref build only)

  void setInputSource(const std::vector<PointXYZL>& cloud, uint64_t key);
  void setInputTarget(const std::vector<PointXYZL>& cloud, uint64_t key);
  void swapSourceAndTarget();
  void clearSource();
  void clearTarget();

  // returns false (and sets error) on too-few points
  bool ensure_covariances();
  bool calculate_covariances(const std::vector<PointXYZL>& cloud, const Search& search, std::vector<M3>& covs,
                             std::vector<int>* neighbors);
  void update_correspondences(const M4& trans);
  double linearize(const M4& trans, double* H36, double* b6);
  double compute_error(const M4& trans);
  bool is_converged(const M4& delta) const;
  bool step_gn(M4& x0, M4& delta);
  bool step_lm(M4& x0, M4& delta);
  bool align(const float* guess_colmajor);
  double fitness(const float* T_colmajor_or_null, double max_range, int* n_in_range, double inlier_sq_thr, int* n_inliers);

  // state (reference fast_apdgicp.hpp:95-121 + pcl::Registration members [ext])
  std::vector<PointXYZL> source, target;
  uint64_t source_key = 0, target_key = 0;
  bool has_source = false, has_target = false;
  std::unique_ptr<Search> source_search, target_search;
  std::vector<M3> source_covs, target_covs;
  std::vector<int> source_neighbors, target_neighbors;  // parity hook: n*k indices
  std::vector<M3> mahalanobis;
  std::vector<int> correspondences;
  std::vector<float> sq_distances;

  // ---- FastVGICP (params.variant == APD_VARIANT_VGICP; oracle/apd_vgicp_oracle.cpp) ----
  // GaussianVoxel (fast_vgicp_voxel.hpp:57-77) of the voxel map, kept in ascending (z, y, x) order of the coordinates
  struct Voxel {
    int coord[3];
    int num_points;
    double mean[3];
    M3 cov;
  };
  std::vector<Voxel> voxels;
  bool voxelmap_valid = false;                 // voxelmap_ != nullptr
  std::vector<int> voxel_correspondences;      // [n_source * n_offsets] index into `voxels` or -1 (the reference keeps a list of pairs)
  std::vector<M3> voxel_mahalanobis;           // same slots
  int n_offsets() const;
  void create_voxelmap();
  int lookup_voxel(int cx, int cy, int cz) const;
  void vgicp_update_correspondences(const M4& trans);
  double vgicp_sums(const M4& trans, double* H36, double* b6);

  M4 final_pose_f64;           // x0 before the float cast
  float final_transformation[16];  // column-major, = float(x0) (lsq_registration_impl.hpp:78)
  double final_hessian[36];
  bool converged = false;
  int nr_iterations = 0;
  double lm_lambda = -1.0;
  std::vector<LmTraceRow> lm_trace;
  int trace_outer = 0;
  std::string error;
  // work counters for the CPU baseline report
  long n_linearize = 0, n_compute_error = 0;
};

}  // namespace apdo
