// The search method the B200 FastAPDGICP shim installs in its pcl::Registration base (no reference counterpart: the
// reference leaves PCL's own CPU kd-tree there).
//
// Why it exists. pcl::Registration (PCL 1.10) owns a pcl::search::KdTree `tree_` over the target: align() ->
// initCompute() rebuilds it for every new target (a FLANN index, milliseconds for a 60 k-point submap, every frame the
// keyframe changes), although FastAPDGICP never searches it; and the callers reach it through the base pointer —
// getFitnessScore() (loop_detector.cpp:229-231, :299; scan_matching_odometry_nodelet.cpp:675) and
// getSearchMethodTarget()->nearestKSearch (:684) — one point at a time. This class is a pcl::search::KdTree whose
// index is the GPU grid the registration has already built over the same target. Installed with
// setSearchMethodTarget(adaptor, /*force_no_recompute=*/true), so initCompute() builds nothing on the CPU.
//
// How the point-at-a-time queries stay cheap. Both callers ask for the nearest target point of the source points
// transformed by final_transformation_, in cloud order. The first k = 1 query after an align fetches that whole
// answer in ONE GPU pass (apd_source_nearest: transformed point, neighbour index, squared distance per source point);
// every query is then matched against the batch at a cursor that walks the cloud. A query that equals the batch
// point bit for bit (the `aligned` cloud align() returned) takes the stored answer; one that differs in the last bits
// (PCL's SSE transformPointCloud associates the sum differently) takes the stored neighbour and recomputes its squared
// distance from the caller's own coordinates; anything else (other points, k > 1) is a single-query call of the GPU
// search (apd_nearest_k) — correct, but a round trip per point.
#ifndef FAST_GICP_APD_SEARCH_HPP
#define FAST_GICP_APD_SEARCH_HPP

#include <cmath>
#include <cstring>
#include <limits>

#include <pcl/search/kdtree.h>

#include <fast_gicp/gicp/apd_shim_common.hpp>

namespace fast_gicp {

// what the adaptor needs to know about the registration object it serves
template <typename PointTarget>
struct ApdSearchOwner {
  virtual ~ApdSearchOwner() {}
  virtual apd_handle* apdSearchHandle() const = 0;
  virtual unsigned long long apdSearchGeneration() const = 0;  // changes whenever the clouds or final_transformation_ do
  virtual std::size_t apdSearchSourceSize() const = 0;
  virtual const PointTarget* apdSearchTargetPoints(std::size_t* n) const = 0;
};

template <typename PointTarget>
class ApdTargetSearch : public pcl::search::KdTree<PointTarget> {
public:
  using Base = pcl::search::KdTree<PointTarget>;
  using PointCloudConstPtr = typename Base::PointCloudConstPtr;
  using Ptr = APD_SHIM_SHARED_PTR<ApdTargetSearch<PointTarget>>;

  explicit ApdTargetSearch(const ApdSearchOwner<PointTarget>* owner) : owner_(owner) {}

  // The registration's GPU grid over the same cloud is the index: nothing is built here.
  void setInputCloud(const PointCloudConstPtr& cloud, const pcl::IndicesConstPtr& = pcl::IndicesConstPtr()) override { this->input_ = cloud; }

  int nearestKSearch(const PointTarget& p, int k, std::vector<int>& k_indices, std::vector<float>& k_sqr_distances) const override {
    k_indices.assign((std::size_t)std::max(k, 0), -1);
    k_sqr_distances.assign((std::size_t)std::max(k, 0), std::numeric_limits<float>::max());
    apd_handle* h = owner_ ? owner_->apdSearchHandle() : nullptr;
    if (!h || k < 1) return 0;
    if (k == 1 && from_batch(h, p, k_indices[0], k_sqr_distances[0])) return 1;
    single_queries_++;
    std::vector<int32_t> idx((std::size_t)k);
    const int rc = apd_nearest_k(h, /*which=target*/ 1, &p.x, 1, (int32_t)sizeof(PointTarget), k, idx.data(), k_sqr_distances.data());
    if (rc != APD_OK) {
      apd_detail::check(h, rc, "nearestKSearch");
      return 0;
    }
    int found = 0;
    for (int j = 0; j < k; j++) {
      k_indices[(std::size_t)j] = idx[(std::size_t)j];
      if (idx[(std::size_t)j] >= 0) found++;
    }
    return found;
  }

  // not used on the registration path (the callers ask for nearest neighbours only)
  int radiusSearch(const PointTarget&, double, std::vector<int>& k_indices, std::vector<float>& k_sqr_distances, unsigned int = 0) const override {
    k_indices.clear();
    k_sqr_distances.clear();
    return 0;
  }

  // instrumentation: GPU passes over the whole source, queries answered from them, queries that took a round trip each
  std::size_t batchPasses() const { return batch_passes_; }
  std::size_t batchHits() const { return batch_hits_; }
  std::size_t singleQueries() const { return single_queries_; }

private:
  bool from_batch(apd_handle* h, const PointTarget& p, int& idx, float& d2) const {
    const std::size_t n = owner_->apdSearchSourceSize();
    if (n == 0) return false;
    const unsigned long long gen = owner_->apdSearchGeneration();
    if (gen != batch_gen_ || b_idx_.size() != n) {
      b_idx_.resize(n);
      b_d2_.resize(n);
      b_xyz_.resize(3 * n);
      batch_gen_ = gen;
      cursor_ = 0;
      batch_ok_ = apd_source_nearest(h, nullptr, b_idx_.data(), b_d2_.data(), b_xyz_.data(), (int32_t)n) == APD_OK;
      batch_passes_++;
    }
    if (!batch_ok_) return false;
    const std::size_t probes[2] = {cursor_ < n ? cursor_ : 0, 0};
    for (std::size_t probe : probes) {
      const float* q = &b_xyz_[3 * probe];
      if (b_idx_[probe] < 0) continue;
      if (std::memcmp(q, &p.x, 12) == 0) {  // the very point the GPU searched for
        idx = b_idx_[probe];
        d2 = b_d2_[probe];
      } else {
        const float tol = 1e-5f * (1.0f + std::fabs(q[0]) + std::fabs(q[1]) + std::fabs(q[2]));
        if (!(std::fabs(q[0] - p.x) <= tol && std::fabs(q[1] - p.y) <= tol && std::fabs(q[2] - p.z) <= tol)) continue;
        std::size_t nt = 0;
        const PointTarget* t = owner_->apdSearchTargetPoints(&nt);
        if (!t || (std::size_t)b_idx_[probe] >= nt) continue;
        const PointTarget& b = t[b_idx_[probe]];
        const float dx = p.x - b.x, dy = p.y - b.y, dz = p.z - b.z;
        idx = b_idx_[probe];
        d2 = (dx * dx + dy * dy) + dz * dz;  // FLANN L2_Simple's order, from the caller's own coordinates
      }
      cursor_ = probe + 1;
      batch_hits_++;
      return true;
    }
    return false;
  }

  const ApdSearchOwner<PointTarget>* owner_;
  mutable std::vector<int32_t> b_idx_;
  mutable std::vector<float> b_d2_, b_xyz_;
  mutable unsigned long long batch_gen_ = ~0ull;
  mutable bool batch_ok_ = false;
  mutable std::size_t cursor_ = 0, batch_passes_ = 0, batch_hits_ = 0, single_queries_ = 0;
};

}  // namespace fast_gicp

#endif
