// TEST INFRASTRUCTURE — stand-in for pcl::search::KdTree as the reference's headers use it. Brute force: the definition.
// [ext] PCL 1.10 kdtree_flann.hpp + FLANN 1.9.1 (L2_Simple<float>): squared distances are fp32, (dx*dx + dy*dy) + dz*dz;
//   nearestKSearch(point, k): the k nearest, ascending by distance — ties by lower index here (FLANN's tie order follows
//     its tree layout and cannot be restated; the oracle and the CUDA library use the same (d2, index) rule);
//   radiusSearch(index, radius): the squared radius handed to FLANN is static_cast<float>(radius * radius), a point is a
//     neighbour when its squared distance is STRICTLY below it (RadiusResultSet), sorted by distance.
#ifndef APDO_REF_STUB_PCL_SEARCH_KDTREE
#define APDO_REF_STUB_PCL_SEARCH_KDTREE
#include <pcl/point_types.h>
#ifdef APDO_REF_NANOFLANN
// the timed build (bench.py --impl reference): nearestKSearch through the kd-tree the reference tree vendors (nanoflann
// 1.3.2, 4DRadarSLAM/include/scan_context/nanoflann.hpp, compiled in place) instead of brute force — a kd-tree of the same
// family as the FLANN index real PCL builds; order at exact ties is the tree's
#include <nanoflann.hpp>
#endif
namespace pcl {
namespace search {
template <typename PointT>
class KdTree {
public:
  using Ptr = std::shared_ptr<KdTree<PointT>>;
  using PointCloudConstPtr = typename pcl::PointCloud<PointT>::ConstPtr;
#ifdef APDO_REF_NANOFLANN
  struct Adaptor {
    const pcl::PointCloud<PointT>* c = nullptr;
    inline size_t kdtree_get_point_count() const { return c->points.size(); }
    inline float kdtree_get_pt(const size_t idx, const size_t dim) const {
      const PointT& p = c->points[idx];
      return dim == 0 ? p.x : (dim == 1 ? p.y : p.z);
    }
    template <class BBOX>
    bool kdtree_get_bbox(BBOX&) const { return false; }
  };
  using Tree = nanoflann::KDTreeSingleIndexAdaptor<nanoflann::L2_Simple_Adaptor<float, Adaptor>, Adaptor, 3, int>;
  void setInputCloud(const PointCloudConstPtr& cloud) {
    input_ = cloud;
    adaptor_.c = cloud.get();
    tree_.reset(new Tree(3, adaptor_, nanoflann::KDTreeSingleIndexAdaptorParams(15)));
    tree_->buildIndex();
  }
  PointCloudConstPtr getInputCloud() const { return input_; }
  int nearestKSearch(const PointT& q, int k, std::vector<int>& k_indices, std::vector<float>& k_sqr_distances) const {
    const size_t kk = std::min<size_t>((size_t)k, input_->points.size());
    k_indices.resize(kk);
    k_sqr_distances.resize(kk);
    const float qq[3] = {q.x, q.y, q.z};
    const size_t found = tree_->knnSearch(qq, kk, k_indices.data(), k_sqr_distances.data());
    k_indices.resize(found);
    k_sqr_distances.resize(found);
    return (int)found;
  }
#else
  void setInputCloud(const PointCloudConstPtr& cloud) { input_ = cloud; }
  PointCloudConstPtr getInputCloud() const { return input_; }
  int nearestKSearch(const PointT& q, int k, std::vector<int>& k_indices, std::vector<float>& k_sqr_distances) const {
    std::vector<std::pair<float, int>> all(input_->points.size());
    for (std::size_t i = 0; i < input_->points.size(); i++) all[i] = std::make_pair(sqdist(q, input_->points[i]), (int)i);
    const std::size_t kk = std::min<std::size_t>((std::size_t)k, all.size());
    std::partial_sort(all.begin(), all.begin() + kk, all.end());
    k_indices.resize(kk);
    k_sqr_distances.resize(kk);
    for (std::size_t i = 0; i < kk; i++) {
      k_indices[i] = all[i].second;
      k_sqr_distances[i] = all[i].first;
    }
    return (int)kk;
  }
#endif
  int radiusSearch(int index, double radius, std::vector<int>& k_indices, std::vector<float>& k_sqr_distances, unsigned int = 0) const {
    const float r2 = static_cast<float>(radius * radius);
    const PointT& q = input_->points[(std::size_t)index];
    std::vector<std::pair<float, int>> found;
    for (std::size_t i = 0; i < input_->points.size(); i++) {
      const float d2 = sqdist(q, input_->points[i]);
      if (d2 < r2) found.emplace_back(d2, (int)i);
    }
    std::sort(found.begin(), found.end());
    k_indices.resize(found.size());
    k_sqr_distances.resize(found.size());
    for (std::size_t i = 0; i < found.size(); i++) {
      k_indices[i] = found[i].second;
      k_sqr_distances[i] = found[i].first;
    }
    return (int)found.size();
  }
private:
  static float sqdist(const PointT& q, const PointT& p) {
    const float dx = q.x - p.x, dy = q.y - p.y, dz = q.z - p.z;
    return (dx * dx + dy * dy) + dz * dz;
  }
  PointCloudConstPtr input_;
#ifdef APDO_REF_NANOFLANN
  Adaptor adaptor_;
  std::unique_ptr<Tree> tree_;
#endif
};
}  // namespace search
}  // namespace pcl
#endif
