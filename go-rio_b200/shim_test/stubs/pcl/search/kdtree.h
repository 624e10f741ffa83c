// MINIMAL STAND-IN for <pcl/search/kdtree.h> (PCL 1.10): the interface of pcl::search::Search / pcl::search::KdTree that
// pcl::Registration holds as tree_ and that callers reach through getSearchMethodTarget(). "Building the tree" is
// counted (a real one is a FLANN index over the cloud) so that a test can see whether align() still builds one per target;
// queries are brute force.
#ifndef APD_STUB_PCL_SEARCH_KDTREE
#define APD_STUB_PCL_SEARCH_KDTREE
#include <cfloat>
#include <memory>
#include <vector>
#include <pcl/point_cloud.h>
namespace pcl {
using IndicesConstPtr = std::shared_ptr<const std::vector<int>>;
namespace search {
template <typename PointT>
class Search {
public:
  using PointCloud = pcl::PointCloud<PointT>;
  using PointCloudConstPtr = typename PointCloud::ConstPtr;
  using Ptr = std::shared_ptr<Search<PointT>>;
  virtual ~Search() {}
  virtual void setInputCloud(const PointCloudConstPtr& cloud, const IndicesConstPtr& = IndicesConstPtr()) { input_ = cloud; }
  virtual PointCloudConstPtr getInputCloud() const { return input_; }
  virtual int nearestKSearch(const PointT& point, int k, std::vector<int>& k_indices, std::vector<float>& k_sqr_distances) const = 0;
  virtual int radiusSearch(const PointT& point, double radius, std::vector<int>& k_indices, std::vector<float>& k_sqr_distances,
                           unsigned int max_nn = 0) const = 0;
protected:
  PointCloudConstPtr input_;
};
template <typename PointT>
class KdTree : public Search<PointT> {
public:
  using PointCloudConstPtr = typename Search<PointT>::PointCloudConstPtr;
  using Ptr = std::shared_ptr<KdTree<PointT>>;
  using ConstPtr = std::shared_ptr<const KdTree<PointT>>;
  static int& builds() { static int n = 0; return n; }  // stub only: CPU tree builds so far
  void setInputCloud(const PointCloudConstPtr& cloud, const IndicesConstPtr& = IndicesConstPtr()) override {
    this->input_ = cloud;
    builds()++;
  }
  int nearestKSearch(const PointT& p, int k, std::vector<int>& idx, std::vector<float>& d2) const override {
    idx.assign(k, -1);
    d2.assign(k, FLT_MAX);
    if (!this->input_) return 0;
    int found = 0;
    for (std::size_t i = 0; i < this->input_->points.size(); i++) {
      const PointT& q = this->input_->points[i];
      const float dx = p.x - q.x, dy = p.y - q.y, dz = p.z - q.z;
      const float d = (dx * dx + dy * dy) + dz * dz;
      int j = k - 1;
      if (!(d < d2[j])) continue;
      while (j > 0 && d < d2[j - 1]) { d2[j] = d2[j - 1]; idx[j] = idx[j - 1]; j--; }
      d2[j] = d; idx[j] = (int)i;
      if (found < k) found++;
    }
    return found;
  }
  int radiusSearch(const PointT&, double, std::vector<int>& idx, std::vector<float>& d2, unsigned int = 0) const override {
    idx.clear(); d2.clear();
    return 0;
  }
};
}  // namespace search
}  // namespace pcl
#endif
