// MINIMAL STAND-IN for <pcl/point_cloud.h>.
#ifndef APD_STUB_PCL_POINT_CLOUD
#define APD_STUB_PCL_POINT_CLOUD
#include <memory>
#include <vector>
#include <pcl/point_types.h>
namespace pcl {
template <typename PointT>
struct PointCloud {
  using PointType = PointT;
  using Ptr = std::shared_ptr<PointCloud<PointT>>;
  using ConstPtr = std::shared_ptr<const PointCloud<PointT>>;
  std::vector<PointT> points;
  std::size_t size() const { return points.size(); }
  bool empty() const { return points.empty(); }
  PointT& operator[](std::size_t i) { return points[i]; }
  const PointT& operator[](std::size_t i) const { return points[i]; }
  PointT& at(std::size_t i) { return points.at(i); }
  const PointT& at(std::size_t i) const { return points.at(i); }
  void push_back(const PointT& p) { points.push_back(p); }
};
}  // namespace pcl
#endif
