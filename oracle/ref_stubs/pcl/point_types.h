// TEST INFRASTRUCTURE — the few PCL types the reference's own headers need, so that they can be compiled IN PLACE from
// /root/reference without PCL (oracle/Makefile: _ref/libapd_ref_dbscan.so, _ref/libapd_ref_apdgicp.so). Not PCL: stand-ins
// with the same member names (pcl::PointXYZINormal's fields and its getVector4fMap / getVector3fMap views, PointCloud,
// PointIndices, PCLHeader).
#ifndef APDO_REF_STUB_PCL_POINT_TYPES
#define APDO_REF_STUB_PCL_POINT_TYPES
#include <Eigen/Core>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>
#include <memory>
#include <string>
#include <vector>

#define PCL_VERSION_CALC(MAJ, MIN, PATCH) (MAJ * 100000 + MIN * 100 + PATCH)
#define PCL_VERSION PCL_VERSION_CALC(1, 10, 0)

namespace pcl {
template <class T>
using shared_ptr = std::shared_ptr<T>;

struct PCLHeader {
  std::uint32_t seq = 0;
  std::uint64_t stamp = 0;
  std::string frame_id;
};
struct PointIndices {
  PCLHeader header;
  std::vector<int> indices;
};

// a view of N consecutive floats of a point as an Eigen vector (Eigen::Map in PCL)
template <int N>
struct FloatMap {
  float* p;
  FloatMap& operator=(const Eigen::Matrix<float, N, 1>& v) {
    for (int i = 0; i < N; i++) p[i] = v.a[i];
    return *this;
  }
  operator Eigen::Matrix<float, N, 1>() const {
    Eigen::Matrix<float, N, 1> v;
    for (int i = 0; i < N; i++) v.a[i] = p[i];
    return v;
  }
  template <class U>
  Eigen::Matrix<U, N, 1> cast() const {
    Eigen::Matrix<U, N, 1> v;
    for (int i = 0; i < N; i++) v.a[i] = static_cast<U>(p[i]);
    return v;
  }
};

// PCL_ADD_POINT4D (data[3] = 1) | PCL_ADD_NORMAL4D | {intensity, curvature}
struct PointXYZINormal {
  union {
    float data[4];
    struct {
      float x, y, z;
    };
  };
  union {
    float data_n[4];
    struct {
      float normal_x, normal_y, normal_z;
    };
  };
  float intensity = 0.f, curvature = 0.f;
  PointXYZINormal() {
    data[0] = data[1] = data[2] = 0.f;
    data[3] = 1.f;
    data_n[0] = data_n[1] = data_n[2] = data_n[3] = 0.f;
  }
  FloatMap<4> getVector4fMap() { return FloatMap<4>{data}; }
  FloatMap<4> getVector4fMap() const { return FloatMap<4>{const_cast<float*>(data)}; }
  FloatMap<3> getVector3fMap() { return FloatMap<3>{data}; }
  FloatMap<3> getVector3fMap() const { return FloatMap<3>{const_cast<float*>(data)}; }
};

template <typename PointT>
struct PointCloud {
  using PointType = PointT;
  using Ptr = std::shared_ptr<PointCloud<PointT>>;
  using ConstPtr = std::shared_ptr<const PointCloud<PointT>>;
  PCLHeader header;
  std::vector<PointT> points;
  std::uint32_t width = 0, height = 1;
  bool is_dense = true;
  std::size_t size() const { return points.size(); }
  const PointT& at(std::size_t i) const { return points.at(i); }
  PointT& at(std::size_t i) { return points.at(i); }
  const PointT& operator[](std::size_t i) const { return points[i]; }
  PointT& operator[](std::size_t i) { return points[i]; }
};
}  // namespace pcl
#endif
