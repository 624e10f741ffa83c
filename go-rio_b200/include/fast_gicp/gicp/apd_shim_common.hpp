// Shared plumbing of the B200 shim headers (no reference counterpart).
#ifndef FAST_GICP_APD_SHIM_COMMON_HPP
#define FAST_GICP_APD_SHIM_COMMON_HPP

#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#include <Eigen/Core>
#include <Eigen/Geometry>

#include <pcl/point_types.h>
#include <pcl/point_cloud.h>
#include <pcl/registration/registration.h>

#include <apdgicp.h>  // the C-ABI (repo include/)

// The aliases every class of the pcl::Registration family publishes; the
// reference spells them out in each class (lsq_registration.hpp:18-35,
// fast_apdgicp.hpp:22-39) — same names, same meaning.
#define APD_SHIM_REGISTRATION_ALIASES(Self, PS, PT)                                      \
  using Scalar = float;                                                                  \
  using PclBase = pcl::Registration<PS, PT, Scalar>;                                     \
  using Matrix4 = typename PclBase::Matrix4;                                             \
  using PointCloudSource = typename PclBase::PointCloudSource;                           \
  using PointCloudSourcePtr = typename PointCloudSource::Ptr;                            \
  using PointCloudSourceConstPtr = typename PointCloudSource::ConstPtr;                  \
  using PointCloudTarget = typename PclBase::PointCloudTarget;                           \
  using PointCloudTargetPtr = typename PointCloudTarget::Ptr;                            \
  using PointCloudTargetConstPtr = typename PointCloudTarget::ConstPtr;                  \
  using Ptr = APD_SHIM_SHARED_PTR<Self<PS, PT>>;                                         \
  using ConstPtr = APD_SHIM_SHARED_PTR<const Self<PS, PT>>

#if defined(PCL_VERSION) && defined(PCL_VERSION_CALC)
#if PCL_VERSION >= PCL_VERSION_CALC(1, 10, 0)
#define APD_SHIM_SHARED_PTR pcl::shared_ptr
#else
#define APD_SHIM_SHARED_PTR boost::shared_ptr
#endif
#else
#define APD_SHIM_SHARED_PTR pcl::shared_ptr
#endif

namespace fast_gicp {
namespace apd_detail {

// byte offset of the cluster label (pcl normal_x, written by the DBSCAN stage,
// 4DRadarSLAM/apps/preprocessing_nodelet_ntu.cpp:561-567) inside a point type,
// or -1 when the point type has no normal_x (the label then reads as 0, which is
// what FastAPDGICP would see for such a type if it compiled: it does not — the
// reference instantiates PointXYZINormal only, fast_apdgicp.cpp:6).
template <typename P, typename = void>
struct LabelOffset {
  static int value() { return -1; }
};
template <typename P>
struct LabelOffset<P, decltype(void(std::declval<P>().normal_x))> {
  static int value() {
    P p;
    return (int)(reinterpret_cast<const char*>(&p.normal_x) - reinterpret_cast<const char*>(&p));
  }
};
template <typename P>
inline int xyz_offset() {
  P p;
  return (int)(reinterpret_cast<const char*>(&p.x) - reinterpret_cast<const char*>(&p));
}

inline void check(apd_handle* h, int rc, const char* what) {
  // The reference never throws on this path: failures surface as
  // hasConverged() == false plus a message on stderr. Only a missing device at
  // construction is fatal, because there is no CPU fallback to hide behind.
  if (rc != APD_OK) std::fprintf(stderr, "[FastAPDGICP/B200] %s failed (%d): %s\n", what, rc, h ? apd_last_error(h) : "");
}

}  // namespace apd_detail
}  // namespace fast_gicp

#endif
