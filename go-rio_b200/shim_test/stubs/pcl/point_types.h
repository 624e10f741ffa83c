// MINIMAL STAND-IN for <pcl/point_types.h>: pcl::PointXYZINormal with PCL 1.10's
// 48-byte layout (x,y,z,pad | normal_x,normal_y,normal_z,pad | intensity,curvature,pad,pad).
// (intensity at byte 32, curvature at 36, as PCL_ADD_POINT4D / PCL_ADD_NORMAL4D / the {intensity, curvature} union lay it out)
#ifndef APD_STUB_PCL_POINT_TYPES
#define APD_STUB_PCL_POINT_TYPES
#define PCL_VERSION_CALC(MAJ, MIN, PATCH) ((MAJ)*100000 + (MIN)*100 + (PATCH))
#define PCL_VERSION PCL_VERSION_CALC(1, 10, 0)
#include <memory>
namespace pcl {
template <typename T> using shared_ptr = std::shared_ptr<T>;
struct alignas(16) PointXYZINormal {
  float x = 0, y = 0, z = 0, data3 = 1.f;
  float normal_x = 0, normal_y = 0, normal_z = 0, data_n3 = 0;
  float intensity = 0, curvature = 0, pad0 = 0, pad1 = 0;
};
static_assert(sizeof(PointXYZINormal) == 48, "PointXYZINormal must be 48 bytes as in PCL");
}  // namespace pcl
#endif
