// B200 drop-in of fast_gicp::FastVGICP (reference fast_apdgicp/include/fast_gicp/gicp/fast_vgicp.hpp:22-81 and
// impl/fast_vgicp_impl.hpp), the class 4DRadarSLAM's select_registration_method builds for registration_method =
// "FAST_VGICP" (registrations.cpp:64-72). FastVGICP is FastGICP with the target replaced by a Gaussian voxel map: a source
// point is matched with the voxel(s) around its transformed position instead of its nearest target point. Same
// covariances and optimizer; the voxel map, the voxel correspondences and their sums are go-rio_b200/csrc/vgicp.cu,
// selected by apd_params.variant = APD_VARIANT_VGICP.
#ifndef FAST_GICP_FAST_VGICP_HPP
#define FAST_GICP_FAST_VGICP_HPP

#include <fast_gicp/gicp/fast_gicp.hpp>

namespace fast_gicp {

template <typename PointSource, typename PointTarget>
class FastVGICP : public FastGICP<PointSource, PointTarget> {
public:
  APD_SHIM_REGISTRATION_ALIASES(FastVGICP, PointSource, PointTarget);

  FastVGICP() {  // fast_vgicp_impl.hpp:19-25
    this->reg_name_ = "FastVGICP";
    this->variant_ = APD_VARIANT_VGICP;
    this->voxel_resolution_ = 1.0;
    this->voxel_search_ = APD_VOXEL_DIRECT1;
    this->voxel_mode_ = APD_VOXEL_ADDITIVE;
  }

  void setResolution(double resolution) { this->voxel_resolution_ = resolution; }                                       // :31-33
  void setNeighborSearchMethod(NeighborSearchMethod method) { this->voxel_search_ = static_cast<int>(method); }         // :36-38
  void setVoxelAccumulationMode(VoxelAccumulationMode mode) { this->voxel_mode_ = static_cast<int>(mode); }             // :41-43
};

}  // namespace fast_gicp

#endif
