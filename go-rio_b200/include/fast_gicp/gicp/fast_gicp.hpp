// B200 drop-in of fast_gicp::FastGICP (reference
// fast_apdgicp/include/fast_gicp/gicp/fast_gicp.hpp:19-104 and impl/fast_gicp_impl.hpp), the class
// 4DRadarSLAM's select_registration_method builds for registration_method = "FAST_GICP"
// (registrations.cpp:28-37). FastGICP is FastAPDGICP without the radar noise term in the combined covariance
// (fast_gicp_impl.hpp:157 against fast_apdgicp_impl.hpp:213-215) and with unit weights (:205 against :276): same
// covariances, correspondences, optimizer and kernels, selected by apd_params.variant = APD_VARIANT_GICP.
// The radar-noise setters of the base are inherited and have no effect.
#ifndef FAST_GICP_FAST_GICP_HPP
#define FAST_GICP_FAST_GICP_HPP

#include <fast_gicp/gicp/fast_apdgicp.hpp>

namespace fast_gicp {

template <typename PointSource, typename PointTarget>
class FastGICP : public FastAPDGICP<PointSource, PointTarget> {
public:
  APD_SHIM_REGISTRATION_ALIASES(FastGICP, PointSource, PointTarget);

  FastGICP() {
    this->reg_name_ = "FastGICP";
    this->variant_ = APD_VARIANT_GICP;
  }
};

}  // namespace fast_gicp

#endif
