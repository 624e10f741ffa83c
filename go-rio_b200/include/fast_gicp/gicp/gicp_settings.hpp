// Drop-in replacement of fast_gicp/gicp/gicp_settings.hpp (reference
// fast_apdgicp/include/fast_gicp/gicp/gicp_settings.hpp:6). Same enumerators in
// the same order; the values are the APD_REG_* codes of include/apdgicp.h.
#ifndef FAST_GICP_GICP_SETTINGS_HPP
#define FAST_GICP_GICP_SETTINGS_HPP

namespace fast_gicp {

enum class RegularizationMethod { NONE = 0, MIN_EIG = 1, NORMALIZED_MIN_EIG = 2, PLANE = 3, FROBENIUS = 4 };

}  // namespace fast_gicp

#endif
