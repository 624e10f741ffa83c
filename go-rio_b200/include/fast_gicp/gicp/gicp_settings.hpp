// Drop-in replacement of fast_gicp/gicp/gicp_settings.hpp (reference
// fast_apdgicp/include/fast_gicp/gicp/gicp_settings.hpp:6). Same enumerators in
// the same order; the values are the APD_REG_* codes of include/apdgicp.h.
#ifndef FAST_GICP_GICP_SETTINGS_HPP
#define FAST_GICP_GICP_SETTINGS_HPP

namespace fast_gicp {

enum class RegularizationMethod { NONE = 0, MIN_EIG = 1, NORMALIZED_MIN_EIG = 2, PLANE = 3, FROBENIUS = 4 };

// FastVGICP (gicp_settings.hpp:10-12); the values are the APD_VOXEL_* codes. DIRECT_RADIUS exists for the reference's
// VGICP_CUDA only and is refused here as it is by the reference's CPU class (fast_vgicp_voxel.hpp:13-15).
enum class NeighborSearchMethod { DIRECT27 = 0, DIRECT7 = 1, DIRECT1 = 2, DIRECT_RADIUS = 3 };

enum class VoxelAccumulationMode { ADDITIVE = 0, ADDITIVE_WEIGHTED = 1, MULTIPLICATIVE = 2 };

}  // namespace fast_gicp

#endif
