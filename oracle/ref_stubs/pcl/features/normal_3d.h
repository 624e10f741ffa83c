// TEST INFRASTRUCTURE — included by the reference's fast_apdgicp_impl.hpp, nothing of it is used
#include <pcl/point_types.h>
