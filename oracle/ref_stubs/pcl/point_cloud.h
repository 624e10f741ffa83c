// TEST INFRASTRUCTURE — see pcl/point_types.h next to this file
#include <pcl/point_types.h>
