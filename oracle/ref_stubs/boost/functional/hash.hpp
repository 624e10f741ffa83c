// TEST INFRASTRUCTURE — stand-in for boost::hash_combine as fast_vgicp_voxel.hpp uses it (the hash only decides where a
// voxel lives in the unordered_map, not what it holds)
#ifndef APDO_REF_STUB_BOOST_HASH
#define APDO_REF_STUB_BOOST_HASH
#include <cstddef>
#include <functional>
namespace boost {
template <class T>
inline void hash_combine(std::size_t& seed, const T& v) {
  seed ^= std::hash<T>()(v) + 0x9e3779b9 + (seed << 6) + (seed >> 2);
}
}  // namespace boost
#endif
