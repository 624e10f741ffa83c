// TEST INFRASTRUCTURE — CPU restatement of the stages on either side of the registration that run on the same
// neighbour-search machinery: the radius searches of the preprocessing nodelet and the submap assembly +
// voxel-grid downsample of the scan-to-map branch. Nothing in the product links or calls this file.
//
//  * apdo_radius_search — what pcl::RadiusOutlierRemoval (preprocessing_nodelet_ntu.cpp:163-172) and DBSCANKdtreeCluster's
//    neighbour queries (:520-532, eps 0.9) ask of pcl::search::KdTree::radiusSearch [ext: PCL 1.10 / FLANN 1.9.1
//    RadiusResultSet: a point is a neighbour when its fp32 squared distance ((dx*dx + dy*dy) + dz*dz, L2_Simple) is
//    STRICTLY below radius^2; the query point itself is among its neighbours]. Brute force: the definition.
//  * apdo_submap_assemble + apdo_voxel_grid — scan_matching_odometry_nodelet.cpp:602-618: every keyframe cloud moved by
//    rel_pose = odom_i^-1 * odom_last with pcl::transformPointCloud(cloud, out, Eigen::Matrix4d) [ext: PCL 1.10
//    transforms.hpp, detail::Transformer<double>::se3 — each coordinate is tf(r,0)*x + tf(r,1)*y + tf(r,2)*z + tf(r,3)
//    summed left to right in DOUBLE and cast to float; normals are not touched], concatenated in keyframe order, then
//    pcl::VoxelGrid<PointT> with leaf = downsample_resolution [ext: PCL 1.10 voxel_grid.hpp applyFilter, downsample_all_data:
//    bounds over the finite points; min_b = floor(min * inv_leaf); voxel index ijk = floor(p * inv_leaf) - (float)min_b
//    (float arithmetic), idx = i + j*dx + k*dx*dy; voxels output in ascending idx; per voxel the CentroidPoint of its
//    members: x,y,z summed in float and divided by the count; the normal summed and NORMALISED (AccumulatorNormal) — with
//    normal_y = normal_z = 0, as this pipeline leaves them, the cluster label normal_x becomes 1 where any member had a
//    positive label and stays 0 otherwise]. PCL's std::sort leaves the order of a voxel's members unspecified; here (and on
//    the GPU) members are summed in ascending original index, so the two agree bit for bit; against PCL itself the
//    centroids agree to float rounding.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <numeric>
#include <vector>

extern "C" {

// pass 1 (indices == nullptr): counts[i] = neighbours of point i; pass 2: indices[offsets[i] ...] in ascending index
int apdo_radius_search(const float* xyzl, int32_t n, double radius, int32_t* counts, const int64_t* offsets, int32_t* indices) {
  const float r2 = (float)(radius * radius);  // [ext] pcl::KdTreeFLANN::radiusSearch: static_cast<float>(radius * radius)
#pragma omp parallel for schedule(dynamic, 64)
  for (int32_t i = 0; i < n; i++) {
    const float qx = xyzl[4 * (size_t)i], qy = xyzl[4 * (size_t)i + 1], qz = xyzl[4 * (size_t)i + 2];
    int32_t c = 0;
    for (int32_t j = 0; j < n; j++) {
      const float dx = qx - xyzl[4 * (size_t)j], dy = qy - xyzl[4 * (size_t)j + 1], dz = qz - xyzl[4 * (size_t)j + 2];
      const float d2 = (dx * dx + dy * dy) + dz * dz;
      if (d2 < r2) {
        if (indices) indices[offsets[i] + c] = j;
        c++;
      }
    }
    if (counts) counts[i] = c;
  }
  return 0;
}

// out: float4 {x,y,z,label} per voxel (capacity n), ascending voxel index; returns 0, or 1 when PCL would refuse the leaf
// size (index overflow: "Leaf size is too small for the input dataset") and pass the cloud through unchanged
int apdo_voxel_grid(const float* xyzl, int32_t n, float leaf, float* out, int32_t* n_out) {
  const float inv = 1.0f / leaf;
  float mn[3] = {std::numeric_limits<float>::max(), std::numeric_limits<float>::max(), std::numeric_limits<float>::max()};
  float mx[3] = {-mn[0], -mn[1], -mn[2]};
  std::vector<int32_t> finite;
  for (int32_t i = 0; i < n; i++) {
    const float* p = xyzl + 4 * (size_t)i;
    if (!std::isfinite(p[0]) || !std::isfinite(p[1]) || !std::isfinite(p[2])) continue;
    finite.push_back(i);
    for (int a = 0; a < 3; a++) {
      mn[a] = std::min(mn[a], p[a]);
      mx[a] = std::max(mx[a], p[a]);
    }
  }
  *n_out = 0;
  if (finite.empty()) return 0;
  const int64_t dx = (int64_t)((mx[0] - mn[0]) * inv) + 1, dy = (int64_t)((mx[1] - mn[1]) * inv) + 1, dz = (int64_t)((mx[2] - mn[2]) * inv) + 1;
  if (dx * dy * dz > (int64_t)std::numeric_limits<int32_t>::max()) {
    std::memcpy(out, xyzl, (size_t)n * 16);
    *n_out = n;
    return 1;
  }
  int min_b[3], div_b[3];
  for (int a = 0; a < 3; a++) {
    min_b[a] = (int)std::floor(mn[a] * inv);
    div_b[a] = (int)std::floor(mx[a] * inv) - min_b[a] + 1;
  }
  const int mul[3] = {1, div_b[0], div_b[0] * div_b[1]};
  std::vector<std::pair<uint32_t, int32_t>> iv;
  iv.reserve(finite.size());
  for (int32_t i : finite) {
    const float* p = xyzl + 4 * (size_t)i;
    const int i0 = (int)(std::floor(p[0] * inv) - (float)min_b[0]);
    const int i1 = (int)(std::floor(p[1] * inv) - (float)min_b[1]);
    const int i2 = (int)(std::floor(p[2] * inv) - (float)min_b[2]);
    iv.emplace_back((uint32_t)(i0 * mul[0] + i1 * mul[1] + i2 * mul[2]), i);
  }
  std::stable_sort(iv.begin(), iv.end(), [](const auto& a, const auto& b) { return a.first < b.first; });  // members in index order
  size_t b = 0;
  while (b < iv.size()) {
    size_t e = b + 1;
    while (e < iv.size() && iv[e].first == iv[b].first) e++;
    float sx = 0.f, sy = 0.f, sz = 0.f, sn = 0.f;
    for (size_t q = b; q < e; q++) {
      const float* p = xyzl + 4 * (size_t)iv[q].second;
      sx += p[0]; sy += p[1]; sz += p[2]; sn += p[3];
    }
    const float cnt = (float)(e - b);
    float* o = out + 4 * (size_t)(*n_out);
    o[0] = sx / cnt; o[1] = sy / cnt; o[2] = sz / cnt;
    const float nn = sn * sn;  // squaredNorm of (sn, 0, 0, 0)
    o[3] = nn > 0.f ? sn / std::sqrt(nn) : sn;  // AccumulatorNormal::get: normalize() (Eigen: only when the norm is positive)
    (*n_out)++;
    b = e;
  }
  return 0;
}

// clouds[c]: float4 points of keyframe c (n[c] of them); poses: n_clouds column-major 4x4 doubles (rel_pose of each
// keyframe). out: capacity sum(n) float4. The assembled, not yet downsampled submap.
int apdo_submap_assemble(const float* const* clouds, const int32_t* n, const double* poses, int32_t n_clouds, float* out, int32_t* n_out) {
  int32_t w = 0;
  for (int32_t c = 0; c < n_clouds; c++) {
    const double* T = poses + 16 * (size_t)c;
    for (int32_t i = 0; i < n[c]; i++) {
      const float* p = clouds[c] + 4 * (size_t)i;
      const double x = p[0], y = p[1], z = p[2];
      float* o = out + 4 * (size_t)w++;
      for (int r = 0; r < 3; r++) o[r] = (float)(T[0 * 4 + r] * x + T[1 * 4 + r] * y + T[2 * 4 + r] * z + T[3 * 4 + r]);
      o[3] = p[3];
    }
  }
  *n_out = w;
  return 0;
}

}  // extern "C"
