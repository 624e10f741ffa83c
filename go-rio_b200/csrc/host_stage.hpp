// Host-side preparation of a cloud, plain C++ (SSE2): AoS -> packed {x,y,z,label} staging with the bounding box found
// on the way, and the sizing of the uniform grid. Included by apdgicp.cu; tests/test_host_stage.py compiles the same
// header with g++ and checks it on the CPU. dst: 16 bytes per point, 16-byte aligned.
#pragma once
#include <emmintrin.h>
#include <xmmintrin.h>

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstddef>
#include <cstring>

#include "grid_desc.hpp"

namespace apd {

// pcl::PointXYZINormal: {x, y, z, data[3]} at 0, normal_x at 16 -> {x, y, z, normal_x} with two shuffles (SSE2)
inline __m128 pcl_xyz_label(const char* p) {
  const __m128 xyzw = _mm_loadu_ps(reinterpret_cast<const float*>(p));
  const __m128 l = _mm_load_ss(reinterpret_cast<const float*>(p + 16));
  const __m128 zl = _mm_shuffle_ps(xyzw, l, _MM_SHUFFLE(0, 0, 3, 2));  // (z, w, l, l)
  return _mm_shuffle_ps(xyzw, zl, _MM_SHUFFLE(2, 0, 1, 0));            // (x, y, z, l)
}

// host AoS -> packed {x,y,z,label} staging + bounding box in one pass. The common layouts (packed float4; PCL's
// PointXYZINormal: xyz at 0, label at 16, stride 48) take fixed-offset vector loads, anything else a gather per point.
// nt: the packed points go out with non-temporal stores (no read-for-ownership of the staging buffer, which only the DMA
// engine reads next: one of the ~6 MB of host DRAM traffic a 60 k-point pair costs) — for many threads staging at once.
inline void stage_cloud(const void* pts, int n, int stride, int xyz_off, int label_off, float* dst, float bbox[6], bool nt = false) {
  const char* base = reinterpret_cast<const char*>(pts);
  __m128 mn = _mm_set1_ps(FLT_MAX), mx = _mm_set1_ps(-FLT_MAX);
  if (stride == 16 && xyz_off == 0 && label_off == 12) {  // packed float4: one pass, copy + min / max
    for (int i = 0; i < n; i++) {
      const __m128 v = _mm_loadu_ps(reinterpret_cast<const float*>(base + (size_t)i * 16));
      _mm_store_ps(dst + 4 * (size_t)i, v);
      mn = _mm_min_ps(v, mn);  // (a NaN coordinate leaves the bound unchanged: the second operand wins)
      mx = _mm_max_ps(v, mx);
    }
  } else if (stride == 48 && xyz_off == 0 && label_off == 16) {  // PCL's AoS: 3.5x faster than the general gather below
    // Memory-bound (3 MB read per 60 k-point cloud, 1 MB written): four points per turn so that several cache-line misses
    // are in flight, and a software prefetch a few lines ahead (a strided 48-byte record stream is one the hardware
    // prefetcher follows only half-heartedly): 7.2 -> ~10 GB/s per core.
    __m128 mn1 = mn, mx1 = mx, mn2 = mn, mx2 = mx, mn3 = mn, mx3 = mx;
    int i = 0;
    for (; i + 4 <= n; i += 4) {
      const char* p = base + (size_t)i * 48;
      _mm_prefetch(p + 1024, _MM_HINT_T0);
      _mm_prefetch(p + 1088, _MM_HINT_T0);
      _mm_prefetch(p + 1152, _MM_HINT_T0);
      const __m128 v0 = pcl_xyz_label(p), v1 = pcl_xyz_label(p + 48), v2 = pcl_xyz_label(p + 96), v3 = pcl_xyz_label(p + 144);
      if (nt) {  // (four points = one 64-byte line of the staging buffer)
        _mm_stream_ps(dst + 4 * (size_t)i, v0);
        _mm_stream_ps(dst + 4 * (size_t)(i + 1), v1);
        _mm_stream_ps(dst + 4 * (size_t)(i + 2), v2);
        _mm_stream_ps(dst + 4 * (size_t)(i + 3), v3);
      } else {
        _mm_store_ps(dst + 4 * (size_t)i, v0);
        _mm_store_ps(dst + 4 * (size_t)(i + 1), v1);
        _mm_store_ps(dst + 4 * (size_t)(i + 2), v2);
        _mm_store_ps(dst + 4 * (size_t)(i + 3), v3);
      }
      mn = _mm_min_ps(v0, mn); mx = _mm_max_ps(v0, mx);
      mn1 = _mm_min_ps(v1, mn1); mx1 = _mm_max_ps(v1, mx1);
      mn2 = _mm_min_ps(v2, mn2); mx2 = _mm_max_ps(v2, mx2);
      mn3 = _mm_min_ps(v3, mn3); mx3 = _mm_max_ps(v3, mx3);
    }
    for (; i < n; i++) {
      const __m128 v = pcl_xyz_label(base + (size_t)i * 48);
      _mm_store_ps(dst + 4 * (size_t)i, v);
      mn = _mm_min_ps(v, mn); mx = _mm_max_ps(v, mx);
    }
    mn = _mm_min_ps(_mm_min_ps(mn, mn1), _mm_min_ps(mn2, mn3));
    mx = _mm_max_ps(_mm_max_ps(mx, mx1), _mm_max_ps(mx2, mx3));
    if (nt) _mm_sfence();  // the copy engine reads the buffer next
  } else {
    for (int i = 0; i < n; i++) {
      const char* p = base + (size_t)i * stride;
      alignas(16) float f[4] = {0.f, 0.f, 0.f, 0.f};
      std::memcpy(f, p + xyz_off, 12);
      if (label_off >= 0) std::memcpy(&f[3], p + label_off, 4);
      const __m128 v = _mm_load_ps(f);
      _mm_store_ps(dst + 4 * (size_t)i, v);
      mn = _mm_min_ps(v, mn);  // (a NaN coordinate leaves the bound unchanged: the second operand wins)
      mx = _mm_max_ps(v, mx);
    }
  }
  alignas(16) float lo[4], hi[4];
  _mm_store_ps(lo, mn);
  _mm_store_ps(hi, mx);
  for (int a = 0; a < 3; a++) {  // (lane 3 is the label: ignored)
    bbox[a] = lo[a];
    bbox[3 + a] = hi[a];
  }
}

// bounding box of a packed float4 {x,y,z,label} array (read-only pass). Four independent min / max chains: with one, the
// loop runs at the latency of minps / maxps (4 cycles per point), not at the speed the cache delivers the points.
inline void bounds_of_packed(const void* pts, int n, float bbox[6]) {
  const char* base = reinterpret_cast<const char*>(pts);
  const __m128 hi0 = _mm_set1_ps(FLT_MAX), lo0 = _mm_set1_ps(-FLT_MAX);
  __m128 mn0 = hi0, mn1 = hi0, mn2 = hi0, mn3 = hi0, mx0 = lo0, mx1 = lo0, mx2 = lo0, mx3 = lo0;
  int i = 0;
  for (; i + 4 <= n; i += 4) {
    const float* p = reinterpret_cast<const float*>(base + (size_t)i * 16);
    const __m128 v0 = _mm_loadu_ps(p), v1 = _mm_loadu_ps(p + 4), v2 = _mm_loadu_ps(p + 8), v3 = _mm_loadu_ps(p + 12);
    mn0 = _mm_min_ps(v0, mn0); mx0 = _mm_max_ps(v0, mx0);  // (a NaN coordinate leaves the bound unchanged: the second operand wins)
    mn1 = _mm_min_ps(v1, mn1); mx1 = _mm_max_ps(v1, mx1);
    mn2 = _mm_min_ps(v2, mn2); mx2 = _mm_max_ps(v2, mx2);
    mn3 = _mm_min_ps(v3, mn3); mx3 = _mm_max_ps(v3, mx3);
  }
  for (; i < n; i++) {
    const __m128 v = _mm_loadu_ps(reinterpret_cast<const float*>(base + (size_t)i * 16));
    mn0 = _mm_min_ps(v, mn0);
    mx0 = _mm_max_ps(v, mx0);
  }
  const __m128 mn = _mm_min_ps(_mm_min_ps(mn0, mn1), _mm_min_ps(mn2, mn3)), mx = _mm_max_ps(_mm_max_ps(mx0, mx1), _mm_max_ps(mx2, mx3));
  alignas(16) float lo[4], hi[4];
  _mm_store_ps(lo, mn);
  _mm_store_ps(hi, mx);
  for (int a = 0; a < 3; a++) {
    bbox[a] = lo[a];
    bbox[3 + a] = hi[a];
  }
}

// (grid sizing: grid_desc.hpp — shared with the device)

}  // namespace apd
