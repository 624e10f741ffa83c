// The stages on either side of the registration that run on the same machinery ("next" rows):
//  * the radius searches of the preprocessing nodelet — pcl::RadiusOutlierRemoval and DBSCANKdtreeCluster's neighbour
//    queries (4DRadarSLAM/apps/preprocessing_nodelet_ntu.cpp:163-172, :520-532) — on the cloud's uniform grid;
//  * the submap assembly of the scan-to-map branch (scan_matching_odometry_nodelet.cpp:602-618): keyframe clouds moved by
//    their relative pose (pcl::transformPointCloud with a Matrix4d), concatenated, and pcl::VoxelGrid-downsampled.
// Semantics of the PCL pieces: oracle/apd_prep_oracle.cpp (the CPU restatement these kernels are tested against).
#include "point_math.cuh"

namespace apd {

namespace {

constexpr int kThreads = 256;

// ---- radius search: one thread per (sorted) point; the cells within ceil(r / cell) of its own, row by row ----
// mode 0: one radius for every query (r2 = float(radius * radius), as pcl::KdTreeFLANN::radiusSearch hands it to FLANN).
// mode 1 / 2: the range-dependent radii of DBSCANSimpleCluster::extract (4DRadarSLAM/include/dbscan/DBSCAN_simple.h):
//   1 (a seed, :35-38):      double norm = sqrt(x*x + y*y + z*z) [float expression, float sqrt]; |norm - 1| / 50 + eps
//   2 (an expansion, :60-63): (sqrt(x*x + y*y + z*z) - 1) / 100 [all float] + eps
__device__ __forceinline__ float query_r2(int mode, float r2_fixed, double eps, const float4& q, float& radius) {
  if (mode == 0) {
    radius = sqrtf(r2_fixed) * 1.0001f;  // (only sizes the reach in cells)
    return r2_fixed;
  }
  const float nf = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(q.x, q.x), __fmul_rn(q.y, q.y)), __fmul_rn(q.z, q.z)));
  double rad;
  if (mode == 1) rad = __dadd_rn(__ddiv_rn(fabs(__dsub_rn((double)nf, 1.0)), 50.0), eps);
  else rad = __dadd_rn((double)__fdiv_rn(__fsub_rn(nf, 1.0f), 100.0f), eps);
  radius = (float)rad * 1.0001f;
  return (float)__dmul_rn(rad, rad);
}
__global__ void __launch_bounds__(kThreads) radius_search_kernel(const float4* __restrict__ spts, const uint32_t* __restrict__ cell_start, GridDesc g, int n,
                                                                 float r2_fixed, int mode, double eps, int32_t* __restrict__ counts,
                                                                 const long long* __restrict__ offsets, int32_t* __restrict__ indices) {
  const int w = blockIdx.x * kThreads + threadIdx.x;
  if (w >= n) return;
  const float4 q = spts[w];
  const int oi = __float_as_int(q.w);
  float radius;
  const float r2 = query_r2(mode, r2_fixed, eps, q, radius);
  // every point within `radius` lies in a cell at most R away (the 0.01 covers the fp32 rounding of the cell coordinate)
  const int R = radius > 0.f ? (int)floorf(radius * g.inv_cell + 0.01f) + 1 : 0;
  const int cx = cell_coord(q.x, g.ox, g.inv_cell, g.nx), cy = cell_coord(q.y, g.oy, g.inv_cell, g.ny), cz = cell_coord(q.z, g.oz, g.inv_cell, g.nz);
  const int x0 = max(cx - R, 0), x1 = min(cx + R, g.nx - 1);
  int c = 0;
  const long long base = offsets ? offsets[oi] : 0;
  for (int z = max(cz - R, 0); z <= min(cz + R, g.nz - 1); z++)
    for (int y = max(cy - R, 0); y <= min(cy + R, g.ny - 1); y++) {
      const int row = (z * g.ny + y) * g.nx;
      const int b = (int)cell_start[row + x0], e = (int)cell_start[row + x1 + 1];
      for (int j = b; j < e; j++) {
        const float4 p = spts[j];
        if (sqdist_rn(q.x, q.y, q.z, p.x, p.y, p.z) < r2) {  // FLANN RadiusResultSet: dist < radius^2
          if (indices) indices[base + c] = __float_as_int(p.w);
          c++;
        }
      }
    }
  if (counts) counts[oi] = c;
}

__global__ void __launch_bounds__(kThreads) transform_d_kernel(const float4* __restrict__ in, int n, PoseD T, float4* __restrict__ out) {
  const int i = blockIdx.x * kThreads + threadIdx.x;
  if (i >= n) return;
  const float4 p = in[i];
  const double x = p.x, y = p.y, z = p.z;
  // pcl::detail::Transformer<double>::se3: tf(r,0)*x + tf(r,1)*y + tf(r,2)*z + tf(r,3), left to right, cast to float
  const float ox = (float)__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(T.r[0], x), __dmul_rn(T.r[1], y)), __dmul_rn(T.r[2], z)), T.t[0]);
  const float oy = (float)__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(T.r[3], x), __dmul_rn(T.r[4], y)), __dmul_rn(T.r[5], z)), T.t[1]);
  const float oz = (float)__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(T.r[6], x), __dmul_rn(T.r[7], y)), __dmul_rn(T.r[8], z)), T.t[2]);
  out[i] = make_float4(ox, oy, oz, p.w);
}

// ---- voxel grid ----
__device__ __forceinline__ unsigned int f2ord(float f) {
  unsigned int u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ bool finite3(const float4& p) { return isfinite(p.x) && isfinite(p.y) && isfinite(p.z); }

// state6: ordered-uint {min x,y,z = 0xffffffff, max x,y,z = 0} prepared by the caller; over the FINITE points
__global__ void __launch_bounds__(kThreads) voxel_bounds_kernel(const float4* __restrict__ pts, int n, unsigned int* state) {
  float mn[3] = {3.4e38f, 3.4e38f, 3.4e38f}, mx[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 p = pts[i];
    if (!finite3(p)) continue;
    mn[0] = fminf(mn[0], p.x); mx[0] = fmaxf(mx[0], p.x);
    mn[1] = fminf(mn[1], p.y); mx[1] = fmaxf(mx[1], p.y);
    mn[2] = fminf(mn[2], p.z); mx[2] = fmaxf(mx[2], p.z);
  }
#pragma unroll
  for (int a = 0; a < 3; a++) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
      mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
    }
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int a = 0; a < 3; a++) {
      if (mn[a] <= mx[a]) {
        atomicMin(&state[a], f2ord(mn[a]));
        atomicMax(&state[3 + a], f2ord(mx[a]));
      }
    }
  }
}

// voxel index of every point (pcl voxel_grid.hpp: floor(p * inv_leaf) - (float)min_b, float arithmetic); non-finite points
// get the largest key and sort to the end
__global__ void __launch_bounds__(kThreads) voxel_keys_kernel(const float4* __restrict__ pts, int n, float inv, int mb0, int mb1, int mb2, int m1, int m2,
                                                              uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  const int i = blockIdx.x * kThreads + threadIdx.x;
  if (i >= n) return;
  const float4 p = pts[i];
  uint32_t key = 0xffffffffu;
  if (finite3(p)) {
    const int i0 = (int)__fsub_rn(floorf(__fmul_rn(p.x, inv)), (float)mb0);
    const int i1 = (int)__fsub_rn(floorf(__fmul_rn(p.y, inv)), (float)mb1);
    const int i2 = (int)__fsub_rn(floorf(__fmul_rn(p.z, inv)), (float)mb2);
    key = (uint32_t)(i0 + i1 * m1 + i2 * m2);
  }
  keys[i] = key;
  vals[i] = (uint32_t)i;
}

__global__ void __launch_bounds__(kThreads) voxel_heads_kernel(const uint32_t* __restrict__ keys, int n, uint32_t* __restrict__ heads) {
  const int i = blockIdx.x * kThreads + threadIdx.x;
  if (i > n) return;
  if (i == n) {  // (one slot past the end: the scan leaves the voxel count there)
    heads[n] = 0u;
    return;
  }
  const uint32_t k = keys[i];
  heads[i] = (k != 0xffffffffu && (i == 0 || keys[i - 1] != k)) ? 1u : 0u;
}

// one thread per voxel head: CentroidPoint of the voxel's members (sorted by key, and — the sort being stable — by original
// index within a voxel): x, y, z summed in float and divided by the count, the label (normal_x) summed and normalised
__global__ void __launch_bounds__(kThreads) voxel_centroids_kernel(const float4* __restrict__ pts, const uint32_t* __restrict__ keys,
                                                                   const uint32_t* __restrict__ vals, const uint32_t* __restrict__ voxel_of, int n,
                                                                   float4* __restrict__ out) {
  const int i = blockIdx.x * kThreads + threadIdx.x;
  if (i >= n) return;
  const uint32_t k = keys[i];
  if (k == 0xffffffffu || (i > 0 && keys[i - 1] == k)) return;
  float sx = 0.f, sy = 0.f, sz = 0.f, sn = 0.f;
  int cnt = 0;
  for (int j = i; j < n && keys[j] == k; j++) {
    const float4 p = pts[vals[j]];
    sx = __fadd_rn(sx, p.x); sy = __fadd_rn(sy, p.y); sz = __fadd_rn(sz, p.z); sn = __fadd_rn(sn, p.w);
    cnt++;
  }
  const float c = (float)cnt;
  const float nn = __fmul_rn(sn, sn);
  out[voxel_of[i]] = make_float4(__fdiv_rn(sx, c), __fdiv_rn(sy, c), __fdiv_rn(sz, c), nn > 0.f ? __fdiv_rn(sn, __fsqrt_rn(nn)) : sn);
}

}  // namespace

void launch_radius_search(const CloudDev& c, float r2_fixed, int mode, double eps, int32_t* d_counts, const long long* d_offsets, int32_t* d_indices,
                          cudaStream_t s, int64_t* launches) {
  if (c.n <= 0) return;
  radius_search_kernel<<<(c.n + kThreads - 1) / kThreads, kThreads, 0, s>>>(c.spts, c.cell_start, c.g, c.n, r2_fixed, mode, eps, d_counts, d_offsets,
                                                                            d_indices);
  (*launches)++;
}

void launch_transform_cloud_d(const float4* in, int n, const double* T, float4* out, cudaStream_t s, int64_t* launches) {
  if (n <= 0) return;
  PoseD P;
  for (int r = 0; r < 3; r++) {
    for (int c = 0; c < 3; c++) P.r[r * 3 + c] = T[c * 4 + r];
    P.t[r] = T[3 * 4 + r];
  }
  transform_d_kernel<<<(n + kThreads - 1) / kThreads, kThreads, 0, s>>>(in, n, P, out);
  (*launches)++;
}

void launch_voxel_bounds(const float4* pts, int n, unsigned int* d_state6, cudaStream_t s, int64_t* launches) {
  const int blocks = max(1, min(148 * 8, (n + kThreads - 1) / kThreads));
  voxel_bounds_kernel<<<blocks, kThreads, 0, s>>>(pts, n, d_state6);
  (*launches)++;
}
void launch_voxel_keys(const float4* pts, int n, float inv_leaf, const int min_b[3], const int mul[3], uint32_t* keys, uint32_t* vals, cudaStream_t s,
                       int64_t* launches) {
  voxel_keys_kernel<<<(n + kThreads - 1) / kThreads, kThreads, 0, s>>>(pts, n, inv_leaf, min_b[0], min_b[1], min_b[2], mul[1], mul[2], keys, vals);
  (*launches)++;
}
void launch_voxel_heads(const uint32_t* keys, int n, uint32_t* heads, cudaStream_t s, int64_t* launches) {
  voxel_heads_kernel<<<(n + 1 + kThreads - 1) / kThreads, kThreads, 0, s>>>(keys, n, heads);
  (*launches)++;
}
void launch_voxel_centroids(const float4* pts, const uint32_t* keys, const uint32_t* vals, const uint32_t* voxel_of, int n, float4* out, cudaStream_t s,
                            int64_t* launches) {
  voxel_centroids_kernel<<<(n + kThreads - 1) / kThreads, kThreads, 0, s>>>(pts, keys, vals, voxel_of, n, out);
  (*launches)++;
}

// Loads this file's kernels into the current context (CUDA loads kernels lazily, at their first launch, and a load may have
// to synchronise with the context: if it happens while another rank's kernel of the same process is spinning on a peer
// — the sharded exchange — neither can proceed. apd_group_create loads everything up front.)
void preload_prep_kernels() {
  cudaFuncAttributes a;
  (void)cudaFuncGetAttributes(&a, radius_search_kernel);
  (void)cudaFuncGetAttributes(&a, transform_d_kernel);
  (void)cudaFuncGetAttributes(&a, voxel_bounds_kernel);
  (void)cudaFuncGetAttributes(&a, voxel_keys_kernel);
  (void)cudaFuncGetAttributes(&a, voxel_heads_kernel);
  (void)cudaFuncGetAttributes(&a, voxel_centroids_kernel);
}

}  // namespace apd
