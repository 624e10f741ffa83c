// FastVGICP behind the same handle (apd_params.variant = APD_VARIANT_VGICP): the voxelised GICP that
// 4DRadarSLAM's select_registration_method builds for registration_method = "FAST_VGICP" (registrations.cpp:64-72).
// Replaces, of reference fast_apdgicp/include/fast_gicp/gicp:
//   fast_vgicp_voxel.hpp:127-185  GaussianVoxelMap (an unordered_map of heap-allocated voxels filled point by point)
//       -> keys + stable radix sort + one thread per voxel (the members of a voxel are summed in index order, as the
//          reference's loop over the cloud does), looked up by binary search in the sorted key list
//   impl/fast_vgicp_impl.hpp:74-118   update_correspondences (point -> voxel(s), (C_B + T C_A T^T)^-1 per pair)
//   impl/fast_vgicp_impl.hpp:121-205  linearize / compute_error (sqrt(num_points)-weighted sums)
// Covariances, the LM / GN loop and the convergence test are FastGICP's, i.e. the code every variant shares.
// Layout: the voxel map is four arrays in ascending key order (key = x + y * dx + z * dx * dy over the box of occupied
// voxel coordinates): key u32, count i32, mean 3 x f64, covariance 6 x f64 (symmetric). Correspondences are a dense table
// [n_source x n_offsets] in ORIGINAL source order (voxel index or -1) with 6 x f64 Mahalanobis per slot.
#include <algorithm>

#include "point_math.cuh"

namespace apd {

namespace {

constexpr int kThreads = 256;

// GaussianVoxelMap::voxel_coord (fast_vgicp_voxel.hpp:160-162): floor(x / resolution - 0.5) in double
__device__ __forceinline__ int voxel_coord_d(double x, double res) { return (int)floor(__dsub_rn(__ddiv_rn(x, res), 0.5)); }

__device__ __forceinline__ bool finite3f(const float4& p) { return isfinite(p.x) && isfinite(p.y) && isfinite(p.z); }

__global__ void __launch_bounds__(kThreads) vgicp_keys_kernel(const float4* __restrict__ pts, int n, VoxelGridDesc vg, uint32_t* __restrict__ keys,
                                                              uint32_t* __restrict__ vals) {
  const int i = blockIdx.x * kThreads + threadIdx.x;
  if (i >= n) return;
  const float4 p = pts[i];
  uint32_t key = 0xffffffffu;  // (a non-finite point belongs to no voxel; it sorts to the end)
  if (finite3f(p)) {
    const int x = voxel_coord_d((double)p.x, vg.res) - vg.mn[0];
    const int y = voxel_coord_d((double)p.y, vg.res) - vg.mn[1];
    const int z = voxel_coord_d((double)p.z, vg.res) - vg.mn[2];
    key = (uint32_t)x + (uint32_t)y * (uint32_t)vg.dim[0] + (uint32_t)z * (uint32_t)vg.dim[0] * (uint32_t)vg.dim[1];
  }
  keys[i] = key;
  vals[i] = (uint32_t)i;
}

// One thread per voxel (the head of a run of equal keys): AdditiveGaussianVoxel / MultiplicativeGaussianVoxel append +
// finalize (fast_vgicp_voxel.hpp:79-124) over the members in ascending original index (the sort is stable).
__global__ void __launch_bounds__(kThreads) vgicp_voxels_kernel(const float4* __restrict__ pts, const int* __restrict__ inv_perm,
                                                                const double* __restrict__ cov, const uint32_t* __restrict__ keys,
                                                                const uint32_t* __restrict__ vals, const uint32_t* __restrict__ voxel_of, int n,
                                                                int multiplicative, uint32_t* __restrict__ vkey, int32_t* __restrict__ vcnt,
                                                                double* __restrict__ vmean, double* __restrict__ vcov) {
  const int i = blockIdx.x * kThreads + threadIdx.x;
  if (i >= n) return;
  const uint32_t k = keys[i];
  if (k == 0xffffffffu || (i > 0 && keys[i - 1] == k)) return;
  double m[3] = {0.0, 0.0, 0.0};
  Sym3 C;
#pragma unroll
  for (int e = 0; e < 6; e++) C.v[e] = 0.0;
  int cnt = 0;
  for (int j = i; j < n && keys[j] == k; j++) {
    const int oi = (int)vals[j];
    const float4 p = pts[oi];
    const double* c = cov + (size_t)inv_perm[oi] * 6;
    Sym3 ci;
#pragma unroll
    for (int e = 0; e < 6; e++) ci.v[e] = c[e];
    const double x = (double)p.x, y = (double)p.y, z = (double)p.z;
    if (!multiplicative) {  // :108-112
      m[0] += x; m[1] += y; m[2] += z;
#pragma unroll
      for (int e = 0; e < 6; e++) C.v[e] += ci.v[e];
    } else {  // :86-93 cov += cov_^-1, mean += cov_^-1 mean_
      const Sym3 inv = sym_inverse(ci);
#pragma unroll
      for (int e = 0; e < 6; e++) C.v[e] += inv.v[e];
      m[0] += (inv.v[0] * x + inv.v[1] * y) + inv.v[2] * z;
      m[1] += (inv.v[1] * x + inv.v[3] * y) + inv.v[4] * z;
      m[2] += (inv.v[2] * x + inv.v[4] * y) + inv.v[5] * z;
    }
    cnt++;
  }
  if (!multiplicative) {  // :114-117
    const double nn = (double)cnt;
#pragma unroll
    for (int a = 0; a < 3; a++) m[a] = __ddiv_rn(m[a], nn);
#pragma unroll
    for (int e = 0; e < 6; e++) C.v[e] = __ddiv_rn(C.v[e], nn);
  } else {  // :95-101
    C = sym_inverse(C);
    const double x = m[0], y = m[1], z = m[2];
    m[0] = (C.v[0] * x + C.v[1] * y) + C.v[2] * z;
    m[1] = (C.v[1] * x + C.v[3] * y) + C.v[4] * z;
    m[2] = (C.v[2] * x + C.v[4] * y) + C.v[5] * z;
  }
  const uint32_t v = voxel_of[i];
  vkey[v] = k;
  vcnt[v] = cnt;
#pragma unroll
  for (int a = 0; a < 3; a++) vmean[(size_t)v * 3 + a] = m[a];
#pragma unroll
  for (int e = 0; e < 6; e++) vcov[(size_t)v * 6 + e] = C.v[e];
}

// neighbor_offsets (fast_vgicp_voxel.hpp:10-44), in the reference's order
__device__ __forceinline__ void offset_of(int method, int o, int& dx, int& dy, int& dz) {
  if (method == kVoxelDirect1) { dx = dy = dz = 0; return; }
  if (method == kVoxelDirect7) {
    dx = o == 1 ? 1 : (o == 2 ? -1 : 0);
    dy = o == 3 ? 1 : (o == 4 ? -1 : 0);
    dz = o == 5 ? 1 : (o == 6 ? -1 : 0);
    return;
  }
  dx = o / 9 - 1; dy = (o / 3) % 3 - 1; dz = o % 3 - 1;
}

// GaussianVoxelMap::lookup_voxel (:169-176): binary search in the sorted keys
__device__ __forceinline__ int lookup_voxel(const uint32_t* __restrict__ vkey, int nv, const VoxelGridDesc& vg, int x, int y, int z) {
  x -= vg.mn[0]; y -= vg.mn[1]; z -= vg.mn[2];
  if (x < 0 || y < 0 || z < 0 || x >= vg.dim[0] || y >= vg.dim[1] || z >= vg.dim[2]) return -1;
  const uint32_t key = (uint32_t)x + (uint32_t)y * (uint32_t)vg.dim[0] + (uint32_t)z * (uint32_t)vg.dim[0] * (uint32_t)vg.dim[1];
  int lo = 0, hi = nv;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(&vkey[mid]) < key) lo = mid + 1;
    else hi = mid;
  }
  return (lo < nv && __ldg(&vkey[lo]) == key) ? lo : -1;
}

// trans * mean_A as Eigen evaluates it: ((r0 x + r1 y) + r2 z) + t, one rounding per operation
__device__ __forceinline__ void transform_d_rn(const PoseD& T, double x, double y, double z, double& ox, double& oy, double& oz) {
  ox = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(T.r[0], x), __dmul_rn(T.r[1], y)), __dmul_rn(T.r[2], z)), T.t[0]);
  oy = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(T.r[3], x), __dmul_rn(T.r[4], y)), __dmul_rn(T.r[5], z)), T.t[1]);
  oz = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(T.r[6], x), __dmul_rn(T.r[7], y)), __dmul_rn(T.r[8], z)), T.t[2]);
}

// FastVGICP::update_correspondences (fast_vgicp_impl.hpp:74-118): one thread per (sorted) source point, its offsets in turn
__global__ void __launch_bounds__(kThreads) vgicp_corr_kernel(const float4* __restrict__ s_spts, const double* __restrict__ s_cov, int n_src,
                                                              const uint32_t* __restrict__ vkey, const double* __restrict__ vcov, int nv,
                                                              VoxelGridDesc vg, int method, int n_off, PoseD T, int32_t* __restrict__ vcorr,
                                                              double* __restrict__ vmaha) {
  const int i = blockIdx.x * kThreads + threadIdx.x;
  if (i >= n_src) return;
  const float4 a = s_spts[i];
  const int oi = __float_as_int(a.w);
  double tx, ty, tz;
  transform_d_rn(T, (double)a.x, (double)a.y, (double)a.z, tx, ty, tz);  // :87
  const int cx = voxel_coord_d(tx, vg.res), cy = voxel_coord_d(ty, vg.res), cz = voxel_coord_d(tz, vg.res);  // :88
  // T cov_A T^T once per point (:111)
  Sym3 rcr0;
  {
    const double* c = s_cov + (size_t)i * 6;
    const double cam[9] = {c[0], c[1], c[2], c[1], c[3], c[4], c[2], c[4], c[5]};
    const double* Rt = T.r;
    double X[9];
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
      for (int cc = 0; cc < 3; cc++) X[r * 3 + cc] = Rt[r * 3 + 0] * cam[0 * 3 + cc] + Rt[r * 3 + 1] * cam[1 * 3 + cc] + Rt[r * 3 + 2] * cam[2 * 3 + cc];
    const int RR[6] = {0, 0, 0, 1, 1, 2}, CC[6] = {0, 1, 2, 1, 2, 2};
#pragma unroll
    for (int e = 0; e < 6; e++) rcr0.v[e] = X[RR[e] * 3 + 0] * Rt[CC[e] * 3 + 0] + X[RR[e] * 3 + 1] * Rt[CC[e] * 3 + 1] + X[RR[e] * 3 + 2] * Rt[CC[e] * 3 + 2];
  }
  for (int o = 0; o < n_off; o++) {
    int dx, dy, dz;
    offset_of(method, o, dx, dy, dz);
    const int v = lookup_voxel(vkey, nv, vg, cx + dx, cy + dy, cz + dz);  // :91
    const size_t slot = (size_t)oi * n_off + o;
    vcorr[slot] = v;
    if (v < 0) continue;
    Sym3 rcr;
#pragma unroll
    for (int e = 0; e < 6; e++) rcr.v[e] = __ldg(&vcov[(size_t)v * 6 + e]) + rcr0.v[e];  // :111 RCR = cov_B + T cov_A T^T
    const Sym3 M = sym_inverse(rcr);                                                      // :114
#pragma unroll
    for (int e = 0; e < 6; e++) vmaha[slot * 6 + e] = M.v[e];
  }
}

// One term of the sums (:141-169): e = mean_B - T a, w e^T M e, w J^T M J, w J^T M e with J = [skew(T a), -I] —
// the structure of point_math.cuh's accumulate_point with the weight folded into M.
template <bool kHB>
__device__ __forceinline__ void vgicp_accumulate(double* acc, double x, double y, double z, double e0, double e1, double e2, const double* m) {
  const double me0 = (m[0] * e0 + m[1] * e1) + m[2] * e2;
  const double me1 = (m[1] * e0 + m[3] * e1) + m[4] * e2;
  const double me2 = (m[2] * e0 + m[4] * e1) + m[5] * e2;
  const double q = (e0 * me0 + e1 * me1) + e2 * me2;
  if (!kHB) {
    acc[0] += q;
    return;
  }
  acc[27] += q;
  const double n00 = m[1] * z - m[2] * y, n01 = m[2] * x - m[0] * z, n02 = m[0] * y - m[1] * x;
  const double n10 = m[3] * z - m[4] * y, n11 = m[4] * x - m[1] * z, n12 = m[1] * y - m[3] * x;
  const double n20 = m[4] * z - m[5] * y, n21 = m[5] * x - m[2] * z, n22 = m[2] * y - m[4] * x;
  acc[0] += z * n10 - y * n20;
  acc[1] += z * n11 - y * n21;
  acc[2] += z * n12 - y * n22;
  acc[6] += x * n21 - z * n01;
  acc[7] += x * n22 - z * n02;
  acc[11] += y * n02 - x * n12;
  acc[3] -= n00;  acc[4] -= n10;  acc[5] -= n20;
  acc[8] -= n01;  acc[9] -= n11;  acc[10] -= n21;
  acc[12] -= n02; acc[13] -= n12; acc[14] -= n22;
  acc[15] += m[0]; acc[16] += m[1]; acc[17] += m[2];
  acc[18] += m[3]; acc[19] += m[4];
  acc[20] += m[5];
  acc[21] += z * me1 - y * me2;
  acc[22] += x * me2 - z * me0;
  acc[23] += y * me0 - x * me1;
  acc[24] -= me0; acc[25] -= me1; acc[26] -= me2;
}

// linearize (:141-181) / compute_error (:186-205) over the correspondence table: per-thread fp64 sums over a grid-stride
// walk -> fixed shuffle tree -> warps in order -> per-block partials -> the last block (ticket) adds them in block order.
template <bool kHB>
__global__ void __launch_bounds__(kThreads) vgicp_reduce_kernel(const float4* __restrict__ s_pts, const int32_t* __restrict__ vcorr,
                                                                const double* __restrict__ vmaha, const int32_t* __restrict__ vcnt,
                                                                const double* __restrict__ vmean, long long slots, int n_off, PoseD T,
                                                                double* __restrict__ partials, double* __restrict__ out28,
                                                                unsigned int* __restrict__ ticket) {
  constexpr int NV = kHB ? kReduceVals : 1;
  double acc[NV];
#pragma unroll
  for (int j = 0; j < NV; j++) acc[j] = 0.0;
  for (long long s = (long long)blockIdx.x * kThreads + threadIdx.x; s < slots; s += (long long)gridDim.x * kThreads) {
    const int v = vcorr[s];
    if (v < 0) continue;
    const float4 a = s_pts[(int)(s / n_off)];
    // (plain expressions: the compiler may contract them — the sums are held to a tolerance, not to the bit)
    const double ax = (double)a.x, ay = (double)a.y, az = (double)a.z;
    const double x = ((T.r[0] * ax + T.r[1] * ay) + T.r[2] * az) + T.t[0];
    const double y = ((T.r[3] * ax + T.r[4] * ay) + T.r[5] * az) + T.t[1];
    const double z = ((T.r[6] * ax + T.r[7] * ay) + T.r[8] * az) + T.t[2];
    const double w = sqrt((double)vcnt[v]);  // :154
    double m[6];
#pragma unroll
    for (int e = 0; e < 6; e++) m[e] = w * vmaha[(size_t)s * 6 + e];
    vgicp_accumulate<kHB>(acc, x, y, z, vmean[(size_t)v * 3 + 0] - x, vmean[(size_t)v * 3 + 1] - y, vmean[(size_t)v * 3 + 2] - z, m);
  }
  __shared__ double sh[kThreads / 32][NV];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < NV; j++) {
    double v = acc[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) sh[warp][j] = v;
  }
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x < NV) {
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; w++) v += sh[w][threadIdx.x];
    partials[(size_t)blockIdx.x * NV + threadIdx.x] = v;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(ticket, 1u);
    last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (last) {
    __threadfence();
    if (threadIdx.x < NV) {
      double v = 0.0;
      for (unsigned int b = 0; b < gridDim.x; b++) v += __ldcg(&partials[(size_t)b * NV + threadIdx.x]);
      out28[kHB ? threadIdx.x : 27] = v;
    }
    if (threadIdx.x == 0) *ticket = 0;
  }
}

// parity hooks: the voxel list as coordinates / 3x3 matrices, the correspondence table's Mahalanobis as 3x3
__global__ void __launch_bounds__(kThreads) vgicp_export_voxels_kernel(const uint32_t* __restrict__ vkey, const double* __restrict__ vcov, int nv,
                                                                       VoxelGridDesc vg, int32_t* __restrict__ coords, double* __restrict__ covs) {
  const int v = blockIdx.x * kThreads + threadIdx.x;
  if (v >= nv) return;
  if (coords) {
    const uint32_t k = vkey[v];
    const uint32_t d0 = (uint32_t)vg.dim[0], d01 = d0 * (uint32_t)vg.dim[1];
    coords[3 * v + 2] = (int)(k / d01) + vg.mn[2];
    coords[3 * v + 1] = (int)((k % d01) / d0) + vg.mn[1];
    coords[3 * v + 0] = (int)(k % d0) + vg.mn[0];
  }
  if (covs) {
    const double* c = vcov + (size_t)v * 6;
    double* o = covs + (size_t)v * 9;
    o[0] = c[0]; o[1] = c[1]; o[2] = c[2]; o[3] = c[1]; o[4] = c[3]; o[5] = c[4]; o[6] = c[2]; o[7] = c[4]; o[8] = c[5];
  }
}
__global__ void __launch_bounds__(kThreads) vgicp_export_maha_kernel(const int32_t* __restrict__ vcorr, const double* __restrict__ vmaha, long long slots,
                                                                     double* __restrict__ out) {
  const long long s = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (s >= slots) return;
  double* o = out + (size_t)s * 9;
  if (vcorr[s] < 0) {
#pragma unroll
    for (int e = 0; e < 9; e++) o[e] = 0.0;
    return;
  }
  const double* c = vmaha + (size_t)s * 6;
  o[0] = c[0]; o[1] = c[1]; o[2] = c[2]; o[3] = c[1]; o[4] = c[3]; o[5] = c[4]; o[6] = c[2]; o[7] = c[4]; o[8] = c[5];
}

}  // namespace

void launch_vgicp_keys(const float4* pts, int n, const VoxelGridDesc& vg, uint32_t* keys, uint32_t* vals, cudaStream_t s, int64_t* launches) {
  vgicp_keys_kernel<<<(n + kThreads - 1) / kThreads, kThreads, 0, s>>>(pts, n, vg, keys, vals);
  (*launches)++;
}
void launch_vgicp_voxels(const CloudDev& tgt, const uint32_t* keys, const uint32_t* vals, const uint32_t* voxel_of, int multiplicative,
                         uint32_t* vkey, int32_t* vcnt, double* vmean, double* vcov, cudaStream_t s, int64_t* launches) {
  vgicp_voxels_kernel<<<(tgt.n + kThreads - 1) / kThreads, kThreads, 0, s>>>(tgt.pts, tgt.inv_perm, tgt.cov, keys, vals, voxel_of, tgt.n, multiplicative,
                                                                             vkey, vcnt, vmean, vcov);
  (*launches)++;
}
void launch_vgicp_correspondences(const CloudDev& src, const VoxelMapDev& vm, int method, int n_off, const PoseD& T, int32_t* vcorr, double* vmaha,
                                  cudaStream_t s, int64_t* launches) {
  if (src.n <= 0) return;
  vgicp_corr_kernel<<<(src.n + kThreads - 1) / kThreads, kThreads, 0, s>>>(src.spts, src.cov, src.n, vm.key, vm.cov, vm.n, vm.g, method, n_off, T, vcorr,
                                                                           vmaha);
  (*launches)++;
}
void launch_vgicp_reduce(const CloudDev& src, const VoxelMapDev& vm, const int32_t* vcorr, const double* vmaha, int n_off, const PoseD& T, bool want_hb,
                         double* partials, int max_blocks, double* d_out28, unsigned int* ticket, cudaStream_t s, int64_t* launches) {
  const long long slots = (long long)src.n * n_off;
  int blocks = (int)std::min<long long>((slots + kThreads - 1) / kThreads, (long long)max_blocks);
  blocks = std::max(blocks, 1);
  if (want_hb) vgicp_reduce_kernel<true><<<blocks, kThreads, 0, s>>>(src.pts, vcorr, vmaha, vm.cnt, vm.mean, slots, n_off, T, partials, d_out28, ticket);
  else vgicp_reduce_kernel<false><<<blocks, kThreads, 0, s>>>(src.pts, vcorr, vmaha, vm.cnt, vm.mean, slots, n_off, T, partials, d_out28, ticket);
  (*launches)++;
}
void launch_vgicp_export_voxels(const VoxelMapDev& vm, int32_t* d_coords, double* d_covs, cudaStream_t s, int64_t* launches) {
  if (vm.n <= 0) return;
  vgicp_export_voxels_kernel<<<(vm.n + kThreads - 1) / kThreads, kThreads, 0, s>>>(vm.key, vm.cov, vm.n, vm.g, d_coords, d_covs);
  (*launches)++;
}
void launch_vgicp_export_maha(const int32_t* vcorr, const double* vmaha, long long slots, double* d_out, cudaStream_t s, int64_t* launches) {
  if (slots <= 0) return;
  vgicp_export_maha_kernel<<<(unsigned)((slots + kThreads - 1) / kThreads), kThreads, 0, s>>>(vcorr, vmaha, slots, d_out);
  (*launches)++;
}

void preload_vgicp_kernels() {
  cudaFuncAttributes a;
  (void)cudaFuncGetAttributes(&a, vgicp_keys_kernel);
  (void)cudaFuncGetAttributes(&a, vgicp_voxels_kernel);
  (void)cudaFuncGetAttributes(&a, vgicp_corr_kernel);
  (void)cudaFuncGetAttributes(&a, vgicp_reduce_kernel<true>);
  (void)cudaFuncGetAttributes(&a, vgicp_reduce_kernel<false>);
  (void)cudaFuncGetAttributes(&a, vgicp_export_voxels_kernel);
  (void)cudaFuncGetAttributes(&a, vgicp_export_maha_kernel);
}

}  // namespace apd
