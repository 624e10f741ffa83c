// K3 — update_correspondences: exact 1-NN of every transformed source point on
// the target grid, the radar range/azimuth/elevation noise covariance at that
// point, the combined covariance and its inverse (the per-point Mahalanobis
// matrix). Replaces FastAPDGICP::update_correspondences (reference
// fast_apdgicp_impl.hpp:160-220). Also the fitness pass (pcl getFitnessScore)
// and the export hooks.
#include "kernels.cuh"

namespace apd {

namespace {

constexpr int kThreads = 128;
constexpr unsigned long long kInfKey = 0xffffffffffffffffull;

// scan one contiguous range of the cell-sorted target points
__device__ __forceinline__ void scan_range(const float4* __restrict__ spts, int b, int e, float qx, float qy, float qz,
                                           unsigned long long& best, int& best_pos) {
  for (int j = b; j < e; j++) {
    const float4 p = __ldg(&spts[j]);
    const unsigned long long key = pack_key(sqdist_rn(qx, qy, qz, p.x, p.y, p.z), __float_as_int(p.w));
    if (key < best) {
      best = key;
      best_pos = j;
    }
  }
}

// Exact nearest neighbour by (d2, original index). Expands Chebyshev shells of
// cells until the best distance is provably final, or until every unscanned
// point is farther than `limit` (then the caller rejects the match anyway).
// G consecutive lanes share one query (G = 1, 2, 4, 8, 16, 32): the x-rows of a
// shell are dealt round-robin to the G lanes and the group's best key is
// min-reduced with shuffles after every shell, so the critical path of a query is
// ~1/G of the single-thread scan (small source clouds are latency-bound).
// All G lanes return the same result.
template <int G>
__device__ __forceinline__ void nn_search(const float4* __restrict__ spts, const uint32_t* __restrict__ cell_start, const GridDesc& g,
                                          float qx, float qy, float qz, double limit_sq, unsigned long long& best, int& best_pos) {
  const int sub = (G == 1) ? 0 : (int)(threadIdx.x & (G - 1));
  const int cx = cell_coord(qx, g.ox, g.inv_cell, g.nx);
  const int cy = cell_coord(qy, g.oy, g.inv_cell, g.ny);
  const int cz = cell_coord(qz, g.oz, g.inv_cell, g.nz);
  best = kInfKey;
  best_pos = -1;
  // groups of one warp leave the shell loop at different times: shuffle within the group's own lanes only
  const unsigned gmask = (G >= 32) ? 0xffffffffu : (((1u << G) - 1u) << ((threadIdx.x & 31) & ~(G - 1)));
  auto group_min = [&]() {
    if (G > 1) {
#pragma unroll
      for (int o = G >> 1; o > 0; o >>= 1) {
        const unsigned long long ob = __shfl_xor_sync(gmask, best, o);
        const int op = __shfl_xor_sync(gmask, best_pos, o);
        if (ob < best) {
          best = ob;
          best_pos = op;
        }
      }
    }
  };
  // ring 0+1: 3x3x3 cube as 9 x-rows
  {
    const int x0 = max(cx - 1, 0), x1 = min(cx + 1, g.nx - 1);
    for (int ri = sub; ri < 9; ri += G) {
      const int y = cy + (ri % 3) - 1, z = cz + (ri / 3) - 1;
      if (y < 0 || y >= g.ny || z < 0 || z >= g.nz) continue;
      const int row = (z * g.ny + y) * g.nx;
      scan_range(spts, (int)__ldg(&cell_start[row + x0]), (int)__ldg(&cell_start[row + x1 + 1]), qx, qy, qz, best, best_pos);
    }
    group_min();
  }
  const float mg = 0.002f * g.cell;
  // thick shells (r, rr]: one cell at a time near the query, then growing ~1.5x (see knn_cov.cu)
  for (int r = 1;;) {
    const float lb = ((float)r - 0.002f) * g.cell;
    const float lb2 = lb * lb;
    if (best != kInfKey && __uint_as_float((unsigned)(best >> 32)) < lb2) break;
    if ((double)lb2 >= limit_sq) break;
    if (cx - r <= 0 && cx + r >= g.nx - 1 && cy - r <= 0 && cy + r >= g.ny - 1 && cz - r <= 0 && cz + r >= g.nz - 1) break;
    const int rr = r < 3 ? r + 1 : r + (r >> 1) + 1;
    const int side = 2 * rr + 1;
    const int x0 = max(cx - rr, 0), x1 = min(cx + rr, g.nx - 1);
    for (int ri = sub; ri < side * side; ri += G) {
      const int dy = ri % side - rr, dz = ri / side - rr;
      const int y = cy + dy, z = cz + dz;
      if (y < 0 || y >= g.ny || z < 0 || z >= g.nz) continue;
      // skip the row if even its nearest point cannot beat the current best / the limit
      const float loy = g.oy + (float)y * g.cell - mg, hiy = g.oy + (float)(y + 1) * g.cell + mg;
      const float loz = g.oz + (float)z * g.cell - mg, hiz = g.oz + (float)(z + 1) * g.cell + mg;
      const float ddy = fmaxf(0.f, fmaxf(loy - qy, qy - hiy)), ddz = fmaxf(0.f, fmaxf(loz - qz, qz - hiz));
      const float dyz2 = (ddy * ddy + ddz * ddz) * 0.9999f;
      if ((double)dyz2 >= limit_sq) continue;
      if (best != kInfKey && dyz2 > __uint_as_float((unsigned)(best >> 32))) continue;
      const int row = (z * g.ny + y) * g.nx;
      if (dy > r || dy < -r || dz > r || dz < -r) {  // row outside the scanned cube: its whole x-range
        scan_range(spts, (int)__ldg(&cell_start[row + x0]), (int)__ldg(&cell_start[row + x1 + 1]), qx, qy, qz, best, best_pos);
      } else {  // row crosses the scanned cube: the two end pieces
        const int xl = min(cx - r - 1, g.nx - 1), xr = max(cx + r + 1, 0);
        if (x0 <= xl) scan_range(spts, (int)__ldg(&cell_start[row + x0]), (int)__ldg(&cell_start[row + xl + 1]), qx, qy, qz, best, best_pos);
        if (xr <= x1) scan_range(spts, (int)__ldg(&cell_start[row + xr]), (int)__ldg(&cell_start[row + x1 + 1]), qx, qy, qz, best, best_pos);
      }
    }
    group_min();
    r = rr;
  }
}

__device__ __forceinline__ PoseF pose_to_f32(const PoseD& T) {
  PoseF f;
#pragma unroll
  for (int i = 0; i < 9; i++) f.r[i] = (float)T.r[i];  // Isometry3d::cast<float>() (:164)
#pragma unroll
  for (int i = 0; i < 3; i++) f.t[i] = (float)T.t[i];
  return f;
}

template <bool kFp64, int G>
__global__ void __launch_bounds__(kThreads) update_corr_kernel(const float4* __restrict__ s_spts, const float* __restrict__ s_label,
                                                               const double* __restrict__ s_cov, int n_src,
                                                               const float4* __restrict__ t_spts, const float* __restrict__ t_label,
                                                               const double* __restrict__ t_cov, const uint32_t* __restrict__ t_cell_start,
                                                               GridDesc tg, PoseD T, NoiseParams np, int* __restrict__ corr,
                                                               float* __restrict__ sqd, void* __restrict__ mahaA, void* __restrict__ mahaB) {
  // G lanes per source point (kThreads and the warp size are multiples of G, so a group never straddles a warp;
  // the grid is sized so that whole warps are either in range or carry clamped duplicates of the last point)
  const int gi = (blockIdx.x * kThreads + threadIdx.x) / G;
  const int i = min(gi, n_src - 1);
  const PoseF Tf = pose_to_f32(T);
  const float4 a = s_spts[i];
  float px, py, pz;
  transform_rn(Tf, a.x, a.y, a.z, px, py, pz);  // :176

  unsigned long long best;
  int pos;
  nn_search<G>(t_spts, t_cell_start, tg, px, py, pz, np.thr_sq, best, pos);  // :178
  if (gi >= n_src || (G > 1 && (threadIdx.x & (G - 1)) != 0)) return;  // one lane per point finishes the job
  const float d2 = (best == kInfKey) ? 3.402823466e38f : __uint_as_float((unsigned)(best >> 32));
  sqd[i] = d2;  // :180
  const bool ok = (best != kInfKey) && ((double)d2 < np.thr_sq);  // :183
  if (!ok) {
    corr[i] = -1;
    return;
  }
  corr[i] = pos | ((t_label[pos] == s_label[i]) ? kCorrLabelBit : 0);  // label test of :271-273, hoisted

  // radar noise covariance at the transformed point (:194-210)
  const double dpx = (double)px, dpy = (double)py, dpz = (double)pz;
  const double dist = sqrt(dpx * dpx + dpy * dpy + dpz * dpz);
  const double s_x = dist * np.dist_var / 400;
  const double s_y = dist * np.sin_az;
  const double s_z = dist * np.sin_el;
  const float rho_xy = __fsqrt_rn(__fadd_rn(__fmul_rn(px, px), __fmul_rn(py, py)));
  // float-valued angles as in the reference (atan2f); evaluated in double and rounded to float
  const double elevation = (double)(float)atan2((double)rho_xy, dpz);
  const double azimuth = (double)(float)atan2(dpy, dpx);
  double sz_, cz_, sy_, cy_;
  sincos(azimuth * 0.5, &sz_, &cz_);
  sincos(elevation * 0.5, &sy_, &cy_);
  // quaternion of AngleAxis(az, Z) * AngleAxis(el, Y) -> rotation matrix (Eigen toRotationMatrix)
  const double qw = cz_ * cy_, qx = -(sz_ * sy_), qy = cz_ * sy_, qz = sz_ * cy_;
  const double tx = 2.0 * qx, ty = 2.0 * qy, tz = 2.0 * qz;
  const double twx = tx * qw, twy = ty * qw, twz = tz * qw;
  const double txx = tx * qx, txy = ty * qx, txz = tz * qx;
  const double tyy = ty * qy, tyz = tz * qy, tzz = tz * qz;
  double R[9];
  R[0] = 1.0 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
  R[3] = txy + twz; R[4] = 1.0 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1.0 - (txx + tyy);
  const double sc[3] = {s_x, s_y, s_z};
  double A[9];
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int c = 0; c < 3; c++) A[r * 3 + c] = R[r * 3 + c] * sc[c];
  Sym3 cr;  // cov_r = A A^T
  cr.v[0] = A[0] * A[0] + A[1] * A[1] + A[2] * A[2];
  cr.v[1] = A[0] * A[3] + A[1] * A[4] + A[2] * A[5];
  cr.v[2] = A[0] * A[6] + A[1] * A[7] + A[2] * A[8];
  cr.v[3] = A[3] * A[3] + A[4] * A[4] + A[5] * A[5];
  cr.v[4] = A[3] * A[6] + A[4] * A[7] + A[5] * A[8];
  cr.v[5] = A[6] * A[6] + A[7] * A[7] + A[8] * A[8];

  // RCR = (cov_B + cov_r) + T (cov_A + cov_r) T^T (:213-215), 3x3 block
  Sym3 ca, cb;
#pragma unroll
  for (int e = 0; e < 6; e++) {
    ca.v[e] = s_cov[(size_t)i * 6 + e] + cr.v[e];
    cb.v[e] = t_cov[(size_t)pos * 6 + e] + cr.v[e];
  }
  // X = Rt * ca (3x3 full), then RCR = cb + X * Rt^T
  const double* Rt = T.r;
  const double cam[9] = {ca.v[0], ca.v[1], ca.v[2], ca.v[1], ca.v[3], ca.v[4], ca.v[2], ca.v[4], ca.v[5]};
  double X[9];
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int c = 0; c < 3; c++) X[r * 3 + c] = Rt[r * 3 + 0] * cam[0 * 3 + c] + Rt[r * 3 + 1] * cam[1 * 3 + c] + Rt[r * 3 + 2] * cam[2 * 3 + c];
  Sym3 rcr;
  const int RR[6] = {0, 0, 0, 1, 1, 2}, CC[6] = {0, 1, 2, 1, 2, 2};
#pragma unroll
  for (int e = 0; e < 6; e++) {
    const int r = RR[e], c = CC[e];
    rcr.v[e] = cb.v[e] + (X[r * 3 + 0] * Rt[c * 3 + 0] + X[r * 3 + 1] * Rt[c * 3 + 1] + X[r * 3 + 2] * Rt[c * 3 + 2]);
  }
  const Sym3 M = sym_inverse(rcr);  // :217-218
  if (kFp64) {
    double2* mA = reinterpret_cast<double2*>(mahaA);
    double2* mB = reinterpret_cast<double2*>(mahaB);
    mA[i] = make_double2(M.v[0], M.v[1]);
    mB[i] = make_double2(M.v[2], M.v[3]);
    mB[(size_t)n_src + i] = make_double2(M.v[4], M.v[5]);
  } else {
    reinterpret_cast<float4*>(mahaA)[i] = make_float4((float)M.v[0], (float)M.v[1], (float)M.v[2], (float)M.v[3]);
    reinterpret_cast<float2*>(mahaB)[i] = make_float2((float)M.v[4], (float)M.v[5]);
  }
}

// ---- fitness -------------------------------------------------------------------
constexpr int kFitThreads = 128;
__global__ void __launch_bounds__(kFitThreads) fitness_kernel(const float4* __restrict__ s_spts, int n_src, const float4* __restrict__ t_spts,
                                                              const uint32_t* __restrict__ t_cell_start, GridDesc tg, PoseF Tf,
                                                              double max_range, double inlier_sq_thr, double* __restrict__ partials,
                                                              double* __restrict__ out3, unsigned int* __restrict__ ticket) {
  double sum = 0.0, nr = 0.0, ni = 0.0;
  for (int i = blockIdx.x * kFitThreads + threadIdx.x; i < n_src; i += gridDim.x * kFitThreads) {
    const float4 a = s_spts[i];
    float px, py, pz;
    transform_rn(Tf, a.x, a.y, a.z, px, py, pz);
    unsigned long long best;
    int pos;
    nn_search<1>(t_spts, t_cell_start, tg, px, py, pz, 1e300, best, pos);
    if (best != kInfKey) {
      const double d2 = (double)__uint_as_float((unsigned)(best >> 32));
      if (d2 <= max_range) { sum += d2; nr += 1.0; }
      if (d2 < inlier_sq_thr) ni += 1.0;
    }
  }
  __shared__ double sh[3][kFitThreads / 32];
  sum = warp_sum(sum); nr = warp_sum(nr); ni = warp_sum(ni);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { sh[0][warp] = sum; sh[1][warp] = nr; sh[2][warp] = ni; }
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) {
    double a = 0, b = 0, c = 0;
    for (int w = 0; w < kFitThreads / 32; w++) { a += sh[0][w]; b += sh[1][w]; c += sh[2][w]; }
    partials[blockIdx.x * 3 + 0] = a; partials[blockIdx.x * 3 + 1] = b; partials[blockIdx.x * 3 + 2] = c;
    __threadfence();
    const unsigned int t = atomicAdd(ticket, 1u);
    last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    double a = 0, b = 0, c = 0;
    for (unsigned int k = 0; k < gridDim.x; k++) {
      a += __ldcg(&partials[k * 3 + 0]); b += __ldcg(&partials[k * 3 + 1]); c += __ldcg(&partials[k * 3 + 2]);
    }
    out3[0] = a; out3[1] = b; out3[2] = c;
    *ticket = 0;
  }
}

// ---- export hooks ----------------------------------------------------------------
template <bool kFp64>
__global__ void __launch_bounds__(256) corr_export_kernel(const float4* __restrict__ s_spts, int n_src, const float4* __restrict__ t_spts,
                                                          const int* __restrict__ corr, const float* __restrict__ sqd,
                                                          const void* __restrict__ mahaA, const void* __restrict__ mahaB,
                                                          int32_t* __restrict__ idx_out, float* __restrict__ sqd_out, double* __restrict__ maha_out) {
  const int s = blockIdx.x * 256 + threadIdx.x;
  if (s >= n_src) return;
  const int oi = __float_as_int(s_spts[s].w);
  const int c = corr[s];
  if (idx_out) idx_out[oi] = c < 0 ? -1 : __float_as_int(t_spts[c & kCorrIndexMask].w);
  if (sqd_out) sqd_out[oi] = sqd[s];
  if (maha_out) {
    double m[6] = {0, 0, 0, 0, 0, 0};
    if (c >= 0) {
      if (kFp64) {
        const double2 a = reinterpret_cast<const double2*>(mahaA)[s];
        const double2 b = reinterpret_cast<const double2*>(mahaB)[s];
        const double2 d = reinterpret_cast<const double2*>(mahaB)[(size_t)n_src + s];
        m[0] = a.x; m[1] = a.y; m[2] = b.x; m[3] = b.y; m[4] = d.x; m[5] = d.y;
      } else {
        const float4 a = reinterpret_cast<const float4*>(mahaA)[s];
        const float2 b = reinterpret_cast<const float2*>(mahaB)[s];
        m[0] = a.x; m[1] = a.y; m[2] = a.z; m[3] = a.w; m[4] = b.x; m[5] = b.y;
      }
    }
    double* o = maha_out + (size_t)oi * 16;
    o[0] = m[0]; o[1] = m[1]; o[2] = m[2]; o[3] = 0.0;
    o[4] = m[1]; o[5] = m[3]; o[6] = m[4]; o[7] = 0.0;
    o[8] = m[2]; o[9] = m[4]; o[10] = m[5]; o[11] = 0.0;
    o[12] = 0.0; o[13] = 0.0; o[14] = 0.0; o[15] = 0.0;
  }
}

__global__ void __launch_bounds__(256) transform_cloud_kernel(const float4* __restrict__ pts, int n, PoseF T, float* __restrict__ xyz) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const float4 p = pts[i];
  float x, y, z;
  transform_rn(T, p.x, p.y, p.z, x, y, z);
  xyz[3 * (size_t)i + 0] = x;
  xyz[3 * (size_t)i + 1] = y;
  xyz[3 * (size_t)i + 2] = z;
}

}  // namespace

void launch_update_correspondences(const CloudDev& src, const CloudDev& tgt, const PoseD& T, const NoiseParams& np,
                                   const CorrOut& out, cudaStream_t s, int64_t* launches) {
  if (src.n <= 0) return;
#define APD_CORR(FP64, GG)                                                                                                              \
  update_corr_kernel<FP64, GG><<<(unsigned)(((size_t)src.n * GG + kThreads - 1) / kThreads), kThreads, 0, s>>>(                           \
      src.spts, src.label, src.cov, src.n, tgt.spts, tgt.label, tgt.cov, tgt.cell_start, tgt.g, T, np, out.corr, out.sqd, out.mahaA, \
      out.mahaB)
  // small source clouds are latency-bound: 8 lanes share a query; large ones are throughput-bound: one lane per query
  const bool wide = src.n < 65536;
  if (out.maha_fp64) {
    if (wide) APD_CORR(true, 8);
    else APD_CORR(true, 1);
  } else {
    if (wide) APD_CORR(false, 8);
    else APD_CORR(false, 1);
  }
#undef APD_CORR
  (*launches)++;
}

void launch_fitness(const CloudDev& src, const CloudDev& tgt, const PoseF& T, double max_range, double inlier_sq_thr,
                    double* d_partials, int max_blocks, double* d_out3, unsigned int* d_ticket, cudaStream_t s, int64_t* launches) {
  int blocks = (src.n + kFitThreads - 1) / kFitThreads;
  blocks = max(1, min(blocks, max_blocks));
  fitness_kernel<<<blocks, kFitThreads, 0, s>>>(src.spts, src.n, tgt.spts, tgt.cell_start, tgt.g, T, max_range, inlier_sq_thr, d_partials,
                                                d_out3, d_ticket);
  (*launches)++;
}

void launch_corr_export(const CloudDev& src, const CloudDev& tgt, const CorrOut& c, int32_t* d_idx, float* d_sqd, double* d_maha4x4,
                        cudaStream_t s, int64_t* launches) {
  if (src.n <= 0) return;
  const int blocks = (src.n + 255) / 256;
  if (c.maha_fp64)
    corr_export_kernel<true><<<blocks, 256, 0, s>>>(src.spts, src.n, tgt.spts, c.corr, c.sqd, c.mahaA, c.mahaB, d_idx, d_sqd, d_maha4x4);
  else
    corr_export_kernel<false><<<blocks, 256, 0, s>>>(src.spts, src.n, tgt.spts, c.corr, c.sqd, c.mahaA, c.mahaB, d_idx, d_sqd, d_maha4x4);
  (*launches)++;
}

void launch_transform_cloud(const float4* pts, int n, const PoseF& T, float* d_xyz, cudaStream_t s, int64_t* launches) {
  if (n <= 0) return;
  transform_cloud_kernel<<<(n + 255) / 256, 256, 0, s>>>(pts, n, T, d_xyz);
  (*launches)++;
}

}  // namespace apd
