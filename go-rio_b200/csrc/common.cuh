// Shared device/host definitions of the B200 FastAPDGICP library.
// sm_100a only; no multi-arch dispatch, no CPU fallback.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "grid_desc.hpp"

namespace apd {

// ---------------------------------------------------------------------------
// Data layout in HBM.
//  A cloud of n points is kept in CELL-SORTED order (stable sort by uniform-grid
//  cell id, ties by original index, so the layout is deterministic):
//    spts[n]    float4 {x, y, z, bits(original index)}                16 B/pt
//    label[n]   float   cluster label (pcl normal_x)                   4 B/pt
//    cov[n]     6 x double, symmetric {xx,xy,xz,yy,yz,zz}             48 B/pt
//    geo[n]     float   sigma3/sigma1 of the regularised covariance    4 B/pt
//    geo64[n]   double  the same, unrounded (fp64-storage mode only)   8 B/pt
//    inv_perm[n] int    original index -> sorted position
//  plus the grid: cell_start[ncells+1] (uint32 prefix of per-cell counts, cells
//  linearised x-fastest so an x-row of cells is ONE contiguous point range).
//  Per source point and linearisation (sorted order of the source):
//    corr[n]    int32: sorted position of the matched target point, bit 30 set
//               when the two cluster labels are equal, -1 for "none"   4 B/pt
//    sqd[n]     float  1-NN squared distance                           4 B/pt
//    maha       symmetric {xx,xy,xz,yy | yz,zz}: float4 + float2 planes
//               (24 B/pt) or double2 x 3 planes (48 B/pt)
// ---------------------------------------------------------------------------

// struct GridDesc {ox, oy, oz, inv_cell, cell, nx, ny, nz}: grid_desc.hpp (included below the layout notes)

constexpr int kCorrLabelBit = 1 << 30;
constexpr int kCorrIndexMask = kCorrLabelBit - 1;

// cell coordinate of a point; the SAME expression is used when the grid is built
// and when it is queried (monotone in x, which the search bound relies on).
__host__ __device__ __forceinline__ int cell_coord(float x, float o, float inv_cell, int n) {
#ifdef __CUDA_ARCH__
  float f = floorf(__fmul_rn(__fsub_rn(x, o), inv_cell));
#else
  float f = floorf((x - o) * inv_cell);
#endif
  // NaN / huge values clamp into the grid
  int c = (f >= (float)n) ? n - 1 : (f > 0.f ? (int)f : 0);
  return c;
}

// fp32 squared distance in the reference's operation order (FLANN L2_Simple:
// ((dx*dx)+(dy*dy))+(dz*dz)), every operation rounded once (no FMA contraction),
// so that the argmin / k-th element agree bit-for-bit with the CPU path.
__device__ __forceinline__ float sqdist_rn(float ax, float ay, float az, float bx, float by, float bz) {
  const float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// (d2, index) packed so that unsigned comparison = lexicographic (d2, index);
// d2 >= 0 so its IEEE bits order like the value.
__device__ __forceinline__ unsigned long long pack_key(float d2, int idx) {
  return ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned int)idx;
}

// Eigen Isometry3f * Vector4f (w = 1): ((r0*x + r1*y) + r2*z) + t per row, fp32,
// no contraction (reference fast_apdgicp_impl.hpp:176).
struct PoseF {
  float r[9];
  float t[3];
};
struct PoseD {
  double r[9];
  double t[3];
};
__device__ __forceinline__ void transform_rn(const PoseF& T, float x, float y, float z, float& ox, float& oy, float& oz) {
  ox = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T.r[0], x), __fmul_rn(T.r[1], y)), __fmul_rn(T.r[2], z)), T.t[0]);
  oy = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T.r[3], x), __fmul_rn(T.r[4], y)), __fmul_rn(T.r[5], z)), T.t[1]);
  oz = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(T.r[6], x), __fmul_rn(T.r[7], y)), __fmul_rn(T.r[8], z)), T.t[2]);
}

// ---- small fp64 symmetric 3x3 helpers (device) -----------------------------
// symmetric storage order: 0 xx, 1 xy, 2 xz, 3 yy, 4 yz, 5 zz
struct Sym3 {
  double v[6];
};

// Symmetric 3x3 eigen-decomposition, cyclic Jacobi, fp64; eigenvalues sorted by
// |value| descending, V columns = eigenvectors. Same algorithm and stopping rule
// as the CPU oracle's svd3_sym so both agree to rounding.
__device__ __forceinline__ void jacobi_eig3(const Sym3& A, double l[3], double V[9]) {
  double a00 = A.v[0], a01 = A.v[1], a02 = A.v[2], a11 = A.v[3], a12 = A.v[4], a22 = A.v[5];
  double v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
#pragma unroll 1
  for (int sweep = 0; sweep < 64; sweep++) {
    const double off = fabs(a01) + fabs(a02) + fabs(a12);
    const double diag = fabs(a00) + fabs(a11) + fabs(a22);
    if (off <= 1e-300 || off <= 1e-22 * diag) break;
    // (p,q) = (0,1)
    if (a01 != 0.0) {
      const double theta = (a11 - a00) / (2.0 * a01);
      const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
      const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
      const double n00 = a00 - t * a01, n11 = a11 + t * a01;
      const double n02 = c * a02 - s * a12, n12 = s * a02 + c * a12;
      a00 = n00; a11 = n11; a01 = 0.0; a02 = n02; a12 = n12;
#pragma unroll
      for (int k = 0; k < 3; k++) {
        const double vp = v[k][0], vq = v[k][1];
        v[k][0] = c * vp - s * vq;
        v[k][1] = s * vp + c * vq;
      }
    }
    // (p,q) = (0,2)
    if (a02 != 0.0) {
      const double theta = (a22 - a00) / (2.0 * a02);
      const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
      const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
      const double n00 = a00 - t * a02, n22 = a22 + t * a02;
      const double n01 = c * a01 - s * a12, n12 = s * a01 + c * a12;
      a00 = n00; a22 = n22; a02 = 0.0; a01 = n01; a12 = n12;
#pragma unroll
      for (int k = 0; k < 3; k++) {
        const double vp = v[k][0], vq = v[k][2];
        v[k][0] = c * vp - s * vq;
        v[k][2] = s * vp + c * vq;
      }
    }
    // (p,q) = (1,2)
    if (a12 != 0.0) {
      const double theta = (a22 - a11) / (2.0 * a12);
      const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
      const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
      const double n11 = a11 - t * a12, n22 = a22 + t * a12;
      const double n01 = c * a01 - s * a02, n02 = s * a01 + c * a02;
      a11 = n11; a22 = n22; a12 = 0.0; a01 = n01; a02 = n02;
#pragma unroll
      for (int k = 0; k < 3; k++) {
        const double vp = v[k][1], vq = v[k][2];
        v[k][1] = c * vp - s * vq;
        v[k][2] = s * vp + c * vq;
      }
    }
  }
  double ev[3] = {a00, a11, a22};
  // sort by |ev| descending, stable
  int i0 = 0, i1 = 1, i2 = 2;
  if (fabs(ev[i1]) > fabs(ev[i0])) { int t = i0; i0 = i1; i1 = t; }
  if (fabs(ev[i2]) > fabs(ev[i1])) { int t = i1; i1 = i2; i2 = t; }
  if (fabs(ev[i1]) > fabs(ev[i0])) { int t = i0; i0 = i1; i1 = t; }
  const int idx[3] = {i0, i1, i2};
#pragma unroll
  for (int j = 0; j < 3; j++) {
    l[j] = ev[idx[j]];
#pragma unroll
    for (int i = 0; i < 3; i++) V[i * 3 + j] = v[i][idx[j]];
  }
}

__device__ __forceinline__ Sym3 sym_inverse(const Sym3& a) {
  // closed-form adjugate / determinant (Eigen Matrix3d::inverse form)
  const double c00 = a.v[3] * a.v[5] - a.v[4] * a.v[4];
  const double c01 = a.v[4] * a.v[2] - a.v[1] * a.v[5];
  const double c02 = a.v[1] * a.v[4] - a.v[3] * a.v[2];
  const double det = a.v[0] * c00 + a.v[1] * c01 + a.v[2] * c02;
  const double id = 1.0 / det;
  Sym3 r;
  r.v[0] = c00 * id;
  r.v[1] = c01 * id;
  r.v[2] = c02 * id;
  r.v[3] = (a.v[0] * a.v[5] - a.v[2] * a.v[2]) * id;
  r.v[4] = (a.v[1] * a.v[2] - a.v[0] * a.v[4]) * id;
  r.v[5] = (a.v[0] * a.v[3] - a.v[1] * a.v[1]) * id;
  return r;
}

// warp-wide sum of a double (butterfly: every lane ends with the total; the
// tree is fixed, so the result is deterministic)
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace apd
