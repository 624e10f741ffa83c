// fp64 helpers of the optimizer loop (the serial part of LsqRegistration: 6x6
// solve, SO(3) exponential, pose composition, convergence test — reference
// lsq_registration_impl.hpp:83-173, so3.hpp:59-78). Plain C++, row-major 4x4
// poses. Every function is __host__ __device__: the host LM loop (large or
// sharded clouds) and the device-resident LM kernel (lm.cu) run the SAME code.
#pragma once
#include <cmath>
#include <cstring>

#ifdef __CUDACC__
#define APD_HD __host__ __device__ __forceinline__
#else
#define APD_HD inline
#endif

namespace apd {
namespace hm {

struct Pose {  // row-major 4x4, last row (0,0,0,1)
  double m[16];
  APD_HD static Pose identity() {
    Pose p;
    for (int i = 0; i < 16; i++) p.m[i] = (i % 5 == 0) ? 1.0 : 0.0;
    return p;
  }
  APD_HD double& operator()(int r, int c) { return m[r * 4 + c]; }
  APD_HD double operator()(int r, int c) const { return m[r * 4 + c]; }
};

APD_HD Pose from_colmajor_f32(const float* g) {
  Pose p;
  for (int r = 0; r < 4; r++)
    for (int c = 0; c < 4; c++) p(r, c) = (double)g[c * 4 + r];
  return p;
}
APD_HD Pose from_colmajor_f64(const double* g) {
  Pose p;
  for (int r = 0; r < 4; r++)
    for (int c = 0; c < 4; c++) p(r, c) = g[c * 4 + r];
  return p;
}

APD_HD double dmax(double a, double b) { return a < b ? b : a; }  // std::max(a, b)
APD_HD void dswap(double& a, double& b) { const double t = a; a = b; b = t; }

// delta * x0 for Isometry3d operands (affine part only; last row stays 0,0,0,1)
APD_HD Pose compose(const Pose& a, const Pose& b) {
  Pose c = Pose::identity();
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) {
      double s = 0.0;
      for (int k = 0; k < 3; k++) s += a(i, k) * b(k, j);
      c(i, j) = s;
    }
    double s = 0.0;
    for (int k = 0; k < 3; k++) s += a(i, k) * b(k, 3);
    c(i, 3) = s + a(i, 3);
  }
  return c;
}

// so3_exp (so3.hpp:59-78) followed by Quaterniond::toRotationMatrix, and the
// translation part of d, as step_gn / step_lm assemble delta (:115-117,:140-142)
APD_HD Pose delta_from_twist(const double d[6]) {
  const double theta_sq = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
  double imag, real;
  if (theta_sq < 1e-10) {
    const double theta_quad = theta_sq * theta_sq;
    imag = 0.5 - 1.0 / 48.0 * theta_sq + 1.0 / 3840.0 * theta_quad;
    real = 1.0 - 1.0 / 8.0 * theta_sq + 1.0 / 384.0 * theta_quad;
  } else {
    const double theta = sqrt(theta_sq);
    const double half = 0.5 * theta;
    imag = sin(half) / theta;
    real = cos(half);
  }
  const double w = real, x = imag * d[0], y = imag * d[1], z = imag * d[2];
  const double tx = 2.0 * x, ty = 2.0 * y, tz = 2.0 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w;
  const double txx = tx * x, txy = ty * x, txz = tz * x;
  const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  Pose p = Pose::identity();
  p(0, 0) = 1.0 - (tyy + tzz); p(0, 1) = txy - twz;         p(0, 2) = txz + twy;
  p(1, 0) = txy + twz;         p(1, 1) = 1.0 - (txx + tzz); p(1, 2) = tyz - twx;
  p(2, 0) = txz - twy;         p(2, 1) = tyz + twx;         p(2, 2) = 1.0 - (txx + tyy);
  p(0, 3) = d[3]; p(1, 3) = d[4]; p(2, 3) = d[5];
  return p;
}

// is_converged (lsq_registration_impl.hpp:83-92)
APD_HD bool is_converged(const Pose& delta, double rot_eps, double trans_eps) {
  double rmax = 0.0, tmax = 0.0;
  for (int r = 0; r < 3; r++) {
    for (int c = 0; c < 3; c++) rmax = dmax(rmax, 1.0 / rot_eps * fabs(delta(r, c) - (r == c ? 1.0 : 0.0)));
    tmax = dmax(tmax, 1.0 / trans_eps * fabs(delta(r, 3)));
  }
  return dmax(rmax, tmax) < 1;
}

// Symmetric 6x6 solve by LDL^T with symmetric pivoting on the largest remaining
// |diagonal| (the pivot rule of Eigen::LDLT, which the reference uses at
// lsq_registration_impl.hpp:112-113,137-138). A is row-major full 6x6.
APD_HD bool ldlt_solve6(const double A_in[36], const double rhs[6], double x[6]) {
  double A[6][6], L[6][6], D[6];
  int perm[6];
  for (int i = 0; i < 6; i++) {
    perm[i] = i;
    for (int j = 0; j < 6; j++) {
      A[i][j] = A_in[i * 6 + j];
      L[i][j] = 0.0;
    }
  }
  for (int k = 0; k < 6; k++) {
    int piv = k;
    double best = fabs(A[k][k]);
    for (int i = k + 1; i < 6; i++)
      if (fabs(A[i][i]) > best) { best = fabs(A[i][i]); piv = i; }
    if (piv != k) {
      for (int j = 0; j < 6; j++) dswap(A[k][j], A[piv][j]);
      for (int i = 0; i < 6; i++) dswap(A[i][k], A[i][piv]);
      for (int j = 0; j < k; j++) dswap(L[k][j], L[piv][j]);
      const int tp = perm[k]; perm[k] = perm[piv]; perm[piv] = tp;
    }
    D[k] = A[k][k];
    L[k][k] = 1.0;
    if (D[k] == 0.0) return false;
    for (int i = k + 1; i < 6; i++) L[i][k] = A[i][k] / D[k];
    for (int i = k + 1; i < 6; i++)
      for (int j = k + 1; j < 6; j++) A[i][j] -= L[i][k] * D[k] * L[j][k];
  }
  double y[6], z[6];
  for (int i = 0; i < 6; i++) {
    double s = rhs[perm[i]];
    for (int j = 0; j < i; j++) s -= L[i][j] * y[j];
    y[i] = s;
  }
  for (int i = 0; i < 6; i++) y[i] /= D[i];
  for (int i = 5; i >= 0; i--) {
    double s = y[i];
    for (int j = i + 1; j < 6; j++) s -= L[j][i] * z[j];
    z[i] = s;
  }
  for (int i = 0; i < 6; i++) x[perm[i]] = z[i];
  return true;
}

// 21 upper-triangular values (row-major, r <= c) -> full symmetric row-major 6x6
APD_HD void unpack_upper(const double u[21], double H[36]) {
  int k = 0;
  for (int r = 0; r < 6; r++)
    for (int c = r; c < 6; c++) {
      H[r * 6 + c] = u[k];
      H[c * 6 + r] = u[k];
      k++;
    }
}

}  // namespace hm
}  // namespace apd
