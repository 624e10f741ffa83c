// TEST INFRASTRUCTURE — part of the CPU oracle (see oracle/README.md).
// Small fixed-size fp64 linear algebra used by the restatement. The reference
// gets these from Eigen 3.3.7 (absent here: SURVEY.md §8c); each routine names
// the Eigen call it stands in for and the documented deviation.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>

namespace apdo {

// Row-major 3x3 / 4x4 / 6x6 helpers. (The C ABI converts to/from Eigen's
// column-major layout at the boundary.)
struct M3 {
  double m[9];
  double& operator()(int r, int c) { return m[r * 3 + c]; }
  double operator()(int r, int c) const { return m[r * 3 + c]; }
};
struct M4 {
  double m[16];
  double& operator()(int r, int c) { return m[r * 4 + c]; }
  double operator()(int r, int c) const { return m[r * 4 + c]; }
  static M4 identity() {
    M4 a;
    std::memset(a.m, 0, sizeof(a.m));
    a(0, 0) = a(1, 1) = a(2, 2) = a(3, 3) = 1.0;
    return a;
  }
};

inline M3 zero3() {
  M3 a;
  std::memset(a.m, 0, sizeof(a.m));
  return a;
}
inline M3 mul3(const M3& a, const M3& b) {
  M3 c;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double s = 0.0;
      for (int k = 0; k < 3; k++) s += a(i, k) * b(k, j);
      c(i, j) = s;
    }
  return c;
}
inline M3 transpose3(const M3& a) {
  M3 c;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) c(i, j) = a(j, i);
  return c;
}
inline M3 add3(const M3& a, const M3& b) {
  M3 c;
  for (int i = 0; i < 9; i++) c.m[i] = a.m[i] + b.m[i];
  return c;
}
// Stands in for Eigen::Matrix3d::inverse() (closed-form cofactors / determinant,
// Eigen/src/LU/InverseImpl.h compute_inverse_size3). Same formula; rounding may
// differ in the last ulp.
inline M3 inverse3(const M3& a) {
  M3 c;
  const double c00 = a(1, 1) * a(2, 2) - a(1, 2) * a(2, 1);
  const double c01 = a(1, 2) * a(2, 0) - a(1, 0) * a(2, 2);
  const double c02 = a(1, 0) * a(2, 1) - a(1, 1) * a(2, 0);
  const double det = a(0, 0) * c00 + a(0, 1) * c01 + a(0, 2) * c02;
  const double id = 1.0 / det;
  c(0, 0) = c00 * id;
  c(1, 0) = c01 * id;
  c(2, 0) = c02 * id;
  c(0, 1) = (a(0, 2) * a(2, 1) - a(0, 1) * a(2, 2)) * id;
  c(1, 1) = (a(0, 0) * a(2, 2) - a(0, 2) * a(2, 0)) * id;
  c(2, 1) = (a(0, 1) * a(2, 0) - a(0, 0) * a(2, 1)) * id;
  c(0, 2) = (a(0, 1) * a(1, 2) - a(0, 2) * a(1, 1)) * id;
  c(1, 2) = (a(0, 2) * a(1, 0) - a(0, 0) * a(1, 2)) * id;
  c(2, 2) = (a(0, 0) * a(1, 1) - a(0, 1) * a(1, 0)) * id;
  return c;
}
inline double frobenius3(const M3& a) {
  double s = 0.0;
  for (int i = 0; i < 9; i++) s += a.m[i] * a.m[i];
  return std::sqrt(s);
}

// Symmetric 3x3 eigen-decomposition by cyclic Jacobi rotations, fp64.
// Stands in for Eigen::JacobiSVD<Matrix3d>(A, ComputeFullU|ComputeFullV) on the
// symmetric matrices this path feeds it (fast_apdgicp_impl.hpp:266,330,385):
// for a symmetric positive semi-definite A the SVD is A = V diag(l) V^T with
// singular values = eigenvalues sorted descending and U = V.
// `A` is symmetrised as (A + A^T)/2 first. Output: l[0] >= l[1] >= l[2] are the
// SINGULAR values (|eigenvalue|), V columns the right singular vectors, and
// U = V with the column sign flipped where the eigenvalue is negative.
inline void svd3_sym(const M3& Ain, double l[3], M3& U, M3& V) {
  double a[3][3];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) a[i][j] = 0.5 * (Ain(i, j) + Ain(j, i));
  double v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int sweep = 0; sweep < 64; sweep++) {
    const double off = std::fabs(a[0][1]) + std::fabs(a[0][2]) + std::fabs(a[1][2]);
    const double diag = std::fabs(a[0][0]) + std::fabs(a[1][1]) + std::fabs(a[2][2]);
    if (off <= 1e-300 || off <= 1e-22 * diag) break;
    for (int p = 0; p < 2; p++)
      for (int q = p + 1; q < 3; q++) {
        const double apq = a[p][q];
        if (apq == 0.0) continue;
        const double theta = (a[q][q] - a[p][p]) / (2.0 * apq);
        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0);
        const double s = t * c;
        // A <- J^T A J
        for (int k = 0; k < 3; k++) {
          const double akp = a[k][p], akq = a[k][q];
          a[k][p] = c * akp - s * akq;
          a[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < 3; k++) {
          const double apk = a[p][k], aqk = a[q][k];
          a[p][k] = c * apk - s * aqk;
          a[q][k] = s * apk + c * aqk;
        }
        a[p][q] = a[q][p] = 0.0;
        for (int k = 0; k < 3; k++) {
          const double vkp = v[k][p], vkq = v[k][q];
          v[k][p] = c * vkp - s * vkq;
          v[k][q] = s * vkp + c * vkq;
        }
      }
  }
  int idx[3] = {0, 1, 2};
  double ev[3] = {a[0][0], a[1][1], a[2][2]};
  // sort by |eigenvalue| descending (stable on ties: lower index first)
  std::stable_sort(idx, idx + 3, [&](int x, int y) { return std::fabs(ev[x]) > std::fabs(ev[y]); });
  for (int j = 0; j < 3; j++) {
    const int s = idx[j];
    l[j] = std::fabs(ev[s]);
    const double sg = ev[s] < 0.0 ? -1.0 : 1.0;
    for (int i = 0; i < 3; i++) {
      V(i, j) = v[i][s];
      U(i, j) = sg * v[i][s];
    }
  }
}

// U * diag(d) * V^T
inline M3 udvt(const M3& U, const double d[3], const M3& V) {
  M3 c;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double s = 0.0;
      for (int k = 0; k < 3; k++) s += U(i, k) * d[k] * V(j, k);
      c(i, j) = s;
    }
  return c;
}

inline M4 mul4(const M4& a, const M4& b) {
  M4 c;
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) {
      double s = 0.0;
      for (int k = 0; k < 4; k++) s += a(i, k) * b(k, j);
      c(i, j) = s;
    }
  return c;
}
// Isometry3d * Isometry3d (Eigen keeps the last row (0,0,0,1) for Isometry mode)
inline M4 mul_isometry(const M4& a, const M4& b) {
  M4 c = M4::identity();
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

// 6x6 symmetric solve. Stands in for Eigen::LDLT<Matrix<double,6,6>>::solve
// (lsq_registration_impl.hpp:112-113,137-138): LDL^T with symmetric pivoting on
// the largest remaining |diagonal| (Eigen's pivot rule), then forward/diagonal/
// back substitution. Returns false if a pivot is exactly zero.
inline bool ldlt_solve6(const double Hin[36], const double rhs[6], double x[6]) {
  const int n = 6;
  double A[6][6];
  int perm[6];
  for (int i = 0; i < n; i++) {
    perm[i] = i;
    for (int j = 0; j < n; j++) A[i][j] = Hin[i * 6 + j];
  }
  double L[6][6] = {{0}};
  double D[6];
  for (int k = 0; k < n; k++) {
    int piv = k;
    double best = std::fabs(A[k][k]);
    for (int i = k + 1; i < n; i++)
      if (std::fabs(A[i][i]) > best) {
        best = std::fabs(A[i][i]);
        piv = i;
      }
    if (piv != k) {
      for (int j = 0; j < n; j++) std::swap(A[k][j], A[piv][j]);
      for (int i = 0; i < n; i++) std::swap(A[i][k], A[i][piv]);
      for (int j = 0; j < k; j++) std::swap(L[k][j], L[piv][j]);
      std::swap(perm[k], perm[piv]);
    }
    D[k] = A[k][k];
    L[k][k] = 1.0;
    if (D[k] == 0.0) return false;
    for (int i = k + 1; i < n; i++) L[i][k] = A[i][k] / D[k];
    for (int i = k + 1; i < n; i++)
      for (int j = k + 1; j < n; j++) A[i][j] -= L[i][k] * D[k] * L[j][k];
  }
  double y[6];
  for (int i = 0; i < n; i++) {
    double s = rhs[perm[i]];
    for (int j = 0; j < i; j++) s -= L[i][j] * y[j];
    y[i] = s;
  }
  for (int i = 0; i < n; i++) y[i] /= D[i];
  double z[6];
  for (int i = n - 1; i >= 0; i--) {
    double s = y[i];
    for (int j = i + 1; j < n; j++) s -= L[j][i] * z[j];
    z[i] = s;
  }
  for (int i = 0; i < n; i++) x[perm[i]] = z[i];
  return true;
}

// Eigen::Quaterniond::toRotationMatrix() (Eigen/src/Geometry/Quaternion.h),
// same expression order.
inline M3 quat_to_rot(double w, double x, double y, double z) {
  M3 r;
  const double tx = 2.0 * x, ty = 2.0 * y, tz = 2.0 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w;
  const double txx = tx * x, txy = ty * x, txz = tz * x;
  const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  r(0, 0) = 1.0 - (tyy + tzz);
  r(0, 1) = txy - twz;
  r(0, 2) = txz + twy;
  r(1, 0) = txy + twz;
  r(1, 1) = 1.0 - (txx + tzz);
  r(1, 2) = tyz - twx;
  r(2, 0) = txz - twy;
  r(2, 1) = tyz + twx;
  r(2, 2) = 1.0 - (txx + tyy);
  return r;
}

// so3_exp -> quaternion (reference so3.hpp:59-78), then toRotationMatrix as the
// callers do (lsq_registration_impl.hpp:116,141).
inline M3 so3_exp_rot(const double omega[3]) {
  const double theta_sq = omega[0] * omega[0] + omega[1] * omega[1] + omega[2] * omega[2];
  double imag_factor, real_factor;
  if (theta_sq < 1e-10) {
    const double theta_quad = theta_sq * theta_sq;
    imag_factor = 0.5 - 1.0 / 48.0 * theta_sq + 1.0 / 3840.0 * theta_quad;
    real_factor = 1.0 - 1.0 / 8.0 * theta_sq + 1.0 / 384.0 * theta_quad;
  } else {
    const double theta = std::sqrt(theta_sq);
    const double half_theta = 0.5 * theta;
    imag_factor = std::sin(half_theta) / theta;
    real_factor = std::cos(half_theta);
  }
  return quat_to_rot(real_factor, imag_factor * omega[0], imag_factor * omega[1], imag_factor * omega[2]);
}

}  // namespace apdo
