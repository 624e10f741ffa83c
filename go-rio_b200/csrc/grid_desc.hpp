// The uniform grid of a cloud (plain C++: shared by the CUDA headers and the host-only code), and its sizing — ONE
// function for the host (clouds staged by the library: the bounding box comes out of the staging pass) and for the
// device (the fused registration kernel sizes the grids of the clouds it was handed itself). Both must choose the SAME
// grid for the same box: the cell-sorted layout, and with it the order of every sum, follows from it. So the sizing
// uses only operations that round identically on both sides — +, -, *, / (spelled with the _rn intrinsics on the
// device so that nothing is contracted into an FMA), floor, integer arithmetic — and no libm call.
#pragma once

#if defined(__CUDACC__)
#define APD_GRID_HD __host__ __device__
#else
#define APD_GRID_HD
#endif

namespace apd {

struct GridDesc {
  float ox, oy, oz;   // origin (min corner)
  float inv_cell;     // 1 / cell size
  float cell;         // cell size (metres)
  int nx, ny, nz;     // dimensions
};

namespace gridmath {
#if defined(__CUDA_ARCH__)
APD_GRID_HD inline double mul(double a, double b) { return __dmul_rn(a, b); }
APD_GRID_HD inline double add(double a, double b) { return __dadd_rn(a, b); }
APD_GRID_HD inline double sub(double a, double b) { return __dsub_rn(a, b); }
APD_GRID_HD inline double div(double a, double b) { return __ddiv_rn(a, b); }
APD_GRID_HD inline float fmulf(float a, float b) { return __fmul_rn(a, b); }
APD_GRID_HD inline float fsubf(float a, float b) { return __fsub_rn(a, b); }
APD_GRID_HD inline double dfloor(double a) { return floor(a); }
APD_GRID_HD inline float ffloor(float a) { return floorf(a); }
#else
inline double mul(double a, double b) { volatile double r = a * b; return r; }  // (volatile: one rounding, no contraction)
inline double add(double a, double b) { volatile double r = a + b; return r; }
inline double sub(double a, double b) { volatile double r = a - b; return r; }
inline double div(double a, double b) { volatile double r = a / b; return r; }
inline float fmulf(float a, float b) { volatile float r = a * b; return r; }
inline float fsubf(float a, float b) { volatile float r = a - b; return r; }
inline double dfloor(double a) { return __builtin_floor(a); }
inline float ffloor(float a) { return __builtin_floorf(a); }
#endif
APD_GRID_HD inline double dmax(double a, double b) { return a < b ? b : a; }
APD_GRID_HD inline bool finite_pos(double a) { return a > 0.0 && a < 1.7e308; }

// cube root of v > 0 by Newton steps from a power of two (exactly rounded operations only: the same bits on host and device)
APD_GRID_HD inline double cbrt_det(double v) {
  if (!(v > 0.0)) return 0.0;
  double x = 1.0;
  while (mul(mul(x, x), x) < v) x = mul(x, 2.0);
  while (mul(mul(x, x), x) > v) x = mul(x, 0.5);  // x^3 <= v < 8 x^3
  for (int it = 0; it < 12; it++) x = div(add(mul(2.0, x), div(v, mul(x, x))), 3.0);
  return x;
}
}  // namespace gridmath

// Grid sizing: cell edge so that the bounding box holds ~cells_per_point (4) * n
// cells (radar clouds live on surfaces, so occupied cells hold several points),
// at most 2048 cells per axis (bounds the fp32 cell-coordinate error the search
// margin covers) and 2^28 cells in total. At most 2 * max(64, cells_per_point * n) cells come out
// (grid_cell_capacity): a caller that cannot wait for the box sizes its arrays for that.
APD_GRID_HD inline void size_grid(const float bbox[6], int n, double cells_per_point, GridDesc& g, int& ncells) {
  using namespace gridmath;
  double ext[3];
  for (int a = 0; a < 3; a++) {
    ext[a] = sub((double)bbox[3 + a], (double)bbox[a]);
    if (!finite_pos(ext[a])) ext[a] = 0.0;
  }
  const double emax = dmax(dmax(ext[0], ext[1]), dmax(ext[2], 1e-3));
  const double target = dmax(64.0, mul(cells_per_point, (double)n));
  double vol = 1.0;
  for (int a = 0; a < 3; a++) vol = mul(vol, dmax(ext[a], mul(emax, 1e-3)));
  double cell = cbrt_det(div(vol, target));
  cell = dmax(cell, div(emax, 2040.0));
  for (int iter = 0; iter < 200; iter++) {
    long long tot = 1;
    bool too_wide = false;
    for (int a = 0; a < 3; a++) {
      const long long d = (long long)dfloor(div(ext[a], cell)) + 1;
      if (d > 2048) too_wide = true;
      tot *= d;
    }
    if (!too_wide && (double)tot <= mul(2.0, target) && tot <= (1ll << 28)) break;
    cell = mul(cell, 1.2599210498948732);  // 2^(1/3)
  }
  g.ox = bbox[0];
  g.oy = bbox[1];
  g.oz = bbox[2];
  g.inv_cell = (float)div(1.0, cell);
  g.cell = (float)div(1.0, (double)g.inv_cell);
  // dims from the SAME fp32 expression the kernels use, so the max corner maps inside
  const int dx = (int)ffloor(fmulf(fsubf(bbox[3], g.ox), g.inv_cell)) + 1;
  const int dy = (int)ffloor(fmulf(fsubf(bbox[4], g.oy), g.inv_cell)) + 1;
  const int dz = (int)ffloor(fmulf(fsubf(bbox[5], g.oz), g.inv_cell)) + 1;
  g.nx = dx > 1 ? dx : 1;
  g.ny = dy > 1 ? dy : 1;
  g.nz = dz > 1 ? dz : 1;
  ncells = g.nx * g.ny * g.nz;
}

// upper bound of the cells size_grid can choose for n points (the dims above may exceed the loop's estimate by one
// per axis: a few per cent of slack on top of the factor two)
inline long long grid_cell_capacity(int n, double cells_per_point) {
  const double target = cells_per_point * (double)n > 64.0 ? cells_per_point * (double)n : 64.0;
  return (long long)(2.6 * target) + 4096;
}

}  // namespace apd
