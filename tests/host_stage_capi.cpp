// TEST INFRASTRUCTURE: exposes the product's host-side cloud preparation (go-rio_b200/csrc/host_stage.hpp) to
// tests/test_host_stage.py. Compiled with g++ by the test.
#include "host_stage.hpp"

extern "C" {
void hs_stage_cloud(const void* pts, int n, int stride, int xyz_off, int label_off, float* dst16, float* bbox6) {
  apd::stage_cloud(pts, n, stride, xyz_off, label_off, dst16, bbox6);
}
void hs_bounds_of_packed(const void* pts, int n, float* bbox6) { apd::bounds_of_packed(pts, n, bbox6); }
// out8: ox, oy, oz, inv_cell, cell, nx, ny, nz (as doubles); returns ncells
int hs_size_grid(const float* bbox6, int n, double cells_per_point, double* out8) {
  apd::GridDesc g;
  int ncells = 0;
  apd::size_grid(bbox6, n, cells_per_point, g, ncells);
  out8[0] = g.ox; out8[1] = g.oy; out8[2] = g.oz; out8[3] = g.inv_cell; out8[4] = g.cell; out8[5] = g.nx; out8[6] = g.ny; out8[7] = g.nz;
  return ncells;
}
}
