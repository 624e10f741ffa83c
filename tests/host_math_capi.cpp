// TEST INFRASTRUCTURE: exposes the product's serial optimizer math (go-rio_b200/csrc/host_math.hpp — the code the host LM
// loop AND the device-resident loop kernel both run) to tests/test_host_math.py. Compiled with g++ by the test.
#include "host_math.hpp"

using namespace apd::hm;

extern "C" {
int hm_ldlt_solve6(const double* A36, const double* rhs6, double* x6) { return ldlt_solve6(A36, rhs6, x6) ? 1 : 0; }
void hm_delta_from_twist(const double* d6, double* pose16_rowmajor) {
  const Pose p = delta_from_twist(d6);
  for (int i = 0; i < 16; i++) pose16_rowmajor[i] = p.m[i];
}
void hm_compose(const double* a16, const double* b16, double* c16) {
  Pose a, b;
  for (int i = 0; i < 16; i++) { a.m[i] = a16[i]; b.m[i] = b16[i]; }
  const Pose c = compose(a, b);
  for (int i = 0; i < 16; i++) c16[i] = c.m[i];
}
int hm_is_converged(const double* delta16, double rot_eps, double trans_eps) {
  Pose d;
  for (int i = 0; i < 16; i++) d.m[i] = delta16[i];
  return is_converged(d, rot_eps, trans_eps) ? 1 : 0;
}
void hm_unpack_upper(const double* u21, double* H36) { unpack_upper(u21, H36); }
void hm_from_colmajor_f32(const float* g16, double* pose16_rowmajor) {
  const Pose p = from_colmajor_f32(g16);
  for (int i = 0; i < 16; i++) pose16_rowmajor[i] = p.m[i];
}
}
