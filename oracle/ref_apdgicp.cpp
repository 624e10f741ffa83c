// TEST INFRASTRUCTURE — runs the REFERENCE's own registration code. The headers below are compiled where they lie under
// /root/reference/fast_apdgicp/include (never copied) against the Eigen / PCL / boost stand-ins of oracle/ref_stubs:
//   fast_gicp/gicp/fast_apdgicp.hpp + impl/fast_apdgicp_impl.hpp     calculate_covariances, update_correspondences,
//                                                                    linearize, compute_error, computeTransformation
//   fast_gicp/gicp/lsq_registration.hpp + impl/lsq_registration_impl.hpp   the LM / GN loop, is_converged
//   fast_gicp/so3/so3.hpp, fast_gicp/gicp/gicp_settings.hpp
// What the reference wrote runs as written; what Eigen / PCL / FLANN would compute underneath is stood in ([ext], see the
// stubs' headers: symmetric Jacobi SVD, pivoted LDL^T, Gauss-Jordan inverse, brute-force neighbour search with (d2, index)
// ties). Output: tests/golden/apdgicp_reference.npz (tests/golden/make_apdgicp_reference.py), which pins the oracle
// (oracle/apd_oracle.cpp) against the reference's own code.
#include <fast_gicp/gicp/fast_apdgicp.hpp>
#include <fast_gicp/gicp/impl/lsq_registration_impl.hpp>
#include <fast_gicp/gicp/impl/fast_apdgicp_impl.hpp>

namespace {
using PointT = pcl::PointXYZINormal;
using Cloud = pcl::PointCloud<PointT>;

struct Probe : public fast_gicp::FastAPDGICP<PointT, PointT> {
  using Base = fast_gicp::FastAPDGICP<PointT, PointT>;
  Cloud::Ptr src{new Cloud()}, tgt{new Cloud()};
  void set_optimizer(int gn) { this->lsq_optimizer_type_ = gn ? fast_gicp::LSQ_OPTIMIZER_TYPE::GaussNewton : fast_gicp::LSQ_OPTIMIZER_TYPE::LevenbergMarquardt; }
  void set_lm(int max_it, double init) {
    this->lm_max_iterations_ = max_it;
    this->lm_init_lambda_factor_ = init;
  }
  void ensure_covs() {  // the first lines of computeTransformation (:149-154)
    if (this->source_covs_.size() != this->input_->size()) this->template calculate_covariances<PointT>(this->input_, *this->source_kdtree_, this->source_covs_);
    if (this->target_covs_.size() != this->target_->size()) this->template calculate_covariances<PointT>(this->target_, *this->target_kdtree_, this->target_covs_);
  }
  double do_linearize(const Eigen::Isometry3d& T, Eigen::Matrix<double, 6, 6>* H, Eigen::Matrix<double, 6, 1>* b) { return this->linearize(T, H, b); }
  double do_compute_error(const Eigen::Isometry3d& T) { return this->compute_error(T); }
  const std::vector<int>& corr() const { return this->correspondences_; }
  const std::vector<float>& sqd() const { return this->sq_distances_; }
  const std::vector<Eigen::Matrix4d>& maha() const { return this->mahalanobis_; }
  int iterations() const { return this->nr_iterations_; }
};

Eigen::Isometry3d pose_of(const double* T16_colmajor) {
  Eigen::Matrix4d m;
  for (int r = 0; r < 4; r++)
    for (int c = 0; c < 4; c++) m(r, c) = T16_colmajor[c * 4 + r];
  return Eigen::Isometry3d(m);
}
void fill(Cloud& c, const float* xyzl, int n) {
  c.points.resize((std::size_t)n);
  for (int i = 0; i < n; i++) {
    PointT& p = c.points[(std::size_t)i];
    p.x = xyzl[4 * i];
    p.y = xyzl[4 * i + 1];
    p.z = xyzl[4 * i + 2];
    p.normal_x = xyzl[4 * i + 3];
    p.intensity = 1.f;
  }
}
}  // namespace

extern "C" {
void* aref_create() { return new Probe(); }
void aref_destroy(void* h) { delete static_cast<Probe*>(h); }
void aref_set_params(void* h, int k, int regularization, double max_corr_dist, double dist_var, double az_var, double el_var, int max_iterations,
                     int gauss_newton, double rot_eps, double trans_eps, int lm_max_iterations, double lm_init_lambda_factor, int threads) {
  Probe* p = static_cast<Probe*>(h);
  p->setNumThreads(threads);
  p->setCorrespondenceRandomness(k);
  p->setRegularizationMethod(static_cast<fast_gicp::RegularizationMethod>(regularization));
  p->setMaxCorrespondenceDistance(max_corr_dist);
  p->setDistVar(dist_var);
  p->setAzimuthVar(az_var);
  p->setElevationVar(el_var);
  p->setMaximumIterations(max_iterations);
  p->set_optimizer(gauss_newton);
  p->setRotationEpsilon(rot_eps);
  p->setTransformationEpsilon(trans_eps);
  p->set_lm(lm_max_iterations, lm_init_lambda_factor);
}
void aref_set_source(void* h, const float* xyzl, int n) {
  Probe* p = static_cast<Probe*>(h);
  p->src.reset(new Cloud());
  fill(*p->src, xyzl, n);
  p->setInputSource(p->src);
}
void aref_set_target(void* h, const float* xyzl, int n) {
  Probe* p = static_cast<Probe*>(h);
  p->tgt.reset(new Cloud());
  fill(*p->tgt, xyzl, n);
  p->setInputTarget(p->tgt);
}
void aref_swap(void* h) {
  Probe* p = static_cast<Probe*>(h);
  p->swapSourceAndTarget();
  std::swap(p->src, p->tgt);
}
// which: 0 source, 1 target; out: n x 9 (the 3x3 block, row-major)
void aref_covariances(void* h, int which, double* out9) {
  Probe* p = static_cast<Probe*>(h);
  p->ensure_covs();
  const auto& covs = which == 0 ? p->getSourceCovariances() : p->getTargetCovariances();
  for (std::size_t i = 0; i < covs.size(); i++)
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) out9[9 * i + 3 * r + c] = covs[i](r, c);
}
double aref_linearize(void* h, const double* T16, double* H36, double* b6) {
  Probe* p = static_cast<Probe*>(h);
  p->ensure_covs();
  Eigen::Matrix<double, 6, 6> H;
  Eigen::Matrix<double, 6, 1> b;
  const double e = p->do_linearize(pose_of(T16), H36 ? &H : nullptr, H36 ? &b : nullptr);
  if (H36) {
    for (int i = 0; i < 36; i++) H36[i] = H.a[i];
    for (int i = 0; i < 6; i++) b6[i] = b.a[i];
  }
  return e;
}
double aref_compute_error(void* h, const double* T16) { return static_cast<Probe*>(h)->do_compute_error(pose_of(T16)); }
void aref_get_correspondences(void* h, int* idx, float* sqd, double* maha9) {
  Probe* p = static_cast<Probe*>(h);
  for (std::size_t i = 0; i < p->corr().size(); i++) {
    idx[i] = p->corr()[i];
    sqd[i] = p->sqd()[i];
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) maha9[9 * i + 3 * r + c] = idx[i] >= 0 ? p->maha()[i](r, c) : 0.0;
  }
}
// the timed protocol of bench.py (fast_apdgicp/src/align.cpp:57-83): clearTarget; clearSource; setInputTarget;
// setInputSource; align — reps times, milliseconds of each into ms_out
int aref_bench_align(void* h, const float* src_xyzl, int ns, const float* tgt_xyzl, int nt, int reps, double* ms_out, float* T_out, int* converged,
                     int* iterations) {
  Probe* p = static_cast<Probe*>(h);
  for (int r = 0; r < reps; r++) {
    const double t0 = omp_get_wtime();
    p->clearTarget();
    p->clearSource();
    p->tgt.reset(new Cloud());
    fill(*p->tgt, tgt_xyzl, nt);
    p->setInputTarget(p->tgt);
    p->src.reset(new Cloud());
    fill(*p->src, src_xyzl, ns);
    p->setInputSource(p->src);
    Cloud out;
    p->align(out, Eigen::Matrix4f::Identity());
    ms_out[r] = 1e3 * (omp_get_wtime() - t0);
  }
  const Eigen::Matrix4f T = p->getFinalTransformation();
  if (T_out)
    for (int r = 0; r < 4; r++)
      for (int c = 0; c < 4; c++) T_out[c * 4 + r] = T(r, c);
  if (converged) *converged = p->hasConverged() ? 1 : 0;
  if (iterations) *iterations = p->iterations();
  return 0;
}
// guess: float[16] column-major or NULL; T_out: float[16] column-major
void aref_align(void* h, const float* guess, float* T_out, int* converged, int* iterations, double* H36) {
  Probe* p = static_cast<Probe*>(h);
  Eigen::Matrix4f g = Eigen::Matrix4f::Identity();
  if (guess)
    for (int r = 0; r < 4; r++)
      for (int c = 0; c < 4; c++) g(r, c) = guess[c * 4 + r];
  Cloud out;
  p->align(out, g);
  const Eigen::Matrix4f T = p->getFinalTransformation();
  for (int r = 0; r < 4; r++)
    for (int c = 0; c < 4; c++) T_out[c * 4 + r] = T(r, c);
  *converged = p->hasConverged() ? 1 : 0;
  *iterations = p->iterations();
  const auto& Hf = p->getFinalHessian();
  for (int i = 0; i < 36; i++) H36[i] = Hf.a[i];
}
}
