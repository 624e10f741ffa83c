// B200 shim of fast_gicp::LsqRegistration — the optimizer base class of
// FastAPDGICP (reference lsq_registration.hpp:15-85, defaults
// lsq_registration_impl.hpp:11-24). The Gauss-Newton / Levenberg-Marquardt loop
// (lsq_registration_impl.hpp:55-173) lives behind the C-ABI (apd_align); this
// class keeps the tunables and results, with the reference's member names so
// that derived code and callers compile unchanged.
#ifndef FAST_GICP_LSQ_REGISTRATION_HPP
#define FAST_GICP_LSQ_REGISTRATION_HPP

#include <fast_gicp/gicp/apd_shim_common.hpp>

namespace fast_gicp {

enum class LSQ_OPTIMIZER_TYPE { GaussNewton, LevenbergMarquardt };

template <typename PointSource, typename PointTarget>
class LsqRegistration : public pcl::Registration<PointSource, PointTarget, float> {
public:
  APD_SHIM_REGISTRATION_ALIASES(LsqRegistration, PointSource, PointTarget);
  EIGEN_MAKE_ALIGNED_OPERATOR_NEW

  LsqRegistration() {
    this->reg_name_ = "LsqRegistration";
    this->max_iterations_ = 64;
    this->transformation_epsilon_ = 5e-4;
    final_hessian_.setIdentity();
  }
  virtual ~LsqRegistration() {}

  void setRotationEpsilon(double eps) { rotation_epsilon_ = eps; }
  void setInitialLambdaFactor(double init_lambda_factor) { lm_init_lambda_factor_ = init_lambda_factor; }
  void setDebugPrint(bool lm_debug_print) { lm_debug_print_ = lm_debug_print; }
  const Eigen::Matrix<double, 6, 6>& getFinalHessian() const { return final_hessian_; }

  // cost (and optionally H, b) at a pose, without running the optimizer
  double evaluateCost(const Eigen::Matrix4f& relative_pose, Eigen::Matrix<double, 6, 6>* H = nullptr, Eigen::Matrix<double, 6, 1>* b = nullptr) {
    return this->linearize(Eigen::Isometry3f(relative_pose).template cast<double>(), H, b);
  }

  virtual void swapSourceAndTarget() {}
  virtual void clearSource() {}
  virtual void clearTarget() {}

protected:
  using PclBase::converged_;
  using PclBase::final_transformation_;
  using PclBase::input_;
  using PclBase::max_iterations_;
  using PclBase::nr_iterations_;
  using PclBase::transformation_epsilon_;

  virtual double linearize(const Eigen::Isometry3d& trans, Eigen::Matrix<double, 6, 6>* H = nullptr, Eigen::Matrix<double, 6, 1>* b = nullptr) = 0;
  virtual double compute_error(const Eigen::Isometry3d& trans) = 0;

  double rotation_epsilon_ = 2e-3;
  LSQ_OPTIMIZER_TYPE lsq_optimizer_type_ = LSQ_OPTIMIZER_TYPE::LevenbergMarquardt;
  int lm_max_iterations_ = 10;
  double lm_init_lambda_factor_ = 1e-9;
  double lm_lambda_ = -1.0;
  bool lm_debug_print_ = false;
  Eigen::Matrix<double, 6, 6> final_hessian_;
};

}  // namespace fast_gicp

#endif
