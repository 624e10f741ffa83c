// B200 drop-in of fast_gicp::FastAPDGICP (reference
// fast_apdgicp/include/fast_gicp/gicp/fast_apdgicp.hpp:19-122 and
// impl/fast_apdgicp_impl.hpp). Same class name, template parameters, public and
// protected member functions; header-only; every operation forwards to the CUDA
// library through the C-ABI of include/apdgicp.h (link with -lapdgicp).
//
// 4DRadarSLAM's select_registration_method (registrations.cpp:38-51) and the
// nodelets that call it through pcl::Registration::Ptr compile against this
// header unchanged. Differences a maintainer should know:
//  - setNumThreads() is accepted and ignored (the work runs on the GPU);
//    setDevice() picks the GPU (default 0).
//  - getFitnessScore() is a NON-virtual member of pcl::Registration; through the
//    base pointer it runs PCL's loop over tree_->nearestKSearch. The shim installs
//    its own search method there (apd_search.hpp: the GPU grid behind the
//    pcl::search::KdTree interface, force_no_recompute), so that loop — and
//    getSearchMethodTarget()->nearestKSearch — is answered from one GPU pass and
//    align() builds no CPU kd-tree per target. getFitnessScoreGPU() gives the same
//    number (and the inlier count) in one call.
//  - The protected members source_kdtree_/target_kdtree_ do not exist: the
//    neighbour search is the GPU grid.
//  - setInputSource / setInputTarget pass the cloud's address as the cache key of
//    the C-ABI (the shim holds the shared_ptr, so the address stays valid): a source
//    promoted to target (scan_matching_odometry_nodelet.cpp:587-588) keeps its grid
//    and covariances on the device instead of rebuilding them.
#ifndef FAST_GICP_FAST_IGICP_HPP
#define FAST_GICP_FAST_IGICP_HPP

#include <fast_gicp/gicp/gicp_settings.hpp>
#include <fast_gicp/gicp/lsq_registration.hpp>
#include <fast_gicp/gicp/apd_search.hpp>

namespace fast_gicp {

template <typename PointSource, typename PointTarget>
class FastAPDGICP : public LsqRegistration<PointSource, PointTarget>, private ApdSearchOwner<PointTarget> {
public:
  APD_SHIM_REGISTRATION_ALIASES(FastAPDGICP, PointSource, PointTarget);
  using CovarianceVector = std::vector<Eigen::Matrix4d, Eigen::aligned_allocator<Eigen::Matrix4d>>;

  FastAPDGICP() {
    this->reg_name_ = "FastAPDGICP";
    this->corr_dist_threshold_ = std::numeric_limits<float>::max();
    open(0);
    // the base class's target search method: the GPU grid instead of a CPU kd-tree rebuilt per target (apd_search.hpp)
    search_.reset(new ApdTargetSearch<PointTarget>(this));
    this->setSearchMethodTarget(search_, /*force_no_recompute=*/true);
  }
  virtual ~FastAPDGICP() override {
    if (handle_) apd_destroy(handle_);
  }
  FastAPDGICP(const FastAPDGICP&) = delete;
  FastAPDGICP& operator=(const FastAPDGICP&) = delete;

  // ---- setters of the reference class (fast_apdgicp_impl.hpp:34-65) ----
  void setNumThreads(int n) { num_threads_ = n; }  // no effect: kept for registrations.cpp:41
  void setCorrespondenceRandomness(int k) { k_correspondences_ = k; }
  void setRegularizationMethod(RegularizationMethod method) { regularization_method_ = method; }
  void setAzimuthVar(double var) { azimuth_variance_ = var; }
  void setElevationVar(double var) { elevation_variance_ = var; }
  void setDistVar(double var) { distance_variance_ = var; }
  // B200 additions
  void setDevice(int device) {
    if (device != device_) open(device);
  }
  void setMahalanobisStorageFp64(bool on) { maha_fp64_ = on; }

  // ---- cloud management (:89-135) ----
  virtual void swapSourceAndTarget() override {
    generation_++;
    input_.swap(target_);
    std::swap(source_covs_valid_, target_covs_valid_);
    source_covs_.swap(target_covs_);
    if (handle_) apd_detail::check(handle_, apd_swap_source_and_target(handle_), "swapSourceAndTarget");
  }
  virtual void clearSource() override {
    generation_++;
    input_.reset();
    source_covs_.clear();
    source_covs_valid_ = false;
    if (handle_) apd_clear_source(handle_);
  }
  virtual void clearTarget() override {
    generation_++;
    target_.reset();
    target_covs_.clear();
    target_covs_valid_ = false;
    if (handle_) apd_clear_target(handle_);
  }
  virtual void setInputSource(const PointCloudSourceConstPtr& cloud) override {
    if (input_ == cloud) return;
    generation_++;
    PclBase::setInputSource(cloud);
    source_covs_.clear();
    source_covs_valid_ = false;
    upload(cloud.get(), /*source=*/true);
  }
  virtual void setInputTarget(const PointCloudTargetConstPtr& cloud) override {
    if (target_ == cloud) return;
    generation_++;
    PclBase::setInputTarget(cloud);
    target_covs_.clear();
    target_covs_valid_ = false;
    upload(cloud.get(), /*source=*/false);
  }
  virtual void setSourceCovariances(const CovarianceVector& covs) {
    source_covs_ = covs;
    source_covs_valid_ = true;
    if (handle_ && !covs.empty()) apd_detail::check(handle_, apd_set_source_covariances(handle_, covs[0].data(), (int32_t)covs.size()), "setSourceCovariances");
  }
  virtual void setTargetCovariances(const CovarianceVector& covs) {
    target_covs_ = covs;
    target_covs_valid_ = true;
    if (handle_ && !covs.empty()) apd_detail::check(handle_, apd_set_target_covariances(handle_, covs[0].data(), (int32_t)covs.size()), "setTargetCovariances");
  }
  // The covariances live on the device; they are downloaded on first request.
  const CovarianceVector& getSourceCovariances() const {
    if (!source_covs_valid_ && handle_ && input_ && input_->size() > 0) {
      source_covs_.resize(input_->size());
      if (apd_get_source_covariances(handle_, source_covs_[0].data(), (int32_t)source_covs_.size()) == APD_OK) source_covs_valid_ = true;
    }
    return source_covs_;
  }
  const CovarianceVector& getTargetCovariances() const {
    if (!target_covs_valid_ && handle_ && target_ && target_->size() > 0) {
      target_covs_.resize(target_->size());
      if (apd_get_target_covariances(handle_, target_covs_[0].data(), (int32_t)target_covs_.size()) == APD_OK) target_covs_valid_ = true;
    }
    return target_covs_;
  }

  // getFitnessScore(max_range) of pcl::Registration, computed on the GPU with the
  // grid that align() already built; optionally the inlier count of the status
  // message (scan_matching_odometry_nodelet.cpp:677-689).
  double getFitnessScoreGPU(double max_range = std::numeric_limits<double>::max(), int* n_inliers = nullptr, double inlier_sq_threshold = 0.25) {
    double score = std::numeric_limits<double>::max();
    int32_t in_range = 0, inl = 0;
    if (handle_) apd_detail::check(handle_, apd_fitness(handle_, nullptr, max_range, &score, &in_range, inlier_sq_threshold, &inl), "getFitnessScoreGPU");
    if (n_inliers) *n_inliers = inl;
    return score;
  }
  apd_handle* handle() { return handle_; }
  // the search method installed in the base class (instrumentation: how its queries were answered)
  const ApdTargetSearch<PointTarget>& targetSearch() const { return *search_; }

protected:
  using PclBase::converged_;
  using PclBase::corr_dist_threshold_;
  using PclBase::final_transformation_;
  using PclBase::input_;
  using PclBase::max_iterations_;
  using PclBase::nr_iterations_;
  using PclBase::reg_name_;
  using PclBase::target_;
  using PclBase::transformation_epsilon_;
  using LsqRegistration<PointSource, PointTarget>::final_hessian_;
  using LsqRegistration<PointSource, PointTarget>::lm_debug_print_;
  using LsqRegistration<PointSource, PointTarget>::lm_init_lambda_factor_;
  using LsqRegistration<PointSource, PointTarget>::lm_max_iterations_;
  using LsqRegistration<PointSource, PointTarget>::lsq_optimizer_type_;
  using LsqRegistration<PointSource, PointTarget>::rotation_epsilon_;

  // pcl::Registration::align() -> here (:148-157 + lsq_registration_impl.hpp:55-80)
  virtual void computeTransformation(PointCloudSource& output, const Matrix4& guess) override {
    converged_ = false;
    generation_++;
    if (!handle_ || !input_ || !target_) return;
    push_params();
    Eigen::Matrix4f T = Eigen::Matrix4f::Identity();
    int32_t conv = 0, iters = 0;
    // (pcl::Registration::align() hands over `output` as a copy of the input; a direct caller may not have)
    if (output.size() != input_->size()) output.points = input_->points;
    std::vector<float> xyz(3 * input_->size());
    const int rc = apd_align(handle_, guess.data(), T.data(), nullptr, final_hessian_.data(), &conv, &iters, xyz.empty() ? nullptr : xyz.data());
    apd_detail::check(handle_, rc, "align");
    if (rc != APD_OK) return;
    final_transformation_ = T;
    converged_ = conv != 0;
    nr_iterations_ = iters;
    generation_++;  // (final_transformation_ changed)
    source_covs_valid_ = source_covs_valid_ && !source_covs_.empty();
    // pcl::transformPointCloud(*input_, output, final_transformation_): align() already copied the
    // input's fields into `output`; only x, y, z change, and the GPU has just computed them.
    for (std::size_t i = 0; i < output.size() && 3 * i + 2 < xyz.size(); i++) {
      output[i].x = xyz[3 * i + 0];
      output[i].y = xyz[3 * i + 1];
      output[i].z = xyz[3 * i + 2];
    }
  }

  virtual void update_correspondences(const Eigen::Isometry3d& trans) {
    if (!handle_) return;
    push_params();
    apd_detail::check(handle_, apd_update_correspondences(handle_, trans.matrix().data()), "update_correspondences");
  }
  virtual double linearize(const Eigen::Isometry3d& trans, Eigen::Matrix<double, 6, 6>* H, Eigen::Matrix<double, 6, 1>* b) override {
    double err = 0.0;
    if (!handle_) return err;
    push_params();
    const bool hb = H != nullptr && b != nullptr;
    apd_detail::check(handle_, apd_linearize(handle_, trans.matrix().data(), hb ? H->data() : nullptr, hb ? b->data() : nullptr, &err), "linearize");
    return err;
  }
  virtual double compute_error(const Eigen::Isometry3d& trans) override {
    double err = 0.0;
    if (handle_) apd_detail::check(handle_, apd_compute_error(handle_, trans.matrix().data(), &err), "compute_error");
    return err;
  }

private:
  void open(int device) {
    if (handle_) apd_destroy(handle_);
    handle_ = nullptr;
    device_ = device;
    const int rc = apd_create(device, &handle_);
    if (rc != APD_OK) {
      handle_ = nullptr;
      throw std::runtime_error("FastAPDGICP (B200): no usable sm_100 CUDA device " + std::to_string(device) + " and there is no CPU fallback");
    }
    if (input_) upload(input_.get(), true);
    if (target_) upload(target_.get(), false);
  }
  template <typename CloudT>
  void upload(const CloudT* cloud, bool source) {
    if (!handle_ || !cloud) return;
    using P = typename CloudT::PointType;
    const void* data = cloud->points.empty() ? static_cast<const void*>(cloud) : static_cast<const void*>(cloud->points.data());
    const int32_t n = (int32_t)cloud->points.size();
    // cache key = the cloud's address: input_ / target_ keep the cloud alive while the key is in use, so identity is as
    // good as the reference's shared_ptr comparison (:116,:128); the library also fingerprints the content
    const uint64_t key = (uint64_t)reinterpret_cast<std::uintptr_t>(cloud);
    const int rc = source ? apd_set_source(handle_, data, n, (int32_t)sizeof(P), apd_detail::xyz_offset<P>(), apd_detail::LabelOffset<P>::value(), key)
                          : apd_set_target(handle_, data, n, (int32_t)sizeof(P), apd_detail::xyz_offset<P>(), apd_detail::LabelOffset<P>::value(), key);
    apd_detail::check(handle_, rc, source ? "setInputSource" : "setInputTarget");
  }
  void push_params() {
    apd_params p;
    apd_default_params(&p);
    p.k_correspondences = k_correspondences_;
    p.regularization = static_cast<int32_t>(regularization_method_);
    p.max_correspondence_distance = corr_dist_threshold_;
    p.dist_var = distance_variance_;
    p.azimuth_var = azimuth_variance_;
    p.elevation_var = elevation_variance_;
    p.max_iterations = max_iterations_;
    p.optimizer = lsq_optimizer_type_ == LSQ_OPTIMIZER_TYPE::GaussNewton ? APD_OPT_GAUSS_NEWTON : APD_OPT_LEVENBERG_MARQUARDT;
    p.rotation_epsilon = rotation_epsilon_;
    p.transformation_epsilon = transformation_epsilon_;
    p.lm_max_iterations = lm_max_iterations_;
    p.lm_debug_print = lm_debug_print_ ? 1 : 0;
    p.lm_init_lambda_factor = lm_init_lambda_factor_;
    p.maha_fp64 = maha_fp64_ ? 1 : 0;
    p.variant = variant_;
    p.voxel_resolution = voxel_resolution_;
    p.voxel_search = voxel_search_;
    p.voxel_mode = voxel_mode_;
    apd_set_params(handle_, &p);
  }

protected:
  int num_threads_ = 0;
  int k_correspondences_ = 20;
  RegularizationMethod regularization_method_ = RegularizationMethod::PLANE;
  double azimuth_variance_ = 0.5;
  double elevation_variance_ = 1.0;
  double distance_variance_ = 0.86;
  mutable CovarianceVector source_covs_, target_covs_;
  mutable bool source_covs_valid_ = false, target_covs_valid_ = false;
  int variant_ = APD_VARIANT_APDGICP;  // fast_gicp.hpp's FastGICP sets APD_VARIANT_GICP, fast_vgicp.hpp's FastVGICP APD_VARIANT_VGICP
  double voxel_resolution_ = 1.0;      // FastVGICP only (fast_vgicp_impl.hpp:22-24)
  int voxel_search_ = APD_VOXEL_DIRECT1;
  int voxel_mode_ = APD_VOXEL_ADDITIVE;

private:
  // ApdSearchOwner: what the installed search method asks of this object
  apd_handle* apdSearchHandle() const override { return (input_ && target_) ? handle_ : nullptr; }
  unsigned long long apdSearchGeneration() const override { return generation_; }
  std::size_t apdSearchSourceSize() const override { return input_ ? input_->size() : 0; }
  const PointTarget* apdSearchTargetPoints(std::size_t* n) const override {
    *n = target_ ? target_->size() : 0;
    return *n ? target_->points.data() : nullptr;
  }

  apd_handle* handle_ = nullptr;
  int device_ = 0;
  bool maha_fp64_ = false;
  unsigned long long generation_ = 0;  // bumps whenever a cloud or final_transformation_ changes (the search batch's validity)
  typename ApdTargetSearch<PointTarget>::Ptr search_;
};

}  // namespace fast_gicp

#endif
