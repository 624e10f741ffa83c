// MINIMAL STAND-IN for <pcl/registration/registration.h>: the members and the
// align() protocol of pcl::Registration (PCL 1.10) that the FastAPDGICP shim and
// its callers (registrations.cpp:38-51, scan_matching_odometry_nodelet.cpp:430-490,
// loop_detector.cpp:222-236) touch.
#ifndef APD_STUB_PCL_REGISTRATION
#define APD_STUB_PCL_REGISTRATION
#include <cfloat>
#include <cmath>
#include <limits>
#include <string>
#include <Eigen/Core>
#include <pcl/point_cloud.h>
namespace pcl {
template <typename PointSource, typename PointTarget, typename Scalar = float>
class Registration {
public:
  using Matrix4 = Eigen::Matrix<Scalar, 4, 4>;
  using PointCloudSource = pcl::PointCloud<PointSource>;
  using PointCloudSourcePtr = typename PointCloudSource::Ptr;
  using PointCloudSourceConstPtr = typename PointCloudSource::ConstPtr;
  using PointCloudTarget = pcl::PointCloud<PointTarget>;
  using PointCloudTargetPtr = typename PointCloudTarget::Ptr;
  using PointCloudTargetConstPtr = typename PointCloudTarget::ConstPtr;
  using Ptr = std::shared_ptr<Registration<PointSource, PointTarget, Scalar>>;
  Registration() { final_transformation_.setIdentity(); }
  virtual ~Registration() {}
  virtual void setInputSource(const PointCloudSourceConstPtr& cloud) { input_ = cloud; }
  virtual void setInputTarget(const PointCloudTargetConstPtr& cloud) { target_ = cloud; }
  void setMaximumIterations(int n) { max_iterations_ = n; }
  void setTransformationEpsilon(double e) { transformation_epsilon_ = e; }
  void setMaxCorrespondenceDistance(double d) { corr_dist_threshold_ = d; }
  bool hasConverged() const { return converged_; }
  Matrix4 getFinalTransformation() const { return final_transformation_; }
  void align(PointCloudSource& output) { align(output, Matrix4::Identity()); }
  void align(PointCloudSource& output, const Matrix4& guess) {
    if (!input_ || !target_) return;
    converged_ = false;
    final_transformation_.setIdentity();
    output.points = input_->points;  // PCL copies the input's fields, then the subclass transforms xyz
    computeTransformation(output, guess);
  }
  // brute-force version of PCL's getFitnessScore, only for the stub
  double getFitnessScore(double max_range = std::numeric_limits<double>::max()) {
    double sum = 0; int nr = 0;
    const Matrix4& T = final_transformation_;
    for (const auto& p : input_->points) {
      const float x = ((T(0,0)*p.x + T(0,1)*p.y) + T(0,2)*p.z) + T(0,3);
      const float y = ((T(1,0)*p.x + T(1,1)*p.y) + T(1,2)*p.z) + T(1,3);
      const float z = ((T(2,0)*p.x + T(2,1)*p.y) + T(2,2)*p.z) + T(2,3);
      float best = FLT_MAX;
      for (const auto& q : target_->points) {
        const float dx = x - q.x, dy = y - q.y, dz = z - q.z;
        const float d = (dx*dx + dy*dy) + dz*dz;
        if (d < best) best = d;
      }
      if ((double)best <= max_range) { sum += best; nr++; }
    }
    return nr ? sum / nr : std::numeric_limits<double>::max();
  }
protected:
  virtual void computeTransformation(PointCloudSource& output, const Matrix4& guess) = 0;
  std::string reg_name_;
  PointCloudSourceConstPtr input_;
  PointCloudTargetConstPtr target_;
  int nr_iterations_ = 0;
  int max_iterations_ = 10;
  Matrix4 final_transformation_;
  double transformation_epsilon_ = 0.0;
  double corr_dist_threshold_ = std::sqrt(std::numeric_limits<double>::max());
  bool converged_ = false;
};
}  // namespace pcl
#endif
