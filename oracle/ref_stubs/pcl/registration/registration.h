// TEST INFRASTRUCTURE — stand-in for pcl::Registration<PointSource, PointTarget, Scalar> (PCL 1.10 registration.h /
// registration.hpp) as the reference's LsqRegistration derives from it: the members it names with `using`, the setters the
// factory calls (4DRadarSLAM registrations.cpp:38-51), align() -> computeTransformation(), and pcl::transformPointCloud.
// [ext] align() of PCL also builds its own kd-tree over the target and resizes the output; neither touches the result.
#ifndef APDO_REF_STUB_PCL_REGISTRATION
#define APDO_REF_STUB_PCL_REGISTRATION
#include <Eigen/Geometry>
#include <pcl/point_types.h>
#include <pcl/search/kdtree.h>
namespace pcl {
// pcl::transformPointCloud(cloud_in, cloud_out, Eigen::Matrix4f) [ext: the SSE path of PCL 1.10 sums in another order;
// the transformed cloud is an output of align(), not an input of anything on this path]
template <typename PointT>
void transformPointCloud(const PointCloud<PointT>& in, PointCloud<PointT>& out, const Eigen::Matrix<float, 4, 4>& T) {
  if (&in != &out) out = in;
  for (auto& p : out.points) {
    const float x = p.x, y = p.y, z = p.z;
    p.x = T(0, 0) * x + T(0, 1) * y + T(0, 2) * z + T(0, 3);
    p.y = T(1, 0) * x + T(1, 1) * y + T(1, 2) * z + T(1, 3);
    p.z = T(2, 0) * x + T(2, 1) * y + T(2, 2) * z + T(2, 3);
  }
}
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
  Registration()
      : nr_iterations_(0), max_iterations_(10), final_transformation_(Matrix4::Identity()), transformation_epsilon_(0.0),
        corr_dist_threshold_(std::sqrt(std::numeric_limits<double>::max())), converged_(false) {}
  virtual ~Registration() {}
  virtual void setInputSource(const PointCloudSourceConstPtr& cloud) { input_ = cloud; }
  virtual void setInputTarget(const PointCloudTargetConstPtr& cloud) { target_ = cloud; }
  void setMaximumIterations(int n) { max_iterations_ = n; }
  void setTransformationEpsilon(double e) { transformation_epsilon_ = e; }
  void setMaxCorrespondenceDistance(double d) { corr_dist_threshold_ = d; }
  Matrix4 getFinalTransformation() const { return final_transformation_; }
  bool hasConverged() const { return converged_; }
  void align(PointCloudSource& output) { align(output, Matrix4::Identity()); }
  void align(PointCloudSource& output, const Matrix4& guess) {
    converged_ = false;
    final_transformation_ = Matrix4::Identity();
    computeTransformation(output, guess);
  }
protected:
  virtual void computeTransformation(PointCloudSource& output, const Matrix4& guess) = 0;
  std::string reg_name_;
  int nr_iterations_;
  int max_iterations_;
  PointCloudSourceConstPtr input_;
  PointCloudTargetConstPtr target_;
  Matrix4 final_transformation_;
  double transformation_epsilon_;
  double corr_dist_threshold_;
  bool converged_;
};
}  // namespace pcl
#endif
