// MINIMAL STAND-IN for <pcl/registration/registration.h>: the members and the
// align() / initCompute() / getFitnessScore() protocol of pcl::Registration (PCL 1.10) that the FastAPDGICP shim and
// its callers (registrations.cpp:38-51, scan_matching_odometry_nodelet.cpp:430-490 and :674-691,
// loop_detector.cpp:222-236) touch — including the search method tree_: the base class (re)builds it over every new
// target inside align() unless a tree was installed with force_no_recompute, and getFitnessScore() /
// getSearchMethodTarget()->nearestKSearch query it one point at a time.
#ifndef APD_STUB_PCL_REGISTRATION
#define APD_STUB_PCL_REGISTRATION
#include <cfloat>
#include <cmath>
#include <limits>
#include <string>
#include <Eigen/Core>
#include <pcl/point_cloud.h>
#include <pcl/search/kdtree.h>
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
  using KdTree = pcl::search::KdTree<PointTarget>;
  using KdTreePtr = typename KdTree::Ptr;
  using Ptr = std::shared_ptr<Registration<PointSource, PointTarget, Scalar>>;
  Registration() : tree_(new KdTree) { final_transformation_.setIdentity(); }
  virtual ~Registration() {}
  virtual void setInputSource(const PointCloudSourceConstPtr& cloud) { input_ = cloud; }
  virtual void setInputTarget(const PointCloudTargetConstPtr& cloud) {
    target_ = cloud;
    target_cloud_updated_ = true;
  }
  // registration.h (PCL 1.10) setSearchMethodTarget
  void setSearchMethodTarget(const KdTreePtr& tree, bool force_no_recompute = false) {
    tree_ = tree;
    if (force_no_recompute) force_no_recompute_ = true;
    target_cloud_updated_ = true;
  }
  KdTreePtr getSearchMethodTarget() const { return tree_; }
  void setMaximumIterations(int n) { max_iterations_ = n; }
  void setTransformationEpsilon(double e) { transformation_epsilon_ = e; }
  void setMaxCorrespondenceDistance(double d) { corr_dist_threshold_ = d; }
  bool hasConverged() const { return converged_; }
  Matrix4 getFinalTransformation() const { return final_transformation_; }
  void align(PointCloudSource& output) { align(output, Matrix4::Identity()); }
  void align(PointCloudSource& output, const Matrix4& guess) {
    if (!initCompute()) return;
    converged_ = false;
    final_transformation_.setIdentity();
    output.points = input_->points;  // PCL resizes the output and copies the input's fields, then the subclass transforms xyz
    computeTransformation(output, guess);
  }
  // registration.hpp (PCL 1.10) getFitnessScore(max_range): the input transformed by final_transformation_ (here with
  // the association of PCL's SSE transformer, (c0 x + c1 y) + (c2 z + c3)), then one tree query per point
  double getFitnessScore(double max_range = std::numeric_limits<double>::max()) {
    double sum = 0;
    int nr = 0;
    const Matrix4& T = final_transformation_;
    std::vector<int> nn_indices(1);
    std::vector<float> nn_dists(1);
    for (const auto& p : input_->points) {
      PointSource q = p;
      q.x = (T(0, 0) * p.x + T(0, 1) * p.y) + (T(0, 2) * p.z + T(0, 3));
      q.y = (T(1, 0) * p.x + T(1, 1) * p.y) + (T(1, 2) * p.z + T(1, 3));
      q.z = (T(2, 0) * p.x + T(2, 1) * p.y) + (T(2, 2) * p.z + T(2, 3));
      tree_->nearestKSearch(q, 1, nn_indices, nn_dists);
      if ((double)nn_dists[0] <= max_range) { sum += nn_dists[0]; nr++; }
    }
    return nr ? sum / nr : std::numeric_limits<double>::max();
  }
protected:
  // registration.hpp (PCL 1.10) initCompute: only a NEW target rebuilds the tree, and only if the tree may be recomputed
  bool initCompute() {
    if (!target_ || !input_) return false;
    if (target_cloud_updated_ && !force_no_recompute_) {
      tree_->setInputCloud(target_);
      target_cloud_updated_ = false;
    }
    return true;
  }
  virtual void computeTransformation(PointCloudSource& output, const Matrix4& guess) = 0;
  std::string reg_name_;
  KdTreePtr tree_;
  PointCloudSourceConstPtr input_;
  PointCloudTargetConstPtr target_;
  int nr_iterations_ = 0;
  int max_iterations_ = 10;
  Matrix4 final_transformation_;
  double transformation_epsilon_ = 0.0;
  double corr_dist_threshold_ = std::sqrt(std::numeric_limits<double>::max());
  bool converged_ = false;
  bool target_cloud_updated_ = true;
  bool force_no_recompute_ = false;
};
}  // namespace pcl
#endif
