// Drives the FastAPDGICP shim exactly the way 4DRadarSLAM does:
// select_registration_method (registrations.cpp:38-51) constructs and configures
// it, then the nodelets use it through pcl::Registration::Ptr
// (scan_matching_odometry_nodelet.cpp:430-490, loop_detector.cpp:222-236).
// usage: test_shim source.f32 n_source target.f32 n_target [gicp]   (float32 [n,4] = x,y,z,label)
// "gicp": the factory's FAST_GICP branch (registrations.cpp:28-37) instead of FAST_APDGICP
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cstring>

#include <fast_gicp/gicp/fast_apdgicp.hpp>
#include <fast_gicp/gicp/fast_gicp.hpp>

using PointT = pcl::PointXYZINormal;

static pcl::PointCloud<PointT>::Ptr load(const char* path, int n) {
  pcl::PointCloud<PointT>::Ptr c(new pcl::PointCloud<PointT>());
  std::vector<float> raw((size_t)n * 4);
  FILE* f = std::fopen(path, "rb");
  if (!f || std::fread(raw.data(), sizeof(float), raw.size(), f) != raw.size()) { std::fprintf(stderr, "cannot read %s\n", path); std::exit(2); }
  std::fclose(f);
  c->points.resize(n);
  for (int i = 0; i < n; i++) {
    c->points[i].x = raw[4 * i]; c->points[i].y = raw[4 * i + 1]; c->points[i].z = raw[4 * i + 2];
    c->points[i].normal_x = raw[4 * i + 3]; c->points[i].intensity = 1.f;
  }
  return c;
}

// the reference factory, registrations.cpp:38-51, with the deployed rosparam values (launch/ntu_loop2.launch:88-99)
static pcl::Registration<PointT, PointT>::Ptr select_registration_method(bool fast_gicp_branch) {
  if (fast_gicp_branch) {  // registrations.cpp:28-37
    fast_gicp::FastGICP<PointT, PointT>::Ptr gicp(new fast_gicp::FastGICP<PointT, PointT>());
    gicp->setNumThreads(0);
    gicp->setTransformationEpsilon(0.1);
    gicp->setMaximumIterations(64);
    gicp->setMaxCorrespondenceDistance(2.0);
    gicp->setCorrespondenceRandomness(20);
    return gicp;
  }
  fast_gicp::FastAPDGICP<PointT, PointT>::Ptr apdgicp(new fast_gicp::FastAPDGICP<PointT, PointT>());
  apdgicp->setNumThreads(0);
  apdgicp->setTransformationEpsilon(0.1);
  apdgicp->setMaximumIterations(64);
  apdgicp->setMaxCorrespondenceDistance(2.0);
  apdgicp->setCorrespondenceRandomness(20);
  apdgicp->setDistVar(0.86);
  apdgicp->setAzimuthVar(0.5);
  apdgicp->setElevationVar(1.0);
  return apdgicp;
}

int main(int argc, char** argv) {
  if (argc < 5) return 2;
  auto source = load(argv[1], std::atoi(argv[2]));
  auto target = load(argv[3], std::atoi(argv[4]));
  pcl::Registration<PointT, PointT>::Ptr registration = select_registration_method(argc > 5 && std::strcmp(argv[5], "gicp") == 0);
  registration->setInputTarget(target);
  registration->setInputSource(source);
  pcl::PointCloud<PointT>::Ptr aligned(new pcl::PointCloud<PointT>());
  registration->align(*aligned);
  const Eigen::Matrix4f T = registration->getFinalTransformation();
  auto* apd = dynamic_cast<fast_gicp::FastAPDGICP<PointT, PointT>*>(registration.get());
  int inliers = 0;
  const double fit_gpu = apd->getFitnessScoreGPU(std::numeric_limits<double>::max(), &inliers);
  const double cost = apd->evaluateCost(T);
  std::printf("{\"converged\": %d, \"T\": [", registration->hasConverged() ? 1 : 0);
  for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) std::printf("%.9g%s", T(r, c), (r == 3 && c == 3) ? "" : ", ");
  std::printf("], \"fitness_gpu\": %.17g, \"inliers\": %d, \"cost\": %.17g, \"aligned0\": [%.9g, %.9g, %.9g], \"n_aligned\": %zu, \"n_cov\": %zu}\n",
              fit_gpu, inliers, cost, (*aligned)[0].x, (*aligned)[0].y, (*aligned)[0].z, aligned->size(), apd->getTargetCovariances().size());
  // swap + re-align exercises swapSourceAndTarget as gicp_test.cpp:157-200 does
  apd->swapSourceAndTarget();
  registration->align(*aligned);
  std::fprintf(stderr, "swapped: converged=%d\n", registration->hasConverged() ? 1 : 0);
  return 0;
}
