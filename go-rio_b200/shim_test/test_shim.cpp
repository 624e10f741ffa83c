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
#include <fast_gicp/gicp/fast_vgicp.hpp>

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
static pcl::Registration<PointT, PointT>::Ptr select_registration_method(const char* method) {
  const bool fast_gicp_branch = std::strcmp(method, "gicp") == 0;
  if (std::strcmp(method, "vgicp") == 0) {  // registrations.cpp:64-72 ("FAST_VGICP"; reg_resolution from the test)
    fast_gicp::FastVGICP<PointT, PointT>::Ptr vgicp(new fast_gicp::FastVGICP<PointT, PointT>());
    vgicp->setNumThreads(0);
    vgicp->setResolution(2.0);
    vgicp->setNeighborSearchMethod(fast_gicp::NeighborSearchMethod::DIRECT7);
    vgicp->setTransformationEpsilon(0.1);
    vgicp->setMaximumIterations(64);
    vgicp->setCorrespondenceRandomness(20);
    return vgicp;
  }
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

static long long knn_launches(apd_handle* h) {
  double ms[APD_K_COUNT];
  int64_t launches[APD_K_COUNT];
  apd_get_kernel_ms(h, ms, launches);
  return (long long)launches[APD_K_KNN_COV];
}

int main(int argc, char** argv) {
  if (argc < 5) return 2;
  auto source = load(argv[1], std::atoi(argv[2]));
  auto target = load(argv[3], std::atoi(argv[4]));
  pcl::Registration<PointT, PointT>::Ptr registration = select_registration_method(argc > 5 ? argv[5] : "apdgicp");
  auto* apd = dynamic_cast<fast_gicp::FastAPDGICP<PointT, PointT>*>(registration.get());
  const long long knn0 = knn_launches(apd->handle());
  registration->setInputTarget(target);
  registration->setInputSource(source);
  pcl::PointCloud<PointT>::Ptr aligned(new pcl::PointCloud<PointT>());
  registration->align(*aligned);
  const long long knn_first = knn_launches(apd->handle()) - knn0;  // both clouds' covariances
  const Eigen::Matrix4f T = registration->getFinalTransformation();
  int inliers = 0;
  const double fit_gpu = apd->getFitnessScoreGPU(std::numeric_limits<double>::max(), &inliers);
  const double cost = apd->evaluateCost(T);

  // --- the callers' use of the base class's search method (scan_matching_odometry_nodelet.cpp:674-691, loop_detector.cpp:229) ---
  const double fit_pcl = registration->getFitnessScore();  // pcl::Registration's loop over tree_->nearestKSearch
  int nodelet_inliers = 0;
  {
    std::vector<int> k_indices;
    std::vector<float> k_sq_dists;
    for (std::size_t i = 0; i < aligned->size(); i++) {
      registration->getSearchMethodTarget()->nearestKSearch(aligned->at(i), 1, k_indices, k_sq_dists);
      if (k_sq_dists[0] < 0.5f * 0.5f) nodelet_inliers++;
    }
  }
  const auto& search = apd->targetSearch();
  const unsigned long long passes = search.batchPasses(), hits = search.batchHits(), singles = search.singleQueries();
  // a query that is not one of the transformed source points, k = 5: the GPU search against brute force
  int knn_equal = 1;
  {
    PointT q = target->at(target->size() / 3);
    q.x += 0.37f; q.y -= 0.21f; q.z += 0.11f;
    std::vector<int> gi, bi;
    std::vector<float> gd, bd;
    registration->getSearchMethodTarget()->nearestKSearch(q, 5, gi, gd);
    pcl::search::KdTree<PointT> brute;
    brute.setInputCloud(target);
    pcl::search::KdTree<PointT>::builds()--;  // (the checker's own tree does not count)
    brute.nearestKSearch(q, 5, bi, bd);
    for (int j = 0; j < 5; j++) knn_equal = knn_equal && gi[j] == bi[j] && gd[j] == bd[j];
  }
  const int tree_builds = pcl::search::KdTree<PointT>::builds();  // CPU kd-tree builds by pcl::Registration::align so far

  std::printf("{\"converged\": %d, \"T\": [", registration->hasConverged() ? 1 : 0);
  for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) std::printf("%.9g%s", T(r, c), (r == 3 && c == 3) ? "" : ", ");
  std::printf("], \"fitness_gpu\": %.17g, \"inliers\": %d, \"cost\": %.17g, \"aligned0\": [%.9g, %.9g, %.9g], \"n_aligned\": %zu, \"n_cov\": %zu, ",
              fit_gpu, inliers, cost, (*aligned)[0].x, (*aligned)[0].y, (*aligned)[0].z, aligned->size(), apd->getTargetCovariances().size());
  std::printf("\"fitness_pcl\": %.17g, \"nodelet_inliers\": %d, \"search_passes\": %llu, \"search_hits\": %llu, \"search_singles\": %llu, "
              "\"knn_equal\": %d, \"tree_builds\": %d, \"knn_launches_first\": %lld, ",
              fit_pcl, nodelet_inliers, passes, hits, singles, knn_equal, tree_builds, knn_first);

  // --- keyframe promotion (scan_matching_odometry_nodelet.cpp:587-588): the registered source becomes the target, a new
  // scan the source. The promoted cloud's grid and covariances stay on the device: only the new scan's are computed.
  pcl::PointCloud<PointT>::Ptr scan2(new pcl::PointCloud<PointT>(*aligned));
  const long long knn1 = knn_launches(apd->handle());
  registration->setInputTarget(source);
  registration->setInputSource(scan2);
  registration->align(*aligned);
  const long long knn_promoted = knn_launches(apd->handle()) - knn1;
  const Eigen::Matrix4f Tp = registration->getFinalTransformation();
  // the same pair on a fresh object (nothing to adopt): must give the same pose
  pcl::Registration<PointT, PointT>::Ptr fresh = select_registration_method(argc > 5 ? argv[5] : "apdgicp");
  pcl::PointCloud<PointT>::Ptr source_copy(new pcl::PointCloud<PointT>(*source));
  fresh->setInputTarget(source_copy);
  fresh->setInputSource(scan2);
  pcl::PointCloud<PointT>::Ptr aligned2(new pcl::PointCloud<PointT>());
  fresh->align(*aligned2);
  const Eigen::Matrix4f Tf = fresh->getFinalTransformation();
  int same_pose = 1;
  for (int i = 0; i < 16; i++) same_pose = same_pose && Tp.data()[i] == Tf.data()[i];
  std::printf("\"knn_launches_promoted\": %lld, \"promoted_same_pose\": %d, \"tree_builds_end\": %d}\n", knn_promoted, same_pose,
              pcl::search::KdTree<PointT>::builds());

  // swap + re-align exercises swapSourceAndTarget as gicp_test.cpp:157-200 does
  apd->swapSourceAndTarget();
  registration->align(*aligned);
  std::fprintf(stderr, "swapped: converged=%d\n", registration->hasConverged() ? 1 : 0);
  return 0;
}
