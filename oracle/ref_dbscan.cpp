// TEST INFRASTRUCTURE — runs the REFERENCE's own DBSCAN code. The two headers below are compiled where they lie under
// /root/reference/4DRadarSLAM/include (never copied) against the PCL stand-in of oracle/ref_stubs; only the neighbour
// search behind search_method_ is ours ([ext] FLANN semantics, ref_stubs/pcl/search/kdtree.h). The cluster ranking that
// follows extract() lives inside a ROS callback (apps/preprocessing_nodelet_ntu.cpp:536-567) and is restated here.
// Output: tests/golden/dbscan_reference.npz (tests/golden/make_dbscan_reference.py), the fixture that pins both the
// NumPy transcription (tests/numpy_restatement.py: dbscan_labels) and apd_dbscan_labels.
#include <pcl/search/kdtree.h>

#include <dbscan/DBSCAN_kdtree.h>

extern "C" int apdo_ref_dbscan(const float* xyz, int n, double eps, int core_min_pts, int min_cluster, int max_cluster, float* labels,
                               int* cluster_of, int* n_clusters) {
  using PointT = pcl::PointXYZINormal;
  pcl::PointCloud<PointT>::Ptr cloud(new pcl::PointCloud<PointT>());
  cloud->points.resize((std::size_t)n);
  for (int i = 0; i < n; i++) {
    cloud->points[(std::size_t)i].x = xyz[3 * i];
    cloud->points[(std::size_t)i].y = xyz[3 * i + 1];
    cloud->points[(std::size_t)i].z = xyz[3 * i + 2];
  }
  pcl::search::KdTree<PointT>::Ptr kdtree(new pcl::search::KdTree<PointT>());
  kdtree->setInputCloud(cloud);
  std::vector<pcl::PointIndices> cluster_indices;
  DBSCANKdtreeCluster<PointT> ec;  // preprocessing_nodelet_ntu.cpp:521-530
  ec.setCorePointMinPts(core_min_pts);
  ec.setClusterTolerance(eps);
  ec.setMinClusterSize(min_cluster);
  ec.setMaxClusterSize(max_cluster);
  ec.setSearchMethod(kdtree);
  ec.setInputCloud(cloud);
  ec.extract(cluster_indices);
  // :533-567 (restated): centroid range of every cluster, ascending order, label = rank + 1
  std::vector<std::pair<int, float>> cluster_distances;
  for (std::size_t i = 0; i < cluster_indices.size(); ++i) {
    float sum_x = 0, sum_y = 0, sum_z = 0;
    const int num_points = (int)cluster_indices[i].indices.size();
    for (int idx : cluster_indices[i].indices) {
      const PointT& point = cloud->points[(std::size_t)idx];
      sum_x += point.x;
      sum_y += point.y;
      sum_z += point.z;
    }
    cluster_distances.emplace_back((int)i, std::hypot(sum_x / num_points, sum_y / num_points, sum_z / num_points));
  }
  std::stable_sort(cluster_distances.begin(), cluster_distances.end(),
                   [](const std::pair<int, float>& a, const std::pair<int, float>& b) { return a.second < b.second; });
  for (int i = 0; i < n; i++) {
    labels[i] = 0.f;
    if (cluster_of) cluster_of[i] = -1;
  }
  for (std::size_t rank = 0; rank < cluster_distances.size(); ++rank)
    for (int idx : cluster_indices[(std::size_t)cluster_distances[rank].first].indices) {
      labels[idx] = static_cast<float>(rank + 1);
      if (cluster_of) cluster_of[idx] = (int)rank;
    }
  if (n_clusters) *n_clusters = (int)cluster_indices.size();
  return 0;
}
