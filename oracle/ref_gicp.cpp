// TEST INFRASTRUCTURE — runs the REFERENCE's own FastGICP and FastVGICP code (the factory's "FAST_GICP" / "FAST_VGICP",
// 4DRadarSLAM registrations.cpp:28-37, :64-72). Compiled where they lie under /root/reference/fast_apdgicp/include against
// the stand-ins of oracle/ref_stubs (see ref_apdgicp.cpp):
//   fast_gicp/gicp/fast_gicp.hpp + impl/fast_gicp_impl.hpp
//   fast_gicp/gicp/fast_vgicp.hpp + impl/fast_vgicp_impl.hpp + fast_vgicp_voxel.hpp
// Output: tests/golden/gicp_reference.npz (tests/golden/make_gicp_reference.py).
#include <fast_gicp/gicp/fast_gicp.hpp>
#include <fast_gicp/gicp/fast_vgicp.hpp>
#include <fast_gicp/gicp/impl/lsq_registration_impl.hpp>
#include <fast_gicp/gicp/impl/fast_gicp_impl.hpp>
#include <fast_gicp/gicp/impl/fast_vgicp_impl.hpp>

#include <map>

namespace {
using PointT = pcl::PointXYZINormal;
using Cloud = pcl::PointCloud<PointT>;

template <class Base>
struct ProbeT : public Base {
  Cloud::Ptr src{new Cloud()}, tgt{new Cloud()};
  void set_optimizer(int gn) { this->lsq_optimizer_type_ = gn ? fast_gicp::LSQ_OPTIMIZER_TYPE::GaussNewton : fast_gicp::LSQ_OPTIMIZER_TYPE::LevenbergMarquardt; }
  void ensure_covs() {  // the first lines of FastGICP::computeTransformation (fast_gicp_impl.hpp:115-120)
    if (this->source_covs_.size() != this->input_->size()) this->template calculate_covariances<PointT>(this->input_, *this->source_kdtree_, this->source_covs_);
    if (this->target_covs_.size() != this->target_->size()) this->template calculate_covariances<PointT>(this->target_, *this->target_kdtree_, this->target_covs_);
  }
  double do_linearize(const Eigen::Isometry3d& T, Eigen::Matrix<double, 6, 6>* H, Eigen::Matrix<double, 6, 1>* b) { return this->linearize(T, H, b); }
  double do_compute_error(const Eigen::Isometry3d& T) { return this->compute_error(T); }
  int iterations() const { return this->nr_iterations_; }
};
struct GicpProbe : ProbeT<fast_gicp::FastGICP<PointT, PointT>> {
  const std::vector<int>& corr() const { return this->correspondences_; }
  const std::vector<float>& sqd() const { return this->sq_distances_; }
  const std::vector<Eigen::Matrix4d>& maha() const { return this->mahalanobis_; }
};
struct VgicpProbe : ProbeT<fast_gicp::FastVGICP<PointT, PointT>> {
  using VoxelPtr = fast_gicp::GaussianVoxel::Ptr;
  // the voxels of the map, in ascending (z, y, x) order of their coordinates (found through the target's own points)
  std::vector<std::pair<std::array<int, 3>, VoxelPtr>> voxels() const {
    std::map<std::array<int, 3>, VoxelPtr> m;  // key (z, y, x)
    if (!this->voxelmap_) return {};
    for (const auto& p : this->target_->points) {
      const Eigen::Vector3i c = this->voxelmap_->voxel_coord(p.getVector4fMap().template cast<double>());
      m[{c[2], c[1], c[0]}] = this->voxelmap_->lookup_voxel(c);
    }
    std::vector<std::pair<std::array<int, 3>, VoxelPtr>> out;
    for (auto& kv : m) out.push_back({{kv.first[2], kv.first[1], kv.first[0]}, kv.second});
    return out;
  }
  const std::vector<std::pair<int, VoxelPtr>>& vcorr() const { return this->voxel_correspondences_; }
  const std::vector<Eigen::Matrix4d>& vmaha() const { return this->voxel_mahalanobis_; }
};

struct Handle {
  int variant;  // 1 FastGICP, 2 FastVGICP
  GicpProbe g;
  VgicpProbe v;
};

Eigen::Isometry3d pose_of(const double* T16) {
  Eigen::Matrix4d m;
  for (int r = 0; r < 4; r++)
    for (int c = 0; c < 4; c++) m(r, c) = T16[c * 4 + r];
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
  }
}
template <class P, class F>
auto with(Handle* h, F f) {
  return h->variant == 1 ? f(h->g) : f(h->v);
}
#define APPLY(h, expr)            \
  do {                            \
    if ((h)->variant == 1) {      \
      auto& p = (h)->g;           \
      expr;                       \
    } else {                      \
      auto& p = (h)->v;           \
      expr;                       \
    }                             \
  } while (0)
}  // namespace

extern "C" {
void* gref_create(int variant) {
  Handle* h = new Handle();
  h->variant = variant;
  return h;
}
void gref_destroy(void* h) { delete static_cast<Handle*>(h); }
void gref_set_params(void* hh, int k, int regularization, double max_corr_dist, int max_iterations, int gauss_newton, double rot_eps, double trans_eps,
                     int threads, double voxel_resolution, int voxel_search, int voxel_mode) {
  Handle* h = static_cast<Handle*>(hh);
  APPLY(h, {
    p.setNumThreads(threads);
    p.setCorrespondenceRandomness(k);
    p.setRegularizationMethod(static_cast<fast_gicp::RegularizationMethod>(regularization));
    p.setMaxCorrespondenceDistance(max_corr_dist);
    p.setMaximumIterations(max_iterations);
    p.set_optimizer(gauss_newton);
    p.setRotationEpsilon(rot_eps);
    p.setTransformationEpsilon(trans_eps);
  });
  if (h->variant == 2) {
    h->v.setResolution(voxel_resolution);
    h->v.setNeighborSearchMethod(static_cast<fast_gicp::NeighborSearchMethod>(voxel_search));
    h->v.setVoxelAccumulationMode(static_cast<fast_gicp::VoxelAccumulationMode>(voxel_mode));
  }
}
void gref_set_clouds(void* hh, const float* src, int ns, const float* tgt, int nt) {
  Handle* h = static_cast<Handle*>(hh);
  APPLY(h, {
    p.tgt.reset(new Cloud());
    fill(*p.tgt, tgt, nt);
    p.setInputTarget(p.tgt);
    p.src.reset(new Cloud());
    fill(*p.src, src, ns);
    p.setInputSource(p.src);
  });
}
void gref_swap(void* hh) {
  Handle* h = static_cast<Handle*>(hh);
  APPLY(h, {
    p.swapSourceAndTarget();
    std::swap(p.src, p.tgt);
  });
}
double gref_linearize(void* hh, const double* T16, double* H36, double* b6) {
  Handle* h = static_cast<Handle*>(hh);
  Eigen::Matrix<double, 6, 6> H;
  Eigen::Matrix<double, 6, 1> b;
  double e = 0;
  APPLY(h, {
    p.ensure_covs();
    e = p.do_linearize(pose_of(T16), &H, &b);
  });
  for (int i = 0; i < 36; i++) H36[i] = H.a[i];
  for (int i = 0; i < 6; i++) b6[i] = b.a[i];
  return e;
}
double gref_compute_error(void* hh, const double* T16) {
  Handle* h = static_cast<Handle*>(hh);
  double e = 0;
  APPLY(h, e = p.do_compute_error(pose_of(T16)));
  return e;
}
// FastGICP: correspondences_, sq_distances_, mahalanobis_ (3x3 blocks)
void gref_get_correspondences(void* hh, int* idx, float* sqd, double* maha9) {
  GicpProbe& p = static_cast<Handle*>(hh)->g;
  for (std::size_t i = 0; i < p.corr().size(); i++) {
    idx[i] = p.corr()[i];
    sqd[i] = p.sqd()[i];
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) maha9[9 * i + 3 * r + c] = idx[i] >= 0 ? p.maha()[i](r, c) : 0.0;
  }
}
// FastVGICP: the voxel map (ascending (z, y, x)); call with coords = NULL for the count
int gref_get_voxels(void* hh, int* coords, int* counts, double* means, double* covs9) {
  VgicpProbe& p = static_cast<Handle*>(hh)->v;
  const auto vox = p.voxels();
  if (coords)
    for (std::size_t i = 0; i < vox.size(); i++) {
      for (int a = 0; a < 3; a++) {
        coords[3 * i + a] = vox[i].first[(std::size_t)a];
        means[3 * i + a] = vox[i].second->mean[a];
      }
      counts[i] = vox[i].second->num_points;
      for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) covs9[9 * i + 3 * r + c] = vox[i].second->cov(r, c);
    }
  return (int)vox.size();
}
// FastVGICP: voxel_correspondences_ as (source index, voxel index in the list above) pairs in list order, with the 3x3
// block of voxel_mahalanobis_; call with src_idx = NULL for the count
int gref_get_voxel_correspondences(void* hh, int* src_idx, int* voxel_idx, double* maha9) {
  VgicpProbe& p = static_cast<Handle*>(hh)->v;
  if (src_idx) {
    const auto vox = p.voxels();
    std::map<const void*, int> index;
    for (std::size_t i = 0; i < vox.size(); i++) index[vox[i].second.get()] = (int)i;
    for (std::size_t i = 0; i < p.vcorr().size(); i++) {
      src_idx[i] = p.vcorr()[i].first;
      voxel_idx[i] = index.at(p.vcorr()[i].second.get());
      for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) maha9[9 * i + 3 * r + c] = p.vmaha()[i](r, c);
    }
  }
  return (int)p.vcorr().size();
}
void gref_align(void* hh, const float* guess, float* T_out, int* converged, int* iterations, double* H36) {
  Handle* h = static_cast<Handle*>(hh);
  Eigen::Matrix4f g = Eigen::Matrix4f::Identity();
  if (guess)
    for (int r = 0; r < 4; r++)
      for (int c = 0; c < 4; c++) g(r, c) = guess[c * 4 + r];
  APPLY(h, {
    Cloud out;
    p.align(out, g);
    const Eigen::Matrix4f T = p.getFinalTransformation();
    for (int r = 0; r < 4; r++)
      for (int c = 0; c < 4; c++) T_out[c * 4 + r] = T(r, c);
    *converged = p.hasConverged() ? 1 : 0;
    *iterations = p.iterations();
    const auto& Hf = p.getFinalHessian();
    for (int i = 0; i < 36; i++) H36[i] = Hf.a[i];
  });
}
}
