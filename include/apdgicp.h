/*
 * apdgicp.h — C-ABI of the B200-native FastAPDGICP registration path.
 *
 * This is the drop-in boundary: everything the reference class
 * fast_gicp::FastAPDGICP<pcl::PointXYZINormal, pcl::PointXYZINormal>
 * (reference fast_apdgicp/include/fast_gicp/gicp/fast_apdgicp.hpp:19-122) and
 * its optimizer base LsqRegistration (lsq_registration.hpp:15-85) do on the hot
 * path is reachable through these entry points. The C++ shim
 * go-rio_b200/include/fast_gicp/gicp/fast_apdgicp.hpp re-creates the reference
 * class on top of them; INTEGRATION.md shows the binding.
 *
 * Rules of the boundary
 *  - extern "C", opaque handle, plain pointers and sizes, int status returns.
 *  - No exception and no C++ type crosses it. All buffers are caller-owned
 *    HOST memory unless a parameter is explicitly named `d_*` (device).
 *  - One CUDA stream per handle. Calls on one handle are not thread-safe,
 *    calls on different handles are (the reference has the same contract: one
 *    instance per nodelet callback thread).
 *  - There is no CPU fallback: every compute entry point fails with
 *    APD_ERR_CUDA when no sm_100 device is usable.
 *
 * Matrix conventions
 *  - 4x4 transforms and 4x4 covariances are COLUMN-MAJOR (Eigen's default
 *    layout, so Eigen::Matrix4f::data() / Matrix4d::data() can be passed as is).
 *  - H is 6x6 column-major (symmetric, so the order is immaterial), b is 6.
 */
#ifndef APDGICP_H_
#define APDGICP_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define APD_ABI_VERSION 3

#if defined(__GNUC__)
#define APD_API __attribute__((visibility("default")))
#else
#define APD_API
#endif

/* status codes */
enum {
  APD_OK = 0,
  APD_ERR_INVALID = 1,   /* bad argument / state (e.g. align without clouds) */
  APD_ERR_CUDA = 2,      /* CUDA runtime error, see apd_last_error           */
  APD_ERR_TOO_FEW = 3,   /* cloud has fewer than k points (reference: UB,
                            fast_apdgicp_impl.hpp:364-369; here: an error)   */
  APD_ERR_UNSUPPORTED = 4,
  APD_ERR_COMM = 5       /* NCCL error                                       */
};

/* reference gicp_settings.hpp:6 — same order, same values */
enum {
  APD_REG_NONE = 0,
  APD_REG_MIN_EIG = 1,
  APD_REG_NORMALIZED_MIN_EIG = 2,
  APD_REG_PLANE = 3,
  APD_REG_FROBENIUS = 4
};

/* Which cost the registration minimises. APDGICP: fast_apdgicp_impl.hpp (radar noise model in the combined
 * covariance, :194-215; per-point weight 1 + geometric + label weight, :266-276). GICP: the reference's FastGICP
 * (fast_gicp_impl.hpp:124-258, selected by registrations.cpp:28-37): the same covariances, correspondences and
 * optimizer, RCR = C_B + T C_A T^T (:157) and unit weights (:205) — dist/azimuth/elevation_var are not read. */
enum { APD_VARIANT_APDGICP = 0, APD_VARIANT_GICP = 1, APD_VARIANT_VGICP = 2 };
/* fast_gicp::NeighborSearchMethod / VoxelAccumulationMode (reference gicp_settings.hpp:10-12), FastVGICP only */
enum { APD_VOXEL_DIRECT27 = 0, APD_VOXEL_DIRECT7 = 1, APD_VOXEL_DIRECT1 = 2 };
enum { APD_VOXEL_ADDITIVE = 0, APD_VOXEL_ADDITIVE_WEIGHTED = 1, APD_VOXEL_MULTIPLICATIVE = 2 };

/* reference lsq_registration.hpp:13 — same order */
enum { APD_OPT_GAUSS_NEWTON = 0, APD_OPT_LEVENBERG_MARQUARDT = 1 };

/* Every tunable the reference class holds on this path.
 * Defaults (apd_default_params) are the reference constructors' values:
 * fast_apdgicp_impl.hpp:14-28, fast_apdgicp.hpp:116-118,
 * lsq_registration_impl.hpp:11-24. */
typedef struct apd_params {
  int32_t k_correspondences;           /* setCorrespondenceRandomness, 20     */
  int32_t regularization;              /* setRegularizationMethod, PLANE      */
  double max_correspondence_distance;  /* pcl corr_dist_threshold_, FLT_MAX   */
  double dist_var;                     /* setDistVar, 0.86                    */
  double azimuth_var;                  /* setAzimuthVar (degrees), 0.5        */
  double elevation_var;                /* setElevationVar (degrees), 1.0      */
  int32_t max_iterations;              /* pcl max_iterations_, 64             */
  int32_t optimizer;                   /* lsq_optimizer_type_, LM             */
  double rotation_epsilon;             /* setRotationEpsilon, 2e-3            */
  double transformation_epsilon;       /* setTransformationEpsilon, 5e-4      */
  int32_t lm_max_iterations;           /* lm_max_iterations_, 10              */
  int32_t lm_debug_print;              /* setDebugPrint, 0                    */
  double lm_init_lambda_factor;        /* setInitialLambdaFactor, 1e-9        */
  int32_t maha_fp64;                   /* 0: store the per-point Mahalanobis
                                          matrix as 6 x fp32 (24 B, default),
                                          1: as 6 x fp64 (48 B). Arithmetic is
                                          fp64 either way.                    */
  int32_t host_loop;                   /* 0 (default): the optimizer loop of
                                          apd_align runs ON THE DEVICE in one
                                          kernel launch when the source cloud has
                                          at most 32768 points and the handle is
                                          not sharded; 1: always drive it from the
                                          host, one launch per stage (what large
                                          or sharded clouds use anyway).         */
  int32_t variant;                     /* APD_VARIANT_APDGICP (default) / _GICP /
                                          _VGICP (fast_gicp::FastVGICP, reference
                                          impl/fast_vgicp_impl.hpp)              */
  int32_t voxel_search;                /* FastVGICP setNeighborSearchMethod,
                                          APD_VOXEL_DIRECT1 (fast_vgicp_impl.hpp:23) */
  double voxel_resolution;             /* FastVGICP setResolution, 1.0 (:22)     */
  int32_t voxel_mode;                  /* FastVGICP setVoxelAccumulationMode,
                                          APD_VOXEL_ADDITIVE (:24)               */
  int32_t reserved_;                   /* 0                                      */
} apd_params;

typedef struct apd_handle apd_handle;

/* ---- life cycle ------------------------------------------------------- */
APD_API int apd_abi_version(void);
APD_API int apd_default_params(apd_params* out);
/* reference FastAPDGICP::FastAPDGICP (fast_apdgicp_impl.hpp:14-28) */
APD_API int apd_create(int device, apd_handle** out);
APD_API int apd_destroy(apd_handle* h);
/* text of the last error on this handle ("" if none). Never NULL. */
APD_API const char* apd_last_error(const apd_handle* h);
/* setters of fast_apdgicp_impl.hpp:34-65 + lsq_registration_impl.hpp:30-42 +
 * the pcl::Registration setters used by registrations.cpp:41-48 */
APD_API int apd_set_params(apd_handle* h, const apd_params* p);
APD_API int apd_get_params(const apd_handle* h, apd_params* out);

/* ---- clouds ----------------------------------------------------------- */
/* reference setInputSource / setInputTarget (fast_apdgicp_impl.hpp:115-135).
 * `pts` is host AoS: point i starts at pts + i*stride_bytes; three floats
 * x,y,z at +xyz_off; one float cluster label (pcl normal_x, written by the
 * DBSCAN stage, preprocessing_nodelet_ntu.cpp:561-567) at +label_off, or
 * label_off < 0 for "all zero". For pcl::PointXYZINormal: stride 48,
 * xyz_off 0, label_off 16.
 * cache_key: the reference early-outs on pointer identity (:116,:128); pass
 * the cloud pointer value (or any id); 0 disables the early-out. A key is only
 * honoured together with the point count and a fingerprint of 64 sampled
 * points, so a recycled address holding other points is a new cloud. A cloud
 * set with the key (and content) the OTHER slot holds is adopted device-to-
 * device with its grid and covariances (a source promoted to target,
 * scan_matching_odometry_nodelet.cpp:587-588). Setting a cloud drops its
 * cached covariances. */
APD_API int apd_set_source(apd_handle* h, const void* pts, int32_t n, int32_t stride_bytes,
                   int32_t xyz_off, int32_t label_off, uint64_t cache_key);
APD_API int apd_set_target(apd_handle* h, const void* pts, int32_t n, int32_t stride_bytes,
                   int32_t xyz_off, int32_t label_off, uint64_t cache_key);
/* same, but the cloud is already on the device as float4 {x,y,z,label}[n]
 * (no reference equivalent; used for HBM-resident benchmarking and sharding) */
APD_API int apd_set_source_device(apd_handle* h, const void* d_xyzl, int32_t n);
APD_API int apd_set_target_device(apd_handle* h, const void* d_xyzl, int32_t n);

/* reference swapSourceAndTarget / clearSource / clearTarget (:89-112) */
APD_API int apd_swap_source_and_target(apd_handle* h);
APD_API int apd_clear_source(apd_handle* h);
APD_API int apd_clear_target(apd_handle* h);

/* reference set/get{Source,Target}Covariances (:138-145, hpp:73-79):
 * n column-major 4x4 doubles (128 B each). Getters compute the covariances
 * first if they are stale (the reference getter would return an empty vector
 * before the first align; computing is a superset). */
APD_API int apd_set_source_covariances(apd_handle* h, const double* covs4x4, int32_t n);
APD_API int apd_set_target_covariances(apd_handle* h, const double* covs4x4, int32_t n);
APD_API int apd_get_source_covariances(apd_handle* h, double* covs4x4, int32_t n);
APD_API int apd_get_target_covariances(apd_handle* h, double* covs4x4, int32_t n);
/* parity hook: the k neighbour indices of every point, ordered by (d2, index);
 * out is int32[n*k]. which: 0 source, 1 target. */
APD_API int apd_get_neighbors(apd_handle* h, int32_t which, int32_t* out, int32_t n, int32_t k);

/* ---- the hot path ----------------------------------------------------- */
/* pcl::Registration::align(output, guess) -> FastAPDGICP::computeTransformation
 * (fast_apdgicp_impl.hpp:148-157) -> LsqRegistration::computeTransformation
 * (lsq_registration_impl.hpp:55-80).
 * guess: float[16] column-major, NULL = identity.
 * T_out: float[16] final_transformation_. T_out_f64: optional double[16], the
 * fp64 pose before the float cast. H_out: optional double[36] final_hessian_.
 * converged/iterations: hasConverged() / nr_iterations_ (= index of the last
 * outer iteration, as the reference sets it at :68) — optional.
 * aligned_xyz: optional float[3*n_source], the transformed source cloud
 * (pcl::transformPointCloud at :79), packed xyz. */
APD_API int apd_align(apd_handle* h, const float* guess, float* T_out, double* T_out_f64,
              double* H_out, int32_t* converged, int32_t* iterations, float* aligned_xyz);

/* FastAPDGICP::linearize (:224-307) incl. update_correspondences (:160-220);
 * T: double[16] column-major. H (36), b (6) may both be NULL (error only).
 * Equivalent to LsqRegistration::evaluateCost when T comes from a float pose.
 * Computes stale covariances first, as align does. */
APD_API int apd_linearize(apd_handle* h, const double* T, double* H, double* b, double* err);
/* FastAPDGICP::compute_error (:310-346): uses the correspondences and
 * Mahalanobis matrices of the LAST linearize. */
APD_API int apd_compute_error(apd_handle* h, const double* T, double* err);
/* FastAPDGICP::update_correspondences (:160-220) alone. */
APD_API int apd_update_correspondences(apd_handle* h, const double* T);
/* parity hook: correspondences_ (int32[n], -1 = none) and sq_distances_
 * (float[n]); either may be NULL. */
APD_API int apd_get_correspondences(apd_handle* h, int32_t* idx, float* sq_dist, int32_t n);
/* parity hook: mahalanobis_ as n column-major 4x4 doubles */
APD_API int apd_get_mahalanobis(apd_handle* h, double* maha4x4, int32_t n);

/* ---- FastVGICP parity hooks (variant = APD_VARIANT_VGICP) -------------------
 * The Gaussian voxel map of the target (reference fast_vgicp_voxel.hpp:127-185,
 * built by the first linearize of an alignment, fast_vgicp_impl.hpp:126-129):
 * voxels in ascending (z, y, x) order of their integer coordinates
 * floor(p / resolution - 0.5). coords: int32[3 * n], counts: int32[n], means:
 * double[3 * n], covs: double[9 * n] row-major 3x3; any may be NULL.
 * *n_voxels receives the number of voxels; at most `capacity` are written. */
APD_API int apd_vgicp_get_voxels(apd_handle* h, int32_t* n_voxels, int32_t* coords, int32_t* counts, double* means, double* covs,
                                 int32_t capacity);
/* voxel_correspondences_ / voxel_mahalanobis_ of the LAST linearize
 * (fast_vgicp_impl.hpp:74-118), per source point (original order) and
 * neighbour offset (the order of neighbor_offsets(), fast_vgicp_voxel.hpp:10-44):
 * voxel[i * n_offsets + o] = index into the list of apd_vgicp_get_voxels or -1;
 * maha: 9 doubles (row-major 3x3) per slot, zeros where there is no voxel. */
APD_API int apd_vgicp_get_correspondences(apd_handle* h, int32_t* voxel, double* maha3x3, int32_t n_source, int32_t n_offsets);

/* pcl::Registration::getFitnessScore(max_range) [PCL 1.10 registration.hpp]
 * over final_transformation_ (or T if not NULL, float[16]); also returns the
 * number of points with d2 <= max_range, and — for the status message of
 * scan_matching_odometry_nodelet.cpp:677-689 — the number of points whose
 * 1-NN squared distance is < inlier_sq_thr. */
APD_API int apd_fitness(apd_handle* h, const float* T, double max_range, double* score,
                int32_t* n_in_range, double inlier_sq_thr, int32_t* n_inliers);

/* The search method behind pcl::Registration::getSearchMethodTarget() (PCL 1.10 registration.h: tree_, a
 * pcl::search::KdTree the base class rebuilds over every new target in initCompute(), and that getFitnessScore() and
 * scan_matching_odometry_nodelet.cpp:684 query one point at a time) on the GPU grid this library builds anyway:
 * exact k nearest neighbours (1 <= k <= 32, ordered by (fp32 d2, index) like every search here) of n host query points
 * (x,y,z floats at queries + i*stride_bytes) in the cloud `which` (0 source, 1 target). idx / sq_dist: [n*k], original
 * point indices and fp32 squared distances; -1 / inf where the cloud has fewer than k points. */
APD_API int apd_nearest_k(apd_handle* h, int32_t which, const void* queries, int32_t n, int32_t stride_bytes, int32_t k, int32_t* idx,
                  float* sq_dist);
/* The same for what those callers actually ask: the nearest target point of EVERY source point under the pose T
 * (float[16] column-major; NULL = final_transformation_), in one pass, in the source's original order. xyz (optional,
 * float[3n]): the transformed points. The shim's search adaptor answers the per-point queries from this batch. */
APD_API int apd_source_nearest(apd_handle* h, const float* T, int32_t* idx, float* sq_dist, float* xyz, int32_t n);

/* ---- the stages on either side of the registration that use the same grid ------------------------ */
/* The radius searches of the preprocessing nodelet — pcl::RadiusOutlierRemoval (preprocessing_nodelet_ntu.cpp:163-172:
 * a point stays when its radiusSearch finds MORE than min_neighbors points, itself included) and the neighbour queries
 * of DBSCANKdtreeCluster (:520-532, eps 0.9, core points >= 10) — on the grid of the cloud `which` (0 source, 1 target),
 * every point of the cloud being a query. pcl / FLANN radiusSearch semantics: a point is a neighbour when its fp32
 * squared distance is STRICTLY below radius^2; the query point counts. counts: int32[n] (optional). offsets: int64[n+1]
 * (optional), the CSR row starts; indices: the neighbours' point indices, row by row (unordered within a row), capacity
 * entries — call once with indices = NULL to learn offsets[n], then again with the array. */
APD_API int apd_radius_search(apd_handle* h, int32_t which, double radius, int32_t* counts, int64_t* offsets, int32_t* indices,
                      int64_t capacity, int32_t n);
/* The cluster labels the registration reads in normal_x: DBSCANKdtreeCluster::extract
 * (4DRadarSLAM/include/dbscan/DBSCAN_simple.h:27-104; eps, core_min_pts, min / max cluster size as
 * preprocessing_nodelet_ntu.cpp:524-530 sets them: 0.9, 10, 20, 25000) followed by the nodelet's ranking (:536-567):
 * clusters ordered by the range of their centroid, every member labelled rank + 1 (a point of several clusters keeps the
 * last rank, as the nodelet's loop leaves it); labels[i] = 0 where the nodelet writes nothing. The radius searches — one
 * per point as a seed (radius |norm - 1| / 50 + eps) and one as an expansion ((norm - 1) / 100 + eps) — run on the GPU
 * grid of cloud `which`; the sequential growth over their lists runs on the host. labels: float[n]. */
APD_API int apd_dbscan_labels(apd_handle* h, int32_t which, double eps, int32_t core_min_pts, int32_t min_cluster, int32_t max_cluster,
                              float* labels, int32_t* n_clusters, int32_t n);
/* pcl::VoxelGrid<PointT> with a cubic leaf (PCL 1.10 voxel_grid.hpp, downsample_all_data): one point per occupied voxel,
 * voxels in ascending voxel index; x, y, z are the float mean of the voxel's points; the cluster label (normal_x) goes
 * through PCL's normal accumulator — summed and normalised — so it becomes 1 where any member had a positive label, 0
 * otherwise (normal_y = normal_z = 0 on this pipeline). out_xyzl: float4 {x,y,z,label}[capacity >= n]. When the leaf is
 * too small for the cloud's extent (voxel index overflow) PCL warns and passes the cloud through: so does this. */
APD_API int apd_voxel_downsample(apd_handle* h, const void* pts, int32_t n, int32_t stride_bytes, int32_t xyz_off, int32_t label_off,
                         double leaf, float* out_xyzl, int32_t capacity, int32_t* n_out);
/* The submap of the scan-to-map branch (scan_matching_odometry_nodelet.cpp:602-618): keyframe cloud c moved by
 * poses[c] (column-major 4x4 doubles, rel_pose = odom_c^-1 * odom_last; pcl::transformPointCloud with a Matrix4d: double
 * arithmetic, cast to float), concatenated in order, then voxel-downsampled with `leaf` (<= 0: not downsampled).
 * out_xyzl (optional): float4[capacity >= sum of the keyframe sizes]. set_as_target != 0: the submap also becomes the
 * handle's target cloud (registration_s2m->setInputTarget) without leaving the device. */
typedef struct apd_cloud_ref {
  const void* pts; /* host AoS, the layout arguments of apd_set_source */
  int32_t n;
} apd_cloud_ref;
APD_API int apd_submap_assemble(apd_handle* h, const apd_cloud_ref* clouds, const double* poses, int32_t n_clouds, int32_t stride_bytes,
                        int32_t xyz_off, int32_t label_off, double leaf, int32_t set_as_target, float* out_xyzl, int32_t capacity,
                        int32_t* n_out);

/* LM trace of the last apd_align: rows of {outer, inner, y0, yi, rho, lambda,
 * |d|, accepted} as 8 doubles, the columns of the reference's lm_debug_print_
 * table (lsq_registration_impl.hpp:148-154). Returns the number of rows
 * written (<= max_rows) in *n_rows. */
APD_API int apd_get_lm_trace(apd_handle* h, double* rows, int32_t max_rows, int32_t* n_rows);

/* ---- batched registrations (config C3; no reference equivalent: the loop
 * detector runs candidates serially, loop_detector.cpp:222-236) ---------- */
typedef struct apd_pair {
  const void* source;  /* host AoS, same layout arguments as apd_set_source  */
  int32_t n_source;
  const void* target;
  int32_t n_target;
  const float* guess;  /* float[16] column-major or NULL                     */
} apd_pair;

typedef struct apd_result {
  float T[16];         /* final_transformation_, column-major                */
  double fitness;      /* getFitnessScore(max_range = DBL_MAX)               */
  int32_t converged;
  int32_t iterations;
  int32_t status;      /* APD_OK or an error code for this pair              */
  int32_t n_inliers;   /* source points with 1-NN d2 < 0.25 m^2 (with_fitness) */
} apd_result;

/* A batch context: `n_workers` registrations in flight (1..512; an internal handle
 * with its own CUDA stream each), driven by a few host threads as non-blocking
 * state machines, that persist across calls, so device buffers, pinned staging
 * and streams are allocated once. 64 nearly saturate a B200 on scan-to-submap pairs
 * (96: +3.5 %).
 * (Streams only run concurrently if each has a hardware work queue, and the
 * driver reads CUDA_DEVICE_MAX_CONNECTIONS (default 8) when the CUDA context is
 * created: export CUDA_DEVICE_MAX_CONNECTIONS=32 for the process. The library
 * never changes the environment when it is loaded; the first apd_batch_create
 * sets the variable only if it is unset and no CUDA context exists yet, and
 * prints a note when it is too late.) apd_batch_align runs n_pairs independent
 * {clearTarget; clearSource; setInputTarget; setInputSource; align
 * [; getFitnessScore]} sequences over them: the staging and H2D copy of one
 * pair overlap the kernels of the others. Packed float4 {x,y,z,label} clouds
 * (stride 16, xyz_off 0, label_off 12) in page-locked host memory are copied
 * without a staging pass. With with_fitness != 0 the result
 * carries getFitnessScore(DBL_MAX) of the final pose and the number of source
 * points whose nearest target point is closer than 0.5 m (the inlier test of
 * scan_matching_odometry_nodelet.cpp:677-689). Results do not depend on
 * n_workers or on the order the pairs are taken. One call at a time per
 * context. */
typedef struct apd_batch apd_batch;
APD_API int apd_batch_create(int device, int32_t n_workers, apd_batch** out);
/* The same over several GPUs of the node from ONE process (apd_align_batch(..., n_devices)): n_workers
 * registrations in flight on EACH of the devices, and ONE shared queue of pairs — a worker of any device takes the next
 * pair when it is free, so pairs that need 3 and pairs that need 30 iterations (loop-closure candidates differ that much)
 * even out by themselves; a static pair -> device split cannot do that. Host clouds only (apd_batch_align); results
 * do not depend on which device took a pair. apd_batch_device_pairs: how many pairs of the last call each device took. */
APD_API int apd_batch_create_multi(const int32_t* devices, int32_t n_devices, int32_t n_workers, apd_batch** out);
APD_API int apd_batch_device_pairs(const apd_batch* b, int64_t* pairs, int32_t n_devices);
APD_API int apd_batch_destroy(apd_batch* b);
APD_API int apd_batch_set_params(apd_batch* b, const apd_params* p);
APD_API int apd_batch_align(apd_batch* b, const apd_pair* pairs, int32_t n_pairs, int32_t stride_bytes,
                    int32_t xyz_off, int32_t label_off, int32_t with_fitness, apd_result* results);
/* same, the clouds of every pair already on the device as float4 {x,y,z,label}
 * (apd_pair.source / .target are device pointers) */
APD_API int apd_batch_align_device(apd_batch* b, const apd_pair* pairs, int32_t n_pairs, int32_t with_fitness,
                           apd_result* results);
APD_API int64_t apd_batch_launch_count(const apd_batch* b);
/* What the pool did since the last reset, measured WHILE it runs at full load (no profiling events): stats[0] pooled
 * registrations, [1] milliseconds of device time summed over their loop kernels (each kernel stamps %globaltimer at its
 * start and end: x CTAs per registration / (296 CTA slots x wall time) = how full the loop keeps the GPU), [2..6] host
 * milliseconds spent per phase: set clouds, wait for the bounding boxes, enqueue grids + covariances, enqueue the loop,
 * wait for the result. n >= 7. */
APD_API int apd_batch_get_load_stats(apd_batch* b, double* stats, int32_t n, int32_t reset);
/* diagnostic: ONE loop-kernel launch over the pairs set on n handles (profiling the kernel under pool-like co-residency) */
APD_API int apd_debug_multi_align(apd_handle* const* handles, int32_t n, int32_t repeat);
/* diagnostic: empty-kernel launches per second from n_threads host threads over n_streams streams (per_second[0]: issue
 * rate, [1]: completion rate) — the driver's ceiling on a pool that issues several launches per registration */
APD_API int apd_debug_launch_rate(int device, int32_t n_streams, int32_t n_threads, int32_t launches_per_thread, double* per_second);
APD_API int apd_batch_set_profiling(apd_batch* b, int32_t enabled);
APD_API int apd_batch_get_kernel_ms(apd_batch* b, double* ms /* [APD_K_COUNT] */, int64_t* launches);

/* One-shot convenience: create a context of n_streams workers (per device), run, destroy. */
APD_API int apd_align_batch_multi(const int32_t* devices, int32_t n_devices, const apd_params* p, const apd_pair* pairs, int32_t n_pairs,
                          int32_t stride_bytes, int32_t xyz_off, int32_t label_off, int32_t n_streams, int32_t with_fitness,
                          apd_result* results);
APD_API int apd_align_batch(int device, const apd_params* p, const apd_pair* pairs, int32_t n_pairs,
                    int32_t stride_bytes, int32_t xyz_off, int32_t label_off,
                    int32_t n_streams, int32_t with_fitness, apd_result* results);

/* ---- one registration sharded over several GPUs (config C4; no reference
 * equivalent) -------------------------------------------------------------- */
/* One process and one handle per GPU; EVERY rank makes the same calls with the
 * same FULL source and target clouds. After apd_comm_init the handle splits the
 * work by interleaved chunks of the cell-sorted points (nranks x 16 equal chunks
 * of a multiple of 256 points; rank r owns chunks r, r + nranks, ...): each rank
 * searches the full grids for its chunks of the covariances (then
 * ncclAllGather), and runs update_correspondences / linearize / compute_error
 * over its chunks of the source in ONE launch each (then ncclAllReduce of 28 / 1
 * doubles on the handle's stream, or the in-kernel exchange below). The LM
 * loop runs redundantly on every rank from the identical reduced values, so all
 * ranks return the same pose; it equals the single-GPU result up to summation
 * order. cl_weight = 1 / (size of the whole source), as in the reference
 * (fast_apdgicp_impl.hpp:273). apd_get_correspondences / apd_get_mahalanobis
 * report the rank's own slice (-1 / 0 elsewhere). id128 is an ncclUniqueId
 * (128 bytes); n_source_total is ignored (kept for ABI v1). */
APD_API int apd_comm_unique_id(void* id128);
APD_API int apd_comm_init(apd_handle* h, const void* id128, int32_t rank, int32_t nranks,
                  int64_t n_source_total);
/* Optional, after apd_comm_init on every rank: fuse the all-reduce of H/b/err into
 * the reduction kernels. Each rank exports a small mailbox in its device memory
 * (handle64: a cudaIpcMemHandle_t, 64 bytes); the caller exchanges the handles
 * between the processes (any transport) and passes all of them, in rank order
 * (nranks x 64 bytes), to apd_comm_peer_attach. From then on the last block of
 * linearize / compute_error pushes its sums into every peer's mailbox over
 * NVLink, waits for the peers' and adds them in rank order: no separate
 * collective launch, bit-identical totals on all ranks. All ranks must attach
 * (or none). At most 16 ranks; GPUs must be peer-accessible (one NVSwitch node). */
APD_API int apd_comm_peer_handle(apd_handle* h, void* handle64);
APD_API int apd_comm_peer_attach(apd_handle* h, const void* handles);
APD_API int apd_comm_destroy(apd_handle* h);

/* The same sharding between handles of ONE process (a C++ caller that owns several GPUs, or several handles on one GPU):
 * no NCCL and no IPC — the n handles (created with apd_create on any devices of one NVSwitch node, 1 <= n <= 16) become
 * ranks 0..n-1; mailboxes and covariance arrays of the peers are plain pointers (peer access is enabled between the
 * devices). The exchange of the H/b/err sums runs inside the reduction kernels as with apd_comm_peer_attach; the
 * covariance chunks are pulled peer-to-peer. Every rank's calls meet inside the library, so each rank must be driven by
 * its own host thread: either the caller's (make the same apd_* calls on every handle, one thread per handle), or the
 * group's — the apd_group_* calls below fan one call out to all ranks and return rank 0's outputs (all ranks hold the
 * same). *_device variants take one device pointer per rank (the cloud as float4 {x,y,z,label} on that rank's device).
 * Destroy the group before its handles; the handles are plain unsharded handles again afterwards. */
typedef struct apd_group apd_group;
APD_API int apd_group_create(apd_handle* const* handles, int32_t n, apd_group** out);
APD_API int apd_group_destroy(apd_group* g);
APD_API int apd_group_size(const apd_group* g);
APD_API int apd_group_set_params(apd_group* g, const apd_params* p);
APD_API int apd_group_set_source(apd_group* g, const void* pts, int32_t n, int32_t stride_bytes, int32_t xyz_off, int32_t label_off);
APD_API int apd_group_set_target(apd_group* g, const void* pts, int32_t n, int32_t stride_bytes, int32_t xyz_off, int32_t label_off);
APD_API int apd_group_set_source_device(apd_group* g, const void* const* d_xyzl, int32_t n);
APD_API int apd_group_set_target_device(apd_group* g, const void* const* d_xyzl, int32_t n);
APD_API int apd_group_align(apd_group* g, const float* guess, float* T_out, double* T_out_f64, double* H_out, int32_t* converged,
                    int32_t* iterations);
APD_API int apd_group_linearize(apd_group* g, const double* T, double* H, double* b, double* err);
APD_API int apd_group_compute_error(apd_group* g, const double* T, double* err);

/* ---- instrumentation --------------------------------------------------- */
/* CUDA stream of the handle (cudaStream_t as void*), for event timing. */
APD_API void* apd_stream(apd_handle* h);
/* Kernel launches issued by this handle since creation (bench.py's
 * gpu_launches), and device milliseconds of the last call per kernel class
 * (CUDA events, only recorded when profiling is enabled). */
APD_API int64_t apd_launch_count(const apd_handle* h);
enum {
  APD_K_GRID = 0,      /* grid build (bounds, count, scan, scatter)          */
  APD_K_KNN_COV = 1,   /* exact kNN + covariance + regularisation            */
  APD_K_CORR = 2,      /* update_correspondences                             */
  APD_K_LINEARIZE = 3, /* linearize H/b/err reduction                        */
  APD_K_ERROR = 4,     /* compute_error reduction                            */
  APD_K_FITNESS = 5,
  APD_K_LM = 6,        /* device-resident optimizer loop (corr + linearize +
                          error trials + solve of ALL iterations, one launch)  */
  APD_K_COUNT = 7
};
APD_API int apd_set_profiling(apd_handle* h, int32_t enabled);
APD_API int apd_get_kernel_ms(apd_handle* h, double* ms /* [APD_K_COUNT] */, int64_t* launches /* [APD_K_COUNT] */);

#ifdef __cplusplus
}
#endif
#endif /* APDGICP_H_ */
