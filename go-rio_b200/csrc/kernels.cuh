// Host-callable launchers of the CUDA kernels (one translation unit per stage).
#pragma once
#include "common.cuh"

namespace apd {

// ---- grid.cu ---------------------------------------------------------------
// Device view of one cloud (all arrays in cell-sorted order, see common.cuh).
struct CloudDev {
  int n = 0;
  const float4* pts = nullptr;  // ORIGINAL order {x,y,z,label}
  GridDesc g{};
  int ncells = 0;
  uint32_t* cell_start = nullptr;  // [ncells+1]
  float4* spts = nullptr;          // [n] {x,y,z,bits(orig idx)}
  float* label = nullptr;          // [n]
  int* inv_perm = nullptr;         // [n] orig idx -> sorted pos
  double* cov = nullptr;           // [n*6]
  float* geo = nullptr;            // [n]
  double* geo64 = nullptr;         // [n] same value unrounded (read when the Mahalanobis storage is fp64)
};

// scratch for the grid build (sized by grid_work_bytes, carved by the caller)
struct GridWork {
  uint32_t* keys[2];
  uint32_t* vals[2];
  uint32_t* hist;       // [256 * sort_blocks]
  uint32_t* scan_tmp;   // block sums for the multi-level scan
  size_t scan_tmp_elems;
  unsigned int* ticket; // zero-initialised counter for the small-cloud scan (reset by the kernel)
  unsigned char* zero_flags = nullptr;  // optional [n]: cleared by the reorder step (Cloud::cov_flag of a target whose
                                        // covariances will be computed on demand — saves the memset call)
};
constexpr int kSmallCloud = 65536;  // at most this many points: rank-by-counting path (see grid.cu)
constexpr int kSortTile = 2048;  // keys per block in the radix sort
constexpr int kScanTile = 2048;  // elements per block in the scan
size_t scan_tmp_elems_for(size_t n);

// bounding box of a device-resident cloud, published by the kernel itself into pinned host memory: h_out = 6 ordered
// uints {minx,miny,minz,maxx,maxy,maxz} (the caller decodes them) followed by a 64-bit sequence number that becomes
// `seq` when the box is complete. d_state: 7 uints of device memory, prepared once by init_bounds_state.
void init_bounds_state(unsigned int* d_state, cudaStream_t s);
void launch_bounds(const float4* pts, int n, unsigned int* d_state, unsigned int* h_out, unsigned long long seq, cudaStream_t s,
                   int64_t* launches);
// keys/counts -> scan -> stable radix sort by cell -> reorder. Fills cell_start,
// spts, label, inv_perm of `c` (c.pts, c.n, c.g, c.ncells must be set).
void launch_grid_build(const CloudDev& c, const GridWork& w, cudaStream_t s, int64_t* launches);

// stable LSD radix sort of (key, value) pairs by the low `bits` bits of the key (8-bit digits); keys / vals: two buffers of
// n each (input in [0]); hist: 256 * ceil(n / kSortTile) uints; returns the index of the buffer that holds the result
int launch_sort_pairs_u32(uint32_t* const keys[2], uint32_t* const vals[2], int n, int bits, uint32_t* hist, uint32_t* scan_tmp, cudaStream_t s,
                          int64_t* launches);
void launch_exclusive_scan_u32(uint32_t* data, size_t n, uint32_t* tmp, cudaStream_t s, int64_t* launches);

// ---- prep_ops.cu: the stages on either side of the registration that use the same grid ------------
// radius search of every point of a cloud on its own grid (pcl radiusSearch semantics: fp32 d2 STRICTLY below r^2, the
// point itself included). counts: [n] at the ORIGINAL index. With offsets != nullptr (CSR, [n] at the original index) the
// neighbours' original ids are written to indices[offsets[i] ...] (in the grid's scan order).
void launch_radius_search(const CloudDev& c, float r2_fixed, int mode, double eps, int32_t* d_counts, const long long* d_offsets, int32_t* d_indices,
                          cudaStream_t s, int64_t* launches);
// pcl::transformPointCloud(cloud, out, Eigen::Matrix4d): double arithmetic, cast to float; the label is carried over
void launch_transform_cloud_d(const float4* in, int n, const double* T16_colmajor_host, float4* out, cudaStream_t s, int64_t* launches);
// voxel grid pieces (pcl::VoxelGrid, see apd_voxel_downsample)
void launch_voxel_bounds(const float4* pts, int n, unsigned int* d_state6, cudaStream_t s, int64_t* launches);
void launch_voxel_keys(const float4* pts, int n, float inv_leaf, const int min_b[3], const int mul[3], uint32_t* keys, uint32_t* vals, cudaStream_t s,
                       int64_t* launches);
void launch_voxel_heads(const uint32_t* keys, int n, uint32_t* heads, cudaStream_t s, int64_t* launches);
void launch_voxel_centroids(const float4* pts, const uint32_t* keys, const uint32_t* vals, const uint32_t* voxel_of, int n, float4* out, cudaStream_t s,
                            int64_t* launches);

// ---- knn_cov.cu --------------------------------------------------------------
// Exact kNN (ties by (d2, index)), in two steps: the search kernels write the neighbour lists d_nb[w*k + j]
// (original ids, ascending) of the sorted points, then launch_cov_regularize computes the covariance of the k
// neighbours (reference fast_apdgicp_impl.hpp:361-372), the regularisation (:374-405) and the geometric weight
// (:266-269), one thread per point. neighbors: optional int32[n*k] in ORIGINAL point order (parity hook).
// warp-per-point search (k <= 32):
// [w0, w0 + wn): the sorted positions to serve (wn < 0: to the end) — a rank of a sharded handle computes its slice only
void launch_knn_cov(const CloudDev& c, int k, int32_t* d_nb, int32_t* neighbors, cudaStream_t s, int64_t* launches, int w0 = 0, int wn = -1);
void launch_cov_regularize(const CloudDev& c, int k, int regularization, const int32_t* d_nb, cudaStream_t s, int64_t* launches, int w0 = 0,
                           int wn = -1);
// thread-per-point variant with the regularisation and geometric weight fused in (k <= 128);
// c.cov == nullptr: neighbours only.
void launch_knn_cov_fused(const CloudDev& c, int k, int regularization, int32_t* neighbors, cudaStream_t s, int64_t* launches, int w0 = 0,
                          int wn = -1);
void launch_regularize(const CloudDev& c, int regularization, cudaStream_t s, int64_t* launches);
// exact kNN (k <= 32) of nq arbitrary query points {x,y,z,-} on the cloud's grid: d_idx / d_d2 [q*k + j] = original ids and
// fp32 squared distances, ascending by (d2, index)
void launch_knn_query(const CloudDev& c, int k, const float4* d_queries, int nq, int32_t* d_idx, float* d_d2, cudaStream_t s, int64_t* launches);
// geo weight only (after set*Covariances)
void launch_geo_weight(const CloudDev& c, cudaStream_t s, int64_t* launches);
// covariance layout conversion for the getters/setters: sorted sym6 <-> 4x4 col-major in original order
void launch_cov_export(const CloudDev& c, double* d_out4x4, cudaStream_t s, int64_t* launches);
void launch_cov_import(const CloudDev& c, const double* d_in4x4, cudaStream_t s, int64_t* launches);

// ---- sharding of one registration over several GPUs ------------------------------
// The cell-sorted source points are cut into nranks * kShardSubMax equal chunks of `chunk` points (a multiple of 256, so
// every chunk starts on a tile and on a 16-byte boundary of every per-point array); rank r owns chunks r, r + nranks, ...
// (interleaved: the cost of a query varies across a scene). A kernel gets the rank's chunks as a table and serves all of
// them in ONE launch. Per-cloud arrays (spts, label, cov, geo) are indexed by the global sorted position begin_of(j) + o,
// per-linearisation arrays (corr, sqd, Mahalanobis) by the rank-local slot j * chunk + o. Unsharded: one chunk,
// begin 0, count n, chunk n — both indices are the sorted position.
constexpr int kShardSubMax = 16;  // chunks per rank: in cell order consecutive chunks are horizontal slabs of the scene whose
                                  // queries differ in cost; 4 per rank left the ranks ~8 % apart at the reduction's exchange
struct ShardTable {
  int nsub = 1;
  int chunk = 0;
  int first = 0;  // sorted position of the rank's first chunk (rank * chunk)
  int step = 0;   // distance between the rank's consecutive chunks (nranks * chunk)
  int n = 0;      // points of the cloud
  int plane = 0;  // local slots in all (= stride between the planes of the fp64 Mahalanobis storage)
  __host__ __device__ int slots() const { return nsub * chunk; }
  // begin / count of chunk j (arithmetic, not a table: a kernel parameter array indexed with a run-time value is copied
  // to local memory first — the linearize kernel lost 13 % to that)
  __host__ __device__ __forceinline__ int begin_of(int j) const { const int b = first + j * step; return b < n ? b : n; }
  __host__ __device__ __forceinline__ int count_of(int j) const { const int r = n - begin_of(j); return r < chunk ? r : chunk; }
};
inline ShardTable whole_cloud(int n) {
  ShardTable t;
  t.nsub = 1; t.chunk = n; t.n = n; t.plane = n;
  return t;
}

// ---- corr.cu -----------------------------------------------------------------
struct NoiseParams {
  double dist_var;      // distance_variance_
  double sin_az;        // sin(azimuth_variance_ / 180 * pi)
  double sin_el;        // sin(elevation_variance_ / 180 * pi)
  double thr_sq;        // corr_dist_threshold_^2 (double product)
  float search_limit;   // radius (m) beyond which no correspondence can pass the threshold; inf if none
  int gicp;             // 1: FastGICP's combined covariance C_B + T C_A T^T (fast_gicp_impl.hpp:157) — no noise term
};
struct CorrOut {
  int* corr;      // [n_src] sorted order of the source
  float* sqd;     // [n_src]
  float* second;  // [n_src] lower bound of the squared distance of every target point other than the match (single-lane search)
  int second_valid;  // `second` was written by the previous pass over the same clouds
  void* mahaA;    // float4[n] or double2[n] {xx,xy | ...} see common.cuh
  void* mahaB;    // float2[n] or double2[2n]
  int maha_fp64;
};
// reference update_correspondences (fast_apdgicp_impl.hpp:160-220) over the source points `sh` lists, in two launches:
// the search (:176-190: fp32 transform, exact 1-NN on the target grid, threshold; writes corr + sqd) and the per-point
// noise model / combined covariance / inverse (:194-218; reads the matched pairs coalesced, writes the Mahalanobis
// planes). T_prev != nullptr: `out` still holds the result of the previous pass over the same clouds at pose T_prev,
// which warm-starts the searches (identical results). lanes: lanes per 1-NN query (0: by cloud size).
// Returns the lanes per query it used (1: `out.second` now holds the bounds of this pass).
int launch_update_correspondences(const CloudDev& src, const CloudDev& tgt, const ShardTable& sh, const PoseD& T, const NoiseParams& np,
                                  const CorrOut& out, const PoseD* T_prev, int lanes, cudaStream_t s, int64_t* launches);
// getFitnessScore + inlier count: d_out = {sum d2 (double), n_in_range (as double), n_inliers (as double)}
void launch_fitness(const CloudDev& src, const CloudDev& tgt, const PoseF& T, double max_range, double inlier_sq_thr,
                    double* d_partials, int max_blocks, double* d_out3, unsigned int* d_ticket, cudaStream_t s,
                    int64_t* launches);
// the nearest target point of every source point under pose T, in the source's ORIGINAL order: the transformed point
// (fp32, pcl::transformPointCloud's arithmetic), the target point's original id (-1: empty target) and the fp32 squared
// distance — what pcl::Registration::getFitnessScore and the nodelet's inlier loop ask the search method for, one
// point at a time (scan_matching_odometry_nodelet.cpp:674-691)
void launch_source_nearest(const CloudDev& src, const CloudDev& tgt, const PoseF& T, int32_t* d_idx, float* d_d2, float* d_xyz, cudaStream_t s,
                           int64_t* launches);
// hooks: correspondences / mahalanobis back to original order and ids
void launch_corr_export(const CloudDev& src, const CloudDev& tgt, const ShardTable& sh, const CorrOut& c, int32_t* d_idx, float* d_sqd,
                        double* d_maha4x4, cudaStream_t s, int64_t* launches);
// transformed source cloud in original order (pcl::transformPointCloud)
void launch_transform_cloud(const float4* pts, int n, const PoseF& T, float* d_xyz, cudaStream_t s, int64_t* launches);

// ---- linearize.cu --------------------------------------------------------------
constexpr int kReduceVals = 28;  // 21 (upper H) + 6 (b) + 1 (err)
// Fused all-reduce of the 28 (1) sums across the ranks of a sharded registration, inside the reduction kernel's last
// block: every rank PUSHES its sums into a mailbox in every peer's memory (plain stores over NVLink peer mappings),
// publishes a sequence number with release semantics at system scope, waits for the numbers of all ranks in its own
// mailbox and adds the contributions in rank order — so all ranks hold bit-identical totals without a separate
// collective launch. Two buffers alternate by sequence parity: a rank can run at most one exchange ahead of a peer.
constexpr int kPeerMaxRanks = 16;
struct PeerMailbox {
  double val[2][kPeerMaxRanks][32];
  unsigned int flag[2][kPeerMaxRanks];
  unsigned int bar[kPeerMaxRanks];  // launch_peer_barrier: bar[r] = the last barrier number rank r has reached
  unsigned int timed_out;  // set if a peer did not show up within the time limit (the sums are then NaN)
};
struct PeerExchange {
  PeerMailbox* box[kPeerMaxRanks];  // box[r]: rank r's mailbox as mapped into this process (box[rank] = the local one)
  int rank = 0, nranks = 1;
  unsigned int seq = 0;             // sequence number of this exchange (0: no exchange in this launch)
};
struct ReduceWork {
  double* partials;     // [max_blocks * 28]
  unsigned int* ticket; // last-block-done counter (zeroed once; the kernel resets it)
  int max_blocks;
  PeerExchange xchg;    // seq != 0: exchange the totals with the peers in the kernel's tail
  // zero-copy result: the last block also writes the totals into pinned host memory (as the device sees it) and then
  // publishes host_seq there; the host polls that word instead of enqueueing a copy and waiting for the stream
  double* host_out = nullptr;               // [28]
  unsigned long long* host_seq_word = nullptr;
  unsigned long long host_seq = 0;
};
// reference linearize (:247-304) / compute_error (:313-343) given the stored
// correspondences and Mahalanobis matrices. out28 = 21 upper-triangular H
// entries (row-major, r<=c), 6 b, 1 err. cl_weight = 1 / n_total. One launch serves all the chunks `sh` lists.
void launch_linearize(const CloudDev& src, const CloudDev& tgt, const ShardTable& sh, const PoseD& T, const CorrOut& c, double n_total,
                      bool want_hb, const ReduceWork& w, double* d_out28, cudaStream_t s, int64_t* launches);
// barrier across the ranks of a sharded registration on their streams (one tiny kernel per rank: publish a sequence
// number in every peer's mailbox, wait for all of them) — orders the peer-to-peer copies of the covariance all-gather
void launch_peer_barrier(const PeerExchange& x, cudaStream_t s, int64_t* launches);

// load every kernel of a translation unit into the current context now instead of at its first launch
void preload_grid_kernels();
void preload_knn_kernels();
void preload_corr_kernels();
void preload_linearize_kernels();
void preload_lm_kernels();
void preload_prep_kernels();
void preload_vgicp_kernels();

// ---- vgicp.cu: FastVGICP (reference impl/fast_vgicp_impl.hpp, fast_vgicp_voxel.hpp) ---------------------------
constexpr int kVoxelDirect27 = 0, kVoxelDirect7 = 1, kVoxelDirect1 = 2;  // = APD_VOXEL_* of include/apdgicp.h
struct VoxelGridDesc {
  double res;  // voxel_resolution_
  int mn[3];   // smallest occupied voxel coordinate per axis
  int dim[3];  // extent of the occupied box in voxels (key = x + y * dim0 + z * dim0 * dim1 < 2^32 - 1)
};
struct VoxelMapDev {
  int n = 0;  // voxels
  VoxelGridDesc g{};
  const uint32_t* key = nullptr;  // ascending
  const int32_t* cnt = nullptr;   // num_points
  const double* mean = nullptr;   // 3 per voxel
  const double* cov = nullptr;    // 6 per voxel (xx, xy, xz, yy, yz, zz)
};
void launch_vgicp_keys(const float4* pts, int n, const VoxelGridDesc& vg, uint32_t* keys, uint32_t* vals, cudaStream_t s, int64_t* launches);
void launch_vgicp_voxels(const CloudDev& tgt, const uint32_t* keys, const uint32_t* vals, const uint32_t* voxel_of, int multiplicative,
                         uint32_t* vkey, int32_t* vcnt, double* vmean, double* vcov, cudaStream_t s, int64_t* launches);
void launch_vgicp_correspondences(const CloudDev& src, const VoxelMapDev& vm, int method, int n_off, const PoseD& T, int32_t* vcorr, double* vmaha,
                                  cudaStream_t s, int64_t* launches);
void launch_vgicp_reduce(const CloudDev& src, const VoxelMapDev& vm, const int32_t* vcorr, const double* vmaha, int n_off, const PoseD& T, bool want_hb,
                         double* partials, int max_blocks, double* d_out28, unsigned int* ticket, cudaStream_t s, int64_t* launches);
void launch_vgicp_export_voxels(const VoxelMapDev& vm, int32_t* d_coords, double* d_covs, cudaStream_t s, int64_t* launches);
void launch_vgicp_export_maha(const int32_t* vcorr, const double* vmaha, long long slots, double* d_out, cudaStream_t s, int64_t* launches);

// an empty kernel (launch-rate diagnostic)
void launch_noop(cudaStream_t s, int64_t* launches);

// ---- lm.cu ---------------------------------------------------------------------
// The device-resident optimizer loop (LsqRegistration::computeTransformation,
// reference lsq_registration_impl.hpp:55-173) of one registration per cluster.
constexpr int kLmTraceRows = 640;  // max_iterations (64) x lm_max_iterations (10) rows of 8 doubles
constexpr int kLmTraceHead = 24;   // rows copied back together with the result header
struct LmResult {
  double pose[12];     // r[9] row-major, t[3]: the fp64 x0 after the last accepted step
  double H[36];        // final_hessian_
  double lm_lambda;
  double fitness[3];   // sum d2, n in range, n inliers (only with want_fitness)
  int converged, nr_iterations, lm_failed, n_trace;
  int hessian_set, pad_;  // hessian_set: H holds a final_hessian_ (some step was accepted)
  unsigned long long t_begin, t_end;  // %globaltimer (ns) when the kernel's first CTA started / just before it published
  // fused prologue (LmJob::prep): the grids the kernel chose, for the host's later calls on the same clouds
  GridDesc grid[2];    // {source, target}
  int ncells[2];
  int prep_status;     // 0 ok; 1 a grid did not fit its arrays: nothing was registered, the host takes the unfused path
  int pad2_;
  // where the kernel's time went, by phase, as CTA 0 saw it (ns; only filled by a library built with -DAPD_LM_PHASE_TIMING):
  // 0 boxes + grids, 1 source covariances, 2 1-NN searches, 3 on-demand kNN searches, 4 on-demand covariances,
  // 5 Mahalanobis matrices, 6 H/b/err sums + reduction + solve, 7 LM trials, 8 fitness tail, 9 (spare)
  unsigned long long phase_ns[10];
  unsigned long long seq; // host copy only: LmJob::seq once the header + first trace rows have arrived (written last)
  double trace[kLmTraceRows * 8];
};
struct LmJob {
  const float4* s_spts; const float* s_label; const double* s_cov; const float* s_geo; const double* s_geo64;
  const float4* t_spts; const float* t_label; const double* t_cov; const uint32_t* t_cell_start;
  GridDesc tg;
  int n_src;
  int* corr; float* sqd; void* mahaA; void* mahaB;
  double cl_w;         // 1 / correspondences_.size() (:273)
  double guess[12];    // r[9] row-major, t[3]
  LmResult* result;    // device memory
  // Zero-copy publication: the kernel's tail writes the result header + the first kLmTraceHead trace rows into pinned
  // host memory and then host_result->seq = seq; the host polls that word — no D2H copy, no stream query per registration
  LmResult* host_result;  // pinned host memory as seen from the device (nullptr: the host copies `result` itself)
  unsigned long long seq;
  // Target covariances on demand (t_cov_flag != nullptr): only the target points that become correspondences ever need
  // one (a scan matches ~2 k of a 60 k-point submap), so the loop computes the covariance of a matched target point the
  // first time it meets it — same kNN search, same arithmetic, same bits as the per-cloud kernels (knn_warp.cuh).
  unsigned char* t_cov_flag;  // [n_tgt] 1: t_cov_rw[pos] is valid. nullptr: all target covariances are (eager)
  double* t_cov_rw;           // = t_cov, writable
  const float4* t_pts;        // target in ORIGINAL order (the covariance gathers its neighbours there)
  int32_t* nb;                // [n_src][k] + [n_src] scratch: neighbour ids between the search pass and the covariance pass, work list
  int k, reg;                 // k_correspondences_, regularization_method_
  // Fused prologue (prep != 0; prep.cuh): the kernel is handed the RAW clouds and builds what the loop needs itself —
  // bit 0: the source's grid + covariances, bit 1: the target's grid (its covariances come on demand). The arrays above
  // are then outputs of the prologue as well; tg is an output when bit 1 is set.
  int prep;
  int n_tgt;
  const float4* s_pts;        // source in ORIGINAL order (t_pts: the target)
  int* s_inv_perm; int* t_inv_perm;
  uint32_t* s_cell_start;     // [s_cell_cap]; t_cell_start: [t_cell_cap]
  int s_cell_cap, t_cell_cap;
  uint32_t* s_scratch;        // [3 n_src]: keys, ranks, tmp of the grid build; t_scratch: [3 n_tgt]
  uint32_t* t_scratch;
  double s_cells_per_point, t_cells_per_point;
  int s_k, s_reg;             // the source's covariance parameters (the target's: k, reg)
  int gicp;                   // FastGICP: unit weights — the geometric weights are written as zeros
  GridDesc sg;                // the source's grid when the prologue does not build it (prep bit 0 clear), for the result
  int s_ncells, t_ncells;     // ... and the cell counts of grids that are not rebuilt
};
struct LmConfig {
  int max_iterations, optimizer, lm_max_iterations, maha_fp64, want_fitness;
  double rotation_epsilon, transformation_epsilon, lm_init_lambda_factor;
  double fitness_max_range, inlier_sq_thr;
  NoiseParams np;
};
constexpr int kLmMaxSource = 32768;  // larger source clouds use the streaming kernels + host loop
// one: job passed by value (d_jobs == nullptr, n_jobs == 1); d_jobs: device array of n_jobs jobs.
// cluster: CTAs per registration (1, 2, 4 or 8).
// min_blocks: 1 = the 128-register build of the kernel (one CTA per SM: a lone registration), 2 = the 64-register build
// (two CTAs per SM: the workers of a batch pool)
// two: a second job passed by value next to `one` (n_jobs becomes 2): two pooled registrations in one launch.
void launch_lm(const LmJob* one, const LmJob* d_jobs, int n_jobs, const LmConfig& cfg, int cluster, int min_blocks, cudaStream_t s,
               int64_t* launches, const LmJob* two = nullptr);

}  // namespace apd
