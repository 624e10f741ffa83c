// C-ABI of the B200 FastAPDGICP library (include/apdgicp.h) and the host side
// of the registration: cloud staging, grid sizing, the LM / GN outer loop
// (reference lsq_registration_impl.hpp:55-173) driving the CUDA kernels, the
// batched and the source-sharded (NCCL) variants.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <sched.h>
#include <emmintrin.h>

#include <atomic>
#include <memory>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstddef>
#include <cctype>
#include <cstring>
#include <functional>
#include <limits>
#include <string>
#include <thread>
#include <vector>

#include "../../include/apdgicp.h"
#include "host_math.hpp"
#include "host_stage.hpp"
#include "kernels.cuh"

using namespace apd;

namespace {

// ------------------------------------------------------------------ NCCL ----
// Loaded lazily with dlopen so that the library has no link-time dependency on
// NCCL (the single-GPU drop-in use case never needs it).
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool load(std::string& err) {
    if (lib) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
      lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (lib) break;
    }
    if (!lib) {
      err = std::string("cannot dlopen libnccl.so.2: ") + dlerror();
      return false;
    }
    GetUniqueId = (decltype(GetUniqueId))dlsym(lib, "ncclGetUniqueId");
    CommInitRank = (decltype(CommInitRank))dlsym(lib, "ncclCommInitRank");
    CommDestroy = (decltype(CommDestroy))dlsym(lib, "ncclCommDestroy");
    AllReduce = (decltype(AllReduce))dlsym(lib, "ncclAllReduce");
    AllGather = (decltype(AllGather))dlsym(lib, "ncclAllGather");
    GetErrorString = (decltype(GetErrorString))dlsym(lib, "ncclGetErrorString");
    if (!GetUniqueId || !CommInitRank || !CommDestroy || !AllReduce || !AllGather) {
      err = "libnccl is missing required symbols";
      return false;
    }
    return true;
  }
};
NcclApi g_nccl;
constexpr int kNcclFloat64 = 8;  // ncclDouble
constexpr int kNcclSum = 0;      // ncclSum

// -------------------------------------------------------------- buffers ----
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 4 + 256;  // grow with slack so repeated set_* calls rarely reallocate
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) return e;
    cap = want;
    return cudaSuccess;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <typename T>
  T* as() const { return reinterpret_cast<T*>(p); }
};

struct PinnedBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMallocHost(&p, want);
    if (e != cudaSuccess) return e;
    cap = want;
    return cudaSuccess;
  }
  void release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
  }
};

struct Cloud {
  int n = 0;
  bool present = false;
  uint64_t key = 0;
  uint64_t print = 0;  // content fingerprint taken when `key` was accepted (see cloud_fingerprint)
  float bbox[6] = {0, 0, 0, 0, 0, 0};
  DevBuf pts, spts, label, inv_perm, cov, geo, geo64, cell_start;
  PinnedBuf stage;                  // host AoS -> float4 staging for the H2D copy
  cudaEvent_t staged = nullptr;     // completion of the last H2D out of `stage`
  unsigned long long staged_seq = 0;  // ... or: any kernel of the handle with a LATER sequence number has been seen to finish
  bool staged_by_seq = false;
  const float4* ext_pts = nullptr;  // user-owned device cloud (apd_set_*_device)
  bool bbox_pending = false;        // ... whose bounding box is not known on the host yet (see finish_bboxes)
  bool bbox_known = false;          // bbox[] holds this cloud's box (staging pass, or a finished device reduction)
  bool bbox_launched = false;       // ... the kernel that reduces it has been launched (it is only launched when a grid is sized on the host:
                                    // the fused registration kernel sizes its grids itself)
  const void* host_packed = nullptr;  // packed float4 cloud in the caller's page-locked memory whose box has not been computed (same reason)
  unsigned long long bbox_seq = 0;  // ... and the sequence number that announces it
  GridDesc g{};
  int ncells = 0;
  bool grid_valid = false;
  bool grid_ordered = true;  // points of a cell lie in ascending original index (always, except a target the fused kernel placed
                             // by atomic ranks: prep.cuh). Only a SOURCE needs it: swapping such a target in rebuilds its grid.
  bool cov_valid = false;
  bool geo_valid = false;
  int geo_variant = APD_VARIANT_APDGICP;  // what geo / geo64 hold: sigma3/sigma1 (APDGICP) or zeros (GICP: unit weights)
  // Covariances on demand (target of the device-resident loop): cov[w] is valid where cov_flag[w] == 1; computed with
  // lazy_k / lazy_reg, the parameters in force when the cloud's first covariance was asked for (the reference computes
  // all target covariances at that moment and keeps them across later parameter changes)
  DevBuf cov_flag;
  bool cov_lazy = false;
  bool flags_fresh = false;  // the grid build that just ran cleared cov_flag
  int lazy_k = 0, lazy_reg = 0;
  void drop_derived() { grid_valid = cov_valid = geo_valid = cov_lazy = bbox_pending = bbox_launched = false; host_packed = nullptr; }
  CloudDev view() const {
    CloudDev c;
    c.n = n;
    c.pts = ext_pts ? ext_pts : pts.as<const float4>();
    c.g = g;
    c.ncells = ncells;
    c.cell_start = cell_start.as<uint32_t>();
    c.spts = spts.as<float4>();
    c.label = label.as<float>();
    c.inv_perm = inv_perm.as<int>();
    c.cov = cov.as<double>();
    c.geo = geo.as<float>();
    c.geo64 = geo64.as<double>();
    return c;
  }
  void release() {
    pts.release(); spts.release(); label.release(); inv_perm.release(); cov.release(); geo.release(); geo64.release(); cell_start.release();
    cov_flag.release();
    stage.release();
    if (staged) cudaEventDestroy(staged);
    staged = nullptr;
  }
};

struct ProfEvent {
  cudaEvent_t a, b;
  int cls;
  int64_t launches;
};

}  // namespace

struct apd_group;

struct apd_handle {
  int device = 0;
  cudaStream_t stream = nullptr;
  apd_params params{};
  std::string error;
  Cloud src, tgt;
  // per-linearisation state (sorted order of the source)
  // FastVGICP (variant = APD_VARIANT_VGICP, vgicp.cu): the target's Gaussian voxel map and the correspondence table
  DevBuf vkey, vcnt, vmean, vcov, vcorr, vmaha;
  bool vox_valid = false;     // voxelmap_ != nullptr
  int n_vox = 0;
  VoxelGridDesc vgd{};
  int vcorr_n = 0, vcorr_noff = 0;  // shape of the table the last update_correspondences wrote
  DevBuf corr, sqd, second, mahaA, mahaB;
  bool second_valid = false;  // `second` holds the bounds of the pass the warm start refers to (single-lane search)
  int corr_n = -1;          // number of source points the buffers describe (-1: none yet)
  // warm start of update_correspondences: the buffers hold a pass over the CURRENT clouds at corr_pose with corr_thr
  bool corr_warm = false;
  PoseD corr_pose{};
  double corr_thr = 0.0;
  int corr_fp64 = 0;
  // scratch
  DevBuf work, scratch, partials, small;  // small: out28 + ticket + fitness
  DevBuf nbuf;  // neighbour lists (n x k original ids) between the kNN search and the covariance kernel
  PinnedBuf h_small;
  PinnedBuf h_query;  // staging of apd_nearest_k's query points
  DevBuf lm_result;   // LmResult of the device-resident optimizer loop
  PinnedBuf h_lm;     // its header + first trace rows on the host
  // Zero-copy results: kernels publish small results (the loop's result header, a device cloud's bounding box) straight
  // into pinned host memory followed by a sequence number; the host polls that word instead of the stream. Per
  // registration this saves the D2H copies, the memsets of the box reduction and every cudaStreamQuery of the wait: a
  // batch pool of 32-64 host threads is bound by the rate of driver calls. APD_ZERO_COPY=0: copies + stream waits.
  bool zero_copy = true;
  unsigned long long reduce_seq = 0;  // last sequence number a reduction kernel was asked to publish (h_small[kReduceSeqSlot])
  LmResult* h_lm_dev = nullptr;         // h_lm as the device sees it
  unsigned int* h_small_dev = nullptr;  // h_small as the device sees it
  unsigned long long seq = 0;           // last sequence number handed to a kernel
  unsigned long long seq_seen = 0;      // highest sequence number the host has seen published (stream order: everything
                                        // enqueued before that kernel has completed too)
  // APD_BATCH_TRACE=1: host wall time per phase of a pooled registration, summed (printed by apd_batch_destroy):
  // 0 set clouds (staging / H2D enqueue), 1 wait for the bounding boxes, 2 enqueue grids + covariances, 3 enqueue the loop,
  // 4 wait for the result
  double phase_s[5] = {0, 0, 0, 0, 0};
  int64_t phase_n = 0;
  double lm_phase_s[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};  // ... by phase (LmResult::phase_ns; diagnostic builds only)
  double lm_kernel_s = 0.0;  // device time of the loop kernels (their own %globaltimer stamps: valid under load, no events needed)
  // CTAs per registration in the device loop (APD_LM_CLUSTER=1|2|4|8|16). A lone registration is latency-bound: 8 CTAs
  // (C2 on B200: lm_kernel 0.42 ms with 4, 0.29 ms with 8); the workers of a batch pool share the SMs: 4.
  int lm_cluster = 8;
  // which build of the loop kernel: 1 = 128 registers, one CTA per SM (shortest lone registration); 2 = 64 registers, two
  // CTAs per SM (a batch pool: +27 % registrations/s). Same arithmetic either way. APD_LM_MINB=1|2.
  int lm_min_blocks = 1;
  bool lm_failed = false;
  // The fused registration kernel (lm.cu + prep.cuh): bounding boxes, grid sizing, both grids and the source covariances
  // inside the loop kernel's launch — one launch per registration, no wait for the boxes. APD_FUSED=0: the separate kernels.
  bool fused = true;
  int fused_inflight = 0;       // prep bits of the loop launch in flight (0: unfused)
  int fused_max_target = 131072;  // larger targets are built by the GPU-wide kernels (the prologue runs on the loop's few SMs)
  // wait for the result of the device loop on a blocking-sync event instead of spinning on the stream: the batch
  // context runs more host threads than it needs cores for (set by apd_batch_create; APD_BLOCKING_SYNC=0|1 overrides)
  bool blocking_wait = false;
  bool pooled = false;  // a worker of a batch pool: throughput matters, not the latency of one registration
  int poll_wait_us = 0;
  cudaEvent_t done_ev = nullptr;
  // Two pooled registrations per launch (batch_worker / PairSpot): a registration whose loop is ready to launch waits
  // (at most pair_hold_us) in its device's spot for the next one that becomes ready, and that one's worker launches both.
  // pair_state: 0 none, 1 waiting in the spot, 2 taken by a partner (launch in progress), 3 launched on launch_stream.
  std::atomic<int> pair_state{0};
  cudaEvent_t pair_ev = nullptr;          // what this handle's stream holds before the loop (cloud copies), for the partner's stream
  cudaStream_t launch_stream = nullptr;   // the stream the loop in flight was launched on (this handle's own, or the partner's)
  LmJob staged_job{};                     // the loop as prepared for launch
  LmConfig staged_cfg{};
  // batch workers ask the device loop to append the getFitnessScore pass (saves a launch and a round trip per pair)
  bool fuse_fitness = false;
  double fuse_inlier_sq_thr = 0.25;
  bool fit_valid = false;  // fit[] describes the last align's final pose
  bool pending_fitness = false;  // the loop in flight carries the fitness pass
  double fit[3] = {0, 0, 0};
  // results of the last align
  hm::Pose final_pose = hm::Pose::identity();
  float final_T[16];        // column-major
  double final_H[36];       // row-major == column-major (symmetric)
  bool converged = false;
  int nr_iterations = 0;
  double lm_lambda = -1.0;
  std::vector<double> trace;  // rows of 8
  // sharding
  ncclComm_t comm = nullptr;
  int comm_rank = 0, comm_size = 1;
  // fused all-reduce over NVLink peer memory (apd_comm_peer_*): this rank's mailbox + the peers' mailboxes as mapped here
  PeerMailbox* mailbox = nullptr;
  PeerMailbox* peer_box[kPeerMaxRanks] = {nullptr};
  bool peers_attached = false;
  unsigned int xchg_seq = 0;
  // instrumentation
  int64_t launches = 0;
  bool profiling = false;
  std::vector<ProfEvent> pending;
  std::vector<cudaEvent_t> event_pool;
  double k_ms[APD_K_COUNT] = {0};
  int64_t k_launches[APD_K_COUNT] = {0};
  int max_reduce_blocks = 148 * 4;
  double cells_per_point = 0.0;  // 0: by size (4 below 500 k points: fewer kNN shells; 8 above: cheaper 1-NN per LM iteration)
  // Scans (<= small_cloud_n points): what their grid is searched for most is their own k = 20 neighbours, and a 1-2 k-point
  // radar scan is so sparse that at 4 cells per point the 3x3x3 cube holds ~7 of them: every query walks several shells.
  // APD_CELLS_PER_POINT_SMALL overrides.
  double cells_per_point_small = 1.0;  // (measured on the C3 pool: 4 -> 29.8 k, 2 -> 30.5 k, 1 -> 31.2 k, 0.5 -> 31.5 k registrations/s)
  int small_cloud_n = 8192;
  // Submaps (small_cloud_n < n < 500 k points): 4 cells per point when every point's k neighbours are searched (round 1, one
  // B200: 0.5 / 1 / 2 / 4 / 8 -> 12.4 / 15.8 / 18.9 / 20.7 / 18.4 k registrations/s). A target whose covariances are
  // computed on demand is searched ~15 k times, not 60 k + 11 k: half the cells — half the cell_start array every resident
  // registration drags through L2, fewer rows per shell — is worth more than the shorter candidate lists (C3 probe,
  // profiles/r02_run46.sh: 1 / 1.5 / 2 / 2.5 / 3 / 4 -> 37.5 / 39.3 / 40.0 / 39.4 / 38.9 / 38.3 k registrations/s).
  // APD_CELLS_PER_POINT_MID / APD_CELLS_PER_POINT_MID_LAZY override. The resolution changes no result (exact searches).
  double cells_per_point_mid = 4.0, cells_per_point_mid_lazy = 2.0;
  double cells_for(int n, bool lazy_cov = false) const {
    if (cells_per_point > 0.0) return cells_per_point;
    if (n >= 500000) return 8.0;
    return n <= small_cloud_n ? cells_per_point_small : (lazy_cov ? cells_per_point_mid_lazy : cells_per_point_mid);
  }
  double max_cells_for(int n) const { return std::max(cells_for(n, false), cells_for(n, true)); }  // (what the arrays are sized for)
  // kNN kernel choice: 0 auto (warp-per-point below knn_thread_min_n points and k <= 32, thread-per-point
  // above: the warp kernel has the shorter critical path, the thread kernel the higher throughput),
  // 1 warp, 2 thread. Override with APD_KNN_MODE=warp|thread.
  int knn_mode = 0;
  int knn_thread_min_n = 500000;
  // Target covariances on demand inside the device-resident loop (see Cloud::cov_lazy): 0 never, 1 whenever the loop
  // runs, -1 auto: in the workers of a batch pool when the target has at least lazy_min_ratio x the source's points (a
  // scan against a submap matches a few percent of the submap: 11.2 k against 7.0 k registrations/s). A lone handle
  // computes all target covariances in one GPU-wide pass: the on-demand searches run on the loop's few SMs and lengthen
  // a single registration (C2: 0.67 ms against 0.64 ms). APD_LAZY_TARGET_COV=0|1|auto.
  int lazy_mode = -1;
  int lazy_min_ratio = 4;
  int knn_max_k() const { return knn_mode == 1 ? 32 : 128; }
  bool knn_use_warp(int n, int k) const { return knn_mode == 1 || (knn_mode == 0 && k <= 32 && n < knn_thread_min_n); }
  // Sharded handle (apd_comm_init across processes, apd_group_create inside one): every rank holds both FULL clouds and
  // grids. The cell-sorted points of a cloud are cut into nranks * kShardSub equal chunks (a multiple of 256 points: whole
  // tiles, 16-byte aligned in every array); rank r owns chunks r, r + nranks, r + 2 nranks, ... — interleaved, because
  // the cost of a 1-NN / kNN query varies across the scene (contiguous halves of one 20 M-point scene differed by 1.5x).
  // Chunk j of every rank forms one contiguous group, so the covariance all-gather runs in place, group by group.
  static constexpr int kShardSub = kShardSubMax;
  apd_group* group = nullptr;  // in-process ranks (apd_group_create): peers are plain pointers, no NCCL
  unsigned int bar_seq = 0;    // last peer barrier this rank has entered
  bool sharded() const { return comm != nullptr || group != nullptr; }
  int shard_subs() const { return sharded() ? kShardSub : 1; }
  int shard_chunk(int n) const {
    if (!sharded()) return n;
    const int c = (n + comm_size * kShardSub - 1) / (comm_size * kShardSub);
    return (c + 255) / 256 * 256;
  }
  int sub_begin(int n, int j) const { return sharded() ? (int)std::min<long long>(n, (long long)(j * comm_size + comm_rank) * shard_chunk(n)) : 0; }
  int sub_count(int n, int j) const {
    return sharded() ? (int)(std::min<long long>(n, (long long)(j * comm_size + comm_rank + 1) * shard_chunk(n)) - sub_begin(n, j)) : n;
  }
  size_t local_n(int n) const { return sharded() ? (size_t)shard_chunk(n) * kShardSub : (size_t)n; }
  size_t padded_n(int n) const { return sharded() ? (size_t)shard_chunk(n) * comm_size * kShardSub : (size_t)n; }
  // the chunks of a cloud of n points this rank serves (unsharded: the whole cloud)
  ShardTable shard_table(int n) const {
    if (!sharded()) return whole_cloud(n);
    ShardTable t;
    t.nsub = kShardSub;
    t.chunk = shard_chunk(n);
    t.first = comm_rank * t.chunk;
    t.step = comm_size * t.chunk;
    t.n = n;
    t.plane = (int)local_n(n);
    return t;
  }
  // lanes per 1-NN query of update_correspondences (0: by cloud size). APD_CORR_LANES=1|2|4|8; APD_CORR_MODE=wide|lane (8 | 1)
  int corr_lanes = 0;
};

// In-process ranks of one sharded registration (apd_group_create): the handles of ONE process, on one device or on
// several, each driven by its own host thread (the group's, for the apd_group_* calls). Peers are plain pointers:
// mailboxes for the in-kernel exchange of the H/b/err sums and for the stream barrier, and the peers' covariance
// arrays, which a rank PULLS its missing chunks from (peer-to-peer copies over NVLink when the devices differ).
struct apd_group {
  std::vector<apd_handle*> ranks;
  // barrier of the rank threads on the host (the all-gather publishes buffer pointers across it)
  std::mutex mu;
  std::condition_variable cv;
  int arrived = 0;
  uint64_t gen = 0;
  void* gather_ptr[apd::kPeerMaxRanks] = {nullptr};
  // the group's own threads (one per rank) for the fan-out calls
  std::vector<std::thread> threads;
  std::condition_variable cv_job, cv_done;
  uint64_t job_gen = 0;
  int job_pending = 0;
  bool stop = false;
  std::function<int(apd_handle*, int)> job;
  std::vector<int> job_rc;
  // waits until every rank thread has arrived; false after 60 s (a rank failed before the meeting point)
  bool host_barrier() {
    std::unique_lock<std::mutex> lk(mu);
    const uint64_t my = gen;
    if (++arrived == (int)ranks.size()) {
      arrived = 0;
      gen++;
      cv.notify_all();
      return true;
    }
    return cv.wait_for(lk, std::chrono::seconds(60), [&] { return gen != my; });
  }
};

namespace {

#define APD_CUDA(h, expr)                                                                 \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      (h)->error = std::string(#expr) + ": " + cudaGetErrorString(_e);                    \
      return APD_ERR_CUDA;                                                                \
    }                                                                                     \
  } while (0)

// Wait for everything enqueued on the handle's stream (the paths that do not publish into pinned memory: host-driven
// loop, parity hooks, APD_ZERO_COPY=0). A lone handle spins (lowest latency); handles of a batch pool sleep and poll the
// stream every poll_wait_us — unlike blocking-sync events that costs no interrupt per wait (8 x B200, 256 waiting
// threads: 46.1 k registrations/s with polling against 22.6 k with blocking-sync events).
cudaError_t wait_stream(apd_handle* h) {
  if (!h->blocking_wait) return cudaStreamSynchronize(h->stream);
  if (h->poll_wait_us > 0) {  // sleep-and-poll: no interrupt per wait (APD_POLL_WAIT_US)
    for (;;) {
      const cudaError_t e = cudaStreamQuery(h->stream);
      if (e != cudaErrorNotReady) return e;
      std::this_thread::sleep_for(std::chrono::microseconds(h->poll_wait_us));
    }
  }
  if (!h->done_ev) {
    cudaError_t e = cudaEventCreateWithFlags(&h->done_ev, cudaEventBlockingSync | cudaEventDisableTiming);
    if (e != cudaSuccess) return e;
  }
  cudaError_t e = cudaEventRecord(h->done_ev, h->stream);
  if (e != cudaSuccess) return e;
  return cudaEventSynchronize(h->done_ev);
}

// Wait until a kernel has published `expect` at *word (pinned host memory, see apd_handle::zero_copy). A lone handle
// spins; pool workers sleep between looks. The stream is queried now and then so that a failed launch / faulting kernel
// ends the wait with its error instead of hanging.
cudaError_t wait_host_seq(apd_handle* h, const volatile unsigned long long* word, unsigned long long expect) {
  const bool sleepy = h->blocking_wait && h->poll_wait_us > 0;
  const auto t0 = std::chrono::steady_clock::now();
  auto next_check = t0 + std::chrono::milliseconds(20);
  for (unsigned spins = 0;; spins++) {
    if (*word == expect) break;
    if (sleepy) std::this_thread::sleep_for(std::chrono::microseconds(h->poll_wait_us));
    else if ((spins & 63) == 63) std::this_thread::yield();
    if (sleepy || (spins & 1023) == 1023) {
      const auto now = std::chrono::steady_clock::now();
      if (now >= next_check) {
        const cudaError_t e = cudaStreamQuery(h->stream);
        if (e != cudaErrorNotReady && e != cudaSuccess) return e;
        if (e == cudaSuccess && *word != expect) return cudaErrorUnknown;  // the stream is idle and nothing was published
        next_check = now + std::chrono::milliseconds(20);
      }
    }
  }
  std::atomic_thread_fence(std::memory_order_acquire);
  return cudaSuccess;
}

int fail(apd_handle* h, int code, const char* msg) {
  h->error = msg;
  return code;
}

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
    else prev = -1;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

// --- profiling: CUDA events around each kernel class --------------------------
struct ProfScope {
  apd_handle* h;
  int cls;
  cudaEvent_t a = nullptr, b = nullptr;
  int64_t l0;
  ProfScope(apd_handle* h_, int cls_) : h(h_), cls(cls_), l0(h_->launches) {
    if (!h->profiling) return;
    auto get = [&]() {
      cudaEvent_t e;
      if (!h->event_pool.empty()) {
        e = h->event_pool.back();
        h->event_pool.pop_back();
      } else {
        cudaEventCreate(&e);
      }
      return e;
    };
    a = get();
    b = get();
    cudaEventRecord(a, h->stream);
  }
  ~ProfScope() {
    if (!h->profiling) {
      h->k_launches[cls] += h->launches - l0;
      return;
    }
    cudaEventRecord(b, h->stream);
    h->pending.push_back(ProfEvent{a, b, cls, h->launches - l0});
  }
};
void flush_prof(apd_handle* h) {
  if (h->pending.empty()) return;
  cudaStreamSynchronize(h->stream);
  for (auto& p : h->pending) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, p.a, p.b);
    h->k_ms[p.cls] += ms;
    h->k_launches[p.cls] += p.launches;
    h->event_pool.push_back(p.a);
    h->event_pool.push_back(p.b);
  }
  h->pending.clear();
}

PoseD to_pose_d(const hm::Pose& p) {
  PoseD T;
  for (int r = 0; r < 3; r++) {
    for (int c = 0; c < 3; c++) T.r[r * 3 + c] = p(r, c);
    T.t[r] = p(r, 3);
  }
  return T;
}
PoseF colmajor_f32_to_pose_f(const float* T) {
  PoseF f;
  for (int r = 0; r < 3; r++) {
    for (int c = 0; c < 3; c++) f.r[r * 3 + c] = T[c * 4 + r];
    f.t[r] = T[3 * 4 + r];
  }
  return f;
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int ensure_small(apd_handle* h);
void ensure_work_queues();
NoiseParams noise_params(const apd_params& p);
int ensure_corr_buffers(apd_handle* h);

// The bounding boxes of device clouds are reduced on the device (set_cloud_device) and read back here, the first time a
// grid needs one: ONE wait serves both clouds of a {setInputTarget; setInputSource; align} sequence.
int launch_bounds_of(apd_handle* h, Cloud& c);
int finish_bboxes(apd_handle* h) {
  for (Cloud* c : {&h->src, &h->tgt})
    if (c->host_packed) {  // a packed host cloud whose box was left for later (the caller's buffer is still valid: same call)
      bounds_of_packed(c->host_packed, c->n, c->bbox);
      c->host_packed = nullptr;
      c->bbox_known = true;
    }
  if (!h->src.bbox_pending && !h->tgt.bbox_pending) return APD_OK;
  for (Cloud* c : {&h->src, &h->tgt})
    if (c->bbox_pending && !c->bbox_launched) {
      const int rc = launch_bounds_of(h, *c);
      if (rc != APD_OK) return rc;
    }
  const auto t_wait0 = std::chrono::steady_clock::now();
  struct Acc {
    apd_handle* h;
    std::chrono::steady_clock::time_point t0;
    ~Acc() { h->phase_s[1] += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }
  } acc{h, t_wait0};
  if (!h->zero_copy) APD_CUDA(h, wait_stream(h));
  for (Cloud* c : {&h->src, &h->tgt}) {
    if (!c->bbox_pending) continue;
    const unsigned int* enc = reinterpret_cast<const unsigned int*>(reinterpret_cast<double*>(h->h_small.p) + (c == &h->src ? 48 : 52));
    if (h->zero_copy) {
      APD_CUDA(h, wait_host_seq(h, reinterpret_cast<const volatile unsigned long long*>(enc + 6), c->bbox_seq));
      h->seq_seen = std::max(h->seq_seen, c->bbox_seq);
    }
    for (int a = 0; a < 6; a++) {
      unsigned int u = enc[a];
      u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
      std::memcpy(&c->bbox[a], &u, 4);
    }
    c->bbox_pending = false;
    c->bbox_launched = false;
    c->bbox_known = true;
  }
  return APD_OK;
}

int ensure_grid(apd_handle* h, Cloud& c, bool clear_flags = false) {
  if (c.grid_valid) return APD_OK;
  if (!c.present || c.n <= 0) return fail(h, APD_ERR_INVALID, "cloud not set");
  {
    if (!c.bbox_known && !c.host_packed) c.bbox_pending = true;  // (e.g. a cloud whose grid the fused kernel had sized itself)
    int rc = finish_bboxes(h);
    if (rc != APD_OK) return rc;
  }
  size_grid(c.bbox, c.n, h->cells_for(c.n), c.g, c.ncells);
  const size_t n = (size_t)c.n;
  APD_CUDA(h, c.spts.ensure(n * sizeof(float4)));
  APD_CUDA(h, c.label.ensure(n * sizeof(float)));
  APD_CUDA(h, c.inv_perm.ensure(n * sizeof(int)));
  APD_CUDA(h, c.cell_start.ensure(((size_t)c.ncells + 1) * sizeof(uint32_t)));
  // scratch: keys x2, vals x2, hist, scan tmp
  const int sblocks = (c.n + kSortTile - 1) / kSortTile;
  const size_t hist_elems = (size_t)256 * sblocks;
  const size_t scan_elems = std::max(scan_tmp_elems_for((size_t)c.ncells + 1), scan_tmp_elems_for(hist_elems));
  const size_t kv = align_up(n * sizeof(uint32_t), 256);
  const size_t total = 4 * kv + align_up(hist_elems * 4, 256) + align_up(scan_elems * 4, 256);
  APD_CUDA(h, h->work.ensure(total));
  GridWork w;
  char* p = h->work.as<char>();
  w.keys[0] = (uint32_t*)p; p += kv;
  w.keys[1] = (uint32_t*)p; p += kv;
  w.vals[0] = (uint32_t*)p; p += kv;
  w.vals[1] = (uint32_t*)p; p += kv;
  w.hist = (uint32_t*)p; p += align_up(hist_elems * 4, 256);
  w.scan_tmp = (uint32_t*)p;
  w.scan_tmp_elems = scan_elems;
  {
    int rc = ensure_small(h);
    if (rc != APD_OK) return rc;
    w.ticket = reinterpret_cast<unsigned int*>(h->small.as<double>() + 42);
  }
  w.zero_flags = clear_flags ? c.cov_flag.as<unsigned char>() : nullptr;  // (allocated by the caller)
  c.flags_fresh = clear_flags;
  {
    ProfScope ps(h, APD_K_GRID);
    launch_grid_build(c.view(), w, h->stream, &h->launches);
  }
  APD_CUDA(h, cudaGetLastError());
  c.grid_valid = true;
  c.grid_ordered = true;
  return APD_OK;
}

// in-place all-gather of a per-point array whose chunks (see apd_handle::shard_chunk) this rank has just written:
// group j = chunks [j * nranks, (j + 1) * nranks) is contiguous and holds one chunk of every rank
int allgather_chunks(apd_handle* h, void* base, int n, size_t elems_per_point, int nccl_type, size_t elem_bytes) {
  const size_t chunk = (size_t)h->shard_chunk(n) * elems_per_point;
  for (int j = 0; j < h->shard_subs(); j++) {
    char* grp = reinterpret_cast<char*>(base) + (size_t)j * h->comm_size * chunk * elem_bytes;
    ncclResult_t r = g_nccl.AllGather(grp + (size_t)h->comm_rank * chunk * elem_bytes, grp, chunk, nccl_type, h->comm, h->stream);
    if (r != 0) return fail(h, APD_ERR_COMM, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "ncclAllGather failed");
  }
  return APD_OK;
}

PeerExchange barrier_ticket(apd_handle* h) {
  PeerExchange x;
  for (int r = 0; r < h->comm_size; r++) x.box[r] = h->peer_box[r];
  x.rank = h->comm_rank;
  x.nranks = h->comm_size;
  x.seq = ++h->bar_seq;
  return x;
}

// the same all-gather between the in-process ranks of a group: publish the array's address, meet the other rank threads,
// order the streams with a barrier kernel (every rank's chunks are complete), then PULL the peers' chunks
int group_allgather_chunks(apd_handle* h, void* base, int n, size_t bytes_per_point) {
  apd_group* g = h->group;
  g->gather_ptr[h->comm_rank] = base;
  if (!g->host_barrier()) return fail(h, APD_ERR_COMM, "a rank of the group did not reach the all-gather");
  launch_peer_barrier(barrier_ticket(h), h->stream, &h->launches);
  const size_t chunk = (size_t)h->shard_chunk(n) * bytes_per_point;
  for (int j = 0; j < h->shard_subs(); j++)
    for (int r = 0; r < h->comm_size; r++) {
      if (r == h->comm_rank) continue;
      const size_t off = ((size_t)j * h->comm_size + r) * chunk;
      APD_CUDA(h, cudaMemcpyAsync(reinterpret_cast<char*>(base) + off, reinterpret_cast<const char*>(g->gather_ptr[r]) + off, chunk,
                                  cudaMemcpyDefault, h->stream));
    }
  // nobody may rewrite its array (the next cloud) while a peer still pulls from it: every stream passes a second barrier
  // behind its own pulls
  launch_peer_barrier(barrier_ticket(h), h->stream, &h->launches);
  APD_CUDA(h, cudaGetLastError());
  // The rank threads share ONE process: whatever one of them does next that synchronises the device (a buffer of the
  // next cloud that grows: cudaFree; a first allocation) would wait for a barrier kernel that is still spinning for a
  // peer whose launch in turn waits for the driver's lock. So: every rank drains its stream, then the threads meet —
  // no barrier kernel of the group is left on the device when any of them goes on. (Seen as a 5 s time-out + NaN sums in
  // one run of the full test suite, never in the sharding tests alone.)
  APD_CUDA(h, cudaStreamSynchronize(h->stream));
  if (!g->host_barrier()) return fail(h, APD_ERR_COMM, "a rank of the group did not finish the all-gather");
  return APD_OK;
}

// FastAPDGICP::calculate_covariances (:351-411)
int ensure_covariances_of(apd_handle* h, Cloud& c) {
  int rc = ensure_grid(h, c);
  if (rc != APD_OK) return rc;
  if (c.cov_valid && c.geo_valid && c.geo_variant == h->params.variant) return APD_OK;
  const size_t np = h->padded_n(c.n);
  APD_CUDA(h, c.cov.ensure(np * 6 * sizeof(double)));
  APD_CUDA(h, c.geo.ensure(np * sizeof(float)));
  APD_CUDA(h, c.geo64.ensure(np * sizeof(double)));
  // Sharded: this rank searches the full grid for its slice of the points only; the covariance slices are then
  // all-gathered (the target's are gathered through arbitrary correspondences, and a swap can make either cloud the
  // target). The geometric weight is only ever read for the rank's own slice of the source, so it stays local.
  if (!c.cov_valid) {
    // a cloud whose covariances were started on demand is completed with the parameters they were started with
    const int k = c.cov_lazy ? c.lazy_k : h->params.k_correspondences;
    const int reg = c.cov_lazy ? c.lazy_reg : h->params.regularization;
    if (k < 1 || k > h->knn_max_k()) return fail(h, APD_ERR_UNSUPPORTED, "k_correspondences out of range (1..128; 1..32 in warp mode)");
    if (c.n < k) return fail(h, APD_ERR_TOO_FEW, "cloud has fewer points than k_correspondences");
    ProfScope ps(h, APD_K_KNN_COV);
    if (h->knn_use_warp(c.n, k)) APD_CUDA(h, h->nbuf.ensure((size_t)c.n * k * sizeof(int32_t)));
    for (int j = 0; j < h->shard_subs(); j++) {
      const int w0 = h->sub_begin(c.n, j), wn = h->sub_count(c.n, j);
      if (h->knn_use_warp(c.n, k)) {
        launch_knn_cov(c.view(), k, h->nbuf.as<int32_t>(), nullptr, h->stream, &h->launches, w0, wn);
        launch_cov_regularize(c.view(), k, reg, h->nbuf.as<int32_t>(), h->stream, &h->launches, w0, wn);
      } else {
        launch_knn_cov_fused(c.view(), k, reg, nullptr, h->stream, &h->launches, w0, wn);
      }
    }
    if (h->group) rc = group_allgather_chunks(h, c.cov.p, c.n, 6 * sizeof(double));
    else if (h->comm) rc = allgather_chunks(h, c.cov.p, c.n, 6, kNcclFloat64, sizeof(double));
    if (rc != APD_OK) return rc;
    c.cov_valid = true;
    c.geo_valid = true;
    c.geo_variant = APD_VARIANT_APDGICP;
    c.cov_lazy = false;
  } else if (!c.geo_valid || (c.geo_variant != h->params.variant && h->params.variant == APD_VARIANT_APDGICP)) {
    // after set*Covariances (every rank holds all covariances, the weight is a local map), or back from GICP
    ProfScope ps(h, APD_K_KNN_COV);
    launch_geo_weight(c.view(), h->stream, &h->launches);
    c.geo_valid = true;
    c.geo_variant = APD_VARIANT_APDGICP;
  }
  if (h->params.variant == APD_VARIANT_GICP && c.geo_variant != APD_VARIANT_GICP) {
    // FastGICP weighs every term by 1 (fast_gicp_impl.hpp:205): 1 + geo + cl with geo = 0 and cl = 0 is exactly 1
    APD_CUDA(h, cudaMemsetAsync(c.geo.p, 0, np * sizeof(float), h->stream));
    APD_CUDA(h, cudaMemsetAsync(c.geo64.p, 0, np * sizeof(double), h->stream));
    c.geo_variant = APD_VARIANT_GICP;
  }
  APD_CUDA(h, cudaGetLastError());
  return APD_OK;
}

int ensure_covariances(apd_handle* h) {
  if (!h->src.present || !h->tgt.present) return fail(h, APD_ERR_INVALID, "source or target cloud not set");
  int rc = ensure_covariances_of(h, h->src);  // :149-151
  if (rc != APD_OK) return rc;
  return ensure_covariances_of(h, h->tgt);    // :152-154
}

bool use_device_loop(const apd_handle* h);

// computeTransformation's covariance step (:148-157) for the device-resident loop: the source's covariances now, the
// target's on demand inside the loop when that pays (lm.cu: corr_phase). Same values either way.
int ensure_covariances_for_loop(apd_handle* h) {
  if (!h->src.present || !h->tgt.present) return fail(h, APD_ERR_INVALID, "source or target cloud not set");
  Cloud& t = h->tgt;
  const int k = t.cov_lazy ? t.lazy_k : h->params.k_correspondences;
  const bool lazy = !t.cov_valid && t.n > 0 && k >= 1 && k <= 32 && t.n >= k &&
                    (t.cov_lazy || h->lazy_mode == 1 || (h->lazy_mode < 0 && h->pooled && (long long)t.n >= (long long)h->lazy_min_ratio * h->src.n));
  if (!lazy) return ensure_covariances(h);
  int rc = ensure_covariances_of(h, h->src);
  if (rc != APD_OK) return rc;
  APD_CUDA(h, t.cov.ensure((size_t)t.n * 6 * sizeof(double)));
  APD_CUDA(h, t.cov_flag.ensure((size_t)t.n));
  APD_CUDA(h, h->nbuf.ensure((size_t)std::max(h->src.n, 1) * (k + 1) * sizeof(int32_t)));
  t.flags_fresh = false;
  rc = ensure_grid(h, t, !t.cov_lazy);  // a grid built now clears the flags on its way (saves the memset call)
  if (rc != APD_OK) return rc;
  if (!t.cov_lazy) {
    if (!t.flags_fresh) APD_CUDA(h, cudaMemsetAsync(t.cov_flag.p, 0, (size_t)t.n, h->stream));
    t.cov_lazy = true;
    t.lazy_k = k;
    t.lazy_reg = h->params.regularization;
    t.geo_valid = false;
  }
  return APD_OK;
}

// Which parts of the preparation the fused registration kernel takes over for the clouds as they are now: bit 0 the
// source's grid + covariances, bit 1 the target's grid; 0: nothing (everything is ready, or the case is not the fused
// kernel's: the separate kernels run). The fused kernel needs the target's covariances either valid already or computed
// on demand inside the loop, i.e. the conditions of ensure_covariances_for_loop.
int fused_prep_bits(const apd_handle* h) {
  if (!h->fused || !use_device_loop(h) || !h->src.present || !h->tgt.present) return 0;
  const Cloud& s = h->src;
  const Cloud& t = h->tgt;
  const int k = h->params.k_correspondences;
  if (k < 1 || k > 32 || s.n < k || t.n < k || t.n > h->fused_max_target || h->knn_mode == 2) return 0;
  const bool src_ready = s.grid_valid && s.cov_valid && s.geo_valid && s.geo_variant == h->params.variant;
  if (!src_ready && s.cov_valid) return 0;  // covariances given by the caller (setSourceCovariances): only the weights are missing
  const int tk = t.cov_lazy ? t.lazy_k : k;
  const bool lazy_ok = tk >= 1 && tk <= 32 && t.n >= tk &&
                       (t.cov_lazy || h->lazy_mode == 1 || (h->lazy_mode < 0 && h->pooled && (long long)t.n >= (long long)h->lazy_min_ratio * s.n));
  if (!t.cov_valid && !lazy_ok) return 0;
  if (t.cov_valid && !t.grid_valid) return 0;
  const int bits = (src_ready ? 0 : 1) | (t.grid_valid ? 0 : 2);
  return bits;
}

// buffers + cloud state for a fused launch with prep bits `bits` (the counterpart of ensure_covariances_for_loop)
int prepare_fused(apd_handle* h, int bits) {
  Cloud& s = h->src;
  Cloud& t = h->tgt;
  const int k = h->params.k_correspondences;
  const double s_cpp = h->max_cells_for(s.n), t_cpp = h->max_cells_for(t.n);  // (capacities: whichever resolution the launch picks)
  if (bits & 1) {
    const size_t n = (size_t)s.n;
    APD_CUDA(h, s.spts.ensure(n * sizeof(float4)));
    APD_CUDA(h, s.label.ensure(n * sizeof(float)));
    APD_CUDA(h, s.inv_perm.ensure(n * sizeof(int)));
    APD_CUDA(h, s.cell_start.ensure((size_t)grid_cell_capacity(s.n, s_cpp) * sizeof(uint32_t)));
    APD_CUDA(h, s.cov.ensure(n * 6 * sizeof(double)));
    APD_CUDA(h, s.geo.ensure(n * sizeof(float)));
    APD_CUDA(h, s.geo64.ensure(n * sizeof(double)));
  }
  if (bits & 2) {
    const size_t n = (size_t)t.n;
    APD_CUDA(h, t.spts.ensure(n * sizeof(float4)));
    APD_CUDA(h, t.label.ensure(n * sizeof(float)));
    APD_CUDA(h, t.inv_perm.ensure(n * sizeof(int)));
    APD_CUDA(h, t.cell_start.ensure((size_t)grid_cell_capacity(t.n, t_cpp) * sizeof(uint32_t)));
  }
  APD_CUDA(h, h->work.ensure(3 * ((size_t)s.n + (size_t)t.n) * sizeof(uint32_t) + 512));
  if (!t.cov_valid) {  // target covariances on demand
    APD_CUDA(h, t.cov.ensure((size_t)t.n * 6 * sizeof(double)));
    APD_CUDA(h, t.cov_flag.ensure((size_t)t.n));
    if (!(bits & 2) && !t.cov_lazy) APD_CUDA(h, cudaMemsetAsync(t.cov_flag.p, 0, (size_t)t.n, h->stream));  // (a grid built in the kernel clears them on its way)
    if (!t.cov_lazy) {
      t.cov_lazy = true;
      t.lazy_k = k;
      t.lazy_reg = h->params.regularization;
      t.geo_valid = false;
    }
  }
  const int kmax = std::max(k, t.cov_lazy ? t.lazy_k : k);
  APD_CUDA(h, h->nbuf.ensure((size_t)std::max(s.n, 1) * (kmax + 1) * sizeof(int32_t)));
  return APD_OK;
}

constexpr int kReduceSeqSlot = 58;  // (double index into h_small: the sequence word of the zero-copy reductions)
int ensure_small(apd_handle* h) {
  // [0..27] out28, [32..34] fitness out3, then tickets (uint) at double index 40, 41
  // device: [0..27] out28, [32..34] fitness out3, tickets (uint) at double index 40-42, the state of the bounding-box
  // reduction at 48 (source) / 52 (target), 56: its output when zero-copy is off. host: the same indices; a box is
  // 6 uints + its 64-bit sequence number (48..51 / 52..55)
  if (!h->small.p) {
    APD_CUDA(h, h->small.ensure(64 * sizeof(double)));
    APD_CUDA(h, cudaMemsetAsync(h->small.p, 0, 64 * sizeof(double), h->stream));
    init_bounds_state(reinterpret_cast<unsigned int*>(h->small.as<double>() + 48), h->stream);
    init_bounds_state(reinterpret_cast<unsigned int*>(h->small.as<double>() + 52), h->stream);
  }
  if (!h->partials.p) APD_CUDA(h, h->partials.ensure((size_t)h->max_reduce_blocks * kReduceVals * sizeof(double)));
  if (!h->h_small.p) {
    APD_CUDA(h, h->h_small.ensure(64 * sizeof(double)));
    std::memset(h->h_small.p, 0, 64 * sizeof(double));
    void* d = nullptr;
    if (h->zero_copy && cudaHostGetDevicePointer(&d, h->h_small.p, 0) == cudaSuccess) h->h_small_dev = reinterpret_cast<unsigned int*>(d);
    else h->zero_copy = false;
  }
  return APD_OK;
}

// per-linearisation arrays of this rank's source points (rank-local slots, see ShardTable)
CorrOut corr_view(apd_handle* h) {
  CorrOut c;
  c.corr = h->corr.as<int>();
  c.sqd = h->sqd.as<float>();
  c.second = h->second.as<float>();
  c.second_valid = h->second_valid ? 1 : 0;
  c.maha_fp64 = h->corr_fp64;
  c.mahaA = h->mahaA.p;
  c.mahaB = h->mahaB.p;  // fp64 storage: two planes of local_n double2
  return c;
}

int vgicp_update_correspondences(apd_handle* h, const hm::Pose& T);
int vgicp_reduce(apd_handle* h, const hm::Pose& T, bool want_hb, double* H36, double* b6, double* err);

// FastAPDGICP::update_correspondences (:160-220)
int do_update_correspondences(apd_handle* h, const hm::Pose& T) {
  if (h->params.variant == APD_VARIANT_VGICP) return vgicp_update_correspondences(h, T);
  int rc0 = ensure_corr_buffers(h);
  if (rc0 != APD_OK) return rc0;
  const NoiseParams np = noise_params(h->params);
  {
    ProfScope ps(h, APD_K_CORR);
    const bool warm = h->corr_warm && h->corr_n == h->src.n && h->corr_thr == h->params.max_correspondence_distance;
    if (!warm) h->second_valid = false;
    const int lanes = launch_update_correspondences(h->src.view(), h->tgt.view(), h->shard_table(h->src.n), to_pose_d(T), np, corr_view(h),
                                                    warm ? &h->corr_pose : nullptr, h->corr_lanes, h->stream, &h->launches);
    h->second_valid = lanes == 1;
  }
  APD_CUDA(h, cudaGetLastError());
  h->corr_n = h->src.n;
  h->corr_warm = true;
  h->corr_pose = to_pose_d(T);
  h->corr_thr = h->params.max_correspondence_distance;
  return APD_OK;
}

// Launches K4/K5, all-reduces across shards if a communicator is set, and
// brings the 28 doubles back. out: H (row-major 36, optional), b (6, optional), err.
int reduce_pass(apd_handle* h, const hm::Pose& T, bool want_hb, double* H36, double* b6, double* err) {
  if (h->params.variant == APD_VARIANT_VGICP) return vgicp_reduce(h, T, want_hb, H36, b6, err);
  int rc = ensure_small(h);
  if (rc != APD_OK) return rc;
  double* d_out = h->small.as<double>();
  bool published = false;
  ReduceWork w;
  w.partials = h->partials.as<double>();
  w.ticket = reinterpret_cast<unsigned int*>(d_out + 40);
  w.max_blocks = h->max_reduce_blocks / 2;  // persistent CTAs: 2 resident per SM
  // correspondences_.size() (:273): the whole source cloud, sharded or not; GICP has no label weight (1 / inf = 0)
  const double n_total = h->params.variant == APD_VARIANT_GICP ? std::numeric_limits<double>::infinity() : (double)h->src.n;
  {
    ProfScope ps(h, want_hb ? APD_K_LINEARIZE : APD_K_ERROR);
    // ranks of one process: nothing that synchronises the device (a buffer that grows: cudaFree) may run on one rank's
    // thread while another rank's kernel already waits for it inside the exchange — meet first, launch after
    if (h->group && !h->group->host_barrier()) return fail(h, APD_ERR_COMM, "a rank of the group did not reach the reduction");
    if (h->peers_attached) {  // the kernel's last block also exchanges the totals with the peers
      for (int r = 0; r < h->comm_size; r++) w.xchg.box[r] = h->peer_box[r];
      w.xchg.rank = h->comm_rank;
      w.xchg.nranks = h->comm_size;
      w.xchg.seq = ++h->xchg_seq;
      if (w.xchg.seq == 0) w.xchg.seq = h->xchg_seq = 2;  // (wrap-around: 0 means "no exchange"; keep the parity alternating)
    }
    // the result goes straight into pinned host memory when nothing else has to happen to it on the stream first
    published = h->zero_copy && h->h_small_dev && !h->profiling && !(h->comm && !h->peers_attached);
    if (published) {
      w.host_out = reinterpret_cast<double*>(h->h_small_dev);
      w.host_seq_word = reinterpret_cast<unsigned long long*>(reinterpret_cast<double*>(h->h_small_dev) + kReduceSeqSlot);
      w.host_seq = ++h->reduce_seq;
    }
    // ONE launch serves all the rank's chunks (a chunk table in the kernel; round 1 launched once per chunk)
    launch_linearize(h->src.view(), h->tgt.view(), h->shard_table(h->src.n), to_pose_d(T), corr_view(h), n_total, want_hb, w, d_out,
                     h->stream, &h->launches);
  }
  APD_CUDA(h, cudaGetLastError());
  if (h->comm && !h->peers_attached) {
    ncclResult_t r = want_hb ? g_nccl.AllReduce(d_out, d_out, kReduceVals, kNcclFloat64, kNcclSum, h->comm, h->stream)
                             : g_nccl.AllReduce(d_out + 27, d_out + 27, 1, kNcclFloat64, kNcclSum, h->comm, h->stream);
    if (r != 0) return fail(h, APD_ERR_COMM, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "ncclAllReduce failed");
  }
  double* hs = reinterpret_cast<double*>(h->h_small.p);
  if (published) {  // the kernel wrote the totals into hs itself: look at the sequence word (no copy, no stream wait)
    APD_CUDA(h, wait_host_seq(h, reinterpret_cast<const volatile unsigned long long*>(hs + kReduceSeqSlot), h->reduce_seq));
  } else {
    APD_CUDA(h, cudaMemcpyAsync(hs, d_out, kReduceVals * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    APD_CUDA(h, wait_stream(h));
  }
  // (ranks of one process: nobody goes on — possibly to a call that synchronises the device — while a peer's kernel may
  // still be waiting inside the exchange for a rank whose launch has not gone through yet; see group_allgather_chunks)
  if (h->group && !h->group->host_barrier()) return fail(h, APD_ERR_COMM, "a rank of the group did not finish the reduction");
  if (want_hb) {
    if (H36) hm::unpack_upper(hs, H36);
    if (b6) std::memcpy(b6, hs + 21, 6 * sizeof(double));
  }
  *err = hs[27];
  return APD_OK;
}

// ---- FastVGICP (vgicp.cu) ----------------------------------------------------------------------------------------
int vgicp_offsets(const apd_params& p) { return p.voxel_search == APD_VOXEL_DIRECT1 ? 1 : (p.voxel_search == APD_VOXEL_DIRECT7 ? 7 : 27); }

VoxelMapDev voxel_view(const apd_handle* h) {
  VoxelMapDev v;
  v.n = h->n_vox;
  v.g = h->vgd;
  v.key = h->vkey.as<uint32_t>();
  v.cnt = h->vcnt.as<int32_t>();
  v.mean = h->vmean.as<double>();
  v.cov = h->vcov.as<double>();
  return v;
}

// GaussianVoxelMap::create_voxelmap (fast_vgicp_voxel.hpp:131-158) over the target and its covariances
int vgicp_build_voxelmap(apd_handle* h) {
  Cloud& t = h->tgt;
  if (!t.present || t.n <= 0) return fail(h, APD_ERR_INVALID, "target cloud not set");
  const int n = t.n;
  int rc = ensure_small(h);
  if (rc != APD_OK) return rc;
  const CloudDev tv = t.view();
  // the box of occupied voxel coordinates, from the box of the finite points (voxel_coord is monotone in x)
  unsigned int* d_state = reinterpret_cast<unsigned int*>(h->small.as<double>() + 60);
  const unsigned int init[6] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u};
  APD_CUDA(h, cudaMemcpyAsync(d_state, init, sizeof(init), cudaMemcpyHostToDevice, h->stream));
  launch_voxel_bounds(tv.pts, n, d_state, h->stream, &h->launches);
  unsigned int enc[6];
  APD_CUDA(h, cudaMemcpyAsync(enc, d_state, sizeof(enc), cudaMemcpyDeviceToHost, h->stream));
  APD_CUDA(h, wait_stream(h));
  if (enc[0] == 0xffffffffu) return fail(h, APD_ERR_INVALID, "the target holds no finite point");
  const double res = h->params.voxel_resolution;
  VoxelGridDesc vg;
  vg.res = res;
  long long cells = 1;
  for (int a = 0; a < 3; a++) {
    float lo, hi;
    unsigned int u = enc[a];
    u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
    std::memcpy(&lo, &u, 4);
    u = enc[3 + a];
    u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
    std::memcpy(&hi, &u, 4);
    const double clo = std::floor((double)lo / res - 0.5), chi = std::floor((double)hi / res - 0.5);
    if (!(std::fabs(clo) < 1e9) || !(std::fabs(chi) < 1e9)) return fail(h, APD_ERR_UNSUPPORTED, "voxel coordinates out of range");
    const long long dim = (long long)(chi - clo) + 1;
    vg.mn[a] = (int)clo;
    vg.dim[a] = (int)std::min<long long>(dim, 0x7fffffffll);
    if (dim >= 0xffffffffll || cells * dim >= 0xffffffffll)
      return fail(h, APD_ERR_UNSUPPORTED, "voxel_resolution too small for the extent of the target (more than 2^32 voxel slots)");
    cells *= dim;
  }
  int bits = 1;
  while (bits < 32 && (1ll << bits) < cells) bits++;
  bits = std::min(32, (bits + 7) / 8 * 8);
  const int sblocks = (n + kSortTile - 1) / kSortTile;
  const size_t hist_elems = (size_t)256 * sblocks;
  const size_t scan_elems = std::max(scan_tmp_elems_for((size_t)n + 1), scan_tmp_elems_for(hist_elems));
  const size_t kv = align_up((size_t)n * sizeof(uint32_t), 256);
  APD_CUDA(h, h->work.ensure(5 * kv + 256 + align_up(hist_elems * 4, 256) + align_up(scan_elems * 4, 256)));
  char* p = h->work.as<char>();
  uint32_t* keys[2] = {(uint32_t*)p, (uint32_t*)(p + kv)};
  uint32_t* vals[2] = {(uint32_t*)(p + 2 * kv), (uint32_t*)(p + 3 * kv)};
  uint32_t* heads = (uint32_t*)(p + 4 * kv);
  uint32_t* hist = (uint32_t*)(p + 5 * kv + 256);
  uint32_t* scan_tmp = (uint32_t*)(p + 5 * kv + 256 + align_up(hist_elems * 4, 256));
  launch_vgicp_keys(tv.pts, n, vg, keys[0], vals[0], h->stream, &h->launches);
  const int cur = launch_sort_pairs_u32(keys, vals, n, 32, hist, scan_tmp, h->stream, &h->launches);
  (void)bits;
  launch_voxel_heads(keys[cur], n, heads, h->stream, &h->launches);
  launch_exclusive_scan_u32(heads, (size_t)n + 1, scan_tmp, h->stream, &h->launches);
  uint32_t nv = 0;
  APD_CUDA(h, cudaMemcpyAsync(&nv, heads + n, sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
  APD_CUDA(h, wait_stream(h));
  APD_CUDA(h, h->vkey.ensure(std::max<size_t>(1, nv) * sizeof(uint32_t)));
  APD_CUDA(h, h->vcnt.ensure(std::max<size_t>(1, nv) * sizeof(int32_t)));
  APD_CUDA(h, h->vmean.ensure(std::max<size_t>(1, nv) * 3 * sizeof(double)));
  APD_CUDA(h, h->vcov.ensure(std::max<size_t>(1, nv) * 6 * sizeof(double)));
  launch_vgicp_voxels(tv, keys[cur], vals[cur], heads, h->params.voxel_mode == APD_VOXEL_MULTIPLICATIVE ? 1 : 0, h->vkey.as<uint32_t>(),
                      h->vcnt.as<int32_t>(), h->vmean.as<double>(), h->vcov.as<double>(), h->stream, &h->launches);
  APD_CUDA(h, cudaGetLastError());
  h->n_vox = (int)nv;
  h->vgd = vg;
  h->vox_valid = true;
  return APD_OK;
}

// FastVGICP::update_correspondences (fast_vgicp_impl.hpp:74-118); builds the voxel map first if there is none (:126-129)
int vgicp_update_correspondences(apd_handle* h, const hm::Pose& T) {
  if (h->sharded()) return fail(h, APD_ERR_UNSUPPORTED, "FastVGICP is not sharded");
  if (!h->vox_valid) {
    ProfScope ps(h, APD_K_GRID);
    const int rc = vgicp_build_voxelmap(h);
    if (rc != APD_OK) return rc;
  }
  const int no = vgicp_offsets(h->params);
  const size_t slots = (size_t)std::max(1, h->src.n) * no;
  APD_CUDA(h, h->vcorr.ensure(slots * sizeof(int32_t)));
  APD_CUDA(h, h->vmaha.ensure(slots * 6 * sizeof(double)));
  {
    ProfScope ps(h, APD_K_CORR);
    launch_vgicp_correspondences(h->src.view(), voxel_view(h), h->params.voxel_search, no, to_pose_d(T), h->vcorr.as<int32_t>(), h->vmaha.as<double>(),
                                 h->stream, &h->launches);
  }
  APD_CUDA(h, cudaGetLastError());
  h->vcorr_n = h->src.n;
  h->vcorr_noff = no;
  return APD_OK;
}

// the sums of FastVGICP::linearize (:141-181) / compute_error (:186-205) over the stored correspondence table
int vgicp_reduce(apd_handle* h, const hm::Pose& T, bool want_hb, double* H36, double* b6, double* err) {
  if (h->vcorr_n != h->src.n || h->vcorr_noff != vgicp_offsets(h->params) || !h->vox_valid)
    return fail(h, APD_ERR_INVALID, "compute_error before linearize (no voxel correspondences)");
  int rc = ensure_small(h);
  if (rc != APD_OK) return rc;
  double* d_out = h->small.as<double>();
  {
    ProfScope ps(h, want_hb ? APD_K_LINEARIZE : APD_K_ERROR);
    launch_vgicp_reduce(h->src.view(), voxel_view(h), h->vcorr.as<int32_t>(), h->vmaha.as<double>(), h->vcorr_noff, to_pose_d(T), want_hb,
                        h->partials.as<double>(), h->max_reduce_blocks, d_out, reinterpret_cast<unsigned int*>(d_out + 40), h->stream, &h->launches);
  }
  APD_CUDA(h, cudaGetLastError());
  double* hs = reinterpret_cast<double*>(h->h_small.p);
  APD_CUDA(h, cudaMemcpyAsync(hs, d_out, kReduceVals * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  APD_CUDA(h, wait_stream(h));
  if (want_hb) {
    if (H36) hm::unpack_upper(hs, H36);
    if (b6) std::memcpy(b6, hs + 21, 6 * sizeof(double));
  }
  *err = hs[27];
  return APD_OK;
}

// FastAPDGICP::linearize (:224-307)
int do_linearize(apd_handle* h, const hm::Pose& T, double* H36, double* b6, double* err) {
  int rc = do_update_correspondences(h, T);  // :226
  if (rc != APD_OK) return rc;
  return reduce_pass(h, T, H36 != nullptr && b6 != nullptr, H36, b6, err);
}

void trace_row(apd_handle* h, int outer, int inner, double y0, double yi, double rho, double lambda, double dn, bool accepted) {
  const double row[8] = {(double)outer, (double)inner, y0, yi, rho, lambda, dn, accepted ? 1.0 : 0.0};
  h->trace.insert(h->trace.end(), row, row + 8);
}

// LsqRegistration::step_gn (lsq_registration_impl.hpp:107-123)
int step_gn(apd_handle* h, int outer, hm::Pose& x0, hm::Pose& delta) {
  double H[36], b[6], y0;
  int rc = do_linearize(h, x0, H, b, &y0);
  if (rc != APD_OK) return rc;
  double nb[6], d[6];
  for (int i = 0; i < 6; i++) nb[i] = -b[i];
  hm::ldlt_solve6(H, nb, d);
  delta = hm::delta_from_twist(d);
  x0 = hm::compose(delta, x0);
  std::memcpy(h->final_H, H, sizeof(H));
  double dn = 0;
  for (int i = 0; i < 6; i++) dn += d[i] * d[i];
  trace_row(h, outer, 0, y0, y0, 0.0, 0.0, std::sqrt(dn), true);
  return APD_OK;
}

// LsqRegistration::step_lm (lsq_registration_impl.hpp:127-173); *ok = its return value
int step_lm(apd_handle* h, int outer, hm::Pose& x0, hm::Pose& delta, bool* ok) {
  double H[36], b[6], y0;
  int rc = do_linearize(h, x0, H, b, &y0);  // :130
  if (rc != APD_OK) return rc;
  if (h->lm_lambda < 0.0) {  // :131-133
    double mx = 0.0;
    for (int i = 0; i < 6; i++) mx = std::max(mx, std::fabs(H[i * 6 + i]));
    h->lm_lambda = h->params.lm_init_lambda_factor * mx;
  }
  double nu = 2.0;
  for (int i = 0; i < h->params.lm_max_iterations; i++) {  // :136
    double Hl[36], nb[6], d[6];
    std::memcpy(Hl, H, sizeof(H));
    for (int j = 0; j < 6; j++) {
      Hl[j * 6 + j] += h->lm_lambda;
      nb[j] = -b[j];
    }
    hm::ldlt_solve6(Hl, nb, d);  // :137-138
    delta = hm::delta_from_twist(d);  // :140-142
    const hm::Pose xi = hm::compose(delta, x0);  // :144
    double yi;
    rc = reduce_pass(h, xi, false, nullptr, nullptr, &yi);  // compute_error(xi) :145
    if (rc != APD_OK) return rc;
    double denom = 0.0, dn = 0.0;
    for (int j = 0; j < 6; j++) {
      denom += d[j] * (h->lm_lambda * d[j] - b[j]);
      dn += d[j] * d[j];
    }
    const double rho = (y0 - yi) / denom;  // :146
    if (h->params.lm_debug_print) {       // :148-154
      if (i == 0) std::printf("--- LM optimization ---\n%5s %15s %15s %15s %15s %15s %5s\n", "i", "y0", "yi", "rho", "lambda", "|delta|", "dec");
      std::printf("%5d %15g %15g %15g %15g %15g %5c\n", i, y0, yi, rho, h->lm_lambda, std::sqrt(dn), rho > 0.0 ? 'x' : ' ');
    }
    trace_row(h, outer, i, y0, yi, rho, h->lm_lambda, std::sqrt(dn), !(rho < 0));
    if (rho < 0) {  // :156-164
      if (hm::is_converged(delta, h->params.rotation_epsilon, h->params.transformation_epsilon)) {
        *ok = true;
        return APD_OK;
      }
      h->lm_lambda = nu * h->lm_lambda;
      nu = 2 * nu;
      continue;
    }
    x0 = xi;  // :166-169
    h->lm_lambda = h->lm_lambda * std::max(1.0 / 3.0, 1 - std::pow(2 * rho - 1, 3));
    std::memcpy(h->final_H, H, sizeof(H));
    *ok = true;
    return APD_OK;
  }
  *ok = false;  // :172
  return APD_OK;
}

// The cloud being set is the one the OTHER slot already holds (same cache key, same size): a frame promoted to keyframe
// (scan_matching_odometry_nodelet.cpp:587-588 sets the just-registered source as the new target). Its grid and
// covariances are pure functions of the cloud, so they are copied device-to-device instead of rebuilt (the reference
// recomputes them).
int adopt_cloud(apd_handle* h, Cloud& dst, const Cloud& src) {
  if (h->src.bbox_launched || h->tgt.bbox_launched) {  // (a box in flight is collected first: the read-back slots belong to the roles)
    const int rc = finish_bboxes(h);
    if (rc != APD_OK) return rc;
  }
  dst.bbox_launched = false;
  dst.host_packed = nullptr;
  dst.bbox_pending = !src.bbox_known;  // (reduced on the device if a grid of the copy is ever sized on the host)
  const size_t n = (size_t)src.n;
  auto copy = [&](DevBuf& d, const DevBuf& s_, size_t bytes) -> cudaError_t {
    if (!s_.p || bytes == 0) return cudaSuccess;
    cudaError_t e = d.ensure(bytes);
    if (e != cudaSuccess) return e;
    return cudaMemcpyAsync(d.p, s_.p, bytes, cudaMemcpyDeviceToDevice, h->stream);
  };
  if (src.ext_pts) {
    dst.ext_pts = src.ext_pts;
  } else {
    dst.ext_pts = nullptr;
    APD_CUDA(h, copy(dst.pts, src.pts, n * sizeof(float4)));
  }
  if (src.grid_valid) {
    APD_CUDA(h, copy(dst.spts, src.spts, n * sizeof(float4)));
    APD_CUDA(h, copy(dst.label, src.label, n * sizeof(float)));
    APD_CUDA(h, copy(dst.inv_perm, src.inv_perm, n * sizeof(int)));
    APD_CUDA(h, copy(dst.cell_start, src.cell_start, ((size_t)src.ncells + 1) * sizeof(uint32_t)));
  }
  if (src.grid_valid && src.cov_valid) APD_CUDA(h, copy(dst.cov, src.cov, h->padded_n(src.n) * 6 * sizeof(double)));
  if (src.grid_valid && src.cov_valid && src.geo_valid) {
    APD_CUDA(h, copy(dst.geo, src.geo, n * sizeof(float)));
    APD_CUDA(h, copy(dst.geo64, src.geo64, n * sizeof(double)));
  }
  dst.n = src.n;
  dst.present = true;
  dst.key = src.key;
  dst.print = src.print;
  std::memcpy(dst.bbox, src.bbox, sizeof(dst.bbox));
  dst.bbox_known = src.bbox_known && !src.host_packed;
  dst.g = src.g;
  dst.ncells = src.ncells;
  dst.grid_valid = src.grid_valid;
  dst.grid_ordered = src.grid_ordered;
  dst.cov_valid = src.grid_valid && src.cov_valid;
  dst.geo_valid = dst.cov_valid && src.geo_valid;
  dst.geo_variant = src.geo_variant;
  dst.cov_lazy = false;  // (a partly computed target is not adopted: the new owner computes what it needs)
  return APD_OK;
}

// The reference early-outs on shared_ptr identity, and the pointer keeps the cloud alive. Here the key is a number the
// caller chose (usually the cloud's address), which can go stale: a freed cloud's address is reused by the next frame.
// A key is therefore only honoured together with the point count and a fingerprint of the content (64 points spread
// over the cloud: xyz + label bits, FNV-1a) taken when the key was accepted — a recycled address with other points in
// it is a new cloud.
uint64_t cloud_fingerprint(const void* pts, int32_t n, int32_t stride, int32_t xyz_off, int32_t label_off) {
  uint64_t hsh = 1469598103934665603ull ^ (uint64_t)(uint32_t)n;
  const int samples = std::min<int32_t>(n, 64);
  for (int q = 0; q < samples; q++) {
    const size_t i = samples > 1 ? (size_t)q * (size_t)(n - 1) / (size_t)(samples - 1) : 0;
    const unsigned char* p = reinterpret_cast<const unsigned char*>(pts) + i * (size_t)stride;
    uint32_t w[4] = {0, 0, 0, 0};
    std::memcpy(w, p + xyz_off, 12);
    if (label_off >= 0) std::memcpy(w + 3, p + label_off, 4);
    for (int e = 0; e < 4; e++) {
      hsh ^= w[e];
      hsh *= 1099511628211ull;
    }
  }
  return hsh;
}

int set_cloud(apd_handle* h, Cloud& c, const void* pts, int32_t n, int32_t stride, int32_t xyz_off, int32_t label_off, uint64_t key) {
  if (!pts || n < 0 || stride < 12 || xyz_off < 0) return fail(h, APD_ERR_INVALID, "bad cloud arguments");
  const uint64_t print = key != 0 ? cloud_fingerprint(pts, n, stride, xyz_off, label_off) : 0;
  if (c.present && key != 0 && key == c.key && n == c.n && print == c.print) return APD_OK;  // pointer-identity early-out (:116,:128)
  DeviceGuard dg(h->device);
  {
    Cloud& other = (&c == &h->src) ? h->tgt : h->src;
    if (key != 0 && other.present && other.key == key && other.n == n && other.print == print && !h->sharded()) {
      h->corr_n = -1;
      h->corr_warm = false;
      return adopt_cloud(h, c, other);
    }
  }
  // A pool copies packed float4 clouds straight out of the caller's memory when that memory is page-locked (the
  // caller's buffers stay valid until apd_batch_align returns, i.e. after the copy has run): no staging pass — at 8
  // GPUs x 20 k registrations/s the 1 MB staging copy per target is ~300 GB/s of host memory traffic by itself.
  if (h->pooled && n > 0 && stride == 16 && xyz_off == 0 && label_off == 12) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, pts) == cudaSuccess && at.type == cudaMemoryTypeHost) {
      c.host_packed = pts;  // (its box is computed only if a grid is sized on the host: finish_bboxes)
      APD_CUDA(h, c.pts.ensure((size_t)n * sizeof(float4)));
      APD_CUDA(h, cudaMemcpyAsync(c.pts.p, pts, (size_t)n * sizeof(float4), cudaMemcpyHostToDevice, h->stream));
      c.ext_pts = nullptr;
      c.bbox_pending = false;
      c.bbox_launched = false;
      c.bbox_known = false;
      c.n = n;
      c.present = true;
      c.key = key;
      c.print = print;
      c.grid_valid = false;
      c.cov_valid = false;  // source_covs_.clear() (:122,:133)
      c.geo_valid = false;
      c.cov_lazy = false;
      h->corr_warm = false;
      return APD_OK;
    }
    (void)cudaGetLastError();  // (older runtimes report unregistered host memory as an error)
  }
  // the previous H2D out of this staging buffer must have completed before it is overwritten
  if (c.staged_by_seq) {
    // (zero-copy handles: no event per copy — a later kernel of this stream was seen to finish, which is the normal case
    // of a pool worker reusing its handle; otherwise wait for the stream)
    if (h->seq_seen <= c.staged_seq) APD_CUDA(h, wait_stream(h));
  } else if (c.staged) {
    if (h->blocking_wait && h->poll_wait_us > 0) {
      cudaError_t e;
      while ((e = cudaEventQuery(c.staged)) == cudaErrorNotReady) std::this_thread::sleep_for(std::chrono::microseconds(h->poll_wait_us));
      APD_CUDA(h, e);
    } else {
      APD_CUDA(h, cudaEventSynchronize(c.staged));
    }
  }
  APD_CUDA(h, c.stage.ensure((size_t)std::max(n, 1) * sizeof(float4)));
  // pool workers stage many clouds at once and the host's memory system is what bounds them: non-temporal stores
  // (APD_STAGE_NT=0|1 overrides; a lone handle keeps ordinary stores — the buffer is small enough to stay in cache)
  static const int nt_env = [] { const char* e = std::getenv("APD_STAGE_NT"); return e ? (std::atoi(e) != 0 ? 1 : 0) : -1; }();
  stage_cloud(pts, n, stride, xyz_off, label_off, reinterpret_cast<float*>(c.stage.p), c.bbox, nt_env >= 0 ? nt_env != 0 : h->pooled);
  APD_CUDA(h, c.pts.ensure((size_t)std::max(n, 1) * sizeof(float4)));
  APD_CUDA(h, cudaMemcpyAsync(c.pts.p, c.stage.p, (size_t)n * sizeof(float4), cudaMemcpyHostToDevice, h->stream));
  if (h->zero_copy) {
    c.staged_by_seq = true;
    c.staged_seq = h->seq;  // kernels enqueued from now on carry larger numbers
  } else {
    c.staged_by_seq = false;
    if (!c.staged) APD_CUDA(h, cudaEventCreateWithFlags(&c.staged, cudaEventDisableTiming | (h->blocking_wait ? cudaEventBlockingSync : 0)));
    APD_CUDA(h, cudaEventRecord(c.staged, h->stream));
  }
  c.ext_pts = nullptr;
  c.bbox_pending = false;
  c.bbox_launched = false;
  c.bbox_known = true;  // (found by the staging pass)
  c.host_packed = nullptr;
  c.n = n;
  c.present = true;
  c.key = key;
  c.print = print;
  c.grid_valid = false;
  c.cov_valid = false;  // source_covs_.clear() (:122,:133)
  c.geo_valid = false;
  c.cov_lazy = false;
  h->corr_warm = false;
  return APD_OK;
}

// the bounding box of a device-resident cloud: reduced by a kernel that publishes it into pinned host memory
int launch_bounds_of(apd_handle* h, Cloud& c) {
  int rc = ensure_small(h);
  if (rc != APD_OK) return rc;
  const int slot = (&c == &h->src) ? 48 : 52;  // per-cloud slots: both boxes may be in flight at once
  unsigned int* d_state = reinterpret_cast<unsigned int*>(h->small.as<double>() + slot);  // prepared by ensure_small
  unsigned int* enc = reinterpret_cast<unsigned int*>(reinterpret_cast<double*>(h->h_small.p) + slot);
  c.bbox_seq = ++h->seq;
  const float4* d_xyzl = c.ext_pts ? c.ext_pts : c.pts.as<const float4>();
  if (h->zero_copy) {
    launch_bounds(d_xyzl, c.n, d_state, h->h_small_dev + 2 * slot, c.bbox_seq, h->stream, &h->launches);
  } else {  // the kernel publishes into device scratch (small[56..59]), copied back in stream order
    unsigned int* d_out = reinterpret_cast<unsigned int*>(h->small.as<double>() + 56);
    launch_bounds(d_xyzl, c.n, d_state, d_out, c.bbox_seq, h->stream, &h->launches);
    APD_CUDA(h, cudaMemcpyAsync(enc, d_out, 6 * sizeof(unsigned int), cudaMemcpyDeviceToHost, h->stream));
  }
  APD_CUDA(h, cudaGetLastError());
  c.bbox_launched = true;
  return APD_OK;
}

int set_cloud_device(apd_handle* h, Cloud& c, const void* d_xyzl, int32_t n) {
  if (!d_xyzl || n < 0) return fail(h, APD_ERR_INVALID, "bad cloud arguments");
  DeviceGuard dg(h->device);
  c.bbox_pending = true;     // (its box is reduced when — and if — a grid is sized on the host: finish_bboxes)
  c.bbox_launched = false;
  c.bbox_known = false;
  c.host_packed = nullptr;
  c.ext_pts = reinterpret_cast<const float4*>(d_xyzl);
  h->corr_warm = false;
  c.n = n;
  c.present = true;
  c.key = 0;
  c.grid_valid = false;
  c.cov_valid = false;
  c.geo_valid = false;
  c.cov_lazy = false;
  return APD_OK;
}

int set_covs(apd_handle* h, Cloud& c, const double* covs, int32_t n) {
  if (!c.present || n != c.n || !covs) return fail(h, APD_ERR_INVALID, "covariance count does not match the cloud");
  DeviceGuard dg(h->device);
  int rc = ensure_grid(h, c);
  if (rc != APD_OK) return rc;
  APD_CUDA(h, c.cov.ensure(h->padded_n(n) * 6 * sizeof(double)));
  APD_CUDA(h, c.geo.ensure(h->padded_n(n) * sizeof(float)));
  APD_CUDA(h, c.geo64.ensure(h->padded_n(n) * sizeof(double)));
  APD_CUDA(h, h->scratch.ensure((size_t)n * 16 * sizeof(double)));
  APD_CUDA(h, cudaMemcpyAsync(h->scratch.p, covs, (size_t)n * 16 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  launch_cov_import(c.view(), h->scratch.as<double>(), h->stream, &h->launches);
  APD_CUDA(h, wait_stream(h));
  c.cov_valid = true;
  c.geo_valid = false;
  c.cov_lazy = false;
  return APD_OK;
}

int get_covs(apd_handle* h, Cloud& c, double* covs, int32_t n) {
  if (!c.present || n != c.n || !covs) return fail(h, APD_ERR_INVALID, "covariance count does not match the cloud");
  DeviceGuard dg(h->device);
  int rc = ensure_covariances_of(h, c);
  if (rc != APD_OK) return rc;
  APD_CUDA(h, h->scratch.ensure((size_t)n * 16 * sizeof(double)));
  launch_cov_export(c.view(), h->scratch.as<double>(), h->stream, &h->launches);
  APD_CUDA(h, cudaMemcpyAsync(covs, h->scratch.p, (size_t)n * 16 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  APD_CUDA(h, wait_stream(h));
  flush_prof(h);
  return APD_OK;
}

void default_params(apd_params* p) {
  std::memset(p, 0, sizeof(*p));
  p->k_correspondences = 20;
  p->regularization = APD_REG_PLANE;
  p->max_correspondence_distance = (double)std::numeric_limits<float>::max();
  p->dist_var = 0.86;
  p->azimuth_var = 0.5;
  p->elevation_var = 1.0;
  p->max_iterations = 64;
  p->optimizer = APD_OPT_LEVENBERG_MARQUARDT;
  p->rotation_epsilon = 2e-3;
  p->transformation_epsilon = 5e-4;
  p->lm_max_iterations = 10;
  p->lm_debug_print = 0;
  p->lm_init_lambda_factor = 1e-9;
  p->maha_fp64 = 0;
  p->host_loop = 0;
  p->variant = APD_VARIANT_APDGICP;
  p->voxel_search = APD_VOXEL_DIRECT1;  // fast_vgicp_impl.hpp:22-24
  p->voxel_resolution = 1.0;
  p->voxel_mode = APD_VOXEL_ADDITIVE;
}

NoiseParams noise_params(const apd_params& p) {
  NoiseParams np;
  np.dist_var = p.dist_var;
  np.sin_az = std::sin(p.azimuth_var / 180 * M_PI);    // :196
  np.sin_el = std::sin(p.elevation_var / 180 * M_PI);  // :197
  const double thr = p.max_correspondence_distance;
  np.thr_sq = thr * thr;  // :183 (double product)
  np.gicp = p.variant == APD_VARIANT_GICP ? 1 : 0;
  np.search_limit = (float)thr;
  return np;
}

int ensure_corr_buffers(apd_handle* h) {
  const size_t n = std::max<size_t>(1, h->local_n(h->src.n));
  const int fp64 = h->params.maha_fp64 ? 1 : 0;
  APD_CUDA(h, h->corr.ensure(n * sizeof(int)));
  APD_CUDA(h, h->sqd.ensure(n * sizeof(float)));
  APD_CUDA(h, h->second.ensure(n * sizeof(float)));
  APD_CUDA(h, h->mahaA.ensure(n * (fp64 ? sizeof(double2) : sizeof(float4))));
  APD_CUDA(h, h->mahaB.ensure(n * (fp64 ? 2 * sizeof(double2) : sizeof(float2))));
  h->corr_fp64 = fp64;
  return APD_OK;
}

LmConfig lm_config(const apd_params& p) {
  LmConfig c{};
  c.max_iterations = p.max_iterations;
  c.optimizer = p.optimizer == APD_OPT_GAUSS_NEWTON ? 0 : 1;
  c.lm_max_iterations = p.lm_max_iterations;
  c.maha_fp64 = p.maha_fp64 ? 1 : 0;
  c.want_fitness = 0;
  c.rotation_epsilon = p.rotation_epsilon;
  c.transformation_epsilon = p.transformation_epsilon;
  c.lm_init_lambda_factor = p.lm_init_lambda_factor;
  c.fitness_max_range = std::numeric_limits<double>::max();
  c.inlier_sq_thr = 0.0;
  c.np = noise_params(p);
  return c;
}

bool use_device_loop(const apd_handle* h) {
  return h->params.variant != APD_VARIANT_VGICP && !h->params.host_loop && !h->sharded() && !h->params.lm_debug_print && h->src.n <= kLmMaxSource;
}

LmJob lm_job(apd_handle* h, const hm::Pose& x0, int prep_bits = 0) {
  const CloudDev s = h->src.view(), t = h->tgt.view();
  LmJob j{};
  j.s_spts = s.spts; j.s_label = s.label; j.s_cov = s.cov; j.s_geo = s.geo; j.s_geo64 = s.geo64;
  j.t_spts = t.spts; j.t_label = t.label; j.t_cov = t.cov; j.t_cell_start = t.cell_start;
  j.tg = t.g;
  j.n_src = s.n;
  j.corr = h->corr.as<int>(); j.sqd = h->sqd.as<float>(); j.mahaA = h->mahaA.p; j.mahaB = h->mahaB.p;
  j.cl_w = h->params.variant == APD_VARIANT_GICP ? 0.0 : 1.0 / (double)s.n;  // 1.0 / correspondences_.size() (:273); GICP: unit weights
  for (int r = 0; r < 3; r++) {
    for (int c = 0; c < 3; c++) j.guess[r * 3 + c] = x0(r, c);
    j.guess[9 + r] = x0(r, 3);
  }
  j.result = h->lm_result.as<LmResult>();
  j.n_tgt = t.n;
  j.sg = s.g;
  j.s_ncells = s.ncells;
  j.t_ncells = t.ncells;
  if (prep_bits) {  // the kernel builds (some of) the grids and the source covariances itself
    j.prep = prep_bits;
    j.s_pts = s.pts;
    j.t_pts = t.pts;
    j.s_inv_perm = s.inv_perm;
    j.t_inv_perm = t.inv_perm;
    j.s_cell_start = s.cell_start;
    j.s_cell_cap = (int)std::min<size_t>(h->src.cell_start.cap / sizeof(uint32_t), 0x7fffffff);
    j.t_cell_cap = (int)std::min<size_t>(h->tgt.cell_start.cap / sizeof(uint32_t), 0x7fffffff);
    j.s_scratch = h->work.as<uint32_t>();
    j.t_scratch = h->work.as<uint32_t>() + 3 * (size_t)s.n;
    j.s_cells_per_point = h->cells_for(s.n);
    j.t_cells_per_point = h->cells_for(t.n, h->tgt.cov_lazy && !h->tgt.cov_valid);
    j.s_k = h->params.k_correspondences;
    j.s_reg = h->params.regularization;
    j.gicp = h->params.variant == APD_VARIANT_GICP ? 1 : 0;
    j.nb = h->nbuf.as<int32_t>();
  }
  if (h->tgt.cov_lazy && !h->tgt.cov_valid) {  // target covariances on demand
    j.t_cov_flag = h->tgt.cov_flag.as<unsigned char>();
    j.t_cov_rw = t.cov;
    j.t_pts = t.pts;
    j.nb = h->nbuf.as<int32_t>();
    j.k = h->tgt.lazy_k;
    j.reg = h->tgt.lazy_reg;
  }
  return j;
}

constexpr size_t kLmHeadBytes = offsetof(LmResult, trace) + (size_t)kLmTraceHead * 8 * sizeof(double);

// unpack a finished LmResult header (host copy) into the handle
int finish_device_align(apd_handle* h, const LmResult* r) {
  if (h->fused_inflight) {
    const int bits = h->fused_inflight;
    h->fused_inflight = 0;
    if (r->prep_status != 0) return APD_ERR_UNSUPPORTED;  // (a grid did not fit: do_align / the pool run the unfused path)
    // the kernel sized and built the grids: take their descriptors, and what is now valid on the device
    if (bits & 1) {
      h->src.g = r->grid[0];
      h->src.ncells = r->ncells[0];
      h->src.grid_valid = h->src.cov_valid = h->src.geo_valid = true;
      h->src.geo_variant = h->params.variant;
      h->src.cov_lazy = false;
    }
    if (bits & 2) {
      h->tgt.g = r->grid[1];
      h->tgt.ncells = r->ncells[1];
      h->tgt.grid_valid = true;
      h->tgt.grid_ordered = false;
    }
    for (Cloud* c : {&h->src, &h->tgt})
      if (c->grid_valid) {  // (the box was never needed on the host)
        c->host_packed = nullptr;
        if (!c->bbox_launched) c->bbox_pending = false;
      }
  }
  hm::Pose x0 = hm::Pose::identity();
  for (int a = 0; a < 3; a++) {
    for (int c = 0; c < 3; c++) x0(a, c) = r->pose[a * 3 + c];
    x0(a, 3) = r->pose[9 + a];
  }
  h->converged = r->converged != 0;
  h->nr_iterations = r->nr_iterations;
  if (r->t_end > r->t_begin) h->lm_kernel_s += 1e-9 * (double)(r->t_end - r->t_begin);
  for (int i = 0; i < 10; i++) h->lm_phase_s[i] += 1e-9 * (double)r->phase_ns[i];
  h->lm_lambda = r->lm_lambda;
  h->lm_failed = r->lm_failed != 0;
  if (r->hessian_set) std::memcpy(h->final_H, r->H, sizeof(h->final_H));  // final_hessian_ changes only when a step was accepted
  const int rows = std::min(r->n_trace, kLmTraceRows);
  h->trace.resize((size_t)rows * 8);
  const int head = std::min(rows, kLmTraceHead);
  std::memcpy(h->trace.data(), r->trace, (size_t)head * 8 * sizeof(double));
  if (rows > head) {  // rare: long LM runs — fetch the remaining rows
    APD_CUDA(h, cudaMemcpyAsync(h->trace.data() + (size_t)head * 8, h->lm_result.as<LmResult>()->trace + (size_t)head * 8,
                                (size_t)(rows - head) * 8 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    APD_CUDA(h, wait_stream(h));
  }
  if (h->lm_failed) std::fprintf(stderr, "lm not converged!!\n");  // lsq :72
  h->final_pose = x0;
  for (int a = 0; a < 4; a++)
    for (int c = 0; c < 4; c++) h->final_T[c * 4 + a] = (float)x0(a, c);  // lsq :78
  h->corr_n = h->src.n;
  h->corr_warm = false;  // (the buffers hold the loop's last linearisation pose, which the host does not track)
  return APD_OK;
}

// LsqRegistration::computeTransformation (lsq :55-80) in one launch (lm.cu)
int enqueue_device_align(apd_handle* h, const float* guess, const LmConfig& cfg, int prep_bits = 0, bool stage_only = false) {
  int rc = ensure_corr_buffers(h);
  if (rc != APD_OK) return rc;
  if (!h->lm_result.p) {
    APD_CUDA(h, h->lm_result.ensure(sizeof(LmResult)));
    APD_CUDA(h, cudaMemsetAsync(h->lm_result.p, 0, sizeof(LmResult), h->stream));
  }
  if (!h->h_lm.p) {
    APD_CUDA(h, h->h_lm.ensure(kLmHeadBytes));
    std::memset(h->h_lm.p, 0, kLmHeadBytes);
    void* d = nullptr;
    if (h->zero_copy && cudaHostGetDevicePointer(&d, h->h_lm.p, 0) == cudaSuccess) h->h_lm_dev = reinterpret_cast<LmResult*>(d);
    else h->zero_copy = false;
  }
  const hm::Pose x0 = guess ? hm::from_colmajor_f32(guess) : hm::Pose::identity();  // lsq :56
  LmJob job = lm_job(h, x0, prep_bits);
  h->fused_inflight = prep_bits;
  job.seq = ++h->seq;
  job.host_result = h->zero_copy ? h->h_lm_dev : nullptr;
  h->launch_stream = h->stream;
  if (stage_only && h->zero_copy) {  // (a pool worker launches it, alone or with a partner: launch_staged)
    h->staged_job = job;
    h->staged_cfg = cfg;
    return APD_OK;
  }
  {
    ProfScope ps(h, APD_K_LM);
    launch_lm(&job, nullptr, 1, cfg, h->lm_cluster, h->lm_min_blocks, h->stream, &h->launches);
  }
  APD_CUDA(h, cudaGetLastError());
  if (!h->zero_copy) APD_CUDA(h, cudaMemcpyAsync(h->h_lm.p, h->lm_result.p, kLmHeadBytes, cudaMemcpyDeviceToHost, h->stream));
  return APD_OK;
}

// The device-resident loop in two halves, so that a pool worker can keep several registrations in flight:
// begin_device_align enqueues covariances + the loop and returns; result_arrived polls; end_device_align unpacks.
int begin_device_align(apd_handle* h, const float* guess, bool stage_only = false) {
  const auto t_a = std::chrono::steady_clock::now();
  const double bbox_before = h->phase_s[1];
  // FastAPDGICP::computeTransformation (:148-157): the covariances — by the separate kernels, or (prep_bits) inside the loop's launch
  const int prep_bits = fused_prep_bits(h);
  int rc = prep_bits ? prepare_fused(h, prep_bits) : ensure_covariances_for_loop(h);
  if (rc != APD_OK) return rc;
  const auto t_b = std::chrono::steady_clock::now();
  h->phase_s[2] += std::chrono::duration<double>(t_b - t_a).count() - (h->phase_s[1] - bbox_before);
  h->fit_valid = false;
  LmConfig cfg = lm_config(h->params);
  if (h->fuse_fitness) {
    cfg.want_fitness = 1;
    cfg.inlier_sq_thr = h->fuse_inlier_sq_thr;
  }
  h->pending_fitness = cfg.want_fitness != 0;
  rc = enqueue_device_align(h, guess, cfg, prep_bits, stage_only);
  h->phase_s[3] += std::chrono::duration<double>(std::chrono::steady_clock::now() - t_b).count();
  return rc;
}
bool result_arrived(const apd_handle* h) {
  return *reinterpret_cast<const volatile unsigned long long*>(&reinterpret_cast<const LmResult*>(h->h_lm.p)->seq) == h->seq;
}
int end_device_align(apd_handle* h) {
  const LmResult* r = reinterpret_cast<const LmResult*>(h->h_lm.p);
  const auto t_c = std::chrono::steady_clock::now();
  if (h->zero_copy && !h->profiling) {
    APD_CUDA(h, wait_host_seq(h, &r->seq, h->seq));
    h->seq_seen = std::max(h->seq_seen, h->seq);
  } else {
    APD_CUDA(h, wait_stream(h));
  }
  h->phase_s[4] += std::chrono::duration<double>(std::chrono::steady_clock::now() - t_c).count();
  h->phase_n++;
  if (h->pending_fitness) {
    h->fit_valid = true;
    for (int i = 0; i < 3; i++) h->fit[i] = r->fitness[i];
  }
  return finish_device_align(h, r);
}

int do_align(apd_handle* h, const float* guess) {
  if (use_device_loop(h)) {
    int rc = begin_device_align(h, guess);
    if (rc == APD_OK) rc = end_device_align(h);
    if (rc == APD_ERR_UNSUPPORTED && h->fused) {  // the fused kernel declined (a grid larger than its arrays): the separate kernels
      h->fused = false;
      h->tgt.cov_lazy = h->tgt.cov_lazy && h->tgt.grid_valid;
      rc = begin_device_align(h, guess);
      if (rc == APD_OK) rc = end_device_align(h);
      h->fused = true;
    }
    return rc;
  }
  if (h->params.variant == APD_VARIANT_VGICP) {
    if (h->sharded()) return fail(h, APD_ERR_UNSUPPORTED, "FastVGICP is not sharded");
    h->vox_valid = false;  // FastVGICP::computeTransformation (fast_vgicp_impl.hpp:66-71)
  }
  int rc = ensure_covariances(h);  // FastAPDGICP::computeTransformation (:148-157)
  if (rc != APD_OK) return rc;
  h->fit_valid = false;
  hm::Pose x0 = guess ? hm::from_colmajor_f32(guess) : hm::Pose::identity();  // lsq :56
  h->lm_lambda = -1.0;   // :58
  h->converged = false;  // :59
  h->trace.clear();
  h->nr_iterations = 0;
  for (int i = 0; i < h->params.max_iterations && !h->converged; i++) {  // :67
    h->nr_iterations = i;                                                 // :68
    hm::Pose delta;
    bool ok = true;
    if (h->params.optimizer == APD_OPT_GAUSS_NEWTON) rc = step_gn(h, i, x0, delta);
    else rc = step_lm(h, i, x0, delta, &ok);
    if (rc != APD_OK) return rc;
    if (!ok) {
      std::fprintf(stderr, "lm not converged!!\n");  // :72
      break;
    }
    h->converged = hm::is_converged(delta, h->params.rotation_epsilon, h->params.transformation_epsilon);  // :75
  }
  h->final_pose = x0;
  for (int r = 0; r < 4; r++)
    for (int c = 0; c < 4; c++) h->final_T[c * 4 + r] = (float)x0(r, c);  // :78
  return APD_OK;
}

int do_fitness(apd_handle* h, const float* T, double max_range, double* score, int32_t* n_in_range, double inlier_sq_thr, int32_t* n_inliers) {
  if (!h->src.present || !h->tgt.present) return fail(h, APD_ERR_INVALID, "source or target cloud not set");
  int rc = ensure_grid(h, h->src);
  if (rc != APD_OK) return rc;
  rc = ensure_grid(h, h->tgt);
  if (rc != APD_OK) return rc;
  rc = ensure_small(h);
  if (rc != APD_OK) return rc;
  double* d_small = h->small.as<double>();
  const PoseF Tf = colmajor_f32_to_pose_f(T ? T : h->final_T);
  {
    ProfScope ps(h, APD_K_FITNESS);
    launch_fitness(h->src.view(), h->tgt.view(), Tf, max_range, inlier_sq_thr, h->partials.as<double>(), h->max_reduce_blocks, d_small + 32,
                   reinterpret_cast<unsigned int*>(d_small + 41), h->stream, &h->launches);
  }
  APD_CUDA(h, cudaGetLastError());
  double* hs = reinterpret_cast<double*>(h->h_small.p) + 32;
  APD_CUDA(h, cudaMemcpyAsync(hs, d_small + 32, 3 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  APD_CUDA(h, wait_stream(h));
  const int nr = (int)hs[1];
  if (score) *score = nr > 0 ? hs[0] / nr : std::numeric_limits<double>::max();
  if (n_in_range) *n_in_range = nr;
  if (n_inliers) *n_inliers = (int)hs[2];
  return APD_OK;
}

}  // namespace

namespace {

// A batch pool drives one CUDA stream per worker (32-64). With the driver's default of 8 hardware work queues the
// streams share queues, and a short kernel of one worker waits behind the 1 ms optimizer kernel of another (measured on
// B200, C2, 32 workers: 6.3 k registrations/s with 8 queues, 15.1 k with 32). CUDA_DEVICE_MAX_CONNECTIONS is read when the
// CUDA context is created and belongs to the whole host process (a nodelet manager has other CUDA users), so the library
// does not touch it at load time (round 1 did): the first pool / group asks for 32 queues only if the variable is unset
// AND no primary context is active yet on its device, and says so once when it is too late for that.
void ensure_work_queues() {
  static std::once_flag once;
  std::call_once(once, [] {
    if (std::getenv("CUDA_DEVICE_MAX_CONNECTIONS")) return;
    bool active = true;  // unknown -> assume the context exists
    if (void* cu = dlopen("libcuda.so.1", RTLD_NOW | RTLD_NOLOAD)) {
      typedef int (*StateFn)(int, unsigned int*, int*);
      typedef int (*InitFn)(unsigned int);
      StateFn state = (StateFn)dlsym(cu, "cuDevicePrimaryCtxGetState");
      InitFn init = (InitFn)dlsym(cu, "cuInit");
      unsigned int flags = 0;
      int act = 1;
      if (state && init && init(0) == 0) {
        active = false;
        int count = 0;
        typedef int (*CountFn)(int*);
        CountFn cnt = (CountFn)dlsym(cu, "cuDeviceGetCount");
        if (cnt) cnt(&count);
        for (int d = 0; d < count; d++)
          if (state(d, &flags, &act) == 0 && act) active = true;
      }
    } else {
      active = false;  // the driver library is not even loaded: no context can exist
    }
    if (!active) {
      setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
    } else {
      std::fprintf(stderr, "[apdgicp] note: the CUDA context already exists and CUDA_DEVICE_MAX_CONNECTIONS is unset: the batch pool's streams "
                           "share the default 8 hardware queues (export CUDA_DEVICE_MAX_CONNECTIONS=32 before the process starts for full throughput)\n");
    }
  });
}

}  // namespace

// =============================================================== C-ABI =====
extern "C" {

int apd_abi_version(void) { return APD_ABI_VERSION; }

int apd_default_params(apd_params* out) {
  if (!out) return APD_ERR_INVALID;
  default_params(out);
  return APD_OK;
}

int apd_create(int device, apd_handle** out) {
  if (!out) return APD_ERR_INVALID;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return APD_ERR_CUDA;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return APD_ERR_CUDA;
  if (prop.major != 10) return APD_ERR_CUDA;  // sm_100a only: no other code path exists
  apd_handle* h = new apd_handle();
  h->device = device;
  default_params(&h->params);
  DeviceGuard dg(device);
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete h;
    return APD_ERR_CUDA;
  }
  h->max_reduce_blocks = prop.multiProcessorCount * 4;
  if (const char* e = std::getenv("APD_CELLS_PER_POINT")) {
    const double v = std::atof(e);
    if (v > 0.01 && v < 1000.0) h->cells_per_point = v;
  }
  if (const char* e = std::getenv("APD_CELLS_PER_POINT_SMALL")) {
    const double v = std::atof(e);
    if (v > 0.01 && v < 1000.0) h->cells_per_point_small = v;
  }
  if (const char* e = std::getenv("APD_CELLS_PER_POINT_MID")) {
    const double v = std::atof(e);
    if (v > 0.01 && v < 1000.0) h->cells_per_point_mid = v;
  }
  if (const char* e = std::getenv("APD_CELLS_PER_POINT_MID_LAZY")) {
    const double v = std::atof(e);
    if (v > 0.01 && v < 1000.0) h->cells_per_point_mid_lazy = v;
  }
  if (const char* e = std::getenv("APD_LM_CLUSTER")) {
    const int v = std::atoi(e);
    if (v == 1 || v == 2 || v == 4 || v == 8 || v == 16) h->lm_cluster = v;
  }
  if (const char* e = std::getenv("APD_LM_MINB")) h->lm_min_blocks = std::atoi(e) >= 2 ? 2 : 1;
  if (const char* e = std::getenv("APD_BLOCKING_SYNC")) h->blocking_wait = std::atoi(e) != 0;
  if (const char* e = std::getenv("APD_POLL_WAIT_US")) h->poll_wait_us = std::atoi(e);
  if (const char* e = std::getenv("APD_LAZY_TARGET_COV")) h->lazy_mode = std::strcmp(e, "auto") == 0 ? -1 : (std::atoi(e) != 0 ? 1 : 0);
  if (const char* e = std::getenv("APD_ZERO_COPY")) h->zero_copy = std::atoi(e) != 0;
  if (const char* e = std::getenv("APD_FUSED")) h->fused = std::atoi(e) != 0;
  if (const char* e = std::getenv("APD_CORR_MODE")) h->corr_lanes = e[0] == 'w' ? 8 : (e[0] == 'l' ? 1 : 0);
  if (const char* e = std::getenv("APD_CORR_LANES")) {
    const int v = std::atoi(e);
    if (v == 1 || v == 2 || v == 4 || v == 8) h->corr_lanes = v;
  }
  if (const char* e = std::getenv("APD_KNN_MODE")) h->knn_mode = std::strcmp(e, "warp") == 0 ? 1 : (std::strcmp(e, "thread") == 0 ? 2 : 0);
  for (int i = 0; i < 16; i++) h->final_T[i] = (i % 5 == 0) ? 1.f : 0.f;
  for (int i = 0; i < 36; i++) h->final_H[i] = (i % 7 == 0) ? 1.0 : 0.0;  // final_hessian_.setIdentity()
  *out = h;
  return APD_OK;
}

int apd_destroy(apd_handle* h) {
  if (!h) return APD_OK;
  DeviceGuard dg(h->device);
  cudaStreamSynchronize(h->stream);
  if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
  for (int r = 0; r < kPeerMaxRanks; r++)
    if (!h->group && h->peer_box[r] && h->peer_box[r] != h->mailbox) cudaIpcCloseMemHandle(h->peer_box[r]);
  if (h->mailbox) cudaFree(h->mailbox);
  for (auto& p : h->pending) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
  for (auto e : h->event_pool) cudaEventDestroy(e);
  if (h->done_ev) cudaEventDestroy(h->done_ev);
  if (h->pair_ev) cudaEventDestroy(h->pair_ev);
  h->src.release(); h->tgt.release();
  h->corr.release(); h->sqd.release(); h->second.release(); h->mahaA.release(); h->mahaB.release();
  h->vkey.release(); h->vcnt.release(); h->vmean.release(); h->vcov.release(); h->vcorr.release(); h->vmaha.release();
  h->work.release(); h->scratch.release(); h->partials.release(); h->small.release();
  h->h_small.release();
  h->h_query.release();
  h->lm_result.release();
  h->nbuf.release();
  h->h_lm.release();
  cudaStreamDestroy(h->stream);
  delete h;
  return APD_OK;
}

const char* apd_last_error(const apd_handle* h) { return h ? h->error.c_str() : "null handle"; }

int apd_set_params(apd_handle* h, const apd_params* p) {
  if (!h || !p) return APD_ERR_INVALID;
  const bool cov_dep = p->k_correspondences != h->params.k_correspondences || p->regularization != h->params.regularization;
  if (p->regularization < APD_REG_NONE || p->regularization > APD_REG_FROBENIUS) return fail(h, APD_ERR_INVALID, "bad regularization");
  if (p->variant != APD_VARIANT_APDGICP && p->variant != APD_VARIANT_GICP && p->variant != APD_VARIANT_VGICP) return fail(h, APD_ERR_INVALID, "bad variant");
  if (p->variant == APD_VARIANT_VGICP) {
    if (!(p->voxel_resolution > 0.0) || !std::isfinite(p->voxel_resolution)) return fail(h, APD_ERR_INVALID, "bad voxel_resolution");
    if (p->voxel_search < APD_VOXEL_DIRECT27 || p->voxel_search > APD_VOXEL_DIRECT1) return fail(h, APD_ERR_INVALID, "bad voxel_search (DIRECT_RADIUS exists on the reference's VGICP_CUDA only)");
    if (p->voxel_mode < APD_VOXEL_ADDITIVE || p->voxel_mode > APD_VOXEL_MULTIPLICATIVE) return fail(h, APD_ERR_INVALID, "bad voxel_mode");
  }
  if (p->variant != h->params.variant) h->corr_warm = false;  // (the stored Mahalanobis matrices are the other variant's)
  if (p->variant != h->params.variant || p->voxel_resolution != h->params.voxel_resolution || p->voxel_mode != h->params.voxel_mode) h->vox_valid = false;
  if (p->variant != h->params.variant || p->voxel_search != h->params.voxel_search) h->vcorr_n = 0;
  h->params = *p;
  (void)cov_dep;  // reference: changing k / regularisation does NOT drop cached covariances either
  return APD_OK;
}
int apd_get_params(const apd_handle* h, apd_params* out) {
  if (!h || !out) return APD_ERR_INVALID;
  *out = h->params;
  return APD_OK;
}

int apd_set_source(apd_handle* h, const void* pts, int32_t n, int32_t stride, int32_t xyz_off, int32_t label_off, uint64_t key) {
  if (!h) return APD_ERR_INVALID;
  return set_cloud(h, h->src, pts, n, stride, xyz_off, label_off, key);
}
int apd_set_target(apd_handle* h, const void* pts, int32_t n, int32_t stride, int32_t xyz_off, int32_t label_off, uint64_t key) {
  if (!h) return APD_ERR_INVALID;
  h->vox_valid = false;  // FastVGICP::setInputTarget (fast_vgicp_impl.hpp:57-64)
  return set_cloud(h, h->tgt, pts, n, stride, xyz_off, label_off, key);
}
int apd_set_source_device(apd_handle* h, const void* d, int32_t n) {
  if (!h) return APD_ERR_INVALID;
  return set_cloud_device(h, h->src, d, n);
}
int apd_set_target_device(apd_handle* h, const void* d, int32_t n) {
  if (!h) return APD_ERR_INVALID;
  h->vox_valid = false;
  return set_cloud_device(h, h->tgt, d, n);
}

int apd_swap_source_and_target(apd_handle* h) {  // :89-98
  if (!h) return APD_ERR_INVALID;
  {
    DeviceGuard dg(h->device);
    // (the read-back slots belong to the roles, not to the clouds: a box in flight is collected before the roles change)
    const int rc = (h->src.bbox_launched || h->tgt.bbox_launched) ? finish_bboxes(h) : APD_OK;
    if (rc != APD_OK) return rc;
  }
  h->vox_valid = false;  // FastVGICP::swapSourceAndTarget (fast_vgicp_impl.hpp:46-54)
  h->vcorr_n = 0;
  std::swap(h->src, h->tgt);
  if (h->src.grid_valid && !h->src.grid_ordered) h->src.drop_derived();  // (sums over source points follow the source's layout: rebuild it in order)
  h->corr_n = -1;  // correspondences_.clear()
  h->corr_warm = false;
  return APD_OK;
}
int apd_clear_source(apd_handle* h) {  // :101-105
  if (!h) return APD_ERR_INVALID;
  h->corr_warm = false;
  h->src.present = false; h->src.n = 0; h->src.key = 0; h->src.ext_pts = nullptr;
  h->src.drop_derived();
  return APD_OK;
}
int apd_clear_target(apd_handle* h) {  // :108-112
  if (!h) return APD_ERR_INVALID;
  h->corr_warm = false;
  h->vox_valid = false;
  h->tgt.present = false; h->tgt.n = 0; h->tgt.key = 0; h->tgt.ext_pts = nullptr;
  h->tgt.drop_derived();
  return APD_OK;
}

int apd_set_source_covariances(apd_handle* h, const double* covs, int32_t n) { return h ? set_covs(h, h->src, covs, n) : APD_ERR_INVALID; }
int apd_set_target_covariances(apd_handle* h, const double* covs, int32_t n) {
  if (!h) return APD_ERR_INVALID;
  h->vox_valid = false;  // (the voxels average the target covariances)
  return set_covs(h, h->tgt, covs, n);
}
int apd_get_source_covariances(apd_handle* h, double* covs, int32_t n) { return h ? get_covs(h, h->src, covs, n) : APD_ERR_INVALID; }
int apd_get_target_covariances(apd_handle* h, double* covs, int32_t n) { return h ? get_covs(h, h->tgt, covs, n) : APD_ERR_INVALID; }

int apd_get_neighbors(apd_handle* h, int32_t which, int32_t* out, int32_t n, int32_t k) {
  if (!h || !out) return APD_ERR_INVALID;
  Cloud& c = which == 0 ? h->src : h->tgt;
  if (!c.present || n != c.n || k != h->params.k_correspondences) return fail(h, APD_ERR_INVALID, "neighbour query does not match the cloud / k");
  if (k < 1 || k > h->knn_max_k()) return fail(h, APD_ERR_UNSUPPORTED, "k_correspondences out of range (1..128; 1..32 in warp mode)");
  if (c.n < k) return fail(h, APD_ERR_TOO_FEW, "cloud has fewer points than k_correspondences");
  DeviceGuard dg(h->device);
  int rc = ensure_grid(h, c);
  if (rc != APD_OK) return rc;
  APD_CUDA(h, h->scratch.ensure((size_t)n * k * sizeof(int32_t)));
  CloudDev v = c.view();
  if (h->knn_use_warp(c.n, k)) {
    APD_CUDA(h, h->nbuf.ensure((size_t)c.n * k * sizeof(int32_t)));
    launch_knn_cov(v, k, h->nbuf.as<int32_t>(), h->scratch.as<int32_t>(), h->stream, &h->launches);
  } else {
    v.cov = nullptr;  // neighbours only
    launch_knn_cov_fused(v, k, h->params.regularization, h->scratch.as<int32_t>(), h->stream, &h->launches);
  }
  APD_CUDA(h, cudaGetLastError());
  APD_CUDA(h, cudaMemcpyAsync(out, h->scratch.p, (size_t)n * k * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
  APD_CUDA(h, wait_stream(h));
  return APD_OK;
}

int apd_align(apd_handle* h, const float* guess, float* T_out, double* T_out_f64, double* H_out, int32_t* converged,
              int32_t* iterations, float* aligned_xyz) {
  if (!h) return APD_ERR_INVALID;
  DeviceGuard dg(h->device);
  int rc = do_align(h, guess);
  if (rc != APD_OK) return rc;
  if (T_out) std::memcpy(T_out, h->final_T, sizeof(h->final_T));
  if (T_out_f64)
    for (int r = 0; r < 4; r++)
      for (int c = 0; c < 4; c++) T_out_f64[c * 4 + r] = h->final_pose(r, c);
  if (H_out) std::memcpy(H_out, h->final_H, sizeof(h->final_H));
  if (converged) *converged = h->converged ? 1 : 0;
  if (iterations) *iterations = h->nr_iterations;
  if (aligned_xyz) {  // pcl::transformPointCloud(*input_, output, final_transformation_) (lsq :79)
    const size_t bytes = (size_t)h->src.n * 3 * sizeof(float);
    APD_CUDA(h, h->scratch.ensure(bytes));
    launch_transform_cloud(h->src.view().pts, h->src.n, colmajor_f32_to_pose_f(h->final_T), h->scratch.as<float>(), h->stream, &h->launches);
    APD_CUDA(h, cudaMemcpyAsync(aligned_xyz, h->scratch.p, bytes, cudaMemcpyDeviceToHost, h->stream));
    APD_CUDA(h, wait_stream(h));
  }
  flush_prof(h);
  return APD_OK;
}

int apd_linearize(apd_handle* h, const double* T, double* H, double* b, double* err) {
  if (!h || !T || !err) return APD_ERR_INVALID;
  DeviceGuard dg(h->device);
  int rc = ensure_covariances(h);
  if (rc != APD_OK) return rc;
  double Hr[36];
  rc = do_linearize(h, hm::from_colmajor_f64(T), (H && b) ? Hr : nullptr, b, err);
  if (rc == APD_OK && H && b) std::memcpy(H, Hr, sizeof(Hr));  // symmetric: row-major == column-major
  flush_prof(h);
  return rc;
}

int apd_compute_error(apd_handle* h, const double* T, double* err) {
  if (!h || !T || !err) return APD_ERR_INVALID;
  if (h->params.variant != APD_VARIANT_VGICP && (h->corr_n != h->src.n || !h->src.present))
    return fail(h, APD_ERR_INVALID, "compute_error needs a prior linearize on the same clouds");
  DeviceGuard dg(h->device);
  int rc = reduce_pass(h, hm::from_colmajor_f64(T), false, nullptr, nullptr, err);
  flush_prof(h);
  return rc;
}

int apd_update_correspondences(apd_handle* h, const double* T) {
  if (!h || !T) return APD_ERR_INVALID;
  DeviceGuard dg(h->device);
  int rc = ensure_covariances(h);
  if (rc != APD_OK) return rc;
  rc = do_update_correspondences(h, hm::from_colmajor_f64(T));
  if (rc != APD_OK) return rc;
  APD_CUDA(h, wait_stream(h));
  flush_prof(h);
  return APD_OK;
}

int apd_get_correspondences(apd_handle* h, int32_t* idx, float* sq_dist, int32_t n) {
  if (!h) return APD_ERR_INVALID;
  if (h->corr_n != n || n != h->src.n) return fail(h, APD_ERR_INVALID, "no correspondences for this source cloud");
  DeviceGuard dg(h->device);
  const size_t ib = align_up((size_t)n * sizeof(int32_t), 256);
  APD_CUDA(h, h->scratch.ensure(2 * ib));
  int32_t* d_idx = h->scratch.as<int32_t>();
  float* d_sq = reinterpret_cast<float*>(h->scratch.as<char>() + ib);
  if (h->sharded()) {  // a sharded handle reports its own slice of the source; the other points read -1 / 0
    APD_CUDA(h, cudaMemsetAsync(d_idx, 0xff, (size_t)n * sizeof(int32_t), h->stream));
    APD_CUDA(h, cudaMemsetAsync(d_sq, 0, (size_t)n * sizeof(float), h->stream));
  }
  launch_corr_export(h->src.view(), h->tgt.view(), h->shard_table(h->src.n), corr_view(h), d_idx, d_sq, nullptr, h->stream, &h->launches);
  if (idx) APD_CUDA(h, cudaMemcpyAsync(idx, d_idx, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
  if (sq_dist) APD_CUDA(h, cudaMemcpyAsync(sq_dist, d_sq, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  APD_CUDA(h, wait_stream(h));
  return APD_OK;
}

int apd_get_mahalanobis(apd_handle* h, double* maha, int32_t n) {
  if (!h || !maha) return APD_ERR_INVALID;
  if (h->corr_n != n || n != h->src.n) return fail(h, APD_ERR_INVALID, "no correspondences for this source cloud");
  DeviceGuard dg(h->device);
  APD_CUDA(h, h->scratch.ensure((size_t)n * 16 * sizeof(double)));
  if (h->sharded()) APD_CUDA(h, cudaMemsetAsync(h->scratch.p, 0, (size_t)n * 16 * sizeof(double), h->stream));
  launch_corr_export(h->src.view(), h->tgt.view(), h->shard_table(h->src.n), corr_view(h), nullptr, nullptr, h->scratch.as<double>(), h->stream,
                     &h->launches);
  APD_CUDA(h, cudaMemcpyAsync(maha, h->scratch.p, (size_t)n * 16 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  APD_CUDA(h, wait_stream(h));
  return APD_OK;
}

int apd_vgicp_get_voxels(apd_handle* h, int32_t* n_voxels, int32_t* coords, int32_t* counts, double* means, double* covs, int32_t capacity) {
  if (!h) return APD_ERR_INVALID;
  const int nv = h->vox_valid ? h->n_vox : 0;
  if (n_voxels) *n_voxels = nv;
  const int m = std::min(nv, std::max(capacity, 0));
  if (m <= 0 || (!coords && !counts && !means && !covs)) return APD_OK;
  DeviceGuard dg(h->device);
  APD_CUDA(h, h->scratch.ensure((size_t)nv * (9 * sizeof(double) + 3 * sizeof(int32_t)) + 256));
  double* d_covs = h->scratch.as<double>();
  int32_t* d_coords = reinterpret_cast<int32_t*>(d_covs + (size_t)nv * 9);
  launch_vgicp_export_voxels(voxel_view(h), coords ? d_coords : nullptr, covs ? d_covs : nullptr, h->stream, &h->launches);
  APD_CUDA(h, cudaGetLastError());
  if (coords) APD_CUDA(h, cudaMemcpyAsync(coords, d_coords, (size_t)m * 3 * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
  if (covs) APD_CUDA(h, cudaMemcpyAsync(covs, d_covs, (size_t)m * 9 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  if (counts) APD_CUDA(h, cudaMemcpyAsync(counts, h->vcnt.p, (size_t)m * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
  if (means) APD_CUDA(h, cudaMemcpyAsync(means, h->vmean.p, (size_t)m * 3 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  APD_CUDA(h, wait_stream(h));
  return APD_OK;
}

int apd_vgicp_get_correspondences(apd_handle* h, int32_t* voxel, double* maha3x3, int32_t n_source, int32_t n_offsets) {
  if (!h) return APD_ERR_INVALID;
  if (h->vcorr_n != n_source || n_source != h->src.n || h->vcorr_noff != n_offsets || n_source <= 0)
    return fail(h, APD_ERR_INVALID, "no voxel correspondences of this shape");
  DeviceGuard dg(h->device);
  const size_t slots = (size_t)n_source * n_offsets;
  if (voxel) APD_CUDA(h, cudaMemcpyAsync(voxel, h->vcorr.p, slots * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
  if (maha3x3) {
    APD_CUDA(h, h->scratch.ensure(slots * 9 * sizeof(double)));
    launch_vgicp_export_maha(h->vcorr.as<int32_t>(), h->vmaha.as<double>(), (long long)slots, h->scratch.as<double>(), h->stream, &h->launches);
    APD_CUDA(h, cudaGetLastError());
    APD_CUDA(h, cudaMemcpyAsync(maha3x3, h->scratch.p, slots * 9 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  }
  APD_CUDA(h, wait_stream(h));
  return APD_OK;
}

int apd_fitness(apd_handle* h, const float* T, double max_range, double* score, int32_t* n_in_range, double inlier_sq_thr, int32_t* n_inliers) {
  if (!h) return APD_ERR_INVALID;
  DeviceGuard dg(h->device);
  int rc = do_fitness(h, T, max_range, score, n_in_range, inlier_sq_thr, n_inliers);
  flush_prof(h);
  return rc;
}

int apd_nearest_k(apd_handle* h, int32_t which, const void* queries, int32_t n, int32_t stride, int32_t k, int32_t* idx, float* sq_dist) {
  if (!h || !queries || n < 0 || stride < 12 || !idx || !sq_dist) return APD_ERR_INVALID;
  if (k < 1 || k > 32) return fail(h, APD_ERR_UNSUPPORTED, "apd_nearest_k serves 1 <= k <= 32");
  Cloud& c = which == 0 ? h->src : h->tgt;
  if (!c.present || c.n <= 0) return fail(h, APD_ERR_INVALID, "cloud not set");
  if (n == 0) return APD_OK;
  DeviceGuard dg(h->device);
  int rc = ensure_grid(h, c);
  if (rc != APD_OK) return rc;
  const size_t qb = align_up((size_t)n * sizeof(float4), 256), ib = align_up((size_t)n * k * sizeof(int32_t), 256);
  APD_CUDA(h, h->scratch.ensure(qb + 2 * ib));
  APD_CUDA(h, h->h_query.ensure(qb));
  float* st = reinterpret_cast<float*>(h->h_query.p);
  for (int32_t i = 0; i < n; i++) {
    std::memcpy(st + 4 * (size_t)i, reinterpret_cast<const char*>(queries) + (size_t)i * stride, 12);
    st[4 * (size_t)i + 3] = 0.f;
  }
  char* d = h->scratch.as<char>();
  APD_CUDA(h, cudaMemcpyAsync(d, st, (size_t)n * sizeof(float4), cudaMemcpyHostToDevice, h->stream));
  launch_knn_query(c.view(), k, reinterpret_cast<const float4*>(d), n, reinterpret_cast<int32_t*>(d + qb), reinterpret_cast<float*>(d + qb + ib),
                   h->stream, &h->launches);
  APD_CUDA(h, cudaGetLastError());
  APD_CUDA(h, cudaMemcpyAsync(idx, d + qb, (size_t)n * k * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
  APD_CUDA(h, cudaMemcpyAsync(sq_dist, d + qb + ib, (size_t)n * k * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  APD_CUDA(h, wait_stream(h));
  return APD_OK;
}

int apd_source_nearest(apd_handle* h, const float* T, int32_t* idx, float* sq_dist, float* xyz, int32_t n) {
  if (!h || !idx || !sq_dist) return APD_ERR_INVALID;
  if (!h->src.present || !h->tgt.present || n != h->src.n) return fail(h, APD_ERR_INVALID, "source or target cloud not set, or n is not the source's size");
  if (n == 0) return APD_OK;
  DeviceGuard dg(h->device);
  int rc = ensure_grid(h, h->src);
  if (rc != APD_OK) return rc;
  rc = ensure_grid(h, h->tgt);
  if (rc != APD_OK) return rc;
  const size_t ib = align_up((size_t)n * sizeof(int32_t), 256);
  APD_CUDA(h, h->scratch.ensure(2 * ib + (size_t)n * 3 * sizeof(float)));
  char* d = h->scratch.as<char>();
  launch_source_nearest(h->src.view(), h->tgt.view(), colmajor_f32_to_pose_f(T ? T : h->final_T), reinterpret_cast<int32_t*>(d),
                        reinterpret_cast<float*>(d + ib), xyz ? reinterpret_cast<float*>(d + 2 * ib) : nullptr, h->stream, &h->launches);
  APD_CUDA(h, cudaGetLastError());
  APD_CUDA(h, cudaMemcpyAsync(idx, d, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
  APD_CUDA(h, cudaMemcpyAsync(sq_dist, d + ib, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  if (xyz) APD_CUDA(h, cudaMemcpyAsync(xyz, d + 2 * ib, (size_t)n * 3 * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  APD_CUDA(h, wait_stream(h));
  return APD_OK;
}

// ---- the radius searches of the preprocessing stage on the cloud's grid ----
}  // extern "C"

namespace {

// counts (and, with want_lists, CSR offsets + neighbour ids in ascending row order of the ORIGINAL index) of one radius
// search per point of cloud c. mode 0: the fixed squared radius r2; 1 / 2: DBSCAN's range-dependent radii (prep_ops.cu)
int radius_csr(apd_handle* h, Cloud& c, int mode, float r2, double eps, std::vector<int32_t>& hc, std::vector<long long>* ho, std::vector<int32_t>* lists_out) {
  const int n = c.n;
  int rc = ensure_grid(h, c);
  if (rc != APD_OK) return rc;
  const size_t cb = align_up((size_t)n * sizeof(int32_t), 256), ob = align_up((size_t)n * sizeof(long long), 256);
  APD_CUDA(h, h->scratch.ensure(cb + ob));
  int32_t* d_counts = h->scratch.as<int32_t>();
  long long* d_off = reinterpret_cast<long long*>(h->scratch.as<char>() + cb);
  launch_radius_search(c.view(), r2, mode, eps, d_counts, nullptr, nullptr, h->stream, &h->launches);
  APD_CUDA(h, cudaGetLastError());
  hc.resize((size_t)n);
  APD_CUDA(h, cudaMemcpyAsync(hc.data(), d_counts, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
  APD_CUDA(h, wait_stream(h));
  if (!ho) return APD_OK;
  ho->resize((size_t)n + 1);
  (*ho)[0] = 0;
  for (int32_t i = 0; i < n; i++) (*ho)[(size_t)i + 1] = (*ho)[(size_t)i] + hc[(size_t)i];
  if (!lists_out) return APD_OK;
  const long long total = (*ho)[(size_t)n];
  lists_out->resize((size_t)total);
  if (total == 0) return APD_OK;
  DevBuf lists;
  APD_CUDA(h, lists.ensure((size_t)total * sizeof(int32_t)));
  APD_CUDA(h, cudaMemcpyAsync(d_off, ho->data(), (size_t)n * sizeof(long long), cudaMemcpyHostToDevice, h->stream));
  launch_radius_search(c.view(), r2, mode, eps, nullptr, d_off, lists.as<int32_t>(), h->stream, &h->launches);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(lists_out->data(), lists.p, (size_t)total * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = wait_stream(h);
  lists.release();
  APD_CUDA(h, e);
  return APD_OK;
}

}  // namespace

extern "C" {

int apd_radius_search(apd_handle* h, int32_t which, double radius, int32_t* counts, int64_t* offsets, int32_t* indices, int64_t capacity, int32_t n) {
  if (!h || !(radius > 0.0) || (!counts && !offsets)) return APD_ERR_INVALID;
  Cloud& c = which == 0 ? h->src : h->tgt;
  if (!c.present || n != c.n || c.n <= 0) return fail(h, APD_ERR_INVALID, "cloud not set, or n is not its size");
  DeviceGuard dg(h->device);
  const float r2 = (float)(radius * radius);  // pcl::KdTreeFLANN::radiusSearch: static_cast<float>(radius * radius)
  std::vector<int32_t> hc, lists;
  std::vector<long long> ho;
  if (offsets && indices) {  // (learn the size first: the caller's capacity must hold offsets[n])
    int rc = radius_csr(h, c, 0, r2, 0.0, hc, &ho, nullptr);
    if (rc != APD_OK) return rc;
    if (capacity < ho[(size_t)n]) return fail(h, APD_ERR_INVALID, "indices: capacity below offsets[n]");
  }
  int rc = radius_csr(h, c, 0, r2, 0.0, hc, offsets ? &ho : nullptr, (offsets && indices) ? &lists : nullptr);
  if (rc != APD_OK) return rc;
  if (counts) std::memcpy(counts, hc.data(), (size_t)n * sizeof(int32_t));
  if (offsets) for (int32_t i = 0; i <= n; i++) offsets[i] = ho[(size_t)i];
  if (offsets && indices && !lists.empty()) std::memcpy(indices, lists.data(), lists.size() * sizeof(int32_t));
  return APD_OK;
}

// DBSCANKdtreeCluster::extract (4DRadarSLAM/include/dbscan/DBSCAN_simple.h:27-104) + the cluster labels of the
// preprocessing nodelet (apps/preprocessing_nodelet_ntu.cpp:520-567): the two radius searches every point can be asked —
// as a seed and as an expansion, with their range-dependent radii — run on the GPU grid, all points at once; the growth of
// the clusters is the reference's sequential loop over those lists (its result does not depend on the order inside a
// neighbour list), then clusters are ranked by the range of their centroid and every member gets rank + 1.
int apd_dbscan_labels(apd_handle* h, int32_t which, double eps, int32_t core_min_pts, int32_t min_cluster, int32_t max_cluster, float* labels,
                      int32_t* n_clusters, int32_t n) {
  if (!h || !labels || !(eps >= 0.0)) return APD_ERR_INVALID;
  Cloud& c = which == 0 ? h->src : h->tgt;
  if (!c.present || n != c.n || c.n <= 0) return fail(h, APD_ERR_INVALID, "cloud not set, or n is not its size");
  DeviceGuard dg(h->device);
  std::vector<int32_t> cnt_seed, cnt_exp, nb_seed, nb_exp;
  std::vector<long long> off_seed, off_exp;
  int rc = radius_csr(h, c, 1, 0.f, eps, cnt_seed, &off_seed, &nb_seed);
  if (rc != APD_OK) return rc;
  rc = radius_csr(h, c, 2, 0.f, eps, cnt_exp, &off_exp, &nb_exp);
  if (rc != APD_OK) return rc;
  // the cloud itself (centroids)
  std::vector<float> pts((size_t)n * 4);
  APD_CUDA(h, cudaMemcpyAsync(pts.data(), c.view().pts, (size_t)n * sizeof(float4), cudaMemcpyDeviceToHost, h->stream));
  APD_CUDA(h, wait_stream(h));
  enum { UN_PROCESSED = 0, PROCESSING = 1, PROCESSED = 2 };
  std::vector<char> is_noise((size_t)n, 0), types((size_t)n, UN_PROCESSED);
  std::vector<std::vector<int>> clusters;
  std::vector<int> queue;
  for (int i = 0; i < n; i++) {  // :32
    if (types[(size_t)i] == PROCESSED) continue;
    if (cnt_seed[(size_t)i] < core_min_pts) {  // :39-42
      is_noise[(size_t)i] = 1;
      continue;
    }
    queue.clear();
    queue.push_back(i);
    types[(size_t)i] = PROCESSED;
    for (long long e = off_seed[(size_t)i]; e < off_seed[(size_t)i + 1]; e++) {  // :48-53 (whatever their type)
      const int j = nb_seed[(size_t)e];
      if (j != i) {
        queue.push_back(j);
        types[(size_t)j] = PROCESSING;
      }
    }
    for (size_t sq = 1; sq < queue.size(); sq++) {  // :55-81
      const int ci = queue[sq];
      if (is_noise[(size_t)ci] || types[(size_t)ci] == PROCESSED) {
        types[(size_t)ci] = PROCESSED;
        continue;
      }
      if (cnt_exp[(size_t)ci] >= core_min_pts) {
        for (long long e = off_exp[(size_t)ci]; e < off_exp[(size_t)ci + 1]; e++) {
          const int j = nb_exp[(size_t)e];
          if (types[(size_t)j] == UN_PROCESSED) {
            queue.push_back(j);
            types[(size_t)j] = PROCESSING;
          }
        }
      }
      types[(size_t)ci] = PROCESSED;
    }
    if ((long long)queue.size() >= min_cluster && (long long)queue.size() <= max_cluster) {  // :82-95
      std::vector<int> r(queue);
      std::sort(r.begin(), r.end());
      r.erase(std::unique(r.begin(), r.end()), r.end());
      clusters.push_back(std::move(r));
    }
  }
  // preprocessing_nodelet_ntu.cpp:536-567: float centroid sums in index order, range by hypot, rank ascending, label rank + 1
  // (later ranks overwrite earlier ones where clusters share a point)
  std::vector<std::pair<float, int>> order;
  for (size_t k = 0; k < clusters.size(); k++) {
    float sx = 0, sy = 0, sz = 0;
    for (int idx : clusters[k]) {
      sx += pts[(size_t)idx * 4];
      sy += pts[(size_t)idx * 4 + 1];
      sz += pts[(size_t)idx * 4 + 2];
    }
    const int np = (int)clusters[k].size();
    order.emplace_back(std::hypot(sx / np, sy / np, sz / np), (int)k);
  }
  std::stable_sort(order.begin(), order.end(), [](const std::pair<float, int>& a, const std::pair<float, int>& b) { return a.first < b.first; });
  std::fill(labels, labels + n, 0.f);
  for (size_t rank = 0; rank < order.size(); rank++)
    for (int idx : clusters[(size_t)order[rank].second]) labels[idx] = (float)(rank + 1);
  if (n_clusters) *n_clusters = (int32_t)clusters.size();
  return APD_OK;
}

}  // extern "C"

namespace {

// pcl::VoxelGrid of n device points d_in (float4 {x,y,z,label}) into d_out (capacity n): *n_out voxels in ascending voxel
// index. status_out: 0 ok, 1 = the leaf is too small for the extent (PCL passes the cloud through unchanged: so does this)
int voxel_downsample_device(apd_handle* h, const float4* d_in, int n, double leaf, float4* d_out, int* n_out, int* refused) {
  *n_out = 0;
  *refused = 0;
  if (n <= 0) return APD_OK;
  int rc = ensure_small(h);
  if (rc != APD_OK) return rc;
  unsigned int* d_state = reinterpret_cast<unsigned int*>(h->small.as<double>() + 60);  // 6 uints of scratch
  const unsigned int init[6] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u};
  APD_CUDA(h, cudaMemcpyAsync(d_state, init, sizeof(init), cudaMemcpyHostToDevice, h->stream));
  launch_voxel_bounds(d_in, n, d_state, h->stream, &h->launches);
  unsigned int enc[6];
  APD_CUDA(h, cudaMemcpyAsync(enc, d_state, sizeof(enc), cudaMemcpyDeviceToHost, h->stream));
  APD_CUDA(h, wait_stream(h));
  if (enc[0] == 0xffffffffu) return APD_OK;  // no finite point
  float mn[3], mx[3];
  for (int a = 0; a < 6; a++) {
    unsigned int u = enc[a];
    u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
    std::memcpy(a < 3 ? &mn[a] : &mx[a - 3], &u, 4);
  }
  const float inv = 1.0f / (float)leaf;
  const long long dx = (long long)((mx[0] - mn[0]) * inv) + 1, dy = (long long)((mx[1] - mn[1]) * inv) + 1, dz = (long long)((mx[2] - mn[2]) * inv) + 1;
  if (dx * dy * dz > (long long)std::numeric_limits<int32_t>::max()) {  // voxel_grid.hpp: "Leaf size is too small for the input dataset"
    APD_CUDA(h, cudaMemcpyAsync(d_out, d_in, (size_t)n * sizeof(float4), cudaMemcpyDeviceToDevice, h->stream));
    *n_out = n;
    *refused = 1;
    return APD_OK;
  }
  int min_b[3], div_b[3];
  for (int a = 0; a < 3; a++) {
    min_b[a] = (int)std::floor(mn[a] * inv);
    div_b[a] = (int)std::floor(mx[a] * inv) - min_b[a] + 1;
  }
  const int mul[3] = {1, div_b[0], div_b[0] * div_b[1]};
  // scratch: keys x2, vals x2, hist, scan tmp, heads
  const int sblocks = (n + kSortTile - 1) / kSortTile;
  const size_t hist_elems = (size_t)256 * sblocks;
  const size_t scan_elems = std::max(scan_tmp_elems_for((size_t)n + 1), scan_tmp_elems_for(hist_elems));
  const size_t kv = align_up((size_t)n * sizeof(uint32_t), 256);
  APD_CUDA(h, h->work.ensure(5 * kv + 256 + align_up(hist_elems * 4, 256) + align_up(scan_elems * 4, 256)));
  char* p = h->work.as<char>();
  uint32_t* keys[2] = {(uint32_t*)p, (uint32_t*)(p + kv)};
  uint32_t* vals[2] = {(uint32_t*)(p + 2 * kv), (uint32_t*)(p + 3 * kv)};
  uint32_t* heads = (uint32_t*)(p + 4 * kv);
  uint32_t* hist = (uint32_t*)(p + 5 * kv + 256);
  uint32_t* scan_tmp = (uint32_t*)(p + 5 * kv + 256 + align_up(hist_elems * 4, 256));
  launch_voxel_keys(d_in, n, inv, min_b, mul, keys[0], vals[0], h->stream, &h->launches);
  const int cur = launch_sort_pairs_u32(keys, vals, n, 32, hist, scan_tmp, h->stream, &h->launches);
  launch_voxel_heads(keys[cur], n, heads, h->stream, &h->launches);
  launch_exclusive_scan_u32(heads, (size_t)n + 1, scan_tmp, h->stream, &h->launches);
  launch_voxel_centroids(d_in, keys[cur], vals[cur], heads, n, d_out, h->stream, &h->launches);
  APD_CUDA(h, cudaGetLastError());
  uint32_t nv = 0;
  APD_CUDA(h, cudaMemcpyAsync(&nv, heads + n, sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
  APD_CUDA(h, wait_stream(h));
  *n_out = (int)nv;
  return APD_OK;
}

}  // namespace

extern "C" {

int apd_voxel_downsample(apd_handle* h, const void* pts, int32_t n, int32_t stride, int32_t xyz_off, int32_t label_off, double leaf, float* out_xyzl,
                         int32_t capacity, int32_t* n_out) {
  if (!h || !pts || n < 0 || stride < 12 || xyz_off < 0 || !(leaf > 0.0) || !out_xyzl || !n_out || capacity < n) return APD_ERR_INVALID;
  *n_out = 0;
  if (n == 0) return APD_OK;
  DeviceGuard dg(h->device);
  PinnedBuf stage;
  DevBuf in, out;
  int rc = APD_OK, refused = 0, nv = 0;
  cudaError_t e = stage.ensure((size_t)n * sizeof(float4));
  if (e == cudaSuccess) e = in.ensure((size_t)n * sizeof(float4));
  if (e == cudaSuccess) e = out.ensure((size_t)n * sizeof(float4));
  if (e == cudaSuccess) {
    float bbox[6];
    stage_cloud(pts, n, stride, xyz_off, label_off, reinterpret_cast<float*>(stage.p), bbox);
    e = cudaMemcpyAsync(in.p, stage.p, (size_t)n * sizeof(float4), cudaMemcpyHostToDevice, h->stream);
  }
  if (e == cudaSuccess) {
    rc = voxel_downsample_device(h, in.as<const float4>(), n, leaf, out.as<float4>(), &nv, &refused);
    if (rc == APD_OK && nv > 0) e = cudaMemcpyAsync(out_xyzl, out.p, (size_t)nv * sizeof(float4), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = wait_stream(h);
  }
  stage.release(); in.release(); out.release();
  if (e != cudaSuccess) {
    h->error = cudaGetErrorString(e);
    return APD_ERR_CUDA;
  }
  if (rc == APD_OK) *n_out = nv;
  return rc;
}

int apd_submap_assemble(apd_handle* h, const apd_cloud_ref* clouds, const double* poses, int32_t n_clouds, int32_t stride, int32_t xyz_off,
                        int32_t label_off, double leaf, int32_t set_as_target, float* out_xyzl, int32_t capacity, int32_t* n_out) {
  if (!h || !clouds || !poses || n_clouds < 1 || stride < 12 || xyz_off < 0 || !n_out) return APD_ERR_INVALID;
  long long total = 0;
  for (int c = 0; c < n_clouds; c++) {
    if (!clouds[c].pts || clouds[c].n < 0) return APD_ERR_INVALID;
    total += clouds[c].n;
  }
  if (total > 0x7fffffffll || (out_xyzl && capacity < total)) return fail(h, APD_ERR_INVALID, "submap: capacity below the sum of the keyframe sizes");
  *n_out = 0;
  if (total == 0) return APD_OK;
  DeviceGuard dg(h->device);
  const int n = (int)total;
  PinnedBuf stage;
  DevBuf raw, moved, vox;
  cudaError_t e = stage.ensure((size_t)n * sizeof(float4));
  if (e == cudaSuccess) e = raw.ensure((size_t)n * sizeof(float4));
  if (e == cudaSuccess) e = moved.ensure((size_t)n * sizeof(float4));
  int rc = APD_OK, nv = n, refused = 0;
  if (e == cudaSuccess) {
    size_t off = 0;
    for (int c = 0; c < n_clouds; c++) {  // gather every keyframe into one pinned buffer, one copy, then one transform launch per keyframe
      float bbox[6];
      stage_cloud(clouds[c].pts, clouds[c].n, stride, xyz_off, label_off, reinterpret_cast<float*>(stage.p) + 4 * off, bbox);
      off += (size_t)clouds[c].n;
    }
    e = cudaMemcpyAsync(raw.p, stage.p, (size_t)n * sizeof(float4), cudaMemcpyHostToDevice, h->stream);
    off = 0;
    for (int c = 0; c < n_clouds && e == cudaSuccess; c++) {
      launch_transform_cloud_d(raw.as<const float4>() + off, clouds[c].n, poses + 16 * (size_t)c, moved.as<float4>() + off, h->stream, &h->launches);
      off += (size_t)clouds[c].n;
    }
    if (e == cudaSuccess) e = cudaGetLastError();
  }
  const float4* result = moved.as<const float4>();
  if (e == cudaSuccess && leaf > 0.0) {
    e = vox.ensure((size_t)n * sizeof(float4));
    if (e == cudaSuccess) {
      rc = voxel_downsample_device(h, moved.as<const float4>(), n, leaf, vox.as<float4>(), &nv, &refused);
      result = vox.as<const float4>();
    }
  }
  if (e == cudaSuccess && rc == APD_OK && out_xyzl && nv > 0) e = cudaMemcpyAsync(out_xyzl, result, (size_t)nv * sizeof(float4), cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess && rc == APD_OK && set_as_target && nv > 0) {
    // registration_s2m->setInputTarget(submap): the submap never leaves the device
    Cloud& t = h->tgt;
    e = t.pts.ensure((size_t)nv * sizeof(float4));
    if (e == cudaSuccess) e = cudaMemcpyAsync(t.pts.p, result, (size_t)nv * sizeof(float4), cudaMemcpyDeviceToDevice, h->stream);
    if (e == cudaSuccess) {
      t.ext_pts = nullptr;
      t.n = nv;
      t.present = true;
      t.key = 0;
      t.print = 0;
      t.drop_derived();
      t.bbox_known = false;
      t.bbox_pending = true;  // (its box is reduced on the device when a grid is sized on the host)
      h->corr_warm = false;
    }
  }
  if (e == cudaSuccess) e = wait_stream(h);
  stage.release(); raw.release(); moved.release(); vox.release();
  if (e != cudaSuccess) {
    h->error = cudaGetErrorString(e);
    return APD_ERR_CUDA;
  }
  if (rc == APD_OK) *n_out = nv;
  return rc;
}

int apd_get_lm_trace(apd_handle* h, double* rows, int32_t max_rows, int32_t* n_rows) {
  if (!h || !rows || !n_rows) return APD_ERR_INVALID;
  const int n = std::min<int>(max_rows, (int)(h->trace.size() / 8));
  std::memcpy(rows, h->trace.data(), (size_t)n * 8 * sizeof(double));
  *n_rows = n;
  return APD_OK;
}

// ---- batched registrations ---------------------------------------------------
}  // extern "C"

// A pool of workers, one handle (CUDA stream) and one host thread each. Pairs are independent (reference
// loop_detector.cpp:222-236 runs them serially), so a worker takes the next pair, stages and copies its clouds,
// enqueues grid build + covariances + the device-resident optimizer loop (+ the fitness pass) and waits for
// the one result copy; the other workers' kernels and copies fill the GPU meanwhile.
struct apd_batch {
  int device = 0;                // the first device
  std::vector<int> devices;      // one pool of workers per device; ALL of them take pairs from the one counter `next`
  std::vector<apd_handle*> handles;  // device d owns handles [d * per_device, (d + 1) * per_device)
  int per_device = 0;
  std::vector<std::thread> threads;
  int n_threads = 0;             // host threads per device
  std::vector<int64_t> pairs_by_device;  // how many pairs of the last call each device took (apd_batch_device_pairs)
  int reserved_s = 0, reserved_t = 0;    // cloud sizes the workers' buffers have been sized for
  bool reserved_host = false;            // ... including the pinned staging buffers of host clouds
  std::mutex mu;
  std::condition_variable cv_work, cv_done;
  uint64_t generation = 0;
  int pending = 0;
  bool stop = false;
  // the current call
  const apd_pair* pairs = nullptr;
  apd_result* results = nullptr;
  int n_pairs = 0, stride = 0, xyz_off = 0, label_off = 0, with_fitness = 0;
  bool device_clouds = false;
  std::atomic<int> next{0};
  // Two registrations per launch. A device runs at most 128 grids at a time; the pool's build of the loop kernel has 148
  // cluster slots (2 CTAs per SM, clusters of 2), so with one registration per launch 20 of them stay empty (CTA-slot
  // occupancy 0.86 = 128 / 148, measured). One spot per device: the handle whose loop is ready and waits for a partner.
  struct PairSpot {
    std::mutex mu;
    apd_handle* h = nullptr;
  };
  std::vector<std::unique_ptr<PairSpot>> spots;
  int pair_hold_us = 200;  // how long a ready registration waits for a partner (APD_PAIR_HOLD_US; 0: one per launch)
};

namespace {

// launch the staged loop of `h` — with the staged loop of `partner` in the same grid when there is one — on h's stream
int launch_staged(apd_handle* h, apd_handle* partner) {
  DeviceGuard dg(h->device);
  cudaError_t e = cudaSuccess;
  if (partner) {
    e = cudaStreamWaitEvent(h->stream, partner->pair_ev, 0);  // the partner's cloud copies
    partner->launch_stream = h->stream;
  }
  if (e == cudaSuccess) {
    launch_lm(&h->staged_job, nullptr, 1, h->staged_cfg, h->lm_cluster, h->lm_min_blocks, h->stream, &h->launches, partner ? &partner->staged_job : nullptr);
    e = cudaGetLastError();
    h->k_launches[APD_K_LM]++;
  }
  // (whatever happened, the partner leaves state 2: its own stall check then finds an idle stream with nothing published)
  if (partner) partner->pair_state.store(3, std::memory_order_release);
  h->pair_state.store(0, std::memory_order_release);
  APD_CUDA(h, e);
  return APD_OK;
}

// can the staged loops of a and b share a launch? (same kernel build, cluster size and configuration — always, within a batch)
bool pairable(const apd_handle* a, const apd_handle* b) {
  return a->device == b->device && a->lm_cluster == b->lm_cluster && a->lm_min_blocks == b->lm_min_blocks &&
         std::memcmp(&a->staged_cfg, &b->staged_cfg, sizeof(LmConfig)) == 0;
}

// result of pair i from the handle that has just finished it (rc: the status so far)
void batch_fill_result(apd_batch* b, apd_handle* h, int i, int rc) {
  apd_result& r = b->results[i];
  if (rc == APD_OK) {
    std::memcpy(r.T, h->final_T, sizeof(h->final_T));
    r.converged = h->converged ? 1 : 0;
    r.iterations = h->nr_iterations;
    if (b->with_fitness) {
      if (h->fit_valid) {  // the device loop already ran the pass on the final pose
        const int nr = (int)h->fit[1];
        r.fitness = nr > 0 ? h->fit[0] / nr : std::numeric_limits<double>::max();
        r.n_inliers = (int)h->fit[2];
      } else {
        rc = apd_fitness(h, nullptr, DBL_MAX, &r.fitness, nullptr, 0.25, &r.n_inliers);
      }
    }
  }
  r.status = rc;
  // a failed pair may leave copies out of the caller's buffers in flight: they must not outlive apd_batch_align
  if (rc != APD_OK) cudaStreamSynchronize(h->stream);
  if (rc == APD_ERR_CUDA || rc == APD_ERR_INVALID)  // (apd_result carries the code only: say what it was)
    std::fprintf(stderr, "[apdgicp] batch pair %d failed (status %d): %s\n", i, rc, h->error.c_str());
}

// A pool worker drives several handles ("slots") and never blocks on one of them: a registration is a short state
// machine — set the clouds (device clouds: launch the bounding-box kernels), wait for the boxes, enqueue grids +
// covariances + the device-resident loop, wait for the result — and the waits are looks at sequence numbers the kernels
// publish into pinned host memory (apd_handle::zero_copy). Few threads issue all the driver calls: with one thread per
// registration in flight (32-64 threads) the calls contend for the driver's context lock and cost ~50 us each.
enum { kSlotIdle = 0, kSlotBbox, kSlotResult, kSlotDone };
struct PoolSlot {
  apd_handle* h = nullptr;
  int pair = -1;
  int state = kSlotIdle;
  int di = 0;  // index of the handle's device in the batch
  std::chrono::steady_clock::time_point since;
  int idle_seen = 0;  // consecutive stall checks that found the stream idle and nothing published (slot_stalled)
};

bool bboxes_arrived(const apd_handle* h) {
  for (const Cloud* c : {&h->src, &h->tgt}) {
    if (!c->bbox_pending) continue;
    const unsigned int* enc = reinterpret_cast<const unsigned int*>(reinterpret_cast<const double*>(h->h_small.p) + (c == &h->src ? 48 : 52));
    if (*reinterpret_cast<const volatile unsigned long long*>(enc + 6) != c->bbox_seq) return false;
  }
  return true;
}

// enqueue the registration of the slot's pair (the clouds are set, their boxes have arrived)
void slot_enqueue(apd_batch* b, PoolSlot& sl) {
  apd_handle* h = sl.h;
  const bool pairing = b->pair_hold_us > 0 && h->zero_copy;
  int rc = begin_device_align(h, b->pairs[sl.pair].guess, pairing);
  if (rc == APD_OK && pairing && h->zero_copy) {
    // ready to launch: with the registration waiting in the device's spot, or wait there for the next one
    apd_batch::PairSpot& spot = *b->spots[(size_t)sl.di];
    apd_handle* partner = nullptr;
    bool waiting = false;
    {
      std::lock_guard<std::mutex> lk(spot.mu);
      if (spot.h && pairable(h, spot.h)) {
        partner = spot.h;
        spot.h = nullptr;
        partner->pair_state.store(2, std::memory_order_release);
      } else if (!spot.h) {
        if (!h->pair_ev) rc = cudaEventCreateWithFlags(&h->pair_ev, cudaEventDisableTiming) == cudaSuccess ? APD_OK : APD_ERR_CUDA;
        if (rc == APD_OK) rc = cudaEventRecord(h->pair_ev, h->stream) == cudaSuccess ? APD_OK : APD_ERR_CUDA;
        if (rc == APD_OK) {
          h->pair_state.store(1, std::memory_order_release);
          spot.h = h;
          waiting = true;
        }
      }
    }
    if (rc == APD_OK && !waiting) rc = launch_staged(h, partner);
  }
  if (rc == APD_OK && !h->zero_copy) rc = end_device_align(h);  // (no zero-copy mapping after all: finish it here)
  else if (rc == APD_OK) {
    sl.state = kSlotResult;
    sl.since = std::chrono::steady_clock::now();
    sl.idle_seen = 0;
    return;
  }
  batch_fill_result(b, h, sl.pair, rc);
  sl.state = kSlotIdle;
}

void slot_begin(apd_batch* b, PoolSlot& sl, int i) {
  apd_handle* h = sl.h;
  sl.pair = i;
  const apd_pair& pr = b->pairs[i];
  std::memset(&b->results[i], 0, sizeof(apd_result));
  // A registration whose loop a partner launched ends with pair_state 3 and launch_stream = the partner's stream. The next
  // one must not inherit them: while it waits for its bounding boxes (device clouds, grids sized on the host — the eager
  // path) the stall check would ask the PARTNER's old stream, find it idle and report a kernel that "finished without
  // publishing its result" (round 2: ~1 pair in a thousand with 256 eager registrations in flight, where a bounds kernel
  // queues for more than the check's 20 ms).
  h->pair_state.store(0, std::memory_order_release);
  h->launch_stream = h->stream;
  const auto t_set = std::chrono::steady_clock::now();
  // the reference benchmark protocol: clearTarget; clearSource; setInputTarget; setInputSource; align (align.cpp:57-83)
  apd_clear_target(h);
  apd_clear_source(h);
  int rc;
  if (b->device_clouds) {
    rc = apd_set_target_device(h, pr.target, pr.n_target);
    if (rc == APD_OK) rc = apd_set_source_device(h, pr.source, pr.n_source);
  } else {
    rc = apd_set_target(h, pr.target, pr.n_target, b->stride, b->xyz_off, b->label_off, 0);
    if (rc == APD_OK) rc = apd_set_source(h, pr.source, pr.n_source, b->stride, b->xyz_off, b->label_off, 0);
  }
  h->phase_s[0] += std::chrono::duration<double>(std::chrono::steady_clock::now() - t_set).count();
  h->fuse_fitness = b->with_fitness != 0;
  if (rc != APD_OK) {
    batch_fill_result(b, h, i, rc);
    sl.state = kSlotIdle;
    return;
  }
  if (!use_device_loop(h) || !h->zero_copy || h->profiling) {  // host-driven loop / copies + stream waits / timing events: blocking
    rc = apd_align(h, pr.guess, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
    batch_fill_result(b, h, i, rc);
    sl.state = kSlotIdle;
    return;
  }
  DeviceGuard dg(h->device);
  if (fused_prep_bits(h) == 0 && (h->src.bbox_pending || h->tgt.bbox_pending)) {
    // the grids are sized on the host: reduce the boxes of the device clouds and come back when they have arrived
    for (Cloud* c : {&h->src, &h->tgt})
      if (c->bbox_pending && !c->bbox_launched) rc = rc != APD_OK ? rc : launch_bounds_of(h, *c);
    if (rc != APD_OK) {
      batch_fill_result(b, h, i, rc);
      sl.state = kSlotIdle;
      return;
    }
    sl.state = kSlotBbox;
    sl.since = std::chrono::steady_clock::now();
    sl.idle_seen = 0;
    return;
  }
  slot_enqueue(b, sl);  // (the fused kernel sizes its grids itself: one launch, no wait for the boxes)
}

// a slot has been waiting for 20 ms: make sure its stream is still alive (a failed launch or a faulting kernel publishes nothing).
// `arrived` is asked AFTER the stream was seen idle: a kernel that finishes between a first look at the result and the
// stream query has published it (round 2: one pair in a thousand of a pool whose registrations take longer than the
// 20 ms — 256 eager ones in flight — was reported as failed this way).
bool slot_stalled(PoolSlot& sl, bool (*arrived)(const apd_handle*)) {
  const auto now = std::chrono::steady_clock::now();
  if (now - sl.since < std::chrono::milliseconds(20)) return false;
  sl.since = now;
  const int ps = sl.h->pair_state.load(std::memory_order_acquire);
  if (ps == 1 || ps == 2) return false;  // not launched yet (waiting for a partner / the partner is launching it)
  const cudaError_t e = cudaStreamQuery(ps == 3 ? sl.h->launch_stream : sl.h->stream);
  if (e == cudaErrorNotReady) {
    sl.idle_seen = 0;
    return false;
  }
  if (e == cudaSuccess && arrived(sl.h)) return false;
  // an idle stream with nothing published is a verdict only when the next check (20 ms later) finds the same: the state a
  // worker reads here is written by other workers (a partner launching this handle's loop), and a false alarm fails a pair
  if (e == cudaSuccess && ++sl.idle_seen < 2) return false;
  sl.h->error = e == cudaSuccess ? "kernel finished without publishing its result" : cudaGetErrorString(e);
  return true;
}

// Host threads that feed a GPU run on the cores of the GPU's NUMA node (staging a 3 MB PCL cloud is a memory-bound loop,
// and the pinned staging buffers are first touched by these threads): best effort from sysfs, within the process's own
// affinity mask; APD_NUMA_PIN=0 leaves the threads where the scheduler puts them.
void pin_to_gpu_node(int device) {
  if (const char* e = std::getenv("APD_NUMA_PIN"))
    if (std::atoi(e) == 0) return;
  char bus[32] = {0};
  if (cudaDeviceGetPCIBusId(bus, sizeof(bus), device) != cudaSuccess) return;
  for (char* p = bus; *p; p++) *p = (char)std::tolower(*p);
  char path[128];
  std::snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/numa_node", bus);
  int node = -1;
  if (FILE* f = std::fopen(path, "r")) {
    if (std::fscanf(f, "%d", &node) != 1) node = -1;
    std::fclose(f);
  }
  if (node < 0) return;
  std::snprintf(path, sizeof(path), "/sys/devices/system/node/node%d/cpulist", node);
  FILE* f = std::fopen(path, "r");
  if (!f) return;
  char list[4096] = {0};
  const bool ok = std::fgets(list, sizeof(list), f) != nullptr;
  std::fclose(f);
  if (!ok) return;
  cpu_set_t mine, want;
  if (sched_getaffinity(0, sizeof(mine), &mine) != 0) return;
  CPU_ZERO(&want);
  for (char* tok = std::strtok(list, ",\n"); tok; tok = std::strtok(nullptr, ",\n")) {
    int a = 0, b2 = 0;
    const int k = std::sscanf(tok, "%d-%d", &a, &b2);
    if (k == 1) b2 = a;
    if (k >= 1)
      for (int c = a; c <= b2 && c < CPU_SETSIZE; c++)
        if (CPU_ISSET(c, &mine)) CPU_SET(c, &want);
  }
  if (CPU_COUNT(&want) > 0) sched_setaffinity(0, sizeof(want), &want);
}

void batch_worker(apd_batch* b, int wi) {
  // thread wi serves device wi / n_threads: every n_threads-th handle of that device
  const int di = wi / b->n_threads, ti = wi % b->n_threads;
  cudaSetDevice(b->devices[(size_t)di]);
  pin_to_gpu_node(b->devices[(size_t)di]);
  std::vector<PoolSlot> slots;
  for (int s = ti; s < b->per_device; s += b->n_threads) {
    PoolSlot sl;
    sl.h = b->handles[(size_t)di * b->per_device + s];
    sl.di = di;
    slots.push_back(sl);
  }
  int64_t taken = 0;
  const int nap_us = std::max(1, slots[0].h->poll_wait_us);
  uint64_t seen = 0;
  for (;;) {
    {
      std::unique_lock<std::mutex> lk(b->mu);
      b->cv_work.wait(lk, [&] { return b->stop || b->generation != seen; });
      if (b->stop) return;
      seen = b->generation;
    }
    for (auto& sl : slots) sl.state = kSlotIdle;
    for (size_t done = 0; done < slots.size();) {
      bool progressed = false;
      done = 0;
      for (auto& sl : slots) {
        apd_handle* h = sl.h;
        switch (sl.state) {
          case kSlotIdle: {
            const int i = b->next.fetch_add(1);  // ONE queue for all devices: whoever is free takes the next pair
            if (i >= b->n_pairs) {
              sl.state = kSlotDone;
            } else {
              slot_begin(b, sl, i);
              taken++;
              progressed = true;
            }
            break;
          }
          case kSlotBbox:
            if (bboxes_arrived(h)) {
              DeviceGuard dg(h->device);
              slot_enqueue(b, sl);
              progressed = true;
            } else if (slot_stalled(sl, bboxes_arrived)) {
              batch_fill_result(b, h, sl.pair, APD_ERR_CUDA);
              sl.state = kSlotIdle;
            }
            break;
          case kSlotResult:
            if (h->pair_state.load(std::memory_order_acquire) == 1 &&
                std::chrono::steady_clock::now() - sl.since > std::chrono::microseconds(b->pair_hold_us)) {
              // no partner came: launch it alone (unless one is taking it right now)
              apd_batch::PairSpot& spot = *b->spots[(size_t)sl.di];
              bool mine = false;
              {
                std::lock_guard<std::mutex> lk(spot.mu);
                if (spot.h == h) {
                  spot.h = nullptr;
                  mine = true;
                }
              }
              if (mine) {
                const int rc = launch_staged(h, nullptr);
                progressed = true;
                if (rc != APD_OK) {
                  batch_fill_result(b, h, sl.pair, rc);
                  sl.state = kSlotIdle;
                  break;
                }
              }
            }
            if (result_arrived(h)) {
              DeviceGuard dg(h->device);
              int rc = end_device_align(h);
              if (rc == APD_ERR_UNSUPPORTED && h->fused) {  // the fused kernel declined this pair (a grid larger than its arrays)
                h->fused = false;
                rc = apd_align(h, b->pairs[sl.pair].guess, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
                h->fused = true;
              }
              batch_fill_result(b, h, sl.pair, rc);
              sl.state = kSlotIdle;
              progressed = true;
            } else if (slot_stalled(sl, result_arrived)) {
              batch_fill_result(b, h, sl.pair, APD_ERR_CUDA);
              sl.state = kSlotIdle;
            }
            break;
          default:
            done++;
            break;
        }
      }
      if (!progressed && done < slots.size()) std::this_thread::sleep_for(std::chrono::microseconds(nap_us));
    }
    {
      std::lock_guard<std::mutex> lk(b->mu);
      b->pairs_by_device[(size_t)di] += taken;
      taken = 0;
      if (--b->pending == 0) b->cv_done.notify_all();
    }
  }
}

// Sizes every buffer a pooled registration of an (ns, nt)-point pair will touch. A buffer that grows later does it with
// cudaFree + cudaMalloc, which synchronise the whole device — i.e. every other registration in flight; with a few pairs
// per worker and clouds of varying size that was what a short batch spent its time on (bench.py --gpus 2: 9.9 k
// registrations/s in the first timed arm, 46 k in the third).
int reserve_pool_buffers(apd_handle* h, int ns, int nt, bool host_clouds) {
  DeviceGuard dg(h->device);
  Cloud& s = h->src;
  Cloud& t = h->tgt;
  const int k = h->params.k_correspondences;
  const size_t S = (size_t)std::max(ns, 1), T = (size_t)std::max(nt, 1);
  APD_CUDA(h, s.pts.ensure(S * sizeof(float4)));
  APD_CUDA(h, t.pts.ensure(T * sizeof(float4)));
  if (host_clouds) {
    APD_CUDA(h, s.stage.ensure(S * sizeof(float4)));
    APD_CUDA(h, t.stage.ensure(T * sizeof(float4)));
  }
  for (int c = 0; c < 2; c++) {
    Cloud& cl = c == 0 ? s : t;
    const size_t n = c == 0 ? S : T;
    APD_CUDA(h, cl.spts.ensure(n * sizeof(float4)));
    APD_CUDA(h, cl.label.ensure(n * sizeof(float)));
    APD_CUDA(h, cl.inv_perm.ensure(n * sizeof(int)));
    APD_CUDA(h, cl.cell_start.ensure((size_t)grid_cell_capacity((int)n, h->max_cells_for((int)n)) * sizeof(uint32_t)));
    APD_CUDA(h, cl.cov.ensure(n * 6 * sizeof(double)));
  }
  APD_CUDA(h, s.geo.ensure(S * sizeof(float)));
  APD_CUDA(h, s.geo64.ensure(S * sizeof(double)));
  APD_CUDA(h, t.cov_flag.ensure(T));
  APD_CUDA(h, h->work.ensure(3 * (S + T) * sizeof(uint32_t) + 512));
  APD_CUDA(h, h->nbuf.ensure(S * (size_t)(std::max(k, 1) + 1) * sizeof(int32_t)));
  APD_CUDA(h, h->corr.ensure(S * sizeof(int)));
  APD_CUDA(h, h->sqd.ensure(S * sizeof(float)));
  APD_CUDA(h, h->second.ensure(S * sizeof(float)));
  APD_CUDA(h, h->mahaA.ensure(S * sizeof(double2)));
  APD_CUDA(h, h->mahaB.ensure(S * 2 * sizeof(double2)));
  return APD_OK;
}

int batch_run(apd_batch* b, const apd_pair* pairs, int32_t n_pairs, int32_t stride, int32_t xyz_off, int32_t label_off, bool device_clouds,
              int32_t with_fitness, apd_result* results) {
  if (!b || !pairs || !results || n_pairs < 0) return APD_ERR_INVALID;
  if (n_pairs == 0) return APD_OK;
  if (device_clouds && b->devices.size() > 1) return APD_ERR_UNSUPPORTED;  // (a device pointer names ONE device's memory)
  {
    int max_s = 0, max_t = 0;
    for (int i = 0; i < n_pairs; i++) {
      max_s = std::max(max_s, pairs[i].n_source);
      max_t = std::max(max_t, pairs[i].n_target);
    }
    if (max_s > b->reserved_s || max_t > b->reserved_t || (!device_clouds && !b->reserved_host)) {  // (before any registration is in flight)
      b->reserved_s = std::max(b->reserved_s, max_s + max_s / 8);
      b->reserved_t = std::max(b->reserved_t, max_t + max_t / 8);
      b->reserved_host = b->reserved_host || !device_clouds;
      for (auto* h : b->handles) (void)reserve_pool_buffers(h, b->reserved_s, b->reserved_t, !device_clouds);  // (best effort: set_cloud reports a real failure)
    }
  }
  {
    std::lock_guard<std::mutex> lk(b->mu);
    std::fill(b->pairs_by_device.begin(), b->pairs_by_device.end(), 0);
    b->pairs = pairs;
    b->results = results;
    b->n_pairs = n_pairs;
    b->stride = stride;
    b->xyz_off = xyz_off;
    b->label_off = label_off;
    b->with_fitness = with_fitness;
    b->device_clouds = device_clouds;
    b->next.store(0);
    b->pending = (int)b->threads.size();
    b->generation++;
  }
  b->cv_work.notify_all();
  std::unique_lock<std::mutex> lk(b->mu);
  b->cv_done.wait(lk, [&] { return b->pending == 0; });
  return APD_OK;
}

}  // namespace

extern "C" {

int apd_batch_create_multi(const int32_t* devices, int32_t n_devices, int32_t n_workers, apd_batch** out) {
  if (!out || !devices || n_devices < 1 || n_devices > 64) return APD_ERR_INVALID;
  *out = nullptr;
  if (n_workers < 1) n_workers = 1;
  if (n_workers > 512) n_workers = 512;
  ensure_work_queues();
  apd_batch* b = new apd_batch();
  b->device = devices[0];
  b->devices.assign(devices, devices + n_devices);
  b->per_device = n_workers;
  b->pairs_by_device.assign((size_t)n_devices, 0);
  for (int d = 0; d < n_devices; d++) b->spots.emplace_back(new apd_batch::PairSpot());
  if (const char* e = std::getenv("APD_PAIR_HOLD_US")) b->pair_hold_us = std::max(0, std::atoi(e));
  for (int d = 0; d < n_devices; d++)
    for (int s = 0; s < n_workers; s++) {
      apd_handle* h = nullptr;
      const int rc = apd_create(devices[d], &h);
      if (rc != APD_OK) {
        for (auto* hh : b->handles) apd_destroy(hh);
        delete b;
        return rc;
      }
      // a pool this large keeps the GPU busy by itself: its waiting threads sleep instead of spinning, so that several
      // ranks' pools can share the host cores (see wait_stream)
      h->pooled = true;
      if (!std::getenv("APD_LM_CLUSTER")) h->lm_cluster = 2;
      if (!std::getenv("APD_LM_MINB")) h->lm_min_blocks = 2;
      if (!std::getenv("APD_BLOCKING_SYNC")) h->blocking_wait = n_workers > 8;
      if (!std::getenv("APD_POLL_WAIT_US")) h->poll_wait_us = 50;  // (a look at pinned host memory: cheap)
      b->handles.push_back(h);
    }
  // Host threads per device: each drives n_workers / n_threads registrations at a time. Default: the host cores this
  // process may use divided by the GPUs sharing them (the devices of this context x the ranks on the node — one process
  // per GPU: LOCAL_WORLD_SIZE as torchrun exports it), between 2 and 16 (staging the 48-byte PCL layout is what the
  // threads are for: 0.3 ms of a core per 60 k-point target). More threads than cores is what to avoid: a
  // thread that holds the CUDA driver's lock and loses its core stalls every other thread of the process (8 x B200 on 32
  // cores, 8 threads per rank: host CPU per registration 0.08 -> 0.29 ms device-resident, 0.18 -> 0.63 ms end to end).
  // APD_BATCH_THREADS overrides.
  int n_threads = 8;
  {
    int cores = (int)std::thread::hardware_concurrency();
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof(set), &set) == 0) cores = CPU_COUNT(&set);
    int ranks = 0;
    if (const char* e = std::getenv("LOCAL_WORLD_SIZE")) ranks = std::atoi(e);
    if (ranks < 1) ranks = 1;
    ranks *= n_devices;
    if (cores > 0 && ranks > 0) n_threads = std::max(2, std::min(16, cores / ranks));
  }
  if (const char* e = std::getenv("APD_BATCH_THREADS")) n_threads = std::atoi(e);
  n_threads = std::max(1, std::min(n_threads, n_workers));
  b->n_threads = n_threads;
  for (int s = 0; s < n_threads * n_devices; s++) b->threads.emplace_back(batch_worker, b, s);
  *out = b;
  return APD_OK;
}

int apd_batch_create(int device, int32_t n_workers, apd_batch** out) {
  const int32_t dev = device;
  return apd_batch_create_multi(&dev, 1, n_workers, out);
}

int apd_batch_device_pairs(const apd_batch* b, int64_t* pairs, int32_t n_devices) {
  if (!b || !pairs || n_devices != (int32_t)b->devices.size()) return APD_ERR_INVALID;
  for (int d = 0; d < n_devices; d++) pairs[d] = b->pairs_by_device[(size_t)d];
  return APD_OK;
}

int apd_batch_destroy(apd_batch* b) {
  if (!b) return APD_OK;
  {
    std::lock_guard<std::mutex> lk(b->mu);
    b->stop = true;
  }
  b->cv_work.notify_all();
  for (auto& t : b->threads) t.join();
  if (const char* e = std::getenv("APD_BATCH_TRACE")) {
    if (std::atoi(e) != 0) {
      double ph[5] = {0, 0, 0, 0, 0};
      int64_t n = 0;
      for (auto* h : b->handles) {
        for (int i = 0; i < 5; i++) ph[i] += h->phase_s[i];
        n += h->phase_n;
      }
      if (n > 0)
        std::fprintf(stderr, "apd_batch trace: %lld registrations, host ms per registration: set %.3f, bbox wait %.3f, enqueue grids+cov %.3f, "
                     "enqueue loop %.3f, result wait %.3f\n", (long long)n, 1e3 * ph[0] / n, 1e3 * ph[1] / n, 1e3 * ph[2] / n, 1e3 * ph[3] / n,
                     1e3 * ph[4] / n);
    }
  }
  for (auto* h : b->handles) apd_destroy(h);
  delete b;
  return APD_OK;
}

int apd_batch_set_params(apd_batch* b, const apd_params* p) {
  if (!b || !p) return APD_ERR_INVALID;
  for (auto* h : b->handles) {
    const int rc = apd_set_params(h, p);
    if (rc != APD_OK) return rc;
  }
  return APD_OK;
}

int apd_batch_align(apd_batch* b, const apd_pair* pairs, int32_t n_pairs, int32_t stride, int32_t xyz_off, int32_t label_off,
                    int32_t with_fitness, apd_result* results) {
  return batch_run(b, pairs, n_pairs, stride, xyz_off, label_off, false, with_fitness, results);
}

int apd_batch_align_device(apd_batch* b, const apd_pair* pairs, int32_t n_pairs, int32_t with_fitness, apd_result* results) {
  return batch_run(b, pairs, n_pairs, 16, 0, 12, true, with_fitness, results);
}

int64_t apd_batch_launch_count(const apd_batch* b) {
  int64_t n = 0;
  if (b)
    for (auto* h : b->handles) n += h->launches;
  return n;
}

int apd_batch_get_load_stats(apd_batch* b, double* stats, int32_t n, int32_t reset) {
  if (!b || !stats || n < 7) return APD_ERR_INVALID;
  for (int i = 0; i < n; i++) stats[i] = 0.0;
  for (auto* h : b->handles) {
    stats[0] += (double)h->phase_n;
    stats[1] += 1e3 * h->lm_kernel_s;
    for (int i = 0; i < 5; i++) stats[2 + i] += 1e3 * h->phase_s[i];
    for (int i = 0; i < 10 && 7 + i < n; i++) stats[7 + i] += 1e3 * h->lm_phase_s[i];
    if (reset) {
      for (int i = 0; i < 10; i++) h->lm_phase_s[i] = 0.0;
      h->phase_n = 0;
      h->lm_kernel_s = 0.0;
      for (int i = 0; i < 5; i++) h->phase_s[i] = 0.0;
    }
  }
  return APD_OK;
}

// Diagnostic: ONE launch of the loop kernel that registers the pairs set on n handles at once (n x cluster CTAs: with n
// >= 74 the GPU is as full as a pool keeps it) — so that ncu, which profiles kernels one at a time, sees the kernel under
// the co-residency it runs with in a pool. The handles' clouds must be set; their results are not unpacked.
int apd_debug_multi_align(apd_handle* const* hs, int32_t n, int32_t repeat) {
  if (!hs || n < 1) return APD_ERR_INVALID;
  apd_handle* h0 = hs[0];
  DeviceGuard dg(h0->device);
  std::vector<LmJob> jobs((size_t)n);
  LmConfig cfg = lm_config(h0->params);
  for (int rep = 0; rep < std::max(1, repeat); rep++) {
    for (int i = 0; i < n; i++) {
      apd_handle* h = hs[i];
      h->pooled = true;
      if (rep > 0) {  // the protocol of a pooled registration: the derived state of both clouds is dropped
        h->src.drop_derived();
        h->tgt.drop_derived();
        h->src.bbox_pending = h->src.ext_pts != nullptr;
        h->tgt.bbox_pending = h->tgt.ext_pts != nullptr;
      }
      const int bits = fused_prep_bits(h);
      int rc = bits ? prepare_fused(h, bits) : ensure_covariances_for_loop(h);
      if (rc == APD_OK) rc = ensure_corr_buffers(h);
      if (rc != APD_OK) return rc;
      if (!h->lm_result.p) APD_CUDA(h, h->lm_result.ensure(sizeof(LmResult)));
      jobs[(size_t)i] = lm_job(h, hm::Pose::identity(), bits);
      jobs[(size_t)i].seq = 0;
      jobs[(size_t)i].host_result = nullptr;
      APD_CUDA(h, cudaStreamSynchronize(h->stream));
    }
    DevBuf d_jobs;
    APD_CUDA(h0, d_jobs.ensure(jobs.size() * sizeof(LmJob)));
    APD_CUDA(h0, cudaMemcpy(d_jobs.p, jobs.data(), jobs.size() * sizeof(LmJob), cudaMemcpyHostToDevice));
    launch_lm(nullptr, d_jobs.as<LmJob>(), n, cfg, h0->lm_cluster, h0->lm_min_blocks, h0->stream, &h0->launches);
    const cudaError_t e = cudaStreamSynchronize(h0->stream);
    d_jobs.release();
    APD_CUDA(h0, e);
  }
  return APD_OK;
}

// Diagnostic: how many (empty) kernel launches per second this process can issue on `device` from n_threads host threads
// over n_streams streams. A pool of registrations issues ~11 launches each; this is the ceiling the driver puts on that.
int apd_debug_launch_rate(int device, int32_t n_streams, int32_t n_threads, int32_t launches_per_thread, double* per_second) {
  if (!per_second || n_streams < 1 || n_threads < 1 || launches_per_thread < 1) return APD_ERR_INVALID;
  DeviceGuard dg(device);
  std::vector<cudaStream_t> st((size_t)n_streams);
  for (auto& s : st)
    if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess) return APD_ERR_CUDA;
  int64_t dummy = 0;
  for (auto& s : st) launch_noop(s, &dummy);
  cudaDeviceSynchronize();
  const auto t0 = std::chrono::steady_clock::now();
  std::vector<std::thread> th;
  for (int t = 0; t < n_threads; t++)
    th.emplace_back([&, t] {
      cudaSetDevice(device);
      int64_t d2 = 0;
      for (int i = 0; i < launches_per_thread; i++) launch_noop(st[(size_t)((t + (int64_t)i * n_threads) % n_streams)], &d2);
    });
  for (auto& t : th) t.join();
  const double issue_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  cudaDeviceSynchronize();
  const double total_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  for (auto& s : st) cudaStreamDestroy(s);
  per_second[0] = (double)n_threads * launches_per_thread / issue_s;
  per_second[1] = (double)n_threads * launches_per_thread / total_s;
  return APD_OK;
}

int apd_batch_set_profiling(apd_batch* b, int32_t enabled) {
  if (!b) return APD_ERR_INVALID;
  for (auto* h : b->handles) apd_set_profiling(h, enabled);
  return APD_OK;
}

int apd_batch_get_kernel_ms(apd_batch* b, double* ms, int64_t* launches) {
  if (!b) return APD_ERR_INVALID;
  for (int i = 0; i < APD_K_COUNT; i++) {
    if (ms) ms[i] = 0.0;
    if (launches) launches[i] = 0;
  }
  for (auto* h : b->handles) {
    double m[APD_K_COUNT];
    int64_t l[APD_K_COUNT];
    apd_get_kernel_ms(h, m, l);
    for (int i = 0; i < APD_K_COUNT; i++) {
      if (ms) ms[i] += m[i];
      if (launches) launches[i] += l[i];
    }
  }
  return APD_OK;
}

int apd_align_batch(int device, const apd_params* p, const apd_pair* pairs, int32_t n_pairs, int32_t stride, int32_t xyz_off,
                    int32_t label_off, int32_t n_streams, int32_t with_fitness, apd_result* results) {
  const int32_t dev = device;
  return apd_align_batch_multi(&dev, 1, p, pairs, n_pairs, stride, xyz_off, label_off, n_streams, with_fitness, results);
}

int apd_align_batch_multi(const int32_t* devices, int32_t n_devices, const apd_params* p, const apd_pair* pairs, int32_t n_pairs, int32_t stride,
                          int32_t xyz_off, int32_t label_off, int32_t n_streams, int32_t with_fitness, apd_result* results) {
  if (!pairs || !results || n_pairs < 0 || !devices || n_devices < 1) return APD_ERR_INVALID;
  if (n_streams < 1) n_streams = 1;
  if (n_streams > n_pairs) n_streams = std::max(1, n_pairs);
  apd_batch* b = nullptr;
  int rc = apd_batch_create_multi(devices, n_devices, n_streams, &b);
  if (rc != APD_OK) return rc;
  if (p) rc = apd_batch_set_params(b, p);
  if (rc == APD_OK) rc = apd_batch_align(b, pairs, n_pairs, stride, xyz_off, label_off, with_fitness, results);
  apd_batch_destroy(b);
  return rc;
}

// ---- source-sharded registration ------------------------------------------------
int apd_comm_unique_id(void* id128) {
  std::string err;
  if (!id128 || !g_nccl.load(err)) return APD_ERR_COMM;
  ncclUniqueId id;
  if (g_nccl.GetUniqueId(&id) != 0) return APD_ERR_COMM;
  std::memcpy(id128, &id, sizeof(id));
  return APD_OK;
}
int apd_comm_init(apd_handle* h, const void* id128, int32_t rank, int32_t nranks, int64_t n_source_total) {
  if (!h || !id128 || rank < 0 || rank >= nranks) return APD_ERR_INVALID;
  if (!g_nccl.load(h->error)) return APD_ERR_COMM;
  DeviceGuard dg(h->device);
  ncclUniqueId id;
  std::memcpy(&id, id128, sizeof(id));
  ncclResult_t r = g_nccl.CommInitRank(&h->comm, nranks, id, rank);
  if (r != 0) return fail(h, APD_ERR_COMM, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "ncclCommInitRank failed");
  h->comm_rank = rank;
  h->comm_size = nranks;
  (void)n_source_total;  // (ABI v1 argument; the handle holds the full source, so the total is src.n)
  {  // NCCL connects lazily on the first collective (~0.2 s): pay for it here, not inside the first registration
    int rc = ensure_small(h);
    if (rc != APD_OK) return rc;
    double* d = h->small.as<double>() + 56;
    if (g_nccl.AllReduce(d, d, 1, kNcclFloat64, kNcclSum, h->comm, h->stream) != 0) return fail(h, APD_ERR_COMM, "ncclAllReduce (warm-up) failed");
    APD_CUDA(h, wait_stream(h));
  }
  h->corr_n = -1;        // correspondences of the unsharded layout are no longer addressable
  h->corr_warm = false;
  return APD_OK;
}
int apd_comm_peer_handle(apd_handle* h, void* handle64) {
  if (!h || !handle64) return APD_ERR_INVALID;
  if (!h->comm) return fail(h, APD_ERR_INVALID, "apd_comm_peer_handle needs apd_comm_init first");
  if (h->comm_size > kPeerMaxRanks) return fail(h, APD_ERR_UNSUPPORTED, "too many ranks for the peer-memory exchange");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  DeviceGuard dg(h->device);
  if (!h->mailbox) {
    APD_CUDA(h, cudaMalloc(reinterpret_cast<void**>(&h->mailbox), sizeof(PeerMailbox)));  // (IPC needs a plain cudaMalloc allocation)
    APD_CUDA(h, cudaMemset(h->mailbox, 0, sizeof(PeerMailbox)));
  }
  cudaIpcMemHandle_t ipc;
  APD_CUDA(h, cudaIpcGetMemHandle(&ipc, h->mailbox));
  std::memcpy(handle64, &ipc, sizeof(ipc));
  return APD_OK;
}

int apd_comm_peer_attach(apd_handle* h, const void* handles) {
  if (!h || !handles) return APD_ERR_INVALID;
  if (!h->comm || !h->mailbox) return fail(h, APD_ERR_INVALID, "apd_comm_peer_attach needs apd_comm_init and apd_comm_peer_handle first");
  DeviceGuard dg(h->device);
  for (int r = 0; r < h->comm_size; r++) {
    if (r == h->comm_rank) {
      h->peer_box[r] = h->mailbox;
      continue;
    }
    cudaIpcMemHandle_t ipc;
    std::memcpy(&ipc, reinterpret_cast<const char*>(handles) + (size_t)r * sizeof(ipc), sizeof(ipc));
    void* p = nullptr;
    APD_CUDA(h, cudaIpcOpenMemHandle(&p, ipc, cudaIpcMemLazyEnablePeerAccess));
    h->peer_box[r] = reinterpret_cast<PeerMailbox*>(p);
  }
  h->xchg_seq = 0;
  h->peers_attached = true;
  return APD_OK;
}

static void release_peers(apd_handle* h) {
  for (int r = 0; r < kPeerMaxRanks; r++) {
    if (h->peer_box[r] && h->peer_box[r] != h->mailbox) cudaIpcCloseMemHandle(h->peer_box[r]);
    h->peer_box[r] = nullptr;
  }
  if (h->mailbox) cudaFree(h->mailbox);
  h->mailbox = nullptr;
  h->peers_attached = false;
}

int apd_comm_destroy(apd_handle* h) {
  if (!h) return APD_ERR_INVALID;
  if (h->comm) {
    DeviceGuard dg(h->device);
    cudaStreamSynchronize(h->stream);
    release_peers(h);
    g_nccl.CommDestroy(h->comm);
    h->comm = nullptr;
    h->corr_n = -1;
    h->corr_warm = false;
  }
  h->comm_size = 1;
  h->comm_rank = 0;
  return APD_OK;
}

// ---- the same sharding between handles of ONE process ---------------------------------
}  // extern "C"

namespace {

void group_thread(apd_group* g, int rank) {
  cudaSetDevice(g->ranks[rank]->device);
  uint64_t seen = 0;
  for (;;) {
    std::function<int(apd_handle*, int)> job;
    {
      std::unique_lock<std::mutex> lk(g->mu);
      g->cv_job.wait(lk, [&] { return g->stop || g->job_gen != seen; });
      if (g->stop) return;
      seen = g->job_gen;
      job = g->job;
    }
    const int rc = job(g->ranks[rank], rank);
    {
      std::lock_guard<std::mutex> lk(g->mu);
      g->job_rc[rank] = rc;
      if (--g->job_pending == 0) g->cv_done.notify_all();
    }
  }
}

// runs `job` on every rank from the group's threads; returns the first non-zero status (rank order)
int group_run(apd_group* g, std::function<int(apd_handle*, int)> job) {
  {
    std::lock_guard<std::mutex> lk(g->mu);
    g->job = std::move(job);
    g->job_pending = (int)g->ranks.size();
    g->job_gen++;
  }
  g->cv_job.notify_all();
  std::unique_lock<std::mutex> lk(g->mu);
  g->cv_done.wait(lk, [&] { return g->job_pending == 0; });
  for (int rc : g->job_rc)
    if (rc != APD_OK) return rc;
  return APD_OK;
}

void group_detach(apd_handle* h) {
  for (int r = 0; r < kPeerMaxRanks; r++) h->peer_box[r] = nullptr;
  if (h->mailbox) {
    DeviceGuard dg(h->device);
    cudaStreamSynchronize(h->stream);
    cudaFree(h->mailbox);
  }
  h->mailbox = nullptr;
  h->peers_attached = false;
  h->group = nullptr;
  h->comm_size = 1;
  h->comm_rank = 0;
  h->corr_n = -1;
  h->corr_warm = false;
  h->src.drop_derived();  // (covariance arrays were sized and gathered for the sharded layout)
  h->tgt.drop_derived();
}

}  // namespace

extern "C" {

int apd_group_create(apd_handle* const* handles, int32_t n, apd_group** out) {
  if (!handles || !out || n < 1 || n > kPeerMaxRanks) return APD_ERR_INVALID;
  *out = nullptr;
  for (int r = 0; r < n; r++)
    if (!handles[r] || handles[r]->sharded()) return APD_ERR_INVALID;
  ensure_work_queues();
  apd_group* g = new apd_group();
  g->ranks.assign(handles, handles + n);
  g->job_rc.assign((size_t)n, APD_OK);
  // mailboxes, and peer access between the devices involved (one NVSwitch node: every pair is peer-accessible)
  for (int r = 0; r < n; r++) {
    apd_handle* h = handles[r];
    DeviceGuard dg(h->device);
    for (int q = 0; q < n; q++) {
      if (handles[q]->device == h->device) continue;
      const cudaError_t e = cudaDeviceEnablePeerAccess(handles[q]->device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
        h->error = std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e);
        for (int t = 0; t <= r; t++) group_detach(handles[t]);
        delete g;
        return APD_ERR_CUDA;
      }
      (void)cudaGetLastError();
    }
    if (cudaMalloc(reinterpret_cast<void**>(&h->mailbox), sizeof(PeerMailbox)) != cudaSuccess ||
        cudaMemset(h->mailbox, 0, sizeof(PeerMailbox)) != cudaSuccess) {
      for (int t = 0; t <= r; t++) group_detach(handles[t]);
      delete g;
      return APD_ERR_CUDA;
    }
  }
  for (int r = 0; r < n; r++) {  // no kernel may be loaded lazily once ranks of this process wait for each other inside kernels
    DeviceGuard dg(handles[r]->device);
    preload_grid_kernels(); preload_knn_kernels(); preload_corr_kernels(); preload_linearize_kernels(); preload_lm_kernels();
    preload_prep_kernels(); preload_vgicp_kernels();
  }
  for (int r = 0; r < n; r++) {
    apd_handle* h = handles[r];
    for (int q = 0; q < n; q++) h->peer_box[q] = handles[q]->mailbox;
    h->group = g;
    h->comm_rank = r;
    h->comm_size = n;
    h->peers_attached = true;
    h->xchg_seq = 0;
    h->bar_seq = 0;
    h->corr_n = -1;
    h->corr_warm = false;
    h->src.drop_derived();
    h->tgt.drop_derived();
  }
  for (int r = 0; r < n; r++) g->threads.emplace_back(group_thread, g, r);
  *out = g;
  return APD_OK;
}

int apd_group_destroy(apd_group* g) {
  if (!g) return APD_OK;
  {
    std::lock_guard<std::mutex> lk(g->mu);
    g->stop = true;
  }
  g->cv_job.notify_all();
  for (auto& t : g->threads) t.join();
  for (auto* h : g->ranks)
    if (h) group_detach(h);
  delete g;
  return APD_OK;
}

int apd_group_size(const apd_group* g) { return g ? (int)g->ranks.size() : 0; }

int apd_group_set_params(apd_group* g, const apd_params* p) {
  if (!g || !p) return APD_ERR_INVALID;
  for (auto* h : g->ranks) {
    const int rc = apd_set_params(h, p);
    if (rc != APD_OK) return rc;
  }
  return APD_OK;
}

int apd_group_set_source(apd_group* g, const void* pts, int32_t n, int32_t stride, int32_t xyz_off, int32_t label_off) {
  if (!g) return APD_ERR_INVALID;
  return group_run(g, [=](apd_handle* h, int) { return apd_set_source(h, pts, n, stride, xyz_off, label_off, 0); });
}
int apd_group_set_target(apd_group* g, const void* pts, int32_t n, int32_t stride, int32_t xyz_off, int32_t label_off) {
  if (!g) return APD_ERR_INVALID;
  return group_run(g, [=](apd_handle* h, int) { return apd_set_target(h, pts, n, stride, xyz_off, label_off, 0); });
}
int apd_group_set_source_device(apd_group* g, const void* const* d_xyzl, int32_t n) {
  if (!g || !d_xyzl) return APD_ERR_INVALID;
  return group_run(g, [=](apd_handle* h, int r) { return apd_set_source_device(h, d_xyzl[r], n); });
}
int apd_group_set_target_device(apd_group* g, const void* const* d_xyzl, int32_t n) {
  if (!g || !d_xyzl) return APD_ERR_INVALID;
  return group_run(g, [=](apd_handle* h, int r) { return apd_set_target_device(h, d_xyzl[r], n); });
}

int apd_group_align(apd_group* g, const float* guess, float* T_out, double* T_out_f64, double* H_out, int32_t* converged, int32_t* iterations) {
  if (!g) return APD_ERR_INVALID;
  return group_run(g, [=](apd_handle* h, int r) {  // every rank takes the same LM decisions on the reduced sums; rank 0 reports
    return r == 0 ? apd_align(h, guess, T_out, T_out_f64, H_out, converged, iterations, nullptr)
                  : apd_align(h, guess, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
  });
}
int apd_group_linearize(apd_group* g, const double* T, double* H, double* b, double* err) {
  if (!g || !T || !err) return APD_ERR_INVALID;
  return group_run(g, [=](apd_handle* h, int r) {
    double Hr[36], br[6], er;
    return r == 0 ? apd_linearize(h, T, H, b, err) : apd_linearize(h, T, (H && b) ? Hr : nullptr, (H && b) ? br : nullptr, &er);
  });
}
int apd_group_compute_error(apd_group* g, const double* T, double* err) {
  if (!g || !T || !err) return APD_ERR_INVALID;
  return group_run(g, [=](apd_handle* h, int r) {
    double er;
    return apd_compute_error(h, T, r == 0 ? err : &er);
  });
}

// ---- instrumentation -------------------------------------------------------------
void* apd_stream(apd_handle* h) { return h ? (void*)h->stream : nullptr; }
int64_t apd_launch_count(const apd_handle* h) { return h ? h->launches : 0; }
int apd_set_profiling(apd_handle* h, int32_t enabled) {
  if (!h) return APD_ERR_INVALID;
  flush_prof(h);
  h->profiling = enabled != 0;
  for (int i = 0; i < APD_K_COUNT; i++) { h->k_ms[i] = 0; h->k_launches[i] = 0; }
  return APD_OK;
}
int apd_get_kernel_ms(apd_handle* h, double* ms, int64_t* launches) {
  if (!h) return APD_ERR_INVALID;
  flush_prof(h);
  for (int i = 0; i < APD_K_COUNT; i++) {
    if (ms) ms[i] = h->k_ms[i];
    if (launches) launches[i] = h->k_launches[i];
  }
  return APD_OK;
}

}  // extern "C"
