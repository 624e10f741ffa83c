// Device-resident optimizer: the WHOLE LsqRegistration loop of one registration
// (reference lsq_registration_impl.hpp:55-173: outer iterations, step_lm / step_gn,
// is_converged) in ONE kernel launch — update_correspondences (:160-220), the
// linearize H/b/err sums (:247-304), the compute_error trials (:313-343), the 6x6
// LDL^T solve, so3_exp and the LM accept/reject logic all stay on the GPU, so a
// registration costs one launch and one device->host copy instead of ~2 host
// round trips per LM iteration.
//
// Mapping: one thread-block CLUSTER per registration (kLmCluster CTAs, co-scheduled
// on one GPC). The source points are split contiguously over the CTAs of the
// cluster. Every reduction is: fp64 per-thread sums (strided, fixed order) -> fixed
// warp-shuffle tree -> warps in order -> per-CTA partial in shared memory ->
// cluster barrier -> EVERY CTA adds the partials of all CTAs in rank order through
// distributed shared memory. All CTAs therefore hold bit-identical sums and run
// the serial part (solve, accept/reject) redundantly: no broadcast, one cluster
// barrier per reduction, and the control flow is uniform across the cluster.
// The per-point arithmetic is the code the streaming kernels use (point_math.cuh).
#include <cooperative_groups.h>

#include <cstddef>

#include "host_math.hpp"
#include "knn_warp.cuh"
#include "point_math.cuh"
#include "prep.cuh"

namespace cg = cooperative_groups;

namespace apd {

namespace {

#ifndef APD_LM_THREADS
#define APD_LM_THREADS 512
#endif
constexpr int kLmThreads = APD_LM_THREADS;
constexpr int kLmWarps = kLmThreads / 32;
#ifndef APD_LM_G
#define APD_LM_G 8
#endif
constexpr int kLmG = APD_LM_G;  // lanes per 1-NN query (the search is latency-bound at these sizes)

__device__ __forceinline__ hm::Pose pose_from(const PoseD& T) {
  hm::Pose p = hm::Pose::identity();
#pragma unroll
  for (int r = 0; r < 3; r++) {
#pragma unroll
    for (int c = 0; c < 3; c++) p(r, c) = T.r[r * 3 + c];
    p(r, 3) = T.t[r];
  }
  return p;
}
__device__ __forceinline__ PoseD pose_to(const hm::Pose& p) {
  PoseD T;
#pragma unroll
  for (int r = 0; r < 3; r++) {
#pragma unroll
    for (int c = 0; c < 3; c++) T.r[r * 3 + c] = p(r, c);
    T.t[r] = p(r, 3);
  }
  return T;
}

// Serial optimizer state. It lives in shared memory and only thread 0 of a CTA touches it (every CTA of
// the cluster holds an identical copy), so it costs the parallel phases no registers.
struct LmSerial {
  hm::Pose x0, delta, xi;
  double H[36], b6[6], d[6];
  double y0, lambda, nu;
  int n_rows, nr_iterations, converged, lm_failed, h_set;
};
struct LmShared {
  double warp[kLmWarps][kReduceVals];
  double part[2][kReduceVals];  // this CTA's partial sums, double-buffered across reductions
  double out[kReduceVals];      // cluster-wide sums (identical in every CTA)
  PoseD T;                      // pose the next phase evaluates
  PoseD T_corr;                 // pose of the last update_correspondences pass (warm start of the next one)
  unsigned long long kbuf[kLmWarps][32];  // on-demand target covariances: the kNN search's candidate buffer, per warp
  int n_need;                   // ... and the number of target points this CTA serves in the current pass
  int next_q, next_li;          // work counters of the search passes: groups / warps take the next query when they are free
  int flag_in, flag_out;        // LM trial decision / outer-loop decision (separate words: each is re-read across one barrier only)
  LmSerial ser;
};

// Phase clock (diagnostic build, -DAPD_LM_PHASE_TIMING): thread 0 of CTA 0 adds the time since the last tick to a phase's
// counter — after the barrier that ends the phase, so waiting for slower warps / CTAs counts towards the phase
struct PhaseClock {
#ifdef APD_LM_PHASE_TIMING
  unsigned long long last = 0, ns[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  bool on = false;
  __device__ __forceinline__ void start(bool writer) {
    on = writer;
    if (on) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(last));
  }
  __device__ __forceinline__ void tick(int i) {
    if (!on) return;
    unsigned long long now;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
    ns[i] += now - last;
    last = now;
  }
  __device__ __forceinline__ void store(LmResult* r) const {
    if (on)
      for (int i = 0; i < 10; i++) r->phase_ns[i] = ns[i];
  }
#else
  __device__ __forceinline__ void start(bool) {}
  __device__ __forceinline__ void tick(int) {}
  __device__ __forceinline__ void store(LmResult*) const {}
#endif
};

// acc[NV] of every thread -> s.out[0..NV) = the sum over the whole cluster
template <int NV>
__device__ __forceinline__ void cluster_reduce(cg::cluster_group& cluster, LmShared& s, const double* acc, int& phase) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int j = 0; j < NV; j++) {
    double v = acc[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) s.warp[warp][j] = v;
  }
  __syncthreads();
  const int buf = phase & 1;
  if (tid < NV) {
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < kLmWarps; w++) v += s.warp[w][tid];
    s.part[buf][tid] = v;
  }
  cluster.sync();
  if (tid < NV) {
    double v = 0.0;
    const unsigned nb = cluster.num_blocks();
    for (unsigned r = 0; r < nb; r++) v += cluster.map_shared_rank(&s.part[buf][0], r)[tid];
    s.out[tid] = v;
  }
  __syncthreads();
  phase++;
}

template <bool kFp64>
__device__ __forceinline__ void load_maha(const LmJob& job, int i, double m[6], double& geo) {
  if (kFp64) {
    const double2 a = reinterpret_cast<const double2*>(job.mahaA)[i];
    const double2 b = reinterpret_cast<const double2*>(job.mahaB)[i];
    const double2 c = reinterpret_cast<const double2*>(job.mahaB)[(size_t)job.n_src + i];
    m[0] = a.x; m[1] = a.y; m[2] = b.x; m[3] = b.y; m[4] = c.x; m[5] = c.y;
    geo = job.s_geo64[i];
  } else {
    const float4 a = reinterpret_cast<const float4*>(job.mahaA)[i];
    const float2 b = reinterpret_cast<const float2*>(job.mahaB)[i];
    m[0] = (double)a.x; m[1] = (double)a.y; m[2] = (double)a.z; m[3] = (double)a.w; m[4] = (double)b.x; m[5] = (double)b.y;
    geo = (double)job.s_geo[i];
  }
}

// Publication of an on-demand covariance: the six values, then the flag with RELEASE semantics at device scope; a
// reader that sees the flag through an ACQUIRE load is guaranteed to see the values (it reads them from L2, __ldcg).
// Several CTAs — of one registration or of registrations sharing a target — may compute the same point concurrently:
// they store the same bits, and the flag only ever goes 0 -> 1.
__device__ __forceinline__ void flag_publish(unsigned char* f) {
  asm volatile("st.release.gpu.global.u8 [%0], %1;" ::"l"(f), "r"(1u) : "memory");
}
__device__ __forceinline__ unsigned flag_acquire(const unsigned char* f) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u8 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
  return v;
}

// covariance + regularisation of target point `pos` from its neighbour ids (the per-cloud kernels' code: same bits);
// __noinline__ keeps the Jacobi sweep's registers out of the optimizer loop
__device__ __noinline__ void lazy_target_covariance(const LmJob& job, const int32_t* nbq, int pos) {
  const Sym3 C = knnw::covariance_of_neighbors(job.t_pts, nbq, job.k, job.reg);
#pragma unroll
  for (int e = 0; e < 6; e++) __stcg(job.t_cov_rw + (size_t)pos * 6 + e, C.v[e]);
  flag_publish(&job.t_cov_flag[pos]);
}

// FastAPDGICP::update_correspondences (:160-220) for this CTA's points [base, base+cnt)
// warm: job.corr / job.sqd hold the previous outer iteration's pass (pose T_prev), which bounds the searches
template <bool kFp64>
__device__ __forceinline__ void corr_phase(const LmJob& job, const LmConfig& cfg, LmShared& smem, const PoseD& T, int base, int cnt, bool warm,
                                           const PoseD& T_prev, PhaseClock& clk) {
  const PoseF Tf = pose_to_f32(T);
  const PoseF Tpf = pose_to_f32(T_prev);
  const int tid = threadIdx.x;
  if (tid == 0) {
    smem.n_need = 0;
    smem.next_q = 0;
    smem.next_li = 0;
  }
  __syncthreads();
  // The searches differ in cost (a cold query walks shells, a warm one checks a row, a provably unmatched one returns at
  // once): the groups of kLmG lanes take queries from a counter instead of a fixed stride, so no warp waits at the
  // barrier below for the unluckiest one. Who serves a query does not matter: it writes that query's own slots.
  {
    const unsigned gmask = nn_group_mask<kLmG>();
    const int leader = (int)(tid & 31) & ~(kLmG - 1);
    for (;;) {
      int q = 0;
      if ((tid & (kLmG - 1)) == 0) q = atomicAdd(&smem.next_q, 1);
      q = __shfl_sync(gmask, q, leader);
      if (q >= cnt) break;
      const int i = base + q;
      const float4 a = job.s_spts[i];
      float px, py, pz;
      transform_rn(Tf, a.x, a.y, a.z, px, py, pz);  // :176
      int seed = -1;
      if (warm) {
        float kept;
        const bool search = warm_start(Tpf, a, px, py, pz, job.corr[i], job.sqd[i], cfg.np.thr_sq, seed, kept);  // (uniform over the lanes of a query)
        nn_group_sync<kLmG>();  // every lane has read corr[i] / sqd[i] before one of them rewrites them
        if (!search) {
          if ((tid & (kLmG - 1)) == 0) job.sqd[i] = kept;  // provably still unmatched; corr[i] stays -1
          continue;
        }
      }
      unsigned long long best;
      int pos;
      float proven2;
      nn_search<kLmG>(job.t_spts, job.t_cell_start, job.tg, px, py, pz, cfg.np.thr_sq, best, pos, seed, proven2);  // :178
      if ((tid & (kLmG - 1)) != 0) continue;
      const float d2 = (best == kInfKey) ? 3.402823466e38f : __uint_as_float((unsigned)(best >> 32));
      const bool ok = (best != kInfKey) && ((double)d2 < cfg.np.thr_sq);  // :183
      if (!ok) {
        job.corr[i] = -1;
        job.sqd[i] = fminf(d2, proven2);  // rejected: the proven lower bound of its distance (see corr.cu)
        continue;
      }
      job.sqd[i] = d2;  // :180
      job.corr[i] = pos | ((job.t_label[pos] == job.s_label[i]) ? kCorrLabelBit : 0);
    }
  }
  __syncthreads();
  clk.tick(2);
  // Target covariances on demand (calculate_covariances(target), :351-411, restricted to the points that are used).
  // The matched target points without a covariance are listed (this CTA's slice of job.nb); pass A, one WARP per listed
  // point: the exact kNN search of the per-cloud kernel (knn_warp.cuh); pass B, one THREAD per listed point: covariance
  // + regularisation. Two threads
  // (or two registrations sharing a target) may compute the same point; they write the same bits, so the race is benign;
  // the flag is published after the values, and readers take the values from L2 (__ldcg), never from a stale L1 line.
  if (job.t_cov_flag) {
    const int lane = tid & 31, warp = tid >> 5;
    int32_t* list = job.nb + (size_t)job.n_src * job.k;  // [n_src] sorted positions of the target points to serve, per CTA slice
    for (int q = tid; q < cnt; q += kLmThreads) {
      const int c = job.corr[base + q];
      if (c < 0) continue;
      const int pos = c & kCorrIndexMask;
      // (the thread that sees the flag here is the one that reads this point's covariance in the last pass below)
      if (flag_acquire(&job.t_cov_flag[pos]) == 0) list[base + atomicAdd(&smem.n_need, 1)] = pos;
    }
    __syncthreads();
    const int n_need = smem.n_need;
    for (;;) {  // warps take listed points from a counter (a kNN search walks 1 to 4 shells)
      int li = 0;
      if (lane == 0) li = atomicAdd(&smem.next_li, 1);
      li = __shfl_sync(0xffffffffu, li, 0);
      if (li >= n_need) break;
      const unsigned long long key = knnw::knn_warp_query(job.t_spts, job.t_cell_start, job.tg, job.k, list[base + li], lane, smem.kbuf[warp]);
      if (lane < job.k) job.nb[(size_t)(base + li) * job.k + lane] = (int)(unsigned)(key & 0xffffffffull);
    }
    __syncthreads();
    clk.tick(3);
    for (int li = tid; li < n_need; li += kLmThreads) lazy_target_covariance(job, job.nb + (size_t)(base + li) * job.k, list[base + li]);
    __threadfence();
    __syncthreads();
    clk.tick(4);
  }
  // second pass, one THREAD per point: the fp64 noise model and Mahalanobis matrix of the matched points (:194-218). In the
  // search pass only one lane in kLmG holds a result; here all lanes work.
  for (int q = tid; q < cnt; q += kLmThreads) {
    const int i = base + q;
    const int c = job.corr[i];
    if (c < 0) continue;
    const int pos = c & kCorrIndexMask;
    const float4 a = job.s_spts[i];
    float px, py, pz;
    transform_rn(Tf, a.x, a.y, a.z, px, py, pz);
    double cb[6];
    if (job.t_cov_flag) {
#pragma unroll
      for (int e = 0; e < 6; e++) cb[e] = __ldcg(job.t_cov + (size_t)pos * 6 + e);
    } else {
#pragma unroll
      for (int e = 0; e < 6; e++) cb[e] = job.t_cov[(size_t)pos * 6 + e];
    }
    const Sym3 M = mahalanobis_of(px, py, pz, job.s_cov + (size_t)i * 6, cb, T, cfg.np);
    if (kFp64) {
      double2* mA = reinterpret_cast<double2*>(job.mahaA);
      double2* mB = reinterpret_cast<double2*>(job.mahaB);
      mA[i] = make_double2(M.v[0], M.v[1]);
      mB[i] = make_double2(M.v[2], M.v[3]);
      mB[(size_t)job.n_src + i] = make_double2(M.v[4], M.v[5]);
    } else {
      reinterpret_cast<float4*>(job.mahaA)[i] = make_float4((float)M.v[0], (float)M.v[1], (float)M.v[2], (float)M.v[3]);
      reinterpret_cast<float2*>(job.mahaB)[i] = make_float2((float)M.v[4], (float)M.v[5]);
    }
  }
}

// the linearize (:247-304) / compute_error (:313-343) sums over this CTA's points
template <bool kFp64, bool kHB>
__device__ __forceinline__ void sum_phase(const LmJob& job, const PoseD& T, int base, int cnt, double* acc) {
  constexpr int NV = kHB ? kReduceVals : 1;
#pragma unroll
  for (int j = 0; j < NV; j++) acc[j] = 0.0;
  for (int q = threadIdx.x; q < cnt; q += kLmThreads) {
    const int i = base + q;
    const int c = job.corr[i];
    const bool valid = c >= 0;
    const float4 a = job.s_spts[i];
    const float4 b = job.t_spts[valid ? (c & kCorrIndexMask) : 0];
    double m[6], geo;
    load_maha<kFp64>(job, i, m, geo);
    accumulate_point<kHB>(acc, valid, a, b, m, geo, c, T, job.cl_w);
  }
}

__device__ __forceinline__ void trace_row(LmResult* res, int& n_rows, int outer, int inner, double y0, double yi, double rho, double lambda,
                                          double dn, bool accepted) {
  if (n_rows < kLmTraceRows) {
    double* row = res->trace + (size_t)n_rows * 8;
    row[0] = (double)outer; row[1] = (double)inner; row[2] = y0; row[3] = yi;
    row[4] = rho; row[5] = lambda; row[6] = dn; row[7] = accepted ? 1.0 : 0.0;
  }
  n_rows++;
}

enum { kLmContinue = 0, kLmStepDone = 1 };

// ---- the serial pieces (thread 0 of every CTA; __noinline__ keeps their registers and stack out of the hot loops) ----
__device__ __noinline__ void serial_after_linearize(LmShared& s, const LmConfig& cfg) {
  LmSerial& z = s.ser;
  hm::unpack_upper(s.out, z.H);
  for (int j = 0; j < 6; j++) z.b6[j] = s.out[21 + j];
  z.y0 = s.out[27];
  if (cfg.optimizer != 0) {
    if (z.lambda < 0.0) {  // lsq :131-133
      double mx = 0.0;
      for (int j = 0; j < 6; j++) mx = hm::dmax(mx, fabs(z.H[j * 6 + j]));
      z.lambda = cfg.lm_init_lambda_factor * mx;
    }
    z.nu = 2.0;
  }
}
// step_gn (lsq :107-123) + the convergence test of the outer loop (lsq :75)
__device__ __noinline__ void serial_step_gn(LmShared& s, const LmConfig& cfg, LmResult* res, bool writer, int it) {
  LmSerial& z = s.ser;
  double nb[6];
  for (int j = 0; j < 6; j++) nb[j] = -z.b6[j];
  hm::ldlt_solve6(z.H, nb, z.d);
  z.delta = hm::delta_from_twist(z.d);
  z.x0 = hm::compose(z.delta, z.x0);
  double dn = 0;
  for (int j = 0; j < 6; j++) dn += z.d[j] * z.d[j];
  if (writer) {
    for (int j = 0; j < 36; j++) res->H[j] = z.H[j];
    z.h_set = 1;
    trace_row(res, z.n_rows, it, 0, z.y0, z.y0, 0.0, 0.0, sqrt(dn), true);
  }
  z.converged = hm::is_converged(z.delta, cfg.rotation_epsilon, cfg.transformation_epsilon) ? 1 : 0;
  s.T = pose_to(z.x0);
  s.flag_out = z.converged;
}
// one LM trial: solve (H + lambda I) d = -b, xi = exp(d) * x0 (lsq :137-144)
__device__ __noinline__ void serial_lm_trial(LmShared& s) {
  LmSerial& z = s.ser;
  double Hl[36], nb[6];
  for (int q = 0; q < 36; q++) Hl[q] = z.H[q];
  for (int q = 0; q < 6; q++) {
    Hl[q * 6 + q] += z.lambda;
    nb[q] = -z.b6[q];
  }
  hm::ldlt_solve6(Hl, nb, z.d);
  z.delta = hm::delta_from_twist(z.d);
  z.xi = hm::compose(z.delta, z.x0);
  s.T = pose_to(z.xi);
}
// rho test and accept / reject (lsq :146-169)
__device__ __noinline__ void serial_lm_decide(LmShared& s, const LmConfig& cfg, LmResult* res, bool writer, int it, int j) {
  LmSerial& z = s.ser;
  const double yi = s.out[0];
  double denom = 0.0, dn = 0.0;
  for (int q = 0; q < 6; q++) {
    denom += z.d[q] * (z.lambda * z.d[q] - z.b6[q]);
    dn += z.d[q] * z.d[q];
  }
  const double rho = (z.y0 - yi) / denom;  // lsq :146
  if (writer) trace_row(res, z.n_rows, it, j, z.y0, yi, rho, z.lambda, sqrt(dn), !(rho < 0));
  int flag = kLmContinue;
  if (rho < 0) {  // lsq :156-164
    if (hm::is_converged(z.delta, cfg.rotation_epsilon, cfg.transformation_epsilon)) {
      flag = kLmStepDone;
    } else {
      z.lambda = z.nu * z.lambda;
      z.nu = 2 * z.nu;
    }
  } else {  // lsq :166-169
    z.x0 = z.xi;
    const double t = 2 * rho - 1;
    z.lambda = z.lambda * hm::dmax(1.0 / 3.0, 1 - pow(t, 3.0));
    if (writer) {
      for (int q = 0; q < 36; q++) res->H[q] = z.H[q];
      z.h_set = 1;
    }
    flag = kLmStepDone;
  }
  s.flag_in = flag;
}
// end of step_lm: "lm not converged!!" (lsq :71-74) or the convergence test (lsq :75)
__device__ __noinline__ void serial_lm_end(LmShared& s, const LmConfig& cfg, bool step_ok) {
  LmSerial& z = s.ser;
  if (!step_ok) {
    z.lm_failed = 1;
    s.flag_out = 2;
  } else {
    z.converged = hm::is_converged(z.delta, cfg.rotation_epsilon, cfg.transformation_epsilon) ? 1 : 0;
    s.flag_out = z.converged;
    s.T = pose_to(z.x0);
  }
}

// kMinB = CTAs per SM the register allocation allows for. 1: 128 registers per thread — the shortest single registration
// (a lone handle). 2: 64 registers (some fp64 state spills to L1-backed local memory) and twice the warps per SM — a batch
// pool is bound by the loop kernel's latency-limited SMs (35 % issue utilisation at 16 warps), and two resident CTAs of
// different registrations hide each other's stalls: 20.7 k -> 26.3 k registrations/s on C2.
template <bool kFp64, int kMinB>
__global__ void __launch_bounds__(kLmThreads, kMinB) lm_kernel(LmJob one, LmJob two, const LmJob* __restrict__ jobs, LmConfig cfg) {
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned C = cluster.num_blocks();
  const unsigned rank = cluster.block_rank();
  // (two registrations of a pool may share a launch — by value, `two` is the second cluster's: a device runs at most 128
  // grids at a time, fewer than the 148 cluster slots the pool's build of this kernel has)
  LmJob job = jobs ? jobs[blockIdx.x / C] : (blockIdx.x < C ? one : two);
  __shared__ LmShared s;
  const int tid = threadIdx.x;
  const bool writer = (rank == 0 && tid == 0);  // the one thread that reports results
  unsigned long long t_begin = 0;
  if (writer) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_begin));
  PhaseClock clk;
  clk.start(writer);

  // ---- fused prologue: from the raw clouds to grids + source covariances, inside this launch (prep.cuh) ----
  int s_ncells = job.s_ncells, t_ncells = job.t_ncells;
  GridDesc s_grid = job.sg;
  if (job.prep) {
    __shared__ prep::PrepShared ps;
    const int gt = (int)rank * kLmThreads + tid, GT = (int)C * kLmThreads;
    const int first = (job.prep & 1) ? 0 : 1;  // clouds [first, last) are built: 0 source, 1 target
    const int last = (job.prep & 2) ? 2 : 1;
    const float4* const pts[2] = {job.s_pts, job.t_pts};
    const int np[2] = {(job.prep & 1) ? job.n_src : 0, (job.prep & 2) ? job.n_tgt : 0};
    prep::bounds_phase(cluster, ps, pts, np, gt, GT);
    if (tid == 0) {
      ps.status = 0;
      ps.grid[0] = job.sg; ps.ncells[0] = job.s_ncells;
      ps.grid[1] = job.tg; ps.ncells[1] = job.t_ncells;
      for (int c = first; c < last; c++) {
        float bbox[6];
        for (int a = 0; a < 6; a++) bbox[a] = prep::ord2f(ps.box[c][a]);
        size_grid(bbox, c == 0 ? job.n_src : job.n_tgt, c == 0 ? job.s_cells_per_point : job.t_cells_per_point, ps.grid[c], ps.ncells[c]);
        if (ps.ncells[c] + 1 > (c == 0 ? job.s_cell_cap : job.t_cell_cap)) ps.status = 1;
      }
    }
    __syncthreads();
    s_grid = ps.grid[0]; s_ncells = ps.ncells[0];
    job.tg = ps.grid[1]; t_ncells = ps.ncells[1];
    if (ps.status != 0) {  // (uniform over the cluster: every CTA computed the same grids from the same boxes)
      if (writer) {
        job.result->prep_status = 1;
        if (job.host_result) {
          job.host_result->prep_status = 1;
          __threadfence_system();
          *reinterpret_cast<volatile unsigned long long*>(&job.host_result->seq) = job.seq;
        }
      }
      cluster.sync();
      return;
    }
    prep::GridJob gj[2];
    gj[0] = prep::GridJob{job.s_pts, job.n_src, job.s_cell_start, job.s_cell_cap, job.s_scratch, job.s_scratch + job.n_src,
                          job.s_scratch + 2 * (size_t)job.n_src, const_cast<float4*>(job.s_spts), const_cast<float*>(job.s_label), job.s_inv_perm,
                          nullptr, true};
    gj[1] = prep::GridJob{job.t_pts, job.n_tgt, const_cast<uint32_t*>(job.t_cell_start), job.t_cell_cap, job.t_scratch, job.t_scratch + job.n_tgt,
                          job.t_scratch + 2 * (size_t)job.n_tgt, const_cast<float4*>(job.t_spts), const_cast<float*>(job.t_label), job.t_inv_perm,
                          job.t_cov_flag, false};
    if (first < last) prep::grids_phase(cluster, ps, gj, first, last, gt, GT);
    clk.tick(0);
    if (job.prep & 1)
      prep::source_cov_phase(cluster, job.s_pts, job.s_spts, job.s_cell_start, s_grid, job.n_src, job.s_k, job.s_reg, job.gicp, job.nb,
                             const_cast<double*>(job.s_cov), const_cast<float*>(job.s_geo), const_cast<double*>(job.s_geo64), s.kbuf[tid >> 5], gt, GT);
    clk.tick(1);
  }

  // this CTA's contiguous slice of the (cell-sorted) source points
  const int per = (job.n_src + (int)C - 1) / (int)C;
  const int base = (int)rank * per;
  const int cnt = max(0, min(per, job.n_src - base));

  if (tid == 0) {
    PoseD g;
#pragma unroll
    for (int i = 0; i < 9; i++) g.r[i] = job.guess[i];
#pragma unroll
    for (int i = 0; i < 3; i++) g.t[i] = job.guess[9 + i];
    LmSerial& z = s.ser;
    z.x0 = pose_from(g);  // lsq :56
    z.delta = hm::Pose::identity();
    z.lambda = -1.0;  // lm_lambda_ reset (lsq :58)
    z.n_rows = 0; z.nr_iterations = 0; z.converged = 0; z.lm_failed = 0; z.h_set = 0;
    s.T = g;
  }
  __syncthreads();
  int phase = 0;
  double acc[kReduceVals];

  for (int it = 0; it < cfg.max_iterations; it++) {  // lsq :67
    if (tid == 0) s.ser.nr_iterations = it;           // lsq :68
    // ---- linearize(x0) (:224-307) ----
    {
      const PoseD Tx0 = s.T;
      corr_phase<kFp64>(job, cfg, s, Tx0, base, cnt, it > 0, s.T_corr, clk);
      __syncthreads();  // the correspondences of this CTA's points are visible to all of its threads; T_corr has been read
      clk.tick(5);
      if (tid == 0) s.T_corr = Tx0;
      sum_phase<kFp64, true>(job, Tx0, base, cnt, acc);
    }
    cluster_reduce<kReduceVals>(cluster, s, acc, phase);
    if (tid == 0) serial_after_linearize(s, cfg);
    clk.tick(6);
    if (cfg.optimizer == 0) {
      if (tid == 0) serial_step_gn(s, cfg, job.result, writer, it);
      __syncthreads();
      if (s.flag_out) break;
      continue;
    }
    // ---- step_lm (lsq :127-173) ----
    bool step_ok = false;
    for (int j = 0; j < cfg.lm_max_iterations; j++) {  // lsq :136
      if (tid == 0) serial_lm_trial(s);
      __syncthreads();
      // ---- compute_error(xi) (:310-346): stale correspondences, trial pose ----
      {
        const PoseD Txi = s.T;
        sum_phase<kFp64, false>(job, Txi, base, cnt, acc);
      }
      cluster_reduce<1>(cluster, s, acc, phase);
      if (tid == 0) serial_lm_decide(s, cfg, job.result, writer, it, j);
      __syncthreads();
      if (s.flag_in == kLmStepDone) {
        step_ok = true;
        break;
      }
    }
    if (tid == 0) serial_lm_end(s, cfg, step_ok);
    __syncthreads();
    clk.tick(7);
    if (s.flag_out) break;
  }

  // ---- optional tail: pcl getFitnessScore of the final (float) pose + inlier count ----
  if (cfg.want_fitness) {
    if (tid == 0) s.T = pose_to(s.ser.x0);
    __syncthreads();
    const PoseD Tfin = s.T;
    const PoseF Tf = pose_to_f32(Tfin);  // final_transformation_ = x0.cast<float>() (lsq :78)
    double f[3] = {0.0, 0.0, 0.0};
    constexpr int kQ = kLmThreads / kLmG;
    for (int p0 = 0; p0 < cnt; p0 += kQ) {
      const int q = p0 + tid / kLmG;
      if (q >= cnt) continue;  // (the lanes of a group share q: the whole group sits this round out)
      const float4 a = job.s_spts[base + q];
      float px, py, pz;
      transform_rn(Tf, a.x, a.y, a.z, px, py, pz);
      unsigned long long best;
      int pos;
      nn_search<kLmG>(job.t_spts, job.t_cell_start, job.tg, px, py, pz, 1e300, best, pos);
      if ((tid & (kLmG - 1)) == 0 && best != kInfKey) {
        const double d2 = (double)__uint_as_float((unsigned)(best >> 32));
        if (d2 <= cfg.fitness_max_range) { f[0] += d2; f[1] += 1.0; }
        if (d2 < cfg.inlier_sq_thr) f[2] += 1.0;
      }
    }
    cluster_reduce<3>(cluster, s, f, phase);
    clk.tick(8);
  }

  if (writer) {
    LmResult* res = job.result;
    const LmSerial& z = s.ser;
    const PoseD T = pose_to(z.x0);
#pragma unroll
    for (int i = 0; i < 9; i++) res->pose[i] = T.r[i];
#pragma unroll
    for (int i = 0; i < 3; i++) res->pose[9 + i] = T.t[i];
    res->lm_lambda = z.lambda;
    res->fitness[0] = cfg.want_fitness ? s.out[0] : 0.0;
    res->fitness[1] = cfg.want_fitness ? s.out[1] : 0.0;
    res->fitness[2] = cfg.want_fitness ? s.out[2] : 0.0;
    res->converged = z.converged;
    res->nr_iterations = z.nr_iterations;
    res->lm_failed = z.lm_failed;
    res->n_trace = z.n_rows;
    res->hessian_set = z.h_set;
    res->t_begin = t_begin;
    res->grid[0] = s_grid; res->grid[1] = job.tg;
    res->ncells[0] = s_ncells; res->ncells[1] = t_ncells;
    res->prep_status = 0;
    clk.store(res);
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(res->t_end));
  }
  if (job.host_result && rank == 0) {
    // publish: header and first trace rows to the host (one 8-byte word per thread), then the sequence number
    __syncthreads();
    constexpr int kWords = (int)((offsetof(LmResult, trace) + (size_t)kLmTraceHead * 8 * sizeof(double)) / 8);
    constexpr int kSeqWord = (int)(offsetof(LmResult, seq) / 8);
    const unsigned long long* srcw = reinterpret_cast<const unsigned long long*>(job.result);
    unsigned long long* dstw = reinterpret_cast<unsigned long long*>(job.host_result);
    for (int w = tid; w < kWords; w += kLmThreads)
      if (w != kSeqWord) dstw[w] = srcw[w];
    __threadfence_system();
    __syncthreads();
    if (tid == 0) *reinterpret_cast<volatile unsigned long long*>(&job.host_result->seq) = job.seq;
  }
  cluster.sync();  // no CTA may exit while a peer can still read its shared memory
}

}  // namespace

void launch_lm(const LmJob* one, const LmJob* d_jobs, int n_jobs, const LmConfig& cfg, int cluster, int min_blocks, cudaStream_t s,
               int64_t* launches, const LmJob* two) {
  if (two && !d_jobs) n_jobs = 2;
  if (n_jobs <= 0) return;
  cudaLaunchConfig_t lc = {};
  lc.gridDim = dim3((unsigned)(n_jobs * cluster), 1, 1);
  lc.blockDim = dim3(kLmThreads, 1, 1);
  lc.dynamicSmemBytes = 0;
  lc.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)cluster;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  lc.attrs = at;
  lc.numAttrs = 1;
  LmJob byval{}, byval2{};
  if (one) byval = *one;
  if (two) byval2 = *two;
  if (cluster > 8) {  // beyond the portable cluster size: opt in once per kernel
    static bool allowed = [] {
      cudaFuncSetAttribute(lm_kernel<true, 1>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
      cudaFuncSetAttribute(lm_kernel<false, 1>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
      cudaFuncSetAttribute(lm_kernel<true, 2>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
      cudaFuncSetAttribute(lm_kernel<false, 2>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
      return true;
    }();
    (void)allowed;
  }
  if (min_blocks >= 2) {
    if (cfg.maha_fp64) cudaLaunchKernelEx(&lc, lm_kernel<true, 2>, byval, byval2, d_jobs, cfg);
    else cudaLaunchKernelEx(&lc, lm_kernel<false, 2>, byval, byval2, d_jobs, cfg);
  } else {
    if (cfg.maha_fp64) cudaLaunchKernelEx(&lc, lm_kernel<true, 1>, byval, byval2, d_jobs, cfg);
    else cudaLaunchKernelEx(&lc, lm_kernel<false, 1>, byval, byval2, d_jobs, cfg);
  }
  (*launches)++;
}

// Loads this file's kernels into the current context (CUDA loads kernels lazily, at their first launch, and a load may have
// to synchronise with the context: if it happens while another rank's kernel of the same process is spinning on a peer
// — the sharded exchange — neither can proceed. apd_group_create loads everything up front.)
void preload_lm_kernels() {
  cudaFuncAttributes a;
  (void)cudaFuncGetAttributes(&a, lm_kernel<false, 1>);
  (void)cudaFuncGetAttributes(&a, lm_kernel<false, 2>);
  (void)cudaFuncGetAttributes(&a, lm_kernel<true, 1>);
  (void)cudaFuncGetAttributes(&a, lm_kernel<true, 2>);
}

}  // namespace apd
