// K3 — update_correspondences: exact 1-NN of every transformed source point on
// the target grid, the radar range/azimuth/elevation noise covariance at that
// point, the combined covariance and its inverse (the per-point Mahalanobis
// matrix). Replaces FastAPDGICP::update_correspondences (reference
// fast_apdgicp_impl.hpp:160-220). Also the fitness pass (pcl getFitnessScore)
// and the export hooks.
#include "point_math.cuh"

namespace apd {

namespace {

constexpr int kCorrLanesLarge = 1;  // lanes per 1-NN query on large source clouds

PoseF pose_to_f32_host(const PoseD& T) {
  PoseF f;
  for (int i = 0; i < 9; i++) f.r[i] = (float)T.r[i];  // Isometry3d::cast<float>() (:164)
  for (int i = 0; i < 3; i++) f.t[i] = (float)T.t[i];
  return f;
}

constexpr int kThreads = 128;

// rank-local slot l -> chunk j and offset o; false if the slot is padding
__device__ __forceinline__ bool slot_of(const ShardTable& sh, int l, int& j, int& o) {
  if (sh.nsub == 1) { j = 0; o = l; }
  else { j = l / sh.chunk; o = l - j * sh.chunk; }
  return j < sh.nsub && o < sh.count_of(j);
}

// ---- pass 1: the search (:176-190). fp32 and index work only: no fp64 state is live, so many warps fit an SM and the
// divergent candidate loop is all the kernel has to hide. G lanes share a query.
template <int G>
__global__ void __launch_bounds__(kThreads) corr_search_kernel(const float4* __restrict__ s_spts, const float* __restrict__ s_label,
                                                               ShardTable sh, const float4* __restrict__ t_spts,
                                                               const float* __restrict__ t_label, const uint32_t* __restrict__ t_cell_start,
                                                               GridDesc tg, PoseF Tf, double thr_sq, int* __restrict__ corr,
                                                               float* __restrict__ sqd, int warm, PoseF Tpf, float* __restrict__ second,
                                                               int second_valid) {
  // kThreads and the warp size are multiples of G, so a group never straddles a warp; the lanes of a group share the
  // slot, so a padding slot retires its whole group (every shuffle below is masked to the group's own lanes)
  int l = (int)(((size_t)blockIdx.x * kThreads + threadIdx.x) / G);
  int j, o;
  int seed = -1;
  if (G == 1 && warm && second_valid) {
    // ---- a pass that follows a single-lane pass over the same clouds: most points need no search ----
    // The previous pass left, next to each match, a lower bound of the squared distance of every OTHER target point
    // (second[l], nn_search_lane_t). The point has moved by delta since: every other target point is still at least
    // sqrt(second) - delta away. If the old match, at its new exact distance, is closer than that (with margins far
    // above the fp32 rounding of the distances), it is still THE nearest neighbour by (d2, index): no search. An
    // unmatched point whose stored bound still clears the correspondence distance needs none either (warm_start).
    // The points that do are then packed into the block's first warps (a warp with 11 searching lanes costs what one
    // with 32 does: after a 5 mm move ~1/3 of the points search, and the search is bound by instruction issue).
    bool need = false;
    if (slot_of(sh, l, j, o)) {
      const float4 a = s_spts[sh.begin_of(j) + o];
      float px, py, pz;
      transform_rn(Tf, a.x, a.y, a.z, px, py, pz);
      const int pc = corr[l];
      if (pc >= 0) {
        const float4 p = t_spts[pc & kCorrIndexMask];
        const float d1 = sqdist_rn(px, py, pz, p.x, p.y, p.z);
        float ox, oy, oz;
        transform_rn(Tpf, a.x, a.y, a.z, ox, oy, oz);
        const float delta = sqrtf(sqdist_rn(px, py, pz, ox, oy, oz)) * 1.0001f;
        const float lb = sqrtf(second[l]) * 0.9999f - delta;
        if (lb > 0.f && d1 * 1.001f < lb * lb && (double)d1 < thr_sq) {
          sqd[l] = d1;          // :180, the distance the search would have found
          second[l] = lb * lb;  // the bound, carried to the new position; corr[l] (match + label bit) stays
        } else {
          need = true;
        }
      } else {
        float kept;
        int unused;
        if (warm_start(Tpf, a, px, py, pz, pc, sqd[l], thr_sq, unused, kept)) need = true;
        else sqd[l] = kept;  // provably still unmatched; corr[l] stays -1
      }
    }
    __shared__ int warp_cnt[kThreads / 32];
    __shared__ int list[kThreads];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned m = __ballot_sync(0xffffffffu, need);
    if (lane == 0) warp_cnt[warp] = __popc(m);
    __syncthreads();
    int base = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; w++) {
      base += w < warp ? warp_cnt[w] : 0;
      total += warp_cnt[w];
    }
    if (need) list[base + __popc(m & ((1u << lane) - 1u))] = l;
    __syncthreads();
    if ((int)threadIdx.x >= total) return;
    l = list[threadIdx.x];
    (void)slot_of(sh, l, j, o);
    const int pc = corr[l];
    seed = pc >= 0 ? (pc & kCorrIndexMask) : -1;
    warm = 0;  // (decided above)
  } else if (!slot_of(sh, l, j, o)) {
    return;
  }
  const int i = sh.begin_of(j) + o;
  const float4 a = s_spts[i];
  float px, py, pz;
  transform_rn(Tf, a.x, a.y, a.z, px, py, pz);  // :176

  // warm: corr / sqd still hold the previous pass over the same clouds (pose T_prev) — see warm_start
  if (warm) {
    float kept;
    if (!warm_start(Tpf, a, px, py, pz, corr[l], sqd[l], thr_sq, seed, kept)) {  // (uniform over the G lanes of a query)
      nn_group_sync<G>();
      if (G == 1 || (threadIdx.x & (G - 1)) == 0) sqd[l] = kept;  // still unmatched; corr[l] stays -1
      return;
    }
    nn_group_sync<G>();  // every lane of the group has read corr[l] / sqd[l] before lane 0 rewrites them
  }
  unsigned long long best;
  int pos;
  float proven2, second2 = 0.f;
  if (G == 1) nn_search_lane_t<true>(t_spts, t_cell_start, tg, px, py, pz, thr_sq, best, pos, seed, proven2, second2);  // :178
  else nn_search<G>(t_spts, t_cell_start, tg, px, py, pz, thr_sq, best, pos, seed, proven2);
  if (G > 1 && (threadIdx.x & (G - 1)) != 0) return;  // one lane per point finishes the job
  const float d2 = (best == kInfKey) ? 3.402823466e38f : __uint_as_float((unsigned)(best >> 32));
  const bool ok = (best != kInfKey) && ((double)d2 < thr_sq);  // :183
  if (!ok) {
    corr[l] = -1;
    sqd[l] = fminf(d2, proven2);  // (:180 stores the found distance; for a rejected point this slot is write-only scratch in the
                                  // reference — here it keeps the proven lower bound of its distance for the next warm start)
    return;
  }
  sqd[l] = d2;  // :180
  if (G == 1) second[l] = second2;
  corr[l] = pos | ((__ldg(&t_label[pos]) == s_label[i]) ? kCorrLabelBit : 0);  // label test of :271-273, hoisted
}

// ---- pass 2: radar noise covariance, combined covariance and its inverse (:194-218), one thread per matched pair:
// coalesced streams (point 16 B, correspondence 4 B, source covariance 48 B, Mahalanobis out 24 / 48 B) + one 48-byte
// gather of the target covariance; fp64 arithmetic.
template <bool kFp64>
__global__ void __launch_bounds__(256) maha_kernel(const float4* __restrict__ s_spts, const double* __restrict__ s_cov, ShardTable sh,
                                                   const double* __restrict__ t_cov, PoseD T, NoiseParams np, const int* __restrict__ corr,
                                                   void* __restrict__ mahaA, void* __restrict__ mahaB) {
  const int l = blockIdx.x * 256 + threadIdx.x;
  int j, o;
  if (!slot_of(sh, l, j, o)) return;
  const int c = corr[l];
  if (c < 0) return;  // (the reductions never read the matrix of an unmatched point)
  const int i = sh.begin_of(j) + o;
  const int pos = c & kCorrIndexMask;
  const float4 a = s_spts[i];
  const PoseF Tf = pose_to_f32(T);
  float px, py, pz;
  transform_rn(Tf, a.x, a.y, a.z, px, py, pz);
  double ca[6], cb[6];
  {
    const double2* pa = reinterpret_cast<const double2*>(s_cov + (size_t)i * 6);
    const double2* pb = reinterpret_cast<const double2*>(t_cov + (size_t)pos * 6);
    const double2 a0 = pa[0], a1 = pa[1], a2 = pa[2];
    const double2 b0 = __ldg(pb), b1 = __ldg(pb + 1), b2 = __ldg(pb + 2);
    ca[0] = a0.x; ca[1] = a0.y; ca[2] = a1.x; ca[3] = a1.y; ca[4] = a2.x; ca[5] = a2.y;
    cb[0] = b0.x; cb[1] = b0.y; cb[2] = b1.x; cb[3] = b1.y; cb[4] = b2.x; cb[5] = b2.y;
  }
  const Sym3 M = mahalanobis_of(px, py, pz, ca, cb, T, np);
  if (kFp64) {
    double2* mA = reinterpret_cast<double2*>(mahaA);
    double2* mB = reinterpret_cast<double2*>(mahaB);
    mA[l] = make_double2(M.v[0], M.v[1]);
    mB[l] = make_double2(M.v[2], M.v[3]);
    mB[(size_t)sh.plane + l] = make_double2(M.v[4], M.v[5]);
  } else {
    reinterpret_cast<float4*>(mahaA)[l] = make_float4((float)M.v[0], (float)M.v[1], (float)M.v[2], (float)M.v[3]);
    reinterpret_cast<float2*>(mahaB)[l] = make_float2((float)M.v[4], (float)M.v[5]);
  }
}

// ---- fitness -------------------------------------------------------------------
constexpr int kFitThreads = 128;
__global__ void __launch_bounds__(kFitThreads) fitness_kernel(const float4* __restrict__ s_spts, int n_src, const float4* __restrict__ t_spts,
                                                              const uint32_t* __restrict__ t_cell_start, GridDesc tg, PoseF Tf,
                                                              double max_range, double inlier_sq_thr, double* __restrict__ partials,
                                                              double* __restrict__ out3, unsigned int* __restrict__ ticket) {
  double sum = 0.0, nr = 0.0, ni = 0.0;
  for (int i = blockIdx.x * kFitThreads + threadIdx.x; i < n_src; i += gridDim.x * kFitThreads) {
    const float4 a = s_spts[i];
    float px, py, pz;
    transform_rn(Tf, a.x, a.y, a.z, px, py, pz);
    unsigned long long best;
    int pos;
    nn_search<1>(t_spts, t_cell_start, tg, px, py, pz, 1e300, best, pos);
    if (best != kInfKey) {
      const double d2 = (double)__uint_as_float((unsigned)(best >> 32));
      if (d2 <= max_range) { sum += d2; nr += 1.0; }
      if (d2 < inlier_sq_thr) ni += 1.0;
    }
  }
  __shared__ double sh[3][kFitThreads / 32];
  sum = warp_sum(sum); nr = warp_sum(nr); ni = warp_sum(ni);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { sh[0][warp] = sum; sh[1][warp] = nr; sh[2][warp] = ni; }
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) {
    double a = 0, b = 0, c = 0;
    for (int w = 0; w < kFitThreads / 32; w++) { a += sh[0][w]; b += sh[1][w]; c += sh[2][w]; }
    partials[blockIdx.x * 3 + 0] = a; partials[blockIdx.x * 3 + 1] = b; partials[blockIdx.x * 3 + 2] = c;
    __threadfence();
    const unsigned int t = atomicAdd(ticket, 1u);
    last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    double a = 0, b = 0, c = 0;
    for (unsigned int k = 0; k < gridDim.x; k++) {
      a += __ldcg(&partials[k * 3 + 0]); b += __ldcg(&partials[k * 3 + 1]); c += __ldcg(&partials[k * 3 + 2]);
    }
    out3[0] = a; out3[1] = b; out3[2] = c;
    *ticket = 0;
  }
}

// ---- per-point nearest neighbour of the transformed source (the search-method adaptor's batch) ----
__global__ void __launch_bounds__(kFitThreads) source_nearest_kernel(const float4* __restrict__ s_spts, int n_src, const float4* __restrict__ t_spts,
                                                                     const uint32_t* __restrict__ t_cell_start, GridDesc tg, PoseF Tf,
                                                                     int32_t* __restrict__ idx, float* __restrict__ d2out, float* __restrict__ xyz) {
  const int i = blockIdx.x * kFitThreads + threadIdx.x;
  if (i >= n_src) return;
  const float4 a = s_spts[i];
  const int oi = __float_as_int(a.w);
  float px, py, pz;
  transform_rn(Tf, a.x, a.y, a.z, px, py, pz);
  unsigned long long best;
  int pos;
  nn_search<1>(t_spts, t_cell_start, tg, px, py, pz, 1e300, best, pos);
  idx[oi] = best == kInfKey ? -1 : (int)(unsigned)(best & 0xffffffffull);
  d2out[oi] = best == kInfKey ? 3.402823466e38f : __uint_as_float((unsigned)(best >> 32));
  if (xyz) {
    xyz[3 * (size_t)oi + 0] = px;
    xyz[3 * (size_t)oi + 1] = py;
    xyz[3 * (size_t)oi + 2] = pz;
  }
}

// ---- export hooks ----------------------------------------------------------------
template <bool kFp64>
__global__ void __launch_bounds__(256) corr_export_kernel(const float4* __restrict__ s_spts, ShardTable sh, const float4* __restrict__ t_spts,
                                                          const int* __restrict__ corr, const float* __restrict__ sqd,
                                                          const void* __restrict__ mahaA, const void* __restrict__ mahaB,
                                                          int32_t* __restrict__ idx_out, float* __restrict__ sqd_out, double* __restrict__ maha_out) {
  const int s = blockIdx.x * 256 + threadIdx.x;  // rank-local slot
  int j, o;
  if (!slot_of(sh, s, j, o)) return;
  const int oi = __float_as_int(s_spts[sh.begin_of(j) + o].w);
  const int c = corr[s];
  if (idx_out) idx_out[oi] = c < 0 ? -1 : __float_as_int(t_spts[c & kCorrIndexMask].w);
  if (sqd_out) sqd_out[oi] = sqd[s];
  if (maha_out) {
    double m[6] = {0, 0, 0, 0, 0, 0};
    if (c >= 0) {
      if (kFp64) {
        const double2 a = reinterpret_cast<const double2*>(mahaA)[s];
        const double2 b = reinterpret_cast<const double2*>(mahaB)[s];
        const double2 d = reinterpret_cast<const double2*>(mahaB)[(size_t)sh.plane + s];
        m[0] = a.x; m[1] = a.y; m[2] = b.x; m[3] = b.y; m[4] = d.x; m[5] = d.y;
      } else {
        const float4 a = reinterpret_cast<const float4*>(mahaA)[s];
        const float2 b = reinterpret_cast<const float2*>(mahaB)[s];
        m[0] = a.x; m[1] = a.y; m[2] = a.z; m[3] = a.w; m[4] = b.x; m[5] = b.y;
      }
    }
    double* o = maha_out + (size_t)oi * 16;
    o[0] = m[0]; o[1] = m[1]; o[2] = m[2]; o[3] = 0.0;
    o[4] = m[1]; o[5] = m[3]; o[6] = m[4]; o[7] = 0.0;
    o[8] = m[2]; o[9] = m[4]; o[10] = m[5]; o[11] = 0.0;
    o[12] = 0.0; o[13] = 0.0; o[14] = 0.0; o[15] = 0.0;
  }
}

__global__ void __launch_bounds__(256) transform_cloud_kernel(const float4* __restrict__ pts, int n, PoseF T, float* __restrict__ xyz) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const float4 p = pts[i];
  float x, y, z;
  transform_rn(T, p.x, p.y, p.z, x, y, z);
  xyz[3 * (size_t)i + 0] = x;
  xyz[3 * (size_t)i + 1] = y;
  xyz[3 * (size_t)i + 2] = z;
}

}  // namespace

int launch_update_correspondences(const CloudDev& src, const CloudDev& tgt, const ShardTable& sh, const PoseD& T, const NoiseParams& np,
                                  const CorrOut& out, const PoseD* T_prev, int lanes, cudaStream_t s, int64_t* launches) {
  const int slots = sh.slots();
  if (slots <= 0) return 0;
  // small source clouds are latency-bound: 8 lanes share a query; large ones are throughput-bound (see DESIGN.md §4.3 for
  // the measured choice). All variants give identical results.
  // (Tried and dropped, 20 M points on B200: staging the union box of a warp's 32 query cubes in shared memory with bulk
  // copies and scanning it densely — 5.7 ms against 4.1 ms: in x-fastest cell order a warp's queries form a strip ~200 cells
  // long, so the box holds ~100x the 4-5 candidates a query needs.)
  if (lanes <= 0) lanes = src.n < 65536 ? 8 : kCorrLanesLarge;
  const PoseF Tf = pose_to_f32_host(T), Tpf = pose_to_f32_host(T_prev ? *T_prev : T);
#define APD_SEARCH(GG)                                                                                                                 \
  corr_search_kernel<GG><<<(unsigned)(((size_t)slots * GG + kThreads - 1) / kThreads), kThreads, 0, s>>>(                               \
      src.spts, src.label, sh, tgt.spts, tgt.label, tgt.cell_start, tgt.g, Tf, np.thr_sq, out.corr, out.sqd, T_prev ? 1 : 0, Tpf, out.second,  \
      (T_prev && out.second_valid && GG == 1) ? 1 : 0)
  switch (lanes) {
    case 1: APD_SEARCH(1); break;
    case 2: APD_SEARCH(2); break;
    case 4: APD_SEARCH(4); break;
    default: APD_SEARCH(8); break;
  }
#undef APD_SEARCH
  (*launches)++;
  const unsigned mblocks = (unsigned)((slots + 255) / 256);
  if (out.maha_fp64) maha_kernel<true><<<mblocks, 256, 0, s>>>(src.spts, src.cov, sh, tgt.cov, T, np, out.corr, out.mahaA, out.mahaB);
  else maha_kernel<false><<<mblocks, 256, 0, s>>>(src.spts, src.cov, sh, tgt.cov, T, np, out.corr, out.mahaA, out.mahaB);
  (*launches)++;
  return lanes;
}

void launch_fitness(const CloudDev& src, const CloudDev& tgt, const PoseF& T, double max_range, double inlier_sq_thr,
                    double* d_partials, int max_blocks, double* d_out3, unsigned int* d_ticket, cudaStream_t s, int64_t* launches) {
  int blocks = (src.n + kFitThreads - 1) / kFitThreads;
  blocks = max(1, min(blocks, max_blocks));
  fitness_kernel<<<blocks, kFitThreads, 0, s>>>(src.spts, src.n, tgt.spts, tgt.cell_start, tgt.g, T, max_range, inlier_sq_thr, d_partials,
                                                d_out3, d_ticket);
  (*launches)++;
}

void launch_source_nearest(const CloudDev& src, const CloudDev& tgt, const PoseF& T, int32_t* d_idx, float* d_d2, float* d_xyz, cudaStream_t s,
                           int64_t* launches) {
  if (src.n <= 0) return;
  source_nearest_kernel<<<(src.n + kFitThreads - 1) / kFitThreads, kFitThreads, 0, s>>>(src.spts, src.n, tgt.spts, tgt.cell_start, tgt.g, T, d_idx,
                                                                                      d_d2, d_xyz);
  (*launches)++;
}

void launch_corr_export(const CloudDev& src, const CloudDev& tgt, const ShardTable& sh, const CorrOut& c, int32_t* d_idx, float* d_sqd,
                        double* d_maha4x4, cudaStream_t s, int64_t* launches) {
  if (sh.slots() <= 0) return;
  const int blocks = (sh.slots() + 255) / 256;
  if (c.maha_fp64)
    corr_export_kernel<true><<<blocks, 256, 0, s>>>(src.spts, sh, tgt.spts, c.corr, c.sqd, c.mahaA, c.mahaB, d_idx, d_sqd, d_maha4x4);
  else
    corr_export_kernel<false><<<blocks, 256, 0, s>>>(src.spts, sh, tgt.spts, c.corr, c.sqd, c.mahaA, c.mahaB, d_idx, d_sqd, d_maha4x4);
  (*launches)++;
}

void launch_transform_cloud(const float4* pts, int n, const PoseF& T, float* d_xyz, cudaStream_t s, int64_t* launches) {
  if (n <= 0) return;
  transform_cloud_kernel<<<(n + 255) / 256, 256, 0, s>>>(pts, n, T, d_xyz);
  (*launches)++;
}

// Loads this file's kernels into the current context (CUDA loads kernels lazily, at their first launch, and a load may have
// to synchronise with the context: if it happens while another rank's kernel of the same process is spinning on a peer
// — the sharded exchange — neither can proceed. apd_group_create loads everything up front.)
void preload_corr_kernels() {
  cudaFuncAttributes a;
  (void)cudaFuncGetAttributes(&a, corr_search_kernel<1>);
  (void)cudaFuncGetAttributes(&a, corr_search_kernel<2>);
  (void)cudaFuncGetAttributes(&a, corr_search_kernel<4>);
  (void)cudaFuncGetAttributes(&a, corr_search_kernel<8>);
  (void)cudaFuncGetAttributes(&a, maha_kernel<false>);
  (void)cudaFuncGetAttributes(&a, maha_kernel<true>);
  (void)cudaFuncGetAttributes(&a, fitness_kernel);
  (void)cudaFuncGetAttributes(&a, source_nearest_kernel);
  (void)cudaFuncGetAttributes(&a, corr_export_kernel<false>);
  (void)cudaFuncGetAttributes(&a, corr_export_kernel<true>);
  (void)cudaFuncGetAttributes(&a, transform_cloud_kernel);
}

}  // namespace apd
