// The prologue of the fused registration kernel (lm.cu): everything between "two raw clouds in HBM" and "the optimizer
// loop can start", done by the loop's own thread-block cluster inside the SAME launch —
//   bounding boxes of both clouds  ->  grid sizing (grid_desc.hpp: the host's function, same bits)  ->  both uniform
//   grids (cell-sorted layout: stable by (cell, original index), the layout grid.cu builds)  ->  kNN + covariance +
//   regularisation + geometric weight of the SOURCE points (knn_warp.cuh: the per-cloud kernels' code)
// — so that a pooled registration is ONE launch and no host round trip (round 1: 11 launches and a wait for the boxes;
// a pool's rate was bound by the launches the host process can issue, not by the GPU). Target covariances are computed
// on demand inside the loop (lm.cu: corr_phase), as before.
//
// Phases are separated by cluster barriers (barrier.cluster arrive.release / wait.acquire: what one CTA wrote to global
// memory before the barrier is visible to the others after it). Arrays written here are read in later phases through
// ordinary (coherent) loads only — never through the read-only path (__ldg / const __restrict__).
#pragma once
#include <cooperative_groups.h>

#include "knn_warp.cuh"

namespace apd {
namespace prep {

namespace cg = cooperative_groups;

__device__ __forceinline__ unsigned int f2ord(float f) {
  unsigned int u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned int u) {
  u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
  return __uint_as_float(u);
}

// scratch of the prologue in the CTA's shared memory (the loop's own shared state is not live yet: it may alias)
struct PrepShared {
  unsigned int box[2][6];     // ordered-uint min / max of {source, target}: this CTA's, then the cluster's
  GridDesc grid[2];           // {source, target}
  int ncells[2];
  unsigned int warp_tot[32];  // scan: totals of this CTA's warps
  unsigned int cta_tot;       // ... and of the CTA
  int status;                 // 0 ok, 1 a grid does not fit the arrays it was given
};

// what one cloud's grid is built from / into
struct GridJob {
  const float4* pts;  // original order {x,y,z,label}
  int n;
  uint32_t* cell_start;  // [cap]
  int cap;
  uint32_t* keys;  // [n] scratch
  uint32_t* rank;  // [n] scratch
  uint32_t* tmp;   // [n] scratch
  float4* spts;
  float* label;
  int* inv_perm;
  unsigned char* zero_flags;  // optional [n]
  // ordered: points of a cell in ascending original index (the deterministic layout of grid.cu). The order of the SOURCE
  // decides the order of every sum over source points; the order of the TARGET inside a cell decides nothing that leaves
  // the library (searches select by the key (d2, original index); covariances, correspondences and neighbour lists are
  // reported by original index) — so the target is placed by its atomic ranks, without the counting pass that costs
  // O(points per cell) per point (5.5 % of the fused kernel's instructions at 6.6 active lanes).
  bool ordered;
};

// ---- bounding boxes of both clouds: every thread strides over the points, warp min/max, CTA atomics in shared memory,
// then every CTA folds in the boxes of its peers through distributed shared memory: all hold the same box
__device__ __forceinline__ void bounds_phase(cg::cluster_group& cluster, PrepShared& ps, const float4* const pts[2], const int n[2], int gt, int GT) {
  const int tid = threadIdx.x;
  if (tid < 12) (&ps.box[0][0])[tid] = (tid % 6) < 3 ? 0xffffffffu : 0u;
  __syncthreads();
  for (int c = 0; c < 2; c++) {
    float mn[3] = {3.4e38f, 3.4e38f, 3.4e38f}, mx[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
    for (int i = gt; i < n[c]; i += GT) {
      const float4 p = pts[c][i];
      mn[0] = fminf(mn[0], p.x); mx[0] = fmaxf(mx[0], p.x);
      mn[1] = fminf(mn[1], p.y); mx[1] = fmaxf(mx[1], p.y);
      mn[2] = fminf(mn[2], p.z); mx[2] = fmaxf(mx[2], p.z);
    }
#pragma unroll
    for (int a = 0; a < 3; a++) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
        mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
      }
    }
    if ((tid & 31) == 0) {
#pragma unroll
      for (int a = 0; a < 3; a++) {
        atomicMin(&ps.box[c][a], f2ord(mn[a]));
        atomicMax(&ps.box[c][3 + a], f2ord(mx[a]));
      }
    }
  }
  cluster.sync();
  unsigned int mine = 0;
  if (tid < 12) {
    const bool is_min = (tid % 6) < 3;
    mine = (&ps.box[0][0])[tid];
    for (unsigned r = 0; r < cluster.num_blocks(); r++) {
      const unsigned int v = cluster.map_shared_rank(&ps.box[0][0], r)[tid];
      mine = is_min ? min(mine, v) : max(mine, v);
    }
  }
  cluster.sync();  // every CTA has read its peers' boxes before anyone overwrites its own
  if (tid < 12) (&ps.box[0][0])[tid] = mine;
  __syncthreads();
}

// ---- exclusive scan of data[0 .. n) in place by the whole cluster. Warp w of W owns a contiguous chunk (a multiple of 32
// entries): pass 1 sums it (coalesced), the warp totals are combined (CTA: shared memory; cluster: distributed shared
// memory), pass 2 scans the chunk 32 entries at a time with a running carry.
__device__ __forceinline__ void scan_phase(cg::cluster_group& cluster, PrepShared& ps, uint32_t* data, int n) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wpc = blockDim.x >> 5;  // warps per CTA
  const int W = wpc * (int)cluster.num_blocks();
  const int gw = (int)cluster.block_rank() * wpc + warp;
  const int chunk = (((n + W - 1) / W) + 31) & ~31;
  const int b = min(gw * chunk, n), e = min(b + chunk, n);
  uint32_t sum = 0;
  for (int j = b + lane; j < e; j += 32) sum += __ldcg(&data[j]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if (lane == 0) ps.warp_tot[warp] = sum;
  __syncthreads();
  if (tid == 0) {
    uint32_t t = 0;
    for (int w = 0; w < wpc; w++) t += ps.warp_tot[w];
    ps.cta_tot = t;
  }
  cluster.sync();
  uint32_t carry = 0;
  for (unsigned r = 0; r < cluster.block_rank(); r++) carry += *cluster.map_shared_rank(&ps.cta_tot, r);
  for (int w = 0; w < warp; w++) carry += ps.warp_tot[w];
  for (int j0 = b; j0 < e; j0 += 32) {
    const int j = j0 + lane;
    const uint32_t v = j < e ? __ldcg(&data[j]) : 0u;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (j < e) __stcg(&data[j], carry + inc - v);
    carry += __shfl_sync(0xffffffffu, inc, 31);
  }
  cluster.sync();  // the scanned counters are complete (and nobody reads cta_tot / warp_tot of this pass any more)
}

// ---- both grids, phase by phase (the clouds share the barriers). The steps of grid.cu's small-cloud path: atomic
// per-cell ranks place the points of a cell in arbitrary order, then every point counts the points of its cell with a
// smaller original index to find its deterministic slot.
__device__ __forceinline__ void grids_phase(cg::cluster_group& cluster, PrepShared& ps, const GridJob job[2], int first, int last, int gt, int GT) {
  for (int c = first; c < last; c++) {
    const int ncs = ps.ncells[c] + 1;
    for (int j = gt; j < ncs; j += GT) __stcg(&job[c].cell_start[j], 0u);
  }
  cluster.sync();
  // (kU points per thread and turn: the loads, then the atomics, then the stores — each step of a point waits a memory
  // round trip for the one before, and under a pool's load a round trip is ~1 us; four in flight per thread instead of one)
  constexpr int kU = 4;
  for (int c = first; c < last; c++) {
    const GridDesc g = ps.grid[c];
    const int n = job[c].n;
    for (int i0 = gt; i0 < n; i0 += kU * GT) {
      float4 p[kU];
#pragma unroll
      for (int u = 0; u < kU; u++) {
        const int i = i0 + u * GT;
        p[u] = job[c].pts[i < n ? i : i0];
      }
      uint32_t key[kU], rk[kU];
#pragma unroll
      for (int u = 0; u < kU; u++) {
        const int cx = cell_coord(p[u].x, g.ox, g.inv_cell, g.nx);
        const int cy = cell_coord(p[u].y, g.oy, g.inv_cell, g.ny);
        const int cz = cell_coord(p[u].z, g.oz, g.inv_cell, g.nz);
        key[u] = (uint32_t)((cz * g.ny + cy) * g.nx + cx);
      }
#pragma unroll
      for (int u = 0; u < kU; u++) rk[u] = (i0 + u * GT < n) ? atomicAdd(&job[c].cell_start[key[u]], 1u) : 0u;
#pragma unroll
      for (int u = 0; u < kU; u++) {
        const int i = i0 + u * GT;
        if (i < n) {
          job[c].keys[i] = key[u];  // (read back by the same thread only)
          job[c].rank[i] = rk[u];
        }
      }
    }
  }
  cluster.sync();
  for (int c = first; c < last; c++) scan_phase(cluster, ps, job[c].cell_start, ps.ncells[c] + 1);
  for (int c = first; c < last; c++)
    if (job[c].ordered)
      for (int i = gt; i < job[c].n; i += GT) __stcg(&job[c].tmp[__ldcg(&job[c].cell_start[job[c].keys[i]]) + job[c].rank[i]], (uint32_t)i);
  cluster.sync();
  for (int c = first; c < last; c++) {
    const int n = job[c].n;
    if (job[c].ordered) {
      for (int i = gt; i < n; i += GT) {
        const uint32_t key = job[c].keys[i];
        const uint32_t b = __ldcg(&job[c].cell_start[key]);
        const uint32_t e = __ldcg(&job[c].cell_start[key + 1]);
        uint32_t r = 0;
        for (uint32_t j = b; j < e; j++) r += (__ldcg(&job[c].tmp[j]) < (uint32_t)i) ? 1u : 0u;
        const int sp = (int)(b + r);
        const float4 p = job[c].pts[i];
        job[c].spts[sp] = make_float4(p.x, p.y, p.z, __int_as_float(i));
        job[c].label[sp] = p.w;
        job[c].inv_perm[i] = sp;
        if (job[c].zero_flags) job[c].zero_flags[i] = 0;
      }
      continue;
    }
    for (int i0 = gt; i0 < n; i0 += kU * GT) {  // (placed by the atomic ranks: kU points per thread and turn, see above)
      uint32_t key[kU], r[kU], b[kU];
      float4 p[kU];
#pragma unroll
      for (int u = 0; u < kU; u++) {
        const int i = i0 + u * GT < n ? i0 + u * GT : i0;
        key[u] = job[c].keys[i];
        r[u] = job[c].rank[i];
        p[u] = job[c].pts[i];
      }
#pragma unroll
      for (int u = 0; u < kU; u++) b[u] = __ldcg(&job[c].cell_start[key[u]]);
#pragma unroll
      for (int u = 0; u < kU; u++) {
        const int i = i0 + u * GT;
        if (i < n) {
          const int sp = (int)(b[u] + r[u]);
          job[c].spts[sp] = make_float4(p[u].x, p[u].y, p[u].z, __int_as_float(i));
          job[c].label[sp] = p[u].w;
          job[c].inv_perm[i] = sp;
          if (job[c].zero_flags) job[c].zero_flags[i] = 0;
        }
      }
    }
  }
  cluster.sync();
}

// ---- covariances of the source cloud (calculate_covariances(source), fast_apdgicp_impl.hpp:351-411): one WARP per
// point for the exact kNN search, then one THREAD per point for covariance + regularisation + geometric weight — the
// two per-cloud kernels of knn_cov.cu, with a cluster barrier between them
__device__ __forceinline__ void source_cov_phase(cg::cluster_group& cluster, const float4* pts, const float4* spts, const uint32_t* cell_start,
                                                 const GridDesc& g, int n, int k, int reg, int gicp, int32_t* nb, double* cov, float* geo, double* geo64,
                                                 unsigned long long* kbuf, int gt, int GT) {
  const int lane = threadIdx.x & 31;
  for (int w = gt >> 5; w < n; w += GT >> 5) {
    const unsigned long long key = knnw::knn_warp_query(spts, cell_start, g, k, w, lane, kbuf);
    if (lane < k) nb[(size_t)w * k + lane] = (int)(unsigned)(key & 0xffffffffull);
  }
  cluster.sync();
  for (int w = gt; w < n; w += GT) {
    const Sym3 out = knnw::covariance_of_neighbors(pts, nb + (size_t)w * k, k, reg);
#pragma unroll
    for (int e = 0; e < 6; e++) cov[(size_t)w * 6 + e] = out.v[e];
    const double gw = gicp ? 0.0 : knnw::geo_weight_of(out);  // FastGICP weighs every term by 1 (fast_gicp_impl.hpp:205)
    geo[w] = (float)gw;
    geo64[w] = gw;
  }
  cluster.sync();
}

}  // namespace prep
}  // namespace apd
