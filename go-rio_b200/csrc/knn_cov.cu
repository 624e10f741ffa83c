// K1/K2 — exact k-nearest-neighbour search on the uniform grid (one warp per
// point) fused with the neighbourhood covariance, then regularisation and the
// per-point geometric weight (one thread per point).
// Replaces FastAPDGICP::calculate_covariances (reference
// fast_apdgicp_impl.hpp:351-411) and hoists the loop-invariant JacobiSVD the
// reference redoes per point in every linearize / compute_error (:266-269,
// :330-333) to once per cloud.
#include "knn_warp.cuh"

namespace apd {

namespace {

constexpr int kThreads = 256;
constexpr int kWarpKnnThreads = 64;  // warp-per-point kernel: small CTAs, so a long search never holds back seven finished warps
using namespace knnw;

// The k neighbours of sorted point w (one (d2, index) key per lane, ascending) -> nb[w*k + j] (original point ids,
// read by cov_regularize_kernel) and, for the parity hook, neighbors[qi*k + j] at the point's ORIGINAL index qi.
__device__ __forceinline__ void emit_point(unsigned long long mykey, int lane, int k, int w, int qi, int32_t* __restrict__ nb,
                                           int32_t* __restrict__ neighbors) {
  if (lane < k) {
    const int nidx = (int)(unsigned)(mykey & 0xffffffffull);
    nb[(size_t)w * k + lane] = nidx;
    if (neighbors) neighbors[(size_t)qi * k + lane] = nidx;
  }
}

__global__ void __launch_bounds__(kWarpKnnThreads) knn_cov_kernel(const float4* __restrict__ spts,
                                                           const uint32_t* __restrict__ cell_start, GridDesc g, int n, int k,
                                                           int32_t* __restrict__ nb, int32_t* __restrict__ neighbors, int w0, int wn) {
  // warp i serves sorted point w0 + i, i < wn
  const int lane = threadIdx.x & 31;
  __shared__ unsigned long long kbuf[kWarpKnnThreads / 32][32];
  const int li = (blockIdx.x * kWarpKnnThreads + threadIdx.x) >> 5;
  if (li >= wn) return;
  const int w = w0 + li;
  const int qi = __float_as_int(spts[w].w);
  const unsigned long long mykey = knn_warp_query(spts, cell_start, g, k, w, lane, kbuf[threadIdx.x >> 5]);
  emit_point(mykey, lane, k, w, qi, nb, neighbors);
}

// exact kNN of arbitrary query points on a cloud's grid (apd_nearest_k: the search method PCL callers reach through
// pcl::Registration::getSearchMethodTarget()): one warp per query, k <= 32; out_idx / out_d2 [q*k + j], ascending by
// (d2, index); entries beyond the cloud's size read -1 / inf
__global__ void __launch_bounds__(kWarpKnnThreads) knn_query_kernel(const float4* __restrict__ spts, const uint32_t* __restrict__ cell_start,
                                                                    GridDesc g, int n, int k, const float4* __restrict__ queries, int nq,
                                                                    int32_t* __restrict__ out_idx, float* __restrict__ out_d2) {
  const int lane = threadIdx.x & 31;
  __shared__ unsigned long long kbuf[kWarpKnnThreads / 32][32];
  const int qi = (blockIdx.x * kWarpKnnThreads + threadIdx.x) >> 5;
  if (qi >= nq) return;
  const unsigned long long key = knn_warp_query_at(spts, cell_start, g, min(k, n), queries[qi], lane, kbuf[threadIdx.x >> 5]);
  if (lane < k) {
    const bool have = lane < n && key != kInfKey;
    out_idx[(size_t)qi * k + lane] = have ? (int)(unsigned)(key & 0xffffffffull) : -1;
    out_d2[(size_t)qi * k + lane] = have ? __uint_as_float((unsigned)(key >> 32)) : __int_as_float(0x7f800000);
  }
}

__global__ void __launch_bounds__(kThreads) regularize_kernel(double* __restrict__ cov, float* __restrict__ geo, double* __restrict__ geo64, int n, int reg) {
  const int i = blockIdx.x * kThreads + threadIdx.x;
  if (i >= n) return;
  Sym3 C;
#pragma unroll
  for (int e = 0; e < 6; e++) C.v[e] = cov[(size_t)i * 6 + e];
  const Sym3 out = regularize_sym3(C, reg);
#pragma unroll
  for (int e = 0; e < 6; e++) cov[(size_t)i * 6 + e] = out.v[e];
  const double gw = geo_weight_of(out);
  geo[i] = (float)gw;
  geo64[i] = gw;
}

// Covariance of the k neighbours of every point + regularisation (:374-405) + geometric weight (:266-269), one thread per
// point of the sorted cloud (knn_warp.cuh: covariance_of_neighbors). nb[w*k + j]: original ids of the neighbours of sorted
// point w, ascending by (d2, index).
__global__ void __launch_bounds__(kThreads) cov_regularize_kernel(const float4* __restrict__ pts, const int32_t* __restrict__ nb, int n, int k,
                                                                  int reg, double* __restrict__ cov, float* __restrict__ geo,
                                                                  double* __restrict__ geo64, int w0, int wn) {
  if (blockIdx.x * kThreads + threadIdx.x >= wn) return;
  const int w = w0 + blockIdx.x * kThreads + threadIdx.x;
  const Sym3 out = covariance_of_neighbors(pts, nb + (size_t)w * k, k, reg);
#pragma unroll
  for (int e = 0; e < 6; e++) cov[(size_t)w * 6 + e] = out.v[e];
  const double gw = geo_weight_of(out);
  geo[w] = (float)gw;
  geo64[w] = gw;
}

// ---------------------------------------------------------------------------
// Thread-per-point exact kNN + covariance + regularisation, fused.
// Each thread keeps its k best (d2, idx) keys in a max-heap in shared memory
// (column tid of two [k][128] uint32 arrays: conflict-free), so one warp serves
// 32 queries at once; neighbouring threads are neighbouring points of the
// cell-sorted cloud and walk nearly the same cells (L1 hits, little divergence).
// ---------------------------------------------------------------------------
constexpr int kKnnT = 128;

__device__ __forceinline__ unsigned long long heap_get(const uint32_t* hd, const uint32_t* hi, int slot, int tid) {
  return ((unsigned long long)hd[slot * kKnnT + tid] << 32) | hi[slot * kKnnT + tid];
}
__device__ __forceinline__ void heap_set(uint32_t* hd, uint32_t* hi, int slot, int tid, unsigned long long key) {
  hd[slot * kKnnT + tid] = (uint32_t)(key >> 32);
  hi[slot * kKnnT + tid] = (uint32_t)key;
}
// put `key` at the root of a max-heap of `size` slots and sift it down
__device__ __forceinline__ void heap_sift_root(uint32_t* hd, uint32_t* hi, int size, int tid, unsigned long long key) {
  int pos = 0;
  for (;;) {
    int c = 2 * pos + 1;
    if (c >= size) break;
    unsigned long long kc = heap_get(hd, hi, c, tid);
    if (c + 1 < size) {
      const unsigned long long k2 = heap_get(hd, hi, c + 1, tid);
      if (k2 > kc) { kc = k2; c = c + 1; }
    }
    if (kc <= key) break;
    heap_set(hd, hi, pos, tid, kc);
    pos = c;
  }
  heap_set(hd, hi, pos, tid, key);
}

__device__ __forceinline__ void knn_scan_range(const float4* __restrict__ spts, int b, int e, float qx, float qy, float qz,
                                               uint32_t* hd, uint32_t* hi, int k, int tid, unsigned long long& top) {
  for (int j = b; j < e; j++) {
    const float4 p = __ldg(&spts[j]);
    const unsigned long long key = pack_key(sqdist_rn(qx, qy, qz, p.x, p.y, p.z), __float_as_int(p.w));
    if (key < top) {
      heap_sift_root(hd, hi, k, tid, key);
      top = heap_get(hd, hi, 0, tid);
    }
  }
}

__global__ void __launch_bounds__(kKnnT) knn_cov_thread_kernel(const float4* __restrict__ spts, const float4* __restrict__ pts,
                                                               const uint32_t* __restrict__ cell_start, GridDesc g, int n, int k, int reg,
                                                               double* __restrict__ cov, float* __restrict__ geo, double* __restrict__ geo64,
                                                               int32_t* __restrict__ neighbors, int w0, int wn) {
  extern __shared__ uint32_t heap_smem[];
  uint32_t* hd = heap_smem;               // [k][128] d2 bits
  uint32_t* hi = heap_smem + k * kKnnT;   // [k][128] original index
  const int tid = threadIdx.x;
  if (blockIdx.x * kKnnT + tid >= wn) return;
  const int w = w0 + blockIdx.x * kKnnT + tid;
  for (int j = 0; j < k; j++) heap_set(hd, hi, j, tid, kInfKey);
  unsigned long long top = kInfKey;

  const float4 q = spts[w];
  const int qi = __float_as_int(q.w);
  const int cx = cell_coord(q.x, g.ox, g.inv_cell, g.nx);
  const int cy = cell_coord(q.y, g.oy, g.inv_cell, g.ny);
  const int cz = cell_coord(q.z, g.oz, g.inv_cell, g.nz);

  // ring 0+1: the 3x3x3 cube as up to 9 x-rows
  {
    const int x0 = max(cx - 1, 0), x1 = min(cx + 1, g.nx - 1);
    for (int z = max(cz - 1, 0); z <= min(cz + 1, g.nz - 1); z++)
      for (int y = max(cy - 1, 0); y <= min(cy + 1, g.ny - 1); y++) {
        const int row = (z * g.ny + y) * g.nx;
        knn_scan_range(spts, (int)__ldg(&cell_start[row + x0]), (int)__ldg(&cell_start[row + x1 + 1]), q.x, q.y, q.z, hd, hi, k, tid, top);
      }
  }
  // shells r = 2, 3, ...: every unscanned point is at least (r - 0.002) cells away
  const float mg = 0.002f * g.cell;
  for (int r = 1;;) {  // thick shells (r, rr], as in the warp kernel
    const float lb = ((float)r - 0.002f) * g.cell;
    if (top != kInfKey && __uint_as_float((unsigned)(top >> 32)) < lb * lb) break;
    if (cx - r <= 0 && cx + r >= g.nx - 1 && cy - r <= 0 && cy + r >= g.ny - 1 && cz - r <= 0 && cz + r >= g.nz - 1) break;
    const int rr = r < 3 ? r + 1 : r + (r >> 1) + 1;
    const int x0 = max(cx - rr, 0), x1 = min(cx + rr, g.nx - 1);
    for (int z = max(cz - rr, 0); z <= min(cz + rr, g.nz - 1); z++) {
      const bool zo = (z < cz - r) || (z > cz + r);
      const float loz = g.oz + (float)z * g.cell - mg, hiz = g.oz + (float)(z + 1) * g.cell + mg;
      const float ddz = fmaxf(0.f, fmaxf(loz - q.z, q.z - hiz));
      for (int y = max(cy - rr, 0); y <= min(cy + rr, g.ny - 1); y++) {
        const float loy = g.oy + (float)y * g.cell - mg, hiy = g.oy + (float)(y + 1) * g.cell + mg;
        const float ddy = fmaxf(0.f, fmaxf(loy - q.y, q.y - hiy));
        const float dyz2 = ddy * ddy + ddz * ddz;
        // prune the whole row if even its nearest point is farther than the current k-th distance
        if (top != kInfKey && dyz2 * 0.9999f > __uint_as_float((unsigned)(top >> 32))) continue;
        const int row = (z * g.ny + y) * g.nx;
        if (zo || y < cy - r || y > cy + r) {
          knn_scan_range(spts, (int)__ldg(&cell_start[row + x0]), (int)__ldg(&cell_start[row + x1 + 1]), q.x, q.y, q.z, hd, hi, k, tid, top);
        } else {
          const int xl = min(cx - r - 1, g.nx - 1), xr = max(cx + r + 1, 0);
          if (x0 <= xl) {
            const float hix = g.ox + (float)(xl + 1) * g.cell + mg;
            const float ddx = fmaxf(0.f, q.x - hix);
            if (top == kInfKey || !((dyz2 + ddx * ddx) * 0.9999f > __uint_as_float((unsigned)(top >> 32))))
              knn_scan_range(spts, (int)__ldg(&cell_start[row + x0]), (int)__ldg(&cell_start[row + xl + 1]), q.x, q.y, q.z, hd, hi, k, tid, top);
          }
          if (xr <= x1) {
            const float lox = g.ox + (float)xr * g.cell - mg;
            const float ddx = fmaxf(0.f, lox - q.x);
            if (top == kInfKey || !((dyz2 + ddx * ddx) * 0.9999f > __uint_as_float((unsigned)(top >> 32))))
              knn_scan_range(spts, (int)__ldg(&cell_start[row + xr]), (int)__ldg(&cell_start[row + x1 + 1]), q.x, q.y, q.z, hd, hi, k, tid, top);
          }
        }
      }
    }
    r = rr;
  }

  // heap-sort in place -> ascending (d2, idx)
  for (int m = k - 1; m >= 1; m--) {
    const unsigned long long last = heap_get(hd, hi, m, tid);
    heap_set(hd, hi, m, tid, heap_get(hd, hi, 0, tid));
    heap_sift_root(hd, hi, m, tid, last);
  }
  if (neighbors)
    for (int j = 0; j < k; j++) neighbors[(size_t)qi * k + j] = (int)hi[j * kKnnT + tid];
  if (!cov) return;

  // covariance of the k neighbours (reference :366-372): fp64, sums in neighbour order with
  // one rounding per operation (the CPU path's operation order), centred, divided by k
  double mx = 0.0, my = 0.0, mz = 0.0;
  for (int j = 0; j < k; j++) {
    const float4 p = __ldg(&pts[hi[j * kKnnT + tid]]);
    mx = __dadd_rn(mx, (double)p.x);
    my = __dadd_rn(my, (double)p.y);
    mz = __dadd_rn(mz, (double)p.z);
  }
  mx /= (double)k; my /= (double)k; mz /= (double)k;
  Sym3 C;
#pragma unroll
  for (int e = 0; e < 6; e++) C.v[e] = 0.0;
  for (int j = 0; j < k; j++) {
    const float4 p = __ldg(&pts[hi[j * kKnnT + tid]]);
    const double dx = __dsub_rn((double)p.x, mx), dy = __dsub_rn((double)p.y, my), dz = __dsub_rn((double)p.z, mz);
    C.v[0] = __dadd_rn(C.v[0], __dmul_rn(dx, dx));
    C.v[1] = __dadd_rn(C.v[1], __dmul_rn(dx, dy));
    C.v[2] = __dadd_rn(C.v[2], __dmul_rn(dx, dz));
    C.v[3] = __dadd_rn(C.v[3], __dmul_rn(dy, dy));
    C.v[4] = __dadd_rn(C.v[4], __dmul_rn(dy, dz));
    C.v[5] = __dadd_rn(C.v[5], __dmul_rn(dz, dz));
  }
#pragma unroll
  for (int e = 0; e < 6; e++) C.v[e] /= (double)k;
  const Sym3 out = regularize_sym3(C, reg);
#pragma unroll
  for (int e = 0; e < 6; e++) cov[(size_t)w * 6 + e] = out.v[e];
  const double gw = geo_weight_of(out);
  geo[w] = (float)gw;
  geo64[w] = gw;
}

__global__ void __launch_bounds__(kThreads) geo_weight_kernel(const double* __restrict__ cov, float* __restrict__ geo, double* __restrict__ geo64, int n) {
  const int i = blockIdx.x * kThreads + threadIdx.x;
  if (i >= n) return;
  Sym3 C;
#pragma unroll
  for (int e = 0; e < 6; e++) C.v[e] = cov[(size_t)i * 6 + e];
  const double gw = geo_weight_of(C);
  geo[i] = (float)gw;
  geo64[i] = gw;
}

// sorted sym6 -> column-major 4x4 (128 B) at the ORIGINAL index
__global__ void __launch_bounds__(kThreads) cov_export_kernel(const double* __restrict__ cov, const float4* __restrict__ spts, int n,
                                                              double* __restrict__ out) {
  const int s = blockIdx.x * kThreads + threadIdx.x;
  if (s >= n) return;
  const int idx = __float_as_int(spts[s].w);
  const double* c = cov + (size_t)s * 6;
  double* o = out + (size_t)idx * 16;
  o[0] = c[0]; o[1] = c[1]; o[2] = c[2]; o[3] = 0.0;
  o[4] = c[1]; o[5] = c[3]; o[6] = c[4]; o[7] = 0.0;
  o[8] = c[2]; o[9] = c[4]; o[10] = c[5]; o[11] = 0.0;
  o[12] = 0.0; o[13] = 0.0; o[14] = 0.0; o[15] = 0.0;
}

// column-major 4x4 at the original index -> sorted sym6 (symmetrised)
__global__ void __launch_bounds__(kThreads) cov_import_kernel(const double* __restrict__ in, const float4* __restrict__ spts, int n,
                                                              double* __restrict__ cov) {
  const int s = blockIdx.x * kThreads + threadIdx.x;
  if (s >= n) return;
  const int idx = __float_as_int(spts[s].w);
  const double* m = in + (size_t)idx * 16;
  double* c = cov + (size_t)s * 6;
  c[0] = m[0];
  c[1] = 0.5 * (m[1] + m[4]);
  c[2] = 0.5 * (m[2] + m[8]);
  c[3] = m[5];
  c[4] = 0.5 * (m[6] + m[9]);
  c[5] = m[10];
}

}  // namespace

void launch_knn_cov_fused(const CloudDev& c, int k, int regularization, int32_t* neighbors, cudaStream_t s, int64_t* launches, int w0,
                          int wn) {
  if (wn < 0) wn = c.n - w0;
  if (c.n <= 0 || wn <= 0) return;
  const size_t smem = (size_t)2 * k * kKnnT * sizeof(uint32_t);
  if (smem > 48 * 1024) cudaFuncSetAttribute(knn_cov_thread_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  knn_cov_thread_kernel<<<(wn + kKnnT - 1) / kKnnT, kKnnT, smem, s>>>(c.spts, c.pts, c.cell_start, c.g, c.n, k, regularization, c.cov, c.geo,
                                                                       c.geo64, neighbors, w0, wn);
  (*launches)++;
}
void launch_knn_cov(const CloudDev& c, int k, int32_t* d_nb, int32_t* neighbors, cudaStream_t s, int64_t* launches, int w0, int wn) {
  if (wn < 0) wn = c.n - w0;
  if (c.n <= 0 || wn <= 0) return;
  const long long threads = (long long)wn * 32;
  const int blocks = (int)((threads + kWarpKnnThreads - 1) / kWarpKnnThreads);
  knn_cov_kernel<<<blocks, kWarpKnnThreads, 0, s>>>(c.spts, c.cell_start, c.g, c.n, k, d_nb, neighbors, w0, wn);
  (*launches)++;
}
void launch_knn_query(const CloudDev& c, int k, const float4* d_queries, int nq, int32_t* d_idx, float* d_d2, cudaStream_t s, int64_t* launches) {
  if (c.n <= 0 || nq <= 0) return;
  const long long threads = (long long)nq * 32;
  knn_query_kernel<<<(unsigned)((threads + kWarpKnnThreads - 1) / kWarpKnnThreads), kWarpKnnThreads, 0, s>>>(c.spts, c.cell_start, c.g, c.n, k,
                                                                                                        d_queries, nq, d_idx, d_d2);
  (*launches)++;
}
void launch_cov_regularize(const CloudDev& c, int k, int regularization, const int32_t* d_nb, cudaStream_t s, int64_t* launches, int w0, int wn) {
  if (wn < 0) wn = c.n - w0;
  if (c.n <= 0 || wn <= 0) return;
  cov_regularize_kernel<<<(wn + kThreads - 1) / kThreads, kThreads, 0, s>>>(c.pts, d_nb, c.n, k, regularization, c.cov, c.geo, c.geo64, w0, wn);
  (*launches)++;
}
void launch_regularize(const CloudDev& c, int regularization, cudaStream_t s, int64_t* launches) {
  if (c.n <= 0) return;
  regularize_kernel<<<(c.n + kThreads - 1) / kThreads, kThreads, 0, s>>>(c.cov, c.geo, c.geo64, c.n, regularization);
  (*launches)++;
}
void launch_geo_weight(const CloudDev& c, cudaStream_t s, int64_t* launches) {
  if (c.n <= 0) return;
  geo_weight_kernel<<<(c.n + kThreads - 1) / kThreads, kThreads, 0, s>>>(c.cov, c.geo, c.geo64, c.n);
  (*launches)++;
}
void launch_cov_export(const CloudDev& c, double* d_out4x4, cudaStream_t s, int64_t* launches) {
  if (c.n <= 0) return;
  cov_export_kernel<<<(c.n + kThreads - 1) / kThreads, kThreads, 0, s>>>(c.cov, c.spts, c.n, d_out4x4);
  (*launches)++;
}
void launch_cov_import(const CloudDev& c, const double* d_in4x4, cudaStream_t s, int64_t* launches) {
  if (c.n <= 0) return;
  cov_import_kernel<<<(c.n + kThreads - 1) / kThreads, kThreads, 0, s>>>(d_in4x4, c.spts, c.n, c.cov);
  (*launches)++;
}

// Loads this file's kernels into the current context (CUDA loads kernels lazily, at their first launch, and a load may have
// to synchronise with the context: if it happens while another rank's kernel of the same process is spinning on a peer
// — the sharded exchange — neither can proceed. apd_group_create loads everything up front.)
void preload_knn_kernels() {
  cudaFuncAttributes a;
  (void)cudaFuncGetAttributes(&a, knn_cov_kernel);
  (void)cudaFuncGetAttributes(&a, knn_query_kernel);
  (void)cudaFuncGetAttributes(&a, regularize_kernel);
  (void)cudaFuncGetAttributes(&a, cov_regularize_kernel);
  (void)cudaFuncGetAttributes(&a, knn_cov_thread_kernel);
  (void)cudaFuncGetAttributes(&a, geo_weight_kernel);
  (void)cudaFuncGetAttributes(&a, cov_export_kernel);
  (void)cudaFuncGetAttributes(&a, cov_import_kernel);
}

}  // namespace apd
