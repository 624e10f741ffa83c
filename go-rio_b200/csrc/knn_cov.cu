// K1/K2 — exact k-nearest-neighbour search on the uniform grid (one warp per
// point) fused with the neighbourhood covariance, then regularisation and the
// per-point geometric weight (one thread per point).
// Replaces FastAPDGICP::calculate_covariances (reference
// fast_apdgicp_impl.hpp:351-411) and hoists the loop-invariant JacobiSVD the
// reference redoes per point in every linearize / compute_error (:266-269,
// :330-333) to once per cloud.
#include "kernels.cuh"

namespace apd {

namespace {

constexpr int kThreads = 256;
constexpr int kWarpKnnThreads = 64;  // warp-per-point kernel: small CTAs, so a long search never holds back seven finished warps
constexpr unsigned kFull = 0xffffffffu;
constexpr unsigned long long kInfKey = 0xffffffffffffffffull;

__device__ __forceinline__ unsigned long long shfl64(unsigned long long v, int src) {
  return __shfl_sync(kFull, v, src);
}
__device__ __forceinline__ unsigned long long shfl_up64(unsigned long long v, int d) {
  return __shfl_up_sync(kFull, v, d);
}

__device__ __forceinline__ unsigned long long umin64(unsigned long long a, unsigned long long b) { return a < b ? a : b; }
__device__ __forceinline__ unsigned long long umax64(unsigned long long a, unsigned long long b) { return a < b ? b : a; }

// ascending bitonic sort of one 64-bit key per lane
__device__ __forceinline__ unsigned long long bitonic_sort32(unsigned long long key, int lane) {
#pragma unroll
  for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      const unsigned long long other = __shfl_xor_sync(kFull, key, stride);
      const bool up = (lane & size) == 0;        // this block sorts ascending
      const bool lower = (lane & stride) == 0;   // this lane is the lower one of the pair
      key = (lower == up) ? umin64(key, other) : umax64(key, other);
    }
  }
  return key;
}
// list (ascending, one key per lane) <- the 32 smallest of list U batch (batch ascending)
__device__ __forceinline__ unsigned long long merge_keep32(unsigned long long list, unsigned long long batch, int lane) {
  const unsigned long long rb = shfl64(batch, 31 - lane);
  unsigned long long c = umin64(list, rb);  // bitonic sequence holding the 32 smallest
#pragma unroll
  for (int stride = 16; stride > 0; stride >>= 1) {
    const unsigned long long other = __shfl_xor_sync(kFull, c, stride);
    c = ((lane & stride) == 0) ? umin64(c, other) : umax64(c, other);
  }
  return c;
}

// Per-warp state of the k-best search: `list` holds the 32 smallest keys seen so far
// (ascending over the lanes; the answer is lanes 0..k-1), `kth` the k-th of them, and
// `buf` (shared memory, 32 keys) collects candidates below `kth` until they are merged
// in one bitonic sort + merge — ~6 instructions per candidate instead of ~22 for
// inserting them one by one.
struct KBest {
  unsigned long long list, kth;
  unsigned long long* buf;
  int buf_n;
  bool empty;  // list holds no key yet
};
__device__ __forceinline__ void kbest_flush(KBest& s, int lane, int k) {
  if (s.buf_n == 0) return;  // warp-uniform
  __syncwarp();
  unsigned long long b = lane < s.buf_n ? s.buf[lane] : kInfKey;
  __syncwarp();
  b = bitonic_sort32(b, lane);
  s.list = s.empty ? b : merge_keep32(s.list, b, lane);  // nothing to merge with on the first flush
  s.empty = false;
  s.kth = shfl64(s.list, k - 1);
  s.buf_n = 0;
}

// Warp-wide candidate scan. Every lane passes one segment [b, b+cnt) of the
// cell-sorted point array (cnt may be 0). The segments are flattened so that
// all 32 lanes test candidates even when segments are short (sparse cells).
__device__ __forceinline__ void scan_segments(const float4* __restrict__ spts, float qx, float qy, float qz, int lane, int k,
                                              int b, int cnt, KBest& s) {
  int incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(kFull, incl, o);
    if (lane >= o) incl += t;
  }
  const int total = __shfl_sync(kFull, incl, 31);
  const int excl = incl - cnt;
  const unsigned lt = (1u << lane) - 1u;
  for (int base = 0; base < total; base += 32) {
    const int t = base + lane;
    const bool valid = t < total;
    const int tt = valid ? t : total - 1;
    // segment containing flat index tt: first lane whose inclusive prefix > tt
    int j = 0;
#pragma unroll
    for (int step = 16; step > 0; step >>= 1) {
      const int v = __shfl_sync(kFull, incl, j + step - 1);
      if (v <= tt) j += step;
    }
    const int bj = __shfl_sync(kFull, b, j);
    const int ej = __shfl_sync(kFull, excl, j);
    const float4 p = spts[bj + (tt - ej)];
    const float d2 = sqdist_rn(qx, qy, qz, p.x, p.y, p.z);
    const unsigned long long key = pack_key(d2, __float_as_int(p.w));
    bool pass = valid && key < s.kth;
    unsigned mask = __ballot_sync(kFull, pass);
    int m = __popc(mask);
    if (m == 0) continue;
    if (s.buf_n + m > 32) {
      kbest_flush(s, lane, k);
      pass = pass && key < s.kth;  // the threshold just dropped
      mask = __ballot_sync(kFull, pass);
      m = __popc(mask);
    }
    if (pass) s.buf[s.buf_n + __popc(mask & lt)] = key;
    s.buf_n += m;
  }
}

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
                                                           int32_t* __restrict__ nb, int32_t* __restrict__ neighbors,
                                                           const int* __restrict__ list, const int* __restrict__ list_n, int w0, int wn) {
  // list == nullptr: warp i serves sorted point w0 + i (+ stride), i < wn; else: the sorted positions in list[0 .. *list_n)
  const int lane = threadIdx.x & 31;
  const int count = list ? min(*list_n, n) : wn;
  const int nwarps = (int)((gridDim.x * kWarpKnnThreads) >> 5);
  __shared__ unsigned long long kbuf[kWarpKnnThreads / 32][32];
  for (int li = (blockIdx.x * kWarpKnnThreads + threadIdx.x) >> 5; li < count; li += nwarps) {
  const int w = list ? list[li] : w0 + li;
  const float4 q = spts[w];
  const int qi = __float_as_int(q.w);
  const int cx = cell_coord(q.x, g.ox, g.inv_cell, g.nx);
  const int cy = cell_coord(q.y, g.oy, g.inv_cell, g.ny);
  const int cz = cell_coord(q.z, g.oz, g.inv_cell, g.nz);

  KBest st;
  st.list = kInfKey;
  st.kth = kInfKey;
  st.buf = kbuf[threadIdx.x >> 5];
  st.buf_n = 0;
  st.empty = true;

  // ring 0+1: the 3x3x3 cube as 9 x-rows
  {
    int b = 0, cnt = 0;
    if (lane < 9) {
      // rows nearest first (centre, the 4 face rows, the 4 corner rows) so that the first
      // 32 candidates already give a tight k-th distance: packed 2-bit (dy+1, dz+1) codes
      // dy = {0,-1,1,0,0,-1,1,-1,1}[lane], dz = {0,0,0,-1,1,-1,-1,1,1}[lane]
      const int y = cy + (int)((0x22161u >> (2 * lane)) & 3u) - 1, z = cz + (int)((0x28215u >> (2 * lane)) & 3u) - 1;
      if (y >= 0 && y < g.ny && z >= 0 && z < g.nz) {
        const int x0 = max(cx - 1, 0), x1 = min(cx + 1, g.nx - 1);
        const int row = (z * g.ny + y) * g.nx;
        b = (int)cell_start[row + x0];
        cnt = (int)cell_start[row + x1 + 1] - b;
      }
    }
    scan_segments(spts, q.x, q.y, q.z, lane, k, b, cnt, st);
    kbest_flush(st, lane, k);
  }
  // shells r = 2, 3, ... until the k-th distance is provably final:
  // every unscanned point is at least (r - 0.002) cells away.
  // Thick shells: after the cube of radius r has been scanned, the next step scans radius (r, rr]. One cell at a
  // time near the query (where most searches end), then growing by ~1.5x: an x-row segment costs the same however
  // many cells it spans, so a search that must reach R cells costs O(R^2) segments instead of O(R^3).
  for (int r = 1;;) {
    const float lb = ((float)r - 0.002f) * g.cell;
    const float kd2 = __uint_as_float((unsigned)(st.kth >> 32));
    if (st.kth != kInfKey && kd2 < lb * lb) break;
    if (cx - r <= 0 && cx + r >= g.nx - 1 && cy - r <= 0 && cy + r >= g.ny - 1 && cz - r <= 0 && cz + r >= g.nz - 1) break;
    const int rr = r < 3 ? r + 1 : r + (r >> 1) + 1;  // outer radius of the shell to scan now
    const int side = 2 * rr + 1;
    const int nslots = 2 * side * side;
    const float inv_side = 1.0f / (float)side;
    for (int sbase = 0; sbase < nslots; sbase += 32) {
      const int slot = sbase + lane;
      const float kd2cur = __uint_as_float((unsigned)(st.kth >> 32));  // shrinks as the shell is scanned
      int b = 0, cnt = 0;
      if (slot < nslots) {
        const int rowid = slot >> 1, which = slot & 1;
        const int qz = __float2int_rd(((float)rowid + 0.5f) * inv_side);  // rowid / side (exact for these small integers)
        const int dy = rowid - qz * side - rr, dz = qz - rr;
        const int y = cy + dy, z = cz + dz;
        if (y >= 0 && y < g.ny && z >= 0 && z < g.nz) {
          const bool outer = (dy > r) || (dy < -r) || (dz > r) || (dz < -r);  // row lies outside the scanned cube
          const int row = (z * g.ny + y) * g.nx;
          int x0 = 1, x1 = 0;
          if (outer) {
            if (which == 0) { x0 = max(cx - rr, 0); x1 = min(cx + rr, g.nx - 1); }
          } else if (which == 0) {
            x0 = max(cx - rr, 0); x1 = min(cx - r - 1, g.nx - 1);
          } else {
            x0 = max(cx + r + 1, 0); x1 = min(cx + rr, g.nx - 1);
          }
          if (x0 <= x1) {
            // prune the segment if its (slightly grown) box is farther than the current k-th distance
            const float m = 0.002f * g.cell;
            const float lox = g.ox + (float)x0 * g.cell - m, hix = g.ox + (float)(x1 + 1) * g.cell + m;
            const float loy = g.oy + (float)y * g.cell - m, hiy = g.oy + (float)(y + 1) * g.cell + m;
            const float loz = g.oz + (float)z * g.cell - m, hiz = g.oz + (float)(z + 1) * g.cell + m;
            const float ddx = fmaxf(0.f, fmaxf(lox - q.x, q.x - hix));
            const float ddy = fmaxf(0.f, fmaxf(loy - q.y, q.y - hiy));
            const float ddz = fmaxf(0.f, fmaxf(loz - q.z, q.z - hiz));
            const float mind2 = (ddx * ddx + ddy * ddy + ddz * ddz) * 0.9999f;
            if (st.kth == kInfKey || !(mind2 > kd2cur)) {
              b = (int)cell_start[row + x0];
              cnt = (int)cell_start[row + x1 + 1] - b;
            }
          }
        }
      }
      if (__ballot_sync(kFull, cnt > 0)) scan_segments(spts, q.x, q.y, q.z, lane, k, b, cnt, st);
    }
    kbest_flush(st, lane, k);
    r = rr;
  }
  const unsigned long long mykey = st.list;

  emit_point(mykey, lane, k, w, qi, nb, neighbors);
  __syncwarp();
  }  // list / stride loop
}

__device__ __forceinline__ double geo_weight_of(const Sym3& C) {
  double l[3], V[9];
  jacobi_eig3(C, l, V);
  return fabs(l[2]) / fabs(l[0]);  // sigma3 / sigma1 (reference :268-269)
}

// regularisation of one covariance (reference :374-405), symmetric storage
__device__ __forceinline__ Sym3 regularize_sym3(const Sym3& C, int reg) {
  Sym3 out = C;
  if (reg == 0) {  // NONE (:374-376)
  } else if (reg == 4) {  // FROBENIUS (:377-383)
    Sym3 Cl = C;
    Cl.v[0] += 1e-3; Cl.v[3] += 1e-3; Cl.v[5] += 1e-3;
    Sym3 Ci = sym_inverse(Cl);
    const double nrm = sqrt(Ci.v[0] * Ci.v[0] + Ci.v[3] * Ci.v[3] + Ci.v[5] * Ci.v[5] +
                            2.0 * (Ci.v[1] * Ci.v[1] + Ci.v[2] * Ci.v[2] + Ci.v[4] * Ci.v[4]));
#pragma unroll
    for (int e = 0; e < 6; e++) Ci.v[e] /= nrm;
    out = sym_inverse(Ci);
  } else {  // SVD-based (:384-407); symmetric PSD input, so U = V up to the sign of negative eigenvalues
    double l[3], V[9];
    jacobi_eig3(C, l, V);
    double val[3], sg[3];
#pragma unroll
    for (int e = 0; e < 3; e++) sg[e] = l[e] < 0.0 ? -1.0 : 1.0;
    const double s0 = fabs(l[0]);
    if (reg == 3) { val[0] = 1.0; val[1] = 1.0; val[2] = 1e-3; }                                   // PLANE
    else if (reg == 1) { for (int e = 0; e < 3; e++) val[e] = fmax(fabs(l[e]), 1e-3); }              // MIN_EIG
    else { for (int e = 0; e < 3; e++) val[e] = fmax(fabs(l[e]) / s0, 1e-3); }                       // NORMALIZED_MIN_EIG
    // U diag(val) V^T with U = V * diag(sg); stored symmetric (upper triangle)
    const int R[6] = {0, 0, 0, 1, 1, 2}, Cc[6] = {0, 1, 2, 1, 2, 2};
#pragma unroll
    for (int e = 0; e < 6; e++) {
      double s = 0.0;
#pragma unroll
      for (int j = 0; j < 3; j++) s += sg[j] * V[R[e] * 3 + j] * val[j] * V[Cc[e] * 3 + j];
      out.v[e] = s;
    }
  }
  return out;
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

// Covariance of the k neighbours of every point (reference :366-372: fp64, centred, divided by k; sums in neighbour
// order with one rounding per operation — the CPU path's operation order, so the raw covariance is bit-identical to
// the oracle's) + regularisation (:374-405) + geometric weight (:266-269), one thread per point of the sorted cloud.
// nb[w*k + j]: original ids of the neighbours of sorted point w, ascending by (d2, index).
__global__ void __launch_bounds__(kThreads) cov_regularize_kernel(const float4* __restrict__ pts, const int32_t* __restrict__ nb, int n, int k,
                                                                  int reg, double* __restrict__ cov, float* __restrict__ geo,
                                                                  double* __restrict__ geo64, int w0, int wn) {
  if (blockIdx.x * kThreads + threadIdx.x >= wn) return;
  const int w = w0 + blockIdx.x * kThreads + threadIdx.x;
  const int32_t* my = nb + (size_t)w * k;
  double mx = 0.0, my_ = 0.0, mz = 0.0;
  for (int j = 0; j < k; j++) {
    const float4 p = __ldg(&pts[my[j]]);
    mx = __dadd_rn(mx, (double)p.x);
    my_ = __dadd_rn(my_, (double)p.y);
    mz = __dadd_rn(mz, (double)p.z);
  }
  mx /= (double)k; my_ /= (double)k; mz /= (double)k;
  Sym3 C;
#pragma unroll
  for (int e = 0; e < 6; e++) C.v[e] = 0.0;
  for (int j = 0; j < k; j++) {
    const float4 p = __ldg(&pts[my[j]]);
    const double dx = __dsub_rn((double)p.x, mx), dy = __dsub_rn((double)p.y, my_), dz = __dsub_rn((double)p.z, mz);
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
  knn_cov_kernel<<<blocks, kWarpKnnThreads, 0, s>>>(c.spts, c.cell_start, c.g, c.n, k, d_nb, neighbors, nullptr, nullptr, w0, wn);
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

}  // namespace apd
