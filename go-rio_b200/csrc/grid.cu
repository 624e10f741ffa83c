// K0 — uniform-grid build: cell keys + per-cell counts, exclusive scan,
// STABLE radix sort of (cell, original index) and the reorder into the
// cell-sorted layout. Replaces the FLANN kd-tree build the reference does in
// setInputSource / setInputTarget (fast_apdgicp_impl.hpp:121,132).
// The sort is stable (ties keep original-index order) so the sorted layout — and
// with it every later reduction order — is bit-deterministic.
#include <cstdlib>

#include "kernels.cuh"

namespace apd {

namespace {

constexpr int kThreads = 256;

// ---------------------------------------------------------------- bounds ----
__device__ __forceinline__ unsigned int f2ord(float f) {
  unsigned int u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// state: 6 ordered-uint slots {min x,y,z = 0xffffffff; max x,y,z = 0} + a ticket, in device memory, initialised ONCE
// (init_bounds_state) — the last block to finish publishes the box and restores the initial state for the next launch.
// host_out: 6 uints + a 64-bit sequence number in pinned host memory, written by the kernel itself (zero-copy): the host
// polls the number — no memset, no copy, no stream query per cloud (a batch pool is bound by the driver's call rate).
__global__ void __launch_bounds__(kThreads) bounds_kernel(const float4* __restrict__ pts, int n, unsigned int* state, unsigned int* host_out,
                                                          unsigned long long seq) {
  float mn[3] = {3.4e38f, 3.4e38f, 3.4e38f}, mx[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 p = pts[i];
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
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int a = 0; a < 3; a++) {
      atomicMin(&state[a], f2ord(mn[a]));
      atomicMax(&state[3 + a], f2ord(mx[a]));
    }
    __threadfence();
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  if (atomicAdd(&state[6], 1u) != gridDim.x - 1) return;
  __threadfence();
#pragma unroll
  for (int a = 0; a < 6; a++) {
    host_out[a] = atomicExch(&state[a], a < 3 ? 0xffffffffu : 0u);  // read the result, restore the initial state
  }
  state[6] = 0u;
  __threadfence_system();
  *reinterpret_cast<volatile unsigned long long*>(host_out + 6) = seq;
}

__global__ void init_bounds_state_kernel(unsigned int* state) {
  if (threadIdx.x < 7) state[threadIdx.x] = threadIdx.x < 3 ? 0xffffffffu : 0u;
}

// ------------------------------------------------------------ keys/counts ----
__global__ void __launch_bounds__(kThreads) cell_keys_kernel(const float4* __restrict__ pts, int n, GridDesc g,
                                                             uint32_t* __restrict__ keys, uint32_t* __restrict__ vals,
                                                             uint32_t* __restrict__ counts) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = pts[i];
  const int cx = cell_coord(p.x, g.ox, g.inv_cell, g.nx);
  const int cy = cell_coord(p.y, g.oy, g.inv_cell, g.ny);
  const int cz = cell_coord(p.z, g.oz, g.inv_cell, g.nz);
  const uint32_t key = (uint32_t)((cz * g.ny + cy) * g.nx + cx);
  keys[i] = key;
  vals[i] = (uint32_t)i;
  atomicAdd(&counts[key], 1u);
}

// ------------------------------------------------------------------ scan ----
// exclusive scan of 2048-element tiles; tile totals go to block_sums
__global__ void __launch_bounds__(kThreads) scan_tiles_kernel(const uint32_t* in, uint32_t* out,  // may alias (in-place)
                                                              size_t n, uint32_t* __restrict__ block_sums) {
  __shared__ uint32_t warp_tot[kThreads / 32];
  const size_t base = (size_t)blockIdx.x * kScanTile + (size_t)threadIdx.x * 8;
  uint32_t v[8];
  uint32_t sum = 0;
#pragma unroll
  for (int j = 0; j < 8; j++) {
    v[j] = (base + j < n) ? in[base + j] : 0u;
    sum += v[j];
  }
  // inclusive warp scan of the per-thread sums
  uint32_t inc = sum;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_tot[warp] = inc;
  __syncthreads();
  uint32_t warp_off = 0, total = 0;
#pragma unroll
  for (int w = 0; w < kThreads / 32; w++) {
    const uint32_t t = warp_tot[w];
    if (w < warp) warp_off += t;
    total += t;
  }
  uint32_t run = warp_off + inc - sum;
#pragma unroll
  for (int j = 0; j < 8; j++) {
    if (base + j < n) out[base + j] = run;
    run += v[j];
  }
  if (threadIdx.x == 0 && block_sums) block_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kThreads) scan_add_kernel(uint32_t* __restrict__ out, size_t n, const uint32_t* __restrict__ block_sums) {
  const size_t base = (size_t)blockIdx.x * kScanTile + (size_t)threadIdx.x * 8;
  const uint32_t off = block_sums[blockIdx.x];
  if (off == 0) return;
#pragma unroll
  for (int j = 0; j < 8; j++)
    if (base + j < n) out[base + j] += off;
}

void exclusive_scan(uint32_t* data, size_t n, uint32_t* tmp, cudaStream_t s, int64_t* launches) {
  if (n == 0) return;
  const size_t blocks = (n + kScanTile - 1) / kScanTile;
  if (blocks == 1) {
    scan_tiles_kernel<<<1, kThreads, 0, s>>>(data, data, n, nullptr);
    (*launches)++;
    return;
  }
  scan_tiles_kernel<<<(unsigned)blocks, kThreads, 0, s>>>(data, data, n, tmp);
  (*launches)++;
  exclusive_scan(tmp, blocks, tmp + blocks, s, launches);
  scan_add_kernel<<<(unsigned)blocks, kThreads, 0, s>>>(data, n, tmp);
  (*launches)++;
}

// ------------------------------------------------------------ radix sort ----
__global__ void __launch_bounds__(kThreads) radix_hist_kernel(const uint32_t* __restrict__ keys, int n, int shift,
                                                              uint32_t* __restrict__ hist, int nblocks) {
  __shared__ uint32_t h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const int base = blockIdx.x * kSortTile;
#pragma unroll
  for (int j = 0; j < kSortTile / kThreads; j++) {
    const int i = base + j * kThreads + threadIdx.x;
    if (i < n) atomicAdd(&h[(keys[i] >> shift) & 255u], 1u);
  }
  __syncthreads();
  hist[(size_t)threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];
}

__global__ void __launch_bounds__(kThreads) radix_scatter_kernel(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                                                                 uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, int n,
                                                                 int shift, const uint32_t* __restrict__ hist_scanned, int nblocks) {
  constexpr int kWarps = kThreads / 32;
  constexpr int kIters = kSortTile / kThreads;  // 8 iterations of 32 keys per warp
  __shared__ uint32_t wcnt[kWarps][256];
  for (int j = threadIdx.x; j < kWarps * 256; j += kThreads) (&wcnt[0][0])[j] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wbase = blockIdx.x * kSortTile + warp * (32 * kIters);
  uint32_t key[kIters], val[kIters], rank[kIters];
  const unsigned lt = (1u << lane) - 1u;
#pragma unroll
  for (int it = 0; it < kIters; it++) {
    const int i = wbase + it * 32 + lane;
    const bool ok = i < n;
    key[it] = ok ? keys_in[i] : 0xffffffffu;
    val[it] = ok ? vals_in[i] : 0u;
    const uint32_t d = ok ? ((key[it] >> shift) & 255u) : 256u;
    const unsigned peers = __match_any_sync(0xffffffffu, d);
    uint32_t before = 0;
    if (ok) before = wcnt[warp][d];
    __syncwarp();
    rank[it] = before + __popc(peers & lt);
    if (ok && (peers & lt) == 0) wcnt[warp][d] = before + __popc(peers);  // lowest lane of the peer group
    __syncwarp();
  }
  __syncthreads();
  {
    const int d = threadIdx.x;  // 256 threads <-> 256 digits
    uint32_t run = hist_scanned[(size_t)d * nblocks + blockIdx.x];
#pragma unroll
    for (int w = 0; w < kWarps; w++) {
      const uint32_t t = wcnt[w][d];
      wcnt[w][d] = run;
      run += t;
    }
  }
  __syncthreads();
#pragma unroll
  for (int it = 0; it < kIters; it++) {
    const int i = wbase + it * 32 + lane;
    if (i < n) {
      const uint32_t d = (key[it] >> shift) & 255u;
      const uint32_t pos = wcnt[warp][d] + rank[it];
      keys_out[pos] = key[it];
      vals_out[pos] = val[it];
    }
  }
}

// ---------------------------------------------------------------- reorder ----
__global__ void __launch_bounds__(kThreads) reorder_kernel(const float4* __restrict__ pts, const uint32_t* __restrict__ vals, int n,
                                                           float4* __restrict__ spts, float* __restrict__ label, int* __restrict__ inv_perm,
                                                           unsigned char* __restrict__ zero_flags) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  if (zero_flags) zero_flags[s] = 0;
  const int idx = (int)vals[s];
  const float4 p = pts[idx];
  spts[s] = make_float4(p.x, p.y, p.z, __int_as_float(idx));
  label[s] = p.w;
  inv_perm[idx] = s;
}

// ------------------------------------------------- small-cloud path ----
// For clouds of at most kSmallCloud points the stable sort is done without radix
// passes (5 launches instead of ~20): atomic per-cell ranks place the points of a
// cell in arbitrary order, then every point counts the points of its cell with a
// smaller original index to find its deterministic slot.
__global__ void __launch_bounds__(kThreads) cell_keys_rank_kernel(const float4* __restrict__ pts, int n, GridDesc g,
                                                                  uint32_t* __restrict__ keys, uint32_t* __restrict__ nd_rank,
                                                                  uint32_t* __restrict__ counts) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = pts[i];
  const int cx = cell_coord(p.x, g.ox, g.inv_cell, g.nx);
  const int cy = cell_coord(p.y, g.oy, g.inv_cell, g.ny);
  const int cz = cell_coord(p.z, g.oz, g.inv_cell, g.nz);
  const uint32_t key = (uint32_t)((cz * g.ny + cy) * g.nx + cx);
  keys[i] = key;
  nd_rank[i] = atomicAdd(&counts[key], 1u);
}

// tile-wise exclusive scan (in place); the last block to finish scans the tile totals
__global__ void __launch_bounds__(kThreads) scan_tiles_last_kernel(uint32_t* data, size_t n, uint32_t* __restrict__ tile_sums,
                                                                   unsigned int* __restrict__ ticket) {
  __shared__ uint32_t warp_tot[kThreads / 32];
  __shared__ bool last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  {
    const size_t base = (size_t)blockIdx.x * kScanTile + (size_t)threadIdx.x * 8;
    uint32_t v[8], sum = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      v[j] = (base + j < n) ? data[base + j] : 0u;
      sum += v[j];
    }
    uint32_t inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    uint32_t warp_off = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; w++) {
      const uint32_t t = warp_tot[w];
      if (w < warp) warp_off += t;
      total += t;
    }
    uint32_t run = warp_off + inc - sum;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      if (base + j < n) data[base + j] = run;
      run += v[j];
    }
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!last) return;
  __threadfence();
  // exclusive scan of up to 2048 tile totals by this block
  const int nt = (int)gridDim.x;
  uint32_t v[8], sum = 0;
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const int t = threadIdx.x * 8 + j;
    v[j] = t < nt ? __ldcg(&tile_sums[t]) : 0u;
    sum += v[j];
  }
  uint32_t inc = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();
  if (lane == 31) warp_tot[warp] = inc;
  __syncthreads();
  uint32_t warp_off = 0;
#pragma unroll
  for (int w = 0; w < kThreads / 32; w++)
    if (w < warp) warp_off += warp_tot[w];
  uint32_t run = warp_off + inc - sum;
#pragma unroll
  for (int j = 0; j < 8; j++) {
    const int t = threadIdx.x * 8 + j;
    if (t < nt) tile_sums[t] = run;
    run += v[j];
  }
  if (threadIdx.x == 0) *ticket = 0;
}

__global__ void __launch_bounds__(kThreads) scatter_nd_kernel(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ nd_rank, int n,
                                                              const uint32_t* __restrict__ cell_start, uint32_t* __restrict__ tmp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  tmp[cell_start[keys[i]] + nd_rank[i]] = (uint32_t)i;
}

__global__ void __launch_bounds__(kThreads) rank_fix_reorder_kernel(const float4* __restrict__ pts, const uint32_t* __restrict__ keys,
                                                                    const uint32_t* __restrict__ tmp, int n,
                                                                    const uint32_t* __restrict__ cell_start, float4* __restrict__ spts,
                                                                    float* __restrict__ label, int* __restrict__ inv_perm,
                                                                    unsigned char* __restrict__ zero_flags) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (zero_flags) zero_flags[i] = 0;
  const uint32_t key = keys[i];
  const uint32_t b = cell_start[key], e = cell_start[key + 1];
  uint32_t r = 0;
  for (uint32_t j = b; j < e; j++) r += (tmp[j] < (uint32_t)i) ? 1u : 0u;
  const int s = (int)(b + r);
  const float4 p = pts[i];
  spts[s] = make_float4(p.x, p.y, p.z, __int_as_float(i));
  label[s] = p.w;
  inv_perm[i] = s;
}

// ------------------------------------------------- tiny-cloud path ----
// A radar scan (1-4 k points, <= 32 k cells) is built by ONE CTA in ONE launch: zero the counters, atomic per-cell
// ranks, block-wide exclusive scan, scatter, deterministic slot by counting (the small-cloud path's steps with CTA
// barriers instead of kernel boundaries — same sorted layout). A batch pool is bound by the driver's launch rate, and
// a lone scan-to-scan registration by launch latency: 1 call instead of 6.
constexpr int kTinyThreads = 1024;
constexpr int kTinyCloud = 4096;
constexpr int kTinyCellsPerThread = 32;
// (no __restrict__ / const on the arrays written here: later phases must not read them through the non-coherent path)
__global__ void __launch_bounds__(kTinyThreads) grid_tiny_kernel(const float4* __restrict__ pts, int n, GridDesc g, int ncs, uint32_t* cell_start,
                                                                 uint32_t* keys, uint32_t* nd_rank, uint32_t* tmp, float4* __restrict__ spts,
                                                                 float* __restrict__ label, int* __restrict__ inv_perm,
                                                                 unsigned char* __restrict__ zero_flags) {
  __shared__ uint32_t warp_tot[kTinyThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int j = tid; j < ncs; j += kTinyThreads) cell_start[j] = 0u;
  __syncthreads();
  for (int i = tid; i < n; i += kTinyThreads) {
    const float4 p = pts[i];
    const int cx = cell_coord(p.x, g.ox, g.inv_cell, g.nx);
    const int cy = cell_coord(p.y, g.oy, g.inv_cell, g.ny);
    const int cz = cell_coord(p.z, g.oz, g.inv_cell, g.nz);
    const uint32_t key = (uint32_t)((cz * g.ny + cy) * g.nx + cx);
    keys[i] = key;
    nd_rank[i] = atomicAdd(&cell_start[key], 1u);
  }
  __syncthreads();
  {  // exclusive scan of cell_start[0 .. ncs): thread t owns `per` consecutive counters
    const int per = (ncs + kTinyThreads - 1) / kTinyThreads;
    const int b = min(tid * per, ncs), e = min(b + per, ncs);
    uint32_t sum = 0;
    for (int j = b; j < e; j++) sum += cell_start[j];
    uint32_t inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    uint32_t run = inc - sum;
    for (int w = 0; w < warp; w++) run += warp_tot[w];
    for (int j = b; j < e; j++) {
      const uint32_t v = cell_start[j];
      cell_start[j] = run;
      run += v;
    }
  }
  __syncthreads();
  for (int i = tid; i < n; i += kTinyThreads) tmp[cell_start[keys[i]] + nd_rank[i]] = (uint32_t)i;
  __syncthreads();
  for (int i = tid; i < n; i += kTinyThreads) {
    const uint32_t key = keys[i];
    const uint32_t b = cell_start[key], e = cell_start[key + 1];
    uint32_t r = 0;
    for (uint32_t j = b; j < e; j++) r += (tmp[j] < (uint32_t)i) ? 1u : 0u;
    const int sp = (int)(b + r);
    const float4 p = pts[i];
    spts[sp] = make_float4(p.x, p.y, p.z, __int_as_float(i));
    label[sp] = p.w;
    inv_perm[i] = sp;
    if (zero_flags) zero_flags[i] = 0;
  }
}

}  // namespace

size_t scan_tmp_elems_for(size_t n) {
  size_t total = 0;
  while (n > (size_t)kScanTile) {
    n = (n + kScanTile - 1) / kScanTile;
    total += n;
  }
  return total + 8;
}

void init_bounds_state(unsigned int* d_state, cudaStream_t s) { init_bounds_state_kernel<<<1, 32, 0, s>>>(d_state); }
void launch_bounds(const float4* pts, int n, unsigned int* d_state, unsigned int* h_out, unsigned long long seq, cudaStream_t s,
                   int64_t* launches) {
  const int blocks = min(148 * 8, (n + kThreads - 1) / kThreads);
  bounds_kernel<<<max(1, blocks), kThreads, 0, s>>>(pts, n, d_state, h_out, seq);
  (*launches)++;
}

void launch_exclusive_scan_u32(uint32_t* data, size_t n, uint32_t* tmp, cudaStream_t s, int64_t* launches) { exclusive_scan(data, n, tmp, s, launches); }

int launch_sort_pairs_u32(uint32_t* const keys[2], uint32_t* const vals[2], int n, int bits, uint32_t* hist, uint32_t* scan_tmp, cudaStream_t s,
                          int64_t* launches) {
  const int passes = (bits + 7) / 8;
  const int sblocks = (n + kSortTile - 1) / kSortTile;
  int cur = 0;
  for (int p = 0; p < passes; p++) {
    radix_hist_kernel<<<sblocks, kThreads, 0, s>>>(keys[cur], n, p * 8, hist, sblocks);
    (*launches)++;
    exclusive_scan(hist, (size_t)256 * sblocks, scan_tmp, s, launches);
    radix_scatter_kernel<<<sblocks, kThreads, 0, s>>>(keys[cur], vals[cur], keys[cur ^ 1], vals[cur ^ 1], n, p * 8, hist, sblocks);
    (*launches)++;
    cur ^= 1;
  }
  return cur;
}

void launch_grid_build(const CloudDev& c, const GridWork& w, cudaStream_t s, int64_t* launches) {
  const int n = c.n;
  if (n <= 0) return;
  const int pblocks = (n + kThreads - 1) / kThreads;
  const size_t ncs = (size_t)c.ncells + 1;
  static const bool tiny_ok = [] { const char* e = getenv("APD_GRID_TINY"); return !e || atoi(e) != 0; }();
  if (tiny_ok && n <= kTinyCloud && ncs <= (size_t)kTinyThreads * kTinyCellsPerThread) {
    grid_tiny_kernel<<<1, kTinyThreads, 0, s>>>(c.pts, n, c.g, (int)ncs, c.cell_start, w.keys[0], w.vals[0], w.vals[1], c.spts, c.label,
                                                c.inv_perm, w.zero_flags);
    (*launches)++;
    return;
  }
  cudaMemsetAsync(c.cell_start, 0, sizeof(uint32_t) * ((size_t)c.ncells + 1), s);
  const size_t tiles = (ncs + kScanTile - 1) / kScanTile;
  if (n <= kSmallCloud && tiles <= (size_t)kScanTile && w.ticket != nullptr) {
    cell_keys_rank_kernel<<<pblocks, kThreads, 0, s>>>(c.pts, n, c.g, w.keys[0], w.vals[0], c.cell_start);
    scan_tiles_last_kernel<<<(unsigned)tiles, kThreads, 0, s>>>(c.cell_start, ncs, w.scan_tmp, w.ticket);
    if (tiles > 1) {
      scan_add_kernel<<<(unsigned)tiles, kThreads, 0, s>>>(c.cell_start, ncs, w.scan_tmp);
      (*launches)++;
    }
    scatter_nd_kernel<<<pblocks, kThreads, 0, s>>>(w.keys[0], w.vals[0], n, c.cell_start, w.vals[1]);
    rank_fix_reorder_kernel<<<pblocks, kThreads, 0, s>>>(c.pts, w.keys[0], w.vals[1], n, c.cell_start, c.spts, c.label, c.inv_perm, w.zero_flags);
    (*launches) += 4;
    return;
  }
  cell_keys_kernel<<<pblocks, kThreads, 0, s>>>(c.pts, n, c.g, w.keys[0], w.vals[0], c.cell_start);
  (*launches)++;
  exclusive_scan(c.cell_start, (size_t)c.ncells + 1, w.scan_tmp, s, launches);

  int bits = 1;
  while ((1ll << bits) < (long long)c.ncells) bits++;
  const int passes = (bits + 7) / 8;
  const int sblocks = (n + kSortTile - 1) / kSortTile;
  int cur = 0;
  for (int p = 0; p < passes; p++) {
    radix_hist_kernel<<<sblocks, kThreads, 0, s>>>(w.keys[cur], n, p * 8, w.hist, sblocks);
    (*launches)++;
    exclusive_scan(w.hist, (size_t)256 * sblocks, w.scan_tmp, s, launches);
    radix_scatter_kernel<<<sblocks, kThreads, 0, s>>>(w.keys[cur], w.vals[cur], w.keys[cur ^ 1], w.vals[cur ^ 1], n, p * 8, w.hist, sblocks);
    (*launches)++;
    cur ^= 1;
  }
  reorder_kernel<<<pblocks, kThreads, 0, s>>>(c.pts, w.vals[cur], n, c.spts, c.label, c.inv_perm, w.zero_flags);
  (*launches)++;
}

// Loads this file's kernels into the current context (CUDA loads kernels lazily, at their first launch, and a load may have
// to synchronise with the context: if it happens while another rank's kernel of the same process is spinning on a peer
// — the sharded exchange — neither can proceed. apd_group_create loads everything up front.)
void preload_grid_kernels() {
  cudaFuncAttributes a;
  (void)cudaFuncGetAttributes(&a, bounds_kernel);
  (void)cudaFuncGetAttributes(&a, init_bounds_state_kernel);
  (void)cudaFuncGetAttributes(&a, cell_keys_kernel);
  (void)cudaFuncGetAttributes(&a, scan_tiles_kernel);
  (void)cudaFuncGetAttributes(&a, scan_add_kernel);
  (void)cudaFuncGetAttributes(&a, radix_hist_kernel);
  (void)cudaFuncGetAttributes(&a, radix_scatter_kernel);
  (void)cudaFuncGetAttributes(&a, reorder_kernel);
  (void)cudaFuncGetAttributes(&a, cell_keys_rank_kernel);
  (void)cudaFuncGetAttributes(&a, scan_tiles_last_kernel);
  (void)cudaFuncGetAttributes(&a, scatter_nd_kernel);
  (void)cudaFuncGetAttributes(&a, rank_fix_reorder_kernel);
  (void)cudaFuncGetAttributes(&a, grid_tiny_kernel);
}

}  // namespace apd
