// K4/K5 — the Gauss-Newton/LM normal-equation build and the cost evaluation:
// per source point e = b - T a, weighted Mahalanobis cost, H_i = J^T M J and
// b_i = J^T M e with J = [skew(T a), -I], summed over all points.
// Replaces the loops of FastAPDGICP::linearize (reference
// fast_apdgicp_impl.hpp:247-304) and compute_error (:313-343).
//
// HBM traffic per source point (fp32 Mahalanobis storage): source float4 16 B +
// correspondence 4 B + gathered target float4 16 B + Mahalanobis 24 B +
// geometric weight 4 B = 64 B; output 28 doubles per launch.
//
// Reduction: fp64 per-thread accumulators over a grid-stride loop -> fixed
// warp-shuffle tree -> fixed per-block order -> per-block partials -> the last
// block (atomic ticket) adds the partials in block order. For a given n the
// launch shape is fixed, so the result is bit-reproducible run to run.
#include "kernels.cuh"

namespace apd {

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

template <bool kFp64>
__device__ __forceinline__ void load_maha(const void* __restrict__ mahaA, const void* __restrict__ mahaB, int i, int n, double m[6]) {
  if (kFp64) {
    const double2 a = __ldg(reinterpret_cast<const double2*>(mahaA) + i);
    const double2 b = __ldg(reinterpret_cast<const double2*>(mahaB) + i);
    const double2 d = __ldg(reinterpret_cast<const double2*>(mahaB) + (size_t)n + i);
    m[0] = a.x; m[1] = a.y; m[2] = b.x; m[3] = b.y; m[4] = d.x; m[5] = d.y;
  } else {
    const float4 a = __ldg(reinterpret_cast<const float4*>(mahaA) + i);
    const float2 b = __ldg(reinterpret_cast<const float2*>(mahaB) + i);
    m[0] = (double)a.x; m[1] = (double)a.y; m[2] = (double)a.z; m[3] = (double)a.w; m[4] = (double)b.x; m[5] = (double)b.y;
  }
}

template <bool kFp64, bool kHB>
__global__ void __launch_bounds__(kThreads, 2)
linearize_kernel(const float4* __restrict__ s_spts, const float* __restrict__ s_geo, const double* __restrict__ s_geo64,
                 const int* __restrict__ corr,
                 const void* __restrict__ mahaA, const void* __restrict__ mahaB, const float4* __restrict__ t_spts, PoseD T,
                 double cl_w, int n, double* __restrict__ partials, double* __restrict__ out28, unsigned int* __restrict__ ticket) {
  constexpr int NV = kHB ? kReduceVals : 1;
  double acc[NV];
#pragma unroll
  for (int j = 0; j < NV; j++) acc[j] = 0.0;

#pragma unroll 2
  for (int i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads) {
    const int c = __ldg(&corr[i]);
    const bool valid = c >= 0;
    const int pos = valid ? (c & kCorrIndexMask) : 0;
    const float4 a = __ldg(&s_spts[i]);
    const float4 b = __ldg(&t_spts[pos]);
    const double geo = kFp64 ? __ldg(&s_geo64[i]) : (double)__ldg(&s_geo[i]);
    double m[6];
    load_maha<kFp64>(mahaA, mahaB, i, n, m);
#pragma unroll
    for (int e = 0; e < 6; e++) m[e] = valid ? m[e] : 0.0;

    const double ax = (double)a.x, ay = (double)a.y, az = (double)a.z;
    // transed_mean_A = T * mean_A ; error = mean_B - transed_mean_A (:262-263)
    const double x = ((T.r[0] * ax + T.r[1] * ay) + T.r[2] * az) + T.t[0];
    const double y = ((T.r[3] * ax + T.r[4] * ay) + T.r[5] * az) + T.t[1];
    const double z = ((T.r[6] * ax + T.r[7] * ay) + T.r[8] * az) + T.t[2];
    const double e0 = (double)b.x - x, e1 = (double)b.y - y, e2 = (double)b.z - z;
    // M e and the weighted cost (:276)
    const double me0 = (m[0] * e0 + m[1] * e1) + m[2] * e2;
    const double me1 = (m[1] * e0 + m[3] * e1) + m[4] * e2;
    const double me2 = (m[2] * e0 + m[4] * e1) + m[5] * e2;
    const double q = (e0 * me0 + e1 * me1) + e2 * me2;
    const double w = (1.0 + geo) + ((c & kCorrLabelBit) ? cl_w : 0.0);
    if (!kHB) {
      acc[0] += w * q;
    } else {
      acc[27] += w * q;
      // N = M * skew(t), t = (x,y,z): N[:,0] = M[:,1] z - M[:,2] y, N[:,1] = M[:,2] x - M[:,0] z, N[:,2] = M[:,0] y - M[:,1] x
      const double n00 = m[1] * z - m[2] * y, n01 = m[2] * x - m[0] * z, n02 = m[0] * y - m[1] * x;
      const double n10 = m[3] * z - m[4] * y, n11 = m[4] * x - m[1] * z, n12 = m[1] * y - m[3] * x;
      const double n20 = m[4] * z - m[5] * y, n21 = m[5] * x - m[2] * z, n22 = m[2] * y - m[4] * x;
      // top-left S^T M S (symmetric): TL[r][c] = sum_k S[k][r] N[k][c]
      acc[0] += z * n10 - y * n20;    // (0,0)
      acc[1] += z * n11 - y * n21;    // (0,1)
      acc[2] += z * n12 - y * n22;    // (0,2)
      acc[6] += x * n21 - z * n01;    // (1,1)
      acc[7] += x * n22 - z * n02;    // (1,2)
      acc[11] += y * n02 - x * n12;   // (2,2)
      // top-right -S^T M = -N^T: H(r, 3+c) = -N[c][r]
      acc[3] -= n00;  acc[4] -= n10;  acc[5] -= n20;    // row 0
      acc[8] -= n01;  acc[9] -= n11;  acc[10] -= n21;   // row 1
      acc[12] -= n02; acc[13] -= n12; acc[14] -= n22;   // row 2
      // bottom-right M
      acc[15] += m[0]; acc[16] += m[1]; acc[17] += m[2];
      acc[18] += m[3]; acc[19] += m[4];
      acc[20] += m[5];
      // b = J^T M e = [S^T (M e); -(M e)]
      acc[21] += z * me1 - y * me2;
      acc[22] += x * me2 - z * me0;
      acc[23] += y * me0 - x * me1;
      acc[24] -= me0; acc[25] -= me1; acc[26] -= me2;
    }
  }

  // warp tree -> block -> partials -> last block
  __shared__ double sh[kWarps][NV];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < NV; j++) {
    double v = acc[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) sh[warp][j] = v;
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    double v = 0.0;
#pragma unroll
    for (int w2 = 0; w2 < kWarps; w2++) v += sh[w2][threadIdx.x];
    partials[(size_t)blockIdx.x * NV + threadIdx.x] = v;
  }
  __shared__ bool last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (last) {
    __threadfence();
    if (threadIdx.x < NV) {
      double v = 0.0;
      for (unsigned int k = 0; k < gridDim.x; k++) v += __ldcg(&partials[(size_t)k * NV + threadIdx.x]);
      if (kHB) out28[threadIdx.x] = v;
      else out28[27] = v;
    }
    if (threadIdx.x == 0) *ticket = 0;
  }
}

}  // namespace

void launch_linearize(const CloudDev& src, const CloudDev& tgt, const PoseD& T, const CorrOut& c, double n_total, bool want_hb,
                      const ReduceWork& w, double* d_out28, cudaStream_t s, int64_t* launches) {
  const int n = src.n;
  int blocks = (n + kThreads - 1) / kThreads;
  blocks = max(1, min(blocks, w.max_blocks));
  const double cl_w = 1.0 / n_total;  // 1.0 / correspondences_.size() (:273)
#define APD_LAUNCH(FP64, HB)                                                                                                      \
  linearize_kernel<FP64, HB><<<blocks, kThreads, 0, s>>>(src.spts, src.geo, src.geo64, c.corr, c.mahaA, c.mahaB, tgt.spts, T, cl_w, n, \
                                                         w.partials, d_out28, w.ticket)
  if (c.maha_fp64) {
    if (want_hb) APD_LAUNCH(true, true);
    else APD_LAUNCH(true, false);
  } else {
    if (want_hb) APD_LAUNCH(false, true);
    else APD_LAUNCH(false, false);
  }
#undef APD_LAUNCH
  (*launches)++;
}

}  // namespace apd
