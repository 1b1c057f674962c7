// texture_kernels.cu - the fork's "texture / mean / variance" CU features and HM's AQ activity on sm_100a.
//
//   tmv_feature_kernel   replaces getTMVFeature (tools_YS.cpp:1682-1839; T3x3Filter::filter :1659-1672,
//                        TMVFeature::getSubBlockMean / getSubBlockVariance tools_YS.h:106-126): 5 directional 3x3 planes
//                        of a CU, mean and mean-absolute-deviation over whole / halves / triangles / quadrants = 5 x 26 doubles.
//   aq_activity_kernel   replaces the unit loop of TEncPreanalyzer::xPreanalyze (TEncPreanalyzer.cpp:64-139): four-quadrant
//                        variance of every AQ unit of every layer, activity = 1 + min variance.
//
// Both are integer reductions followed by a handful of IEEE double operations per region.  The doubles are part of the
// reference's results (they are printed / compared as doubles), so the kernels use only +, -, *, / in round-to-nearest
// with contraction switched off (__dmul_rn / __dsub_rn): bit-identical to the x86-64 build of the reference.
// Streaming kernels: every source sample is read once from HBM/L2 (TMV: once per CU that contains it).
#include <cuda_runtime.h>
#include "kernels.h"

namespace cucd {

namespace {
// feature slots of tools_YS.cpp:1735-1836 for the 12 regions of tmv_feature_kernel
__constant__ int kMeanSlot[12] = {0, 2, 3, 6, 7, 10, 11, 14, 15, 18, 19, 21};
__constant__ int kDevSlot[12] = {1, 4, 5, 8, 9, 12, 13, 16, 17, 22, 23, 25};
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
  return v;
}
// value of directional plane k at (i, j) of the staged n x n CU: 0 on the one-sample ring (tools_YS.cpp:1662-1663)
__device__ __forceinline__ int tmv_plane(const int16_t* t, int n, int k, int i, int j) {
  if (i == 0 || j == 0 || i == n - 1 || j == n - 1) return 0;
  const int16_t* c = t + i * n + j;
  int v;
  switch (k) {
    case 0: v = c[0]; break;
    case 1: v = c[-1] - c[1]; break;                  // aiHorMask  :1705
    case 2: v = c[-n] - c[n]; break;                  // aiVerMask  :1706
    case 3: v = c[-n + 1] - c[n - 1]; break;          // aiDiagMask :1707
    default: v = c[-n - 1] - c[n + 1]; break;         // aiAntDMask :1708
  }
  return (int)(int16_t)v;                             // the reference accumulates in Pel
}
}  // namespace

// regions: 0 whole, 1 top half, 2 bottom half, 3 left half, 4 right half, 5..8 triangles (j < n-i, j >= n-i-1, j >= i, j <= i),
//          9 top-left, 10 top-right, 11 bottom-right quadrant.  The reference's "bottom-left quadrant" entries (feature 20 / 24)
//          are computed over the whole bottom half (tools_YS.cpp:1830-1832) and are copied from region 2.
__global__ void __launch_bounds__(128)
tmv_feature_kernel(const int16_t* __restrict__ org, int stride, const TmvCu* __restrict__ cus, double* __restrict__ feat) {
  __shared__ __align__(16) int16_t tile[64 * 64];
  __shared__ int sSum[12], sDev[12];
  __shared__ double sMean[12];
  const int tid = threadIdx.x, lane = tid & 31;
  const TmvCu cu = cus[blockIdx.x];
  const int n = 1 << cu.log2n, half = n >> 1, shift = cu.log2n;
  const int16_t* src = org + (size_t)cu.y * stride + cu.x;
  for (int p = tid * 4; p < n * n; p += 128 * 4) {       // n >= 8 and x is a multiple of 8: 8-byte loads
    const int i = p >> shift, j = p & (n - 1);
    *reinterpret_cast<uint2*>(&tile[p]) = *reinterpret_cast<const uint2*>(&src[(size_t)i * stride + j]);
  }
  double* out = feat + (size_t)blockIdx.x * 130;
  for (int k = 0; k < 5; k++) {
    if (tid < 12) { sSum[tid] = 0; sDev[tid] = 0; }
    __syncthreads();
    int acc[12];
#pragma unroll
    for (int r = 0; r < 12; r++) acc[r] = 0;
    for (int p = tid; p < n * n; p += 128) {
      const int i = p >> shift, j = p & (n - 1);
      const int v = tmv_plane(tile, n, k, i, j);
      const bool top = i < half, left = j < half;
      acc[0] += v;
      acc[1] += top ? v : 0; acc[2] += top ? 0 : v;
      acc[3] += left ? v : 0; acc[4] += left ? 0 : v;
      acc[5] += (j < n - i) ? v : 0; acc[6] += (j >= n - i - 1) ? v : 0;
      acc[7] += (j >= i) ? v : 0; acc[8] += (j <= i) ? v : 0;
      acc[9] += (top && left) ? v : 0; acc[10] += (top && !left) ? v : 0; acc[11] += (!top && !left) ? v : 0;
    }
#pragma unroll
    for (int r = 0; r < 12; r++) { const int s = warp_sum(acc[r]); if (lane == 0 && s) atomicAdd(&sSum[r], s); }
    __syncthreads();
    if (tid < 12) {
      const double s = (double)sSum[tid];
      double m;
      if (tid >= 5 && tid <= 8) m = s / (double)(((unsigned)n * (unsigned)n) >> 1);                        // :1756, divisor N*N>>1
      else {
        const int rows = (tid == 0 || tid == 3 || tid == 4) ? n : half, cols = (tid <= 2) ? n : half;
        m = s / (double)rows / (double)cols;                                                               // tools_YS.h:114
      }
      sMean[tid] = m;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 12; r++) acc[r] = 0;
    const double m0 = sMean[0], m1 = sMean[1], m2 = sMean[2], m3 = sMean[3], m4 = sMean[4], m9 = sMean[9], m10 = sMean[10], m11 = sMean[11];
    for (int p = tid; p < n * n; p += 128) {
      const int i = p >> shift, j = p & (n - 1);
      const double v = (double)tmv_plane(tile, n, k, i, j);
      const bool top = i < half, left = j < half;
      const int dW = abs(__double2int_rz(__dsub_rn(v, m0)));                                                // Int tmp = sample - mean
      const int dV = abs(__double2int_rz(__dsub_rn(v, top ? m1 : m2)));
      const int dH = abs(__double2int_rz(__dsub_rn(v, left ? m3 : m4)));
      const int dQ = abs(__double2int_rz(__dsub_rn(v, top ? (left ? m9 : m10) : m11)));
      acc[0] += dW;
      acc[1] += top ? dV : 0; acc[2] += top ? 0 : dV;
      acc[3] += left ? dH : 0; acc[4] += left ? 0 : dH;
      acc[9] += (top && left) ? dQ : 0; acc[10] += (top && !left) ? dQ : 0; acc[11] += (!top && !left) ? dQ : 0;
    }
#pragma unroll
    for (int r = 0; r < 12; r++) {
      if (r >= 5 && r <= 8) continue;
      const int s = warp_sum(acc[r]); if (lane == 0 && s) atomicAdd(&sDev[r], s);
    }
    __syncthreads();
    if (tid < 12) {
      double* d = out + k * 26;
      double dev;
      if (tid >= 5 && tid <= 8) {
        // the reference ASSIGNS inside the triangle loops (:1762-1766): only the last visited sample counts, and that is
        // always a ring sample (value 0) -> |(Int)(0 - mean)| / (N*N>>1)
        dev = (double)abs(__double2int_rz(__dsub_rn(0.0, sMean[tid]))) / (double)(((unsigned)n * (unsigned)n) >> 1);
      } else {
        const int rows = (tid == 0 || tid == 3 || tid == 4) ? n : half, cols = (tid <= 2) ? n : half;
        dev = (double)sDev[tid] / (double)rows / (double)cols;
      }
      const double mean = sMean[tid];
      d[kMeanSlot[tid]] = mean; d[kDevSlot[tid]] = dev;
      if (tid == 2) { d[20] = mean; d[24] = dev; }
    }
    __syncthreads();
  }
}

cudaError_t launch_tmv_features(const int16_t* org, int stride, const TmvCu* cus, int nCu, double* feat, cudaStream_t st, int* launches) {
  if (nCu <= 0) return cudaSuccess;
  tmv_feature_kernel<<<nCu, 128, 0, st>>>(org, stride, cus, feat);
  if (launches) *launches += 1;
  return cudaGetLastError();
}

// one warp per AQ unit; layers back to back in `out` (layer d at layers.off[d]); 4 units per CTA
__global__ void __launch_bounds__(128)
aq_activity_kernel(const int16_t* __restrict__ org, int stride, int W, int H, const AqLayers layers, double* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int unit = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (unit >= layers.total) return;
  int d = 0;
  while (d + 1 < layers.count && unit >= layers.off[d + 1]) d++;
  const int part = layers.part[d], nw = (W + part - 1) / part;
  const int u = unit - layers.off[d];
  const int x0 = (u % nw) * part, y0 = (u / nw) * part;
  const int w = min(part, W - x0), h = min(part, H - y0);
  const int hw = w >> 1, hh = h >> 1;
  uint32_t s[4] = {0, 0, 0, 0};
  unsigned long long ss[4] = {0, 0, 0, 0};
  const int16_t* p0 = org + (size_t)y0 * stride + x0;
  for (int by = 0; by < h; by++) {
    const int qy = by < hh ? 0 : 2;
    for (int bx = lane; bx < w; bx += 32) {
      const int p = p0[(size_t)by * stride + bx];
      const int q = qy + (bx < hw ? 0 : 1);
      const uint32_t sq = (uint32_t)(p * p);
#pragma unroll
      for (int r = 0; r < 4; r++) { s[r] += (q == r) ? (uint32_t)p : 0u; ss[r] += (q == r) ? sq : 0u; }
    }
  }
#pragma unroll
  for (int r = 0; r < 4; r++) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) { s[r] += __shfl_xor_sync(0xffffffffu, s[r], m); ss[r] += __shfl_xor_sync(0xffffffffu, ss[r], m); }
  }
  if (lane == 0) {
    const double npix = (double)(unsigned)(w * h);
    double minVar = 1.7976931348623157e308;
#pragma unroll
    for (int r = 0; r < 4; r++) {
      const double avg = (double)s[r] / npix;                                        // TEncPreanalyzer.cpp:128
      const double var = __dsub_rn((double)ss[r] / npix, __dmul_rn(avg, avg));       // :129, no FMA contraction
      minVar = var < minVar ? var : minVar;
    }
    out[unit] = __dadd_rn(1.0, minVar);
  }
}

cudaError_t launch_aq_activity(const int16_t* org, int stride, int W, int H, const AqLayers& layers, double* out, cudaStream_t st, int* launches) {
  if (layers.total <= 0) return cudaSuccess;
  aq_activity_kernel<<<(layers.total + 3) / 4, 128, 0, st>>>(org, stride, W, H, layers, out);
  if (launches) *launches += 1;
  return cudaGetLastError();
}

}  // namespace cucd
