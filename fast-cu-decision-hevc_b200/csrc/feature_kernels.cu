// feature_kernels.cu - the fork's per-picture texture features on sm_100a.
//
// Replaces TEncSlice::getOutlierWithDCT (TEncSlice.cpp:878-1173), the per-CU OBF block sums of
// TEncCu.cpp:589-600 and the source-only Hadamard cost of TEncCu.cpp:1780-1893.  The truncated
// Laplacian fit between the two passes (TCMprocessOneSequence, TEncSlice.cpp:291-392, double
// precision exp/log) stays on the host in tcm_host.cpp, fed by the histograms of pass 1.
//
// Thread mapping for both DCT passes: one CTA = one 64x64 CTU, one thread = one 4x4 block
// (16 lanes read 128 contiguous bytes of a sample row).  These kernels are HBM/L2 streaming
// kernels: 8 B of source per thread row, 32 B of Outlier output per thread.
#include <cuda_runtime.h>
#include <algorithm>
#include "feature_core.cuh"
#include "kernels.h"

namespace cucd {

constexpr int kLowBins = 512;   // bins kept in shared memory; rarer, larger magnitudes go to global atomics

__global__ void __launch_bounds__(256)
feature_hist_kernel(const FeaturePlanes fp, uint32_t* __restrict__ hist) {
  __shared__ uint32_t sh[15 * kLowBins];
  const int tid = threadIdx.x, pic = blockIdx.y;
  for (int i = tid; i < 15 * kLowBins; i += 256) sh[i] = 0;
  __syncthreads();
  const int16_t* org = fp.org + (size_t)pic * fp.orgPicStride;
  uint32_t* gh = hist + (size_t)pic * kHistFreqs * kHistBins;
  const int bw = fp.W >> 2, bh = fp.H >> 2;           // whole 4x4 blocks only (TEncSlice.cpp:930-931)
  for (int ctu = blockIdx.x; ctu < fp.ctusPerPic; ctu += gridDim.x) {
    const int bx = (ctu % fp.ctusPerRow) * 16 + (tid & 15), by = (ctu / fp.ctusPerRow) * 16 + (tid >> 4);
    if (bx < bw && by < bh) {
      int c[16];
      dct4x4(org + (size_t)(by * 4) * fp.orgStride + bx * 4, fp.orgStride, fp.bitDepth, c);
#pragma unroll
      for (int f = 1; f < 16; f++) {
        const int bin = min(coeff_bin(c[f]), kHistBins - 1);
        if (bin < kLowBins) atomicAdd(&sh[(f - 1) * kLowBins + bin], 1u);
        else atomicAdd(&gh[f * kHistBins + bin], 1u);
      }
    }
  }
  __syncthreads();
  for (int i = tid; i < 15 * kLowBins; i += 256) {
    const uint32_t v = sh[i];
    if (v) atomicAdd(&gh[(i / kLowBins + 1) * kHistBins + (i % kLowBins)], v);
  }
}

__global__ void __launch_bounds__(256)
feature_obf_kernel(const FeaturePlanes fp, const int32_t* __restrict__ thr, const FeatureOut out) {
  __shared__ int16_t cell[16][16];      // OBF count of every 4x4 block of the CTU (0 outside the picture)
  __shared__ int32_t sNum[64 + 16 + 4], sSum[64 + 16 + 4];
  const int tid = threadIdx.x, pic = blockIdx.y, ctu = blockIdx.x;
  const int cx = tid & 15, cy = tid >> 4;
  const int ctuX = (ctu % fp.ctusPerRow) * 64, ctuY = (ctu / fp.ctusPerRow) * 64;
  const int bw = fp.W >> 2, bh = fp.H >> 2;
  const int bx = ctuX / 4 + cx, by = ctuY / 4 + cy;
  const int32_t* t = thr + pic * kHistFreqs;
  int cnt = 0;
  if (bx < bw && by < bh) {
    const int16_t* org = fp.org + (size_t)pic * fp.orgPicStride;
    int c[16];
    dct4x4(org + (size_t)(by * 4) * fp.orgStride + bx * 4, fp.orgStride, fp.bitDepth, c);
    uint32_t o[8];
#pragma unroll
    for (int f = 0; f < 16; f++) {
      int v = 0;
      if (f > 0 && coeff_is_outlier(c[f], t[f])) { v = c[f]; cnt++; }     // DC dropped, TEncSlice.cpp:987-988
      int16_t p = (int16_t)v;                                              // Pel = short, TEncSlice.cpp:1106
      if (p < 0) p = (int16_t)-p;
      const uint32_t q = (uint32_t)(uint16_t)(int16_t)(p / 100);
      if (f & 1) o[f >> 1] |= q << 16; else o[f >> 1] = q;
    }
    if (out.outlier) {
      int16_t* dst = out.outlier + (size_t)pic * out.outlierPicStride + (size_t)(by * 4) * fp.W + bx * 4;
#pragma unroll
      for (int y = 0; y < 4; y++) *reinterpret_cast<uint2*>(dst + (size_t)y * fp.W) = make_uint2(o[2 * y], o[2 * y + 1]);
    }
    if (out.outlier8) {                   // |AC coefficient| / 100 <= 163: exact as a byte (include/cucudecide.h)
      uint8_t* dst = out.outlier8 + (size_t)pic * out.outlierPicStride + (size_t)(by * 4) * fp.W + bx * 4;
#pragma unroll
      for (int y = 0; y < 4; y++) *reinterpret_cast<uint32_t*>(dst + (size_t)y * fp.W) = __byte_perm(o[2 * y], o[2 * y + 1], 0x6420);
    }
    if (out.obf) out.obf[(size_t)pic * out.obfPicStride + (size_t)by * bw + bx] = (int16_t)cnt;
    if (out.obf8) out.obf8[(size_t)pic * out.obfPicStride + (size_t)by * bw + bx] = (uint8_t)cnt;
  }
  cell[cy][cx] = (int16_t)cnt;
  __syncthreads();
  // depth 3 (8x8 CUs = 2x2 cells), then 2, 1, 0 hierarchically: Num_OBF = #cells > 0, N_Outlier = sum of cells
  if (tid < 64) {
    const int ux = tid & 7, uy = tid >> 3;
    int num = 0, sum = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) { const int v = cell[uy * 2 + (k >> 1)][ux * 2 + (k & 1)]; num += v > 0; sum += v; }
    sNum[tid] = num; sSum[tid] = sum;
  }
  __syncthreads();
  if (tid < 16) {
    const int ux = tid & 3, uy = tid >> 2;
    int num = 0, sum = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) { const int i = (uy * 2 + (k >> 1)) * 8 + ux * 2 + (k & 1); num += sNum[i]; sum += sSum[i]; }
    sNum[64 + tid] = num; sSum[64 + tid] = sum;
  }
  __syncthreads();
  if (tid < 4) {
    const int ux = tid & 1, uy = tid >> 1;
    int num = 0, sum = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) { const int i = 64 + (uy * 2 + (k >> 1)) * 4 + ux * 2 + (k & 1); num += sNum[i]; sum += sSum[i]; }
    sNum[80 + tid] = num; sSum[80 + tid] = sum;
  }
  __syncthreads();
  // scatter: a CU is reported only if it lies completely inside the picture
  if (tid < 85) {
    int depth, ux, uy, num, sum;
    if (tid < 64) { depth = 3; ux = tid & 7; uy = tid >> 3; num = sNum[tid]; sum = sSum[tid]; }
    else if (tid < 80) { depth = 2; ux = (tid - 64) & 3; uy = (tid - 64) >> 2; num = sNum[tid]; sum = sSum[tid]; }
    else if (tid < 84) { depth = 1; ux = (tid - 80) & 1; uy = (tid - 80) >> 1; num = sNum[tid]; sum = sSum[tid]; }
    else { depth = 0; ux = 0; uy = 0; num = sNum[80] + sNum[81] + sNum[82] + sNum[83]; sum = sSum[80] + sSum[81] + sSum[82] + sSum[83]; }
    const int size = 64 >> depth;
    const int gx = ctuX / size + ux, gy = ctuY / size + uy;
    const int cw = fp.W / size, ch = fp.H / size;
    if (gx < cw && gy < ch) {
      const size_t o = (size_t)pic * out.cuPicStride[depth] + (size_t)gy * cw + gx;
      out.numObf[depth][o] = num; out.nOutlier[depth][o] = sum;
    }
  }
}

// One warp per CTU; a lane takes 8x8 blocks lane, lane+32.  Cost of a block = (sum|H s| - |DC| + 2) >> 2.
__global__ void __launch_bounds__(128)
ctu_src_had_kernel(const FeaturePlanes fp, int32_t* __restrict__ ctuHad) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, pic = blockIdx.y;
  const int ctu = blockIdx.x * 4 + warp;
  if (ctu >= fp.ctusPerPic) return;
  const int ctuX = (ctu % fp.ctusPerRow) * 64, ctuY = (ctu / fp.ctusPerRow) * 64;
  const int16_t* org = fp.org + (size_t)pic * fp.orgPicStride;
  int sum = 0;
#pragma unroll
  for (int k = 0; k < 2; k++) {
    const int b = lane + 32 * k, x = ctuX + (b & 7) * 8, y = ctuY + (b >> 3) * 8;
    if (x + 8 <= fp.W && y + 8 <= fp.H) {
      Tile t;
      tile_load(t, org + (size_t)y * fp.orgStride + x, fp.orgStride);
      // the packed transform of non-negative samples; word 0 ends up holding the two halves of the DC term
      uint32_t* d = t.r;
#pragma unroll
      for (int j = 0; j < 4; j++)
#pragma unroll
        for (int len = 1; len < 8; len <<= 1)
#pragma unroll
          for (int yy = 0; yy < 8; yy++)
            if (!(yy & len)) { const uint32_t a = d[yy * 4 + j], c = d[(yy + len) * 4 + j]; d[yy * 4 + j] = a + c; d[(yy + len) * 4 + j] = a - c; }
#pragma unroll
      for (int yy = 0; yy < 8; yy++)
#pragma unroll
        for (int len = 1; len < 4; len <<= 1)
#pragma unroll
          for (int j = 0; j < 4; j++)
            if (!(j & len)) { const uint32_t a = d[yy * 4 + j], c = d[yy * 4 + j + len]; d[yy * 4 + j] = a + c; d[yy * 4 + j + len] = a - c; }
      int acc = 0;
#pragma unroll
      for (int i = 0; i < 32; i++) acc += pair_max_abs(d[i]);
      const int lo = (int)(int16_t)(d[0] & 0xffffu), hi = ((int)d[0] - lo) >> 16;
      sum += (2 * acc - iabs32(lo + hi) + 2) >> 2;
    }
  }
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, m);
  if (lane == 0) ctuHad[(size_t)pic * fp.ctusPerPic + ctu] = sum;
}

__global__ void copy_words_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, size_t n4) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}
cudaError_t launch_copy_words(const uint32_t* src, uint32_t* dst, size_t nWords, cudaStream_t st, int* launches) {
  if (nWords == 0) return cudaSuccess;
  copy_words_kernel<<<64, 256, 0, st>>>(reinterpret_cast<const uint4*>(src), reinterpret_cast<uint4*>(dst), nWords / 4);   // callers pass multiples of 4 words
  if (launches) *launches += 1;
  return cudaGetLastError();
}

// 16 samples per thread: u8 -> int16 (upload of 8-bit content held as bytes)
__global__ void widen_u8_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, size_t n16) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
    const uint4 v = src[i];
    dst[2 * i] = make_uint4(__byte_perm(v.x, 0, 0x4140), __byte_perm(v.x, 0, 0x4342), __byte_perm(v.y, 0, 0x4140), __byte_perm(v.y, 0, 0x4342));
    dst[2 * i + 1] = make_uint4(__byte_perm(v.z, 0, 0x4140), __byte_perm(v.z, 0, 0x4342), __byte_perm(v.w, 0, 0x4140), __byte_perm(v.w, 0, 0x4342));
  }
}
static int convert_blocks(size_t n16) { return (int)std::min<size_t>((n16 + 255) / 256, 148 * 8); }
cudaError_t launch_widen_u8(const uint8_t* src, int16_t* dst, size_t nSamples, cudaStream_t st, int* launches) {
  if (nSamples == 0) return cudaSuccess;
  widen_u8_kernel<<<convert_blocks(nSamples / 16), 256, 0, st>>>(reinterpret_cast<const uint4*>(src), reinterpret_cast<uint4*>(dst), nSamples / 16);
  if (launches) *launches += 1;
  return cudaGetLastError();
}
// one thread per CTU walks its 85 CUs (prune_mask_ctu)
__global__ void prune_mask_kernel(const FeaturePlanes fp, const FeatureOut sums, const PruneSwitches sw, uint8_t* __restrict__ needed, int nPics) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nPics * fp.ctusPerPic) return;
  const int pic = i / fp.ctusPerPic, ctu = i - pic * fp.ctusPerPic;
  auto numObf = [&](int d, int cx, int cy) { return sums.numObf[d][(size_t)pic * sums.cuPicStride[d] + (size_t)cy * (fp.W / (64 >> d)) + cx]; };
  uint8_t m[341];
  prune_mask_ctu((ctu % fp.ctusPerRow) * 64, (ctu / fp.ctusPerRow) * 64, fp.W, fp.H, sw.skip2Nx2N, sw.terminateCU, numObf, m);
  uint8_t* dst = needed + (size_t)i * 341;
  for (int k = 0; k < 341; k++) dst[k] = m[k];
}
cudaError_t launch_prune_mask(const FeaturePlanes& fp, int nPics, const FeatureOut& sums, PruneSwitches sw, uint8_t* needed, cudaStream_t st, int* launches) {
  const int n = nPics * fp.ctusPerPic;
  prune_mask_kernel<<<(n + 63) / 64, 64, 0, st>>>(fp, sums, sw, needed, nPics);
  if (launches) *launches += 1;
  return cudaGetLastError();
}

cudaError_t launch_feature_hist(const FeaturePlanes& fp, int nPics, uint32_t* hist, cudaStream_t st, int* launches) {
  cudaError_t e = cudaMemsetAsync(hist, 0, (size_t)nPics * kHistFreqs * kHistBins * sizeof(uint32_t), st);
  if (e != cudaSuccess) return e;
  int gx = (148 * 4 + nPics - 1) / nPics;           // ~4 CTAs per SM over the whole launch
  if (gx > fp.ctusPerPic) gx = fp.ctusPerPic;
  if (gx < 1) gx = 1;
  feature_hist_kernel<<<dim3(gx, nPics), 256, 0, st>>>(fp, hist);
  if (launches) *launches += 1;
  return cudaGetLastError();
}
cudaError_t launch_feature_obf(const FeaturePlanes& fp, int nPics, const int32_t* thr, const FeatureOut& out, cudaStream_t st, int* launches) {
  feature_obf_kernel<<<dim3(fp.ctusPerPic, nPics), 256, 0, st>>>(fp, thr, out);
  if (launches) *launches += 1;
  return cudaGetLastError();
}
cudaError_t launch_ctu_src_had(const FeaturePlanes& fp, int nPics, int32_t* ctuHad, cudaStream_t st, int* launches) {
  ctu_src_had_kernel<<<dim3((fp.ctusPerPic + 3) / 4, nPics), 128, 0, st>>>(fp, ctuHad);
  if (launches) *launches += 1;
  return cudaGetLastError();
}

}  // namespace cucd
