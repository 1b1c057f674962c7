// rmd_kernels.cu - intra rough-mode-decision cost tables on sm_100a.
//
// One CTA of 256 threads per chunk (see rmd_chunk.cuh): borders are built cooperatively in shared
// memory, every lane keeps its 8x8 source tile in registers, the 35 modes are spread over the 8 warps
// (mode is warp-uniform), predictions and Hadamard butterflies stay in registers, per-PU costs are
// accumulated in shared memory and leave the SM as one coalesced block of 35*PUS uint32.
// Replaces, per PU, the reference loop TEncSearch.cpp:2327-2361 (initAdiPatternChType ->
// predIntraAng x35 -> xGetHADs).
#include <cuda_runtime.h>
#include <algorithm>
#include "rmd_chunk.cuh"
#include "satd_tc.cuh"
#include "kernels.h"

namespace cucd {

extern __shared__ __align__(16) unsigned char smem_raw[];

// The mode loop of a chunk.  Not templated on the PU size: N = 8..64 share this code, N = 4 takes the
// region branch.  __noinline__ keeps exactly one copy in the frame kernel.
template <bool FRAME>
__device__ __noinline__ void phase_modes(const int log2n, const int chunk, const FrameSource& fs, const BatchSource& bs, const int bitDepth,
                                         const int16_t* orgPic, const int ctuX, const int ctuY) {
  unsigned char* smem = smem_raw;
  const RtGeo g = make_rt_geo_rt(log2n);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cls = warp_class(warp), half = warp_half(warp), par = warp & 1;
  const int N = g.n, pusPerChunk = 4096 >> (2 * g.log2n);
  const uint8_t* valid = smem + g.validOff;
  uint32_t* acc = reinterpret_cast<uint32_t*>(smem + g.accOff);
  LaneGeo lg; lg.init(g, half, lane);
  const bool ok = valid[lg.pu] == kPuEvaluate;     // N = 4: the four PUs of a region share their state (W, H multiples of 8; NxN is pruned as a whole)
  Tile src;
  if (ok) {
    Tile raw;
    if (FRAME) {
      int px, py;
      if (g.log2n == 2) { demorton(lg.pu >> 2, px, py); px *= 8; py *= 8; }
      else { demorton(lg.pu, px, py); px = px * N + lg.tx0; py = py * N + lg.ty0; }
      tile_load(raw, orgPic + (size_t)(ctuY + py) * fs.orgStride + ctuX + px, fs.orgStride);
    } else {
      if (g.log2n == 2) {
#pragma unroll
        for (int s = 0; s < 4; s++) {
          const bool okS = chunk * pusPerChunk + lg.pu + s < bs.count;
          const int16_t* base = bs.org + (okS ? (size_t)bs.pus[chunk * pusPerChunk + lg.pu + s].orgOff : 0);
#pragma unroll
          for (int y = 0; y < 4; y++) {
            uint2 v = make_uint2(0u, 0u);
            if (okS) v = *reinterpret_cast<const uint2*>(base + y * 4);
            raw.r[((s >> 1) * 4 + y) * 4 + (s & 1) * 2 + 0] = v.x;
            raw.r[((s >> 1) * 4 + y) * 4 + (s & 1) * 2 + 1] = v.y;
          }
        }
      } else {
        const int16_t* base = bs.org + (size_t)bs.pus[chunk * pusPerChunk + lg.pu].orgOff;
        tile_load(raw, base + lg.ty0 * N + lg.tx0, N);
      }
    }
    if (cls == 0) src = raw;
    else if (g.log2n == 2) tile_transpose4x4(raw, src);
    else tile_transpose8(raw, src);
  } else {
#pragma unroll
    for (int k = 0; k < 32; k++) src.r[k] = 0;
  }

  const int nModes = class_num_modes(cls);
  for (int i = par; i < nModes; i += 2) {
    const int mode = class_mode(cls, i);
    const bool neg = mode >= 2 && mode_angle(mode) < 0;
    if (neg) {
      if (ok) lane_build_ext(g, smem, warp, lg, cls, mode);
      __syncwarp();
    }
    if (g.log2n == 2) {
      if (ok) {
        uint32_t c4[4];
        lane_eval_region4(g, smem, warp, lg, cls, mode, bitDepth, src, c4);
#pragma unroll
        for (int s = 0; s < 4; s++) acc[(lg.pu + s) * kNumModes + mode] = c4[s];
      }
    } else {
      uint32_t v = 0;
      if (ok) v = lane_eval_tile(g, smem, warp, lg, cls, mode, bitDepth, src);
      for (int m = 1; m < g.lanesPerPu; m <<= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
      if (ok && lg.subLane == 0) {
        if (g.tilesPerPu > 32) atomicAdd(&acc[lg.pu * kNumModes + mode], v);
        else acc[lg.pu * kNumModes + mode] = v;
      }
    }
    if (neg) __syncwarp();
  }
}

template <int LOG2N, bool FRAME>
__device__ __forceinline__ void rmd_body(const int chunk, const FrameSource& fs, const BatchSource& bs, const int bitDepth, const int strong) {
  typedef Geo<LOG2N> G;
  constexpr int N = G::N;
  SmemView<LOG2N> sm; sm.base = smem_raw;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- chunk placement --------------------------------------------------------------------
  int pic = 0, ctu = 0, ctuX = 0, ctuY = 0;
  const int16_t* orgPic = nullptr; const int16_t* recPic = nullptr;
  if (FRAME) {
    pic = chunk / fs.ctusPerPic; ctu = chunk - pic * fs.ctusPerPic;
    ctuX = (ctu % fs.ctusPerRow) * 64; ctuY = (ctu / fs.ctusPerRow) * 64;
    orgPic = fs.org + (size_t)pic * fs.orgPicStride;
    recPic = fs.rec + (size_t)pic * fs.recPicStride;
  }

  // ---- phase A: validity + unfiltered linear borders --------------------------------------
  int anyEval = 0;
  for (int p = tid; p < G::PUS; p += kRmdThreads) {
    bool ok;
    if (FRAME) { int px, py; demorton(p, px, py); ok = (ctuX + (px + 1) * N <= fs.W) && (ctuY + (py + 1) * N <= fs.H); }
    else ok = chunk * G::PUS + p < bs.count;
    uint8_t st = ok ? kPuEvaluate : kPuOutside;
    if (FRAME && ok && fs.needed && !fs.needed[(size_t)chunk * kPusPerCtu + pu_offset_of_depth(6 - LOG2N) + p]) st = kPuPruned;
    sm.valid()[p] = st;
    anyEval |= st == kPuEvaluate;
  }
  if (FRAME && fs.needed && !__syncthreads_or(anyEval)) {
    // fork-aware mode: nothing to evaluate in this chunk - only the table codes are written
    auto val = [&](int i) -> uint32_t { return sm.valid()[i / kNumModes] == kPuPruned ? kCostPruned : kCostOutside; };
    if (fs.out) {
      uint32_t* o = fs.out + ((size_t)chunk * kPusPerCtu + pu_offset_of_depth(6 - LOG2N)) * kNumModes;
      for (int i = tid; i < G::PUS * kNumModes; i += kRmdThreads) o[i] = val(i);
    }
    if (fs.outPacked) store_packed_depth<LOG2N>(fs.outPacked + (size_t)chunk * kPackedCtuBytes, tid, kRmdThreads, val);
    return;
  }
  if (FRAME) {
    border_gather_frame<LOG2N>(tid, kRmdThreads, recPic, fs.recStride, fs.W, fs.H, ctuX, ctuY, sm.lin(), sm.flags());
  } else {
    const int first = chunk * G::PUS;
    const int npu = min(G::PUS, bs.count - first);
    for (int idx = tid; idx < npu * (4 * N + 1); idx += kRmdThreads) {
      const int p = idx / (4 * N + 1), i = idx - p * (4 * N + 1);
      sm.lin()[p * G::LIN + i] = bs.border[(size_t)bs.pus[first + p].borderOff + i];
    }
  }
  __syncthreads();
  if (FRAME) {
    border_substitute<LOG2N>(tid, kRmdThreads, bitDepth, sm.lin(), sm.flags());
    __syncthreads();
  }
  // ---- phase C: ascending ref arrays, smoothed copies, DC ---------------------------------
  border_derive<LOG2N>(tid, kRmdThreads, bitDepth, strong, sm.lin(), sm.arrs());
  border_pad<LOG2N>(tid, kRmdThreads, sm.arrs());
  __syncthreads();
  border_dc<LOG2N>(tid, kRmdThreads, sm.arrs(), sm.dc());
  for (int i = tid; i < G::PUS * kNumModes; i += kRmdThreads) sm.acc()[i] = 0;   // aliases lin/flags: derive is done
  __syncthreads();

  // ---- phase E: modes (one runtime-sized code path for every N, see RtGeo) ----------------------
  phase_modes<FRAME>(LOG2N, chunk, fs, bs, bitDepth, orgPic, ctuX, ctuY);
  __syncthreads();

  // ---- phase F: coalesced cost-table store --------------------------------------------------
  const int shift = bitDepth - 8;     // xGetHADs' final DISTORTION_PRECISION_ADJUSTMENT, TComRdCost.cpp:1603
  if (FRAME) {
    auto val = [&](int i) -> uint32_t {
      const uint8_t v = sm.valid()[i / kNumModes];
      return v == kPuEvaluate ? (sm.acc()[i] >> shift) : (v == kPuPruned ? kCostPruned : kCostOutside);
    };
    if (fs.out) {
      uint32_t* o = fs.out + ((size_t)chunk * kPusPerCtu + pu_offset_of_depth(6 - LOG2N)) * kNumModes;
      for (int i = tid; i < G::PUS * kNumModes; i += kRmdThreads) o[i] = val(i);
    }
    if (fs.outPacked) store_packed_depth<LOG2N>(fs.outPacked + (size_t)chunk * kPackedCtuBytes, tid, kRmdThreads, val);
  } else {
    const int first = chunk * G::PUS;
    for (int i = tid; i < G::PUS * kNumModes; i += kRmdThreads) {
      const int p = i / kNumModes, m = i - p * kNumModes;
      if (sm.valid()[p] == kPuEvaluate) bs.out[(size_t)bs.pus[first + p].outIndex * kNumModes + m] = sm.acc()[i] >> shift;
    }
  }
}

// Frame (replay) mode: ONE launch covers every (picture, CTU, depth).
__global__ void __launch_bounds__(kRmdThreads, 2)
rmd_frame_kernel(const FrameSource fs, const int bitDepth, const int strong) {
  const BatchSource bs = {};
  // depth-major block order: at any moment most SMs run the same size variant (instruction-cache footprint)
  const int chunks = gridDim.x / 5;
  const int depth = blockIdx.x / chunks, chunk = blockIdx.x - depth * chunks;
  switch (depth) {
    case 0: rmd_body<6, true>(chunk, fs, bs, bitDepth, strong); break;
    case 1: rmd_body<5, true>(chunk, fs, bs, bitDepth, strong); break;
    case 2: rmd_body<4, true>(chunk, fs, bs, bitDepth, strong); break;
    case 3: rmd_body<3, true>(chunk, fs, bs, bitDepth, strong); break;
    default: rmd_body<2, true>(chunk, fs, bs, bitDepth, strong); break;
  }
}

// +-(H8 (x) H8) and the block-diagonal +-(H4 (x) H4) per 4x4 quadrant, as s8 in the UMMA K-major layout:
// four 4 KB matrices back to back
__global__ void hadamard_operands_kernel(int8_t* dst) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= 64 * 64) return;
  const int j = tid >> 6, k = tid & 63;
  const int y = k >> 3, x = k & 7;
  const int off = (k >> 4) * 1024 + (j >> 3) * 128 + (j & 7) * 16 + (k & 15);
  const int s8 = (__popc((j >> 3) & y) + __popc((j & 7) & x)) & 1;
  dst[off] = (int8_t)(s8 ? -1 : 1);
  dst[4096 + off] = (int8_t)(s8 ? 1 : -1);
  const int q = j >> 4, u = (j >> 2) & 3, v = j & 3, qk = (y >> 2) * 2 + (x >> 2);
  const int s4 = (__popc(u & (y & 3)) + __popc(v & (x & 3))) & 1;
  const int e = q == qk ? (s4 ? -1 : 1) : 0;
  dst[8192 + off] = (int8_t)e;
  dst[12288 + off] = (int8_t)(-e);
}
cudaError_t launch_hadamard_operands(int8_t* dst, cudaStream_t st) {
  hadamard_operands_kernel<<<16, 256, 0, st>>>(dst);
  return cudaGetLastError();
}

template <int LOG2N>
__global__ void __launch_bounds__(kRmdThreads, 2)
rmd_batch_kernel(const BatchSource bs, const int bitDepth, const int strong) {
  const FrameSource fs = {};
  rmd_body<LOG2N, false>(blockIdx.x, fs, bs, bitDepth, strong);
}

constexpr int cmax(int a, int b) { return a > b ? a : b; }
constexpr int kFrameSmem = cmax(cmax(cmax(Smem<2>::TOTAL, Smem<3>::TOTAL), cmax(Smem<4>::TOTAL, Smem<5>::TOTAL)), Smem<6>::TOTAL);

cudaError_t launch_rmd_frames(const FrameSource& fs, int nPics, int bitDepth, int strong, cudaStream_t st, int* launches) {
  const int chunks = nPics * fs.ctusPerPic;
  if (chunks <= 0) return cudaSuccess;
  rmd_frame_kernel<<<chunks * 5, kRmdThreads, kFrameSmem, st>>>(fs, bitDepth, strong);
  if (launches) *launches += 1;
  return cudaGetLastError();
}

template <int LOG2N>
static cudaError_t launch_batch(int chunks, const BatchSource& bs, int bitDepth, int strong, cudaStream_t st) {
  constexpr int bytes = Smem<LOG2N>::TOTAL;
  rmd_batch_kernel<LOG2N><<<chunks, kRmdThreads, bytes, st>>>(bs, bitDepth, strong);
  return cudaGetLastError();
}

// The opt-in to > 48 KB of dynamic shared memory is a per-DEVICE function attribute: cucd_create calls this after
// cudaSetDevice for every handle, so a process holding handles on several GPUs configures each of them.
cudaError_t configure_rmd_kernels() {
  cudaError_t e = cudaFuncSetAttribute(rmd_frame_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFrameSmem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(rmd_batch_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<2>::TOTAL);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(rmd_batch_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<3>::TOTAL);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(rmd_batch_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<4>::TOTAL);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(rmd_batch_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<5>::TOTAL);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(rmd_batch_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<6>::TOTAL);
  return e;
}

cudaError_t launch_rmd_batch(int log2n, const BatchSource& bs, int bitDepth, int strong, cudaStream_t st, int* launches) {
  if (bs.count <= 0) return cudaSuccess;
  const int n = 1 << log2n, pus = 4096 / (n * n);
  const int chunks = (bs.count + pus - 1) / pus;
  cudaError_t e;
  switch (log2n) {
    case 2: e = launch_batch<2>(chunks, bs, bitDepth, strong, st); break;
    case 3: e = launch_batch<3>(chunks, bs, bitDepth, strong, st); break;
    case 4: e = launch_batch<4>(chunks, bs, bitDepth, strong, st); break;
    case 5: e = launch_batch<5>(chunks, bs, bitDepth, strong, st); break;
    case 6: e = launch_batch<6>(chunks, bs, bitDepth, strong, st); break;
    default: return cudaErrorInvalidValue;
  }
  if (e == cudaSuccess && launches) *launches += 1;
  return e;
}

}  // namespace cucd
