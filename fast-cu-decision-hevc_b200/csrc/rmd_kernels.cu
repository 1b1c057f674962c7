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
  const bool ok = valid[lg.pu] != 0;     // N = 4: the four PUs of a region share validity (W, H multiples of 8)
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

// ------------------------------------------------------------------------------------------------
// Tensor-core variant of the mode loop (8-bit content): predictions are produced exactly as in
// phase_modes, but instead of the register butterflies every lane stores its predicted tile as one
// 64-byte row of a [128 tiles x 64 pixels] u8 operand and tcgen05.mma (kind::i8) multiplies it - and the
// lane's source tile - by +-(H8 (x) H8) (or the block-diagonal 4x4 version for N = 4) into TMEM; the
// epilogue is 64 x |.| per lane (satd_tc.cuh).  Warps 0-3 and 4-7 form two independent 128-row groups.
// Rounds are software pipelined: the MMA of round i runs while the warps read back round i-1.
// ------------------------------------------------------------------------------------------------
constexpr int kTcGroupBytes = 3 * 8192;            // A_O, A_P[2]
constexpr int kTcBytes = 2 * 4096 + 2 * kTcGroupBytes + 64;
template <int LOG2N>
struct TcLayout {
  typedef Smem<LOG2N> S;
  // N = 4 does not stage costs in shared memory (they go straight to global), so the operand buffers start
  // right behind the border-construction scratch
  static constexpr int BASE = ((LOG2N == 2 ? (S::FLAGS_OFF + S::FLAGS_BYTES) : S::TOTAL) + 127) & ~127;
  static constexpr int TOTAL = BASE + kTcBytes;
};
CUCD_HD int tc_base_rt(int log2n) {
  switch (log2n) {
    case 2: return TcLayout<2>::BASE;
    case 3: return TcLayout<3>::BASE;
    case 4: return TcLayout<4>::BASE;
    case 5: return TcLayout<5>::BASE;
    default: return TcLayout<6>::BASE;
  }
}

__device__ __forceinline__ uint32_t pack_bytes(uint32_t a, uint32_t b) { return __byte_perm(a, b, 0x6420); }  // 4 x (value < 256)
// store the lane's 8x8 tile (32 packed 16-bit pairs) as row `row` of a K-major u8 operand
__device__ __forceinline__ void store_row_u8(unsigned char* opnd, int row, const uint32_t* w) {
  unsigned char* dst = opnd + (row >> 3) * 128 + (row & 7) * 16;
#pragma unroll
  for (int c = 0; c < 4; c++) {
    uint4 v;
    v.x = pack_bytes(w[8 * c + 0], w[8 * c + 1]); v.y = pack_bytes(w[8 * c + 2], w[8 * c + 3]);
    v.z = pack_bytes(w[8 * c + 4], w[8 * c + 5]); v.w = pack_bytes(w[8 * c + 6], w[8 * c + 7]);
    *reinterpret_cast<uint4*>(dst + c * 2048) = v;
  }
}

__device__ __noinline__ void phase_modes_tc(const int log2n, const int chunk, const FrameSource& fs, const int bitDepth,
                                            const int16_t* orgPic, const int ctuX, const int ctuY, const uint4* __restrict__ hadamard) {
  using namespace tc;
  unsigned char* smem = smem_raw;
  const RtGeo g = make_rt_geo_rt(log2n);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cls = warp_class(warp), half = warp_half(warp), par = warp & 1;
  const int grp = warp >> 2, row = tid & 127;
  const int N = g.n;
  unsigned char* tcb = smem + tc_base_rt(log2n);
  unsigned char* sB = tcb;                                  // +B at 0, -B at 4096
  unsigned char* sAO = tcb + 8192 + grp * kTcGroupBytes;
  unsigned char* sAP = sAO + 8192;                          // two buffers of 8192
  uint64_t* mbar = reinterpret_cast<uint64_t*>(tcb + 8192 + 2 * kTcGroupBytes);   // [grp][buf]
  uint32_t* tmemSlot = reinterpret_cast<uint32_t*>(mbar + 4);
  const uint8_t* valid = smem + g.validOff;
  uint32_t* acc = reinterpret_cast<uint32_t*>(smem + g.accOff);

  // ---- one-time setup: barriers, TMEM, the two Hadamard operands, the lane's source row -------------
  if (tid == 0) { for (int i = 0; i < 4; i++) mbar_init(&mbar[i], 1); asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
  if (warp == 0) tmem_alloc(tmemSlot, 256);
  {
    const uint4* src = hadamard + (log2n == 2 ? 512 : 0);   // [8x8: +B, -B][4x4 blocks: +B, -B], 256 uint4 each
    uint4* dst = reinterpret_cast<uint4*>(sB);
    dst[tid] = src[tid]; dst[tid + 256] = src[tid + 256];
  }
  LaneGeo lg; lg.init(g, half, lane);
  const bool ok = valid[lg.pu] != 0;
  {
    Tile raw, src;
    if (ok) {
      int px, py;
      if (g.log2n == 2) { demorton(lg.pu >> 2, px, py); px *= 8; py *= 8; }
      else { demorton(lg.pu, px, py); px = px * N + lg.tx0; py = py * N + lg.ty0; }
      tile_load(raw, orgPic + (size_t)(ctuY + py) * fs.orgStride + ctuX + px, fs.orgStride);
      if (cls == 0) src = raw;
      else if (g.log2n == 2) tile_transpose4x4(raw, src);
      else tile_transpose8(raw, src);
    } else {
#pragma unroll
      for (int k = 0; k < 32; k++) src.r[k] = 0;
    }
    store_row_u8(sAO, row, src.r);
  }
  tc_fence_before();
  fence_async_smem();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmemBase = *tmemSlot;
  const uint32_t idesc = make_idesc_i8(128, 64, 0);
  const uint64_t dBp = make_desc(smem_u32(sB), 1024, 128), dBn = make_desc(smem_u32(sB + 4096), 1024, 128);
  const uint64_t dAO = make_desc(smem_u32(sAO), 2048, 128);
  const uint32_t myTmem = tmemBase + ((uint32_t)((warp & 3) * 32) << 16) + grp * 128;
  const int nModes = class_num_modes(cls);
  constexpr int kRounds = 9;
  uint32_t* outN4 = fs.out + ((size_t)chunk * kPusPerCtu + pu_offset_of_depth(4)) * kNumModes;

  auto epilogue = [&](int j) {
    const int buf = j & 1, i = j * 2 + par;
    mbar_wait(&mbar[grp * 2 + buf], (uint32_t)(j >> 1) & 1u);
    tc_fence_after();
    uint32_t q[4];
#pragma unroll
    for (int c = 0; c < 4; c++) {
      uint32_t v[16];
      tmem_ld16(myTmem + buf * 64 + c * 16, v);
      tmem_ld_wait();
      uint32_t a = 0;
#pragma unroll
      for (int k = 0; k < 16; k++) a = __sad((int)v[k], 0, a);
      q[c] = a;
    }
    tc_fence_before();
    if (i >= nModes) return;
    const int mode = class_mode(cls, i);
    if (g.log2n == 2) {
      if (ok) {
#pragma unroll
        for (int c = 0; c < 4; c++) outN4[(lg.pu + c) * kNumModes + mode] = (q[c] + 1u) >> 1;      // xCalcHADs4x4 rounding, 8-bit: no shift
      }
    } else {
      uint32_t v = ok ? ((q[0] + q[1] + q[2] + q[3] + 2u) >> 2) : 0u;                                // xCalcHADs8x8 rounding
      for (int m = 1; m < g.lanesPerPu; m <<= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
      if (ok && lg.subLane == 0) {
        if (g.tilesPerPu > 32) atomicAdd(&acc[lg.pu * kNumModes + mode], v);
        else acc[lg.pu * kNumModes + mode] = v;
      }
    }
  };

  for (int j = 0; j < kRounds; j++) {
    const int i = j * 2 + par, buf = j & 1;
    if (i < nModes) {                                        // warp-uniform
      const int mode = class_mode(cls, i);
      const bool neg = mode >= 2 && mode_angle(mode) < 0;
      if (neg) {
        if (ok) lane_build_ext(g, smem, warp, lg, cls, mode);
        __syncwarp();
      }
      if (ok) {
        uint32_t p[32];
        if (g.log2n == 2) {
#pragma unroll
          for (int s = 0; s < 4; s++) {
            uint32_t q4[8];
            block_predict<4, 2>(g, smem, warp, lg.pu + s, lg.extSlot + 32 * s, 0, 0, cls, mode, bitDepth, q4);
#pragma unroll
            for (int y = 0; y < 4; y++) { p[((s >> 1) * 4 + y) * 4 + (s & 1) * 2] = q4[y * 2]; p[((s >> 1) * 4 + y) * 4 + (s & 1) * 2 + 1] = q4[y * 2 + 1]; }
          }
        } else {
          block_predict<8, 4>(g, smem, warp, lg.pu, lg.extSlot, cls ? lg.ty0 : lg.tx0, cls ? lg.tx0 : lg.ty0, cls, mode, bitDepth, p);
        }
        store_row_u8(sAP + buf * 8192, row, p);
      }
      __syncwarp();                                          // the ext scratch is rewritten by the next negative-angle mode
    }
    fence_async_smem();
    tc_fence_before();
    asm volatile("bar.sync %0, 128;\n" :: "r"(grp + 1) : "memory");
    if (row == 0) {
      tc_fence_after();
      const uint64_t dAP = make_desc(smem_u32(sAP + buf * 8192), 2048, 128);
      const uint32_t d = tmemBase + grp * 128 + buf * 64;
      mma_i8(d, dAO, dBp, idesc, 0u);
      mma_i8(d, dAO + ((2 * 2048) >> 4), dBp + ((2 * 1024) >> 4), idesc, 1u);
      mma_i8(d, dAP, dBn, idesc, 1u);
      mma_i8(d, dAP + ((2 * 2048) >> 4), dBn + ((2 * 1024) >> 4), idesc, 1u);
      mma_commit(&mbar[grp * 2 + buf]);
    }
    if (j > 0) epilogue(j - 1);
  }
  epilogue(kRounds - 1);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmemBase, 256);
}

template <int LOG2N, bool FRAME, bool TC = false>
__device__ __forceinline__ void rmd_body(const int chunk, const FrameSource& fs, const BatchSource& bs, const int bitDepth, const int strong, const uint4* hadamard = nullptr) {
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
  for (int p = tid; p < G::PUS; p += kRmdThreads) {
    bool ok;
    if (FRAME) { int px, py; demorton(p, px, py); ok = (ctuX + (px + 1) * N <= fs.W) && (ctuY + (py + 1) * N <= fs.H); }
    else ok = chunk * G::PUS + p < bs.count;
    sm.valid()[p] = ok ? 1 : 0;
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
  constexpr bool kStaged = !(TC && LOG2N == 2);          // the tensor-core N = 4 path writes its costs straight to global
  if (kStaged) for (int i = tid; i < G::PUS * kNumModes; i += kRmdThreads) sm.acc()[i] = 0;   // aliases lin/flags: derive is done
  __syncthreads();

  // ---- phase E: modes (one runtime-sized code path for every N, see RtGeo) ----------------------
  if (TC) phase_modes_tc(LOG2N, chunk, fs, bitDepth, orgPic, ctuX, ctuY, hadamard);
  else phase_modes<FRAME>(LOG2N, chunk, fs, bs, bitDepth, orgPic, ctuX, ctuY);
  __syncthreads();

  // ---- phase F: coalesced cost-table store --------------------------------------------------
  const int shift = bitDepth - 8;     // xGetHADs' final DISTORTION_PRECISION_ADJUSTMENT, TComRdCost.cpp:1603
  if (FRAME) {
    uint32_t* o = fs.out + ((size_t)chunk * kPusPerCtu + pu_offset_of_depth(6 - LOG2N)) * kNumModes;
    for (int i = tid; i < G::PUS * kNumModes; i += kRmdThreads) {
      const int p = i / kNumModes;
      if (kStaged) o[i] = sm.valid()[p] ? (sm.acc()[i] >> shift) : 0xffffffffu;
      else if (!sm.valid()[p]) o[i] = 0xffffffffu;
    }
  } else {
    const int first = chunk * G::PUS;
    for (int i = tid; i < G::PUS * kNumModes; i += kRmdThreads) {
      const int p = i / kNumModes, m = i - p * kNumModes;
      if (sm.valid()[p]) bs.out[(size_t)bs.pus[first + p].outIndex * kNumModes + m] = sm.acc()[i] >> shift;
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

// Tensor-core frame kernel (bit depth 8): same phases, phase_modes_tc for the mode loop.
__global__ void __launch_bounds__(kRmdThreads, 2)
rmd_frame_tc_kernel(const FrameSource fs, const int strong, const uint4* __restrict__ hadamard) {
  const BatchSource bs = {};
  const int chunks = gridDim.x / 5;
  const int depth = blockIdx.x / chunks, chunk = blockIdx.x - depth * chunks;
  switch (depth) {
    case 0: rmd_body<6, true, true>(chunk, fs, bs, 8, strong, hadamard); break;
    case 1: rmd_body<5, true, true>(chunk, fs, bs, 8, strong, hadamard); break;
    case 2: rmd_body<4, true, true>(chunk, fs, bs, 8, strong, hadamard); break;
    case 3: rmd_body<3, true, true>(chunk, fs, bs, 8, strong, hadamard); break;
    default: rmd_body<2, true, true>(chunk, fs, bs, 8, strong, hadamard); break;
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

constexpr int cmax2(int a, int b) { return a > b ? a : b; }
constexpr int kFrameTcSmem = cmax2(cmax2(cmax2(TcLayout<2>::TOTAL, TcLayout<3>::TOTAL), cmax2(TcLayout<4>::TOTAL, TcLayout<5>::TOTAL)), TcLayout<6>::TOTAL);

cudaError_t launch_rmd_frames_tc(const FrameSource& fs, int nPics, int strong, const int8_t* hadamard, cudaStream_t st, int* launches) {
  const int chunks = nPics * fs.ctusPerPic;
  if (chunks <= 0) return cudaSuccess;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(rmd_frame_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFrameTcSmem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  rmd_frame_tc_kernel<<<chunks * 5, kRmdThreads, kFrameTcSmem, st>>>(fs, strong, reinterpret_cast<const uint4*>(hadamard));
  if (launches) *launches += 1;
  return cudaGetLastError();
}

// cost tables in the packed host format: PUs 0..20 stay uint32, PUs 21..340 (8x8, 4x4) become uint16
constexpr int kPackWide = 21, kPackCtuBytes = kPackWide * kNumModes * 4 + (kPusPerCtu - kPackWide) * kNumModes * 2;
__global__ void pack_costs_kernel(const uint32_t* __restrict__ cost, uint8_t* __restrict__ packed, const int nCtus) {
  const int perCtu = kPusPerCtu * kNumModes;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)nCtus * perCtu; i += (long long)gridDim.x * blockDim.x) {
    const int ctu = (int)(i / perCtu), e = (int)(i - (long long)ctu * perCtu);
    const uint32_t c = cost[i];
    uint8_t* base = packed + (size_t)ctu * kPackCtuBytes;
    if (e < kPackWide * kNumModes) reinterpret_cast<uint32_t*>(base)[e] = c;
    else reinterpret_cast<uint16_t*>(base + kPackWide * kNumModes * 4)[e - kPackWide * kNumModes] = c == 0xffffffffu ? (uint16_t)0xffffu : (uint16_t)c;
  }
}
cudaError_t launch_pack_costs(const uint32_t* cost, uint8_t* packed, int nCtus, cudaStream_t st, int* launches) {
  if (nCtus <= 0) return cudaSuccess;
  const long long n = (long long)nCtus * kPusPerCtu * kNumModes;
  const int blocks = (int)std::min<long long>((n + 255) / 256, 148 * 16);
  pack_costs_kernel<<<blocks, 256, 0, st>>>(cost, packed, nCtus);
  if (launches) *launches += 1;
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
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(rmd_frame_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFrameSmem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  rmd_frame_kernel<<<chunks * 5, kRmdThreads, kFrameSmem, st>>>(fs, bitDepth, strong);
  if (launches) *launches += 1;
  return cudaGetLastError();
}

template <int LOG2N>
static cudaError_t launch_batch(int chunks, const BatchSource& bs, int bitDepth, int strong, cudaStream_t st) {
  constexpr int bytes = Smem<LOG2N>::TOTAL;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(rmd_batch_kernel<LOG2N>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  rmd_batch_kernel<LOG2N><<<chunks, kRmdThreads, bytes, st>>>(bs, bitDepth, strong);
  return cudaGetLastError();
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
