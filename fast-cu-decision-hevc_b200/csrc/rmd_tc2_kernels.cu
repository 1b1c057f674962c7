// rmd_tc2_kernels.cu - tensor-core RMD frame kernel for 8-bit content (sm_100a, tcgen05 kind::i8).
//
// One CTA (512 threads = 4 row groups x 128 TMEM lanes) evaluates 4 CTUs at one depth.  A thread owns one
// 8x8 tile in one orientation for the whole CTA ("row"); per mode round and row group:
//     gather 32-byte reference window -> TMEM (A1)        [N = 4: static 64-byte record row]
//     MMA 1: D = A1 x weights(angle, phase)               prediction * 256 in byte 1 of every accumulator
//     epilogue 1: tcgen05.ld.pack::16b + 16 PRMT -> 64 predicted bytes -> TMEM (A2)
//     MMA 2: D = A2 x (H8 (x) H8)
//     epilogue 2: sum |D - Ho| against the row's transformed source tile (64 registers), HM rounding
// See rmd_tc2.cuh for the arithmetic and the reference citations; planar and DC are predicted on the ALU.
// Replaces, per PU, the reference loop TEncSearch.cpp:2327-2361.
#include <cuda_runtime.h>
#include "rmd_tc2.cuh"
#include "satd_tc.cuh"
#include "kernels.h"

namespace cucd {

extern __shared__ __align__(128) unsigned char smem2[];

using namespace tc;
using namespace tc2;

namespace {

struct Tc2Args {
  FrameSource fs;
  int strong, totalCtus;
  const uint8_t* tabWin; const uint8_t* tabN4; const int8_t* had;
};

__device__ __forceinline__ uint32_t make_idesc_i8x(int M, int N, int aSigned, int bSigned) {
  uint32_t d = 0;
  d |= 2u << 4;                                   // D = S32
  d |= (uint32_t)(aSigned ? 1 : 0) << 7;
  d |= (uint32_t)(bSigned ? 1 : 0) << 10;
  d |= (uint32_t)(N >> 3) << 17;
  d |= (uint32_t)(M >> 4) << 24;
  return d;
}
// A operand in TMEM (lane = row, 4 K-bytes per 32-bit column), B through a shared-memory descriptor
__device__ __forceinline__ void mma_i8_ts(uint32_t tmemD, uint32_t tmemA, uint64_t descB, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}\n"
      :: "r"(tmemD), "r"(tmemA), "l"(descB), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t addr, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};\n"
      :: "r"(addr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t addr, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};\n"
      :: "r"(addr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
         "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }
// 32 accumulator columns -> 16 registers: (col 2j & 0xffff) | (col 2j+1 << 16)
__device__ __forceinline__ void tmem_ld16_pack(uint32_t addr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.pack::16b.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(addr) : "memory");
}
__device__ __forceinline__ void group_bar(int grp) { asm volatile("bar.sync %0, 128;\n" :: "r"(grp + 1) : "memory"); }

// ---- prologue: reference arrays of the CTA's four CTUs -------------------------------------------------
template <int LOG2N>
__device__ __forceinline__ void tc2_prologue(const Tc2Args& a, const Geo2& g, const int group) {
  typedef Geo<LOG2N> G;
  constexpr int N = G::N;
  unsigned char* smem = smem2;
  SmemView<LOG2N> sm; sm.base = smem + g.scratchOff;
  const int tid = threadIdx.x;
  const FrameSource& fs = a.fs;
  for (int c = 0; c < kCtus; c++) {
    const int cg = group * kCtus + c;
    uint8_t* valid = smem + g.validOff + c * 256;
    if (cg >= a.totalCtus) {                                   // CTA-uniform
      for (int p = tid; p < G::PUS; p += kThreads) valid[p] = 0;
      continue;
    }
    const int pic = cg / fs.ctusPerPic, ctu = cg - pic * fs.ctusPerPic;
    const int ctuX = (ctu % fs.ctusPerRow) * 64, ctuY = (ctu / fs.ctusPerRow) * 64;
    const int16_t* recPic = fs.rec + (size_t)pic * fs.recPicStride;
    for (int p = tid; p < G::PUS; p += kThreads) {
      int px, py; demorton(p, px, py);
      valid[p] = ((ctuX + (px + 1) * N <= fs.W) && (ctuY + (py + 1) * N <= fs.H)) ? 1 : 0;
    }
    border_gather_frame<LOG2N>(tid, kThreads, recPic, fs.recStride, fs.W, fs.H, ctuX, ctuY, sm.lin(), sm.flags());
    __syncthreads();
    border_substitute<LOG2N>(tid, kThreads, 8, sm.lin(), sm.flags());
    __syncthreads();
    border_derive<LOG2N>(tid, kThreads, 8, a.strong, sm.lin(), sm.arrs());
    border_pad<LOG2N>(tid, kThreads, sm.arrs());
    __syncthreads();
    border_dc<LOG2N>(tid, kThreads, sm.arrs(), sm.dc());
    __syncthreads();
    convert_arrays<LOG2N>(tid, kThreads, g, c, sm.arrs(), sm.dc(), smem);
    __syncthreads();
  }
}

// ---- the mode rounds (one copy of the code for every PU size) --------------------------------------------
// Per row group the rounds are software pipelined so that the integer ALU always has work while an MMA is in
// flight (a lone tcgen05.mma + commit takes ~350 cycles, profiles/ubench):
//     wait MMA1(i) | epilogue 1 -> A2, projected refs of round i+1 | issue MMA2(i) | stage B1/A1 of round i+1 |
//     wait MMA2(i) | issue MMA1(i+1) | epilogue 2 of round i (costs)
// TMEM per row group: D1 = columns [0, 64) (A2 aliases its first 16 once they have been read), D2 = [64, 128).
__device__ __noinline__ void tc2_modes(const Tc2Args& a, const int log2n, const int group) {
  const Geo2 g = make_geo2_rt(log2n);
  unsigned char* smem = smem2;
  const int tid = threadIdx.x, grp = tid >> 7, rowTid = tid & 127, warp = tid >> 5, lane = tid & 31;
  const Row r = row_map(log2n, tid);
  unsigned char* store = smem + g.storeOff;
  unsigned char* sB1 = smem + g.b1Off + grp * 2 * g.b1Bytes;
  unsigned char* sA1 = smem + g.a1Off + grp * 8192;
  uint64_t* mbar1 = reinterpret_cast<uint64_t*>(smem + g.barOff) + grp;
  uint64_t* mbar2 = mbar1 + 4;
  uint32_t* tmemSlot = reinterpret_cast<uint32_t*>(smem + g.barOff + 64);
  uint32_t* acc = reinterpret_cast<uint32_t*>(smem + g.accOff);
  const FrameSource& fs = a.fs;
  const int cg = group * kCtus + r.ctu;
  const bool ok = smem[g.validOff + r.ctu * 256 + (log2n == 2 ? 4 * r.pu : r.pu)] != 0;   // N = 4: the region's PUs share validity (W, H multiples of 8)
  const int slot = pu_slot2(log2n, g.pus, r.ctu, r.pu);

  // ---- one-time setup ---------------------------------------------------------------------------------
  if (tid == 0) {
    for (int i = 0; i < 8; i++) mbar_init(reinterpret_cast<uint64_t*>(smem + g.barOff) + i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 0) tmem_alloc(tmemSlot, 512);
  if (tid < 256) reinterpret_cast<uint4*>(smem + g.hadOff)[tid] = reinterpret_cast<const uint4*>(a.had + (log2n == 2 ? 8192 : 0))[tid];
  if (g.accStaged) for (int i = tid; i < kCtus * g.pus * kNumModes; i += kThreads) acc[i] = 0;

  uint32_t p[16];                                   // the row's current byte tile: word 2*v + h = pixels (v, 4h..4h+3)
  if (ok) {
    const int pic = cg / fs.ctusPerPic, ctu = cg - pic * fs.ctusPerPic;
    int px, py; demorton(r.pu, px, py);
    if (log2n == 2) { px *= 8; py *= 8; }
    else { px = px * g.n + (r.o ? r.v0 : r.u0); py = py * g.n + (r.o ? r.u0 : r.v0); }
    const int16_t* src = fs.org + (size_t)pic * fs.orgPicStride + (size_t)((ctu / fs.ctusPerRow) * 64 + py) * fs.orgStride + (ctu % fs.ctusPerRow) * 64 + px;
    uint32_t raw[16];
#pragma unroll
    for (int y = 0; y < 8; y++) {
      const uint4 v = *reinterpret_cast<const uint4*>(src + (size_t)y * fs.orgStride);
      raw[2 * y] = __byte_perm(v.x, v.y, 0x6420); raw[2 * y + 1] = __byte_perm(v.z, v.w, 0x6420);
    }
    if (r.o) tile_transpose_bytes(raw, p, log2n != 2);
    else {
#pragma unroll
      for (int i = 0; i < 16; i++) p[i] = raw[i];
    }
  } else {
#pragma unroll
    for (int i = 0; i < 16; i++) p[i] = 0;
  }
  tc_fence_before();
  fence_async_smem();
  __syncthreads();
  tc_fence_after();

  const uint32_t tmemBase = *tmemSlot;
  const uint32_t laneOff = (uint32_t)((warp & 3) * 32) << 16;
  const uint32_t tD1 = tmemBase + grp * 128, tA2 = tD1, tD2 = tD1 + 64;
  const uint32_t idescPred = make_idesc_i8x(128, 64, 0, 0), idescHad = make_idesc_i8x(128, 64, 0, 1);
  const uint64_t dHad = make_desc(smem_u32(smem + g.hadOff), 1024, 128);
  constexpr uint64_t kStepB = (2 * 1024) >> 4;      // descriptor advance of one K = 32 step: two 16-byte chunks of 64 rows
  constexpr uint64_t kStepA = (2 * 2048) >> 4;      // ... of 128 rows
  uint32_t ph1 = 0, ph2 = 0;
  uint32_t ho[64];

  // A2 (already stored to TMEM by every thread) x H -> D2
  auto issue_mma2 = [&]() {
    tmem_st_wait();
    tc_fence_before();
    group_bar(grp);
    if (rowTid == 0) {
      tc_fence_after();
      mma_i8_ts(tD2, tA2, dHad, idescHad, 0u);
      mma_i8_ts(tD2, tA2 + 8, dHad + kStepB, idescHad, 1u);
      mma_commit(mbar2);
    }
  };
  auto wait_mma2 = [&]() { mbar_wait(mbar2, ph2); ph2 ^= 1u; tc_fence_after(); };
  // window / record operand (shared memory) x weights -> D1
  auto issue_mma1 = [&](int buf) {
    fence_async_smem();
    tc_fence_before();
    group_bar(grp);
    if (rowTid == 0) {
      tc_fence_after();
      const uint64_t dB = make_desc(smem_u32(sB1 + buf * g.b1Bytes), 1024, 128);
      if (log2n == 2) {
        const uint64_t dA = make_desc(smem_u32(sA1), 2048, 128);
        mma_i8(tD1, dA, dB, idescPred, 0u);
        mma_i8(tD1, dA + kStepA, dB + kStepB, idescPred, 1u);
      } else {
        mma_i8(tD1, make_desc(smem_u32(sA1 + buf * 4096), 2048, 128), dB, idescPred, 0u);
      }
      mma_commit(mbar1);
    }
  };
  auto wait_mma1 = [&]() { mbar_wait(mbar1, ph1); ph1 ^= 1u; tc_fence_after(); };
  // weights of round `am` from global memory (L2 resident) into registers, one round ahead of their use
  uint4 nb0 = make_uint4(0, 0, 0, 0), nb1 = nb0;
  auto prefetch_b1 = [&](int am) {
    const int ai = am + 8;
    if (log2n == 2) {
      const uint4* t = reinterpret_cast<const uint4*>(a.tabN4 + ai * 4096);
      nb0 = __ldg(t + rowTid); nb1 = __ldg(t + rowTid + 128);
    } else {
      const int fc = group_frac0(log2n, grp, angle_of_am(am)) >> 3;
      nb0 = __ldg(reinterpret_cast<const uint4*>(a.tabWin + (ai * 4 + fc) * 2048) + rowTid);
    }
  };
  // operands of round `am` into buffer `buf`: weights from the prefetch registers, the row's reference window
  auto stage = [&](int am, int buf) {
    uint4* b = reinterpret_cast<uint4*>(sB1 + buf * g.b1Bytes);
    b[rowTid] = nb0;
    if (log2n == 2) { b[rowTid + 128] = nb1; return; }
    const int filt = mode_uses_filtered_rt(log2n, 26 + am) ? 1 : 0;
    uint32_t w8[8];
    gather_window(store, arr_k0_off(g, grp, slot, r.o, filt) + win_k0(angle_of_am(am), r.u0, r.v0), w8);
    uint4* d = reinterpret_cast<uint4*>(sA1 + buf * 4096 + (rowTid >> 3) * 128 + (rowTid & 7) * 16);
    d[0] = make_uint4(w8[0], w8[1], w8[2], w8[3]);
    d[128] = make_uint4(w8[4], w8[5], w8[6], w8[7]);          // second 16-byte chunk: + 128 rows * 16 B
  };
  // epilogue 2 + cost hand-over for mode `mode` (has = the row has a mode in this round)
  uint32_t* outN4 = fs.out + ((size_t)cg * kPusPerCtu + pu_offset_of_depth(4) + 4 * r.pu) * kNumModes;
  auto cost_out = [&](int mode, bool has) {
    uint32_t q[4];
#pragma unroll
    for (int c = 0; c < 4; c++) {
      uint32_t v[16];
      tmem_ld16(tD2 + laneOff + c * 16, v);
      tmem_ld_wait();
      uint32_t s = 0;
#pragma unroll
      for (int k = 0; k < 16; k++) s = sad_acc(v[k], ho[c * 16 + k], s);
      q[c] = s;
    }
    tc_fence_before();
    if (log2n == 2) {
      if (ok && has) {
#pragma unroll
        for (int c = 0; c < 4; c++) outN4[c * kNumModes + mode] = (q[c] + 1u) >> 1;      // xCalcHADs4x4 rounding; 8-bit: no final shift
      }
    } else {
      uint32_t v = ok ? ((q[0] + q[1] + q[2] + q[3] + 2u) >> 2) : 0u;                    // xCalcHADs8x8 rounding
      for (int m = 1; m < r.seg; m <<= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
      if (ok && has && (lane & (r.seg - 1)) == 0) {
        uint32_t* dst = &acc[(r.ctu * g.pus + r.pu) * kNumModes + mode];
        if (log2n >= 4) atomicAdd(dst, v); else *dst = v;
      }
    }
  };

  // ---- Ho = H x source tile ------------------------------------------------------------------------------
  tmem_st16(tA2 + laneOff, p);
  issue_mma2();
  prefetch_b1(8);
  const unsigned char* rec4 = store + rec_off(r.ctu, r.o, 4 * r.pu);
  if (log2n == 2) {                                 // N = 4: the record row is the A operand of every mode (4 chunks)
    uint4* d = reinterpret_cast<uint4*>(sA1 + (rowTid >> 3) * 128 + (rowTid & 7) * 16);
#pragma unroll
    for (int i = 0; i < 4; i++) d[i * 128] = reinterpret_cast<const uint4*>(rec4)[i];
  }
  // round 0 prediction while the MMA runs: planar (true orientation rows) / DC (transposed rows) on the ALU
  const int unfMain = arr_k0_off(g, grp, slot, r.o, 0), unfSide = arr_k0_off(g, grp, slot, r.o ^ 1, 0);
  if (ok) {
    if (log2n == 2) { if (r.o == 0) planar_region4(rec4, p); else dc_region4(rec4, p); }
    else if (r.o == 0) {
      const int f = g.hasFilt;                      // planar reads the smoothed border for N = 8, 16, 32 (TComPattern.cpp:523-548)
      planar_tile(log2n, store + arr_k0_off(g, grp, slot, 0, f), store + arr_k0_off(g, grp, slot, 1, f), r.u0, r.v0, p);
    } else {
      dc_tile(reinterpret_cast<const int16_t*>(smem + g.dcOff)[r.ctu * 64 + r.pu], g.n <= 16, store + unfMain, store + unfSide, r.u0, r.v0, p);
    }
  }
  wait_mma2();
#pragma unroll
  for (int c = 0; c < 4; c++) tmem_ld16(tD2 + laneOff + c * 16, ho + c * 16);
  tmem_ld_wait();
  tc_fence_before();

  // ---- round 0 -------------------------------------------------------------------------------------------------
  tmem_st16(tA2 + laneOff, p);
  issue_mma2();
  stage(8, 0);
  prefetch_b1(7);
  wait_mma2();
  issue_mma1(0);
  cost_out(r.o ? 1 : 0, true);

  // ---- angular rounds -----------------------------------------------------------------------------------------
  for (int am = 8; am >= -8; --am) {
    const int angle = angle_of_am(am), buf = (8 - am) & 1;
    wait_mma1();
    // epilogue 1: byte 1 of every accumulator is the predicted pixel
#pragma unroll
    for (int h = 0; h < 2; h++) {
      uint32_t v[16];
      tmem_ld16_pack(tD1 + laneOff + h * 32, v);
      tmem_ld_wait();
      pack_pred(v, p + 8 * h, 8);
    }
    if (angle == 0 && g.n <= 16 && ok) {
      if (log2n == 2) patch_edge0_region4(rec4, p);
      else if (r.u0 == 0) patch_edge0_tile(store + unfMain, store + unfSide, r.v0, p);
    }
    tmem_st16(tA2 + laneOff, p);
    if (am > -8 && log2n != 2 && angle_of_am(am - 1) < 0)   // the gathers of round am are done (barrier of its MMA 1)
      build_ext_group(rowTid, g, grp, angle_of_am(am - 1), inv_angle_of_am(am - 1), mode_uses_filtered_rt(log2n, 25 + am) ? 1 : 0, store);
    issue_mma2();
    if (am > -8) {
      stage(am - 1, buf ^ 1);
      if (am > -7) prefetch_b1(am - 2);
    }
    wait_mma2();
    if (am > -8) issue_mma1(buf ^ 1);
    cost_out(r.o ? 10 - am : 26 + am, !(r.o && am == -8));
  }

  // ---- costs leave the SM ------------------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  for (int c = 0; c < kCtus; c++) {
    const int cgc = group * kCtus + c;
    if (cgc >= a.totalCtus) break;
    const uint8_t* valid = smem + g.validOff + c * 256;
    uint32_t* o = fs.out + ((size_t)cgc * kPusPerCtu + pu_offset_of_depth(6 - log2n)) * kNumModes;
    for (int i = tid; i < g.pus * kNumModes; i += kThreads) {
      const bool v = valid[i / kNumModes] != 0;
      if (g.accStaged) o[i] = v ? acc[c * g.pus * kNumModes + i] : 0xffffffffu;
      else if (!v) o[i] = 0xffffffffu;
    }
  }
  if (warp == 0) tmem_dealloc(tmemBase, 512);
}

template <int LOG2N>
__device__ __forceinline__ void tc2_body(const Tc2Args& a, const int group) {
  const Geo2 g = make_geo2<LOG2N>();
  tc2_prologue<LOG2N>(a, g, group);
  tc2_modes(a, LOG2N, group);
}

__global__ void __launch_bounds__(kThreads, 1)
rmd_frame_tc2_kernel(const Tc2Args a) {
  // depth-major block order, as rmd_frame_kernel
  const int groups = gridDim.x / 5;
  const int depth = blockIdx.x / groups, group = blockIdx.x - depth * groups;
  switch (depth) {
    case 0: tc2_body<6>(a, group); break;
    case 1: tc2_body<5>(a, group); break;
    case 2: tc2_body<4>(a, group); break;
    case 3: tc2_body<3>(a, group); break;
    default: tc2_body<2>(a, group); break;
  }
}

constexpr int cmax2(int a, int b) { return a > b ? a : b; }

}  // namespace

int rmd_tc2_smem_bytes() {
  int m = 0;
  for (int l = 2; l <= 6; l++) m = cmax2(m, make_geo2_rt(l).total);
  return m;
}

cudaError_t launch_rmd_frames_tc2(const FrameSource& fs, int nPics, int strong, const uint8_t* tabWin, const uint8_t* tabN4, const int8_t* hadamard,
                                  cudaStream_t st, int* launches) {
  const int total = nPics * fs.ctusPerPic;
  if (total <= 0) return cudaSuccess;
  const int smemBytes = rmd_tc2_smem_bytes();
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(rmd_frame_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smemBytes);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  Tc2Args a;
  a.fs = fs; a.strong = strong; a.totalCtus = total; a.tabWin = tabWin; a.tabN4 = tabN4; a.had = hadamard;
  const int groups = (total + kCtus - 1) / kCtus;
  rmd_frame_tc2_kernel<<<groups * 5, kThreads, smemBytes, st>>>(a);
  if (launches) *launches += 1;
  return cudaGetLastError();
}

}  // namespace cucd
