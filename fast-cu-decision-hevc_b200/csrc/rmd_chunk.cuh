// rmd_chunk.cuh - how one CTA (8 warps) processes one RMD chunk (4096 samples of PUs of one size).
//
// Host/device like rmd_core.cuh: the phases are plain functions of (tid, nthreads) or (warp, lane)
// so that tests/emul can replay them in loops.  Shared-memory layout, warp->(mode, tile) schedule
// and the per-lane evaluation live here; rmd_kernels.cu only adds __syncthreads / shuffles.
#pragma once
#include "rmd_core.cuh"

namespace cucd {

constexpr int kRmdThreads = 256;
constexpr int kRmdWarps = kRmdThreads / 32;
constexpr int kNumModes = 35;
constexpr int kPusPerCtu = 341;

// first PU index of each depth inside a CTU's 341-entry table (depth-major, z-order inside a depth)
CUCD_HD int pu_offset_of_depth(int d) { return d == 0 ? 0 : (d == 1 ? 1 : (d == 2 ? 5 : (d == 3 ? 21 : 85))); }

// ---------------------------------------------------------------------------------------------
// shared memory carve-up (all offsets in bytes, every region 16-byte aligned)
// ---------------------------------------------------------------------------------------------
template <int LOG2N>
struct Smem {
  typedef Geo<LOG2N> G;
  static constexpr int al16(int v) { return (v + 15) & ~15; }
  static constexpr int ARRS_OFF = 0;
  static constexpr int ARRS_BYTES = al16(G::PUS * G::PU_STRIDE * 2);
  static constexpr int DC_OFF = ARRS_OFF + ARRS_BYTES;
  static constexpr int DC_BYTES = al16(G::PUS * 2);
  static constexpr int VALID_OFF = DC_OFF + DC_BYTES;
  static constexpr int VALID_BYTES = al16(G::PUS);
  static constexpr int EXT_OFF = VALID_OFF + VALID_BYTES;
  static constexpr int EXT_BYTES = al16(kRmdWarps * G::EXT_PER_WARP * 2);
  // union: {linear borders + unit flags} during border construction, cost accumulators afterwards
  static constexpr int UNI_OFF = EXT_OFF + EXT_BYTES;
  static constexpr int LIN_BYTES = al16(G::PUS * G::LIN * 2);
  static constexpr int FLAGS_OFF = UNI_OFF + LIN_BYTES;
  static constexpr int FLAGS_BYTES = al16(G::PUS * (G::N + 1));
  static constexpr int ACC_BYTES = al16(G::PUS * kNumModes * 4);
  static constexpr int UNI_BYTES = (LIN_BYTES + FLAGS_BYTES) > ACC_BYTES ? (LIN_BYTES + FLAGS_BYTES) : ACC_BYTES;
  static constexpr int TOTAL = UNI_OFF + UNI_BYTES;
};

template <int LOG2N>
struct SmemView {
  typedef Smem<LOG2N> S;
  unsigned char* base;
  CUCD_HD int16_t* arrs() const { return reinterpret_cast<int16_t*>(base + S::ARRS_OFF); }
  CUCD_HD int16_t* dc() const { return reinterpret_cast<int16_t*>(base + S::DC_OFF); }
  CUCD_HD uint8_t* valid() const { return base + S::VALID_OFF; }
  CUCD_HD int16_t* ext() const { return reinterpret_cast<int16_t*>(base + S::EXT_OFF); }
  CUCD_HD int16_t* lin() const { return reinterpret_cast<int16_t*>(base + S::UNI_OFF); }
  CUCD_HD uint8_t* flags() const { return base + S::FLAGS_OFF; }
  CUCD_HD uint32_t* acc() const { return reinterpret_cast<uint32_t*>(base + S::UNI_OFF); }
  CUCD_HD const int16_t* s16() const { return reinterpret_cast<const int16_t*>(base); }
  CUCD_HD const uint32_t* s32() const { return reinterpret_cast<const uint32_t*>(base); }
  // int16 index (relative to the start of shared memory) of a region
  static constexpr int ARRS16 = S::ARRS_OFF / 2;
  static constexpr int EXT16 = S::EXT_OFF / 2;
};

// ---------------------------------------------------------------------------------------------
// warp schedule.  Warps 0-3 evaluate in the true orientation (planar + modes 18..34), warps 4-7 on
// the transposed tile (DC + modes 2..17).  Inside a class: bit 1 of the warp id selects the half of
// the chunk's 64 tiles, bit 0 the parity of the position in the class's mode list.
// ---------------------------------------------------------------------------------------------
CUCD_HD int warp_class(int warp) { return warp >> 2; }
CUCD_HD int warp_half(int warp) { return (warp >> 1) & 1; }
CUCD_HD int class_num_modes(int cls) { return cls == 0 ? 18 : 17; }
CUCD_HD int class_mode(int cls, int i) { return cls == 0 ? (i == 0 ? 0 : 17 + i) : (i == 0 ? 1 : 1 + i); }

// which PU / tile a lane owns
template <int LOG2N>
struct LaneGeo {
  typedef Geo<LOG2N> G;
  int pu;        // chunk-local PU of the tile (N >= 8) or first of the region's four PUs (N = 4)
  int tx0, ty0;  // tile origin inside its PU, true orientation
  int extSlot;   // index of the lane's first extended-ref array inside the warp's scratch
  int lanesPerPu, subLane;
  CUCD_HD void init(int half, int lane) {
    const int t = half * 32 + lane;
    if constexpr (LOG2N == 2) { pu = 4 * t; tx0 = 0; ty0 = 0; extSlot = 4 * lane; lanesPerPu = 1; subLane = 0; }
    else {
      constexpr int TPP = G::TILES_PER_PU, TPR = G::N / 8;
      pu = t / TPP;
      const int q = t % TPP;
      tx0 = (q % TPR) * 8; ty0 = (q / TPR) * 8;
      lanesPerPu = TPP < 32 ? TPP : 32;
      subLane = lane % lanesPerPu;
      extSlot = lane / lanesPerPu;
    }
  }
};

// Build the extended main reference(s) of a negative-angle mode for the PU(s) this lane works on.
template <int LOG2N>
CUCD_HD void lane_build_ext(const SmemView<LOG2N>& sm, int warp, const LaneGeo<LOG2N>& lg, int cls, int mode) {
  typedef Geo<LOG2N> G;
  constexpr int N = G::N;
  const int angle = mode_angle(mode), inv = mode_inv_angle(mode);
  const int lastIdx = (N * angle) >> 5;
  const int fo = mode_uses_filtered<LOG2N>(mode) ? 2 * G::AS : 0;
  int16_t* ext = sm.ext() + warp * G::EXT_PER_WARP;
  const int nsub = LOG2N == 2 ? 4 : 1;
  for (int s = 0; s < nsub; s++) {
    const int arr0 = SmemView<LOG2N>::ARRS16 + (lg.pu + s) * G::PU_STRIDE + fo;
    const int main0 = arr0 + (cls ? G::AS : 0), side0 = arr0 + (cls ? 0 : G::AS);
    int16_t* e = ext + (lg.extSlot + s) * G::XS + N;       // element k = 0
    for (int i = lg.subLane; i < G::XS; i += lg.lanesPerPu) {
      const int k = i - N;
      int16_t v = 0;
      if (k <= N && k > lastIdx) v = ext_ref_sample<LOG2N>(sm.s16(), main0, side0, inv, k);
      e[k] = v;
    }
  }
}

// Residual + SATD of the lane's tile (N >= 8) for one mode.  `src` is the lane's source tile in the
// orientation of its warp class.
template <int LOG2N>
CUCD_HD uint32_t lane_eval_tile(const SmemView<LOG2N>& sm, int warp, const LaneGeo<LOG2N>& lg, int cls, int mode, int bitDepth, const Tile& src) {
  typedef Geo<LOG2N> G;
  constexpr int N = G::N;
  constexpr bool EDGE = N <= 16;
  const int fo = mode_uses_filtered<LOG2N>(mode) ? 2 * G::AS : 0;
  const int arr0 = SmemView<LOG2N>::ARRS16 + lg.pu * G::PU_STRIDE + fo;
  const int main0 = arr0 + (cls ? G::AS : 0), side0 = arr0 + (cls ? 0 : G::AS);
  const int x0 = cls ? lg.ty0 : lg.tx0, y0 = cls ? lg.tx0 : lg.ty0;
  uint32_t d[32];
  if (mode == 0) resid_planar<LOG2N, 8, 4>(sm.s16(), main0, side0, x0, y0, src.r, 4, d);
  else if (mode == 1) resid_dc<8, 4>(sm.s16(), main0, side0, x0, y0, sm.dc()[lg.pu], EDGE, src.r, 4, d);
  else {
    const int angle = mode_angle(mode);
    if (angle == 0) resid_angular_pure<8, 4>(sm.s32(), sm.s16(), main0, side0, x0, y0, EDGE, (1 << bitDepth) - 1, src.r, 4, d);
    else {
      const int m0 = angle < 0 ? SmemView<LOG2N>::EXT16 + warp * G::EXT_PER_WARP + lg.extSlot * G::XS + N : main0;
      if (angle == 32 || angle == -32) resid_angular_int<8, 4>(sm.s32(), m0, x0, y0, angle, src.r, 4, d);
      else resid_angular_frac<8, 4>(sm.s32(), m0, x0, y0, angle, src.r, 4, d);
    }
  }
  return satd8x8_packed(d);
}

// N = 4: the lane's region holds four independent 4x4 PUs; cost[s] for PU lg.pu + s
CUCD_HD void lane_eval_region4(const SmemView<2>& sm, int warp, const LaneGeo<2>& lg, int cls, int mode, int bitDepth, const Tile& src, uint32_t* cost) {
  typedef Geo<2> G;
  constexpr int N = 4;
#pragma unroll
  for (int s = 0; s < 4; s++) {
    const int arr0 = SmemView<2>::ARRS16 + (lg.pu + s) * G::PU_STRIDE;
    const int main0 = arr0 + (cls ? G::AS : 0), side0 = arr0 + (cls ? 0 : G::AS);
    const uint32_t* sp = &src.r[((s >> 1) * 4) * 4 + (s & 1) * 2];
    uint32_t d[8];
    if (mode == 0) resid_planar<2, 4, 2>(sm.s16(), main0, side0, 0, 0, sp, 4, d);
    else if (mode == 1) resid_dc<4, 2>(sm.s16(), main0, side0, 0, 0, sm.dc()[lg.pu + s], true, sp, 4, d);
    else {
      const int angle = mode_angle(mode);
      if (angle == 0) resid_angular_pure<4, 2>(sm.s32(), sm.s16(), main0, side0, 0, 0, true, (1 << bitDepth) - 1, sp, 4, d);
      else {
        const int m0 = angle < 0 ? SmemView<2>::EXT16 + warp * G::EXT_PER_WARP + (lg.extSlot + s) * G::XS + N : main0;
        if (angle == 32 || angle == -32) resid_angular_int<4, 2>(sm.s32(), m0, 0, 0, angle, sp, 4, d);
        else resid_angular_frac<4, 2>(sm.s32(), m0, 0, 0, angle, sp, 4, d);
      }
    }
    cost[s] = satd4x4_packed(d);
  }
}

// ---------------------------------------------------------------------------------------------
// chunk sources
// ---------------------------------------------------------------------------------------------
// Frame (replay) mode: a chunk is CTU `ctu` of picture `frame` at one depth; borders come from a
// reconstruction plane with z-scan availability.
struct FrameSource {
  const int16_t* org; const int16_t* rec;   // picture 0, sample (0,0)
  long long orgPicStride, recPicStride;     // samples between pictures
  int orgStride, recStride;                 // samples between rows
  int W, H, ctusPerRow, ctusPerPic;
  uint32_t* out;                            // [pic][ctu][341][35]
};
// Batch mode: `count` host-described PUs of one size, tightly packed source blocks and borders.
struct BatchPu { int32_t orgOff, borderOff, outIndex, pad; };   // sample offsets into org / border, row of the cost table
struct BatchSource {
  const int16_t* org;      // all source blocks of the batch, back to back
  const int16_t* border;   // all 4N+1 borders of the batch, back to back
  const BatchPu* pus;      // [count] the PUs of ONE size class
  uint32_t* out;           // [nPU][35]
  int count;
};

}  // namespace cucd
