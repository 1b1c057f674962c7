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

// ---------------------------------------------------------------------------------------------
// Runtime geometry.  The border phases are templated on the PU size (short, run once per chunk); the
// mode loop - where the time goes - is ONE piece of code for N = 8..64 that takes the size at run time,
// so that the five depths of a frame launch share a single hot instruction footprint.
// ---------------------------------------------------------------------------------------------
struct RtGeo {
  int log2n, n, as, xs, puStride, tilesPerPu, tilesPerRow, lanesPerPu, extPerWarp;
  int arrs16, ext16;                 // int16 index of the ref arrays / the ext scratch in shared memory
  int dcOff, validOff, accOff;       // byte offsets
  int edge;                          // luma edge filters apply (N <= 16)
};
template <int LOG2N>
CUCD_HD RtGeo make_rt_geo() {
  typedef Geo<LOG2N> G; typedef Smem<LOG2N> S;
  RtGeo g;
  g.log2n = LOG2N; g.n = G::N; g.as = G::AS; g.xs = G::XS; g.puStride = G::PU_STRIDE;
  g.tilesPerPu = G::TILES_PER_PU; g.tilesPerRow = G::N >= 8 ? G::N / 8 : 1;
  g.lanesPerPu = G::TILES_PER_PU < 32 ? G::TILES_PER_PU : 32; g.extPerWarp = G::EXT_PER_WARP;
  g.arrs16 = S::ARRS_OFF / 2; g.ext16 = S::EXT_OFF / 2; g.dcOff = S::DC_OFF; g.validOff = S::VALID_OFF; g.accOff = S::UNI_OFF;
  g.edge = G::N <= 16;
  return g;
}

CUCD_HD RtGeo make_rt_geo_rt(int log2n) {
  switch (log2n) {
    case 2: return make_rt_geo<2>();
    case 3: return make_rt_geo<3>();
    case 4: return make_rt_geo<4>();
    case 5: return make_rt_geo<5>();
    default: return make_rt_geo<6>();
  }
}

// which PU / tile a lane owns
struct LaneGeo {
  int pu;        // chunk-local PU of the tile (N >= 8) or first of the region's four PUs (N = 4)
  int tx0, ty0;  // tile origin inside its PU, true orientation
  int extSlot;   // index of the lane's first extended-ref array inside the warp's scratch
  int subLane;   // position among the lanes that share the PU
  CUCD_HD void init(const RtGeo& g, int half, int lane) {
    const int t = half * 32 + lane;
    if (g.log2n == 2) { pu = 4 * t; tx0 = 0; ty0 = 0; extSlot = lane; subLane = 0; }   // N = 4: sub-PU s uses ext slot s*32 + lane
    else {
      pu = t / g.tilesPerPu;
      const int q = t - pu * g.tilesPerPu;
      const int row = q / g.tilesPerRow;
      tx0 = (q - row * g.tilesPerRow) * 8; ty0 = row * 8;
      subLane = lane % g.lanesPerPu;
      extSlot = lane / g.lanesPerPu;
    }
  }
};

// Build the extended main reference(s) of a negative-angle mode for the PU(s) this lane works on
// (TComPrediction.cpp:300-322): ref[k] = side[(128 + |k|*invAngle) >> 8] for k = -1 .. lastIdx+1 and a
// copy of main[0..N] behind it so that a row window can straddle k = 0.  Lanes that share a PU split
// the entries.  Entries outside [lastIdx+1, N] are never consumed (they are only touched as the unused
// half of an aligned word), so they are left as they are.
CUCD_HD void lane_build_ext(const RtGeo& g, unsigned char* smem, int warp, const LaneGeo& lg, int cls, int mode) {
  const int N = g.n;
  const int angle = mode_angle(mode), inv = mode_inv_angle(mode);
  const int nNeg = -((N * angle) >> 5) - 1;          // projected entries k = -1 .. -nNeg
  const int fo = mode_uses_filtered_rt(g.log2n, mode) ? 2 * g.as : 0;
  int16_t* s16 = reinterpret_cast<int16_t*>(smem);
  const int nsub = g.log2n == 2 ? 4 : 1, step = g.log2n == 2 ? 1 : g.lanesPerPu;
  for (int s = 0; s < nsub; s++) {
    const int arr0 = g.arrs16 + pu_slot_rt(g.log2n, lg.pu + s) * g.puStride + fo;
    const int main0 = arr0 + (cls ? g.as : 0), side0 = arr0 + (cls ? 0 : g.as);
    const int e0 = g.ext16 + warp * g.extPerWarp + (lg.extSlot + 32 * s) * g.xs + N;  // element k = 0
    for (int j = 1 + lg.subLane; j <= nNeg; j += step) s16[e0 - j] = s16[side0 + ((128 + j * inv) >> 8)];
    for (int k = lg.subLane; k <= N; k += step) s16[e0 + k] = s16[main0 + k];
  }
}

// Prediction of one block (8x8 tile: ROWS 8 / WORDS 4, 4x4 PU: ROWS 4 / WORDS 2) for one mode, as packed pairs.
template <int ROWS, int WORDS>
CUCD_HD void block_predict(const RtGeo& g, const unsigned char* smem, int warp, int pu, int extSlot, int x0, int y0, int cls, int mode,
                           int bitDepth, uint32_t* p) {
  const int16_t* s16 = reinterpret_cast<const int16_t*>(smem);
  const uint32_t* s32 = reinterpret_cast<const uint32_t*>(smem);
  const int fo = mode_uses_filtered_rt(g.log2n, mode) ? 2 * g.as : 0;
  const int arr0 = g.arrs16 + pu_slot_rt(g.log2n, pu) * g.puStride + fo;
  const int main0 = arr0 + (cls ? g.as : 0), side0 = arr0 + (cls ? 0 : g.as);
  if (mode == 0) pred_planar<ROWS, WORDS>(g.log2n, s16, main0, side0, x0, y0, p);
  else if (mode == 1) pred_dc<ROWS, WORDS>(s16, main0, side0, x0, y0, reinterpret_cast<const int16_t*>(smem + g.dcOff)[pu], g.edge != 0, p);
  else {
    const int angle = mode_angle(mode);
    const int m0 = angle < 0 ? g.ext16 + warp * g.extPerWarp + extSlot * g.xs + g.n : main0;
    pred_angular<ROWS, WORDS>(s32, m0, x0, y0, angle, p);
    if (angle == 0 && g.edge && x0 == 0) patch_pure_edge<ROWS, WORDS>(s16, main0, side0, y0, (1 << bitDepth) - 1, p);
  }
}

// N >= 8: SATD of the lane's tile for one mode.  `src` is the tile in the orientation of the warp class.
CUCD_HD uint32_t lane_eval_tile(const RtGeo& g, const unsigned char* smem, int warp, const LaneGeo& lg, int cls, int mode, int bitDepth, const Tile& src) {
  uint32_t d[32];
  block_predict<8, 4>(g, smem, warp, lg.pu, lg.extSlot, cls ? lg.ty0 : lg.tx0, cls ? lg.tx0 : lg.ty0, cls, mode, bitDepth, d);
#pragma unroll
  for (int k = 0; k < 32; k++) d[k] = src.r[k] - d[k];
  return satd8x8_packed(d);
}

// N = 4: the lane's region holds four independent 4x4 PUs; cost[s] for PU lg.pu + s
CUCD_HD void lane_eval_region4(const RtGeo& g, const unsigned char* smem, int warp, const LaneGeo& lg, int cls, int mode, int bitDepth, const Tile& src, uint32_t* cost) {
#pragma unroll
  for (int s = 0; s < 4; s++) {
    uint32_t d[8];
    block_predict<4, 2>(g, smem, warp, lg.pu + s, lg.extSlot + 32 * s, 0, 0, cls, mode, bitDepth, d);
    const uint32_t* sp = &src.r[((s >> 1) * 4) * 4 + (s & 1) * 2];
#pragma unroll
    for (int y = 0; y < 4; y++) { d[y * 2] = sp[y * 4] - d[y * 2]; d[y * 2 + 1] = sp[y * 4 + 1] - d[y * 2 + 1]; }
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
  uint32_t* out;                            // [pic][ctu][341][35], or null
  uint8_t* outPacked;                       // [pic][ctu][kPackedCtuBytes]: the packed CTU tables of include/cucudecide.h, or null
  const uint8_t* needed;                    // [pic][ctu][341] or null: fork-aware mode, 0 = this PU is pruned by the early decisions
};
// per-PU state inside a chunk / CTA
constexpr uint8_t kPuOutside = 0, kPuEvaluate = 1, kPuPruned = 2;
constexpr uint32_t kCostOutside = 0xffffffffu, kCostPruned = 0xfffffffeu;   // table codes (include/cucudecide.h)

// ---------------------------------------------------------------------------------------------
// packed CTU cost table (include/cucudecide.h): uint32 for PUs 0..20, uint16 for the 8x8 PUs, 13-bit stream for the 4x4 PUs
// ---------------------------------------------------------------------------------------------
constexpr int kPackedWidePus = 21;
constexpr int kPackedU16Off = kPackedWidePus * kNumModes * 4;               // 2940
constexpr int kPackedB13Off = kPackedU16Off + 64 * kNumModes * 2;           // 7420
constexpr int kPackedCtuBytes = kPackedB13Off + 256 * kNumModes * 13 / 8;   // 21980
// Store the cost block of one depth (PUs of size 1 << LOG2N, element i = pu * 35 + mode, PUS * 35 elements) of one CTU.
// val(i) returns the cost, or 0xffffffff for a PU that is not inside the picture.  Every thread writes whole 32-bit words.
template <int LOG2N, class Val>
CUCD_HD void store_packed_depth(uint8_t* ctuBase, int tid, int nthreads, Val val) {
  constexpr int PUS = 4096 >> (2 * LOG2N), ELEMS = PUS * kNumModes;
  if (LOG2N >= 4) {
    uint32_t* o = reinterpret_cast<uint32_t*>(ctuBase) + pu_offset_of_depth(6 - LOG2N) * kNumModes;
    for (int i = tid; i < ELEMS; i += nthreads) o[i] = val(i);
  } else if (LOG2N == 3) {
    uint32_t* o = reinterpret_cast<uint32_t*>(ctuBase + kPackedU16Off);
    for (int w = tid; w < ELEMS / 2; w += nthreads) o[w] = (val(2 * w) & 0xffffu) | (val(2 * w + 1) << 16);
  } else {
    uint32_t* o = reinterpret_cast<uint32_t*>(ctuBase + kPackedB13Off);
    for (int w = tid; w < ELEMS * 13 / 32; w += nthreads) {
      const int bit0 = 32 * w, e0 = bit0 / 13, sh = bit0 - 13 * e0;
      unsigned long long acc = 0;
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const int e = e0 + k;
        const unsigned long long v = e < ELEMS ? (unsigned long long)(val(e) & 0x1fffu) : 0ull;
        acc |= v << (13 * k);
      }
      o[w] = (uint32_t)(acc >> sh);
    }
  }
}
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
