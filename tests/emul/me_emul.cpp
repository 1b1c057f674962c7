// me_emul.cpp - CPU replay of the dy-lane integer-ME SAD kernel (TEST HARNESS, not a product path).
//
// me_sad_dy_kernel (csrc/me_kernels.cu) is a thin shell around the __host__ __device__ code of csrc/me_core.cuh: tile
// decomposition of a window, staging of the PU and of the reference window into shared memory, and the per-lane sliding SAD.  This
// file replays one CTA per M / E tile the way the kernel runs it - staging for all 8 warps x 32 lanes, then every warp and lane -
// so that `pytest -m "not gpu"` can compare its arithmetic and its tiling with the oracle.  The candidates of the O tiles (the
// dx-lane kernels, which have no host restatement) are marked with a sentinel and counted; the test checks that every candidate
// is written exactly once.
#include <cstdint>
#include <cstring>
#include <vector>
#include "me_core.cuh"

using namespace cucd;

namespace {

template <bool U8>
void replay_dy_tile(int kind, int x0, int y0, int bitDepth, const int16_t* cur, int curStride, int w, int h, const int16_t* refWin, int refStride,
                    int cols, int rows, int subShift, uint32_t* out, uint8_t* hits) {
  constexpr int P = U8 ? kMeDyPitch8 : kMeDyPitch16, SPW = U8 ? 4 : 2;
  std::vector<uint32_t> sCur(64 * 64 / SPW, 0xA5A5A5A5u), sRef((size_t)kMeDyRows * P, 0xA5A5A5A5u);      // poison: stale shared memory
  const int step = 1 << subShift;
  const MeDyTile tl = me_dy_tile(kind, x0, y0, w, h, cols, rows);
  for (int warp = 0; warp < 8; warp++)
    for (int lane = 0; lane < 32; lane++) {
      me_dy_stage_cur<U8>(warp, lane, cur, curStride, w, h, step, sCur.data());
      me_dy_stage_ref<U8>(warp, lane, refWin + (long long)y0 * refStride + x0, refStride, tl.winW, tl.winH, sRef.data());
    }
  const int shiftOut = U8 ? 0 : bitDepth - 8;
  for (int warp = 0; warp < 8; warp++) {
    const MeDyWarp q = me_dy_warp<U8>(kind, warp, tl.cr);
    if (!(q.active && 32 * q.blk < tl.nLam)) continue;
    for (int lane = 0; lane < 32; lane++) {
      const int lam = 32 * q.blk + lane;
      uint32_t sad[kMeDyK];
      const int K = kind == kMeKindM ? kMeDyK : 1;
      if (kind == kMeKindM) {
        if (U8) me_dy_sad_u8<kMeDyK>(sCur.data(), sRef.data(), w >> 2, h, step, lam, q.wbase, q.s, sad);
        else me_dy_sad_s16<kMeDyK>(sCur.data(), sRef.data(), w >> 1, h, step, lam, q.wbase, q.s, me_fold_rows(bitDepth, w >> 1), sad);
      } else {
        if (U8) me_dy_sad_u8<1>(sCur.data(), sRef.data(), w >> 2, h, step, lam, q.wbase, q.s, sad);
        else me_dy_sad_s16<1>(sCur.data(), sRef.data(), w >> 1, h, step, lam, q.wbase, q.s, me_fold_rows(bitDepth, w >> 1), sad);
      }
      if (lam >= tl.nLam) continue;
      for (int k = 0; k < K; k++) {
        const size_t o = (size_t)(y0 + lam) * cols + x0 + q.delta0 + SPW * k;
        out[o] = (sad[k] << subShift) >> shiftOut;
        if (hits[o] < 255) hits[o]++;
      }
    }
  }
}

}  // namespace

extern "C" {

// refWin: the reference sample at the window's (left, top) displacement of the PU origin.  out[rows * cols]; hits[rows * cols] counts the
// writes per candidate; counts[3] = number of O / M / E tiles.  tileRowsO as in capi_batch.cu (16 for 8-bit content, 8 otherwise).
void emul_me_sad_surface(int bitDepth, const int16_t* cur, int curStride, int w, int h, const int16_t* refWin, int refStride, int cols, int rows,
                         int subShift, uint32_t* out, uint8_t* hits, int* counts) {
  std::memset(hits, 0, (size_t)rows * cols);
  counts[0] = counts[1] = counts[2] = 0;
  const int tileRowsO = bitDepth == 8 ? 16 : 8;
  me_enum_tiles(cols, rows, true, tileRowsO, [&](int kind, int x0, int y0) {
    int k2, xx, yy;
    me_tile_unpack(me_tile_pack(kind, x0, y0), k2, xx, yy);                 // through the record format the kernel decodes
    counts[k2]++;
    if (k2 == kMeKindO) {
      for (int y = yy; y < yy + tileRowsO && y < rows; y++)
        for (int x = xx; x < xx + 32 && x < cols; x++) { out[(size_t)y * cols + x] = 0xffffffffu; if (hits[(size_t)y * cols + x] < 255) hits[(size_t)y * cols + x]++; }
    } else if (bitDepth == 8) replay_dy_tile<true>(k2, xx, yy, bitDepth, cur, curStride, w, h, refWin, refStride, cols, rows, subShift, out, hits);
    else replay_dy_tile<false>(k2, xx, yy, bitDepth, cur, curStride, w, h, refWin, refStride, cols, rows, subShift, out, hits);
  });
}

}  // extern "C"

// ---- fractional-pel refinement of a small PU: the phases of me_subpel_small_kernel, every phase for all 256 thread ids ----------
extern "C" int emul_subpel_small(int bitDepth, const int16_t* cur, int curStride, int w, int h, const int16_t* refAtIntMv, int refStride,
                                 int useHadamard, uint32_t* out /*49*/) {
  if (!subpel_is_small(w, h)) return 0;
  std::vector<int16_t> sCur(256, 0x5a5a), sWin(kSpSmallWinCap, 0x5a5a), sHor(kSpHorCap, 0x5a5a), sPred(kSpPredCap, 0x5a5a);
  int sSum[64];
  const SubpelGeo g = subpel_geo(w, h, bitDepth);
  for (int tid = 0; tid < 256; tid++) subpel_small_stage(tid, g, cur, curStride, refAtIntMv - 4 * (long long)refStride - 4, refStride, sCur.data(), sWin.data());
  for (int tid = 0; tid < 256; tid++) subpel_small_hor(tid, g, sWin.data(), sHor.data());
  const int perGroup = kSpPredCap / g.wh < 49 ? kSpPredCap / g.wh : 49;
  for (int pBase = 0; pBase < 49; pBase += perGroup) {
    const int nP = 49 - pBase < perGroup ? 49 - pBase : perGroup;
    for (int tid = 0; tid < 256; tid++) { if (tid < nP) sSum[tid] = 0; subpel_small_ver(tid, g, pBase, nP, sHor.data(), sPred.data()); }
    for (int tid = 0; tid < 256; tid++) subpel_small_dist(tid, g, nP, useHadamard, sCur.data(), sPred.data(), sSum);
    for (int tid = 0; tid < nP; tid++) out[pBase + tid] = (uint32_t)sSum[tid] >> (bitDepth - 8);
  }
  return 1;
}
