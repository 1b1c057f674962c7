// rmd_emul.cpp - CPU replay of the CUDA kernels' per-thread code (TEST HARNESS, not a product path).
//
// The arithmetic of the sm_100a kernels lives in __host__ __device__ headers
// (csrc/rmd_core.cuh, rmd_chunk.cuh, feature_core.cuh).  This file compiles those headers with g++
// and replays one CTA the way rmd_kernels.cu / feature_kernels.cu run it: every phase for all 256
// thread ids, then the next phase (a __syncthreads boundary), warps one after the other, shuffles
// replaced by plain sums.  It lets `pytest -m "not gpu"` compare the kernels' arithmetic with the
// oracle in a container that has no GPU.  It is built only by tests/ and never loaded by the library.
#include <cstdint>
#include <cstring>
#include <vector>
#include <algorithm>
#include "rmd_chunk.cuh"
#include "feature_core.cuh"

using namespace cucd;

namespace {

template <int LOG2N, bool FRAME>
void emul_chunk(int chunk, const FrameSource& fs, const BatchSource& bs, int bitDepth, int strong) {
  typedef Geo<LOG2N> G;
  constexpr int N = G::N;
  std::vector<unsigned char> smemStore(Smem<LOG2N>::TOTAL + 64, 0xA5);   // poison: uninitialised reads show up
  SmemView<LOG2N> sm;
  sm.base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smemStore.data()) + 15) & ~uintptr_t(15));
  int pic = 0, ctu = 0, ctuX = 0, ctuY = 0;
  const int16_t* orgPic = nullptr; const int16_t* recPic = nullptr;
  if (FRAME) {
    pic = chunk / fs.ctusPerPic; ctu = chunk - pic * fs.ctusPerPic;
    ctuX = (ctu % fs.ctusPerRow) * 64; ctuY = (ctu / fs.ctusPerRow) * 64;
    orgPic = fs.org + (size_t)pic * fs.orgPicStride; recPic = fs.rec + (size_t)pic * fs.recPicStride;
  }
  // phase A
  for (int tid = 0; tid < kRmdThreads; tid++) {
    for (int p = tid; p < G::PUS; p += kRmdThreads) {
      bool ok;
      if (FRAME) { int px, py; demorton(p, px, py); ok = (ctuX + (px + 1) * N <= fs.W) && (ctuY + (py + 1) * N <= fs.H); }
      else ok = chunk * G::PUS + p < bs.count;
      sm.valid()[p] = ok ? 1 : 0;
    }
    if (FRAME) border_gather_frame<LOG2N>(tid, kRmdThreads, recPic, fs.recStride, fs.W, fs.H, ctuX, ctuY, sm.lin(), sm.flags());
    else {
      const int first = chunk * G::PUS, npu = std::min(G::PUS, bs.count - first);
      for (int idx = tid; idx < npu * (4 * N + 1); idx += kRmdThreads) {
        const int p = idx / (4 * N + 1), i = idx - p * (4 * N + 1);
        sm.lin()[p * G::LIN + i] = bs.border[(size_t)bs.pus[first + p].borderOff + i];
      }
    }
  }
  if (FRAME) for (int tid = 0; tid < kRmdThreads; tid++) border_substitute<LOG2N>(tid, kRmdThreads, bitDepth, sm.lin(), sm.flags());
  for (int tid = 0; tid < kRmdThreads; tid++) { border_derive<LOG2N>(tid, kRmdThreads, bitDepth, strong, sm.lin(), sm.arrs()); border_pad<LOG2N>(tid, kRmdThreads, sm.arrs()); }
  for (int tid = 0; tid < kRmdThreads; tid++) {
    border_dc<LOG2N>(tid, kRmdThreads, sm.arrs(), sm.dc());
    for (int i = tid; i < G::PUS * kNumModes; i += kRmdThreads) sm.acc()[i] = 0;
  }
  // phase E: warps one after the other; inside a warp the ext build of all lanes precedes the evaluation
  const RtGeo g = make_rt_geo_rt(LOG2N);
  unsigned char* smem = sm.base;
  const int pusPerChunk = G::PUS;
  for (int warp = 0; warp < kRmdWarps; warp++) {
    const int cls = warp_class(warp), half = warp_half(warp), par = warp & 1;
    LaneGeo lg[32]; bool ok[32]; Tile src[32];
    for (int lane = 0; lane < 32; lane++) {
      lg[lane].init(g, half, lane);
      ok[lane] = sm.valid()[lg[lane].pu] != 0;
      std::memset(&src[lane], 0, sizeof(Tile));
      if (!ok[lane]) continue;
      Tile raw;
      if (FRAME) {
        int px, py;
        if (g.log2n == 2) { demorton(lg[lane].pu >> 2, px, py); px *= 8; py *= 8; }
        else { demorton(lg[lane].pu, px, py); px = px * N + lg[lane].tx0; py = py * N + lg[lane].ty0; }
        tile_load(raw, orgPic + (size_t)(ctuY + py) * fs.orgStride + ctuX + px, fs.orgStride);
      } else {
        if (g.log2n == 2) {
          for (int s = 0; s < 4; s++) {
            const bool okS = chunk * pusPerChunk + lg[lane].pu + s < bs.count;
            const int16_t* base = bs.org + (okS ? (size_t)bs.pus[chunk * pusPerChunk + lg[lane].pu + s].orgOff : 0);
            for (int y = 0; y < 4; y++) {
              uint32_t v0 = 0, v1 = 0;
              if (okS) { v0 = (uint16_t)base[y * 4] | ((uint32_t)(uint16_t)base[y * 4 + 1] << 16); v1 = (uint16_t)base[y * 4 + 2] | ((uint32_t)(uint16_t)base[y * 4 + 3] << 16); }
              raw.r[((s >> 1) * 4 + y) * 4 + (s & 1) * 2 + 0] = v0;
              raw.r[((s >> 1) * 4 + y) * 4 + (s & 1) * 2 + 1] = v1;
            }
          }
        } else {
          const int16_t* base = bs.org + (size_t)bs.pus[chunk * pusPerChunk + lg[lane].pu].orgOff;
          tile_load(raw, base + lg[lane].ty0 * N + lg[lane].tx0, N);
        }
      }
      if (cls == 0) src[lane] = raw;
      else if (g.log2n == 2) tile_transpose4x4(raw, src[lane]);
      else tile_transpose8(raw, src[lane]);
    }
    for (int i = par; i < class_num_modes(cls); i += 2) {
      const int mode = class_mode(cls, i);
      const bool neg = mode >= 2 && mode_angle(mode) < 0;
      if (neg) for (int lane = 0; lane < 32; lane++) if (ok[lane]) lane_build_ext(g, smem, warp, lg[lane], cls, mode);
      for (int lane = 0; lane < 32; lane++) {
        if (!ok[lane]) continue;
        if (g.log2n == 2) {
          uint32_t c4[4];
          lane_eval_region4(g, smem, warp, lg[lane], cls, mode, bitDepth, src[lane], c4);
          for (int s = 0; s < 4; s++) sm.acc()[(lg[lane].pu + s) * kNumModes + mode] = c4[s];
        } else {
          sm.acc()[lg[lane].pu * kNumModes + mode] += lane_eval_tile(g, smem, warp, lg[lane], cls, mode, bitDepth, src[lane]);
        }
      }
    }
  }
  // phase F
  const int shift = bitDepth - 8;
  if (FRAME) {
    uint32_t* o = fs.out + ((size_t)chunk * kPusPerCtu + pu_offset_of_depth(6 - LOG2N)) * kNumModes;
    for (int i = 0; i < G::PUS * kNumModes; i++) o[i] = sm.valid()[i / kNumModes] ? (sm.acc()[i] >> shift) : 0xffffffffu;
  } else {
    const int first = chunk * G::PUS;
    for (int i = 0; i < G::PUS * kNumModes; i++) {
      const int p = i / kNumModes, m = i - p * kNumModes;
      if (sm.valid()[p]) bs.out[(size_t)bs.pus[first + p].outIndex * kNumModes + m] = sm.acc()[i] >> shift;
    }
  }
}

}  // namespace

extern "C" {

// frame (replay) mode over CTUs [ctuBegin, ctuEnd) of one picture; out as the kernel writes it
int emul_rmd_frame(int bitDepth, int strong, const int16_t* org, int orgStride, const int16_t* rec, int recStride, int W, int H,
                   int ctuBegin, int ctuEnd, uint32_t* outAllCtus) {
  FrameSource fs;
  fs.org = org; fs.rec = rec; fs.orgPicStride = 0; fs.recPicStride = 0; fs.orgStride = orgStride; fs.recStride = recStride;
  fs.W = W; fs.H = H; fs.ctusPerRow = (W + 63) / 64; fs.ctusPerPic = fs.ctusPerRow * ((H + 63) / 64); fs.out = outAllCtus; fs.outPacked = nullptr;
  BatchSource bs = {};
  for (int c = ctuBegin; c < ctuEnd; c++) {
    emul_chunk<6, true>(c, fs, bs, bitDepth, strong);
    emul_chunk<5, true>(c, fs, bs, bitDepth, strong);
    emul_chunk<4, true>(c, fs, bs, bitDepth, strong);
    emul_chunk<3, true>(c, fs, bs, bitDepth, strong);
    emul_chunk<2, true>(c, fs, bs, bitDepth, strong);
  }
  return 0;
}

// batch mode: `count` PUs of one size, tightly packed org blocks and borders, out[count][35]
int emul_rmd_batch(int bitDepth, int strong, int log2n, int count, const int16_t* org, const int16_t* border, uint32_t* out) {
  const int n = 1 << log2n, pusPerChunk = 4096 / (n * n);
  std::vector<BatchPu> pus(count);
  for (int i = 0; i < count; i++) { pus[i].orgOff = i * n * n; pus[i].borderOff = i * (4 * n + 1); pus[i].outIndex = i; pus[i].pad = 0; }
  BatchSource bs; bs.org = org; bs.border = border; bs.pus = pus.data(); bs.out = out; bs.count = count;
  FrameSource fs = {};
  const int chunks = (count + pusPerChunk - 1) / pusPerChunk;
  for (int c = 0; c < chunks; c++) {
    switch (log2n) {
      case 2: emul_chunk<2, false>(c, fs, bs, bitDepth, strong); break;
      case 3: emul_chunk<3, false>(c, fs, bs, bitDepth, strong); break;
      case 4: emul_chunk<4, false>(c, fs, bs, bitDepth, strong); break;
      case 5: emul_chunk<5, false>(c, fs, bs, bitDepth, strong); break;
      case 6: emul_chunk<6, false>(c, fs, bs, bitDepth, strong); break;
      default: return -1;
    }
  }
  return 0;
}

// feature pass 1 + pass 2 arithmetic (thread = one 4x4 block), thresholds given
int emul_feature_hist(int bitDepth, const int16_t* org, int stride, int W, int H, uint32_t* hist /*16*4096*/) {
  std::memset(hist, 0, 16 * 4096 * sizeof(uint32_t));
  for (int by = 0; by < H / 4; by++)
    for (int bx = 0; bx < W / 4; bx++) {
      int c[16];
      dct4x4(org + (size_t)(by * 4) * stride + bx * 4, stride, bitDepth, c);
      for (int f = 1; f < 16; f++) hist[f * 4096 + std::min(coeff_bin(c[f]), 4095)]++;
    }
  return 0;
}
int emul_feature_obf(int bitDepth, const int16_t* org, int stride, int W, int H, const int32_t* thr, int16_t* obf, int16_t* outlier) {
  for (int by = 0; by < H / 4; by++)
    for (int bx = 0; bx < W / 4; bx++) {
      int c[16], cnt = 0;
      dct4x4(org + (size_t)(by * 4) * stride + bx * 4, stride, bitDepth, c);
      for (int f = 0; f < 16; f++) {
        int v = 0;
        if (f > 0 && coeff_is_outlier(c[f], thr[f])) { v = c[f]; cnt++; }
        int16_t p = (int16_t)v;
        if (p < 0) p = (int16_t)-p;
        outlier[(size_t)(by * 4 + f / 4) * W + bx * 4 + (f & 3)] = (int16_t)(p / 100);
      }
      obf[(size_t)by * (W / 4) + bx] = (int16_t)cnt;
    }
  return 0;
}


// the kernels' packed-table store (rmd_chunk.cuh store_packed_depth) for one CTU: cost = uint32[341][35] in, packed table out,
// written depth by depth by `nthreads` emulated threads exactly as the CTAs do
int emul_store_packed_ctu(const uint32_t* cost, int nthreads, uint8_t* packed) {
  for (int tid = 0; tid < nthreads; tid++) {
    store_packed_depth<6>(packed, tid, nthreads, [&](int i) { return cost[(size_t)pu_offset_of_depth(0) * kNumModes + i]; });
    store_packed_depth<5>(packed, tid, nthreads, [&](int i) { return cost[(size_t)pu_offset_of_depth(1) * kNumModes + i]; });
    store_packed_depth<4>(packed, tid, nthreads, [&](int i) { return cost[(size_t)pu_offset_of_depth(2) * kNumModes + i]; });
    store_packed_depth<3>(packed, tid, nthreads, [&](int i) { return cost[(size_t)pu_offset_of_depth(3) * kNumModes + i]; });
    store_packed_depth<2>(packed, tid, nthreads, [&](int i) { return cost[(size_t)pu_offset_of_depth(4) * kNumModes + i]; });
  }
  return kPackedCtuBytes;
}


// the kernel-side fork-aware enumeration (feature_core.cuh prune_mask_ctu) for every CTU of a picture; numObf[d] = per-depth grids of whole CUs
int emul_prune_mask(int W, int H, const int32_t* n0, const int32_t* n1, const int32_t* n2, const int32_t* n3, const uint8_t* swSkip, const uint8_t* swTerm, uint8_t* needed) {
  const int32_t* grids[4] = {n0, n1, n2, n3};
  const int wc = (W + 63) / 64, hc = (H + 63) / 64;
  for (int ctu = 0; ctu < wc * hc; ctu++) {
    auto numObf = [&](int d, int cx, int cy) { return grids[d][(size_t)cy * (W / (64 >> d)) + cx]; };
    prune_mask_ctu((ctu % wc) * 64, (ctu / wc) * 64, W, H, swSkip, swTerm, numObf, needed + (size_t)ctu * 341);
  }
  return 0;
}

}  // extern "C"
