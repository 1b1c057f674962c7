/* hmref_driver.cpp - C driver around the REFERENCE's own hot-path functions (test infrastructure).
 *
 * Compiled by oracle/build_ref.sh into oracle/_ref/libhmref.so together with the patched copy of
 * the reference's TLibCommon / TLibEncoder objects.  It exists so that
 *   (1) oracle/cucd_oracle.c (the plain-C restatement) can be cross-checked, on arbitrary seeded
 *       inputs, against the functions the reference encoder really executes, and
 *   (2) bench.py can time the reference's CPU implementation of the path (`--impl reference`,
 *       `cpu_baseline.kind = "reference"`).
 * It is never linked or loaded by the product (libcucudecide.so).
 *
 * What is the reference's own compiled code here:
 *   fillReferenceSamples            TComPattern.cpp:314-521      (free function, called as is)
 *   TComPrediction::xPredIntraAng   TComPrediction.cpp:250-410   (via a derived class)
 *   TComPrediction::xPredIntraPlanar                :755-805
 *   TComPrediction::xDCPredFiltering                :818-841
 *   TComPrediction::filteringIntraReferenceSamples  TComPattern.cpp:523-548
 *   TComRdCost::setDistParam + DistFunc (xGetHADs / xGetSAD*)  TComRdCost.cpp:306-431,465-1604
 *   TEncSlice::getOutlierWithDCT (+ partialButterfly, TCMprocessOneSequence) TEncSlice.cpp:55-392,878-1173
 *   xCalcHADs8x8_ISlice             TEncCu.cpp:1780-1872
 *   xTrMxN / xITrMxN (partial butterflies, DST)   TComTrQuant.cpp:388-985   (free functions, called as is)
 *   TComRdCost::getDistPart -> xGetSSE*           TComRdCost.cpp:433-455, 970-1315
 *   getTMVFeature (+ T3x3Filter, TMVFeature)  tools_YS.cpp:1659-1839   (on a TComDataCU shell carrying position, size, picture)
 *   TEncPreanalyzer::xPreanalyze    TEncPreanalyzer.cpp:64-139         (on a TEncPic shell carrying the source plane + AQ layers)
 * What the driver has to restate because the reference only has it inline in functions that need a
 * live TComDataCU/TComTU/TComPic (all three restatements are pinned by the encoder KAT dumps in
 * tests/golden/, which come from the reference's real call sites):
 *   - the [1 2 1] / strong reference-sample smoothing of initAdiPatternChType  TComPattern.cpp:185-283
 *   - neighbour availability in frame (replay) mode = HEVC z-scan order rule   TComPattern.cpp:550-727
 *   - the mode loop glue of estIntraPredLumaQT                                 TEncSearch.cpp:2327-2361
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <stdint.h>
#include <vector>
#include <thread>
#include <atomic>
#include <algorithm>
#include <iostream>
#include <fstream>
#include <sstream>
#include <string>
#include <list>
#include <map>

#define private public
#define protected public
#include "TLibCommon/TComRom.h"
#include "TLibCommon/TComPic.h"
#include "TLibCommon/TComPicYuv.h"
#include "TLibCommon/TComPattern.h"
#include "TLibCommon/TComPrediction.h"
#include "TLibCommon/TComRdCost.h"
#include "TLibEncoder/TEncSlice.h"
#include "TLibEncoder/TEncPic.h"
#include "TLibEncoder/TEncPreanalyzer.h"
#include "TLibCommon/TComDataCU.h"
#include "TLibCommon/tools_YS.h"
#undef private
#undef protected

Void fillReferenceSamples(const Int bitDepth, TComDataCU* pcCU, const Pel* piRoiOrigin, Pel* piAdiTemp, const Bool* bNeighborFlags,
                          const Int iNumIntraNeighbor, const Int unitWidth, const Int unitHeight, const Int iAboveUnits, const Int iLeftUnits,
                          const UInt uiCuWidth, const UInt uiCuHeight, const UInt uiWidth, const UInt uiHeight, const Int iPicStride,
                          const ChannelType chType, const ChromaFormat chFmt);
Int xCalcHADs8x8_ISlice(Pel* piOrg, Int iStrideOrg);
Void xTrMxN(Int bitDepth, TCoeff* block, TCoeff* coeff, Int iWidth, Int iHeight, Bool useDST, const Int maxTrDynamicRange);
Void xITrMxN(Int bitDepth, TCoeff* coeff, TCoeff* block, Int iWidth, Int iHeight, Bool useDST, const Int maxTrDynamicRange);

namespace {

bool g_inited = false;

void set_globals(int bitDepth) {
  if (!g_inited) {
    g_uiMaxCUWidth = 64; g_uiMaxCUHeight = 64; g_uiMaxCUDepth = 4; g_uiAddCUDepth = 1;
    initROM();
    g_inited = true;
  }
  for (int c = 0; c < MAX_NUM_CHANNEL_TYPE; c++) {
    g_bitDepth[c] = bitDepth;
#if O0043_BEST_EFFORT_DECODING
    g_bitDepthInStream[c] = bitDepth;
#endif
    g_maxTrDynamicRange[c] = 15; /* extended_precision_processing off: TAppEncCfg.cpp */
  }
}

/* Restatement of TComPattern.cpp:185-283 on the L-shaped (2N+1)-stride arrays. */
void filter_border(const Pel* unf, Pel* fil, int n, int bitDepth, bool strongEnabled) {
  const int n2 = 2 * n, stride = n2 + 1;
  const Pel* src = unf + stride * n2;
  Pel* dst = fil + stride * n2;
  bool strong = strongEnabled;
  const Pel bottomLeft = unf[stride * n2], topLeft = unf[0], topRight = unf[n2];
  if (strong) {
    const int thr = 1 << (bitDepth - 5);
    const bool bl = abs((bottomLeft + topLeft) - 2 * unf[stride * n]) < thr;
    const bool ba = abs((topLeft + topRight) - 2 * unf[n]) < thr;
    if (n < 32 || !bl || !ba) strong = false;
  }
  *dst = *src; dst -= stride; src -= stride;
  if (strong) {
    const int shift = g_aucConvertToBit[n] + 3;
    for (int i = 1; i < n2; i++, dst -= stride) *dst = ((n2 - i) * bottomLeft + i * topLeft + n) >> shift;
    src -= stride * (n2 - 1);
  } else {
    for (int i = 1; i < n2; i++, dst -= stride, src -= stride) *dst = (src[stride] + 2 * src[0] + src[-stride] + 2) >> 2;
  }
  if (strong) *dst = src[0]; else *dst = (src[stride] + 2 * src[0] + src[1] + 2) >> 2;
  dst++; src++;
  if (strong) {
    const int shift = g_aucConvertToBit[n] + 3;
    for (int i = 1; i < n2; i++, dst++) *dst = ((n2 - i) * topLeft + i * topRight + n) >> shift;
    src += n2 - 1;
  } else {
    for (int i = 1; i < n2; i++, dst++, src++) *dst = (src[1] + 2 * src[0] + src[-1] + 2) >> 2;
  }
  *dst = *src;
}

struct RefPred : public TComPrediction {
  /* glue of TEncSearch.cpp:2327-2339 + TComPrediction::predIntraAng :412-496 for luma */
  void predict(int bitDepth, const Pel* unf, const Pel* fil, int n, int mode, Pel* pred, int predStride) {
    const int sw = 2 * n + 1;
    const bool useFilter = TComPrediction::filteringIntraReferenceSamples(COMPONENT_Y, mode, n, n, CHROMA_420, false);
    const Pel* ptrSrc = useFilter ? fil : unf;
    if (mode == PLANAR_IDX) {
      xPredIntraPlanar(ptrSrc + sw + 1, sw, pred, predStride, n, n, CHANNEL_TYPE_LUMA, CHROMA_420);
    } else {
      xPredIntraAng(bitDepth, ptrSrc + sw + 1, sw, pred, predStride, n, n, CHANNEL_TYPE_LUMA, CHROMA_420, mode, true, true, true);
      if (mode == DC_IDX) xDCPredFiltering(ptrSrc + sw + 1, sw, pred, predStride, n, n, CHANNEL_TYPE_LUMA);
    }
  }
};

struct Worker {
  RefPred pred;
  TComRdCost rd;
  std::vector<Pel> unf, fil, predBuf, orgBuf;
  Worker() : unf(129 * 129), fil(129 * 129), predBuf(64 * 64), orgBuf(64 * 64) {}
  /* one PU: L-shaped unfiltered border already in unf */
  void rmd(int bitDepth, int n, int strong, const Pel* org, int orgStride, uint32_t* sad) {
    filter_border(unf.data(), fil.data(), n, bitDepth, strong != 0);
    for (int mode = 0; mode < 35; mode++) {
      pred.predict(bitDepth, unf.data(), fil.data(), n, mode, predBuf.data(), n);
      DistParam dp;
      rd.setDistParam(dp, bitDepth, const_cast<Pel*>(org), orgStride, predBuf.data(), n, n, n, true);
      dp.bApplyWeight = false;
      sad[mode] = dp.DistFunc(&dp);
    }
  }
};

inline uint32_t morton(uint32_t x, uint32_t y) {
  uint32_t r = 0;
  for (int b = 0; b < 4; b++) r |= ((x >> b) & 1) << (2 * b) | ((y >> b) & 1) << (2 * b + 1);
  return r;
}

/* HEVC 6.4.1 z-scan availability for the 4x4 unit at (xn,yn) seen from a PU whose first unit is at (xc,yc). */
inline bool unit_available(int xc, int yc, int xn, int yn, int W, int H) {
  if (xn < 0 || yn < 0 || xn >= W || yn >= H) return false;
  const int wc = (W + 63) >> 6;
  const int ctuC = (yc >> 6) * wc + (xc >> 6), ctuN = (yn >> 6) * wc + (xn >> 6);
  if (ctuN != ctuC) return ctuN < ctuC;
  return morton((xn & 63) >> 2, (yn & 63) >> 2) < morton((xc & 63) >> 2, (yc & 63) >> 2);
}

}  // namespace

extern "C" {

int hmref_init(int bitDepth) { set_globals(bitDepth); return 0; }

/* linear border (bottom-left .. top-left .. top-right, 4N+1) -> 35 SATD costs */
int hmref_rmd_pu(int bitDepth, int n, int strongSmoothing, const int16_t* org, int orgStride, const int16_t* border, uint32_t* sad) {
  set_globals(bitDepth);
  static thread_local Worker* w = 0;
  if (!w) w = new Worker;
  const int sw = 2 * n + 1;
  for (int i = 0; i < 2 * n; i++) w->unf[(2 * n - i) * sw] = border[i];
  for (int i = 0; i < sw; i++) w->unf[i] = border[2 * n + i];
  w->rmd(bitDepth, n, strongSmoothing, org, orgStride, sad);
  return 0;
}

/* the reference's filtered border for a linear unfiltered border (restated filter; see header) */
int hmref_filter_border(int bitDepth, int n, int strongSmoothing, const int16_t* border, int16_t* filtered) {
  set_globals(bitDepth);
  const int sw = 2 * n + 1;
  std::vector<Pel> unf(sw * sw), fil(sw * sw);
  for (int i = 0; i < 2 * n; i++) unf[(2 * n - i) * sw] = border[i];
  for (int i = 0; i < sw; i++) unf[i] = border[2 * n + i];
  filter_border(unf.data(), fil.data(), n, bitDepth, strongSmoothing != 0);
  for (int i = 0; i < 2 * n; i++) filtered[i] = fil[(2 * n - i) * sw];
  for (int i = 0; i < sw; i++) filtered[2 * n + i] = fil[i];
  return 0;
}

/* one prediction block, for mode-by-mode checks of the restatement */
int hmref_predict(int bitDepth, int n, int mode, const int16_t* border, int useFiltered_unused, int strongSmoothing, int16_t* pred) {
  set_globals(bitDepth);
  const int sw = 2 * n + 1;
  std::vector<Pel> unf(sw * sw), fil(sw * sw);
  for (int i = 0; i < 2 * n; i++) unf[(2 * n - i) * sw] = border[i];
  for (int i = 0; i < sw; i++) unf[i] = border[2 * n + i];
  filter_border(unf.data(), fil.data(), n, bitDepth, strongSmoothing != 0);
  RefPred p;
  p.predict(bitDepth, unf.data(), fil.data(), n, mode, pred, n);
  (void)useFiltered_unused;
  return 0;
}

/* fillReferenceSamples as the reference runs it: recon plane + flags -> linear border */
int hmref_fill_border(int bitDepth, int n, const int16_t* recOrigin, int recStride, const uint8_t* flags, int16_t* border) {
  set_globals(bitDepth);
  const int units = n / 4, total = 4 * units + 1, sw = 2 * n + 1;
  Bool bf[4 * 32 + 1];
  int num = 0;
  for (int i = 0; i < total; i++) { bf[i] = flags[i] != 0; num += bf[i] ? 1 : 0; }
  std::vector<Pel> ext(sw * sw);
  fillReferenceSamples(bitDepth, 0, recOrigin, ext.data(), bf, num, 4, 4, 2 * units, 2 * units, n, n, sw, sw, recStride, CHANNEL_TYPE_LUMA, CHROMA_420);
  for (int i = 0; i < 2 * n; i++) border[i] = ext[(2 * n - i) * sw];
  for (int i = 0; i < sw; i++) border[2 * n + i] = ext[i];
  return 0;
}

uint32_t hmref_hads(int bitDepth, const int16_t* org, int orgStride, const int16_t* cur, int curStride, int w, int h) {
  set_globals(bitDepth);
  TComRdCost rd; DistParam dp;
  rd.setDistParam(dp, bitDepth, const_cast<Pel*>(org), orgStride, const_cast<Pel*>(cur), curStride, w, h, true);
  dp.bApplyWeight = false;
  return dp.DistFunc(&dp);
}

/* SAD exactly as xTZSearchHelp / xPatternSearch set it up (TEncSearch.cpp:336-360, 3886-3925) */
uint32_t hmref_sad(int bitDepth, const int16_t* org, int orgStride, const int16_t* ref, int refStride, int w, int h, int subShift) {
  set_globals(bitDepth);
  TComRdCost rd; DistParam dp; TComPattern pat;
  pat.initPattern(const_cast<Pel*>(org), w, h, orgStride);
  rd.setDistParam(&pat, const_cast<Pel*>(ref), refStride, dp);
  dp.iSubShift = subShift; dp.bitDepth = bitDepth; dp.bApplyWeight = false;
  return dp.DistFunc(&dp);
}

/* SAD surface over an integer window, raster order y-major: out[(dy-top)*(right-left+1) + (dx-left)] */
int hmref_sad_surface(int bitDepth, const int16_t* org, int orgStride, int w, int h, const int16_t* refAtZeroMv, int refStride,
                      int left, int right, int top, int bottom, int subShift, uint32_t* out) {
  set_globals(bitDepth);
  TComRdCost rd; DistParam dp; TComPattern pat;
  pat.initPattern(const_cast<Pel*>(org), w, h, orgStride);
  rd.setDistParam(&pat, const_cast<Pel*>(refAtZeroMv), refStride, dp);
  dp.iSubShift = subShift; dp.bitDepth = bitDepth; dp.bApplyWeight = false;
  const int cols = right - left + 1;
  for (int y = top; y <= bottom; y++)
    for (int x = left; x <= right; x++) {
      dp.pCur = const_cast<Pel*>(refAtZeroMv) + y * refStride + x;
      out[(y - top) * cols + (x - left)] = dp.DistFunc(&dp);
    }
  return 0;
}

int hmref_src_had8x8(const int16_t* org, int stride) { return xCalcHADs8x8_ISlice(const_cast<Pel*>(org), stride); }

/* The reference's whole per-picture feature pass: TEncSlice::getOutlierWithDCT on a TComPic that
 * carries only the three planes the function touches. obf: (W/4)x(H/4), outlier: WxH, tight. */
int hmref_outlier_frame(int bitDepth, const int16_t* org, int orgStride, int W, int H, int16_t* obf, int16_t* outlier) {
  set_globals(bitDepth);
  TComPicYuv yOrg, yOut, yObf;
  yOrg.create(W, H, CHROMA_420, 64, 64, 4);
  yOut.create(W, H, CHROMA_420, 64, 64, 4);
  yObf.create(W / 4, H / 4, CHROMA_420, 64, 64, 4);
  for (int y = 0; y < H; y++) memcpy(yOrg.getAddr(COMPONENT_Y) + y * yOrg.getStride(COMPONENT_Y), org + y * orgStride, W * sizeof(Pel));
  for (int y = 0; y < H; y++) memset(yOut.getAddr(COMPONENT_Y) + y * yOut.getStride(COMPONENT_Y), 0, W * sizeof(Pel));
  for (int y = 0; y < H / 4; y++) memset(yObf.getAddr(COMPONENT_Y) + y * yObf.getStride(COMPONENT_Y), 0, (W / 4) * sizeof(Pel));
  TComPic* pic = new TComPic;
  pic->m_apcPicYuv[TComPic::PIC_YUV_ORG] = &yOrg;
  pic->m_apcPicYuvOutlier = &yOut;
  pic->m_apcOBF = &yObf;
  TEncSlice* slice = new TEncSlice;
  slice->getOutlierWithDCT(pic);
  for (int y = 0; y < H; y++) memcpy(outlier + y * W, yOut.getAddr(COMPONENT_Y) + y * yOut.getStride(COMPONENT_Y), W * sizeof(Pel));
  for (int y = 0; y < H / 4; y++) memcpy(obf + y * (W / 4), yObf.getAddr(COMPONENT_Y) + y * yObf.getStride(COMPONENT_Y), (W / 4) * sizeof(Pel));
  pic->m_apcPicYuv[TComPic::PIC_YUV_ORG] = 0;
  pic->m_apcPicYuvOutlier = 0;
  pic->m_apcOBF = 0;
  /* the TComPic / TEncSlice shells are leaked on purpose: their destructors assume a full create() */
  yOrg.destroy(); yOut.destroy(); yObf.destroy();
  return 0;
}

/* Frame (replay) mode full enumeration: for every CTU in [ctuBegin,ctuEnd) and every PU
 * (depth-major, z-order inside a depth; 341 per CTU) that lies inside the picture, build the
 * border from `rec` with z-scan availability, run the 35-mode RMD. out[(ctu-ctuBegin)*341*35 ...];
 * PUs outside the picture get 0xFFFFFFFF. */
int hmref_rmd_frame(int bitDepth, int strongSmoothing, const int16_t* org, int orgStride, const int16_t* rec, int recStride,
                    int W, int H, int ctuBegin, int ctuEnd, int nthreads, uint32_t* out) {
  set_globals(bitDepth);
  const int wc = (W + 63) >> 6;
  std::atomic<int> next(ctuBegin);
  auto body = [&]() {
    Worker w;
    Bool bf[4 * 32 + 1];
    for (;;) {
      const int ctu = next.fetch_add(1);
      if (ctu >= ctuEnd) break;
      const int cx = (ctu % wc) * 64, cy = (ctu / wc) * 64;
      uint32_t* o = out + (size_t)(ctu - ctuBegin) * 341 * 35;
      int pu = 0;
      for (int d = 0; d < 5; d++) {
        const int n = 64 >> d, per = 1 << d, units = n / 4, sw = 2 * n + 1;
        for (int z = 0; z < per * per; z++, pu++) {
          int px = 0, py = 0;
          for (int b = 0; b < d; b++) { px |= ((z >> (2 * b)) & 1) << b; py |= ((z >> (2 * b + 1)) & 1) << b; }
          const int x0 = cx + px * n, y0 = cy + py * n;
          uint32_t* sad = o + pu * 35;
          if (x0 + n > W || y0 + n > H) { for (int m = 0; m < 35; m++) sad[m] = 0xFFFFFFFFu; continue; }
          int num = 0;
          for (int u = 0; u < 2 * units; u++) {      /* below-left .. left, bottom to top */
            bf[u] = unit_available(x0, y0, x0 - 1, y0 + (2 * units - 1 - u) * 4, W, H); num += bf[u];
          }
          bf[2 * units] = unit_available(x0, y0, x0 - 1, y0 - 1, W, H); num += bf[2 * units];
          for (int u = 0; u < 2 * units; u++) {      /* above .. above-right */
            bf[2 * units + 1 + u] = unit_available(x0, y0, x0 + u * 4, y0 - 1, W, H); num += bf[2 * units + 1 + u];
          }
          fillReferenceSamples(bitDepth, 0, rec + (size_t)y0 * recStride + x0, w.unf.data(), bf, num, 4, 4, 2 * units, 2 * units,
                               n, n, sw, sw, recStride, CHANNEL_TYPE_LUMA, CHROMA_420);
          w.rmd(bitDepth, n, strongSmoothing, org + (size_t)y0 * orgStride + x0, orgStride, sad);
        }
      }
    }
  };
  if (nthreads <= 1) { body(); return 0; }
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; t++) th.emplace_back(body);
  for (auto& t : th) t.join();
  return 0;
}

/* The reference's forward / inverse core transform of one n x n block (what TComTrQuant::xT / xIT call). */
int hmref_fwd_transform(int bitDepth, int n, int useDST, const int32_t* block, int32_t* coeff) {
  set_globals(bitDepth);
  std::vector<TCoeff> in(block, block + n * n);
  xTrMxN(bitDepth, in.data(), coeff, n, n, useDST != 0, 15);
  return 0;
}
int hmref_inv_transform(int bitDepth, int n, int useDST, const int32_t* coeff, int32_t* block) {
  set_globals(bitDepth);
  std::vector<TCoeff> in(coeff, coeff + n * n);
  xITrMxN(bitDepth, in.data(), block, n, n, useDST != 0, 15);
  return 0;
}
uint32_t hmref_sse(int bitDepth, const int16_t* org, int orgStride, const int16_t* cur, int curStride, int w, int h) {
  set_globals(bitDepth);
  static TComRdCost rd;          /* the constructor runs init() (TComRdCost.cpp:46-49) */
  return rd.getDistPart(bitDepth, const_cast<Pel*>(cur), curStride, const_cast<Pel*>(org), orgStride, w, h, COMPONENT_Y, DF_SSE);
}

/* The reference's getTMVFeature on the CU at (x,y) of size n inside a W x H source plane: feat[5][26]. */
int hmref_tmv_features(const int16_t* org, int orgStride, int W, int H, int x, int y, int n, double* feat) {
  set_globals(8);
  TComPicYuv yOrg;
  yOrg.create(W, H, CHROMA_420, 64, 64, 4);
  for (int r = 0; r < H; r++) memcpy(yOrg.getAddr(COMPONENT_Y) + r * yOrg.getStride(COMPONENT_Y), org + r * orgStride, W * sizeof(Pel));
  TComPic* pic = new TComPic;
  pic->m_apcPicYuv[TComPic::PIC_YUV_ORG] = &yOrg;
  TComDataCU* cu = new TComDataCU;     /* shell: only the members getTMVFeature reads */
  UChar size = (UChar)n;
  cu->m_pcPic = pic; cu->m_uiCUPelX = x; cu->m_uiCUPelY = y; cu->m_puhWidth = &size; cu->m_puhHeight = &size;
  TMVFeature* f = getTMVFeature(cu);
  memcpy(feat, f->m_adFeature, sizeof(double) * 5 * 26);
  delete f;
  cu->m_puhWidth = 0; cu->m_puhHeight = 0; cu->m_pcPic = 0;
  pic->m_apcPicYuv[TComPic::PIC_YUV_ORG] = 0;
  yOrg.destroy();   /* the TComPic / TComDataCU shells are leaked on purpose (destructors assume create()) */
  return 0;
}

/* The reference's TEncPreanalyzer::xPreanalyze with maxAQDepth layers (unit = 64 >> d); activity[d] = ceil(W/unit)*ceil(H/unit)
 * doubles in raster order (NULL = skip), avg[d] = the layer's average activity. */
int hmref_aq_activity(const int16_t* org, int orgStride, int W, int H, int maxAQDepth, double* const* activity, double* avg) {
  set_globals(8);
  TComPicYuv yOrg;
  yOrg.create(W, H, CHROMA_420, 64, 64, 4);
  for (int r = 0; r < H; r++) memcpy(yOrg.getAddr(COMPONENT_Y) + r * yOrg.getStride(COMPONENT_Y), org + r * orgStride, W * sizeof(Pel));
  TEncPic* pic = new TEncPic;
  pic->m_apcPicYuv[TComPic::PIC_YUV_ORG] = &yOrg;
  pic->m_uiMaxAQDepth = maxAQDepth;
  pic->m_acAQLayer = new TEncPicQPAdaptationLayer[maxAQDepth];          /* what TEncPic::create does, TEncPic.cpp:128-137 */
  for (int d = 0; d < maxAQDepth; d++) pic->m_acAQLayer[d].create(W, H, 64 >> d, 64 >> d);
  TEncPreanalyzer pre;
  pre.xPreanalyze(pic);
  for (int d = 0; d < maxAQDepth; d++) {
    TEncPicQPAdaptationLayer* l = pic->getAQLayer(d);
    const int n = l->getNumAQPartInWidth() * l->getNumAQPartInHeight();
    if (activity && activity[d]) for (int i = 0; i < n; i++) activity[d][i] = l->getQPAdaptationUnit()[i].getActivity();
    if (avg) avg[d] = l->getAvgActivity();
  }
  delete[] pic->m_acAQLayer; pic->m_acAQLayer = 0; pic->m_uiMaxAQDepth = 0;
  pic->m_apcPicYuv[TComPic::PIC_YUV_ORG] = 0;
  yOrg.destroy();
  return 0;
}

}  /* extern "C" */
