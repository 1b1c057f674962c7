/* Known-answer-test dump hooks for the patched copy of the reference (test infrastructure).
 *
 * oracle/build_ref.sh force-includes this header into the /tmp copy of the reference and inserts
 * one-line calls at the hot-path seams listed in SURVEY.md §8c.  Every hook is a no-op unless the
 * matching environment variable names an output file, so the instrumented TAppEncoder produces
 * the same bitstream as the un-instrumented one.  Records are little-endian int32 words followed
 * by int16 sample payloads; tests/golden/gen_golden.py is the only reader.
 *
 *   CUCD_DUMP_RMD  : one record per rough-mode-decision PU   (TEncSearch.cpp:2252-2361)
 *   CUCD_DUMP_ME   : every CUCD_DUMP_ME_EVERY-th integer-ME SAD probe (TEncSearch.cpp:336-437)
 *   CUCD_DUMP_OBF  : one record per picture of the outlier feature pass (TEncSlice.cpp:878-1173)
 *   CUCD_DUMP_CU   : one record per CU visited by xCompressCU (TEncCu.cpp:589-600)
 *   CUCD_DUMP_TU   : every CUCD_DUMP_TU_EVERY-th luma TU of xIntraCodingTUBlock (TEncSearch.cpp:1092-1387)
 */
#ifndef CUCD_DUMP_H
#define CUCD_DUMP_H
#ifdef __cplusplus
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdint.h>

struct CucdDump {
  FILE* f;
  bool  tried;
  CucdDump() : f(0), tried(false) {}
  FILE* get(const char* env) {
    if (!tried) { tried = true; const char* p = getenv(env); if (p && *p) f = fopen(p, "wb"); }
    return f;
  }
};
inline void cucd_w32(FILE* f, int32_t v) { fwrite(&v, 4, 1, f); }

/* ---- integration build (-DCUCD_INTEGRATION): the lines INTEGRATION.md asks a maintainer to add, so that the patched
 *      reference encoder computes its S1 / S2 numbers on the GPU through libcucudecide.so (oracle/_ref/TAppEncoderCucd).
 *      Used by tests/test_gpu_encoder_md5.py to show that the bitstream stays byte-identical. ------------------------- */
#ifdef CUCD_INTEGRATION
#include <vector>
#include "cucudecide.h"
#include "cucd_ipc.h"
struct CucdShim { cucd_handle* h; int W, H, bd, strong; uint32_t sad[35]; long rmdCalls, frameCalls, tmvCalls; bool curForTmv; bool ipcOpen; bool rmdActive; int minN; long rmdCpu; };
inline CucdShim& cucd_shim() { static CucdShim s = {0, 0, 0, 0, 0, {0}, 0, 0, 0, false, false, false, -1, 0}; return s; }
/* Server mode (CUCD_SERVER=<shared-memory name>): this encoder instance is one of several processes whose requests cucd_server
 * coalesces (include/cucd_ipc.h); the instance itself never touches CUDA.  Otherwise the library is called in-process. */
inline cucd_ipc_client& cucd_ipc() { static cucd_ipc_client c; return c; }
inline bool cucd_ipc_mode() {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("CUCD_SERVER");
    mode = (e && *e) ? 1 : 0;
    if (mode && cucd_ipc().connect(e) != CUCD_OK) { fprintf(stderr, "cucd shim: cannot reach cucd_server: %s\n", cucd_ipc().err); exit(1); }
  }
  return mode == 1;
}
inline bool cucd_shim_ready() { return cucd_shim().h != 0 || cucd_shim().ipcOpen; }
inline void cucd_shim_die(const char* what) {   /* HM convention: fatal error -> exit(1) (CommonDef.h:141-164) */
  fprintf(stderr, "cucd shim: %s failed: %s\n", what, cucd_ipc_mode() ? cucd_ipc().err : cucd_last_error(cucd_shim().h));
  exit(1);
}
/* S0: TEncTop::create would do this once; the shim opens lazily at the first picture */
inline void cucd_shim_open(int W, int H, int bd, int strong) {
  CucdShim& s = cucd_shim();
  if (cucd_ipc_mode()) {
    if (!s.ipcOpen || s.W != W || s.H != H || s.bd != bd || s.strong != strong) {
      if (cucd_ipc().open(W, H, bd, strong) != CUCD_OK) cucd_shim_die("cucd_server OPEN");
      s.W = W; s.H = H; s.bd = bd; s.strong = strong; s.ipcOpen = true;
    }
    return;
  }
  if (s.h && (s.W != W || s.H != H || s.bd != bd || s.strong != strong)) { cucd_destroy(s.h); s.h = 0; }
  if (!s.h) {
    cucd_config cfg = {W, H, bd, 64, 4, strong, 0, 1, 0};
    if (cucd_create(&cfg, &s.h) != CUCD_OK) cucd_shim_die("cucd_create");
    s.W = W; s.H = H; s.bd = bd; s.strong = strong;
  }
}
/* S1: replaces TEncSlice::getOutlierWithDCT(pcPic) at TEncGOP.cpp:1095-1096 */
inline void cucd_shim_outlier(int W, int H, int bd, int strong, const short* org, int orgStride, short* obf, int obfStride,
                              short* outl, int outlStride) {
  cucd_shim_open(W, H, bd, strong);
  CucdShim& s = cucd_shim();
  std::vector<int16_t> o((size_t)(W / 4) * (H / 4)), t((size_t)W * H);
  if (cucd_ipc_mode()) {
    if (cucd_ipc().frame(W, H, org, orgStride, o.data(), t.data()) != CUCD_OK) cucd_shim_die("cucd_server FRAME");
  } else {
    cucd_frame_out fo; memset(&fo, 0, sizeof fo);
    fo.obf = o.data(); fo.outlier = t.data();
    if (cuCUDecide_frame(s.h, org, orgStride, 0, 0, 0, &fo) != CUCD_OK) cucd_shim_die("cuCUDecide_frame");
    if (cucd_set_cur_picture(s.h, org, orgStride) != CUCD_OK) cucd_shim_die("cucd_set_cur_picture");   /* a12 check + S3 of this picture */
    s.curForTmv = true;
  }
  for (int r = 0; r < H / 4; r++) memcpy(obf + (size_t)r * obfStride, &o[(size_t)r * (W / 4)], (W / 4) * sizeof(short));
  for (int r = 0; r < H; r++) memcpy(outl + (size_t)r * outlStride, &t[(size_t)r * W], W * sizeof(short));
  s.frameCalls++;
}
/* S2: replaces predIntraAng + DistFunc of the 35-mode loop TEncSearch.cpp:2327-2361 by one batch call per PU */
/* Offload policy: CUCD_SHIM_MIN_N=<size> keeps PUs smaller than <size> on the encoder's own CPU loop.  One PU at a time is all the live
 * encoder can offer (its border is the reconstruction of the PU before it), so a request costs a host<->device round trip of tens of
 * microseconds - more than the CPU needs for the 35 modes of a 4x4 or 8x8 PU, much less than it needs for a 32x32 or 64x64 one
 * (profiles/r02_encoder_wallclock.md).  Default 0: everything goes to the GPU (what the byte-identity tests exercise). */
inline bool cucd_shim_rmd_active() { return cucd_shim().rmdActive; }
inline void cucd_shim_rmd(int n, const short* unfExt, const short* org, int orgStride) {
  CucdShim& s = cucd_shim();
  if (s.minN < 0) { const char* e = getenv("CUCD_SHIM_MIN_N"); s.minN = e ? atoi(e) : 0; }
  s.rmdActive = n >= s.minN;
  if (!s.rmdActive) { s.rmdCpu++; return; }
  int16_t border[4 * 64 + 1], blk[64 * 64];
  const int sw = 2 * n + 1;
  for (int i = 0; i < 2 * n; i++) border[i] = unfExt[(2 * n - i) * sw];
  for (int i = 0; i < sw; i++) border[2 * n + i] = unfExt[i];
  for (int r = 0; r < n; r++) memcpy(blk + r * n, org + (size_t)r * orgStride, n * sizeof(short));
  int lg = 0; while ((1 << lg) < n) lg++;
  cucd_pu_desc d = {(uint8_t)lg, {0, 0, 0}};
  if ((cucd_ipc_mode() ? cucd_ipc().intra_rmd_batch(1, &d, blk, border, s.sad) : cucd_intra_rmd_batch(s.h, 1, &d, blk, border, s.sad)) != CUCD_OK)
    cucd_shim_die("cucd_intra_rmd_batch");
  s.rmdCalls++;
}
inline unsigned cucd_shim_rmd_sad(int mode) { return cucd_shim().sad[mode]; }
/* S3: integer motion estimation (uni-directional search).  xMotionEstimation opens a PU/reference pair; every SAD the
 * search asks for (xTZSearchHelp TEncSearch.cpp:421, xPatternSearch :3924) is read from SAD surfaces computed by
 * cucd_me_sad_surface.  The TZ search re-centres its window on the best start point (TEncSearch.cpp:4034-4047), so the shim
 * asks for 128x128-candidate tiles of the legal MV area (TComDataCU::clipMv) on demand instead of one fixed window. */
#include <map>
struct CucdMeShim {
  bool active; int curPoc, nRefs; const void* refKey[64];
  cucd_me_desc d; int minX, maxX, minY, maxY;
  std::map<long, std::vector<uint32_t> > tiles;
  long pus, tilesComputed, probes, biPus;
  bool ownKey; std::vector<short> key;       /* bi-predictive search: the pattern key 2 * org - other prediction (w x h, packed) */
  CucdMeShim() : active(false), curPoc(-1000000), nRefs(0), pus(0), tilesComputed(0), probes(0), biPus(0), ownKey(false) {}
};
inline CucdMeShim& cucd_me_shim() { static CucdMeShim s; return s; }
inline bool cucd_shim_me_active() { return cucd_me_shim().active; }
/* the current picture (once per POC) and the reference plane (once per picture object) become resident; returns the reference's slot */
inline int cucd_shim_me_pictures(int W, int H, int bd, int strong, int curPoc, const short* orgY, int orgStride, const void* refKey, const short* refY,
                                 int refStride, int marginX, int marginY) {
  cucd_shim_open(W, H, bd, strong);
  CucdShim& s = cucd_shim(); CucdMeShim& m = cucd_me_shim();
  if (curPoc != m.curPoc) {
    if ((cucd_ipc_mode() ? cucd_ipc().set_cur_picture(W, H, orgY, orgStride) : cucd_set_cur_picture(s.h, orgY, orgStride)) != CUCD_OK) cucd_shim_die("cucd_set_cur_picture");
    m.curPoc = curPoc; m.nRefs = 0;
  }
  int slot = -1;
  for (int i = 0; i < m.nRefs; i++) if (m.refKey[i] == refKey) slot = i;
  if (slot < 0) {
    slot = m.nRefs++; m.refKey[slot] = refKey;
    if ((cucd_ipc_mode() ? cucd_ipc().set_ref_picture(W, H, slot, refY, refStride, marginX, marginY) : cucd_set_ref_picture(s.h, slot, refY, refStride, marginX, marginY)) != CUCD_OK)
      cucd_shim_die("cucd_set_ref_picture");
  }
  return slot;
}
/* Distortion between the source PU and its uni-directional motion-compensated prediction at quarter-pel MV (mvxQ, mvyQ): what
 * xGetTemplateCost (AMVP candidate check, TEncSearch.cpp:3719-3760: xPredInterBlk + getDistPart DF_SAD) and xGetInterPredictionError
 * (merge candidates, :2905-2926: motionCompensation + SAD / Hadamard) compute.  HM's motion compensation of a PU is the separable
 * 8-tap interpolation at (mv & 3) around the integer part (mv >> 2): one entry of cucd_me_subpel_cost's 49-point table. */
struct CucdMcShim { long amvp, merge; CucdMcShim() : amvp(0), merge(0) {} };
inline CucdMcShim& cucd_mc_shim() { static CucdMcShim s; return s; }
inline bool cucd_shim_mc_enabled() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("CUCD_SHIM_MC"); on = (e && *e == '0') ? 0 : 1; }
  return on != 0;
}
inline unsigned cucd_shim_mc_dist(int W, int H, int bd, int strong, int curPoc, const short* orgY, int orgStride, const void* refKey, const short* refY,
                                  int refStride, int marginX, int marginY, int puX, int puY, int w, int h, int mvxQ, int mvyQ, int useHadamard, int isMerge) {
  const int slot = cucd_shim_me_pictures(W, H, bd, strong, curPoc, orgY, orgStride, refKey, refY, refStride, marginX, marginY);
  cucd_subpel_desc d = {puX, puY, w, h, slot, mvxQ >> 2, mvyQ >> 2, useHadamard};
  uint32_t cost[49];
  if ((cucd_ipc_mode() ? cucd_ipc().me_subpel_cost(1, &d, cost) : cucd_me_subpel_cost(cucd_shim().h, 1, &d, cost)) != CUCD_OK) cucd_shim_die("cucd_me_subpel_cost (MC distortion)");
  if (isMerge) cucd_mc_shim().merge++; else cucd_mc_shim().amvp++;
  return cost[((mvyQ & 3) + 3) * 7 + (mvxQ & 3) + 3];
}
/* keyBlock != 0: bi-predictive refinement (if (bBi), TEncSearch.cpp:3787-3797): the search key is the caller's block, not the picture's */
inline void cucd_shim_me_begin(int W, int H, int bd, int strong, int curPoc, const short* orgY, int orgStride, const void* refKey, const short* refY,
                               int refStride, int marginX, int marginY, int cuX, int cuY, int puX, int puY, int w, int h, int subShift,
                               const short* keyBlock = 0, int keyStride = 0) {
  CucdMeShim& m = cucd_me_shim();
  m.ownKey = keyBlock != 0;
  if (m.ownKey) {
    m.key.resize((size_t)w * h);
    for (int y = 0; y < h; y++) memcpy(&m.key[(size_t)y * w], keyBlock + (size_t)y * keyStride, (size_t)w * sizeof(short));
    m.biPus++;
  }
  const int slot = cucd_shim_me_pictures(W, H, bd, strong, curPoc, orgY, orgStride, refKey, refY, refStride, marginX, marginY);
  m.d.x = puX; m.d.y = puY; m.d.w = w; m.d.h = h; m.d.ref_idx = slot; m.d.sub_shift = subShift;
  m.minX = -(64 + 8 + cuX - 1); m.maxX = W + 8 - cuX - 1;          /* TComDataCU::clipMv, TComDataCU.cpp:2946-2958, in integer pels */
  m.minY = -(64 + 8 + cuY - 1); m.maxY = H + 8 - cuY - 1;
  m.tiles.clear(); m.active = true; m.pus++;
}
inline void cucd_shim_me_end() { cucd_me_shim().active = false; }
inline unsigned cucd_shim_me_sad(int x, int y) {
  CucdShim& s = cucd_shim(); CucdMeShim& m = cucd_me_shim();
  if (x < m.minX || x > m.maxX || y < m.minY || y > m.maxY) { fprintf(stderr, "cucd shim: ME probe (%d,%d) outside the legal MV area\n", x, y); exit(1); }
  /* tiles of 128 x 128 candidates for the TZ search, 32 x 32 for the bi-predictive refinement (+-4 around its start, xPatternSearch) */
  const int sh = m.ownKey ? 5 : 7, span = (1 << sh) - 1;
  const int tx = (x + 4096) >> sh, ty = (y + 4096) >> sh;
  const long key = (long)ty * 4096 + tx;
  cucd_me_desc d = m.d;
  d.left = (tx << sh) - 4096; d.right = d.left + span; d.top = (ty << sh) - 4096; d.bottom = d.top + span;
  if (d.left < m.minX) d.left = m.minX; if (d.right > m.maxX) d.right = m.maxX;
  if (d.top < m.minY) d.top = m.minY; if (d.bottom > m.maxY) d.bottom = m.maxY;
  std::map<long, std::vector<uint32_t> >::iterator it = m.tiles.find(key);
  if (it == m.tiles.end()) {
    std::vector<uint32_t> surf((size_t)(d.right - d.left + 1) * (d.bottom - d.top + 1));
    if (m.ownKey) { if (cucd_me_sad_surface_src(s.h, 1, &d, &m.key[0], surf.data()) != CUCD_OK) cucd_shim_die("cucd_me_sad_surface_src"); }
    else if ((cucd_ipc_mode() ? cucd_ipc().me_sad_surface(1, &d, surf.data()) : cucd_me_sad_surface(s.h, 1, &d, surf.data())) != CUCD_OK) cucd_shim_die("cucd_me_sad_surface");
    it = m.tiles.insert(std::make_pair(key, std::vector<uint32_t>())).first;
    it->second.swap(surf);
    m.tilesComputed++;
  }
  m.probes++;
  return it->second[(size_t)(y - d.top) * (d.right - d.left + 1) + (x - d.left)];
}
/* 8f.3: fractional-pel refinement of the uni-directional searches: xPatternSearchFracDIF asks once for the 49 quarter-pel
 * positions around the integer MV, xPatternRefinement (TEncSearch.cpp:851) reads its 2 x 9 candidates from that table */
struct CucdFracShim { bool active; uint32_t cost[49]; long calls; CucdFracShim() : active(false), calls(0) {} };
inline CucdFracShim& cucd_frac_shim() { static CucdFracShim s; return s; }
inline bool cucd_shim_frac_active() { return cucd_frac_shim().active; }
inline void cucd_shim_frac_begin(int biPred, int mvx, int mvy, int useHadamard) {
  CucdFracShim& f = cucd_frac_shim(); CucdMeShim& m = cucd_me_shim();
  f.active = false;
  if (!cucd_shim_ready() || m.pus == 0) return;                    /* the PU / reference of the integer search that just ended */
  if (biPred && !m.ownKey) return;                                 /* (server mode keeps the bi-predictive refinement on the CPU) */
  cucd_subpel_desc d = {m.d.x, m.d.y, m.d.w, m.d.h, m.d.ref_idx, mvx, mvy, useHadamard};
  if (m.ownKey) { if (cucd_me_subpel_cost_src(cucd_shim().h, 1, &d, &m.key[0], f.cost) != CUCD_OK) cucd_shim_die("cucd_me_subpel_cost_src"); }
  else if ((cucd_ipc_mode() ? cucd_ipc().me_subpel_cost(1, &d, f.cost) : cucd_me_subpel_cost(cucd_shim().h, 1, &d, f.cost)) != CUCD_OK) cucd_shim_die("cucd_me_subpel_cost");
  f.active = true; f.calls++;
}
inline void cucd_shim_frac_end() { cucd_frac_shim().active = false; }
inline unsigned cucd_shim_frac_cost(int horVal, int verVal) { return cucd_frac_shim().cost[(verVal + 3) * 7 + horVal + 3]; }
/* 8f.2 / 8f.4: intra TU coding around the host's quantiser (CUCD_SHIM_TU=1).  Per TU of xIntraCodingTUBlock (TEncSearch.cpp:1092-1387):
 *   cucd_intra_tu_forward  -> the prediction REPLACES piPred; the transform output is compared with m_plTempCoeff after transformNxN
 *   (host: RDOQ / xQuant as shipped)
 *   cucd_intra_tu_recon    -> the reconstruction REPLACES piReco / the picture samples; the SSE is compared with getDistPart
 * Any difference is fatal, and a difference in the substituted samples would change the bitstream (tests/test_gpu_encoder_md5.py). */
struct CucdTuShim {
  int enabled; bool live; cucd_tu_desc d; int n;
  int16_t org[32 * 32], border[4 * 32 + 1], pix[32 * 32];
  int32_t coef[32 * 32];
  uint32_t dist; long tus;
  CucdTuShim() : enabled(-1), live(false), n(0), dist(0), tus(0) {}
};
inline CucdTuShim& cucd_tu_shim() { static CucdTuShim s; return s; }
inline bool cucd_shim_tu_enabled() {
  CucdTuShim& t = cucd_tu_shim();
  if (t.enabled < 0) { const char* e = getenv("CUCD_SHIM_TU"); t.enabled = (e && *e && *e != '0') ? 1 : 0; }
  return t.enabled == 1 && cucd_shim().h != 0;
}
inline void cucd_shim_tu_forward(int compID, int n, int mode, int qp, int transformSkip, const short* unfExt, const short* org, short* pred, int stride) {
  CucdTuShim& t = cucd_tu_shim();
  t.live = false;
  if (!cucd_shim_tu_enabled() || n > 32) return;
  int lg = 0; while ((1 << lg) < n) lg++;
  t.d.log2_size = (uint8_t)lg; t.d.mode = (uint8_t)mode; t.d.qp = (int8_t)qp;
  t.d.flags = (uint8_t)((transformSkip ? CUCD_TU_TRANSFORM_SKIP : 0) | (compID ? CUCD_TU_CHROMA : 0));
  t.n = n;
  const int sw = 2 * n + 1;
  for (int i = 0; i < 2 * n; i++) t.border[i] = unfExt[(2 * n - i) * sw];
  for (int i = 0; i < sw; i++) t.border[2 * n + i] = unfExt[i];
  for (int r = 0; r < n; r++) memcpy(t.org + r * n, org + (size_t)r * stride, n * sizeof(short));
  if (cucd_intra_tu_forward(cucd_shim().h, 1, &t.d, t.org, t.border, t.coef, t.pix) != CUCD_OK) cucd_shim_die("cucd_intra_tu_forward");
  for (int r = 0; r < n; r++) memcpy(pred + (size_t)r * stride, t.pix + r * n, n * sizeof(short));
  t.live = true; t.tus++;
}
inline void cucd_shim_tu_after_quant(const int* cpuCoef, const int* level) {
  CucdTuShim& t = cucd_tu_shim();
  if (!t.live) return;
  if (memcmp(cpuCoef, t.coef, sizeof(int32_t) * t.n * t.n) != 0) { fprintf(stderr, "cucd shim: transform output of a %dx%d TU differs from the reference\n", t.n, t.n); exit(1); }
  if (cucd_intra_tu_recon(cucd_shim().h, 1, &t.d, t.org, t.border, level, t.pix, &t.dist) != CUCD_OK) cucd_shim_die("cucd_intra_tu_recon");
}
inline void cucd_shim_tu_reco(short* reco, int stride, short* recQt, int recQtStride, short* recPic, int recPicStride, unsigned cpuSse) {
  CucdTuShim& t = cucd_tu_shim();
  if (!t.live) return;
  t.live = false;
  if (cpuSse != t.dist) { fprintf(stderr, "cucd shim: SSE of a %dx%d TU differs from the reference (%u vs %u)\n", t.n, t.n, t.dist, cpuSse); exit(1); }
  for (int r = 0; r < t.n; r++) {
    memcpy(reco + (size_t)r * stride, t.pix + r * t.n, t.n * sizeof(short));
    memcpy(recQt + (size_t)r * recQtStride, t.pix + r * t.n, t.n * sizeof(short));
    memcpy(recPic + (size_t)r * recPicStride, t.pix + r * t.n, t.n * sizeof(short));
  }
}
/* a12: getTMVFeature(rpcBestCU) (TEncCu.cpp:1561) -> cucd_tmv_features; compared bit for bit with the reference's doubles, which
 * then go to the insight file as before.  The S1 shim uploads the picture being encoded. */
inline void cucd_shim_tmv_check(int x, int y, int size, const double* ref130) {
  CucdShim& s = cucd_shim();
  if (!s.h || !s.curForTmv) return;
  static int enabled = -1;          /* CUCD_SHIM_TMV=0: no verification calls (wall-clock measurements) */
  if (enabled < 0) { const char* e = getenv("CUCD_SHIM_TMV"); enabled = (e && *e == '0') ? 0 : 1; }
  if (!enabled) return;
  int lg = 0; while ((1 << lg) < size) lg++;
  cucd_cu_desc cu = {x, y, lg};
  double got[CUCD_TMV_FEATURES];
  if (cucd_tmv_features(s.h, 1, &cu, got) != CUCD_OK) cucd_shim_die("cucd_tmv_features");
  if (memcmp(got, ref130, sizeof got) != 0) { fprintf(stderr, "cucd shim: TMV features of the %dx%d CU at (%d,%d) differ from the reference\n", size, size, x, y); exit(1); }
  s.tmvCalls++;
}
struct CucdShimReport { ~CucdShimReport() { CucdShim& s = cucd_shim();
  if (s.rmdCpu) fprintf(stderr, "cucd shim: %ld RMD PUs smaller than %d kept on the CPU (CUCD_SHIM_MIN_N)\n", s.rmdCpu, s.minN);
  if (s.ipcOpen) { fprintf(stderr, "cucd shim: %ld pictures, %ld RMD PUs on the GPU, %ld ME searches (%ld SAD tiles, %ld probes) on the GPU, %ld sub-pel refinements on the GPU, through cucd_server\n", s.frameCalls, s.rmdCalls, cucd_me_shim().pus, cucd_me_shim().tilesComputed, cucd_me_shim().probes, cucd_frac_shim().calls); cucd_ipc().close_client(); }
  if (s.h) { fprintf(stderr, "cucd shim: %ld pictures, %ld RMD PUs on the GPU, %ld ME searches (%ld SAD tiles, %ld probes) on the GPU, %ld of them bi-predictive, %ld sub-pel refinements on the GPU, %ld AMVP + %ld merge candidate distortions on the GPU, %ld TUs coded on the GPU, %ld TMV feature sets verified, %lld kernel launches\n", s.frameCalls, s.rmdCalls, cucd_me_shim().pus, cucd_me_shim().tilesComputed, cucd_me_shim().probes, cucd_me_shim().biPus, cucd_frac_shim().calls, cucd_mc_shim().amvp, cucd_mc_shim().merge, cucd_tu_shim().tus, s.tmvCalls, cucd_launch_count(s.h)); cucd_destroy(s.h); s.h = 0; } } };
static CucdShimReport cucd_shim_report_at_exit;
#endif

/* ---- RMD ------------------------------------------------------------------------------- */
struct CucdRmdState {
  CucdDump out;
  unsigned char flags[4 * 32 + 1];
  int nflags, nintra;
  uint32_t sad[35];
  int32_t hdr[8];
  const short *unf, *fil, *org; int orgStride;
};
inline CucdRmdState& cucd_rmd() { static CucdRmdState s; return s; }

/* TComPattern.cpp:134-139 — neighbour availability per 4-sample unit, luma only */
inline void cucd_hook_flags(int isLuma, const bool* flags, int n, int numIntra) {
  CucdRmdState& s = cucd_rmd();
  if (!isLuma) return;
  s.nflags = n; s.nintra = numIntra;
  for (int i = 0; i < n; i++) s.flags[i] = flags[i] ? 1 : 0;
}
/* TEncSearch.cpp:2311-2327 — before the 35-mode loop */
inline void cucd_hook_rmd_begin(int poc, int x, int y, int n, int bitDepth, const short* unf, const short* fil,
                                const short* org, int orgStride) {
  CucdRmdState& s = cucd_rmd();
  s.hdr[0] = 0x444d5243; s.hdr[1] = poc; s.hdr[2] = x; s.hdr[3] = y; s.hdr[4] = n; s.hdr[5] = bitDepth;
  s.unf = unf; s.fil = fil; s.org = org; s.orgStride = orgStride;
#ifdef CUCD_INTEGRATION
  cucd_shim_rmd(n, unf, org, orgStride);
#endif
}
inline void cucd_hook_rmd_mode(int mode, unsigned sad) { cucd_rmd().sad[mode] = sad; }
/* L-shaped (2N+1)-stride array -> linear bottom-left .. top-left .. top-right, 4N+1 samples */
inline void cucd_write_border(FILE* f, const short* ext, int n) {
  const int sw = 2 * n + 1;
  for (int i = 2 * n; i >= 1; i--) fwrite(&ext[i * sw], 2, 1, f);
  fwrite(ext, 2, sw, f);
}
inline void cucd_hook_rmd_end() {
  CucdRmdState& s = cucd_rmd();
  FILE* f = s.out.get("CUCD_DUMP_RMD");
  if (!f) return;
  const int n = s.hdr[4];
  s.hdr[6] = s.nflags; s.hdr[7] = s.nintra;
  fwrite(s.hdr, 4, 8, f);
  fwrite(s.flags, 1, s.nflags, f);
  cucd_write_border(f, s.unf, n);
  cucd_write_border(f, s.fil, n);
  for (int r = 0; r < n; r++) fwrite(s.org + r * s.orgStride, 2, n, f);
  fwrite(s.sad, 4, 35, f);
}

/* ---- integer ME probe ------------------------------------------------------------------ */
#ifdef CUCD_INTEGRATION
#define cucd_hook_me(...) ((void)0)       /* its last argument evaluates the CPU SAD */
#else
inline void cucd_hook_me(const short* org, int orgStride, const short* ref, int refStride, int cols, int rows,
                         int subShift, int bitDepth, int mvx, int mvy, unsigned sad) {
  static CucdDump out; static long cnt = 0; static long every = -1;
  FILE* f = out.get("CUCD_DUMP_ME");
  if (!f) return;
  if (every < 0) { const char* e = getenv("CUCD_DUMP_ME_EVERY"); every = e ? atol(e) : 97; if (every < 1) every = 1; }
  if ((cnt++ % every) != 0) return;
  int32_t hdr[8] = {0x454d5243, cols, rows, subShift, bitDepth, mvx, mvy, (int32_t)sad};
  fwrite(hdr, 4, 8, f);
  for (int r = 0; r < rows; r++) fwrite(org + r * orgStride, 2, cols, f);
  for (int r = 0; r < rows; r++) fwrite(ref + r * refStride, 2, cols, f);
}
#endif

/* ---- OBF / outlier picture pass -------------------------------------------------------- */
inline void cucd_hook_obf_yc(const double* yc, int n) {
  static CucdDump out; FILE* f = out.get("CUCD_DUMP_OBF_YC");
  if (!f) return;
  cucd_w32(f, n); fwrite(yc, 8, n, f); fflush(f);
}
inline void cucd_hook_obf(int poc, int w, int h, int bitDepth, const short* org, int orgStride, const short* obf,
                          int obfStride, const short* outl, int outlStride) {
  static CucdDump out; FILE* f = out.get("CUCD_DUMP_OBF");
  if (!f) return;
  int32_t hdr[5] = {0x46424f43, poc, w, h, bitDepth};
  fwrite(hdr, 4, 5, f);
  for (int r = 0; r < h; r++) fwrite(org + r * orgStride, 2, w, f);
  for (int r = 0; r < h / 4; r++) fwrite(obf + r * obfStride, 2, w / 4, f);
  for (int r = 0; r < h; r++) fwrite(outl + r * outlStride, 2, w, f);
  fflush(f);
}

/* ---- per-CU block sums ----------------------------------------------------------------- */
inline void cucd_hook_cu(int poc, int depth, int x, int y, int size, int numObf, int nOutlier) {
  static CucdDump out; FILE* f = out.get("CUCD_DUMP_CU");
  if (!f) return;
  int32_t rec[7] = {poc, depth, x, y, size, numObf, nOutlier};
  fwrite(rec, 4, 7, f);
}

/* ---- the fork's per-depth decision switches as they stand while a picture is coded (g_bDecisionSwitch, set by SetDecisionSwitch
 *      tools_YS.cpp:1123-1154 after the verify picture): one record per POC: poc, skip2Nx2N[4], terminateCU[4] ---- */
inline void cucd_hook_switches(int poc, int mainModel, bool*** sw) {
  static CucdDump out; static int lastPoc = -1000000;
  FILE* f = out.get("CUCD_DUMP_SWITCHES");
  if (!f || poc == lastPoc || !sw) return;
  lastPoc = poc;
  int32_t rec[9]; rec[0] = poc;
  for (int d = 0; d < 4; d++) { rec[1 + d] = sw[d][mainModel][1] ? 1 : 0; rec[5 + d] = sw[d][mainModel][2] ? 1 : 0; }   /* DecisionType: Skip2Nx2N = 1, TerminateCU = 2 */
  fwrite(rec, 4, 9, f); fflush(f);
}

/* ---- intra luma TU coding: prediction -> residual -> transform/quant -> inverse -> recon -> SSE ---- */
struct CucdTuState {
  CucdDump out; long cnt, every; bool live;
  int32_t hdr[16];
  short border[4 * 64 + 1], pred[32 * 32], org[32 * 32];
  int32_t coef[32 * 32], level[32 * 32];
  CucdTuState() : cnt(0), every(-1), live(false) {}
};
inline CucdTuState& cucd_tu() { static CucdTuState s; return s; }
/* TEncSearch.cpp:1207 - after the prediction is in piPred, before the residual */
inline void cucd_hook_tu_pred(int compID, int poc, int x, int y, int n, int mode, int bitDepth, int transformSkip, int loadMode,
                              const short* unfExt, const short* pred, const short* org, int stride) {
  CucdTuState& s = cucd_tu();
  s.live = false;
  if (n > 32) return;
  FILE* f = s.out.get("CUCD_DUMP_TU");
  if (!f) return;
  if (s.every < 0) { const char* e = getenv("CUCD_DUMP_TU_EVERY"); s.every = e ? atol(e) : 53; if (s.every < 1) s.every = 1; }
  if ((s.cnt++ % s.every) != 0) return;
  s.live = true;
  s.hdr[0] = 0x55545243; s.hdr[1] = poc; s.hdr[2] = x; s.hdr[3] = y; s.hdr[4] = n; s.hdr[5] = mode; s.hdr[6] = bitDepth;
  s.hdr[7] = transformSkip; s.hdr[8] = loadMode; s.hdr[15] = compID;
  const int sw = 2 * n + 1;
  for (int i = 0; i < 2 * n; i++) s.border[i] = unfExt[(2 * n - i) * sw];
  for (int i = 0; i <= 2 * n; i++) s.border[2 * n + i] = unfExt[i];
  for (int r = 0; r < n; r++) { memcpy(s.pred + r * n, pred + r * stride, 2 * n); memcpy(s.org + r * n, org + r * stride, 2 * n); }
}
/* TEncSearch.cpp:1263-1268 - right after transformNxN: m_plTempCoeff still holds the transform output */
inline void cucd_hook_tu_coeff(int qp, int sliceIsIntra, int signHide, int rdoq, const int* tmpCoeff, const int* level, int absSum) {
  CucdTuState& s = cucd_tu();
  if (!s.live) return;
  const int n = s.hdr[4];
  s.hdr[9] = qp; s.hdr[10] = sliceIsIntra; s.hdr[11] = signHide; s.hdr[12] = rdoq; s.hdr[13] = absSum;
  memcpy(s.coef, tmpCoeff, 4 * n * n); memcpy(s.level, level, 4 * n * n);
}
/* TEncSearch.cpp:1385-1386 - reconstruction done, distortion of this TU */
inline void cucd_hook_tu_end(const short* reco, int stride, unsigned dist) {
  CucdTuState& s = cucd_tu();
  if (!s.live) return;
  s.live = false;
  FILE* f = s.out.get("CUCD_DUMP_TU");
  const int n = s.hdr[4];
  s.hdr[14] = (int32_t)dist;
  fwrite(s.hdr, 4, 16, f);
  fwrite(s.border, 2, 4 * n + 1, f);
  fwrite(s.org, 2, n * n, f); fwrite(s.pred, 2, n * n, f);
  fwrite(s.coef, 4, n * n, f); fwrite(s.level, 4, n * n, f);
  for (int r = 0; r < n; r++) fwrite(reco + r * stride, 2, n, f);
}

/* ---- fractional-pel ME refinement: one record per sampled candidate of xPatternRefinement (TEncSearch.cpp:808-865) ---- */
struct CucdFracState { const short* ref; int stride; CucdFracState() : ref(0), stride(0) {} };
inline CucdFracState& cucd_frac() { static CucdFracState s; return s; }
/* TEncSearch.cpp:4353-4358: the reference block at the integer MV */
inline void cucd_hook_frac_begin(const short* refAtIntMv, int stride) { cucd_frac().ref = refAtIntMv; cucd_frac().stride = stride; }
/* TEncSearch.cpp:849-851: (horVal, verVal) = the candidate's offset from the integer MV in quarter pels */
#ifdef CUCD_INTEGRATION
#define cucd_hook_frac_cand(...) ((void)0)     /* its last argument evaluates the CPU distortion */
#else
inline void cucd_hook_frac_cand(const short* org, int orgStride, int w, int h, int bitDepth, int hadamard, int horVal, int verVal, unsigned dist) {
  static CucdDump out; static long cnt = 0; static long every = -1;
  FILE* f = out.get("CUCD_DUMP_FRAC");
  if (!f || !cucd_frac().ref) return;
  if (every < 0) { const char* e = getenv("CUCD_DUMP_FRAC_EVERY"); every = e ? atol(e) : 211; if (every < 1) every = 1; }
  if ((cnt++ % every) != 0) return;
  int32_t hdr[8] = {0x52465243, w, h, bitDepth, hadamard, horVal, verVal, (int32_t)dist};
  fwrite(hdr, 4, 8, f);
  for (int r = 0; r < h; r++) fwrite(org + r * orgStride, 2, w, f);
  const short* ref = cucd_frac().ref; const int rs = cucd_frac().stride;
  for (int r = -4; r < h + 5; r++) fwrite(ref + r * rs - 4, 2, w + 9, f);      /* window rows/cols -4 .. size+4 */
}
#endif
#endif /* __cplusplus */
#endif
