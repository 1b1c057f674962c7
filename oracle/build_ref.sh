#!/usr/bin/env bash
# Build the REAL reference (Jiraiya812/Fast-CU-Decision-HEVC, HM-16.3 fork) for use as a checker.
#
# TEST INFRASTRUCTURE ONLY.  Nothing in the product path links, loads or executes what this builds.
#
# The reference is MSVC-only as shipped, so this script
#   1. copies /root/reference into a scratch directory under /tmp  (the reference is read-only and
#      its sources are never copied into this repository),
#   2. applies the mechanical g++ patch of SURVEY.md §8c (trailing `;` on fork macros, <String>,
#      OpenCV stub, two undeclared globals in a dead template),
#   3. inserts one-line, env-var-gated KAT dump hooks (oracle/ref_shims/cucd_dump.h) at the seams,
#   4. compiles with plain g++/gcc via a generated Makefile (the reference has no build system),
#   5. writes ONLY binaries to oracle/_ref/ :
#        TAppEncoder, TAppDecoder   whole patched encoder/decoder (KAT generation, MD5 round trip)
#        libhmref.so                the reference's own hot-path functions behind a C driver
#                                   (oracle/ref_shims/hmref_driver.cpp) - CPU baseline + cross-check
#
# oracle/_ref/ is git-ignored (it still travels to the GPU box with gpurun).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${CUCD_REFERENCE:-/root/reference}"
WORK="${CUCD_REF_WORK:-/tmp/cucd_ref_build}"
OUT="$HERE/_ref"
JOBS="${CUCD_JOBS:-$(nproc)}"

if [ ! -d "$REF/Lib/TLibCommon" ]; then
  echo "build_ref.sh: reference not present at $REF - keeping any prebuilt oracle/_ref as is" >&2
  exit 0
fi
mkdir -p "$WORK" "$OUT"
STAMP="$WORK/.stamp"
SIG="$(cat "$HERE/build_ref.sh" "$HERE"/ref_shims/* "$HERE/../include/cucudecide.h" | md5sum | cut -d' ' -f1)"
if [ -f "$STAMP" ] && [ "$(cat "$STAMP")" = "$SIG" ] && [ -x "$OUT/TAppEncoder" ] && [ -f "$OUT/libhmref.so" ] && [ -x "$OUT/TAppEncoderCucd" ]; then
  echo "build_ref.sh: oracle/_ref up to date"
  exit 0
fi

rm -rf "$WORK/src" "$WORK/obj"
cp -r "$REF" "$WORK/src"
chmod -R u+w "$WORK/src"
SRC="$WORK/src"

# ---- 2. g++ patch ------------------------------------------------
sed -i -E '55,131s/^(#define[[:space:]]+[A-Za-z_0-9]+[[:space:]]+[0-9]+)[[:space:]]*;/\1/; 131s/#endif;/#endif/' "$SRC/Lib/TLibCommon/TypeDef.h"
sed -i 's/#include <String>/#include <string>/' "$SRC/Lib/TLibCommon/globals_YS.h" "$SRC/Lib/TLibCommon/tools_YS.h"
cp "$HERE/ref_shims/cvheaders.h" "$SRC/Lib/TLibCommon/cvheaders.h"
sed -i -E 's/g_dJdx(01|10)\[uiDepth\]\[x_OBF\]/0/' "$SRC/Lib/TLibCommon/tools_YS.cpp"
cp "$HERE/ref_shims/cucd_dump.h" "$SRC/Lib/TLibCommon/cucd_dump.h"
cp "$HERE/ref_shims/hmref_driver.cpp" "$SRC/hmref_driver.cpp"

# ---- 3. KAT dump hooks ---------------------------------------------------------------------
python3 - "$SRC" <<'PY'
import sys, re, io
src = sys.argv[1]
def patch(path, edits):
    p = f"{src}/{path}"
    s = open(p, encoding="latin-1").read()
    for anchor, repl, where in edits:
        n = s.count(anchor)
        assert n == 1, (path, anchor, n)
        s = s.replace(anchor, (anchor + repl) if where == "after" else (repl + anchor))
    open(p, "w", encoding="latin-1").write(s)

# neighbour flags (TComPattern.cpp:134-139)
patch("Lib/TLibCommon/TComPattern.cpp", [
  ("  bAbove = true;\n  bLeft  = true;\n",
   "  cucd_hook_flags(compID == COMPONENT_Y, bNeighborFlags, iLeftUnits + iAboveUnits + 1, iNumIntraNeighbor);\n", "after"),
])
# RMD loop (TEncSearch.cpp:2327-2361) and integer-ME probe (TEncSearch.cpp:421-424)
patch("Lib/TLibEncoder/TEncSearch.cpp", [
  ("      for (Int modeIdx = 0; modeIdx < numModesAvailable; modeIdx++)\n      {\n        UInt       uiMode = modeIdx;\n",
   "      cucd_hook_rmd_begin(g_iPOC, pcCU->getCUPelX() + puRect.x0, pcCU->getCUPelY() + puRect.y0, puRect.width, g_bitDepth[CHANNEL_TYPE_LUMA],\n"
   "                          m_piYuvExt[COMPONENT_Y][PRED_BUF_UNFILTERED], m_piYuvExt[COMPONENT_Y][PRED_BUF_FILTERED], piOrg, uiStride);\n", "before"),
  ("        // do intra prediciton \n        predIntraAng(COMPONENT_Y, uiMode, piOrg, uiStride, piPred, uiStride, tuRecurseWithPU,",
   "#ifdef CUCD_INTEGRATION\n        if (cucd_shim_rmd_active()) uiSad += cucd_shim_rmd_sad(modeIdx);      /* INTEGRATION.md S2 */\n        else {\n#endif\n", "before"),
  ("        uiSad += distParam.DistFunc(&distParam);  // DistFunc is a member of DistParam class \n",
   "#ifdef CUCD_INTEGRATION\n        }\n#endif\n        cucd_hook_rmd_mode(modeIdx, uiSad);\n", "after"),
  ("      //////////////  End of RMD ///",
   "      cucd_hook_rmd_end();\n", "before"),
  ("    uiSad = m_cDistParam.DistFunc(&m_cDistParam);\n\n    // motion cost\n    uiSad += m_pcRdCost->getCost(iSearchX, iSearchY);\n\n    if (uiSad < rcStruct.uiBestSad)",
   "    cucd_hook_me(m_cDistParam.pOrg, m_cDistParam.iStrideOrg, m_cDistParam.pCur, m_cDistParam.iStrideCur, m_cDistParam.iCols, m_cDistParam.iRows,\n"
   "                 m_cDistParam.iSubShift, m_cDistParam.bitDepth, iSearchX, iSearchY, m_cDistParam.DistFunc(&m_cDistParam));\n"
   "#ifdef CUCD_INTEGRATION\n    if (cucd_shim_me_active()) uiSad = cucd_shim_me_sad(iSearchX, iSearchY) + m_pcRdCost->getCost(iSearchX, iSearchY); else\n#endif\n"
   "    {\n", "before"),
  ("    uiSad = m_cDistParam.DistFunc(&m_cDistParam);\n\n    // motion cost\n    uiSad += m_pcRdCost->getCost(iSearchX, iSearchY);\n",
   "    }\n", "after"),
])
# intra luma TU coding (TEncSearch.cpp:1092-1387)
patch("Lib/TLibCommon/TComTrQuant.h", [
  ("  TCoeff* m_plTempCoeff;\n", "public:\n  TCoeff* cucdTempCoeff() { return m_plTempCoeff; }\nprotected:\n", "after"),
])
patch("Lib/TLibEncoder/TEncSearch.cpp", [
  ("  //===== get residual signal =====\n",
   "  cucd_hook_tu_pred(compID, g_iPOC, pcCU->getCUPelX() + blkX, pcCU->getCUPelY() + blkY, uiWidth, uiChFinalMode, g_bitDepth[chType], useTransformSkip, default0Save1Load2 == 2,\n"
   "                    m_piYuvExt[compID][PRED_BUF_UNFILTERED], piPred, piOrg, uiStride);\n"
   "#ifdef CUCD_INTEGRATION\n"
   "  cucd_shim_tu_forward(compID, uiWidth, uiChFinalMode, QpParam(*pcCU, compID).Qp - 6 * (g_bitDepth[chType] - 8), useTransformSkip,\n"
   "                       m_piYuvExt[compID][PRED_BUF_UNFILTERED], piOrg, piPred, uiStride);   /* INTEGRATION.md, TU coding */\n"
   "#endif\n", "before"),
  ("  //--- inverse transform ---\n",
   "  cucd_hook_tu_coeff(cQP.Qp - 6 * (g_bitDepth[chType] - 8), pcCU->getSlice()->getSliceType() == I_SLICE, pcCU->getSlice()->getPPS()->getSignHideFlag(),\n"
   "                                  useTransformSkip ? m_pcEncCfg->getUseRDOQTS() : m_pcEncCfg->getUseRDOQ(), m_pcTrQuant->cucdTempCoeff(), pcCoeff, uiAbsSum);\n"
   "#ifdef CUCD_INTEGRATION\n  cucd_shim_tu_after_quant(m_pcTrQuant->cucdTempCoeff(), pcCoeff);\n#endif\n", "before"),
  ("  //===== update distortion =====\n",
   "#ifdef CUCD_INTEGRATION\n"
   "  cucd_shim_tu_reco(piReco, uiStride, piRecQt, uiRecQtStride, piRecIPred, uiRecIPredStride,\n"
   "                    m_pcRdCost->getDistPart(g_bitDepth[chType], piReco, uiStride, piOrg, uiStride, uiWidth, uiHeight, COMPONENT_Y));\n"
   "#endif\n", "before"),
  ("  //===== update distortion =====\n  ruiDist += m_pcRdCost->getDistPart(g_bitDepth[chType], piReco, uiStride, piOrg, uiStride, uiWidth, uiHeight, compID);\n",
   "  cucd_hook_tu_end(piReco, uiStride, m_pcRdCost->getDistPart(g_bitDepth[chType], piReco, uiStride, piOrg, uiStride, uiWidth, uiHeight, COMPONENT_Y));   /* unweighted SSE */\n", "after"),
])
# fractional-pel refinement candidates (TEncSearch.cpp:808-865, 4340-4376)
patch("Lib/TLibEncoder/TEncSearch.cpp", [
  ("  //  Half-pel refinement\n  xExtDIFUpSamplingH(&cPatternRoi, biPred);\n",
   "  cucd_hook_frac_begin(piRefY + iOffset, iRefStride);\n"
   "#ifdef CUCD_INTEGRATION\n  cucd_shim_frac_begin(biPred, pcMvInt->getHor(), pcMvInt->getVer(), m_pcEncCfg->getUseHADME() && !bIsLosslessCoded);   /* INTEGRATION.md 8f.3 */\n#endif\n", "before"),
  ("  ruiCost = xPatternRefinement(pcPatternKey, baseRefMv, 1, rcMvQter, !bIsLosslessCoded);\n",
   "#ifdef CUCD_INTEGRATION\n  cucd_shim_frac_end();\n#endif\n", "after"),
  ("    uiDist = m_cDistParam.DistFunc(&m_cDistParam);\n    uiDist += m_pcRdCost->getCost(cMvTest.getHor(), cMvTest.getVer());\n",
   "    cucd_hook_frac_cand(m_cDistParam.pOrg, m_cDistParam.iStrideOrg, m_cDistParam.iCols, m_cDistParam.iRows, m_cDistParam.bitDepth,\n"
   "                        m_pcEncCfg->getUseHADME() && bAllowUseOfHadamard, horVal, verVal, m_cDistParam.DistFunc(&m_cDistParam));\n"
   "#ifdef CUCD_INTEGRATION\n    if (cucd_shim_frac_active()) { uiDist = cucd_shim_frac_cost(horVal, verVal); uiDist += m_pcRdCost->getCost(cMvTest.getHor(), cMvTest.getVer()); } else\n#endif\n    {\n", "before"),
  ("    uiDist = m_cDistParam.DistFunc(&m_cDistParam);\n    uiDist += m_pcRdCost->getCost(cMvTest.getHor(), cMvTest.getVer());\n", "    }\n", "after"),
])
# S3: integer ME through SAD surfaces (integration build only)
patch("Lib/TLibEncoder/TEncSearch.cpp", [
  ("  setWpScalingDistParam(pcCU, iRefIdxPred, eRefPicList);\n  //  Do integer search\n",
   "#ifdef CUCD_INTEGRATION\n"
   "  if ((!bBi || !cucd_ipc_mode()) && m_pcEncCfg->getFastSearch() != SELECTIVE) {   /* INTEGRATION.md S3; bBi: the key is m_cYuvPredTemp */\n"
   "    TComPicYuv* cucdRef = pcCU->getSlice()->getRefPic(eRefPicList, iRefIdxPred)->getPicYuvRec();\n"
   "    TComPicYuv* cucdOrg = pcCU->getPic()->getPicYuvOrg();\n"
   "    const long cucdOff = (long)(piRefY - cucdRef->getAddr(COMPONENT_Y));\n"
   "    const int cucdPuY = (int)(cucdOff / iRefStride), cucdPuX = (int)(cucdOff - (long)cucdPuY * iRefStride);\n"
   "    cucd_shim_me_begin(cucdOrg->getWidth(COMPONENT_Y), cucdOrg->getHeight(COMPONENT_Y), g_bitDepth[CHANNEL_TYPE_LUMA],\n"
   "                       pcCU->getSlice()->getSPS()->getUseStrongIntraSmoothing() ? 1 : 0, pcCU->getSlice()->getPOC(),\n"
   "                       cucdOrg->getAddr(COMPONENT_Y), cucdOrg->getStride(COMPONENT_Y), cucdRef, cucdRef->getAddr(COMPONENT_Y), iRefStride,\n"
   "                       cucdRef->getMarginX(COMPONENT_Y), cucdRef->getMarginY(COMPONENT_Y), pcCU->getCUPelX(), pcCU->getCUPelY(), cucdPuX, cucdPuY,\n"
   "                       iRoiWidth, iRoiHeight, (m_pcEncCfg->getUseFastEnc() && iRoiHeight > 8) ? 1 : 0,\n"
   "                       bBi ? pcYuv->getAddr(COMPONENT_Y, uiPartAddr) : (const Pel*)0, bBi ? (int)pcYuv->getStride(COMPONENT_Y) : 0);\n"
   "  }\n#endif\n", "after"),
  ("  m_pcRdCost->setCostScale(1);\n\n  const Bool bIsLosslessCoded",
   "#ifdef CUCD_INTEGRATION\n  cucd_shim_me_end();\n#endif\n", "before"),
  ("      uiSad = m_cDistParam.DistFunc(&m_cDistParam);\n\n      // motion cost\n      uiSad += m_pcRdCost->getCost(x, y);\n",
   "#ifdef CUCD_INTEGRATION\n      if (cucd_shim_me_active()) uiSad = cucd_shim_me_sad(x, y) + m_pcRdCost->getCost(x, y);\n#endif\n", "after"),
])
# AMVP candidate check and merge candidates: source vs uni-directional motion-compensated prediction (integration build only)
patch("Lib/TLibEncoder/TEncSearch.cpp", [
  ("  uiCost = m_pcRdCost->getDistPart(g_bitDepth[CHANNEL_TYPE_LUMA], pcTemplateCand->getAddr(COMPONENT_Y, uiPartAddr), pcTemplateCand->getStride(COMPONENT_Y), pcOrgYuv->getAddr(COMPONENT_Y, uiPartAddr), pcOrgYuv->getStride(COMPONENT_Y), iSizeX, iSizeY, COMPONENT_Y, DF_SAD);\n",
   "#ifdef CUCD_INTEGRATION\n"
   "  if (cucd_shim_mc_enabled() && !(pcCU->getSlice()->testWeightPred() && pcCU->getSlice()->getSliceType() == P_SLICE)) {   /* xGetTemplateCost */\n"
   "    TComPicYuv* cucdOrg = pcCU->getPic()->getPicYuvOrg();\n"
   "    const long cucdOff = (long)(pcPicYuvRef->getAddr(COMPONENT_Y, pcCU->getCtuRsAddr(), pcCU->getZorderIdxInCtu() + uiPartAddr) - pcPicYuvRef->getAddr(COMPONENT_Y));\n"
   "    const int cucdRs = pcPicYuvRef->getStride(COMPONENT_Y), cucdPuY = (int)(cucdOff / cucdRs), cucdPuX = (int)(cucdOff - (long)cucdPuY * cucdRs);\n"
   "    uiCost = cucd_shim_mc_dist(cucdOrg->getWidth(COMPONENT_Y), cucdOrg->getHeight(COMPONENT_Y), g_bitDepth[CHANNEL_TYPE_LUMA],\n"
   "                               pcCU->getSlice()->getSPS()->getUseStrongIntraSmoothing() ? 1 : 0, pcCU->getSlice()->getPOC(),\n"
   "                               cucdOrg->getAddr(COMPONENT_Y), cucdOrg->getStride(COMPONENT_Y), pcPicYuvRef, pcPicYuvRef->getAddr(COMPONENT_Y), cucdRs,\n"
   "                               pcPicYuvRef->getMarginX(COMPONENT_Y), pcPicYuvRef->getMarginY(COMPONENT_Y), cucdPuX, cucdPuY, iSizeX, iSizeY,\n"
   "                               cMvCand.getHor(), cMvCand.getVer(), 0, 0);\n"
   "  } else\n#endif\n", "before"),
  ("  motionCompensation(pcCU, &m_tmpYuvPred, REF_PIC_LIST_X, iPartIdx);\n\n  UInt uiAbsPartIdx = 0;\n",
   "#ifdef CUCD_INTEGRATION\n"
   "  if (cucd_shim_mc_enabled()) {   /* xGetInterPredictionError: uni-directional candidates (bi-predictive ones average two 14-bit predictions: host) */\n"
   "    UInt cucdPart = 0; Int cucdW = 0, cucdH = 0;\n"
   "    pcCU->getPartIndexAndSize(iPartIdx, cucdPart, cucdW, cucdH);\n"
   "    const Int cucdR0 = pcCU->getCUMvField(REF_PIC_LIST_0)->getRefIdx(cucdPart), cucdR1 = pcCU->getCUMvField(REF_PIC_LIST_1)->getRefIdx(cucdPart);\n"
   "    const Bool cucdWp = (pcCU->getSlice()->getPPS()->getUseWP() && pcCU->getSlice()->getSliceType() == P_SLICE) || (pcCU->getSlice()->getPPS()->getWPBiPred() && pcCU->getSlice()->getSliceType() == B_SLICE);\n"
   "    if (((cucdR0 >= 0) != (cucdR1 >= 0)) && !cucdWp) {\n"
   "      const RefPicList cucdList = cucdR0 >= 0 ? REF_PIC_LIST_0 : REF_PIC_LIST_1;\n"
   "      TComMv cucdMv = pcCU->getCUMvField(cucdList)->getMv(cucdPart);\n"
   "      pcCU->clipMv(cucdMv);\n"
   "      TComPicYuv* cucdRef = pcCU->getSlice()->getRefPic(cucdList, cucdR0 >= 0 ? cucdR0 : cucdR1)->getPicYuvRec();\n"
   "      TComPicYuv* cucdOrg = pcCU->getPic()->getPicYuvOrg();\n"
   "      const long cucdOff = (long)(cucdRef->getAddr(COMPONENT_Y, pcCU->getCtuRsAddr(), pcCU->getZorderIdxInCtu() + cucdPart) - cucdRef->getAddr(COMPONENT_Y));\n"
   "      const int cucdRs = cucdRef->getStride(COMPONENT_Y), cucdPuY = (int)(cucdOff / cucdRs), cucdPuX = (int)(cucdOff - (long)cucdPuY * cucdRs);\n"
   "      ruiErr = cucd_shim_mc_dist(cucdOrg->getWidth(COMPONENT_Y), cucdOrg->getHeight(COMPONENT_Y), g_bitDepth[CHANNEL_TYPE_LUMA],\n"
   "                                 pcCU->getSlice()->getSPS()->getUseStrongIntraSmoothing() ? 1 : 0, pcCU->getSlice()->getPOC(),\n"
   "                                 cucdOrg->getAddr(COMPONENT_Y), cucdOrg->getStride(COMPONENT_Y), cucdRef, cucdRef->getAddr(COMPONENT_Y), cucdRs,\n"
   "                                 cucdRef->getMarginX(COMPONENT_Y), cucdRef->getMarginY(COMPONENT_Y), cucdPuX, cucdPuY, cucdW, cucdH,\n"
   "                                 cucdMv.getHor(), cucdMv.getVer(), (m_pcEncCfg->getUseHADME() && (pcCU->getCUTransquantBypass(iPartIdx) == 0)) ? 1 : 0, 1);\n"
   "      return;\n"
   "    }\n"
   "  }\n#endif\n", "before"),
])
# a12: TMV features (TEncCu.cpp:1561), integration build only (a13 / AdaptiveQP: the reference itself crashes with --AdaptiveQP=1, no in-situ test)
patch("Lib/TLibEncoder/TEncCu.cpp", [
  ("       TMVFeature* feature_x = getTMVFeature(rpcBestCU);\n",
   "#ifdef CUCD_INTEGRATION\n       cucd_shim_tmv_check(rpcBestCU->getCUPelX(), rpcBestCU->getCUPelY(), rpcBestCU->getWidth(0), &feature_x->m_adFeature[0][0]);\n#endif\n", "after"),
])
# S1 call site (TEncGOP.cpp:1095-1096)
patch("Lib/TLibEncoder/TEncGOP.cpp", [
  ("\t\t  m_pcSliceEncoder->getOutlierWithDCT(pcPic);\n",
   "#ifdef CUCD_INTEGRATION\n"
   "      cucd_shim_outlier(pcPic->getPicYuvOrg()->getWidth(COMPONENT_Y), pcPic->getPicYuvOrg()->getHeight(COMPONENT_Y), g_bitDepth[CHANNEL_TYPE_LUMA],\n"
   "                        pcSlice->getSPS()->getUseStrongIntraSmoothing() ? 1 : 0,\n"
   "                        pcPic->getPicYuvOrg()->getAddr(COMPONENT_Y), pcPic->getPicYuvOrg()->getStride(COMPONENT_Y),\n"
   "                        pcPic->getOBF()->getAddr(COMPONENT_Y), pcPic->getOBF()->getStride(COMPONENT_Y),\n"
   "                        pcPic->getPicYuvOutlier()->getAddr(COMPONENT_Y), pcPic->getPicYuvOutlier()->getStride(COMPONENT_Y));   /* INTEGRATION.md S1 */\n"
   "#else\n", "before"),
  ("\t\t  m_pcSliceEncoder->getOutlierWithDCT(pcPic);\n", "#endif\n", "after"),
])
# outlier picture pass (TEncSlice.cpp:878-1173)
patch("Lib/TLibEncoder/TEncSlice.cpp", [
  ("\t\tdelete[] Yc;\n", "\t\tcucd_hook_obf_yc(Yc, FrequencySize);\n", "before"),
  ("#if OUT_OUTLIER  // write files\n",
   "        cucd_hook_obf(rpcPic->getPOC(), uiFrameWidth, uiFrameHeight, bitDepth, pData, uiStrideSrc, pOBF, uiStrideOBF, pDst, uiStrideDst);\n", "before"),
])
# per-CU OBF block sums (TEncCu.cpp:589-600)
patch("Lib/TLibEncoder/TEncCu.cpp", [
  ("\t\tuiHasOutlier = Num_OBF>0 ? 1 : 0;\n\t\tN_NonZeroFeature = Num_OBF;\n",
   "\t\tif (!bBoundary) cucd_hook_cu(g_iPOC, uiDepth, uiLPelX, uiTPelY, BlockSize, Num_OBF, N_Outlier);\n"
   "\t\tcucd_hook_switches(g_iPOC, (int)g_mainModelType, g_bDecisionSwitch);\n", "after"),
])
PY

# ---- 4. build --------------------------------------------------------------------------------
cat > "$WORK/build.mk" <<'EOF'
CXX := g++
CXXFLAGS := -std=c++11 -O3 -w -fpermissive -fPIC -include limits -include cstring -include cucd_dump.h -I$(SRC)/Lib -I$(SRC)/App -I$(SRC)/Lib/TLibCommon
CFLAGS := -O3 -w -fPIC
COMMON_CPP := $(filter-out %/basic_tools_YS.cpp %/svm.cpp %/train_linear.cpp,$(wildcard $(SRC)/Lib/TLibCommon/*.cpp))
COMMON_C := $(wildcard $(SRC)/Lib/TLibCommon/*.c) $(wildcard $(SRC)/Lib/libmd5/*.c)
SUPPORT_CPP := $(wildcard $(SRC)/Lib/TLibVideoIO/*.cpp) $(wildcard $(SRC)/Lib/TAppCommon/*.cpp)
ENC_CPP := $(wildcard $(SRC)/Lib/TLibEncoder/*.cpp)
APP_CPP := $(wildcard $(SRC)/App/TAppEncoder/*.cpp)
DEC_CPP := $(wildcard $(SRC)/Lib/TLibDecoder/*.cpp) $(wildcard $(SRC)/App/TAppDecoder/*.cpp)
o = $(patsubst $(SRC)/%,$(OBJ)/%.o,$(1))
BASE_O := $(call o,$(COMMON_CPP) $(SUPPORT_CPP)) $(call o,$(COMMON_C))
ENC_O := $(call o,$(ENC_CPP))
APP_O := $(call o,$(APP_CPP))
DEC_O := $(call o,$(DEC_CPP))
# encmain.cpp defines main() AND the fork's global output streams; the driver library reuses the
# object with main renamed so that those globals exist without a second definition.
APP_LIB_O := $(patsubst %.o,%.lib.o,$(APP_O))
# integration build: the same patched sources with -DCUCD_INTEGRATION, linked against the product library
INT_O := $(patsubst %.o,%.int.o,$(ENC_O) $(APP_O))
all: $(OUT)/TAppEncoder $(OUT)/TAppDecoder $(OUT)/libhmref.so $(if $(wildcard $(PKGDIR)/libcucudecide.so),$(OUT)/TAppEncoderCucd)
$(OBJ)/%.cpp.o: $(SRC)/%.cpp
	@mkdir -p $(dir $@)
	$(CXX) $(CXXFLAGS) -c $< -o $@
$(OBJ)/%.cpp.lib.o: $(SRC)/%.cpp
	@mkdir -p $(dir $@)
	$(CXX) $(CXXFLAGS) -Dmain=hm_encoder_main -c $< -o $@
$(OBJ)/%.cpp.int.o: $(SRC)/%.cpp
	@mkdir -p $(dir $@)
	$(CXX) $(CXXFLAGS) -DCUCD_INTEGRATION -I$(INCDIR) -c $< -o $@
$(OBJ)/%.c.o: $(SRC)/%.c
	@mkdir -p $(dir $@)
	gcc $(CFLAGS) -c $< -o $@
$(OUT)/TAppEncoder: $(BASE_O) $(ENC_O) $(APP_O)
	$(CXX) -o $@ $^ -lm
$(OUT)/TAppDecoder: $(BASE_O) $(DEC_O)
	$(CXX) -o $@ $^ -lm
$(OUT)/TAppEncoderCucd: $(BASE_O) $(INT_O) $(PKGDIR)/libcucudecide.so
	$(CXX) -o $@ $(BASE_O) $(INT_O) -L$(PKGDIR) -lcucudecide -Wl,-rpath,'$$ORIGIN/../../fast-cu-decision-hevc_b200' -lm -lpthread -lrt
$(OUT)/libhmref.so: $(BASE_O) $(ENC_O) $(APP_LIB_O) $(OBJ)/hmref_driver.cpp.o
	$(CXX) -shared -o $@ $^ -lm -lpthread
EOF
make -s -f "$WORK/build.mk" -j"$JOBS" SRC="$SRC" OBJ="$WORK/obj" OUT="$OUT" INCDIR="$HERE/../include" PKGDIR="$HERE/../fast-cu-decision-hevc_b200" all
echo "$SIG" > "$STAMP"
ls -la "$OUT"
