/* cucd_oracle.h - CPU restatement of the reference's CU-decision cost arithmetic.
 *
 * TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * load this.  The product (libcucudecide.so) never links or calls it.
 *
 * Parity status: PINNED.  Every function below is checked (tests/test_oracle_golden.py) against
 * known-answer vectors dumped from the reference encoder's own call sites (the .npz files in tests/golden/,
 * made by tests/golden/gen_golden.py from oracle/_ref/TAppEncoder) and, where oracle/_ref is
 * present, against the reference's compiled functions on random inputs (tests/test_oracle_vs_ref.py).
 *
 * Conventions shared with include/cucudecide.h:
 *   sample     int16_t (HM `Pel`, TypeDef.h:769), 8..12-bit content
 *   border     linear array of 4N+1 samples: [0..2N-1] left column bottom(below-left end)->top,
 *              [2N] top-left corner, [2N+1..4N] above row left->right (above-right end last)
 *   flags      one byte per 4-sample unit in the same order: N/2 left units, 1 corner, N/2 above
 *   cost table uint32_t[35], index = HEVC intra mode (0 planar, 1 DC, 2..34 angular)
 *   PU order   inside a CTU: depth-major (64,32,16,8,4), z-order (Morton) inside a depth -> 341 PUs
 */
#ifndef CUCD_ORACLE_H
#define CUCD_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define ORACLE_PUS_PER_CTU 341
#define ORACLE_NUM_MODES 35

/* TComPattern.cpp:550-727 in frame (replay) mode == HEVC 6.4.1 z-scan availability */
int  oracle_unit_available(int xc, int yc, int xn, int yn, int W, int H);
void oracle_neighbour_flags(int x0, int y0, int n, int W, int H, uint8_t* flags);
/* TComPattern.cpp:314-521 */
void oracle_fill_border(int bitDepth, int n, const int16_t* recOrigin, int recStride, const uint8_t* flags, int16_t* border);
/* TComPattern.cpp:185-283 */
void oracle_filter_border(int bitDepth, int n, int strongSmoothing, const int16_t* border, int16_t* filtered);
/* TComPattern.cpp:523-548, TComPrediction.cpp:50-67 */
int  oracle_use_filtered(int n, int mode);
/* TComPrediction.cpp:183-222, 250-410, 412-496, 755-841; pred is n x n, stride n */
void oracle_predict(int bitDepth, int n, int mode, const int16_t* unfiltered, const int16_t* filtered, int16_t* pred);
/* TComRdCost.cpp:1321-1604 */
uint32_t oracle_satd(int bitDepth, const int16_t* org, int orgStride, const int16_t* cur, int curStride, int w, int h);
/* TComRdCost.cpp:465-962 */
uint32_t oracle_sad(int bitDepth, const int16_t* org, int orgStride, const int16_t* ref, int refStride, int w, int h, int subShift);
/* TEncSearch.cpp:2327-2361 without the mode-bit term */
void oracle_rmd_pu(int bitDepth, int n, int strongSmoothing, const int16_t* org, int orgStride, const int16_t* border, uint32_t* sad35);
/* whole-frame replay enumeration; out[(ctu-ctuBegin)*341*35 + pu*35 + mode], 0xFFFFFFFF for PUs outside the picture */
void oracle_rmd_frame(int bitDepth, int strongSmoothing, const int16_t* org, int orgStride, const int16_t* rec, int recStride,
                      int W, int H, int ctuBegin, int ctuEnd, uint32_t* out);
/* TEncSearch.cpp:3886-3943 raw SAD part; out[(dy-top)*(right-left+1)+(dx-left)] */
void oracle_sad_surface(int bitDepth, const int16_t* org, int orgStride, int w, int h, const int16_t* refAtZeroMv, int refStride,
                        int left, int right, int top, int bottom, int subShift, uint32_t* out);
/* TEncSlice.cpp:55-77 + 944-966: forward 4x4 DCT of one block -> coeff[16] (row = vertical frequency) */
void oracle_dct4x4(int bitDepth, const int16_t* blk, int stride, int32_t* coeff);
/* TEncSlice.cpp:291-392 on a histogram of |coeff/8| : count[0..peak], len samples -> Yc */
double oracle_tcm_yc(const uint32_t* count, int peak, int len);
/* TEncSlice.cpp:878-1173. obf (W/4)x(H/4) tight, outlier WxH tight, yc[16] (yc[0] unused) */
void oracle_outlier_frame(int bitDepth, const int16_t* org, int orgStride, int W, int H, int16_t* obf, int16_t* outlier, double* yc);
/* TEncCu.cpp:589-600 for every CU of depth d (size 64>>d) fully inside the picture, raster order over
 * the (W/size)x(H/size) grid of whole CUs: numObf = #cells>0, nOutlier = sum of cells */
void oracle_cu_sums(const int16_t* obf, int W, int H, int depth, int32_t* numObf, int32_t* nOutlier);
/* TEncCu.cpp:1780-1893: per CTU sum of DC-less 8x8 source Hadamard costs (whole 8x8 blocks inside the picture) */
void oracle_ctu_src_had(const int16_t* org, int orgStride, int W, int H, int32_t* perCtu);
/* tools_YS.cpp:1659-1839 (getTMVFeature): feat[5][26] of one n x n CU (n = 8..64) */
void oracle_tmv_features(const int16_t* cu, int stride, int n, double* feat);
/* TEncPreanalyzer.cpp:64-139 for one AQ layer of part x part units; activity ceil(W/part) x ceil(H/part); returns the average */
double oracle_aq_activity(const int16_t* org, int stride, int W, int H, int part, double* activity);


/* ---- intra luma TU coding, xIntraCodingTUBlock TEncSearch.cpp:1092-1387 (n = 4..32) -------------------- */
int  oracle_dct_coef(int n, int k, int x);                                              /* g_aiT<n>[k][x], TComRom.cpp:356-460 */
void oracle_fwd_transform(int bitDepth, int n, int useDST, const int16_t* resi, int stride, int32_t* coeff);   /* TComTrQuant.cpp:860-919 */
void oracle_inv_transform(int bitDepth, int n, int useDST, const int32_t* coeff, int16_t* resi, int stride);   /* :927-985 */
void oracle_transform_skip(int bitDepth, int n, const int16_t* resi, int stride, int32_t* coeff);              /* :1933-1978 */
void oracle_inv_transform_skip(int bitDepth, int n, const int32_t* coeff, int16_t* resi, int stride);          /* :1980-2031 */
int  oracle_scan_idx(int n, int mode);                                                  /* TComDataCU.cpp:3356-3410; 0 diag 1 hor 2 ver */
void oracle_scan_order(int scanIdx, int n, uint16_t* scan);                             /* TComRom.cpp:53-228, grouped 4x4 */
int  oracle_quant(int bitDepth, int n, int qp, int intraSlice, int signHiding, int scanIdx, const int32_t* coef, int32_t* level); /* :991-1240 */
void oracle_dequant(int bitDepth, int n, int qp, const int32_t* level, int32_t* coef);  /* :1242-1352 */
uint32_t oracle_sse(int bitDepth, const int16_t* org, int os, const int16_t* cur, int cs, int w, int h);       /* TComRdCost.cpp:970-1315 */
/* stage 0: prediction + residual + transform (coef, pred out); 1: whole chain with the plain quantiser; 2: reconstruction from given levels */
void oracle_intra_tu(int bitDepth, int n, int mode, int qp, int transformSkip, int strongSmoothing, int intraSlice, int signHiding, int stage,
                     const int16_t* org, int orgStride, const int16_t* border, int32_t* coef, int32_t* level, int16_t* pred, int16_t* reco,
                     uint32_t* dist, int32_t* absSum);

/* ---- fractional-pel ME refinement (TEncSearch.cpp:808-865, 4340-4376, 5431-5637; TComInterpolationFilter.cpp) ---- */
void oracle_interp_luma(int bitDepth, const int16_t* ref, int stride, int w, int h, int fx, int fy, int16_t* out);
uint32_t oracle_subpel_cost(int bitDepth, const int16_t* org, int orgStride, int w, int h, const int16_t* refAtZeroMv, int refStride,
                            int qx, int qy, int useHadamard);
void oracle_subpel_surface(int bitDepth, const int16_t* org, int orgStride, int w, int h, const int16_t* refAtZeroMv, int refStride,
                           int mvx, int mvy, int useHadamard, uint32_t* out);
/* chroma blocks of a 4:2:0 picture (estIntraPredChromaQT -> xIntraCodingTUBlock): unfiltered references, no DC / edge
 * filters, DCT only, mode-dependent scan for 4x4 only */
void oracle_predict_chroma(int bitDepth, int n, int mode, const int16_t* unfiltered, int16_t* pred);
int  oracle_scan_idx_c(int n, int mode, int chroma);
void oracle_intra_tu_c(int bitDepth, int n, int mode, int qp, int transformSkip, int chroma, int strongSmoothing, int intraSlice, int signHiding, int stage,
                       const int16_t* org, int orgStride, const int16_t* border, int32_t* coef, int32_t* level, int16_t* pred, int16_t* reco,
                       uint32_t* dist, int32_t* absSum);

/* ---- fork-aware enumeration: which luma RMD PUs TEncCu::xCompressCU still evaluates on a Testing picture (intra slice, default Naive
 * model on N_OBF, tools_YS.cpp:686-695) given the per-depth switches g_bDecisionSwitch[d][model][Skip2Nx2N / TerminateCU]
 * (TEncCu.cpp:488-489, 645-675, 951-996, 1040, 1140-1143, 1257-1260).  obf = the picture's OBF plane; needed[nCtu][341], table PU order. */
void oracle_prune_mask(const int16_t* obf, int W, int H, const uint8_t* swSkip2Nx2N, const uint8_t* swTerminateCU, uint8_t* needed);

#ifdef __cplusplus
}
#endif
#endif
