/* cucd_oracle.c - plain-C restatement of the reference's CU-decision cost arithmetic.
 *
 * TEST INFRASTRUCTURE (see cucd_oracle.h): the checker for the CUDA path, never the product.
 * Parity status: PINNED against reference-encoder KAT dumps (tests/golden/) and oracle/_ref.
 * All citations are file:line in /root/reference (Jiraiya812/Fast-CU-Decision-HEVC).
 * Scalar, single-threaded, written for clarity rather than speed.
 */
#include "cucd_oracle.h"
#include <stdlib.h>
#include <string.h>
#include <math.h>

static int ilog2(int n) { int l = 0; while ((1 << l) < n) l++; return l; }
static int iabs(int v) { return v < 0 ? -v : v; }
static int clip3(int lo, int hi, int v) { return v < lo ? lo : (v > hi ? hi : v); }

/* ------------------------------------------------------------------------------------------
 * Neighbour availability (frame / replay mode).
 * The reference asks TComDataCU (TComPattern.cpp:550-727 -> getPUAbove/Left/AboveRight/BelowLeft)
 * which, for one slice, one tile and constrained_intra_pred off, is the HEVC 6.4.1 rule: a
 * neighbouring 4x4 unit is available iff it is inside the picture and precedes the current block
 * in decoding order (CTU raster order, z-scan inside the CTU).
 * ------------------------------------------------------------------------------------------ */
static unsigned zscan4(unsigned ux, unsigned uy) {
  unsigned r = 0; int b;
  for (b = 0; b < 4; b++) r |= ((ux >> b) & 1u) << (2 * b) | ((uy >> b) & 1u) << (2 * b + 1);
  return r;
}

int oracle_unit_available(int xc, int yc, int xn, int yn, int W, int H) {
  int wc, ctuC, ctuN;
  if (xn < 0 || yn < 0 || xn >= W || yn >= H) return 0;
  wc = (W + 63) >> 6;
  ctuC = (yc >> 6) * wc + (xc >> 6);
  ctuN = (yn >> 6) * wc + (xn >> 6);
  if (ctuN != ctuC) return ctuN < ctuC;
  return zscan4((unsigned)(xn & 63) >> 2, (unsigned)(yn & 63) >> 2) < zscan4((unsigned)(xc & 63) >> 2, (unsigned)(yc & 63) >> 2);
}

/* flag order of bNeighborFlags, TComPattern.cpp:134-139 */
void oracle_neighbour_flags(int x0, int y0, int n, int W, int H, uint8_t* flags) {
  const int half = n / 2; /* units per side = 2N/4 */
  int u;
  for (u = 0; u < half; u++) flags[u] = (uint8_t)oracle_unit_available(x0, y0, x0 - 1, y0 + (half - 1 - u) * 4, W, H);
  flags[half] = (uint8_t)oracle_unit_available(x0, y0, x0 - 1, y0 - 1, W, H);
  for (u = 0; u < half; u++) flags[half + 1 + u] = (uint8_t)oracle_unit_available(x0, y0, x0 + u * 4, y0 - 1, W, H);
}

/* ------------------------------------------------------------------------------------------
 * Reference-sample gathering and substitution: fillReferenceSamples, TComPattern.cpp:314-521.
 * Sample-wise statement of the same process: walk the border from the below-left end to the
 * above-right end; if nothing is available use 1<<(bd-1); an unavailable first unit takes the
 * first available sample found walking forward; every other unavailable sample repeats its
 * predecessor.
 * ------------------------------------------------------------------------------------------ */
void oracle_fill_border(int bitDepth, int n, const int16_t* rec, int stride, const uint8_t* flags, int16_t* b) {
  const int n2 = 2 * n, total = 4 * n + 1, units = n + 1; /* n/2 + 1 + n/2 units of 4 -> but corner is 1 sample */
  uint8_t av[4 * 64 + 1];
  int i, any = 0;
  (void)units;
  for (i = 0; i < n2; i++) { av[i] = flags[i >> 2]; }
  av[n2] = flags[n / 2];
  for (i = 0; i < n2; i++) { av[n2 + 1 + i] = flags[n / 2 + 1 + (i >> 2)]; }
  for (i = 0; i < total; i++) any |= av[i];
  if (!any) { for (i = 0; i < total; i++) b[i] = (int16_t)(1 << (bitDepth - 1)); return; }
  for (i = 0; i < n2; i++) if (av[i]) b[i] = rec[(n2 - 1 - i) * stride - 1];
  if (av[n2]) b[n2] = rec[-stride - 1];
  for (i = 0; i < n2; i++) if (av[n2 + 1 + i]) b[n2 + 1 + i] = rec[-stride + i];
  if (!av[0]) {
    int j = 1;
    while (!av[j]) j++;
    for (i = 0; i < j; i++) b[i] = b[j];
    /* samples 0..j-1 are now defined; mark so the forward pass keeps them */
    for (i = 0; i < j; i++) av[i] = 1;
  }
  for (i = 1; i < total; i++) if (!av[i]) b[i] = b[i - 1];
}

/* ------------------------------------------------------------------------------------------
 * Reference-sample smoothing: TComPattern.cpp:185-283.
 * ------------------------------------------------------------------------------------------ */
void oracle_filter_border(int bitDepth, int n, int strongSmoothing, const int16_t* b, int16_t* f) {
  const int n2 = 2 * n, last = 4 * n;
  int i, strong = 0;
  if (strongSmoothing && n >= 32) {                                   /* :201-214 */
    const int thr = 1 << (bitDepth - 5);
    const int bl = b[0], tl = b[n2], tr = b[last];
    strong = iabs(bl + tl - 2 * b[n]) < thr && iabs(tl + tr - 2 * b[n2 + n]) < thr;
  }
  f[0] = b[0]; f[last] = b[last];                                     /* :216, :283 */
  if (strong) {                                                        /* :224-272 */
    const int shift = ilog2(n2), bl = b[0], tl = b[n2], tr = b[last];
    for (i = 1; i < n2; i++) f[i] = (int16_t)(((n2 - i) * bl + i * tl + n) >> shift);
    f[n2] = b[n2];
    for (i = 1; i < n2; i++) f[n2 + i] = (int16_t)(((n2 - i) * tl + i * tr + n) >> shift);
  } else {                                                             /* :237-279; the corner uses its two neighbours across the bend */
    for (i = 1; i < last; i++) f[i] = (int16_t)((b[i - 1] + 2 * b[i] + b[i + 1] + 2) >> 2);
  }
}

/* TComPattern.cpp:523-548 with m_aucIntraFilter TComPrediction.cpp:50-67 (luma row) */
int oracle_use_filtered(int n, int mode) {
  static const int thr[5] = {10, 7, 1, 0, 10};
  int d10, d26, diff;
  if (mode == 1) return 0;
  d10 = iabs(mode - 10); d26 = iabs(mode - 26);
  diff = d10 < d26 ? d10 : d26;
  return diff > thr[ilog2(n) - 2];
}

/* ------------------------------------------------------------------------------------------
 * Prediction: predIntraAng TComPrediction.cpp:412-496 for luma, bAbove=bLeft=true
 * (TComPattern.cpp:141-142), edge filters enabled (:481).
 * ------------------------------------------------------------------------------------------ */
static void predict_planar(int n, const int16_t* b, int16_t* pred) {   /* :755-805 */
  const int n2 = 2 * n, sh = ilog2(n);
  const int16_t* top = b + n2 + 1;
  const int tr = top[n], bl = b[n2 - 1 - n];
  int x, y;
  for (y = 0; y < n; y++) {
    const int l = b[n2 - 1 - y];
    for (x = 0; x < n; x++)
      pred[y * n + x] = (int16_t)(((n - 1 - x) * l + (x + 1) * tr + (n - 1 - y) * top[x] + (y + 1) * bl + n) >> (sh + 1));
  }
}

static void predict_dc(int n, const int16_t* b, int16_t* pred, int luma) {        /* :183-222, 266-276, 818-841 */
  const int n2 = 2 * n;
  const int16_t* top = b + n2 + 1;
  int sum = 0, i, x, y, dc;
  for (i = 0; i < n; i++) sum += top[i] + b[n2 - 1 - i];
  dc = (sum + n) / (2 * n);
  for (i = 0; i < n * n; i++) pred[i] = (int16_t)dc;
  if (luma && n <= 16) {                                                /* xDCPredFiltering: luma only (:822) */
    pred[0] = (int16_t)((top[0] + b[n2 - 1] + 2 * dc + 2) >> 2);
    for (x = 1; x < n; x++) pred[x] = (int16_t)((top[x] + 3 * dc + 2) >> 2);
    for (y = 1; y < n; y++) pred[y * n] = (int16_t)((b[n2 - 1 - y] + 3 * dc + 2) >> 2);
  }
}

static void predict_angular(int bitDepth, int n, int mode, const int16_t* b, int16_t* pred, int luma) { /* :278-409 */
  static const int angTable[9] = {0, 2, 5, 9, 13, 17, 21, 26, 32};
  static const int invAngTable[9] = {0, 4096, 1638, 910, 630, 482, 390, 315, 256};
  const int n2 = 2 * n, vertical = mode >= 18;
  const int am = vertical ? mode - 26 : -(mode - 10);
  const int angle = (am < 0 ? -1 : 1) * angTable[iabs(am)], invAngle = invAngTable[iabs(am)];
  int16_t mainBuf[3 * 64 + 2], side[2 * 64 + 1];
  int16_t* ref = mainBuf + 64;  /* ref[0] = corner, ref[1..] along the main edge, ref[-k] projected side samples */
  int i, x, y;
  /* main = above row for vertical modes, left column for horizontal ones (then the block is transposed) */
  for (i = 0; i <= n2; i++) {
    const int16_t t = b[n2 + i];        /* corner, top[0..] */
    const int16_t l = b[n2 - i];        /* corner, left[0..] */
    ref[i] = vertical ? t : l;
    side[i] = vertical ? l : t;
  }
  if (angle < 0) {                      /* :300-322 */
    const int lastIdx = (n * angle) >> 5;
    int k, acc = 128;
    for (k = -1; k > lastIdx; k--) { acc += invAngle; ref[k] = side[acc >> 8]; }
  }
  for (y = 0; y < n; y++) {
    const int delta = (y + 1) * angle, di = delta >> 5, df = delta & 31;
    for (x = 0; x < n; x++) {
      int v;
      if (angle == 0) v = ref[x + 1];
      else if (df) v = ((32 - df) * ref[x + di + 1] + df * ref[x + di + 2] + 16) >> 5;
      else v = ref[x + di + 1];
      if (angle == 0 && x == 0 && n <= 16 && luma)                  /* :284, 356-362: edge filter is luma only */
        v = clip3(0, (1 << bitDepth) - 1, v + ((side[y + 1] - side[0]) >> 1));
      if (vertical) pred[y * n + x] = (int16_t)v; else pred[x * n + y] = (int16_t)v;  /* :397-408 */
    }
  }
}

void oracle_predict(int bitDepth, int n, int mode, const int16_t* unf, const int16_t* fil, int16_t* pred) {
  const int16_t* b = oracle_use_filtered(n, mode) ? fil : unf;
  if (mode == 0) predict_planar(n, b, pred);
  else if (mode == 1) predict_dc(n, b, pred, 1);
  else predict_angular(bitDepth, n, mode, b, pred, 1);
}
/* chroma of 4:2:0 / 4:2:2: unfiltered references only (TComChromaFormat.h:147-150), no DC / edge filters */
void oracle_predict_chroma(int bitDepth, int n, int mode, const int16_t* unf, int16_t* pred) {
  if (mode == 0) predict_planar(n, unf, pred);
  else if (mode == 1) predict_dc(n, unf, pred, 0);
  else predict_angular(bitDepth, n, mode, unf, pred, 0);
}

/* ------------------------------------------------------------------------------------------
 * SATD: xGetHADs TComRdCost.cpp:1537-1604 with xCalcHADs8x8 :1439-1534, 4x4 :1343-1437, 2x2 :1321-1341.
 * Only sum|coeff| matters, so a generic in-place Walsh-Hadamard is used.
 * ------------------------------------------------------------------------------------------ */
static uint32_t had_block(const int16_t* o, int os, const int16_t* c, int cs, int s) {
  int32_t m[64];
  int i, j, len, x, y, cnt = s * s;
  uint32_t sum = 0;
  for (y = 0; y < s; y++) for (x = 0; x < s; x++) m[y * s + x] = o[y * os + x] - c[y * cs + x];
  for (y = 0; y < s; y++)                               /* rows */
    for (len = 1; len < s; len <<= 1)
      for (i = 0; i < s; i += 2 * len)
        for (j = i; j < i + len; j++) { int32_t a = m[y * s + j], b2 = m[y * s + j + len]; m[y * s + j] = a + b2; m[y * s + j + len] = a - b2; }
  for (x = 0; x < s; x++)                               /* columns */
    for (len = 1; len < s; len <<= 1)
      for (i = 0; i < s; i += 2 * len)
        for (j = i; j < i + len; j++) { int32_t a = m[j * s + x], b2 = m[(j + len) * s + x]; m[j * s + x] = a + b2; m[(j + len) * s + x] = a - b2; }
  for (i = 0; i < cnt; i++) sum += (uint32_t)iabs(m[i]);
  if (s == 8) return (sum + 2) >> 2;                    /* :1531 */
  if (s == 4) return (sum + 1) >> 1;                    /* :1434 */
  return sum;                                           /* 2x2 :1335-1340 */
}

uint32_t oracle_satd(int bitDepth, const int16_t* org, int os, const int16_t* cur, int cs, int w, int h) {
  const int s = (w % 8 == 0 && h % 8 == 0) ? 8 : ((w % 4 == 0 && h % 4 == 0) ? 4 : 2);
  uint32_t sum = 0; int x, y;
  for (y = 0; y < h; y += s) for (x = 0; x < w; x += s) sum += had_block(org + y * os + x, os, cur + y * cs + x, cs, s);
  return sum >> (bitDepth - 8);                         /* :1603, DISTORTION_PRECISION_ADJUSTMENT TypeDef.h:343-347 */
}

/* ------------------------------------------------------------------------------------------
 * SAD: TComRdCost.cpp:465-962.  The width-specialised variants (4,8,12,16,24,32,48,64 and
 * multiples of 16 through xGetSAD16N) honour iSubShift; the generic xGetSAD (:465-491), reached
 * for any other width, ignores it.
 * ------------------------------------------------------------------------------------------ */
uint32_t oracle_sad(int bitDepth, const int16_t* org, int os, const int16_t* ref, int rs, int w, int h, int subShift) {
  const int special = (w == 4 || w == 8 || w == 12 || w == 16 || w == 24 || w == 32 || w == 48 || w == 64);
  const int sh = special ? subShift : 0, step = 1 << sh;
  uint32_t sum = 0; int x, y;
  for (y = 0; y < h; y += step) for (x = 0; x < w; x++) sum += (uint32_t)iabs(org[y * os + x] - ref[y * rs + x]);
  sum <<= sh;
  return sum >> (bitDepth - 8);
}

void oracle_sad_surface(int bitDepth, const int16_t* org, int os, int w, int h, const int16_t* ref0, int rs,
                        int left, int right, int top, int bottom, int subShift, uint32_t* out) {
  const int cols = right - left + 1; int x, y;
  for (y = top; y <= bottom; y++) for (x = left; x <= right; x++)
    out[(y - top) * cols + (x - left)] = oracle_sad(bitDepth, org, os, ref0 + y * rs + x, rs, w, h, subShift);
}

/* ------------------------------------------------------------------------------------------
 * Rough mode decision for one PU: TEncSearch.cpp:2327-2361 minus xModeBitsIntra.
 * ------------------------------------------------------------------------------------------ */
void oracle_rmd_pu(int bitDepth, int n, int strong, const int16_t* org, int os, const int16_t* border, uint32_t* sad) {
  int16_t fil[4 * 64 + 1], pred[64 * 64];
  int mode;
  oracle_filter_border(bitDepth, n, strong, border, fil);
  for (mode = 0; mode < 35; mode++) {
    oracle_predict(bitDepth, n, mode, border, fil, pred);
    sad[mode] = oracle_satd(bitDepth, org, os, pred, n, n, n);
  }
}

void oracle_rmd_frame(int bitDepth, int strong, const int16_t* org, int os, const int16_t* rec, int rs,
                      int W, int H, int ctuBegin, int ctuEnd, uint32_t* out) {
  const int wc = (W + 63) >> 6;
  int ctu;
  for (ctu = ctuBegin; ctu < ctuEnd; ctu++) {
    const int cx = (ctu % wc) * 64, cy = (ctu / wc) * 64;
    uint32_t* o = out + (size_t)(ctu - ctuBegin) * ORACLE_PUS_PER_CTU * 35;
    int d, pu = 0;
    for (d = 0; d < 5; d++) {
      const int n = 64 >> d, cnt = 1 << (2 * d);
      int z;
      for (z = 0; z < cnt; z++, pu++) {
        int px = 0, py = 0, bit, m;
        uint8_t flags[65]; int16_t border[4 * 64 + 1];
        for (bit = 0; bit < d; bit++) { px |= ((z >> (2 * bit)) & 1) << bit; py |= ((z >> (2 * bit + 1)) & 1) << bit; }
        if (cx + (px + 1) * n > W || cy + (py + 1) * n > H) { for (m = 0; m < 35; m++) o[pu * 35 + m] = 0xFFFFFFFFu; continue; }
        oracle_neighbour_flags(cx + px * n, cy + py * n, n, W, H, flags);
        oracle_fill_border(bitDepth, n, rec + (size_t)(cy + py * n) * rs + cx + px * n, rs, flags, border);
        oracle_rmd_pu(bitDepth, n, strong, org + (size_t)(cy + py * n) * os + cx + px * n, os, border, o + pu * 35);
      }
    }
  }
}

/* ------------------------------------------------------------------------------------------
 * Outlier feature pass: TEncSlice.cpp:55-77 (4-point butterfly with g_aiT4 = {64,83,36},
 * TComRom.cpp:464-468), :878-1173.
 * ------------------------------------------------------------------------------------------ */
static void dct4_stage(const int32_t* src, int32_t* dst, int shift) {   /* :55-77, line = 4 */
  const int add = shift > 0 ? 1 << (shift - 1) : 0;
  int j;
  for (j = 0; j < 4; j++) {
    const int32_t e0 = src[4 * j] + src[4 * j + 3], o0 = src[4 * j] - src[4 * j + 3];
    const int32_t e1 = src[4 * j + 1] + src[4 * j + 2], o1 = src[4 * j + 1] - src[4 * j + 2];
    dst[j] = (64 * e0 + 64 * e1 + add) >> shift;
    dst[8 + j] = (64 * e0 - 64 * e1 + add) >> shift;
    dst[4 + j] = (83 * o0 + 36 * o1 + add) >> shift;
    dst[12 + j] = (36 * o0 - 83 * o1 + add) >> shift;
  }
}

void oracle_dct4x4(int bitDepth, const int16_t* blk, int stride, int32_t* coeff) {
  const int shift1 = 2 + bitDepth + 6 - 15, shift2 = 2 + 6;            /* :925-926, g_maxTrDynamicRange = 15 */
  int32_t in[16], tmp[16]; int x, y;
  for (y = 0; y < 4; y++) for (x = 0; x < 4; x++) in[y * 4 + x] = blk[y * stride + x];
  dct4_stage(in, tmp, shift1);
  dct4_stage(tmp, coeff, shift2);
}

/* TCM: TEncSlice.cpp:202-392.  Works from the histogram of |c| (the reference builds the same
 * histogram from the coefficient list at :315-333). */
static double tcm_lambda(double yc, double sumYi, double total) {      /* :202-230 */
  const double c = sumYi / total;
  double lam, old; int k;
  if (c / yc >= 0.95) return -1.0;
  old = c;
  lam = c - yc * (1.0 - 1.0 / (1.0 - exp(-yc / old)));
  for (k = 0; k < 5; k++) { old = lam; lam = c - yc * (1.0 - 1.0 / (1.0 - exp(-yc / old))); }
  while (fabs(lam - old) > 0.1) { old = lam; lam = c - yc * (1.0 - 1.0 / (1.0 - exp(-yc / old))); }
  return lam;
}

double oracle_tcm_yc(const uint32_t* count, int peak, int len) {
  double *accAmp, *accNum, best, like, ycBest;
  int k, start, N = len;
  if (peak == 0 || peak >= 65536) return 0.0;                           /* :304-313 */
  accAmp = (double*)calloc((size_t)peak + 1, sizeof(double));
  accNum = (double*)calloc((size_t)peak + 1, sizeof(double));
  accNum[0] = count[0];
  for (k = 1; k <= peak; k++) { accAmp[k] = accAmp[k - 1] + (double)k * count[k]; accNum[k] = accNum[k - 1] + count[k]; }
  /* FindStartPoint :233-256 */
  for (k = peak; k > 0; k--) { if (count[k] == 0) continue; if (accNum[k] < N * (1.0 - 0.1)) break; }
  if (peak >= 3 && (int)count[0] > N / 100 && (int)count[1] > N / 100 && (int)count[2] > N / 100 && (int)count[3] > N / 100) { if (k < 3) k = 3; }
  else if (peak >= 2 && (int)count[0] > N / 100 && (int)count[1] > N / 100 && (int)count[2] > N / 100) { if (k < 2) k = 2; }
  else { if (k < 1) k = 1; }
  start = k;
  best = 0; ycBest = 0;
  for (k = start; k <= peak; k++) {                                     /* :356-374 */
    double n1, n2, yc, sumYi, lam, prob;
    if (k != start && count[k] == 0) continue;
    n1 = accNum[k]; n2 = N - n1; yc = k; sumYi = accAmp[k];
    lam = tcm_lambda(yc, sumYi, n1);                                    /* ComputeLikelyhood :258-289 */
    prob = n1 / (double)N;
    if (lam > 0)
      like = n2 * log(1 - prob) + n1 * log(prob) - n2 * log((peak - yc) * 2.0) - n1 * log(1 - exp(-yc / lam)) - n1 * log(2 * lam) - sumYi / lam;
    else
      like = 1.e30;                                                     /* -MinLikelyhood :284 */
    if (k == start || like > best) { best = like; ycBest = k; }
  }
  free(accAmp); free(accNum);
  return best > -1.e30 ? ycBest : 0.0;                                  /* :376-389 */
}

void oracle_outlier_frame(int bitDepth, const int16_t* org, int os, int W, int H, int16_t* obf, int16_t* outlier, double* yc) {
  const int bw = W / 4, bh = H / 4, nblk = bw * bh;
  int32_t* coeff = (int32_t*)malloc((size_t)nblk * 16 * sizeof(int32_t));
  uint32_t* hist = (uint32_t*)malloc(65536 * sizeof(uint32_t));
  int bx, by, f, i;
  for (by = 0; by < bh; by++) for (bx = 0; bx < bw; bx++) oracle_dct4x4(bitDepth, org + (by * 4) * os + bx * 4, os, coeff + (size_t)(by * bw + bx) * 16);
  yc[0] = 0;
  for (f = 1; f < 16; f++) {                                            /* :998-1004 on CoeffFrequency = coeff/8.0 truncated (:962) */
    int peak = 0;
    memset(hist, 0, 65536 * sizeof(uint32_t));
    for (i = 0; i < nblk; i++) { int a = iabs(coeff[(size_t)i * 16 + f] / 8); if (a > 65535) a = 65535; if (a > peak) peak = a; hist[a]++; }
    yc[f] = oracle_tcm_yc(hist, peak, nblk);
  }
  for (by = 0; by < bh; by++) for (bx = 0; bx < bw; bx++) {
    const int32_t* c = coeff + (size_t)(by * bw + bx) * 16;
    int cnt = 0;
    for (f = 0; f < 16; f++) {
      int32_t v = 0;
      if (f > 0) {                                                      /* DC dropped :987-988 */
        const double t = yc[f] * 8.0;
        v = c[f];
        if ((double)v < t && (double)v > -t) v = 0;                     /* :1010 */
        if (v != 0) cnt++;                                              /* :1027-1035 */
      }
      {                                                                 /* :1106-1112, coefficient domain, Pel = short */
        int16_t p = (int16_t)v;
        if (p < 0) p = (int16_t)-p;
        outlier[(by * 4 + f / 4) * W + bx * 4 + (f & 3)] = (int16_t)(p / 100);
      }
    }
    obf[by * bw + bx] = (int16_t)cnt;
  }
  /* samples right of / below the last whole 4x4 block are not written by the reference (:930-931) */
  for (by = 0; by < H; by++) for (bx = 0; bx < W; bx++) if (bx >= bw * 4 || by >= bh * 4) outlier[by * W + bx] = 0;
  free(coeff); free(hist);
}

void oracle_cu_sums(const int16_t* obf, int W, int H, int depth, int32_t* numObf, int32_t* nOutlier) {
  const int size = 64 >> depth, cells = size / 4, cw = W / size, ch = H / size, bw = W / 4;
  int cx, cy, x, y;
  for (cy = 0; cy < ch; cy++) for (cx = 0; cx < cw; cx++) {
    int32_t num = 0, sum = 0;
    for (y = 0; y < cells; y++) for (x = 0; x < cells; x++) {          /* TEncCu.cpp:589-598 */
      const int v = obf[(cy * cells + y) * bw + cx * cells + x];
      if (v > 0) { num++; sum += v; }
    }
    numObf[cy * cw + cx] = num; nOutlier[cy * cw + cx] = sum;
  }
}

void oracle_ctu_src_had(const int16_t* org, int os, int W, int H, int32_t* perCtu) {
  const int wc = (W + 63) >> 6, hc = (H + 63) >> 6;
  static const int16_t zero[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int cx, cy, x, y, i, j;
  for (cy = 0; cy < hc; cy++) for (cx = 0; cx < wc; cx++) {
    const int w = W - cx * 64 < 64 ? W - cx * 64 : 64, h = H - cy * 64 < 64 ? H - cy * 64 : 64;
    int32_t sum = 0;
    for (y = 0; y + 8 <= h; y += 8) for (x = 0; x + 8 <= w; x += 8) {  /* TEncCu.cpp:1874-1893 */
      const int16_t* p = org + (cy * 64 + y) * os + cx * 64 + x;
      int32_t m[64], tot = 0; int len;
      for (i = 0; i < 8; i++) for (j = 0; j < 8; j++) m[i * 8 + j] = p[i * os + j] - zero[j];
      for (i = 0; i < 8; i++) for (len = 1; len < 8; len <<= 1) for (j = 0; j < 8; j++) if (!(j & len)) { int32_t a = m[i * 8 + j], b = m[i * 8 + j + len]; m[i * 8 + j] = a + b; m[i * 8 + j + len] = a - b; }
      for (j = 0; j < 8; j++) for (len = 1; len < 8; len <<= 1) for (i = 0; i < 8; i++) if (!(i & len)) { int32_t a = m[i * 8 + j], b = m[(i + len) * 8 + j]; m[i * 8 + j] = a + b; m[(i + len) * 8 + j] = a - b; }
      for (i = 1; i < 64; i++) tot += iabs(m[i]);                       /* DC removed :1868 */
      sum += (tot + 2) >> 2;                                            /* :1869 */
    }
    perCtu[cy * wc + cx] = sum;
  }
}

/* ---------------------------------------------------------------------------------------------------
 * a12: TMVFeature / getTMVFeature (tools_YS.cpp:1659-1839, tools_YS.h:99-127).
 * Five 3x3 masks applied INSIDE the CU (the one-sample ring of every filtered plane stays 0, :1659-1672),
 * then for each plane 13 regions: mean = integer sum / rows / cols as double (tools_YS.h:106-115), "variance" =
 * mean of |(Int)(sample - mean)| (:117-126).  The four triangle entries keep the reference's quirks: the divisor
 * is N*N>>1 (not the triangle's sample count) and the deviation loop ASSIGNS instead of accumulating, so only the
 * last visited sample - always a ring sample, i.e. 0 - contributes (:1757-1816).
 * ------------------------------------------------------------------------------------------------- */
static double tmv_mean(const int16_t* b, int n, int i0, int i1, int j0, int j1) {
  int sum = 0, i, j;
  for (i = i0; i < i1; i++) for (j = j0; j < j1; j++) sum += b[i * n + j];
  return ((double)sum) / (i1 - i0) / (j1 - j0);
}
static double tmv_dev(const int16_t* b, double mean, int n, int i0, int i1, int j0, int j1) {
  int sum = 0, i, j;
  for (i = i0; i < i1; i++) for (j = j0; j < j1; j++) { const int t = (int)(b[i * n + j] - mean); sum += t > 0 ? t : -t; }
  return ((double)sum) / (i1 - i0) / (j1 - j0);
}
void oracle_tmv_features(const int16_t* cu, int stride, int n, double* feat) {
  static const int mask[5][9] = {{0, 0, 0, 0, 1, 0, 0, 0, 0}, {0, 0, 0, 1, 0, -1, 0, 0, 0}, {0, 1, 0, 0, 0, 0, 0, -1, 0},
                                 {0, 0, 1, 0, 0, 0, -1, 0, 0}, {1, 0, 0, 0, 0, 0, 0, 0, -1}};
  int16_t* f = (int16_t*)malloc((size_t)n * n * sizeof(int16_t));
  const int h = n >> 1;
  const unsigned tri = ((unsigned)n * (unsigned)n) >> 1;
  int k, x, y, i, j;
  for (k = 0; k < 5; k++) {
    double* d = feat + k * 26;
    double m, v;
    memset(f, 0, (size_t)n * n * sizeof(int16_t));
    for (y = 1; y < n - 1; y++) for (x = 1; x < n - 1; x++)
      for (i = 0; i < 3; i++) for (j = 0; j < 3; j++) f[y * n + x] = (int16_t)(f[y * n + x] + mask[k][i * 3 + j] * cu[(y - 1 + i) * stride + (x - 1 + j)]);
    m = tmv_mean(f, n, 0, n, 0, n); d[0] = m; d[1] = tmv_dev(f, m, n, 0, n, 0, n);
    m = tmv_mean(f, n, 0, h, 0, n); d[2] = m; d[4] = tmv_dev(f, m, n, 0, h, 0, n);
    m = tmv_mean(f, n, h, n, 0, n); d[3] = m; d[5] = tmv_dev(f, m, n, h, n, 0, n);
    m = tmv_mean(f, n, 0, n, 0, h); d[6] = m; d[8] = tmv_dev(f, m, n, 0, n, 0, h);
    m = tmv_mean(f, n, 0, n, h, n); d[7] = m; d[9] = tmv_dev(f, m, n, 0, n, h, n);
#define TRI(JLO, JHI, MI, VI)                                                                      \
    m = 0; for (i = 0; i < n; i++) for (j = (JLO); j < (JHI); j++) m += f[i * n + j];              \
    m /= tri; d[MI] = m; v = 0;                                                                    \
    for (i = 0; i < n; i++) for (j = (JLO); j < (JHI); j++) { const int t = (int)(f[i * n + j] - m); v = t > 0 ? t : -t; } \
    v /= tri; d[VI] = v;
    TRI(0, n - i, 10, 12)
    TRI(n - i - 1, n, 11, 13)
    TRI(i, n, 14, 16)
    TRI(0, i + 1, 15, 17)
#undef TRI
    m = tmv_mean(f, n, 0, h, 0, h); d[18] = m; d[22] = tmv_dev(f, m, n, 0, h, 0, h);
    m = tmv_mean(f, n, 0, h, h, n); d[19] = m; d[23] = tmv_dev(f, m, n, 0, h, h, n);
    m = tmv_mean(f, n, h, n, 0, n); d[20] = m; d[24] = tmv_dev(f, m, n, h, n, 0, n);       /* sic: columns 0..n (:1830-1832) */
    m = tmv_mean(f, n, h, n, h, n); d[21] = m; d[25] = tmv_dev(f, m, n, h, n, h, n);
  }
  free(f);
}

/* ---------------------------------------------------------------------------------------------------
 * a13: TEncPreanalyzer::xPreanalyze (TEncPreanalyzer.cpp:64-139) for ONE AQ layer: units of part x part samples
 * (clipped at the picture edge), four quadrants split at half of the CLIPPED unit size, activity = 1 + min variance
 * where both moments are divided by the pixel count of the whole unit (:128-131).  Returns the layer average (:137).
 * activity: ceil(W/part) x ceil(H/part), raster order.
 * ------------------------------------------------------------------------------------------------- */
double oracle_aq_activity(const int16_t* org, int stride, int W, int H, int part, double* activity) {
  const int nw = (W + part - 1) / part, nh = (H + part - 1) / part;
  double sumAct = 0.0;
  int ux, uy, q;
  for (uy = 0; uy < nh; uy++) for (ux = 0; ux < nw; ux++) {
    const int x0 = ux * part, y0 = uy * part;
    const int w = W - x0 < part ? W - x0 : part, h = H - y0 < part ? H - y0 : part;
    uint64_t s[4] = {0, 0, 0, 0}, ss[4] = {0, 0, 0, 0};
    unsigned npix = 0;
    double minVar = 1.7976931348623157e308;
    int bx, by;
    for (by = 0; by < h; by++) for (bx = 0; bx < w; bx++, npix++) {
      const int p = org[(y0 + by) * stride + x0 + bx];
      q = (by < (h >> 1) ? 0 : 2) + (bx < (w >> 1) ? 0 : 1);
      s[q] += (uint64_t)(int64_t)p; ss[q] += (uint64_t)(int64_t)(p * p);
    }
    for (q = 0; q < 4; q++) {
      const double avg = (double)s[q] / npix;
      const double var = (double)ss[q] / npix - avg * avg;
      if (var < minVar) minVar = var;
    }
    activity[uy * nw + ux] = 1.0 + minVar;
    sumAct += 1.0 + minVar;
  }
  return sumAct / (double)(unsigned)(nw * nh);
}

/* ===================================================================================================
 * SURVEY.md 8f.2 - intra luma TU coding: xIntraCodingTUBlock TEncSearch.cpp:1092-1387
 * =================================================================================================== */

/* The HEVC core transform matrices g_aiT4/8/16/32 (TComRom.cpp:356-460, 6-bit precision): row k of the N-point matrix is
 * round(64*sqrt(2)*cos(k(2n+1)pi/2N)) with the standard's hand-tuned values; all of them are sub-sampled rows of the
 * 32-point matrix, whose 31 distinct magnitudes are c[m] ~ 64*sqrt(2)*cos(m*pi/64). */
static const int kDctC[33] = {64, 90, 90, 90, 89, 88, 87, 85, 83, 82, 80, 78, 75, 73, 70, 67, 64,
                              61, 57, 54, 50, 46, 43, 38, 36, 31, 25, 22, 18, 13, 9, 4, 0};
int oracle_dct_coef(int n, int k, int x) {         /* g_aiT<n>[TRANSFORM_FORWARD][k][x] */
  int a;
  if (k == 0) return 64;
  a = (k * (32 / n) * (2 * x + 1)) & 127;           /* angle in units of pi/64 */
  if (a > 64) a = 128 - a;
  return a <= 32 ? kDctC[a] : -kDctC[64 - a];
}
static const int kDst4[4][4] = {{29, 55, 74, 84}, {74, 74, 0, -74}, {84, -29, -74, 55}, {55, -84, 74, -29}};   /* TComRom.cpp:343-349 */
static int tr_coef(int n, int useDST, int k, int x) { return useDST ? kDst4[k][x] : oracle_dct_coef(n, k, x); }

/* TComTrQuant::xT -> xTrMxN (TComTrQuant.cpp:860-919, 1857-1882): the partial butterflies are exact integer matrix products,
 * dst[k*line + j] = (sum_x T[k][x] * src[j*N + x] + add) >> shift, twice (rows, then columns) */
void oracle_fwd_transform(int bitDepth, int n, int useDST, const int16_t* resi, int stride, int32_t* coeff) {
  const int lg = ilog2(n), shift1 = lg + bitDepth + 6 - 15, shift2 = lg + 6;
  const int add1 = shift1 > 0 ? 1 << (shift1 - 1) : 0, add2 = 1 << (shift2 - 1);
  int32_t tmp[32 * 32];
  int j, k, x;
  for (j = 0; j < n; j++) for (k = 0; k < n; k++) {
    int32_t s = 0;
    for (x = 0; x < n; x++) s += tr_coef(n, useDST, k, x) * resi[j * stride + x];
    tmp[k * n + j] = (s + add1) >> shift1;
  }
  for (j = 0; j < n; j++) for (k = 0; k < n; k++) {
    int32_t s = 0;
    for (x = 0; x < n; x++) s += tr_coef(n, useDST, k, x) * tmp[j * n + x];
    coeff[k * n + j] = (s + add2) >> shift2;
  }
}
/* TComTrQuant::xIT -> xITrMxN (:927-985, 1892-1924): columns first (clip to 16 bit), then rows (clip to Pel) */
void oracle_inv_transform(int bitDepth, int n, int useDST, const int32_t* coeff, int16_t* resi, int stride) {
  const int shift1 = 7, shift2 = 20 - bitDepth;
  int32_t tmp[32 * 32];
  int j, k, x;
  for (j = 0; j < n; j++) for (x = 0; x < n; x++) {
    int32_t s = 0;
    for (k = 0; k < n; k++) s += tr_coef(n, useDST, k, x) * coeff[k * n + j];
    tmp[j * n + x] = clip3(-32768, 32767, (s + (1 << (shift1 - 1))) >> shift1);
  }
  for (j = 0; j < n; j++) for (x = 0; x < n; x++) {
    int32_t s = 0;
    for (k = 0; k < n; k++) s += tr_coef(n, useDST, k, x) * tmp[k * n + j];
    resi[j * stride + x] = (int16_t)clip3(-32768, 32767, (s + (1 << (shift2 - 1))) >> shift2);
  }
}
/* xTransformSkip / xITransformSkip (:1933-2031), no rotation, bit depths <= 10 (shift = 15 - bd - log2 n >= 0) */
void oracle_transform_skip(int bitDepth, int n, const int16_t* resi, int stride, int32_t* coeff) {
  const int sh = 15 - bitDepth - ilog2(n); int x, y;
  for (y = 0; y < n; y++) for (x = 0; x < n; x++) coeff[y * n + x] = (int32_t)resi[y * stride + x] << sh;
}
void oracle_inv_transform_skip(int bitDepth, int n, const int32_t* coeff, int16_t* resi, int stride) {
  const int sh = 15 - bitDepth - ilog2(n), off = sh == 0 ? 0 : 1 << (sh - 1); int x, y;
  for (y = 0; y < n; y++) for (x = 0; x < n; x++) resi[y * stride + x] = (int16_t)((coeff[y * n + x] + off) >> sh);
}

/* TComDataCU::getCoefScanIdx (TComDataCU.cpp:3356-3410) for intra blocks: 0 diagonal, 1 horizontal, 2 vertical; mode-dependent
 * scans up to 8x8 luma / 4x4 chroma (4:2:0), `mode` = the final prediction mode of the component */
int oracle_scan_idx_c(int n, int mode, int chroma) {
  if (n > (chroma ? 4 : 8)) return 0;
  if (iabs(mode - 26) <= 4) return 1;
  if (iabs(mode - 10) <= 4) return 2;
  return 0;
}
int oracle_scan_idx(int n, int mode) {
  if (n > 8) return 0;
  if (iabs(mode - 26) <= 4) return 1;
  if (iabs(mode - 10) <= 4) return 2;
  return 0;
}
/* g_scanOrder[SCAN_GROUPED_4x4][scanIdx][log2 n][log2 n] (TComRom.cpp:53-228): 4x4 coefficient groups visited in the
 * scan's order, the same scan inside each group; scan[pos] = raster index */
static void scan_next(int type, int w, int h, int* line, int* col) {
  if (type == 0) {
    if (*col == w - 1 || *line == 0) { *line += *col + 1; *col = 0; if (*line >= h) { *col += *line - (h - 1); *line = h - 1; } }
    else { (*col)++; (*line)--; }
  } else if (type == 1) { if (*col == w - 1) { (*line)++; *col = 0; } else (*col)++; }
  else { if (*line == h - 1) { (*col)++; *line = 0; } else (*line)++; }
}
void oracle_scan_order(int scanIdx, int n, uint16_t* scan) {
  const int g = n >> 2;
  int gl = 0, gc = 0, gi, p;
  for (gi = 0; gi < g * g; gi++) {
    int l = 0, c = 0;
    for (p = 0; p < 16; p++) { scan[gi * 16 + p] = (uint16_t)((gl * 4 + l) * n + gc * 4 + c); scan_next(scanIdx, 4, 4, &l, &c); }
    scan_next(scanIdx, g, g, &gl, &gc);
  }
}

static const int kQuantScales[6] = {26214, 23302, 20560, 18396, 16384, 14564};   /* TComRom.cpp:328-336 */
static const int kInvQuantScales[6] = {40, 45, 51, 57, 64, 72};

/* TComTrQuant::xQuant without RDOQ (:1126-1240, flat scaling) + signBitHidingHDQ (:991-1123). qp = CU luma QP. Returns uiAbsSum. */
int oracle_quant(int bitDepth, int n, int qp, int intraSlice, int signHiding, int scanIdx, const int32_t* coef, int32_t* level) {
  const int baseQp = qp + 6 * (bitDepth - 8), per = baseQp / 6, rem = baseQp % 6;
  const int tshift = 15 - bitDepth - ilog2(n), qbits = 14 + per + tshift;
  const int64_t add = (int64_t)(intraSlice ? 171 : 85) << (qbits - 9);
  const int scale = kQuantScales[rem], total = n * n;
  int32_t deltaU[32 * 32];
  uint16_t scan[32 * 32];
  int absSum = 0, i;
  for (i = 0; i < total; i++) {
    const int64_t t = (int64_t)iabs(coef[i]) * scale;
    const int32_t q = (int32_t)((t + add) >> qbits);
    deltaU[i] = (int32_t)((t - ((int64_t)q << qbits)) >> (qbits - 8));
    absSum += q;
    level[i] = clip3(-32768, 32767, coef[i] < 0 ? -q : q);
  }
  if (signHiding && absSum >= 2) {
    int lastCG = -1, subSet;
    oracle_scan_order(scanIdx, n, scan);
    for (subSet = (total - 1) >> 4; subSet >= 0; subSet--) {
      const int subPos = subSet << 4;
      int firstNZ = 16, lastNZ = -1, sum = 0, k;
      for (k = 15; k >= 0; k--) if (level[scan[k + subPos]]) { lastNZ = k; break; }
      for (k = 0; k < 16; k++) if (level[scan[k + subPos]]) { firstNZ = k; break; }
      for (k = firstNZ; k <= lastNZ; k++) sum += level[scan[k + subPos]];
      if (lastNZ >= 0 && lastCG == -1) lastCG = 1;
      if (lastNZ - firstNZ >= 4) {
        const int signbit = level[scan[subPos + firstNZ]] > 0 ? 0 : 1;
        if (signbit != (sum & 1)) {
          int32_t curCost = 0x7fffffff, minCostInc = 0x7fffffff;
          int minPos = -1, finalChange = 0, curChange = 0;
          for (k = (lastCG == 1 ? lastNZ : 15); k >= 0; k--) {
            const int blk = scan[k + subPos];
            if (level[blk] != 0) {
              if (deltaU[blk] > 0) { curCost = -deltaU[blk]; curChange = 1; }
              else if (k == firstNZ && iabs(level[blk]) == 1) curCost = 0x7fffffff;
              else { curCost = deltaU[blk]; curChange = -1; }
            } else if (k < firstNZ) {
              const int thisSign = coef[blk] >= 0 ? 0 : 1;
              if (thisSign != signbit) curCost = 0x7fffffff;
              else { curCost = -deltaU[blk]; curChange = 1; }
            } else { curCost = -deltaU[blk]; curChange = 1; }
            if (curCost < minCostInc) { minCostInc = curCost; finalChange = curChange; minPos = blk; }
          }
          if (level[minPos] == 32767 || level[minPos] == -32768) finalChange = -1;
          if (coef[minPos] >= 0) level[minPos] += finalChange; else level[minPos] -= finalChange;
        }
      }
      if (lastCG == 1) lastCG = 0;
    }
  }
  return absSum;
}
/* TComTrQuant::xDeQuant, flat scaling (:1242-1352) */
void oracle_dequant(int bitDepth, int n, int qp, const int32_t* level, int32_t* coef) {
  const int baseQp = qp + 6 * (bitDepth - 8), per = baseQp / 6, rem = baseQp % 6;
  const int tshift = 15 - bitDepth - ilog2(n), rightShift = 6 - (tshift + per), scale = kInvQuantScales[rem];
  const int bitsIn = 16 < 32 + rightShift - 7 ? 16 : 32 + rightShift - 7;
  const int inMin = -(1 << (bitsIn - 1)), inMax = (1 << (bitsIn - 1)) - 1;
  int i;
  for (i = 0; i < n * n; i++) {
    const int32_t q = clip3(inMin, inMax, level[i]);
    const int32_t v = rightShift > 0 ? (q * scale + (1 << (rightShift - 1))) >> rightShift : (int32_t)((uint32_t)(q * scale) << -rightShift);
    coef[i] = clip3(-32768, 32767, v);
  }
}
/* TComRdCost::xGetSSE* (TComRdCost.cpp:970-1315) */
uint32_t oracle_sse(int bitDepth, const int16_t* org, int os, const int16_t* cur, int cs, int w, int h) {
  const int sh = (bitDepth - 8) << 1; uint32_t sum = 0; int x, y;
  for (y = 0; y < h; y++) for (x = 0; x < w; x++) { const int d = org[y * os + x] - cur[y * cs + x]; sum += (uint32_t)((d * d) >> sh); }
  return sum;
}

/* One luma TU through xIntraCodingTUBlock. stage: 0 = forward only (prediction + residual + transform; `coef` and `pred` out),
 * 1 = whole chain with the plain quantiser (level/reco/dist/absSum out), 2 = reconstruction from GIVEN levels (level in). */
void oracle_intra_tu(int bitDepth, int n, int mode, int qp, int transformSkip, int strongSmoothing, int intraSlice, int signHiding, int stage,
                     const int16_t* org, int orgStride, const int16_t* border, int32_t* coef, int32_t* level, int16_t* pred, int16_t* reco,
                     uint32_t* dist, int32_t* absSum) {
  oracle_intra_tu_c(bitDepth, n, mode, qp, transformSkip, 0, strongSmoothing, intraSlice, signHiding, stage, org, orgStride, border, coef, level, pred, reco, dist, absSum);
}
/* the same for a chroma block of a 4:2:0 picture when chroma != 0: qp = the component's mapped QP minus the bit-depth offset
 * (QpParam, TComTrQuant.cpp:66-118), the distortion is returned before the chroma weight (TComRdCost.cpp:447-450) */
void oracle_intra_tu_c(int bitDepth, int n, int mode, int qp, int transformSkip, int chroma, int strongSmoothing, int intraSlice, int signHiding, int stage,
                       const int16_t* org, int orgStride, const int16_t* border, int32_t* coef, int32_t* level, int16_t* pred, int16_t* reco,
                       uint32_t* dist, int32_t* absSum) {
  int16_t fil[4 * 32 + 1], p[32 * 32], resi[32 * 32];
  int32_t c[32 * 32], deq[32 * 32];
  const int useDST = n == 4 && !chroma;                                 /* TComTU::useDST: intra luma 4x4 */
  int i, x, y, sum = 0;
  if (chroma) oracle_predict_chroma(bitDepth, n, mode, border, p);
  else { oracle_filter_border(bitDepth, n, strongSmoothing, border, fil); oracle_predict(bitDepth, n, mode, border, fil, p); }
  if (pred) memcpy(pred, p, (size_t)n * n * sizeof(int16_t));
  if (stage != 2) {
    for (y = 0; y < n; y++) for (x = 0; x < n; x++) resi[y * n + x] = (int16_t)(org[y * orgStride + x] - p[y * n + x]);   /* :1207-1224 */
    if (transformSkip) oracle_transform_skip(bitDepth, n, resi, n, c); else oracle_fwd_transform(bitDepth, n, useDST, resi, n, c);
    if (coef) memcpy(coef, c, (size_t)n * n * sizeof(int32_t));
    if (stage == 0) return;
    sum = oracle_quant(bitDepth, n, qp, intraSlice, signHiding, oracle_scan_idx_c(n, mode, chroma), c, level);
  } else {
    for (i = 0; i < n * n; i++) sum += iabs(level[i]);                  /* only its being non-zero matters (:1271-1291) */
  }
  if (absSum) *absSum = sum;
  if (sum > 0) {
    oracle_dequant(bitDepth, n, qp, level, deq);
    if (transformSkip) oracle_inv_transform_skip(bitDepth, n, deq, resi, n); else oracle_inv_transform(bitDepth, n, useDST, deq, resi, n);
  } else {
    if (stage == 1) memset(level, 0, (size_t)n * n * sizeof(int32_t));
    memset(resi, 0, sizeof(resi));
  }
  for (i = 0; i < n * n; i++) reco[i] = (int16_t)clip3(0, (1 << bitDepth) - 1, p[i] + resi[i]);   /* :1360-1381 */
  *dist = oracle_sse(bitDepth, reco, n, org, orgStride, n, n);                                     /* :1385-1386 */
}

/* ===================================================================================================
 * SURVEY.md 8f.3 - fractional-pel motion-estimation refinement: xPatternSearchFracDIF TEncSearch.cpp:4340-4376,
 * xExtDIFUpSamplingH / Q :5431-5637, xPatternRefinement :808-865, TComInterpolationFilter.cpp:57-290.
 * Every candidate block of the half- and quarter-pel refinement is the separable 8-tap luma interpolation of the reference at
 * that quarter-pel position: rows first into 14-bit intermediates (filterHor, isFirst, !isLast), then columns (filterVer,
 * !isFirst, isLast) with rounding and clipping - also for zero fractions, where the "filter" is filterCopy.
 * =================================================================================================== */
static const int kLumaFilter[4][8] = {{0, 0, 0, 64, 0, 0, 0, 0}, {-1, 4, -10, 58, 17, -5, 1, 0}, {-1, 4, -11, 40, 40, -11, 4, -1}, {0, 1, -5, 17, 58, -10, 4, -1}};
/* ref points at the block's top-left sample at the INTEGER part of the position; fx, fy = quarter-pel fractions 0..3 */
void oracle_interp_luma(int bitDepth, const int16_t* ref, int stride, int w, int h, int fx, int fy, int16_t* out) {
  const int head = 14 - bitDepth;                       /* IF_INTERNAL_PREC - bitDepth (>= 2 for bit depths <= 12) */
  int16_t tmp[(64 + 7) * 64];
  int r, c, k;
  for (r = 0; r < h + 7; r++) for (c = 0; c < w; c++) {               /* rows -3 .. h+3 */
    const int16_t* s = ref + (r - 3) * stride + c;
    int v;
    if (fx == 0) v = (s[0] << head) - 8192;                           /* filterCopy, isFirst */
    else { int sum = 0; for (k = 0; k < 8; k++) sum += s[k - 3] * kLumaFilter[fx][k]; v = (sum - (8192 << (6 - head))) >> (6 - head); }
    tmp[r * w + c] = (int16_t)v;
  }
  for (r = 0; r < h; r++) for (c = 0; c < w; c++) {
    int v;
    if (fy == 0) v = (tmp[(r + 3) * w + c] + 8192 + (1 << (head - 1))) >> head;      /* filterCopy, isLast */
    else { int sum = 0; for (k = 0; k < 8; k++) sum += tmp[(r + k) * w + c] * kLumaFilter[fy][k]; v = (sum + (1 << (6 + head - 1)) + (8192 << 6)) >> (6 + head); }
    out[r * w + c] = (int16_t)clip3(0, (1 << bitDepth) - 1, v);
  }
}
/* distortion of one candidate: quarter-pel MV (qx, qy) relative to refAtZeroMv; Hadamard (xGetHADs) or SAD (xGetSAD, no sub-sampling) */
uint32_t oracle_subpel_cost(int bitDepth, const int16_t* org, int orgStride, int w, int h, const int16_t* refAtZeroMv, int refStride,
                            int qx, int qy, int useHadamard) {
  int16_t pred[64 * 64];
  oracle_interp_luma(bitDepth, refAtZeroMv + (qy >> 2) * refStride + (qx >> 2), refStride, w, h, qx & 3, qy & 3, pred);
  if (useHadamard) return oracle_satd(bitDepth, org, orgStride, pred, w, w, h);
  { uint32_t sum = 0; int x, y; for (y = 0; y < h; y++) for (x = 0; x < w; x++) sum += (uint32_t)iabs(org[y * orgStride + x] - pred[y * w + x]); return sum >> (bitDepth - 8); }
}
/* all 49 candidates within +-3 quarter-pels of the integer MV (mvx, mvy): out[(dy+3)*7 + (dx+3)] */
void oracle_subpel_surface(int bitDepth, const int16_t* org, int orgStride, int w, int h, const int16_t* refAtZeroMv, int refStride,
                           int mvx, int mvy, int useHadamard, uint32_t* out) {
  int dx, dy;
  for (dy = -3; dy <= 3; dy++) for (dx = -3; dx <= 3; dx++)
    out[(dy + 3) * 7 + dx + 3] = oracle_subpel_cost(bitDepth, org, orgStride, w, h, refAtZeroMv, refStride, 4 * mvx + dx, 4 * mvy + dy, useHadamard);
}


/* ---- fork-aware enumeration (TEST INFRASTRUCTURE like everything in this file) ---------------------------------------------------
 * A recursive walk that mirrors TEncCu::xCompressCU for an intra slice in the Testing state:
 *   xCompressCU(CU, depth):                                                        TEncCu.cpp
 *     bBoundary = CU not completely inside the picture                             :488-489
 *     if !bBoundary: Num_OBF over the CU's OBF cells                               :589-600
 *                    Naive model: Num_OBF == 0 -> TerminateCU else Skip2Nx2N       tools_YS.cpp:686-695, TEncCu.cpp:645-675
 *                    bEarlyTerminate / bSkip2Nx2N gated by the per-depth switches  :973-980
 *                    intra 2Nx2N unless bSkip2Nx2N                                 :1040
 *                    depth 3: intra NxN unless bEarlyTerminate                     :1140-1143
 *     sub-CUs unless bEarlyTerminate (boundary CUs always recurse)                 :1257-1260, 1290-1330 */
static int pu_index_in_ctu(int depth, int xInCtu, int yInCtu) {
  static const int first[5] = {0, 1, 5, 21, 85};
  const int size = 64 >> depth, px = xInCtu / size, py = yInCtu / size;
  int z = 0, b;
  for (b = 0; b < 4; b++) z |= ((px >> b) & 1) << (2 * b) | ((py >> b) & 1) << (2 * b + 1);
  return first[depth] + z;
}
static void prune_walk(const int16_t* obf, int W, int H, const uint8_t* swSkip, const uint8_t* swTerm, int depth, int x, int y, uint8_t* ctuMask) {
  const int size = 64 >> depth, bw = W / 4;
  int term = 0, q;
  if (x >= W || y >= H) return;                                  /* the recursion only enters sub-CUs whose origin is inside the picture */
  if (x + size <= W && y + size <= H) {
    int num = 0, cx, cy, skip;
    for (cy = 0; cy < size / 4; cy++) for (cx = 0; cx < size / 4; cx++) if (obf[(y / 4 + cy) * bw + x / 4 + cx] > 0) num++;
    term = swTerm[depth] && num == 0;
    skip = swSkip[depth] && num > 0;
    if (!skip) ctuMask[pu_index_in_ctu(depth, x & 63, y & 63)] = 1;
    if (depth == 3 && !term) for (q = 0; q < 4; q++) ctuMask[pu_index_in_ctu(4, (x & 63) + 4 * (q & 1), (y & 63) + 4 * (q >> 1))] = 1;
  }
  if (depth == 3 || term) return;
  for (q = 0; q < 4; q++) prune_walk(obf, W, H, swSkip, swTerm, depth + 1, x + (q & 1) * (size / 2), y + (q >> 1) * (size / 2), ctuMask);
}
void oracle_prune_mask(const int16_t* obf, int W, int H, const uint8_t* swSkip2Nx2N, const uint8_t* swTerminateCU, uint8_t* needed) {
  const int wc = (W + 63) >> 6, hc = (H + 63) >> 6;
  int cx, cy;
  memset(needed, 0, (size_t)wc * hc * 341);
  for (cy = 0; cy < hc; cy++) for (cx = 0; cx < wc; cx++) prune_walk(obf, W, H, swSkip2Nx2N, swTerminateCU, 0, cx * 64, cy * 64, needed + (size_t)(cy * wc + cx) * 341);
}
