// rmd_core.cuh - per-thread arithmetic of the intra rough-mode-decision (RMD) kernels.
//
// Everything in this header is `__host__ __device__`: the CUDA kernels in rmd_kernels.cu call it
// from sm_100a threads, and tests/emul/ compiles the very same code with g++ to replay a CTA's
// phases lane by lane on the CPU (there is no GPU in the build container), so arithmetic bugs are
// found before GPU time is spent.  The emulation harness is a test, not a product path.
//
// What is computed (reference file:line, see DESIGN.md for the mapping):
//   reference-sample substitution + [1 2 1]/strong smoothing   TComPattern.cpp:185-283, 314-521
//   which border a mode reads                                   TComPattern.cpp:523-548
//   planar / DC(+edge) / 33 angular predictions                 TComPrediction.cpp:183-222,250-410,755-841
//   8x8 / 4x4 Hadamard SATD with HM's per-tile rounding         TComRdCost.cpp:1343-1604
//
// Data model on the SM
//   * a work item ("chunk") is 4096 luma samples = 64 tiles of 8x8 that belong to 4096/N^2 PUs of
//     one size N (a CTU at one depth in frame mode, 4096/N^2 host-submitted PUs in batch mode);
//   * a lane owns ONE 8x8 tile (N>=8) or one 8x8 region = four 4x4 PUs (N=4) for the whole chunk and
//     keeps its source samples in 32 registers as packed int16 pairs; the mode is warp-uniform;
//   * predictions never leave registers: residual pairs are formed as 32-bit differences of packed
//     words (lo + 65536*hi as one integer, the borrow is undone when the halves are separated), so a
//     Hadamard butterfly on two coefficients is one IADD/ISUB;  5 of the 6 stages run packed, the
//     6th (between the halves of a word) is folded into |a+b|+|a-b| = 2*max(|a|,|b|);
//   * horizontal-class modes (2..17) are evaluated on the transposed tile: SATD is transpose
//     invariant, the lane keeps a transposed copy of its source tile instead.
// Valid for bit depths 8..10 (5 packed stages need 32*(2^bd-1) < 32768).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define CUCD_HD __host__ __device__ __forceinline__
#else
#define CUCD_HD inline
#endif

namespace cucd {

// ---------------------------------------------------------------------------------------------
// small portable intrinsics
// ---------------------------------------------------------------------------------------------
CUCD_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t s) {   // s in [0,31]
#if defined(__CUDA_ARCH__)
  return __funnelshift_r(lo, hi, s);
#else
  return s ? ((lo >> s) | (hi << (32 - s))) : lo;
#endif
}
CUCD_HD uint32_t funnel_rc(uint32_t lo, uint32_t hi, uint32_t s) {  // s in [0,32], clamped
#if defined(__CUDA_ARCH__)
  return __funnelshift_rc(lo, hi, s);
#else
  return s >= 32 ? hi : (s ? ((lo >> s) | (hi << (32 - s))) : lo);
#endif
}
CUCD_HD uint32_t pack_lo(uint32_t a, uint32_t b) {                  // (a.lo, b.lo)
#if defined(__CUDA_ARCH__)
  return __byte_perm(a, b, 0x5410);
#else
  return (a & 0xffffu) | (b << 16);
#endif
}
CUCD_HD uint32_t pack_hi(uint32_t a, uint32_t b) {                  // (a.hi, b.hi)
#if defined(__CUDA_ARCH__)
  return __byte_perm(a, b, 0x7632);
#else
  return (a >> 16) | (b & 0xffff0000u);
#endif
}
CUCD_HD int iabs32(int v) { return v < 0 ? -v : v; }
CUCD_HD int imax32(int a, int b) { return a > b ? a : b; }
CUCD_HD int imin32(int a, int b) { return a < b ? a : b; }

// ---------------------------------------------------------------------------------------------
// mode tables
// ---------------------------------------------------------------------------------------------
// signed 1/32-sample displacement per row of an angular mode (TComPrediction.cpp:280-291)
// The two 9-entry tables {0,2,5,9,13,17,21,26,32} and {0,4096,1638,910,630,482,390,315,256} are packed
// into 64-bit immediates (8 bits / 16 bits per entry): a table in local memory would cost a store and a
// load per lookup inside the mode loop.
CUCD_HD int mode_angle(int mode) {
  const int am = mode >= 18 ? mode - 26 : 10 - mode;
  const int a = am < 0 ? -am : am;
  const unsigned long long lo = 0x1a15110d09050200ull;        // entries 0..7
  const int v = a == 8 ? 32 : (int)((lo >> (8 * a)) & 0xffu);
  return am < 0 ? -v : v;
}
CUCD_HD int mode_inv_angle(int mode) {
  const int am = mode >= 18 ? mode - 26 : 10 - mode;
  const int a = am < 0 ? -am : am;
  const unsigned long long t0 = 0x038e066610000000ull;        // entries 0..3: 0, 4096, 1638, 910
  const unsigned long long t1 = 0x013b018601e20276ull;        // entries 4..7: 630, 482, 390, 315
  if (a == 8) return 256;
  return (int)(((a < 4 ? t0 : t1) >> (16 * (a & 3))) & 0xffffu);
}
// TComPattern.cpp:523-548 with the luma row of m_aucIntraFilter (TComPrediction.cpp:50-67)
template <int LOG2N>
CUCD_HD bool mode_uses_filtered(int mode) {
  if (LOG2N == 2 || LOG2N == 6 || mode == 1) return false;
  const int thr = LOG2N == 3 ? 7 : (LOG2N == 4 ? 1 : 0);
  const int d10 = mode > 10 ? mode - 10 : 10 - mode, d26 = mode > 26 ? mode - 26 : 26 - mode;
  return (d10 < d26 ? d10 : d26) > thr;
}

CUCD_HD bool mode_uses_filtered_rt(int log2n, int mode) {
  if (log2n == 2 || log2n == 6 || mode == 1) return false;
  const int thr = log2n == 3 ? 7 : (log2n == 4 ? 1 : 0);
  const int d10 = mode > 10 ? mode - 10 : 10 - mode, d26 = mode > 26 ? mode - 26 : 26 - mode;
  return (d10 < d26 ? d10 : d26) > thr;
}

// ---------------------------------------------------------------------------------------------
// shared-memory geometry of one chunk
// ---------------------------------------------------------------------------------------------
template <int LOG2N>
struct Geo {
  static constexpr int N = 1 << LOG2N;
  static constexpr int PUS = 4096 / (N * N);            // PUs per chunk
  static constexpr int LIN = 4 * N + 2;                 // linear border (4N+1) padded to even
  static constexpr int AS = 2 * N + 6;                  // one ascending ref array: corner + 2N samples + over-read pad
  static constexpr bool HAS_FILT = (LOG2N >= 3 && LOG2N <= 5);
  static constexpr int NARR = HAS_FILT ? 4 : 2;         // Tu, Lu [, Tf, Lf]
  // int16 per PU, padded so that the stride is an ODD number of 32-bit words: lanes that read the same
  // offset of different PUs then hit 32 different shared-memory banks
  static constexpr int PU_STRIDE = (NARR * AS) + ((((NARR * AS) / 2) & 1) ? 0 : 2);
  static constexpr int TILES_PER_PU = (N >= 8) ? (N / 8) * (N / 8) : 1;
  static constexpr int PUS_PER_WARP = (N >= 8) ? ((32 / TILES_PER_PU) > 0 ? (32 / TILES_PER_PU) : 1) : 128;
  static constexpr int XS = 2 * N + 2;                  // extended ref array [-N .. N+1]: N+1 words, an odd count for every N
  static constexpr int EXT_PER_WARP = PUS_PER_WARP * XS;
};

// Where PU p of a chunk keeps its arrays.  For N = 4 a lane owns PUs 4t..4t+3; storing them
// sub-PU-major (slot = (p & 3) * 64 + (p >> 2)) makes the lane stride one PU_STRIDE (odd words) again.
template <int LOG2N>
CUCD_HD int pu_slot(int p) { return LOG2N == 2 ? ((p & 3) * 64 + (p >> 2)) : p; }
CUCD_HD int pu_slot_rt(int log2n, int p) { return log2n == 2 ? ((p & 3) * 64 + (p >> 2)) : p; }

// ---------------------------------------------------------------------------------------------
// border construction (phases; `tid`/`nthreads` make them replayable on the host)
// ---------------------------------------------------------------------------------------------
CUCD_HD uint32_t zscan4(uint32_t ux, uint32_t uy) {       // interleave the low 4 bits of ux (even) and uy (odd)
  uint32_t x = ux & 15u, y = uy & 15u;
  x = (x | (x << 2)) & 0x33u; x = (x | (x << 1)) & 0x55u;
  y = (y | (y << 2)) & 0x33u; y = (y | (y << 1)) & 0x55u;
  return x | (y << 1);
}
CUCD_HD void demorton(int z, int& px, int& py) {
  uint32_t x = (uint32_t)z & 0x55u, y = ((uint32_t)z >> 1) & 0x55u;
  x = (x | (x >> 1)) & 0x33u; x = (x | (x >> 2)) & 0x0fu;
  y = (y | (y >> 1)) & 0x33u; y = (y | (y >> 2)) & 0x0fu;
  px = (int)x; py = (int)y;
}
// HEVC 6.4.1 z-scan availability (what TComPattern.cpp:550-727 evaluates for 1 slice / 1 tile)
CUCD_HD bool unit_available(int xc, int yc, int xn, int yn, int W, int H) {
  if (xn < 0 || yn < 0 || xn >= W || yn >= H) return false;
  const int wc = (W + 63) >> 6;
  const int ctuC = (yc >> 6) * wc + (xc >> 6), ctuN = (yn >> 6) * wc + (xn >> 6);
  if (ctuN != ctuC) return ctuN < ctuC;
  return zscan4((uint32_t)(xn & 63) >> 2, (uint32_t)(yn & 63) >> 2) < zscan4((uint32_t)(xc & 63) >> 2, (uint32_t)(yc & 63) >> 2);
}

// first/last linear-border sample of availability unit u (N/2 left units, corner, N/2 above units)
template <int LOG2N>
CUCD_HD void unit_range(int u, int& first, int& count) {
  constexpr int N = 1 << LOG2N, HALF = N / 2;
  if (u < HALF) { first = 4 * u; count = 4; }
  else if (u == HALF) { first = 2 * N; count = 1; }
  else { first = 2 * N + 1 + 4 * (u - HALF - 1); count = 4; }
}

// Phase A (frame mode): availability flag of every unit of every PU + gather of available samples
// from the reconstruction plane into the linear border.  One (PU, unit) per loop trip.
template <int LOG2N>
CUCD_HD void border_gather_frame(int tid, int nthreads, const int16_t* rec, int recStride, int W, int H, int ctuX, int ctuY,
                                 int16_t* lin /*[PUS][LIN]*/, uint8_t* flags /*[PUS][N+1]*/) {
  typedef Geo<LOG2N> G;
  constexpr int N = G::N, UNITS = N + 1, HALF = N / 2;
  for (int idx = tid; idx < G::PUS * UNITS; idx += nthreads) {
    const int p = idx / UNITS, u = idx - p * UNITS;
    int px, py; demorton(p, px, py);
    const int x0 = ctuX + px * N, y0 = ctuY + py * N;
    bool av = false;
    if (x0 + N <= W && y0 + N <= H) {
      int xn, yn;
      if (u < HALF) { xn = x0 - 1; yn = y0 + (HALF - 1 - u) * 4; }
      else if (u == HALF) { xn = x0 - 1; yn = y0 - 1; }
      else { xn = x0 + (u - HALF - 1) * 4; yn = y0 - 1; }
      av = unit_available(x0, y0, xn, yn, W, H);
      if (av) {
        int16_t* b = lin + p * G::LIN;
        if (u < HALF) {           // left column, stored bottom -> top: b[i] = rec(x0-1, y0 + 2N-1-i)
          for (int k = 0; k < 4; k++) { const int i = 4 * u + k; b[i] = rec[(size_t)(y0 + 2 * N - 1 - i) * recStride + x0 - 1]; }
        } else if (u == HALF) {
          b[2 * N] = rec[(size_t)(y0 - 1) * recStride + x0 - 1];
        } else {
          const int xs = (u - HALF - 1) * 4;
          for (int k = 0; k < 4; k++) b[2 * N + 1 + xs + k] = rec[(size_t)(y0 - 1) * recStride + x0 + xs + k];
        }
      }
    }
    flags[idx] = av ? 1 : 0;
  }
}

// Phase B: substitution of unavailable units (TComPattern.cpp:325-336, 442-504): nothing available
// -> 1<<(bd-1); a leading unavailable run takes the first available sample; any other unavailable
// unit repeats the last sample of the nearest available unit below it.
template <int LOG2N>
CUCD_HD void border_substitute(int tid, int nthreads, int bitDepth, int16_t* lin, const uint8_t* flags) {
  typedef Geo<LOG2N> G;
  constexpr int N = G::N, UNITS = N + 1;
  for (int idx = tid; idx < G::PUS * UNITS; idx += nthreads) {
    const int p = idx / UNITS, u = idx - p * UNITS;
    const uint8_t* f = flags + p * UNITS;
    if (f[u]) continue;
    int16_t* b = lin + p * G::LIN;
    int v = u - 1;
    while (v >= 0 && !f[v]) v--;
    int val;
    if (v >= 0) { int first, cnt; unit_range<LOG2N>(v, first, cnt); val = b[first + cnt - 1]; }
    else {
      v = u + 1;
      while (v < UNITS && !f[v]) v++;
      if (v < UNITS) { int first, cnt; unit_range<LOG2N>(v, first, cnt); val = b[first]; }
      else val = 1 << (bitDepth - 1);
    }
    int first, cnt; unit_range<LOG2N>(u, first, cnt);
    for (int k = 0; k < cnt; k++) b[first + k] = (int16_t)val;
  }
}

// Phase C: from the linear unfiltered border derive the ascending ref arrays the predictors read:
//   Tu[k] = corner, above[0..2N-1];  Lu[k] = corner, left[0..2N-1]  (k = 0..2N)  and, for N=8/16/32,
//   the smoothed copies Tf/Lf (TComPattern.cpp:185-283).  One border sample per loop trip.
template <int LOG2N>
CUCD_HD void border_derive(int tid, int nthreads, int bitDepth, int strongEnabled, const int16_t* lin, int16_t* arrs /*[PUS][NARR][AS]*/) {
  typedef Geo<LOG2N> G;
  constexpr int N = G::N, LEN = 4 * N + 1;
  for (int idx = tid; idx < G::PUS * LEN; idx += nthreads) {
    const int p = idx / LEN, i = idx - p * LEN;
    const int16_t* b = lin + p * G::LIN;
    int16_t* a = arrs + pu_slot<LOG2N>(p) * G::PU_STRIDE;
    const int v = b[i];
    // position in the ascending arrays
    if (i >= 2 * N) a[0 * G::AS + (i - 2 * N)] = (int16_t)v;           // Tu
    if (i <= 2 * N) a[1 * G::AS + (2 * N - i)] = (int16_t)v;           // Lu
    if (G::HAS_FILT) {
      int fv;
      if (i == 0 || i == 4 * N) fv = v;
      else {
        bool strong = false;
        if (LOG2N == 5 && strongEnabled) {
          const int thr = 1 << (bitDepth - 5);
          const int bl = b[0], tl = b[2 * N], tr = b[4 * N];
          strong = iabs32(bl + tl - 2 * b[N]) < thr && iabs32(tl + tr - 2 * b[3 * N]) < thr;
        }
        if (strong) {
          const int bl = b[0], tl = b[2 * N], tr = b[4 * N];
          if (i < 2 * N) fv = ((2 * N - i) * bl + i * tl + N) >> (LOG2N + 1);
          else if (i == 2 * N) fv = tl;
          else fv = ((4 * N - i) * tl + (i - 2 * N) * tr + N) >> (LOG2N + 1);
        } else {
          fv = (b[i - 1] + 2 * v + b[i + 1] + 2) >> 2;
        }
      }
      if (i >= 2 * N) a[2 * G::AS + (i - 2 * N)] = (int16_t)fv;        // Tf
      if (i <= 2 * N) a[3 * G::AS + (2 * N - i)] = (int16_t)fv;        // Lf
    }
  }
}
// pad words after each array so that the 5-word row window never reads uninitialised memory
template <int LOG2N>
CUCD_HD void border_pad(int tid, int nthreads, int16_t* arrs) {
  typedef Geo<LOG2N> G;
  constexpr int N = G::N, PAD = G::AS - (2 * N + 1);
  for (int idx = tid; idx < G::PUS * G::NARR * PAD; idx += nthreads) {
    const int arr = idx / PAD, k = idx - arr * PAD;
    const int p = arr / G::NARR, which = arr - p * G::NARR;
    arrs[p * G::PU_STRIDE + which * G::AS + 2 * N + 1 + k] = 0;
  }
}
// DC value of every PU (TComPrediction.cpp:183-222 with bAbove = bLeft = true), from the unfiltered arrays
template <int LOG2N>
CUCD_HD void border_dc(int tid, int nthreads, const int16_t* arrs, int16_t* dc /*[PUS]*/) {
  typedef Geo<LOG2N> G;
  constexpr int N = G::N;
  for (int p = tid; p < G::PUS; p += nthreads) {
    const int16_t* a = arrs + pu_slot<LOG2N>(p) * G::PU_STRIDE;
    int sum = N;
    for (int k = 1; k <= N; k++) sum += a[k] + a[G::AS + k];
    dc[p] = (int16_t)(sum >> (LOG2N + 1));
  }
}

// ---------------------------------------------------------------------------------------------
// source tile in registers
// ---------------------------------------------------------------------------------------------
struct Tile { uint32_t r[32]; };   // r[y*4 + j] = (s[y][2j], s[y][2j+1])

// 8 rows x 8 int16 from memory (16-byte aligned rows)
CUCD_HD void tile_load(Tile& t, const int16_t* p, int stride) {
#pragma unroll
  for (int y = 0; y < 8; y++) {
#if defined(__CUDA_ARCH__)
    const uint4 v = *reinterpret_cast<const uint4*>(p + (size_t)y * stride);
    t.r[y * 4 + 0] = v.x; t.r[y * 4 + 1] = v.y; t.r[y * 4 + 2] = v.z; t.r[y * 4 + 3] = v.w;
#else
    const int16_t* q = p + (size_t)y * stride;
    for (int j = 0; j < 4; j++) t.r[y * 4 + j] = (uint32_t)(uint16_t)q[2 * j] | ((uint32_t)(uint16_t)q[2 * j + 1] << 16);
#endif
  }
}
// whole 8x8 transpose of packed pairs
CUCD_HD void tile_transpose8(const Tile& s, Tile& d) {
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int jp = 0; jp < 4; jp++) {
      const uint32_t a = s.r[(2 * jp) * 4 + (i >> 1)], b = s.r[(2 * jp + 1) * 4 + (i >> 1)];
      d.r[i * 4 + jp] = (i & 1) ? pack_hi(a, b) : pack_lo(a, b);
    }
}
// transpose each 4x4 quadrant in place (N = 4: the four PUs of a region are independent)
CUCD_HD void tile_transpose4x4(const Tile& s, Tile& d) {
#pragma unroll
  for (int qy = 0; qy < 2; qy++)
#pragma unroll
    for (int qx = 0; qx < 2; qx++)
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int jp = 0; jp < 2; jp++) {
          const uint32_t a = s.r[(qy * 4 + 2 * jp) * 4 + qx * 2 + (i >> 1)], b = s.r[(qy * 4 + 2 * jp + 1) * 4 + qx * 2 + (i >> 1)];
          d.r[(qy * 4 + i) * 4 + qx * 2 + jp] = (i & 1) ? pack_hi(a, b) : pack_lo(a, b);
        }
}

// ---------------------------------------------------------------------------------------------
// Hadamard on packed residual pairs
// ---------------------------------------------------------------------------------------------
// v = lo + 65536*hi as one integer  ->  max(|lo|, |hi|)
CUCD_HD int pair_max_abs(uint32_t v) {
  const int lo = (int)(int16_t)(v & 0xffffu);
  const int hi = ((int)v - lo) >> 16;
  return imax32(iabs32(lo), iabs32(hi));
}
// 8x8 tile: (sum|H d H^T| + 2) >> 2   (TComRdCost.cpp:1439-1534)
CUCD_HD uint32_t satd8x8_packed(uint32_t* d /*32, destroyed*/) {
  // vertical: 3 stages across the 8 rows, per column pair
#pragma unroll
  for (int j = 0; j < 4; j++) {
#pragma unroll
    for (int len = 1; len < 8; len <<= 1)
#pragma unroll
      for (int y = 0; y < 8; y++)
        if (!(y & len)) { const uint32_t a = d[y * 4 + j], b = d[(y + len) * 4 + j]; d[y * 4 + j] = a + b; d[(y + len) * 4 + j] = a - b; }
  }
  // horizontal: 2 stages across the 4 words of a row
#pragma unroll
  for (int y = 0; y < 8; y++) {
#pragma unroll
    for (int len = 1; len < 4; len <<= 1)
#pragma unroll
      for (int j = 0; j < 4; j++)
        if (!(j & len)) { const uint32_t a = d[y * 4 + j], b = d[y * 4 + j + len]; d[y * 4 + j] = a + b; d[y * 4 + j + len] = a - b; }
  }
  // last stage inside each word: |lo+hi| + |lo-hi| = 2*max(|lo|,|hi|)
  int acc = 0;
#pragma unroll
  for (int k = 0; k < 32; k++) acc += pair_max_abs(d[k]);
  return ((uint32_t)(2 * acc) + 2u) >> 2;
}
// 4x4 block held in 8 words d[y*2 + j]: (sum + 1) >> 1   (TComRdCost.cpp:1343-1437)
CUCD_HD uint32_t satd4x4_packed(uint32_t* d /*8, destroyed*/) {
#pragma unroll
  for (int j = 0; j < 2; j++) {
#pragma unroll
    for (int len = 1; len < 4; len <<= 1)
#pragma unroll
      for (int y = 0; y < 4; y++)
        if (!(y & len)) { const uint32_t a = d[y * 2 + j], b = d[(y + len) * 2 + j]; d[y * 2 + j] = a + b; d[(y + len) * 2 + j] = a - b; }
  }
  int acc = 0;
#pragma unroll
  for (int y = 0; y < 4; y++) {
    const uint32_t a = d[y * 2], b = d[y * 2 + 1];
    acc += pair_max_abs(a + b) + pair_max_abs(a - b);
  }
  return ((uint32_t)(2 * acc) + 1u) >> 1;
}

// ---------------------------------------------------------------------------------------------
// predictors.  All of them produce the residual  d = src - pred  as packed pairs for one 8x8 tile
// (ROWS = 8, WORDS = 4) or one 4x4 block (ROWS = 4, WORDS = 2) whose top-left corner sits at
// (x0, y0) inside the PU, in the orientation being computed (transposed for modes 2..17).
//   smem16 : the CTA's shared memory viewed as int16;  every ref array starts on an even index
//   main0  : int16 index of element 0 (the corner) of the array the mode reads along x
//   side0  : same for the array along y
// ---------------------------------------------------------------------------------------------
template <int ROWS, int WORDS>
CUCD_HD void load_row_window(const uint32_t* smem32, int a, uint32_t* A /*WORDS*/, uint32_t* B /*WORDS*/) {
  // A[j] = (r[a+2j], r[a+2j+1]),  B[j] = (r[a+2j+1], r[a+2j+2])
  const int wi = a >> 1;
  const uint32_t e16 = ((uint32_t)a & 1u) << 4;
  uint32_t w[WORDS + 1];
#pragma unroll
  for (int j = 0; j <= WORDS; j++) w[j] = smem32[wi + j];
#pragma unroll
  for (int j = 0; j < WORDS; j++) {
    A[j] = funnel_r(w[j], w[j + 1], e16);
    B[j] = funnel_rc(w[j], w[j + 1], e16 + 16);
  }
}

// All predictors write the PREDICTION as packed pairs p[y*WORDS + j] = (pred[y][2j], pred[y][2j+1]);
// the ALU path then forms src - p, the tensor-core path stores p as bytes (rmd_tc.cuh).
//
// Angular, any angle: two-tap 1/32 interpolation (TComPrediction.cpp:368-383).  With f == 0 the
// interpolation degenerates to an exact copy ((32*a + 16) >> 5 == a), so |angle| == 32 and angle == 0
// need no code of their own (one hot loop body keeps the instruction footprint small - the kernel is
// instruction-fetch sensitive).
template <int ROWS, int WORDS>
CUCD_HD void pred_angular(const uint32_t* smem32, int main0, int x0, int y0, int angle, uint32_t* p) {
#pragma unroll
  for (int y = 0; y < ROWS; y++) {
    const int delta = (y0 + y + 1) * angle;
    const int di = delta >> 5;
    const uint32_t f = (uint32_t)delta & 31u;
    uint32_t A[WORDS], B[WORDS];
    load_row_window<ROWS, WORDS>(smem32, main0 + x0 + di + 1, A, B);
#pragma unroll
    for (int j = 0; j < WORDS; j++) {
      const uint32_t t = A[j] * (32u - f) + (B[j] * f + 0x00100010u);
      p[y * WORDS + j] = (t >> 5) & 0x07ff07ffu;
    }
  }
}
// Pure vertical / horizontal modes patch the first column for the luma edge filter of N <= 16
// (TComPrediction.cpp:346-363): pred[y][0] = clip(ref[1] + ((side[y+1] - side[0]) >> 1)).
template <int ROWS, int WORDS>
CUCD_HD void patch_pure_edge(const int16_t* smem16, int main0, int side0, int y0, int maxVal, uint32_t* p) {
  const int s0 = smem16[side0], m1 = smem16[main0 + 1];
#pragma unroll
  for (int y = 0; y < ROWS; y++) {
    int v = m1 + ((smem16[side0 + y0 + y + 1] - s0) >> 1);
    v = imin32(imax32(v, 0), maxVal);
    p[y * WORDS] = (p[y * WORDS] & 0xffff0000u) | (uint32_t)v;
  }
}
// DC with edge smoothing for N <= 16 (TComPrediction.cpp:266-276, 818-841); orientation-symmetric
template <int ROWS, int WORDS>
CUCD_HD void pred_dc(const int16_t* smem16, int main0, int side0, int x0, int y0, int dc, bool edge, uint32_t* p) {
  const uint32_t dc2 = (uint32_t)dc * 0x00010001u;
#pragma unroll
  for (int y = 0; y < ROWS; y++)
#pragma unroll
    for (int j = 0; j < WORDS; j++) {
      uint32_t v = dc2;
      if (edge) {
        const int yy = y0 + y;
        if (yy == 0) {             // first row of the PU: filtered against the main-side neighbours
          const int xa = x0 + 2 * j;
          int lo = (smem16[main0 + 1 + xa] + 3 * dc + 2) >> 2;
          const int hi = (smem16[main0 + 2 + xa] + 3 * dc + 2) >> 2;
          if (xa == 0) lo = (smem16[main0 + 1] + smem16[side0 + 1] + 2 * dc + 2) >> 2;
          v = (uint32_t)lo | ((uint32_t)hi << 16);
        } else if (x0 == 0 && j == 0) {
          const int lo = (smem16[side0 + 1 + yy] + 3 * dc + 2) >> 2;
          v = (v & 0xffff0000u) | (uint32_t)lo;
        }
      }
      p[y * WORDS + j] = v;
    }
}
// planar (TComPrediction.cpp:755-805); T = array along x, L = array along y
template <int ROWS, int WORDS>
CUCD_HD void pred_planar(const int LOG2N, const int16_t* smem16, int t0, int l0, int x0, int y0, uint32_t* p) {
  const int N = 1 << LOG2N;
  const int tr = smem16[t0 + 1 + N], bl = smem16[l0 + 1 + N];
  int vert[2 * WORDS], vstep[2 * WORDS];
#pragma unroll
  for (int x = 0; x < 2 * WORDS; x++) {
    const int t = smem16[t0 + 1 + x0 + x];
    vert[x] = (N - 1 - y0) * t + (y0 + 1) * bl;
    vstep[x] = bl - t;
  }
#pragma unroll
  for (int y = 0; y < ROWS; y++) {
    const int l = smem16[l0 + 1 + y0 + y];
    int hor = (N - 1 - x0) * l + (x0 + 1) * tr + N;
    const int hstep = tr - l;
#pragma unroll
    for (int j = 0; j < WORDS; j++) {
      const int p0 = (hor + vert[2 * j]) >> (LOG2N + 1); hor += hstep;
      const int p1 = (hor + vert[2 * j + 1]) >> (LOG2N + 1); hor += hstep;
      p[y * WORDS + j] = (uint32_t)p0 | ((uint32_t)p1 << 16);
    }
#pragma unroll
    for (int x = 0; x < 2 * WORDS; x++) vert[x] += vstep[x];
  }
}

// negative-angle modes: element k of the extended main reference, k in [-N, N]
// (TComPrediction.cpp:300-322: ref[k<0] = side[(128 + |k|*invAngle) >> 8])
CUCD_HD int16_t ext_ref_sample(const int16_t* smem16, int main0, int side0, int invAngle, int k) {
  if (k >= 0) return smem16[main0 + k];
  return smem16[side0 + ((128 - k * invAngle) >> 8)];
}

}  // namespace cucd
