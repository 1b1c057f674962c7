// feature_core.cuh - per-thread arithmetic of the fork's outlier ("OBF") texture features.
// Host/device so that tests/emul can replay it.  Reference: TEncSlice.cpp:55-77 (4-point forward
// butterfly with g_aiT4 = {64, 83, 36}, TComRom.cpp:464-468), :922-966 (two-stage 4x4 DCT of every
// source block), :1006-1042 (outlier test and OBF count), :1106-1112 (Outlier plane).
#pragma once
#include "rmd_core.cuh"

namespace cucd {

// one butterfly stage over 4 lines; in[line*4 + k], out[k*4 + line]
CUCD_HD void dct4_stage(const int* in, int* out, int shift) {
  const int add = 1 << (shift - 1);
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const int e0 = in[4 * j] + in[4 * j + 3], o0 = in[4 * j] - in[4 * j + 3];
    const int e1 = in[4 * j + 1] + in[4 * j + 2], o1 = in[4 * j + 1] - in[4 * j + 2];
    out[j] = (64 * e0 + 64 * e1 + add) >> shift;
    out[8 + j] = (64 * e0 - 64 * e1 + add) >> shift;
    out[4 + j] = (83 * o0 + 36 * o1 + add) >> shift;
    out[12 + j] = (36 * o0 - 83 * o1 + add) >> shift;
  }
}
// 4x4 source block (rows of 4 int16) -> 16 coefficients; shift_1st = 2 + bd + 6 - 15, shift_2nd = 8
CUCD_HD void dct4x4(const int16_t* blk, int stride, int bitDepth, int* coeff) {
  int in[16], tmp[16];
#pragma unroll
  for (int y = 0; y < 4; y++) {
#if defined(__CUDA_ARCH__)
    const uint2 v = *reinterpret_cast<const uint2*>(blk + (size_t)y * stride);
    in[y * 4 + 0] = (int)(int16_t)(v.x & 0xffffu); in[y * 4 + 1] = (int)(int16_t)(v.x >> 16);
    in[y * 4 + 2] = (int)(int16_t)(v.y & 0xffffu); in[y * 4 + 3] = (int)(int16_t)(v.y >> 16);
#else
    for (int x = 0; x < 4; x++) in[y * 4 + x] = blk[(size_t)y * stride + x];
#endif
  }
  dct4_stage(in, tmp, bitDepth - 7);
  dct4_stage(tmp, coeff, 8);
}
// histogram bin of a coefficient: |(Int)(coeff / 8.0)|  (TEncSlice.cpp:962, truncation toward zero)
CUCD_HD int coeff_bin(int c) { return iabs32(c) >> 3; }
// outlier test TEncSlice.cpp:1010 with thr = Yc*8 (Yc is integer valued): kept iff not strictly inside
CUCD_HD bool coeff_is_outlier(int c, int thr) { return c != 0 && !(c < thr && c > -thr); }

// ---------------------------------------------------------------------------------------------
// Fork-aware PU enumeration of one CTU (Testing pictures).  Restates the control flow of TEncCu::xCompressCU around the
// early decisions for an intra picture with the default Naive model (tools_YS.cpp:686-695: Num_OBF == 0 -> TerminateCU,
// else Skip2Nx2N):
//   - a CU that is not completely inside the picture is never predicted and never evaluated, its sub-CUs are visited
//     (bBoundary, TEncCu.cpp:488-489, 645);
//   - bSkip2Nx2N = switch[depth][Skip2Nx2N] && Num_OBF > 0   -> the 2Nx2N PU is not evaluated (:951-996, 1040);
//   - bEarlyTerminate = switch[depth][TerminateCU] && Num_OBF == 0 -> no NxN at depth 3 (:1140-1143), no sub-CUs (:1257-1260).
// numObf(d, cuX, cuY) returns Num_OBF of the whole CU at depth d with CU coordinates (cuX, cuY).
// needed[341]: PU order of the cost tables (depth-major, z-order inside a depth).
// ---------------------------------------------------------------------------------------------
template <class NumObf>
CUCD_HD void prune_mask_ctu(int ctuX, int ctuY, int W, int H, const uint8_t* swSkip, const uint8_t* swTerm, NumObf numObf, uint8_t* needed) {
  for (int i = 0; i < 341; i++) needed[i] = 0;
  unsigned long long visited = 1;       // bit i: CU i (z-order) of the current depth is reached by the recursion
  int off = 0;                          // first PU index of the depth
  for (int d = 0; d < 4; d++) {
    const int n = 1 << (2 * d), size = 64 >> d;
    unsigned long long next = 0;        // the (at most 64) CUs of the next depth
    for (int i = 0; i < n; i++) {
      if (!((visited >> i) & 1)) continue;
      int px, py; demorton(i, px, py);
      const int x = ctuX + px * size, y = ctuY + py * size;
      if (x >= W || y >= H) continue;                                   // outside the picture: not coded at all
      bool recurse = true;                                              // boundary CU: forced split
      if (x + size <= W && y + size <= H) {
        const int num = numObf(d, x / size, y / size);
        const bool term = swTerm[d] && num == 0, skip = swSkip[d] && num > 0;
        if (!skip) needed[off + i] = 1;
        recurse = !term;
      }
      if (!recurse) continue;
      if (d == 3) { for (int q = 0; q < 4; q++) needed[85 + 4 * i + q] = 1; }   // NxN of the 8x8 CU (W, H multiples of 8: never a boundary CU)
      else next |= 0xfull << (4 * i);
    }
    off += n;
    visited = next;
  }
}

}  // namespace cucd