// me_kernels.cu - integer-pel motion-estimation SAD surfaces on sm_100a.
//
// Replaces the distortion part of TEncSearch::xTZSearchHelp (TEncSearch.cpp:336-437) and of the
// full/raster search xPatternSearch (:3886-3943): for a PU and an inclusive integer MV window the
// kernel returns SAD(cur PU, ref + mv) for EVERY mv, computed exactly like TComRdCost::xGetSAD*
// (TComRdCost.cpp:465-962): rows stepped by 1<<subShift, sum << subShift, then >> (bitDepth-8).
// The host walks the TZ pattern over the table and adds getCost(mv) itself, so ties break as in HM.
//
// One CTA = one 32x8 tile of candidates of one PU.  The PU and the (w+32)x(h+8) reference window are
// staged once in shared memory; a lane owns one dx, a warp one dy; samples are handled as packed
// int16 pairs: |a-b| per half = max-min (VIMNMX.S16x2 x2 + ISUB), accumulated with IDP.2A.
#include <cuda_runtime.h>
#include "rmd_core.cuh"
#include "kernels.h"

namespace cucd {

constexpr int kMeTileX = 32, kMeTileY = 8;
constexpr int kMeRefPitch = 64 + kMeTileX + 4;       // int16 per staged reference row (even)
constexpr int kMeRefRows = 64 + kMeTileY;

__global__ void __launch_bounds__(256)
me_sad_kernel(const MePlanes mp, const MeJob* __restrict__ jobs, const int32_t* __restrict__ tileJob, const int32_t* __restrict__ tileIdx,
              uint32_t* __restrict__ out) {
  __shared__ __align__(16) int16_t sCur[64 * 64];
  __shared__ __align__(16) int16_t sRef[kMeRefRows * kMeRefPitch];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const MeJob job = jobs[tileJob[blockIdx.x]];
  const int cols = job.right - job.left + 1, rows = job.bottom - job.top + 1;
  const int tilesX = (cols + kMeTileX - 1) / kMeTileX;
  const int t = tileIdx[blockIdx.x];
  const int dx0 = (t % tilesX) * kMeTileX, dy0 = (t / tilesX) * kMeTileY;     // relative to (left, top)
  const int w = job.w, h = job.h;

  const int16_t* cur = mp.cur + job.curOff;
  for (int i = tid; i < w * h; i += 256) { const int y = i / w, x = i - y * w; sCur[y * w + x] = cur[(size_t)y * mp.curStride + x]; }
  const int refStride = mp.refStride[job.refSlot];
  const int16_t* ref = mp.ref[job.refSlot] + job.refOff + (long long)(job.top + dy0) * refStride + (job.left + dx0);
  const int winW = min(w + kMeTileX - 1, w + cols - dx0 - 1), winH = min(h + kMeTileY - 1, h + rows - dy0 - 1);
  for (int i = tid; i < kMeRefRows * kMeRefPitch; i += 256) {
    const int y = i / kMeRefPitch, x = i - y * kMeRefPitch;
    sRef[i] = (y < winH && x < winW) ? ref[(long long)y * refStride + x] : (int16_t)0;
  }
  __syncthreads();

  const int dx = dx0 + lane, dy = dy0 + warp;
  if (dx >= cols || dy >= rows) return;
  const int step = 1 << job.subShift;
  const uint32_t* cw = reinterpret_cast<const uint32_t*>(sCur);
  const uint32_t* rw = reinterpret_cast<const uint32_t*>(sRef);
  const uint32_t e16 = ((uint32_t)lane & 1u) << 4;
  const int pairs = w >> 1;
  int acc = 0;
  for (int y = 0; y < h; y += step) {
    const int rbase = ((y + warp) * kMeRefPitch + lane) >> 1;
    const int cbase = (y * w) >> 1;
    uint32_t prev = rw[rbase];
    for (int j = 0; j < pairs; j++) {
      const uint32_t next = rw[rbase + j + 1];
      const uint32_t r = funnel_r(prev, next, e16);
      const uint32_t c = cw[cbase + j];
      const uint32_t d = __vmaxs2(c, r) - __vmins2(c, r);
      acc = __dp2a_lo((int)d, 0x0101, acc);
      prev = next;
    }
  }
  out[job.outOff + (long long)dy * cols + dx] = ((uint32_t)acc << job.subShift) >> (mp.bitDepth - 8);
}

cudaError_t launch_me_sad(const MePlanes& mp, const MeJob* jobs, int nJobs, const int32_t* tileJob, const int32_t* tileIdx, int nTiles,
                          uint32_t* out, cudaStream_t st, int* launches) {
  (void)nJobs;
  if (nTiles <= 0) return cudaSuccess;
  me_sad_kernel<<<nTiles, 256, 0, st>>>(mp, jobs, tileJob, tileIdx, out);
  if (launches) *launches += 1;
  return cudaGetLastError();
}

}  // namespace cucd
