// tcm_host.h - host-side TCM fit between the two GPU passes of the outlier feature.
#pragma once
#include <stdint.h>
#include <stddef.h>
namespace cucd {
constexpr int kTcmBins = 4096;
// histogram of |coeff/8| of one frequency (kTcmBins counts) -> Yc (TEncSlice.cpp:291-392)
double tcm_fit_one(const uint32_t* count, int nSamples);
// 16 x kTcmBins histograms of one picture -> yc[16], thr[16] = (int)(Yc*8)
void tcm_fit_picture(const uint32_t* hist, int nBlocks, double* yc, int32_t* thr);
}
