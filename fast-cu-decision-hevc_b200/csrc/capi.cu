// capi.cu - the C ABI (include/cucudecide.h) and the host runtime behind it: device buffers, pinned
// staging, streams, the two-pass feature pipeline with the host TCM fit in between, size-class
// bucketing of batched RMD requests and tiling of ME search windows.  There is no CPU compute path:
// every entry point either runs the CUDA kernels or fails.
#include <cuda_runtime.h>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <string>
#include <thread>
#include <vector>
#include <algorithm>
#include "../../include/cucudecide.h"
#include "kernels.h"
#include "rmd_tc2.cuh"
#include "tcm_host.h"

using namespace cucd;

namespace {
std::string g_createError;

template <class T> struct DevBuf {
  T* p = nullptr; size_t n = 0;
  cudaError_t reserve(size_t count) {
    if (count <= n) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; n = 0;
    cudaError_t e = cudaMalloc(&p, count * sizeof(T));
    if (e == cudaSuccess) n = count;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
};
template <class T> struct PinBuf {
  T* p = nullptr; size_t n = 0;
  cudaError_t reserve(size_t count) {
    if (count <= n) return cudaSuccess;
    if (p) cudaFreeHost(p);
    p = nullptr; n = 0;
    cudaError_t e = cudaMallocHost(&p, count * sizeof(T));
    if (e == cudaSuccess) n = count;
    return e;
  }
  void release() { if (p) cudaFreeHost(p); p = nullptr; n = 0; }
};

struct RefPlane { DevBuf<int16_t> buf; int stride = 0, marginX = 0, marginY = 0; bool set = false; };
}  // namespace

struct cucd_handle {
  cucd_config cfg;
  int ctusPerRow = 0, ctusPerCol = 0, ctusPerPic = 0, pitch = 0;
  size_t planeSamples = 0;
  cudaStream_t sMain = nullptr, sFeat = nullptr, sGrp[2] = {nullptr, nullptr};   // sGrp: [0] RMD compute, [1] cost-table download
  cudaStream_t sUp = nullptr;                     // picture upload
  static constexpr int kGroups = 8;               // sub-groups of pictures one cuCUDecide_frames call is pipelined over
  cudaEvent_t evFork = nullptr, evJoin = nullptr;   // cucd_dev_frames: feature path beside the RMD kernel
  cudaEvent_t evUp = nullptr, evHist = nullptr, evUpG[kGroups] = {}, evRmdG[kGroups] = {};
  static constexpr int kTimeRing = 64;
  cudaEvent_t evRmd0[kTimeRing] = {}, evRmd1[kTimeRing] = {};
  cudaEvent_t evK0 = nullptr, evK1 = nullptr;     // around the kernels of the last batch call (cucd_last_kernel_time)
  bool kTimed = false;
  long long rmdCalls = 0;
  int launches = 0;
  long long launchTotal = 0;
  std::string err;
  int hostThreads = 1;
  // frame path
  DevBuf<int16_t> dOrg, dRec, dObf, dOutlier;
  DevBuf<uint32_t> dCost, dHist;
  DevBuf<uint8_t> dCostPacked;    // CUCD_PACKED_CTU_BYTES per CTU (cucd_frame_out.rmd_cost_packed)
  DevBuf<int8_t> dHadamard;       // +-(H8 x H8), +-(blockdiag H4 x H4) operands of the tensor-core SATD
  DevBuf<uint8_t> dTc2Tables;     // interpolation-weight operands of the tensor-core prediction (rmd_tc2.cuh)
  int useTensor = 0;              // 8-bit content (cucd_set_rmd_path): 1 = predictions + Hadamard on tcgen05, 2 = Hadamard only, 0 = integer ALU
  DevBuf<int32_t> dThr, dNum[4], dSum[4], dCtuHad;
  PinBuf<uint32_t> hHist;
  PinBuf<int32_t> hThr;
  size_t cuCount[4] = {0, 0, 0, 0};
  // batch RMD path
  DevBuf<int16_t> bOrg, bBorder;
  DevBuf<BatchPu> bPus;
  PinBuf<uint8_t> hStage; DevBuf<uint8_t> bStage;   // S2: one pinned staging block up, one down
  DevBuf<uint32_t> bOut;
  // ME path
  std::vector<RefPlane> refs;
  DevBuf<int16_t> dCur; int curStride = 0; bool curSet = false;
  DevBuf<const int16_t*> dRefPtr; DevBuf<int32_t> dRefStride;
  DevBuf<SubpelJob> dSubJobs;
  DevBuf<TuJob> tJobs; DevBuf<int32_t> tCoef, tAbs; DevBuf<int16_t> tPix; DevBuf<uint32_t> tDist;   // TU coding path
  DevBuf<TmvCu> dTmvCus; DevBuf<double> dDoubles;   // texture features / AQ activity
  DevBuf<MeJob> dJobs; DevBuf<int32_t> dTileJob, dTileIdx; DevBuf<uint32_t> dSad;
};

namespace {

int fail(cucd_handle* h, int code, const std::string& msg) {
  if (h) h->err = msg; else g_createError = msg;
  return code;
}
int cuda_fail(cucd_handle* h, cudaError_t e, const char* what) {
  return fail(h, CUCD_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define CK(call)                                                         \
  do {                                                                   \
    cudaError_t e__ = (call);                                            \
    if (e__ != cudaSuccess) return cuda_fail(h, e__, #call);             \
  } while (0)

void flush_launches(cucd_handle* h) { h->launchTotal += h->launches; h->launches = 0; }

FrameSource make_frame_source(const cucd_handle* h, const int16_t* org, long long orgPic, int orgStride, const int16_t* rec, long long recPic,
                              int recStride, uint32_t* out) {
  FrameSource fs;
  fs.org = org; fs.rec = rec; fs.orgPicStride = orgPic; fs.recPicStride = recPic; fs.orgStride = orgStride; fs.recStride = recStride;
  fs.W = h->cfg.width; fs.H = h->cfg.height; fs.ctusPerRow = h->ctusPerRow; fs.ctusPerPic = h->ctusPerPic; fs.out = out;
  return fs;
}
FeaturePlanes make_feature_planes(const cucd_handle* h, const int16_t* org, long long orgPic, int orgStride) {
  FeaturePlanes fp;
  fp.org = org; fp.orgPicStride = orgPic; fp.orgStride = orgStride; fp.W = h->cfg.width; fp.H = h->cfg.height;
  fp.ctusPerRow = h->ctusPerRow; fp.ctusPerPic = h->ctusPerPic; fp.bitDepth = h->cfg.bit_depth;
  return fp;
}

cudaError_t launch_rmd_auto(cucd_handle* h, const FrameSource& fs, int nPics, cudaStream_t st) {
  if (h->useTensor == 1)
    return launch_rmd_frames_tc2(fs, nPics, h->cfg.strong_intra_smoothing, h->dTc2Tables.p, h->dTc2Tables.p + tc2::kWinTableBytes, h->dHadamard.p, st, &h->launches);
  if (h->useTensor == 2) return launch_rmd_frames_tc(fs, nPics, h->cfg.strong_intra_smoothing, h->dHadamard.p, st, &h->launches);
  return launch_rmd_frames(fs, nPics, h->cfg.bit_depth, h->cfg.strong_intra_smoothing, st, &h->launches);
}

// parallel-for over [0, n) on up to `threads` host threads
template <class F> void parallel_for(int n, int threads, F fn) {
  if (threads <= 1 || n <= 1) { for (int i = 0; i < n; i++) fn(i); return; }
  std::atomic<int> next(0);
  auto body = [&]() { for (;;) { const int i = next.fetch_add(1); if (i >= n) break; fn(i); } };
  const int nt = std::min(threads, n);
  std::vector<std::thread> pool;
  for (int t = 1; t < nt; t++) pool.emplace_back(body);
  body();
  for (auto& t : pool) t.join();
}

}  // namespace

extern "C" {

int cucd_abi_version(void) { return CUCD_ABI_VERSION; }

uint32_t cucd_packed_cost(const uint8_t* ctu_table, int pu, int mode) {
  if (pu < CUCD_PACKED_WIDE_PUS) return reinterpret_cast<const uint32_t*>(ctu_table)[pu * 35 + mode];
  const uint16_t v = reinterpret_cast<const uint16_t*>(ctu_table + CUCD_PACKED_WIDE_PUS * 35 * 4)[(pu - CUCD_PACKED_WIDE_PUS) * 35 + mode];
  return v == 0xFFFFu ? 0xFFFFFFFFu : (uint32_t)v;
}

const char* cucd_last_error(const cucd_handle* h) { return h ? h->err.c_str() : g_createError.c_str(); }
long long cucd_launch_count(const cucd_handle* h) { return h ? h->launchTotal + h->launches : 0; }

int cucd_create(const cucd_config* cfg, cucd_handle** out) {
  if (!cfg || !out) return fail(nullptr, CUCD_ERR_INVALID, "cucd_create: null argument");
  *out = nullptr;
  if (cfg->ctu_size != 64 || cfg->max_depth != 4) return fail(nullptr, CUCD_ERR_UNSUPPORTED, "cucd_create: only CTU 64 / depth 4 is built");
  if (cfg->bit_depth < 8 || cfg->bit_depth > 10) return fail(nullptr, CUCD_ERR_UNSUPPORTED, "cucd_create: bit depth must be 8..10");
  if (cfg->width < 8 || cfg->height < 8 || (cfg->width & 7) || (cfg->height & 7)) return fail(nullptr, CUCD_ERR_INVALID, "cucd_create: width/height must be multiples of 8");
  if (cfg->max_pictures < 1) return fail(nullptr, CUCD_ERR_INVALID, "cucd_create: max_pictures < 1");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0) return fail(nullptr, CUCD_ERR_NO_DEVICE, std::string("cucd_create: no CUDA device (") + cudaGetErrorString(e) + "); this library has no CPU path");
  if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, CUCD_ERR_INVALID, "cucd_create: bad device ordinal");
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, cfg->device)) != cudaSuccess) return cuda_fail(nullptr, e, "cudaGetDeviceProperties");
  if (prop.major != 10) return fail(nullptr, CUCD_ERR_NO_DEVICE, "cucd_create: kernels are built for sm_100a only, device is sm_" + std::to_string(prop.major * 10 + prop.minor));
  if ((e = cudaSetDevice(cfg->device)) != cudaSuccess) return cuda_fail(nullptr, e, "cudaSetDevice");

  cucd_handle* h = new cucd_handle;
  h->cfg = *cfg;
  h->ctusPerRow = (cfg->width + 63) / 64; h->ctusPerCol = (cfg->height + 63) / 64; h->ctusPerPic = h->ctusPerRow * h->ctusPerCol;
  h->pitch = (cfg->width + 63) & ~63;
  h->planeSamples = (size_t)h->pitch * cfg->height;
  // TCM fits of a 16-picture step are ~4 ms of CPU work; 8 threads hide them behind the RMD kernel.  More would only add thread start-up
  // cost per call (the pool is spawned per call) and oversubscribe the host when several ranks / instances share it.
  h->hostThreads = cfg->host_threads > 0 ? cfg->host_threads : (int)std::min(8u, std::max(1u, std::thread::hardware_concurrency()));
  for (int d = 0; d < 4; d++) h->cuCount[d] = (size_t)(cfg->width / (64 >> d)) * (cfg->height / (64 >> d));
  const size_t P = (size_t)cfg->max_pictures;
  int prioLow = 0, prioHigh = 0;
  cudaDeviceGetStreamPriorityRange(&prioLow, &prioHigh);
  bool ok = cudaStreamCreateWithFlags(&h->sMain, cudaStreamNonBlocking) == cudaSuccess &&
            cudaStreamCreateWithPriority(&h->sFeat, cudaStreamNonBlocking, prioHigh) == cudaSuccess &&   // small feature kernels go ahead of queued RMD blocks
            cudaStreamCreateWithFlags(&h->sGrp[0], cudaStreamNonBlocking) == cudaSuccess &&
            cudaStreamCreateWithFlags(&h->sGrp[1], cudaStreamNonBlocking) == cudaSuccess &&
            cudaStreamCreateWithFlags(&h->sUp, cudaStreamNonBlocking) == cudaSuccess &&
            cudaEventCreateWithFlags(&h->evUp, cudaEventDisableTiming) == cudaSuccess &&
            cudaEventCreateWithFlags(&h->evHist, cudaEventDisableTiming) == cudaSuccess &&
            cudaEventCreateWithFlags(&h->evFork, cudaEventDisableTiming) == cudaSuccess &&
            cudaEventCreateWithFlags(&h->evJoin, cudaEventDisableTiming) == cudaSuccess;
  for (int i = 0; i < cucd_handle::kGroups; i++)
    ok = ok && cudaEventCreateWithFlags(&h->evUpG[i], cudaEventDisableTiming) == cudaSuccess && cudaEventCreateWithFlags(&h->evRmdG[i], cudaEventDisableTiming) == cudaSuccess;
  ok = ok && cudaEventCreate(&h->evK0) == cudaSuccess && cudaEventCreate(&h->evK1) == cudaSuccess;
  for (int i = 0; i < cucd_handle::kTimeRing; i++) ok = ok && cudaEventCreate(&h->evRmd0[i]) == cudaSuccess && cudaEventCreate(&h->evRmd1[i]) == cudaSuccess;
  ok = ok && h->dOrg.reserve(P * h->planeSamples) == cudaSuccess && h->dRec.reserve(P * h->planeSamples) == cudaSuccess;
  ok = ok && h->dCost.reserve(P * h->ctusPerPic * kPusPerCtu * kNumModes) == cudaSuccess;
  ok = ok && h->dCostPacked.reserve(P * h->ctusPerPic * (size_t)CUCD_PACKED_CTU_BYTES) == cudaSuccess;
  ok = ok && h->dHist.reserve(P * kHistFreqs * kHistBins) == cudaSuccess && h->dThr.reserve(P * kHistFreqs) == cudaSuccess;
  ok = ok && h->dObf.reserve(P * (size_t)(cfg->width / 4) * (cfg->height / 4)) == cudaSuccess;
  ok = ok && h->dOutlier.reserve(P * (size_t)cfg->width * cfg->height) == cudaSuccess;
  for (int d = 0; d < 4; d++) ok = ok && h->dNum[d].reserve(P * std::max<size_t>(1, h->cuCount[d])) == cudaSuccess && h->dSum[d].reserve(P * std::max<size_t>(1, h->cuCount[d])) == cudaSuccess;
  ok = ok && h->dCtuHad.reserve(P * h->ctusPerPic) == cudaSuccess;
  ok = ok && h->dHadamard.reserve(16384) == cudaSuccess && launch_hadamard_operands(h->dHadamard.p, h->sMain) == cudaSuccess &&
       cudaStreamSynchronize(h->sMain) == cudaSuccess;
  if (ok) {
    std::vector<uint8_t> tab(tc2::kWinTableBytes + tc2::kN4TableBytes);
    tc2::fill_win_tables(tab.data()); tc2::fill_n4_tables(tab.data() + tc2::kWinTableBytes);
    ok = h->dTc2Tables.reserve(tab.size()) == cudaSuccess && cudaMemcpy(h->dTc2Tables.p, tab.data(), tab.size(), cudaMemcpyHostToDevice) == cudaSuccess;
  }
  h->useTensor = cfg->bit_depth == 8 ? 1 : 0;
  { const char* e = getenv("CUCD_RMD_PATH");
    if (e && !strcmp(e, "alu")) h->useTensor = 0;
    if (e && !strcmp(e, "tc1") && cfg->bit_depth == 8) h->useTensor = 2; }
  ok = ok && h->hHist.reserve(P * kHistFreqs * kHistBins) == cudaSuccess && h->hThr.reserve(P * kHistFreqs) == cudaSuccess;
  if (!ok) {
    const std::string msg = std::string("cucd_create: allocation failed: ") + cudaGetErrorString(cudaGetLastError());
    cucd_destroy(h);
    return fail(nullptr, CUCD_ERR_NOMEM, msg);
  }
  *out = h;
  return CUCD_OK;
}

int cucd_destroy(cucd_handle* h) {
  if (!h) return CUCD_OK;
  cudaSetDevice(h->cfg.device);
  cudaDeviceSynchronize();
  h->dOrg.release(); h->dRec.release(); h->dObf.release(); h->dOutlier.release(); h->dCost.release(); h->dCostPacked.release(); h->dHist.release(); h->dThr.release();
  for (int d = 0; d < 4; d++) { h->dNum[d].release(); h->dSum[d].release(); }
  h->dCtuHad.release(); h->hHist.release(); h->hThr.release(); h->dHadamard.release(); h->dTc2Tables.release();
  h->bOrg.release(); h->bBorder.release(); h->bPus.release(); h->bOut.release(); h->hStage.release(); h->bStage.release();
  for (auto& r : h->refs) r.buf.release();
  h->dTmvCus.release(); h->dDoubles.release();
  h->dSubJobs.release(); h->tJobs.release(); h->tCoef.release(); h->tAbs.release(); h->tPix.release(); h->tDist.release();
  h->dCur.release(); h->dRefPtr.release(); h->dRefStride.release(); h->dJobs.release(); h->dTileJob.release(); h->dTileIdx.release(); h->dSad.release();
  for (int i = 0; i < cucd_handle::kTimeRing; i++) { if (h->evRmd0[i]) cudaEventDestroy(h->evRmd0[i]); if (h->evRmd1[i]) cudaEventDestroy(h->evRmd1[i]); }
  for (int i = 0; i < cucd_handle::kGroups; i++) { if (h->evUpG[i]) cudaEventDestroy(h->evUpG[i]); if (h->evRmdG[i]) cudaEventDestroy(h->evRmdG[i]); }
  if (h->evK0) cudaEventDestroy(h->evK0);
  if (h->evK1) cudaEventDestroy(h->evK1);
  if (h->sUp) cudaStreamDestroy(h->sUp);
  if (h->evUp) cudaEventDestroy(h->evUp);
  if (h->evHist) cudaEventDestroy(h->evHist);
  if (h->evFork) cudaEventDestroy(h->evFork);
  if (h->evJoin) cudaEventDestroy(h->evJoin);
  for (int i = 0; i < 2; i++) if (h->sGrp[i]) cudaStreamDestroy(h->sGrp[i]);
  if (h->sMain) cudaStreamDestroy(h->sMain);
  if (h->sFeat) cudaStreamDestroy(h->sFeat);
  delete h;
  return CUCD_OK;
}

int cucd_set_rmd_path(cucd_handle* h, int use_tensor_cores) {
  if (!h) return CUCD_ERR_INVALID;
  if (use_tensor_cores && h->cfg.bit_depth != 8) return fail(h, CUCD_ERR_UNSUPPORTED, "cucd_set_rmd_path: the tcgen05 kind::i8 path needs 8-bit content");
  if (use_tensor_cores < 0 || use_tensor_cores > 2) return fail(h, CUCD_ERR_INVALID, "cucd_set_rmd_path: path must be 0, 1 or 2");
  h->useTensor = use_tensor_cores;
  return CUCD_OK;
}

int cucd_rmd_kernel_time(cucd_handle* h, int nCalls, float* avg_ms) {
  if (!h || !avg_ms || nCalls < 1) return fail(h, CUCD_ERR_INVALID, "cucd_rmd_kernel_time: bad argument");
  const int n = (int)std::min<long long>(std::min<long long>(nCalls, h->rmdCalls), cucd_handle::kTimeRing);
  if (n < 1) return fail(h, CUCD_ERR_INVALID, "cucd_rmd_kernel_time: no timed call yet");
  double sum = 0;
  for (int i = 0; i < n; i++) {
    const int slot = (int)((h->rmdCalls - 1 - i) % cucd_handle::kTimeRing);
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, h->evRmd0[slot], h->evRmd1[slot]));
    sum += ms;
  }
  *avg_ms = (float)(sum / n);
  return n;
}

int cucd_last_kernel_time(cucd_handle* h, float* ms) {
  if (!h || !ms) return fail(h, CUCD_ERR_INVALID, "cucd_last_kernel_time: bad argument");
  if (!h->kTimed) return fail(h, CUCD_ERR_INVALID, "cucd_last_kernel_time: no batch call yet");
  CK(cudaEventElapsedTime(ms, h->evK0, h->evK1));
  return CUCD_OK;
}

int cucd_tcm_fit(const uint32_t* hist, int nBlocks, double* yc, int32_t* thr) {
  if (!hist || !yc || !thr || nBlocks <= 0) return CUCD_ERR_INVALID;
  tcm_fit_picture(hist, nBlocks, yc, thr);
  return CUCD_OK;
}

// ------------------------------------------------------------------------------------------------
// device-resident entry points
// ------------------------------------------------------------------------------------------------
int cucd_dev_rmd_frames(cucd_handle* h, void* stream, int nPics, const int16_t* d_org, long long orgPicStride, int orgStride,
                        const int16_t* d_rec, long long recPicStride, int recStride, uint32_t* d_rmd_cost) {
  if (!h || nPics < 1 || !d_org || !d_rec || !d_rmd_cost) return fail(h, CUCD_ERR_INVALID, "cucd_dev_rmd_frames: bad argument");
  if ((orgStride & 7) || (orgPicStride & 7) || ((uintptr_t)d_org & 15)) return fail(h, CUCD_ERR_INVALID, "cucd_dev_rmd_frames: source plane must be 16-byte aligned with strides multiple of 8");
  CK(cudaSetDevice(h->cfg.device));
  const FrameSource fs = make_frame_source(h, d_org, orgPicStride, orgStride, d_rec, recPicStride, recStride, d_rmd_cost);
  CK(launch_rmd_auto(h, fs, nPics, (cudaStream_t)stream));
  flush_launches(h);
  return CUCD_OK;
}

int cucd_dev_feature_hist(cucd_handle* h, void* stream, int nPics, const int16_t* d_org, long long orgPicStride, int orgStride, uint32_t* d_hist) {
  if (!h || nPics < 1 || !d_org || !d_hist) return fail(h, CUCD_ERR_INVALID, "cucd_dev_feature_hist: bad argument");
  if ((orgStride & 3) || (orgPicStride & 3) || ((uintptr_t)d_org & 7)) return fail(h, CUCD_ERR_INVALID, "cucd_dev_feature_hist: source plane must be 8-byte aligned with strides multiple of 4");
  CK(cudaSetDevice(h->cfg.device));
  CK(launch_feature_hist(make_feature_planes(h, d_org, orgPicStride, orgStride), nPics, d_hist, (cudaStream_t)stream, &h->launches));
  flush_launches(h);
  return CUCD_OK;
}

int cucd_dev_feature_obf(cucd_handle* h, void* stream, int nPics, const int16_t* d_org, long long orgPicStride, int orgStride,
                         const int32_t* d_thr, int16_t* d_obf, int16_t* d_outlier, int32_t* const d_num_obf[4],
                         int32_t* const d_n_outlier[4], int32_t* d_ctu_src_had) {
  if (!h || nPics < 1 || !d_org || !d_thr || !d_obf || !d_outlier || !d_num_obf || !d_n_outlier) return fail(h, CUCD_ERR_INVALID, "cucd_dev_feature_obf: bad argument");
  if ((orgStride & 7) || (orgPicStride & 7) || ((uintptr_t)d_org & 15)) return fail(h, CUCD_ERR_INVALID, "cucd_dev_feature_obf: source plane must be 16-byte aligned with strides multiple of 8");
  CK(cudaSetDevice(h->cfg.device));
  const FeaturePlanes fp = make_feature_planes(h, d_org, orgPicStride, orgStride);
  FeatureOut fo;
  fo.obf = d_obf; fo.obfPicStride = (long long)(h->cfg.width / 4) * (h->cfg.height / 4);
  fo.outlier = d_outlier; fo.outlierPicStride = (long long)h->cfg.width * h->cfg.height;
  for (int d = 0; d < 4; d++) {
    if (!d_num_obf[d] || !d_n_outlier[d]) return fail(h, CUCD_ERR_INVALID, "cucd_dev_feature_obf: null per-depth output");
    fo.numObf[d] = d_num_obf[d]; fo.nOutlier[d] = d_n_outlier[d]; fo.cuPicStride[d] = (long long)h->cuCount[d];
  }
  CK(launch_feature_obf(fp, nPics, d_thr, fo, (cudaStream_t)stream, &h->launches));
  if (d_ctu_src_had) CK(launch_ctu_src_had(fp, nPics, d_ctu_src_had, (cudaStream_t)stream, &h->launches));
  flush_launches(h);
  return CUCD_OK;
}

int cucd_dev_frames(cucd_handle* h, void* stream, int nPics, const int16_t* d_org, long long orgPicStride, int orgStride,
                    const int16_t* d_rec, long long recPicStride, int recStride, const cucd_dev_out* out, double* yc_host) {
  if (!h || nPics < 1 || nPics > h->cfg.max_pictures || !d_org || !out) return fail(h, CUCD_ERR_INVALID, "cucd_dev_frames: bad argument (nPics must be <= max_pictures)");
  if ((orgStride & 7) || (orgPicStride & 7) || ((uintptr_t)d_org & 15)) return fail(h, CUCD_ERR_INVALID, "cucd_dev_frames: source plane must be 16-byte aligned with strides multiple of 8");
  if ((out->rmd_cost != nullptr) != (d_rec != nullptr)) return fail(h, CUCD_ERR_INVALID, "cucd_dev_frames: d_rec and rmd_cost go together");
  CK(cudaSetDevice(h->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;
  const int W = h->cfg.width, H = h->cfg.height;
  const bool wantFeat = out->obf || out->outlier || out->ctu_src_had || yc_host;
  const FeaturePlanes fp = make_feature_planes(h, d_org, orgPicStride, orgStride);
  // The feature path (two small streaming kernels around a host fit) runs on the library's high-priority stream beside the RMD
  // kernel instead of in front of / behind it: fork from the caller's stream here, join before returning.  Its CTAs slip in as
  // RMD CTAs retire, so a step costs about the RMD launch alone.
  cudaStream_t sf = st;
  if (wantFeat && d_rec) {
    sf = h->sFeat;
    CK(cudaEventRecord(h->evFork, st));
    CK(cudaStreamWaitEvent(sf, h->evFork, 0));
  }
  if (wantFeat) {
    CK(launch_feature_hist(fp, nPics, h->dHist.p, sf, &h->launches));
    CK(cudaMemcpyAsync(h->hHist.p, h->dHist.p, (size_t)nPics * kHistFreqs * kHistBins * sizeof(uint32_t), cudaMemcpyDeviceToHost, sf));
    CK(cudaEventRecord(h->evHist, sf));
  }
  if (d_rec) {
    const FrameSource fs = make_frame_source(h, d_org, orgPicStride, orgStride, d_rec, recPicStride, recStride, out->rmd_cost);
    const int slot = (int)(h->rmdCalls % cucd_handle::kTimeRing);
    CK(cudaEventRecord(h->evRmd0[slot], st));
    CK(launch_rmd_auto(h, fs, nPics, st));
    CK(cudaEventRecord(h->evRmd1[slot], st));
    h->rmdCalls++;
  }
  if (wantFeat) {
    CK(cudaEventSynchronize(h->evHist));      // the RMD kernel keeps the GPU busy while the host fits
    const int nBlocks = (W / 4) * (H / 4);
    std::vector<double> yc((size_t)nPics * 16, 0.0);
    parallel_for(nPics * 15, h->hostThreads, [&](int i) {
      const int p = i / 15, f = 1 + i % 15;
      const double y = tcm_fit_one(h->hHist.p + ((size_t)p * kHistFreqs + f) * kHistBins, nBlocks);
      yc[(size_t)p * 16 + f] = y;
      h->hThr.p[p * kHistFreqs + f] = (int32_t)(y * 8.0);
    });
    for (int p = 0; p < nPics; p++) h->hThr.p[p * kHistFreqs] = 0;
    if (yc_host) memcpy(yc_host, yc.data(), yc.size() * sizeof(double));
    CK(cudaMemcpyAsync(h->dThr.p, h->hThr.p, (size_t)nPics * kHistFreqs * sizeof(int32_t), cudaMemcpyHostToDevice, sf));
    FeatureOut fo;
    fo.obf = out->obf ? out->obf : h->dObf.p; fo.obfPicStride = (long long)(W / 4) * (H / 4);
    fo.outlier = out->outlier ? out->outlier : h->dOutlier.p; fo.outlierPicStride = (long long)W * H;
    for (int d = 0; d < 4; d++) {
      fo.numObf[d] = out->num_obf[d] ? out->num_obf[d] : h->dNum[d].p;
      fo.nOutlier[d] = out->n_outlier[d] ? out->n_outlier[d] : h->dSum[d].p;
      fo.cuPicStride[d] = (long long)h->cuCount[d];
    }
    CK(launch_feature_obf(fp, nPics, h->dThr.p, fo, sf, &h->launches));
    if (out->ctu_src_had) CK(launch_ctu_src_had(fp, nPics, out->ctu_src_had, sf, &h->launches));
    if (sf != st) {
      CK(cudaEventRecord(h->evJoin, sf));
      CK(cudaStreamWaitEvent(st, h->evJoin, 0));
    }
  }
  flush_launches(h);
  return CUCD_OK;
}

// ------------------------------------------------------------------------------------------------
// S1/S4 (+ replay S2) with host buffers
// ------------------------------------------------------------------------------------------------
static int frames_group(cucd_handle* h, int nPics, const int16_t* const* orgY, int strideY, const int16_t* const* recY, int strideRec,
                        cucd_frame_out* outs) {
  const int W = h->cfg.width, H = h->cfg.height;
  // CUCD_TRACE=1: host-side timeline of the call on stderr (development aid)
  static const bool trace = getenv("CUCD_TRACE") != nullptr;
  const auto t0 = std::chrono::steady_clock::now();
  auto stamp = [&](const char* what) {
    if (trace) fprintf(stderr, "[cucd] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
  };
  bool wantRmd = false;
  bool wantPacked = false;
  if (recY) for (int p = 0; p < nPics; p++) { wantRmd = wantRmd || outs[p].rmd_cost != nullptr || outs[p].rmd_cost_packed != nullptr; wantPacked = wantPacked || outs[p].rmd_cost_packed != nullptr; }
  // ---- three-stage pipeline over sub-groups of pictures: upload (sUp) -> RMD + pack (sGrp[0]) -> cost-table
  //      download (sGrp[1]); PCIe is full duplex, so uploads, kernels and the (dominant) downloads overlap.
  //      The feature path (sFeat) needs every source plane and runs beside it. ---------------------------------
  const size_t perPic = (size_t)h->ctusPerPic * kPusPerCtu * kNumModes;
  const size_t perPicPacked = (size_t)h->ctusPerPic * CUCD_PACKED_CTU_BYTES;
  static const int nGroups = [] { const char* e = getenv("CUCD_GROUPS"); const int g = e ? atoi(e) : 8; return std::min(std::max(g, 1), (int)cucd_handle::kGroups); }();
  const int grp = std::max(1, (nPics + nGroups - 1) / nGroups);
  int gi = 0;
  for (int first = 0; first < nPics; first += grp, gi++) {
    const int n = std::min(grp, nPics - first);
    for (int p = first; p < first + n; p++) {
      CK(cudaMemcpy2DAsync(h->dOrg.p + (size_t)p * h->planeSamples, (size_t)h->pitch * 2, orgY[p], (size_t)strideY * 2, (size_t)W * 2, H,
                           cudaMemcpyHostToDevice, h->sUp));
      if (wantRmd)
        CK(cudaMemcpy2DAsync(h->dRec.p + (size_t)p * h->planeSamples, (size_t)h->pitch * 2, recY[p], (size_t)strideRec * 2, (size_t)W * 2, H,
                             cudaMemcpyHostToDevice, h->sUp));
    }
    CK(cudaEventRecord(h->evUpG[gi], h->sUp));
    if (!wantRmd) continue;
    CK(cudaStreamWaitEvent(h->sGrp[0], h->evUpG[gi], 0));
    const FrameSource fs = make_frame_source(h, h->dOrg.p + (size_t)first * h->planeSamples, (long long)h->planeSamples, h->pitch,
                                             h->dRec.p + (size_t)first * h->planeSamples, (long long)h->planeSamples, h->pitch,
                                             h->dCost.p + (size_t)first * perPic);
    CK(launch_rmd_auto(h, fs, n, h->sGrp[0]));
    if (wantPacked) CK(launch_pack_costs(h->dCost.p + (size_t)first * perPic, h->dCostPacked.p + (size_t)first * perPicPacked, n * h->ctusPerPic, h->sGrp[0], &h->launches));
    CK(cudaEventRecord(h->evRmdG[gi], h->sGrp[0]));
    CK(cudaStreamWaitEvent(h->sGrp[1], h->evRmdG[gi], 0));
    for (int p = first; p < first + n; p++) {
      if (outs[p].rmd_cost) CK(cudaMemcpyAsync(outs[p].rmd_cost, h->dCost.p + p * perPic, perPic * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->sGrp[1]));
      if (outs[p].rmd_cost_packed) CK(cudaMemcpyAsync(outs[p].rmd_cost_packed, h->dCostPacked.p + p * perPicPacked, perPicPacked, cudaMemcpyDeviceToHost, h->sGrp[1]));
    }
  }
  CK(cudaEventRecord(h->evUp, h->sUp));
  // ---- feature pass 1 on sFeat -----------------------------------------------------------------
  CK(cudaStreamWaitEvent(h->sFeat, h->evUp, 0));
  const FeaturePlanes fp = make_feature_planes(h, h->dOrg.p, (long long)h->planeSamples, h->pitch);
  CK(launch_feature_hist(fp, nPics, h->dHist.p, h->sFeat, &h->launches));
  // histograms to the host by an SM copy (pinned memory is device-visible under UVA): a cudaMemcpyAsync would wait in the
  // copy-engine queue behind the cost-table downloads and delay the TCM fit, i.e. the whole feature path
  CK(launch_copy_words(h->dHist.p, h->hHist.p, (size_t)nPics * kHistFreqs * kHistBins, h->sFeat, &h->launches));
  CK(cudaEventRecord(h->evHist, h->sFeat));
  stamp("enqueued uploads + RMD");
  // ---- host: TCM fit per picture and frequency -------------------------------------------------
  CK(cudaEventSynchronize(h->evHist));
  stamp("histograms on host");
  const int nBlocks = (W / 4) * (H / 4);
  std::vector<double> yc((size_t)nPics * 16, 0.0);
  parallel_for(nPics * 15, h->hostThreads, [&](int i) {
    const int p = i / 15, f = 1 + i % 15;
    const double y = tcm_fit_one(h->hHist.p + ((size_t)p * kHistFreqs + f) * kHistBins, nBlocks);
    yc[(size_t)p * 16 + f] = y;
    h->hThr.p[p * kHistFreqs + f] = (int32_t)(y * 8.0);
  });
  for (int p = 0; p < nPics; p++) { h->hThr.p[p * kHistFreqs] = 0; if (outs[p].yc) memcpy(outs[p].yc, &yc[(size_t)p * 16], 16 * sizeof(double)); }
  stamp("TCM fits done");
  // ---- feature pass 2 on sFeat -----------------------------------------------------------------
  CK(cudaMemcpyAsync(h->dThr.p, h->hThr.p, (size_t)nPics * kHistFreqs * sizeof(int32_t), cudaMemcpyHostToDevice, h->sFeat));
  FeatureOut fo;
  fo.obf = h->dObf.p; fo.obfPicStride = (long long)(W / 4) * (H / 4);
  fo.outlier = h->dOutlier.p; fo.outlierPicStride = (long long)W * H;
  for (int d = 0; d < 4; d++) { fo.numObf[d] = h->dNum[d].p; fo.nOutlier[d] = h->dSum[d].p; fo.cuPicStride[d] = (long long)h->cuCount[d]; }
  CK(launch_feature_obf(fp, nPics, h->dThr.p, fo, h->sFeat, &h->launches));
  CK(launch_ctu_src_had(fp, nPics, h->dCtuHad.p, h->sFeat, &h->launches));
  for (int p = 0; p < nPics; p++) {
    const cucd_frame_out& o = outs[p];
    if (o.obf) CK(cudaMemcpyAsync(o.obf, h->dObf.p + (size_t)p * fo.obfPicStride, (size_t)fo.obfPicStride * 2, cudaMemcpyDeviceToHost, h->sFeat));
    if (o.outlier) CK(cudaMemcpyAsync(o.outlier, h->dOutlier.p + (size_t)p * fo.outlierPicStride, (size_t)fo.outlierPicStride * 2, cudaMemcpyDeviceToHost, h->sFeat));
    for (int d = 0; d < 4; d++) {
      if (!h->cuCount[d]) continue;
      if (o.num_obf[d]) CK(cudaMemcpyAsync(o.num_obf[d], h->dNum[d].p + (size_t)p * h->cuCount[d], h->cuCount[d] * 4, cudaMemcpyDeviceToHost, h->sFeat));
      if (o.n_outlier[d]) CK(cudaMemcpyAsync(o.n_outlier[d], h->dSum[d].p + (size_t)p * h->cuCount[d], h->cuCount[d] * 4, cudaMemcpyDeviceToHost, h->sFeat));
    }
    if (o.ctu_src_had) CK(cudaMemcpyAsync(o.ctu_src_had, h->dCtuHad.p + (size_t)p * h->ctusPerPic, (size_t)h->ctusPerPic * 4, cudaMemcpyDeviceToHost, h->sFeat));
  }
  stamp("enqueued pass 2 + copies");
  CK(cudaStreamSynchronize(h->sFeat));
  stamp("feature stream done");
  CK(cudaStreamSynchronize(h->sUp));
  CK(cudaStreamSynchronize(h->sGrp[0]));
  CK(cudaStreamSynchronize(h->sGrp[1]));
  stamp("RMD streams done");
  flush_launches(h);
  return CUCD_OK;
}

int cuCUDecide_frames(cucd_handle* h, int nPics, const int16_t* const* orgY, int strideY, const int16_t* const* recY, int strideRec,
                      cucd_frame_out* outs) {
  if (!h || nPics < 1 || !orgY || !outs || strideY < h->cfg.width) return fail(h, CUCD_ERR_INVALID, "cuCUDecide_frames: bad argument");
  if (recY && strideRec < h->cfg.width) return fail(h, CUCD_ERR_INVALID, "cuCUDecide_frames: bad reconstruction stride");
  for (int p = 0; p < nPics; p++) {
    if (!orgY[p] || (recY && !recY[p])) return fail(h, CUCD_ERR_INVALID, "cuCUDecide_frames: null plane");
    if ((outs[p].rmd_cost || outs[p].rmd_cost_packed) && !recY) return fail(h, CUCD_ERR_INVALID, "cuCUDecide_frames: rmd_cost wanted but no reconstruction plane given");
  }
  CK(cudaSetDevice(h->cfg.device));
  for (int first = 0; first < nPics; first += h->cfg.max_pictures) {
    const int n = std::min(h->cfg.max_pictures, nPics - first);
    const int rc = frames_group(h, n, orgY + first, strideY, recY ? recY + first : nullptr, strideRec, outs + first);
    if (rc != CUCD_OK) return rc;
  }
  return CUCD_OK;
}

int cuCUDecide_frame(cucd_handle* h, const int16_t* orgY, int strideY, const int16_t* recY, int strideRec, int poc, cucd_frame_out* out) {
  (void)poc;   // the train/verify/test schedule keyed on POC (tools_YS.cpp:1237-1242) is host-side state
  const int16_t* o[1] = {orgY};
  const int16_t* r[1] = {recY};
  return cuCUDecide_frames(h, 1, o, strideY, recY ? r : nullptr, strideRec, out);
}

// ------------------------------------------------------------------------------------------------
// S2: batched RMD with caller-supplied borders
// ------------------------------------------------------------------------------------------------
int cucd_intra_rmd_batch(cucd_handle* h, int nPU, const cucd_pu_desc* desc, const int16_t* org, const int16_t* border, uint32_t* sad) {
  if (!h || nPU < 0 || (nPU > 0 && (!desc || !org || !border || !sad))) return fail(h, CUCD_ERR_INVALID, "cucd_intra_rmd_batch: bad argument");
  if (nPU == 0) return CUCD_OK;
  CK(cudaSetDevice(h->cfg.device));
  // bucket by size; offsets follow the caller's back-to-back packing
  std::vector<BatchPu> pus[7];
  size_t orgOff = 0, borderOff = 0;
  for (int i = 0; i < nPU; i++) {
    const int l = desc[i].log2_size;
    if (l < 2 || l > 6) return fail(h, CUCD_ERR_INVALID, "cucd_intra_rmd_batch: log2_size must be 2..6");
    const int n = 1 << l;
    BatchPu b; b.orgOff = (int32_t)orgOff; b.borderOff = (int32_t)borderOff; b.outIndex = i; b.pad = 0;
    pus[l].push_back(b);
    orgOff += (size_t)n * n; borderOff += (size_t)4 * n + 1;
    if (orgOff > 0x7fffffffull) return fail(h, CUCD_ERR_INVALID, "cucd_intra_rmd_batch: batch too large");
  }
  std::vector<BatchPu> all; all.reserve(nPU);
  size_t first[7] = {0};
  for (int l = 2; l <= 6; l++) { first[l] = all.size(); all.insert(all.end(), pus[l].begin(), pus[l].end()); }
  // one pinned staging block, one upload: [source blocks | borders | PU records], every part 16-byte aligned; one pinned block back.
  // (three pageable copies in, one out cost ~30 us per call - most of a small request's latency)
  auto up16 = [](size_t v) { return (v + 15) & ~(size_t)15; };
  const size_t orgBytes = up16((orgOff + 64) * 2), brdBytes = up16((borderOff + 8) * 2), puBytes = up16(all.size() * sizeof(BatchPu));
  const size_t inBytes = orgBytes + brdBytes + puBytes, outWords = (size_t)nPU * kNumModes;
  CK(h->hStage.reserve(std::max(inBytes, outWords * sizeof(uint32_t)))); CK(h->bStage.reserve(inBytes)); CK(h->bOut.reserve(outWords));
  memcpy(h->hStage.p, org, orgOff * 2);
  memcpy(h->hStage.p + orgBytes, border, borderOff * 2);
  memcpy(h->hStage.p + orgBytes + brdBytes, all.data(), all.size() * sizeof(BatchPu));
  CK(cudaMemcpyAsync(h->bStage.p, h->hStage.p, inBytes, cudaMemcpyHostToDevice, h->sMain));
  const int16_t* dOrgB = reinterpret_cast<const int16_t*>(h->bStage.p);
  const int16_t* dBrdB = reinterpret_cast<const int16_t*>(h->bStage.p + orgBytes);
  const BatchPu* dPus = reinterpret_cast<const BatchPu*>(h->bStage.p + orgBytes + brdBytes);
  CK(cudaEventRecord(h->evK0, h->sMain));
  for (int l = 6; l >= 2; l--) {
    if (pus[l].empty()) continue;
    BatchSource bs;
    bs.org = dOrgB; bs.border = dBrdB; bs.pus = dPus + first[l]; bs.out = h->bOut.p; bs.count = (int)pus[l].size();
    if (h->useTensor == 1)
      CK(launch_rmd_batch_tc2(l, bs, h->cfg.strong_intra_smoothing, h->dTc2Tables.p, h->dTc2Tables.p + tc2::kWinTableBytes, h->dHadamard.p, h->sMain, &h->launches));
    else
      CK(launch_rmd_batch(l, bs, h->cfg.bit_depth, h->cfg.strong_intra_smoothing, h->sMain, &h->launches));
  }
  CK(cudaEventRecord(h->evK1, h->sMain)); h->kTimed = true;
  CK(cudaMemcpyAsync(h->hStage.p, h->bOut.p, outWords * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->sMain));
  CK(cudaStreamSynchronize(h->sMain));
  memcpy(sad, h->hStage.p, outWords * sizeof(uint32_t));
  flush_launches(h);
  return CUCD_OK;
}

// ------------------------------------------------------------------------------------------------
// S2, asynchronous and coalescing: a worker thread turns everything that is pending into one batch
// ------------------------------------------------------------------------------------------------
struct cucd_queue {
  struct Request {
    uint64_t ticket; int nPU;
    std::vector<cucd_pu_desc> desc; std::vector<int16_t> org, border;
    uint32_t* sad;
  };
  cucd_handle* h = nullptr;
  std::mutex m;
  std::condition_variable cvWork, cvDone;
  std::deque<Request> pending;
  uint64_t nextTicket = 1, doneUpTo = 0;
  int lastStatus = CUCD_OK;
  uint64_t failedFrom = 0;                 // first ticket of a failed batch (0 = none)
  bool stop = false;
  long long requests = 0, pus = 0, batches = 0;
  std::thread worker;

  void run() {
    std::vector<cucd_pu_desc> desc; std::vector<int16_t> org, border; std::vector<uint32_t> sad;
    for (;;) {
      std::deque<Request> batch;
      {
        std::unique_lock<std::mutex> lk(m);
        cvWork.wait(lk, [&] { return stop || !pending.empty(); });
        if (pending.empty()) return;       // stop requested and nothing left
        batch.swap(pending);
      }
      desc.clear(); org.clear(); border.clear();
      int total = 0;
      for (const Request& r : batch) {
        desc.insert(desc.end(), r.desc.begin(), r.desc.end());
        org.insert(org.end(), r.org.begin(), r.org.end());
        border.insert(border.end(), r.border.begin(), r.border.end());
        total += r.nPU;
      }
      sad.resize((size_t)total * kNumModes);
      const int rc = total ? cucd_intra_rmd_batch(h, total, desc.data(), org.data(), border.data(), sad.data()) : CUCD_OK;
      size_t off = 0;
      for (const Request& r : batch) {
        if (rc == CUCD_OK) memcpy(r.sad, sad.data() + off, (size_t)r.nPU * kNumModes * sizeof(uint32_t));
        off += (size_t)r.nPU * kNumModes;
      }
      {
        std::lock_guard<std::mutex> lk(m);
        if (rc != CUCD_OK && !failedFrom) { failedFrom = batch.front().ticket; lastStatus = rc; }
        doneUpTo = batch.back().ticket;
        batches++;
      }
      cvDone.notify_all();
    }
  }
};

int cucd_queue_create(cucd_handle* h, cucd_queue** out) {
  if (!h || !out) return fail(h, CUCD_ERR_INVALID, "cucd_queue_create: null argument");
  cucd_queue* q = new cucd_queue;
  q->h = h;
  q->worker = std::thread([q] { q->run(); });
  *out = q;
  return CUCD_OK;
}

int cucd_queue_destroy(cucd_queue* q) {
  if (!q) return CUCD_OK;
  { std::lock_guard<std::mutex> lk(q->m); q->stop = true; }
  q->cvWork.notify_all();
  if (q->worker.joinable()) q->worker.join();
  delete q;
  return CUCD_OK;
}

int cucd_queue_submit(cucd_queue* q, int nPU, const cucd_pu_desc* desc, const int16_t* org, const int16_t* border, uint32_t* sad, uint64_t* ticket) {
  if (!q || !ticket || nPU < 0 || (nPU > 0 && (!desc || !org || !border || !sad))) return CUCD_ERR_INVALID;
  cucd_queue::Request r;
  r.nPU = nPU; r.sad = sad;
  size_t orgN = 0, brdN = 0;
  for (int i = 0; i < nPU; i++) {
    const int l = desc[i].log2_size;
    if (l < 2 || l > 6) return CUCD_ERR_INVALID;
    orgN += (size_t)1 << (2 * l); brdN += ((size_t)4 << l) + 1;
  }
  r.desc.assign(desc, desc + nPU); r.org.assign(org, org + orgN); r.border.assign(border, border + brdN);
  {
    std::lock_guard<std::mutex> lk(q->m);
    if (q->stop) return CUCD_ERR_INVALID;
    r.ticket = *ticket = q->nextTicket++;
    q->requests++; q->pus += nPU;
    q->pending.push_back(std::move(r));
  }
  q->cvWork.notify_one();
  return CUCD_OK;
}

int cucd_queue_wait(cucd_queue* q, uint64_t ticket) {
  if (!q || ticket == 0) return CUCD_ERR_INVALID;
  std::unique_lock<std::mutex> lk(q->m);
  if (ticket >= q->nextTicket) return CUCD_ERR_INVALID;
  q->cvDone.wait(lk, [&] { return q->doneUpTo >= ticket; });
  return (q->failedFrom && ticket >= q->failedFrom) ? q->lastStatus : CUCD_OK;
}

int cucd_queue_stats(cucd_queue* q, long long* requests, long long* pus, long long* batches) {
  if (!q) return CUCD_ERR_INVALID;
  std::lock_guard<std::mutex> lk(q->m);
  if (requests) *requests = q->requests;
  if (pus) *pus = q->pus;
  if (batches) *batches = q->batches;
  return CUCD_OK;
}

// ------------------------------------------------------------------------------------------------
// S3: integer ME SAD surfaces
// ------------------------------------------------------------------------------------------------
int cucd_set_ref_picture(cucd_handle* h, int ref_idx, const int16_t* recY, int stride, int marginX, int marginY) {
  if (!h || ref_idx < 0 || ref_idx >= 64 || !recY || marginX < 0 || marginY < 0 || stride < h->cfg.width + 2 * marginX)
    return fail(h, CUCD_ERR_INVALID, "cucd_set_ref_picture: bad argument");
  CK(cudaSetDevice(h->cfg.device));
  if ((int)h->refs.size() <= ref_idx) h->refs.resize(ref_idx + 1);
  RefPlane& r = h->refs[ref_idx];
  const int pw = h->cfg.width + 2 * marginX, ph = h->cfg.height + 2 * marginY;
  const int pitch = (pw + 7) & ~7;
  CK(r.buf.reserve((size_t)pitch * ph));
  CK(cudaMemcpy2DAsync(r.buf.p, (size_t)pitch * 2, recY - (ptrdiff_t)marginY * stride - marginX, (size_t)stride * 2, (size_t)pw * 2, ph,
                       cudaMemcpyHostToDevice, h->sMain));
  CK(cudaStreamSynchronize(h->sMain));
  r.stride = pitch; r.marginX = marginX; r.marginY = marginY; r.set = true;
  return CUCD_OK;
}

int cucd_set_cur_picture(cucd_handle* h, const int16_t* orgY, int stride) {
  if (!h || !orgY || stride < h->cfg.width) return fail(h, CUCD_ERR_INVALID, "cucd_set_cur_picture: bad argument");
  CK(cudaSetDevice(h->cfg.device));
  CK(h->dCur.reserve(h->planeSamples));
  CK(cudaMemcpy2DAsync(h->dCur.p, (size_t)h->pitch * 2, orgY, (size_t)stride * 2, (size_t)h->cfg.width * 2, h->cfg.height, cudaMemcpyHostToDevice, h->sMain));
  CK(cudaStreamSynchronize(h->sMain));
  h->curStride = h->pitch; h->curSet = true;
  return CUCD_OK;
}

int cucd_me_sad_surface(cucd_handle* h, int nPU, const cucd_me_desc* desc, uint32_t* sadOut) {
  if (!h || nPU < 0 || (nPU > 0 && (!desc || !sadOut))) return fail(h, CUCD_ERR_INVALID, "cucd_me_sad_surface: bad argument");
  if (nPU == 0) return CUCD_OK;
  if (!h->curSet) return fail(h, CUCD_ERR_INVALID, "cucd_me_sad_surface: cucd_set_cur_picture not called");
  CK(cudaSetDevice(h->cfg.device));
  const int W = h->cfg.width, H = h->cfg.height;
  std::vector<MeJob> jobs(nPU);
  std::vector<int32_t> tileJob, tileIdx;
  long long total = 0;
  for (int i = 0; i < nPU; i++) {
    const cucd_me_desc& d = desc[i];
    if (d.w < 4 || d.h < 4 || d.w > 64 || d.h > 64 || (d.w & 1) || d.x < 0 || d.y < 0 || d.x + d.w > W || d.y + d.h > H || d.left > d.right || d.top > d.bottom ||
        d.sub_shift < 0 || d.sub_shift > 4)
      return fail(h, CUCD_ERR_INVALID, "cucd_me_sad_surface: bad PU / window");
    if (d.ref_idx < 0 || d.ref_idx >= (int)h->refs.size() || !h->refs[d.ref_idx].set) return fail(h, CUCD_ERR_INVALID, "cucd_me_sad_surface: reference picture not set");
    const RefPlane& r = h->refs[d.ref_idx];
    if (d.x + d.left < -r.marginX || d.y + d.top < -r.marginY || d.x + d.w - 1 + d.right > W - 1 + r.marginX || d.y + d.h - 1 + d.bottom > H - 1 + r.marginY)
      return fail(h, CUCD_ERR_INVALID, "cucd_me_sad_surface: window leaves the padded reference picture");
    MeJob& j = jobs[i];
    j.curOff = d.y * h->curStride + d.x;
    j.refOff = (d.y + r.marginY) * r.stride + d.x + r.marginX;
    j.refSlot = d.ref_idx; j.w = (int16_t)d.w; j.h = (int16_t)d.h;
    j.left = (int16_t)d.left; j.right = (int16_t)d.right; j.top = (int16_t)d.top; j.bottom = (int16_t)d.bottom;
    // the width-specialised xGetSAD* honour iSubShift, the generic xGetSAD (TComRdCost.cpp:465-491) does not
    const bool special = d.w == 4 || d.w == 8 || d.w == 12 || d.w == 16 || d.w == 24 || d.w == 32 || d.w == 48 || d.w == 64;
    j.subShift = (int16_t)(special ? d.sub_shift : 0); j.pad = 0;
    j.outOff = total;
    const int cols = d.right - d.left + 1, rows = d.bottom - d.top + 1;
    const int tileRows = h->cfg.bit_depth == 8 ? 16 : 8;     // me_sad_u8_kernel covers 32 x 16 candidates per CTA, me_sad_kernel 32 x 8
    const int tiles = ((cols + 31) / 32) * ((rows + tileRows - 1) / tileRows);
    for (int t = 0; t < tiles; t++) { tileJob.push_back(i); tileIdx.push_back(t); }
    total += (long long)cols * rows;
  }
  std::vector<const int16_t*> refPtr(h->refs.size(), nullptr);
  std::vector<int32_t> refStride(h->refs.size(), 0);
  for (size_t i = 0; i < h->refs.size(); i++) { refPtr[i] = h->refs[i].buf.p; refStride[i] = h->refs[i].stride; }
  CK(h->dRefPtr.reserve(refPtr.size())); CK(h->dRefStride.reserve(refStride.size()));
  CK(h->dJobs.reserve(jobs.size())); CK(h->dTileJob.reserve(tileJob.size())); CK(h->dTileIdx.reserve(tileIdx.size())); CK(h->dSad.reserve((size_t)total));
  CK(cudaMemcpyAsync(h->dRefPtr.p, refPtr.data(), refPtr.size() * sizeof(void*), cudaMemcpyHostToDevice, h->sMain));
  CK(cudaMemcpyAsync(h->dRefStride.p, refStride.data(), refStride.size() * 4, cudaMemcpyHostToDevice, h->sMain));
  CK(cudaMemcpyAsync(h->dJobs.p, jobs.data(), jobs.size() * sizeof(MeJob), cudaMemcpyHostToDevice, h->sMain));
  CK(cudaMemcpyAsync(h->dTileJob.p, tileJob.data(), tileJob.size() * 4, cudaMemcpyHostToDevice, h->sMain));
  CK(cudaMemcpyAsync(h->dTileIdx.p, tileIdx.data(), tileIdx.size() * 4, cudaMemcpyHostToDevice, h->sMain));
  MePlanes mp;
  mp.cur = h->dCur.p; mp.curStride = h->curStride; mp.ref = h->dRefPtr.p; mp.refStride = h->dRefStride.p; mp.bitDepth = h->cfg.bit_depth;
  CK(cudaEventRecord(h->evK0, h->sMain));
  CK(launch_me_sad(mp, h->dJobs.p, nPU, h->dTileJob.p, h->dTileIdx.p, (int)tileJob.size(), h->dSad.p, h->sMain, &h->launches));
  CK(cudaEventRecord(h->evK1, h->sMain)); h->kTimed = true;
  CK(cudaMemcpyAsync(sadOut, h->dSad.p, (size_t)total * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->sMain));
  CK(cudaStreamSynchronize(h->sMain));
  flush_launches(h);
  return CUCD_OK;
}

// ------------------------------------------------------------------------------------------------
// Fractional-pel refinement: distortion of the 49 quarter-pel positions around an integer MV
// ------------------------------------------------------------------------------------------------
int cucd_me_subpel_cost(cucd_handle* h, int nPU, const cucd_subpel_desc* desc, uint32_t* cost) {
  if (!h || nPU < 0 || (nPU > 0 && (!desc || !cost))) return fail(h, CUCD_ERR_INVALID, "cucd_me_subpel_cost: bad argument");
  if (nPU == 0) return CUCD_OK;
  if (!h->curSet) return fail(h, CUCD_ERR_INVALID, "cucd_me_subpel_cost: cucd_set_cur_picture not called");
  CK(cudaSetDevice(h->cfg.device));
  const int W = h->cfg.width, H = h->cfg.height;
  std::vector<SubpelJob> jobs(nPU);
  for (int i = 0; i < nPU; i++) {
    const cucd_subpel_desc& d = desc[i];
    if (d.w < 4 || d.h < 4 || d.w > 64 || d.h > 64 || (d.w & 3) || (d.h & 3) || d.x < 0 || d.y < 0 || d.x + d.w > W || d.y + d.h > H)
      return fail(h, CUCD_ERR_INVALID, "cucd_me_subpel_cost: bad PU");
    if (d.ref_idx < 0 || d.ref_idx >= (int)h->refs.size() || !h->refs[d.ref_idx].set) return fail(h, CUCD_ERR_INVALID, "cucd_me_subpel_cost: reference picture not set");
    const RefPlane& r = h->refs[d.ref_idx];
    if (d.x + d.mvx - 4 < -r.marginX || d.y + d.mvy - 4 < -r.marginY || d.x + d.mvx + d.w + 4 > W - 1 + r.marginX || d.y + d.mvy + d.h + 4 > H - 1 + r.marginY)
      return fail(h, CUCD_ERR_INVALID, "cucd_me_subpel_cost: the interpolation support leaves the padded reference picture");
    SubpelJob& j = jobs[i];
    j.curOff = d.y * h->curStride + d.x;
    j.refOff = (d.y + d.mvy + r.marginY) * r.stride + d.x + d.mvx + r.marginX;
    j.refSlot = d.ref_idx; j.w = (int16_t)d.w; j.h = (int16_t)d.h; j.useHadamard = d.use_hadamard ? 1 : 0;
  }
  std::vector<const int16_t*> refPtr(h->refs.size(), nullptr);
  std::vector<int32_t> refStride(h->refs.size(), 0);
  for (size_t i = 0; i < h->refs.size(); i++) { refPtr[i] = h->refs[i].buf.p; refStride[i] = h->refs[i].stride; }
  CK(h->dRefPtr.reserve(refPtr.size())); CK(h->dRefStride.reserve(refStride.size()));
  CK(h->dSubJobs.reserve(jobs.size())); CK(h->dSad.reserve((size_t)nPU * CUCD_SUBPEL_POINTS));
  CK(cudaMemcpyAsync(h->dRefPtr.p, refPtr.data(), refPtr.size() * sizeof(void*), cudaMemcpyHostToDevice, h->sMain));
  CK(cudaMemcpyAsync(h->dRefStride.p, refStride.data(), refStride.size() * 4, cudaMemcpyHostToDevice, h->sMain));
  CK(cudaMemcpyAsync(h->dSubJobs.p, jobs.data(), jobs.size() * sizeof(SubpelJob), cudaMemcpyHostToDevice, h->sMain));
  MePlanes mp;
  mp.cur = h->dCur.p; mp.curStride = h->curStride; mp.ref = h->dRefPtr.p; mp.refStride = h->dRefStride.p; mp.bitDepth = h->cfg.bit_depth;
  CK(cudaEventRecord(h->evK0, h->sMain));
  CK(launch_me_subpel(mp, h->dSubJobs.p, nPU, h->dSad.p, h->sMain, &h->launches));
  CK(cudaEventRecord(h->evK1, h->sMain)); h->kTimed = true;
  CK(cudaMemcpyAsync(cost, h->dSad.p, (size_t)nPU * CUCD_SUBPEL_POINTS * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->sMain));
  CK(cudaStreamSynchronize(h->sMain));
  flush_launches(h);
  return CUCD_OK;
}

// ------------------------------------------------------------------------------------------------
// Intra luma TU coding (xIntraCodingTUBlock): forward half, reconstruction half, or the whole chain with the plain quantiser
// ------------------------------------------------------------------------------------------------
static int tu_batch(cucd_handle* h, const char* who, int stage, int flags, int nTU, const cucd_tu_desc* desc, const int16_t* org, const int16_t* border,
                    int32_t* coefOut, const int32_t* levelIn, int16_t* pixOut, uint32_t* dist, int32_t* absSum) {
  if (!h || nTU < 0 || (nTU > 0 && (!desc || !org || !border))) return fail(h, CUCD_ERR_INVALID, std::string(who) + ": bad argument");
  if (nTU == 0) return CUCD_OK;
  CK(cudaSetDevice(h->cfg.device));
  std::vector<TuJob> jobs[6];
  size_t orgOff = 0, borderOff = 0;
  for (int i = 0; i < nTU; i++) {
    const cucd_tu_desc& d = desc[i];
    if (d.log2_size < 2 || d.log2_size > 5) return fail(h, CUCD_ERR_INVALID, std::string(who) + ": log2_size must be 2..5");
    if (d.mode > 34 || d.qp < 0 || d.qp > 51) return fail(h, CUCD_ERR_INVALID, std::string(who) + ": mode must be 0..34 and qp 0..51");
    if ((d.flags & CUCD_TU_TRANSFORM_SKIP) && d.log2_size != 2) return fail(h, CUCD_ERR_UNSUPPORTED, std::string(who) + ": transform skip is a 4x4 tool (log2MaxTransformSkipSize = 2)");
    const int n = 1 << d.log2_size;
    TuJob j; j.orgOff = (int32_t)orgOff; j.borderOff = (int32_t)borderOff; j.outIndex = i; j.mode = d.mode; j.ts = d.flags & (CUCD_TU_TRANSFORM_SKIP | CUCD_TU_CHROMA); j.qp = d.qp; j.pad = 0;
    jobs[d.log2_size].push_back(j);
    orgOff += (size_t)n * n; borderOff += (size_t)4 * n + 1;
    if (orgOff > 0x7fffffffull) return fail(h, CUCD_ERR_INVALID, std::string(who) + ": batch too large");
  }
  std::vector<TuJob> all; all.reserve(nTU);
  size_t first[6] = {0};
  for (int l = 2; l <= 5; l++) { first[l] = all.size(); all.insert(all.end(), jobs[l].begin(), jobs[l].end()); }
  CK(h->bOrg.reserve(orgOff + 64)); CK(h->bBorder.reserve(borderOff + 8)); CK(h->tJobs.reserve(all.size()));
  CK(h->tCoef.reserve(orgOff)); CK(h->tPix.reserve(orgOff)); CK(h->tDist.reserve(nTU)); CK(h->tAbs.reserve(nTU));
  CK(cudaMemcpyAsync(h->bOrg.p, org, orgOff * 2, cudaMemcpyHostToDevice, h->sMain));
  CK(cudaMemcpyAsync(h->bBorder.p, border, borderOff * 2, cudaMemcpyHostToDevice, h->sMain));
  CK(cudaMemcpyAsync(h->tJobs.p, all.data(), all.size() * sizeof(TuJob), cudaMemcpyHostToDevice, h->sMain));
  if (stage == 2) CK(cudaMemcpyAsync(h->tCoef.p, levelIn, orgOff * sizeof(int32_t), cudaMemcpyHostToDevice, h->sMain));
  CK(cudaEventRecord(h->evK0, h->sMain));
  for (int l = 5; l >= 2; l--) {
    if (jobs[l].empty()) continue;
    TuBatch tb;
    tb.org = h->bOrg.p; tb.border = h->bBorder.p; tb.jobs = h->tJobs.p + first[l]; tb.count = (int)jobs[l].size();
    tb.stage = stage; tb.bitDepth = h->cfg.bit_depth; tb.strong = h->cfg.strong_intra_smoothing;
    tb.intraSlice = (flags & CUCD_TU_INTRA_SLICE) ? 1 : 0; tb.signHiding = (flags & CUCD_TU_SIGN_HIDING) ? 1 : 0;
    tb.coef = h->tCoef.p; tb.pred = (stage == 0 && pixOut) ? h->tPix.p : nullptr; tb.reco = h->tPix.p; tb.dist = h->tDist.p; tb.absSum = h->tAbs.p;
    CK(launch_intra_tu(l, tb, h->sMain, &h->launches));
  }
  CK(cudaEventRecord(h->evK1, h->sMain)); h->kTimed = true;
  if (coefOut) CK(cudaMemcpyAsync(coefOut, h->tCoef.p, orgOff * sizeof(int32_t), cudaMemcpyDeviceToHost, h->sMain));
  if (pixOut) CK(cudaMemcpyAsync(pixOut, h->tPix.p, orgOff * sizeof(int16_t), cudaMemcpyDeviceToHost, h->sMain));
  if (dist) CK(cudaMemcpyAsync(dist, h->tDist.p, (size_t)nTU * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->sMain));
  if (absSum) CK(cudaMemcpyAsync(absSum, h->tAbs.p, (size_t)nTU * sizeof(int32_t), cudaMemcpyDeviceToHost, h->sMain));
  CK(cudaStreamSynchronize(h->sMain));
  flush_launches(h);
  return CUCD_OK;
}

int cucd_intra_tu_forward(cucd_handle* h, int nTU, const cucd_tu_desc* desc, const int16_t* org, const int16_t* border, int32_t* coef, int16_t* pred) {
  if (nTU > 0 && !coef) return fail(h, CUCD_ERR_INVALID, "cucd_intra_tu_forward: coef is NULL");
  return tu_batch(h, "cucd_intra_tu_forward", 0, 0, nTU, desc, org, border, coef, nullptr, pred, nullptr, nullptr);
}
int cucd_intra_tu_recon(cucd_handle* h, int nTU, const cucd_tu_desc* desc, const int16_t* org, const int16_t* border, const int32_t* level,
                        int16_t* reco, uint32_t* dist) {
  if (nTU > 0 && (!level || !reco || !dist)) return fail(h, CUCD_ERR_INVALID, "cucd_intra_tu_recon: NULL level / reco / dist");
  return tu_batch(h, "cucd_intra_tu_recon", 2, 0, nTU, desc, org, border, nullptr, level, reco, dist, nullptr);
}
int cucd_intra_tu_code(cucd_handle* h, int nTU, const cucd_tu_desc* desc, const int16_t* org, const int16_t* border, int flags,
                       int32_t* level, int16_t* reco, uint32_t* dist, int32_t* abs_sum) {
  if (nTU > 0 && (!level || !reco || !dist)) return fail(h, CUCD_ERR_INVALID, "cucd_intra_tu_code: NULL level / reco / dist");
  return tu_batch(h, "cucd_intra_tu_code", 1, flags, nTU, desc, org, border, level, nullptr, reco, dist, abs_sum);
}

// ------------------------------------------------------------------------------------------------
// CU texture features (getTMVFeature) and AQ activity (TEncPreanalyzer) of the current picture
// ------------------------------------------------------------------------------------------------
int cucd_tmv_features(cucd_handle* h, int nCU, const cucd_cu_desc* cus, double* feat) {
  if (!h || nCU < 0 || (nCU > 0 && (!cus || !feat))) return fail(h, CUCD_ERR_INVALID, "cucd_tmv_features: bad argument");
  if (nCU == 0) return CUCD_OK;
  if (!h->curSet) return fail(h, CUCD_ERR_INVALID, "cucd_tmv_features: cucd_set_cur_picture not called");
  CK(cudaSetDevice(h->cfg.device));
  std::vector<TmvCu> v(nCU);
  for (int i = 0; i < nCU; i++) {
    const cucd_cu_desc& c = cus[i];
    if (c.log2_size < 3 || c.log2_size > 6) return fail(h, CUCD_ERR_INVALID, "cucd_tmv_features: log2_size must be 3..6");
    const int n = 1 << c.log2_size;
    if (c.x < 0 || c.y < 0 || (c.x & (n - 1)) || (c.y & (n - 1)) || c.x + n > h->cfg.width || c.y + n > h->cfg.height)
      return fail(h, CUCD_ERR_INVALID, "cucd_tmv_features: CU must be aligned to its size and lie inside the picture");
    v[i].x = c.x; v[i].y = c.y; v[i].log2n = c.log2_size; v[i].pad = 0;
  }
  CK(h->dTmvCus.reserve(nCU)); CK(h->dDoubles.reserve((size_t)nCU * CUCD_TMV_FEATURES));
  CK(cudaMemcpyAsync(h->dTmvCus.p, v.data(), v.size() * sizeof(TmvCu), cudaMemcpyHostToDevice, h->sMain));
  CK(cudaEventRecord(h->evK0, h->sMain));
  CK(launch_tmv_features(h->dCur.p, h->curStride, h->dTmvCus.p, nCU, h->dDoubles.p, h->sMain, &h->launches));
  CK(cudaEventRecord(h->evK1, h->sMain)); h->kTimed = true;
  CK(cudaMemcpyAsync(feat, h->dDoubles.p, (size_t)nCU * CUCD_TMV_FEATURES * sizeof(double), cudaMemcpyDeviceToHost, h->sMain));
  CK(cudaStreamSynchronize(h->sMain));
  flush_launches(h);
  return CUCD_OK;
}

int cucd_aq_activity(cucd_handle* h, int max_aq_depth, double* const* activity, double* avg_activity) {
  if (!h || max_aq_depth < 1 || max_aq_depth > 4 || (!activity && !avg_activity)) return fail(h, CUCD_ERR_INVALID, "cucd_aq_activity: bad argument");
  if (!h->curSet) return fail(h, CUCD_ERR_INVALID, "cucd_aq_activity: cucd_set_cur_picture not called");
  CK(cudaSetDevice(h->cfg.device));
  const int W = h->cfg.width, H = h->cfg.height;
  AqLayers L; L.count = max_aq_depth; L.total = 0;
  for (int d = 0; d < 4; d++) { L.part[d] = 0; L.off[d] = 0; }
  for (int d = 0; d < max_aq_depth; d++) {
    L.part[d] = h->cfg.ctu_size >> d; L.off[d] = L.total;
    L.total += ((W + L.part[d] - 1) / L.part[d]) * ((H + L.part[d] - 1) / L.part[d]);
  }
  for (int d = max_aq_depth; d <= 4; d++) L.off[d] = L.total;
  CK(h->dDoubles.reserve(L.total));
  CK(cudaEventRecord(h->evK0, h->sMain));
  CK(launch_aq_activity(h->dCur.p, h->curStride, W, H, L, h->dDoubles.p, h->sMain, &h->launches));
  CK(cudaEventRecord(h->evK1, h->sMain)); h->kTimed = true;
  std::vector<double> act(L.total);
  CK(cudaMemcpyAsync(act.data(), h->dDoubles.p, (size_t)L.total * sizeof(double), cudaMemcpyDeviceToHost, h->sMain));
  CK(cudaStreamSynchronize(h->sMain));
  for (int d = 0; d < max_aq_depth; d++) {
    const int n = L.off[d + 1] - L.off[d];
    if (activity && activity[d]) memcpy(activity[d], act.data() + L.off[d], (size_t)n * sizeof(double));
    if (avg_activity) {          // dSumAct accumulates in raster order (TEncPreanalyzer.cpp:132): a sequential double sum, kept on the host
      double sum = 0.0;
      for (int i = 0; i < n; i++) sum += act[L.off[d] + i];
      avg_activity[d] = sum / (double)(unsigned)n;
    }
  }
  flush_launches(h);
  return CUCD_OK;
}

}  // extern "C"
