// Savitzky-Golay smoothing + output post-ops of ARTAvatarInferEngine.inference (inference.py:52-56,89-95) on the device:
// window 5 / poly 2 on every dim, window 9 / poly 3 on dims 100:103, scipy mode 'interp' (edge samples come from
// the polynomial fitted to the first / last window). The hat matrices H5, H9 (row i = weights giving the fitted
// value at window position i) are built on the host in fp64 and passed by value.
#include <mutex>
#include "kernels.cuh"

namespace artalk {

struct SavgolTables { float h5[5][5]; float h9[9][9]; };
static SavgolTables g_tables;
static std::atomic<bool> g_tables_set{false};
static std::mutex g_tables_mu;      // the tables are constants of the filter; concurrent engines may race to set them

void set_savgol_tables(const float* h5, const float* h9) {
  std::lock_guard<std::mutex> lk(g_tables_mu);
  for (int i = 0; i < 25; ++i) (&g_tables.h5[0][0])[i] = h5[i];
  for (int i = 0; i < 81; ++i) (&g_tables.h9[0][0])[i] = h9[i];
  g_tables_set.store(true, std::memory_order_release);
}

template <int W>
__device__ __forceinline__ float savgol_at(const float* __restrict__ x, int64_t stride, int t, int T, const float (*H)[W]) {
  constexpr int HALF = W / 2;
  int row, start;
  if (t < HALF) { row = t; start = 0; }
  else if (t >= T - HALF) { row = W - (T - t); start = T - W; }
  else { row = HALF; start = t - HALF; }
  float a = 0.f;
#pragma unroll
  for (int j = 0; j < W; ++j) a = fmaf(H[row][j], x[(int64_t)(start + j) * stride], a);
  return a;
}

__global__ void __launch_bounds__(256) savgol_post_kernel(const float* __restrict__ motion, float* __restrict__ out, int T,
                                                          int T_out, int dim, int fix_pose, int zero_tail, SavgolTables tb, int64_t total) {
  pdl_enter();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % dim);
    int64_t rt = i / dim;
    int t = (int)(rt % T_out), clip = (int)(rt / T_out);
    const float* x = motion + (int64_t)clip * T * dim + c;
    float v;
    if (c >= 104 && zero_tail) v = 0.f;                      // inference.py:56
    else if (c >= 100 && c < 103) v = fix_pose ? 0.f : savgol_at<9>(x, dim, t, T, tb.h9);
    else v = savgol_at<5>(x, dim, t, T, tb.h5);
    out[i] = v;
  }
}

int launch_savgol_post(const float* motion, float* out, int n_clips, int T, int T_out, int dim, int fix_pose,
                       int zero_tail, cudaStream_t st) {
  if (n_clips <= 0 || T_out <= 0) return AT_OK;
  AT_REQUIRE(g_tables_set.load(std::memory_order_acquire), "savgol: tables not set");
  AT_REQUIRE(T >= 9, "savgol: window_length 9 must be <= number of frames (%d)", T);
  AT_REQUIRE(T_out <= T && dim == 106, "savgol: bad shape");
  int64_t total = (int64_t)n_clips * T_out * dim;
  int grid = (int)((total + 255) / 256);
  if (grid > 148 * 16) grid = 148 * 16;
  AT_CUDA(launch_k(savgol_post_kernel, dim3(grid), dim3(256), 0, st, motion, out, T, T_out, dim, fix_pose, zero_tail, g_tables, total));
  AT_LAUNCH_CHECK();
  return AT_OK;
}

// ---------------------------------------------------------------- forehead EMA of the GAGAvatar point builder
// app/GAGAvatar/models.py:120-125, run frame by frame by the reference: the first frame ever initialises the state with its
// own points, every later frame does u <- 0.98 u + 0.02 c and the frame's forehead vertices are REPLACED by u. Batched here
// as a scan over the frames of one call: one thread per (forehead vertex, coordinate), sequential over frames (a first-order
// IIR), state carried across calls in `state` [n_idx][3]. The recurrence is sequential but the loads are not: every thread requests
// the values of the next 32 frames (60 KB apart: one HBM / L2 line each) before it walks them, so a call costs one memory
// round trip per 32 frames instead of one per frame (16 000 frames: 8.7 ms -> measured in profiles/).
__global__ void __launch_bounds__(128) ema_scan_kernel(float* __restrict__ points, int64_t frame_stride, const int* __restrict__ idx,
                                                       int n_idx, int n_frames, float* __restrict__ state, int has_state, float keep) {
  pdl_enter();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_idx * 3) return;
  const int v = idx[t / 3], c = t % 3;
  float* p = points + (int64_t)v * 3 + c;
  float u = has_state ? state[t] : 0.f;
  constexpr int PF = 32;
  for (int f0 = 0; f0 < n_frames; f0 += PF) {
    float cur[PF];
#pragma unroll
    for (int i = 0; i < PF; ++i) cur[i] = (f0 + i < n_frames) ? p[(int64_t)i * frame_stride] : 0.f;       // independent loads
#pragma unroll
    for (int i = 0; i < PF; ++i) {
      if (f0 + i >= n_frames) break;
      if (f0 + i == 0 && !has_state) u = cur[i];             // models.py:120-121: no blending on the very first frame
      else { u = keep * u + (1.0f - keep) * cur[i]; p[(int64_t)i * frame_stride] = u; }     // models.py:123-125
    }
    p += (int64_t)PF * frame_stride;
  }
  state[t] = u;
}

int launch_ema_scan(float* points, int64_t frame_stride, const int* idx, int n_idx, int n_frames, float* state, int has_state,
                    float keep, cudaStream_t st) {
  if (n_idx <= 0 || n_frames <= 0) return AT_OK;
  AT_REQUIRE(points && idx && state && keep >= 0.f && keep <= 1.f, "ema_scan: bad argument");
  AT_CUDA(launch_k(ema_scan_kernel, dim3(ceil_div(n_idx * 3, 128)), dim3(128), 0, st, points, frame_stride, idx, n_idx, n_frames, state,
                   has_state, keep));
  AT_LAUNCH_CHECK();
  return AT_OK;
}

}  // namespace artalk
