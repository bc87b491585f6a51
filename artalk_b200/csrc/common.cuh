// Shared device/host helpers for the artalk_b200 sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <math.h>
#include <atomic>

namespace artalk {

typedef __nv_bfloat16 bf16;

// status codes of the C ABI (include/artalk_b200.h)
enum : int { AT_OK = 0, AT_EINVAL = 1, AT_ECUDA = 2, AT_ENOMEM = 3, AT_EMISSING = 4, AT_ESTATE = 5 };

void set_last_error(const char* fmt, ...);

#define AT_CUDA(expr)                                                                        \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      ::artalk::set_last_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,                 \
                               cudaGetErrorString(_e));                                      \
      return ::artalk::AT_ECUDA;                                                             \
    }                                                                                        \
  } while (0)

// Per-device state shared by the launchers: the library may drive several GPUs of one process (one engine per device), so
// SM counts, the barrier-timeout flag the tcgen05 kernels report into, and the "dynamic shared memory opt-in done" marks are
// kept per device ordinal (a process-wide static would hand device 1 a pointer allocated on device 0).
struct DevCtx { int dev; int num_sms; unsigned int* err_flag; };
int dev_ctx(const DevCtx** out);                          // context of the CURRENT device, created on first use
int ensure_dyn_smem(const void* kernel, int bytes);       // cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (device, kernel)
int& per_device_slot(int (&slots)[16]);                    // slots[current device] (ordinals >= 16 share the last slot)

extern std::atomic<unsigned long long> g_launch_count;      // kernels launched by this library (bench.py's gpu_launches)
extern std::atomic<unsigned int> g_option_epoch;            // bumped by artalk_set_option / artalk_enable_pdl: captured graphs are stale
// optional launch trace (artalk_trace_begin/end): one CUDA event after every launch on the launching stream `st`;
// the time between consecutive events (kernel + any idle gap before it) is attributed to the launching function.
extern bool g_trace_on;
extern int g_trace_dims[3];              // set by launchers that want their problem size in the trace (GEMM, attention)
void trace_event(const char* func, cudaStream_t st);
#define AT_LAUNCH_CHECK()                                         \
  do {                                                            \
    ::artalk::g_launch_count.fetch_add(1, std::memory_order_relaxed); \
    if (::artalk::g_trace_on) ::artalk::trace_event(__func__, st); \
    AT_CUDA(cudaGetLastError());                                  \
  } while (0)

#define AT_REQUIRE(cond, ...)                                                                \
  do {                                                                                       \
    if (!(cond)) {                                                                           \
      ::artalk::set_last_error(__VA_ARGS__);                                                 \
      return ::artalk::AT_EINVAL;                                                            \
    }                                                                                        \
  } while (0)

#define AT_TRY(expr)                                                                         \
  do {                                                                                       \
    int _s = (expr);                                                                         \
    if (_s != ::artalk::AT_OK) return _s;                                                    \
  } while (0)

// row r of a logical [rows, cols] matrix lives at base + (r / rpb) * bs + (r % rpb) * rs (elements)
struct RowMap {
  int rpb;        // rows per batch (<=0: plain, offset = r * rs)
  int64_t bs;     // batch stride
  int64_t rs;     // row stride
  __host__ __device__ inline int64_t off(int r) const {
    if (rpb <= 0) return (int64_t)r * rs;
    int b = r / rpb;
    return (int64_t)b * bs + (int64_t)(r - b * rpb) * rs;
  }
};
__host__ __device__ static inline RowMap plain_rows(int64_t rs) { RowMap m; m.rpb = 0; m.bs = 0; m.rs = rs; return m; }
__host__ __device__ static inline RowMap batched_rows(int rpb, int64_t bs, int64_t rs) { RowMap m; m.rpb = rpb; m.bs = bs; m.rs = rs; return m; }

enum Act : int { ACT_NONE = 0, ACT_GELU_ERF = 1, ACT_GELU_TANH = 2, ACT_LEAKY02 = 3, ACT_SILU = 4 };

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float gelu_tanh(float x) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  float inner = k0 * (x + k1 * x * x * x);
  return 0.5f * x * (1.0f + tanhf(inner));
}
// bf16-path erf-GELU (results are rounded to bf16, spacing 2^-8 relative): erf(z) = tanh(z (c1 + c3 z^2 + c5 z^4)) with a
// minimax fit of the GELU error over |x| <= 6 (max |gelu - exact| = 6.3e-5 with an exact tanh; the hardware tanh.approx
// adds <= 2^-11 relative, i.e. <= 5e-4 absolute at |x| ~ 2 where one bf16 rounding step is 8e-3). One MUFU and 7 FP ops per
// element instead of two MUFUs and 14: the wav2vec FFN1 epilogue and the conv LN+GELU kernels are bound by this function.
// x^2 is clamped so the odd polynomial keeps its sign for large |x| (tanh saturates to +-1 there).
__device__ __forceinline__ float gelu_erf_fast(float x) {
  const float x2 = fminf(x * x, 49.0f);
  float p = fmaf(-3.8652387e-4f, x2, 3.7255297e-2f);
  p = fmaf(p, x2, 7.9716741e-1f);
  p *= x;
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(p));
  const float hx = 0.5f * x;
  return fmaf(hx, t, hx);
}
// bf16-path variant: hardware tanh.approx (rel. error ~2^-11, below bf16 rounding of the result)
__device__ __forceinline__ float gelu_tanh_fast(float x) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  float inner = k0 * (x + k1 * x * x * x), t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(inner));
  return 0.5f * x * (1.0f + t);
}
__device__ __forceinline__ float silu(float x) { return x / (1.0f + expf(-x)); }
__device__ __forceinline__ float apply_act(float x, int act) {
  switch (act) {
    case ACT_GELU_ERF: return gelu_erf(x);
    case ACT_GELU_TANH: return gelu_tanh(x);
    case ACT_LEAKY02: return x > 0.f ? x : 0.2f * x;
    case ACT_SILU: return silu(x);
    default: return x;
  }
}

__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f32<bf16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// block-wide sum for blockDim.x <= 1024 (multiple of 32); `red` is >= 32 floats of shared memory
__device__ __forceinline__ float block_sum(float v, float* red) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float t = (threadIdx.x < nw) ? red[threadIdx.x] : 0.f;
  if (w == 0) {
    t = warp_sum(t);
    if (lane == 0) red[0] = t;
  }
  __syncthreads();
  return red[0];
}

// 4-wide vector access helpers (fp32: 16 B, bf16: 8 B)
__device__ __forceinline__ void load4(const float* p, float (&v)[4]) {
  float4 t = *reinterpret_cast<const float4*>(p);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ __forceinline__ void load4(const bf16* p, float (&v)[4]) {
  uint2 t = *reinterpret_cast<const uint2*>(p);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&t.x), b = *reinterpret_cast<__nv_bfloat162*>(&t.y);
  v[0] = __low2float(a); v[1] = __high2float(a); v[2] = __low2float(b); v[3] = __high2float(b);
}
__device__ __forceinline__ void store4(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void store4(bf16* p, const float (&v)[4]) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
  uint2 t;
  t.x = *reinterpret_cast<uint32_t*>(&a);
  t.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = t;
}

// ---- programmatic dependent launch (PDL): kernels are launched with cudaLaunchAttributeProgrammaticStreamSerialization
// (launch_k below), so a kernel's prologue overlaps the tail of its predecessor in the stream / CUDA graph. Every kernel
// launched that way calls pdl_launch_dependents() early and pdl_wait() before it touches memory written by earlier
// kernels (both are no-ops for a normally launched kernel); because every kernel waits before it finishes, completion
// stays transitive along the stream.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter() { pdl_launch_dependents(); pdl_wait(); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

extern bool g_pdl;        // host switch (artalk_enable_pdl); default on
extern int g_pdl_w2v_max_chunks;   // wav2vec sub-batches larger than this run without PDL (option "pdl_w2v_max_chunks")
extern int g_skinny_tokens;   // option "skinny_tokens" (engine.cu)
extern int g_attn_split;      // option "attn_split" (engine.cu)
extern int g_conv0_fold;      // option "conv0_fold" (engine.cu)
extern int g_posconv4;        // option "posconv4" (engine.cu)
extern int g_attn_bound;      // option "attn_bound" (engine.cu)
extern int g_w2v_graph_chunks;   // option "w2v_graph_chunks" (engine.cu)
extern int g_pdl_mask;    // per kernel class (option "pdl_mask"): 1 = tcgen05 GEMM, 2 = tcgen05 attention, 4 = everything else
#ifndef ARTALK_PDL_CLASS
#define ARTALK_PDL_CLASS 4
#endif
static inline bool pdl_on() { return g_pdl && (g_pdl_mask & ARTALK_PDL_CLASS); }
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl_on() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline int64_t align_up(int64_t a, int64_t b) { return (a + b - 1) / b * b; }

}  // namespace artalk
