// Memory-bound row kernels: LayerNorm variants, AdaLN modulate, audio normalisation statistics,
// conv layer 0 (+LN+GELU), multi-scale audio pooling, activation/cast.
// One warp owns one row; every global access is a 16-byte (fp32) / 8-byte (bf16) vector, 512 B per warp.
#include "kernels.cuh"

namespace artalk {

// ---------------------------------------------------------------- LayerNorm
// V4 = float4s per lane -> cols = 128 * V4
template <int V4, typename TO>
__global__ void __launch_bounds__(128) ln_affine_kernel(const float* __restrict__ x, int64_t x_rs, TO* __restrict__ out,
                                                        int64_t out_rs, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, int rows, float eps, int act) {
  pdl_enter();
  int row = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + (int64_t)row * x_rs;
  float v[V4][4];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < V4; ++i) {
    load4(xr + (i * 32 + lane) * 4, v[i]);
    s += (v[i][0] + v[i][1]) + (v[i][2] + v[i][3]);
  }
  const float inv_n = 1.0f / (float)(V4 * 128);
  float mean = warp_sum(s) * inv_n;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < V4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { float d = v[i][j] - mean; q += d * d; }
  float rstd = rsqrtf(warp_sum(q) * inv_n + eps);
  TO* orow = out + (int64_t)row * out_rs;
#pragma unroll
  for (int i = 0; i < V4; ++i) {
    int c = (i * 32 + lane) * 4;
    float g[4] = {1.f, 1.f, 1.f, 1.f}, b[4] = {0.f, 0.f, 0.f, 0.f}, o[4];
    if (gamma) load4(gamma + c, g);
    if (beta) load4(beta + c, b);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float y = (v[i][j] - mean) * rstd * g[j] + b[j];
      o[j] = (sizeof(TO) == 2 && act == ACT_GELU_ERF) ? gelu_erf_fast(y) : apply_act(y, act);   // bf16 output: fast erf
    }
    store4(orow + c, o);
  }
}

template <int V4, typename TO>
static int ln_launch(const float* x, int64_t x_rs, void* out, int64_t out_rs, const float* g, const float* b, int rows,
                     float eps, int act, cudaStream_t st) {
  AT_CUDA(launch_k(ln_affine_kernel<V4, TO>, dim3(ceil_div(rows, 4)), dim3(128), 0, st, x, x_rs, (TO*)out, out_rs, g, b, rows, eps, act));
  AT_LAUNCH_CHECK();
  return AT_OK;
}

int launch_layernorm(const float* x, int64_t x_rs, void* out, int out_dt, int64_t out_rs, const float* gamma,
                     const float* beta, int rows, int cols, float eps, int act, cudaStream_t st) {
  if (rows <= 0) return AT_OK;
  AT_REQUIRE(cols == 128 || cols == 512 || cols == 768 || cols == 1024, "layernorm: unsupported width %d", cols);
  AT_REQUIRE(x_rs % 4 == 0 && out_rs % 4 == 0, "layernorm: row strides must be multiples of 4");
#define LN_CASE(V4)                                                                                          \
  return out_dt == DT_F32 ? ln_launch<V4, float>(x, x_rs, out, out_rs, gamma, beta, rows, eps, act, st)      \
                          : ln_launch<V4, bf16>(x, x_rs, out, out_rs, gamma, beta, rows, eps, act, st)
  switch (cols) {
    case 128: LN_CASE(1);
    case 512: LN_CASE(4);
    case 768: LN_CASE(6);
    default: LN_CASE(8);
  }
#undef LN_CASE
}

// ---------------------------------------------------------------- AdaLN modulate
// Optionally folds the previous sub-layer's gated residual update in first (app/transformer.py:37,41):
//   x[r] <- fma(gate[r], y[r], x[r])   with y = the fp32 output of the projection / FFN GEMM, gate = ada[map(r)][gate_off + c]
// so those GEMMs run with the plain epilogue (no gate / residual loads per output row) and the read-modify-write of the fp32
// residual stream happens here, one warp per row with 16-byte accesses. Same arithmetic (one fma per element) as the fused
// GEMM epilogue, so results are bit-identical.
template <int V4, typename TA, typename TO>
__global__ void __launch_bounds__(128) adaln_kernel(float* __restrict__ x, const TA* __restrict__ ada, RowMap ada_map,
                                                    int scale_off, int shift_off, TO* __restrict__ out, int rows,
                                                    float eps, const float* __restrict__ y, int gate_off) {
  pdl_enter();
  constexpr int C = V4 * 128;
  int row = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  float* xr = x + (int64_t)row * C;
  const TA* ar = ada + ada_map.off(row);
  float v[V4][4], sc[V4][4], sh[V4][4];
  // every operand of the row is requested before the first reduction: for the few-row launches of the recurrence this kernel
  // is one memory round trip (x, y, gate, scale, shift together) instead of two dependent ones
#pragma unroll
  for (int i = 0; i < V4; ++i) {
    const int c = (i * 32 + lane) * 4;
    load4(xr + c, v[i]);
    load4(ar + scale_off + c, sc[i]);
    load4(ar + shift_off + c, sh[i]);
  }
  float s = 0.f;
  if (y) {
    float yv[V4][4], gv[V4][4];
#pragma unroll
    for (int i = 0; i < V4; ++i) {
      const int c = (i * 32 + lane) * 4;
      load4(y + (int64_t)row * C + c, yv[i]);
      load4(ar + gate_off + c, gv[i]);
    }
#pragma unroll
    for (int i = 0; i < V4; ++i) {
      const int c = (i * 32 + lane) * 4;
#pragma unroll
      for (int j = 0; j < 4; ++j) v[i][j] = fmaf(yv[i][j], gv[i][j], v[i][j]);
      store4(xr + c, v[i]);
    }
  }
#pragma unroll
  for (int i = 0; i < V4; ++i) s += (v[i][0] + v[i][1]) + (v[i][2] + v[i][3]);
  float mean = warp_sum(s) * (1.0f / C);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < V4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { float d = v[i][j] - mean; q += d * d; }
  float rstd = rsqrtf(warp_sum(q) * (1.0f / C) + eps);
  TO* orow = out + (int64_t)row * C;
#pragma unroll
  for (int i = 0; i < V4; ++i) {
    int c = (i * 32 + lane) * 4;
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = (v[i][j] - mean) * rstd * (1.0f + sc[i][j]) + sh[i][j];
    store4(orow + c, o);
  }
}

int launch_adaln_modulate(float* x, const void* ada, int ada_dt, RowMap ada_map, int scale_off, int shift_off,
                          void* out, int out_dt, int rows, int cols, float eps, cudaStream_t st, const float* y, int gate_off) {
  if (rows <= 0) return AT_OK;
  AT_REQUIRE(cols == 768, "adaln: width must be 768 (got %d)", cols);
  dim3 grid(ceil_div(rows, 4));
  if (ada_dt == DT_F32 && out_dt == DT_F32)
    AT_CUDA(launch_k(adaln_kernel<6, float, float>, dim3(grid), dim3(128), 0, st, x, (const float*)ada, ada_map, scale_off, shift_off, (float*)out, rows, eps, y, gate_off));
  else if (ada_dt == DT_BF16 && out_dt == DT_BF16)
    AT_CUDA(launch_k(adaln_kernel<6, bf16, bf16>, dim3(grid), dim3(128), 0, st, x, (const bf16*)ada, ada_map, scale_off, shift_off, (bf16*)out, rows, eps, y, gate_off));
  else if (ada_dt == DT_F32 && out_dt == DT_BF16)
    AT_CUDA(launch_k(adaln_kernel<6, float, bf16>, dim3(grid), dim3(128), 0, st, x, (const float*)ada, ada_map, scale_off, shift_off, (bf16*)out, rows, eps, y, gate_off));
  else
    AT_CUDA(launch_k(adaln_kernel<6, bf16, float>, dim3(grid), dim3(128), 0, st, x, (const bf16*)ada, ada_map, scale_off, shift_off, (float*)out, rows, eps, y, gate_off));
  AT_LAUNCH_CHECK();
  return AT_OK;
}

// ---------------------------------------------------------------- audio statistics
__global__ void __launch_bounds__(1024) audio_stats_kernel(const float* __restrict__ audio, int n_samples,
                                                           float2* __restrict__ stats) {
  pdl_enter();
  __shared__ float red[32];
  const float* a = audio + (int64_t)blockIdx.x * n_samples;
  float s = 0.f;
  for (int i = threadIdx.x * 4; i < n_samples; i += blockDim.x * 4) {
    if (i + 3 < n_samples) {
      float4 t = *reinterpret_cast<const float4*>(a + i);
      s += (t.x + t.y) + (t.z + t.w);
    } else {
      for (int j = i; j < n_samples; ++j) s += a[j];
    }
  }
  float mean = block_sum(s, red) / (float)n_samples;
  float q = 0.f;
  for (int i = threadIdx.x * 4; i < n_samples; i += blockDim.x * 4) {
    if (i + 3 < n_samples) {
      float4 t = *reinterpret_cast<const float4*>(a + i);
      float d0 = t.x - mean, d1 = t.y - mean, d2 = t.z - mean, d3 = t.w - mean;
      q += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
    } else {
      for (int j = i; j < n_samples; ++j) { float d = a[j] - mean; q += d * d; }
    }
  }
  float var = block_sum(q, red) / (float)(n_samples - 1);     // unbiased (torch.std default)
  if (threadIdx.x == 0) stats[blockIdx.x] = make_float2(mean, 1.0f / (sqrtf(var) + 1e-6f));
}

int launch_audio_stats(const float* audio, int n_chunks, int n_samples, float2* stats, cudaStream_t st) {
  if (n_chunks <= 0) return AT_OK;
  AT_REQUIRE(n_samples % 4 == 0 && n_samples > 1, "audio_stats: n_samples must be a multiple of 4");
  AT_CUDA(launch_k(audio_stats_kernel, dim3(n_chunks), dim3(1024), 0, st, audio, n_samples, stats));
  AT_LAUNCH_CHECK();
  return AT_OK;
}

// ---------------------------------------------------------------- conv0 + LN + GELU
// transformers modeling_wav2vec2.py:291-299 for layer 0 (Cin = 1). One warp per output time step; lane owns
// channels {i*128 + lane*4 .. +3 : i < 4}. Weights [k][512] and the LN affine live in shared memory.
template <typename TO>
__global__ void __launch_bounds__(256) conv0_kernel(const float* __restrict__ audio, const float2* __restrict__ stats,
                                                    const float* __restrict__ w_kc, const float* __restrict__ bias,
                                                    const float* __restrict__ ln_g, const float* __restrict__ ln_b,
                                                    TO* __restrict__ out, int n_chunks, int n_samples, int l_out,
                                                    int ksz, int stride, float eps) {
  pdl_enter();
  extern __shared__ float sm[];
  float* sw = sm;                    // [ksz][512]
  float* sb = sw + ksz * 512;        // bias, gamma, beta: 3 x 512
  for (int i = threadIdx.x; i < ksz * 512; i += blockDim.x) sw[i] = w_kc[i];
  for (int i = threadIdx.x; i < 512; i += blockDim.x) {
    sb[i] = bias[i]; sb[512 + i] = ln_g[i]; sb[1024 + i] = ln_b[i];
  }
  __syncthreads();
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int64_t total = (int64_t)n_chunks * l_out;
  for (int64_t idx = (int64_t)blockIdx.x * 8 + warp; idx < total; idx += (int64_t)gridDim.x * 8) {
    int chunk = (int)(idx / l_out), t = (int)(idx - (int64_t)chunk * l_out);
    float2 stt = stats[chunk];
    const float* a = audio + (int64_t)chunk * n_samples + (int64_t)t * stride;
    float xv = (lane < ksz) ? (a[lane] - stt.x) * stt.y : 0.f;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) load4(sb + i * 128 + lane * 4, acc[i]);
    for (int k = 0; k < ksz; ++k) {
      float xk = __shfl_sync(0xffffffffu, xv, k);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float w[4];
        load4(sw + k * 512 + i * 128 + lane * 4, w);
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(w[j], xk, acc[i][j]);
      }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) s += (acc[i][0] + acc[i][1]) + (acc[i][2] + acc[i][3]);
    float mean = warp_sum(s) * (1.0f / 512.0f);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) { float d = acc[i][j] - mean; q += d * d; }
    float rstd = rsqrtf(warp_sum(q) * (1.0f / 512.0f) + eps);
    TO* orow = out + idx * 512;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int c = i * 128 + lane * 4;
      float g[4], b[4], o[4];
      load4(sb + 512 + c, g);
      load4(sb + 1024 + c, b);
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = gelu_erf((acc[i][j] - mean) * rstd * g[j] + b[j]);
      store4(orow + c, o);
    }
  }
}

// Same computation with the lane's 16 channels x 10 taps of weights held in registers (kernel size <= 10): no
// shared-memory weight traffic in the inner loop, x broadcast by shuffles. FAST = bf16 output path: erf-GELU through
// the A&S 7.1.26 rational/exp form (|err| <= 1.5e-7) instead of erff.
template <typename TO, bool FAST>
__global__ void __launch_bounds__(256, 1) conv0_reg_kernel(const float* __restrict__ audio, const float2* __restrict__ stats,
                                                           const float* __restrict__ w_kc, const float* __restrict__ bias,
                                                           const float* __restrict__ ln_g, const float* __restrict__ ln_b,
                                                           TO* __restrict__ out, int n_chunks, int n_samples, int l_out,
                                                           int ksz, int stride, float eps) {
  pdl_enter();
  constexpr int KMAX = 10;
  __shared__ __align__(16) float sb[3 * 512];
  for (int i = threadIdx.x; i < 512; i += blockDim.x) {
    sb[i] = bias[i]; sb[512 + i] = ln_g[i]; sb[1024 + i] = ln_b[i];
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float w[KMAX][16];
#pragma unroll
  for (int k = 0; k < KMAX; ++k)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float t[4] = {0.f, 0.f, 0.f, 0.f};
      if (k < ksz) load4(w_kc + k * 512 + i * 128 + lane * 4, t);
      w[k][i * 4 + 0] = t[0]; w[k][i * 4 + 1] = t[1]; w[k][i * 4 + 2] = t[2]; w[k][i * 4 + 3] = t[3];
    }
  __syncthreads();
  const int64_t total = (int64_t)n_chunks * l_out;
  const int64_t step = (int64_t)gridDim.x * 8;
  // software prefetch of the next time step's samples (only 8 warps per SM: hide the global-load latency explicitly)
  auto fetch = [&](int64_t i, float& raw, float2& st2) {
    raw = 0.f; st2 = make_float2(0.f, 0.f);
    if (i < total) {
      const int ch = (int)(i / l_out), tt = (int)(i - (int64_t)ch * l_out);
      st2 = stats[ch];
      if (lane < ksz) raw = audio[(int64_t)ch * n_samples + (int64_t)tt * stride + lane];
    }
  };
  float raw_n; float2 st_n;
  fetch((int64_t)blockIdx.x * 8 + warp, raw_n, st_n);
  for (int64_t idx = (int64_t)blockIdx.x * 8 + warp; idx < total; idx += step) {
    const float2 stt = st_n;
    const float xv = (lane < ksz) ? (raw_n - stt.x) * stt.y : 0.f;
    fetch(idx + step, raw_n, st_n);
    float acc[16];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float b4[4];
      load4(sb + i * 128 + lane * 4, b4);
      acc[i * 4] = b4[0]; acc[i * 4 + 1] = b4[1]; acc[i * 4 + 2] = b4[2]; acc[i * 4 + 3] = b4[3];
    }
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      const float xk = __shfl_sync(0xffffffffu, xv, k);
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[j] = fmaf(w[k][j], xk, acc[j]);
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) s += acc[j];
    const float mean = warp_sum(s) * (1.0f / 512.0f);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) { float d = acc[j] - mean; q += d * d; }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / 512.0f) + eps);
    TO* orow = out + idx * 512;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = i * 128 + lane * 4;
      float g[4], b[4], o[4];
      load4(sb + 512 + c, g);
      load4(sb + 1024 + c, b);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float y = (acc[i * 4 + j] - mean) * rstd * g[j] + b[j];
        o[j] = FAST ? gelu_erf_fast(y) : gelu_erf(y);
      }
      store4(orow + c, o);
    }
  }
}

int launch_conv0_ln_gelu(const float* audio, const float2* stats, const float* w_kc, const float* bias,
                         const float* ln_g, const float* ln_b, void* out, int out_dt, int n_chunks, int n_samples,
                         int l_out, int kernel, int stride, float eps, cudaStream_t st) {
  if (n_chunks <= 0) return AT_OK;
  AT_REQUIRE(kernel <= 32, "conv0: kernel size %d > 32", kernel);
  int64_t total = (int64_t)n_chunks * l_out;
  if (kernel <= 10) {
    int grid = (int)((total + 7) / 8);
    if (grid > 148) grid = 148;                      // one 8-warp block per SM (weights live in registers)
    if (out_dt == DT_F32)
      AT_CUDA(launch_k(conv0_reg_kernel<float, false>, dim3(grid), dim3(256), 0, st, audio, stats, w_kc, bias, ln_g, ln_b, (float*)out, n_chunks, n_samples,
                                                           l_out, kernel, stride, eps));
    else
      AT_CUDA(launch_k(conv0_reg_kernel<bf16, true>, dim3(grid), dim3(256), 0, st, audio, stats, w_kc, bias, ln_g, ln_b, (bf16*)out, n_chunks, n_samples,
                                                         l_out, kernel, stride, eps));
    AT_LAUNCH_CHECK();
    return AT_OK;
  }
  int grid = (int)((total + 7) / 8);
  if (grid > 148 * 8) grid = 148 * 8;
  size_t smem = (size_t)(kernel * 512 + 3 * 512) * sizeof(float);
  if (out_dt == DT_F32)
    AT_CUDA(launch_k(conv0_kernel<float>, dim3(grid), dim3(256), smem, st, audio, stats, w_kc, bias, ln_g, ln_b, (float*)out, n_chunks, n_samples,
                                                 l_out, kernel, stride, eps));
  else
    AT_CUDA(launch_k(conv0_kernel<bf16>, dim3(grid), dim3(256), smem, st, audio, stats, w_kc, bias, ln_g, ln_b, (bf16*)out, n_chunks, n_samples,
                                                l_out, kernel, stride, eps));
  AT_LAUNCH_CHECK();
  return AT_OK;
}

// ---------------------------------------------------------------- multi-scale audio pooling
struct PoolLevels { int n; int pn[8]; int cum[8]; };

__global__ void __launch_bounds__(256) audio_pool_kernel(const float* __restrict__ x, float* __restrict__ cond, int l_in,
                                                         int cols, PoolLevels lv, int l_out) {
  pdl_enter();
  int chunk = blockIdx.y, orow = blockIdx.x;
  int level = 0;
  while (level + 1 < lv.n && orow >= lv.cum[level]) ++level;
  int i = orow - (level ? lv.cum[level - 1] : 0), o = lv.pn[level];
  int start = (int)(((int64_t)i * l_in) / o);
  int end = (int)((((int64_t)(i + 1)) * l_in + o - 1) / o);
  float inv = 1.0f / (float)(end - start);
  const float* xb = x + (int64_t)chunk * l_in * cols;
  for (int c = threadIdx.x * 4; c < cols; c += blockDim.x * 4) {
    float a[4] = {0.f, 0.f, 0.f, 0.f};
    for (int r = start; r < end; ++r) {
      float v[4];
      load4(xb + (int64_t)r * cols + c, v);
#pragma unroll
      for (int j = 0; j < 4; ++j) a[j] += v[j];
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) a[j] *= inv;
    store4(cond + ((int64_t)chunk * l_out + orow) * cols + c, a);
  }
}

int launch_audio_pool(const float* x, float* cond, int n, int l_in, int cols, const int* patch_nums, int n_levels,
                      cudaStream_t st) {
  if (n <= 0) return AT_OK;
  AT_REQUIRE(n_levels >= 1 && n_levels <= 8 && cols % 4 == 0, "audio_pool: bad levels/cols");
  PoolLevels lv;
  lv.n = n_levels;
  int c = 0;
  for (int i = 0; i < n_levels; ++i) { lv.pn[i] = patch_nums[i]; c += patch_nums[i]; lv.cum[i] = c; }
  dim3 grid(c, n);
  AT_CUDA(launch_k(audio_pool_kernel, dim3(grid), dim3(256), 0, st, x, cond, l_in, cols, lv, c));
  AT_LAUNCH_CHECK();
  return AT_OK;
}

// ---------------------------------------------------------------- activation + cast
template <typename TO>
__global__ void __launch_bounds__(256) act_cast_kernel(const float* __restrict__ x, RowMap x_map, TO* __restrict__ out,
                                                       int rows, int cols, int act) {
  pdl_enter();
  int c4 = cols >> 2;
  int64_t total = (int64_t)rows * c4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int r = (int)(i / c4), c = (int)(i - (int64_t)r * c4) * 4;
    float v[4];
    load4(x + x_map.off(r) + c, v);
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = apply_act(v[j], act);
    store4(out + (int64_t)r * cols + c, v);
  }
}

int launch_act_cast(const float* x, RowMap x_map, void* out, int out_dt, int rows, int cols, int act, cudaStream_t st) {
  if (rows <= 0) return AT_OK;
  AT_REQUIRE(cols % 4 == 0, "act_cast: cols %% 4");
  int64_t total = (int64_t)rows * (cols / 4);
  int grid = (int)((total + 255) / 256);
  if (grid > 148 * 16) grid = 148 * 16;
  if (out_dt == DT_F32) AT_CUDA(launch_k(act_cast_kernel<float>, dim3(grid), dim3(256), 0, st, x, x_map, (float*)out, rows, cols, act));
  else AT_CUDA(launch_k(act_cast_kernel<bf16>, dim3(grid), dim3(256), 0, st, x, x_map, (bf16*)out, rows, cols, act));
  AT_LAUNCH_CHECK();
  return AT_OK;
}

}  // namespace artalk
