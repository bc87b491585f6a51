// wav2vec2 feature-extractor layer 0 of the bf16 path: Conv1d(1 -> 512, k = 10, stride 5) + LayerNorm(512) + GELU
// (transformers modeling_wav2vec2.py:291-299) with the LayerNorm folded through the convolution.
//
// The conv is linear in the 10 samples x of a frame, so the LayerNorm statistics over the 512 channels are functions of x alone:
//   y_c - mean_c(y) = wc_c . x + bc_c            (wc = w - mean_c(w), bc = b - mean_c(b): channel-centred filter / bias)
//   var_c(y)        = [x; 1]^T Q [x; 1],  Q = (1/512) sum_c [wc_c; bc_c][wc_c; bc_c]^T   (11 x 11, PSD)
//                   = sum_i (qf_i . [x; 1])^2    (qf_i = sqrt(lambda_i) v_i from the eigen-decomposition of Q, fp64 at load)
//   out_c           = gelu(rstd . (wq_c . x + bq_c) + beta_c),   wq = wc * gamma, bq = bc * gamma
// i.e. 132 FMAs per FRAME for the statistics instead of a mean / centred-square / normalise pass per ELEMENT, and no cross-lane
// reduction, so a thread can own 4 channels for many frames with its filters in registers. The per-element work (10 FMA conv,
// 1 FMA normalise, erf-GELU) runs on packed pairs (fma.rn.f32x2 -> FFMA2: two channels per instruction): 12.5 issued
// instructions per element. Measured (ncu, profiles/r2_ncu_new_kernels.md): 1.6x faster than the direct kernel, issue 48 %,
// FMA pipe 27 %; FFMA2 halves the issue slots of the arithmetic but occupies the FMA pipe like two FFMAs, and a third block
// per SM (80 registers) measured equal.
// The previous kernel (norms.cu conv0_reg_kernel: one warp per frame, ~31 issued instructions per element at 57 % issue
// utilisation, profiles/r2_ncu_kernel_table.md) stays for the fp32 / parity-grade modes and as the A/B reference (option "conv0_fold").
// The sum of squares is evaluated with the same conditioning as the direct form (every term is a dot product with x).
#include "kernels.cuh"

namespace artalk {

namespace {

typedef unsigned long long u64;
__device__ __forceinline__ u64 pk2(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk2(u64 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

// gelu_erf_fast (common.cuh) on a pair
__device__ __forceinline__ u64 gelu2(u64 x) {
  u64 x2 = mul2(x, x);
  float a, b;
  upk2(x2, a, b);
  x2 = pk2(fminf(a, 49.0f), fminf(b, 49.0f));
  u64 p = fma2(pk2(-3.8652387e-4f, -3.8652387e-4f), x2, pk2(3.7255297e-2f, 3.7255297e-2f));
  p = fma2(p, x2, pk2(7.9716741e-1f, 7.9716741e-1f));
  p = mul2(p, x);
  upk2(p, a, b);
  float ta, tb;
  asm("tanh.approx.f32 %0, %1;" : "=f"(ta) : "f"(a));
  asm("tanh.approx.f32 %0, %1;" : "=f"(tb) : "f"(b));
  const u64 hx = mul2(x, pk2(0.5f, 0.5f));
  return fma2(hx, pk2(ta, tb), hx);
}

constexpr int TT = 64;                 // frames per tile
constexpr int NS = 5 * TT + 5;         // samples a tile touches (stride 5, kernel 10)
constexpr int QF_LD = 12;              // row pitch of the 11 x 11 variance factor

__global__ void __launch_bounds__(256) conv0_fold_kernel(const float* __restrict__ audio, const float2* __restrict__ stats,
                                                         const float* __restrict__ wq, const float* __restrict__ bq,
                                                         const float* __restrict__ beta, const float* __restrict__ qf,
                                                         bf16* __restrict__ out, int n_chunks, int n_samples, int l_out, float eps) {
  pdl_enter();
  __shared__ __align__(16) float2 xs2[NS + 3];        // normalised samples, duplicated (x, x): operands of the packed FMAs
  __shared__ __align__(16) float2 rs2[TT];            // 1 / sqrt(var + eps) per frame, duplicated
  __shared__ float sq[11 * QF_LD];
  const int tid = threadIdx.x, cg = tid & 127, th = tid >> 7;
  // this thread's 4 channels: filters, centred bias and LayerNorm shift as pairs (c, c+1), (c+2, c+3)
  u64 w[10][2];
#pragma unroll
  for (int k = 0; k < 10; ++k) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(wq + k * 512) + cg);
    w[k][0] = pk2(t.x, t.y); w[k][1] = pk2(t.z, t.w);
  }
  const float4 b4 = __ldg(reinterpret_cast<const float4*>(bq) + cg), e4 = __ldg(reinterpret_cast<const float4*>(beta) + cg);
  const u64 bq0 = pk2(b4.x, b4.y), bq1 = pk2(b4.z, b4.w), be0 = pk2(e4.x, e4.y), be1 = pk2(e4.z, e4.w);
  for (int i = tid; i < 11 * QF_LD; i += 256) sq[i] = qf[i];
  const int tpc = (l_out + TT - 1) / TT;
  const int total_tiles = n_chunks * tpc;
  for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    const int chunk = tile / tpc, t0 = (tile - chunk * tpc) * TT;
    const int nt = min(TT, l_out - t0);
    __syncthreads();                                   // the previous tile's readers are done (first tile: sq is staged)
    {
      const float2 st = stats[chunk];
      const float* a = audio + (int64_t)chunk * n_samples + 5 * t0;
      const int ns = 5 * nt + 5;                       // last sample index 5 (t0 + nt - 1) + 9 < n_samples
      for (int i = tid; i < NS + 3; i += 256) {
        const float v = (i < ns) ? (a[i] - st.x) * st.y : 0.f;
        xs2[i] = make_float2(v, v);
      }
    }
    __syncthreads();
    {
      // variance of the frame's 512 conv outputs from its 10 samples: four threads per frame share the 11 factor rows
      const int t = tid >> 2, part = tid & 3;
      float acc = 0.f;
      if (t < nt) {
        float z[10];
#pragma unroll
        for (int k = 0; k < 10; ++k) z[k] = xs2[5 * t + k].x;
        for (int i = part; i < 11; i += 4) {
          float u = sq[i * QF_LD + 10];
#pragma unroll
          for (int k = 0; k < 10; ++k) u = fmaf(sq[i * QF_LD + k], z[k], u);
          acc = fmaf(u, u, acc);
        }
      }
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      if (part == 0) { const float r = rsqrtf(acc + eps); rs2[t] = make_float2(r, r); }
    }
    __syncthreads();
    // 4 channels x 4 frames per step; thread half `th` takes frames [32 th, 32 th + 32) of the tile
    bf16* orow = out + ((int64_t)chunk * l_out + t0) * 512 + cg * 4;
#pragma unroll 1
    for (int tg = 0; tg < 8; ++tg) {
      const int tb = th * 32 + tg * 4;
      if (tb >= nt) break;
      u64 xd[26];
      const float4* xp = reinterpret_cast<const float4*>(xs2 + 5 * tb);        // 40 tb bytes: 16-byte aligned (tb % 4 == 0)
#pragma unroll
      for (int j = 0; j < 13; ++j) { const float4 v = xp[j]; xd[2 * j] = pk2(v.x, v.y); xd[2 * j + 1] = pk2(v.z, v.w); }
      u64 acc[4][2];
#pragma unroll
      for (int tt = 0; tt < 4; ++tt) { acc[tt][0] = bq0; acc[tt][1] = bq1; }
#pragma unroll
      for (int k = 0; k < 10; ++k)
#pragma unroll
        for (int tt = 0; tt < 4; ++tt) {
          acc[tt][0] = fma2(w[k][0], xd[5 * tt + k], acc[tt][0]);
          acc[tt][1] = fma2(w[k][1], xd[5 * tt + k], acc[tt][1]);
        }
#pragma unroll
      for (int tt = 0; tt < 4; ++tt) {
        if (tb + tt < nt) {
          const float2 r = rs2[tb + tt];
          const u64 r2 = pk2(r.x, r.y);
          const u64 g0 = gelu2(fma2(acc[tt][0], r2, be0)), g1 = gelu2(fma2(acc[tt][1], r2, be1));
          float o0, o1, o2, o3;
          upk2(g0, o0, o1); upk2(g1, o2, o3);
          __nv_bfloat162 h0 = __floats2bfloat162_rn(o0, o1), h1 = __floats2bfloat162_rn(o2, o3);
          uint2 pk;
          pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
          *reinterpret_cast<uint2*>(orow + (int64_t)(tb + tt) * 512) = pk;
        }
      }
    }
  }
}

}  // namespace

// wq [10][512] = (w - mean_c w) * gamma, bq [512] = (b - mean_c b) * gamma, beta [512], qf [11][12] variance factor rows
// (weights.conv0_fold). out bf16 [n_chunks][l_out][512].
int launch_conv0_fold(const float* audio, const float2* stats, const float* wq, const float* bq, const float* beta, const float* qf,
                      void* out, int n_chunks, int n_samples, int l_out, float eps, cudaStream_t st) {
  if (n_chunks <= 0) return AT_OK;
  AT_REQUIRE(l_out == (n_samples - 10) / 5 + 1 && l_out > 0, "conv0_fold: kernel 10 / stride 5 only (l_out=%d, samples=%d)", l_out, n_samples);
  AT_REQUIRE(((uintptr_t)wq % 16 == 0) && ((uintptr_t)bq % 16 == 0) && ((uintptr_t)beta % 16 == 0) && ((uintptr_t)out % 8 == 0),
             "conv0_fold: operands must be 16-byte aligned");
  const DevCtx* dc = nullptr;
  AT_TRY(dev_ctx(&dc));
  static int per_sm_dev[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  int& per_sm = per_device_slot(per_sm_dev);
  if (per_sm <= 0) {
    int n = 0;
    AT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, conv0_fold_kernel, 256, 0));
    per_sm = n > 0 ? n : 1;
  }
  const int64_t tiles = (int64_t)n_chunks * ((l_out + TT - 1) / TT);
  const int64_t cap = (int64_t)dc->num_sms * per_sm;
  const int grid = (int)(tiles < cap ? tiles : cap);
  AT_CUDA(launch_k(conv0_fold_kernel, dim3(grid), dim3(256), 0, st, audio, stats, wq, bq, beta, qf, (bf16*)out, n_chunks, n_samples, l_out, eps));
  AT_LAUNCH_CHECK();
  return AT_OK;
}

}  // namespace artalk
