// fp32 CUDA-core GEMM with the path's fused epilogues:  out = resid + gate * act(A W^T + bias).
// This is the arithmetic of the fp32 precision mode (north-star tolerance 1e-3) and of the tiny once-per-clip
// style encoder; the bf16 mode runs the same GemmArgs through the tcgen05 kernel in gemm_tc.cu.
// 128x128x16 tiles, 256 threads, 8x8 register tile per thread, register-prefetched double buffering; a 32x128 variant
// (2x8 per thread) keeps the small GEMMs of the style encoder (M = 50 rows per clip) from running on a handful of SMs.
#include "kernels.cuh"

namespace artalk {

namespace {
constexpr int BN = 128, BK = 16, PADM = 4;

struct SimtParams {
  const float* A; RowMap a_map; const float* W; int64_t ldw; int M, N, K;
  int tap_w, tap_pad; int64_t a_gs, w_gs, c_gs; int bias_gs;
  const float* bias; int act;
  const void* gate; int gate_dt; RowMap gate_map;
  const float* resid; RowMap resid_map;
  float* out32; void* out_act; int out_act_dt; RowMap c_map;
  int vec_ok;
};

__device__ __forceinline__ float gate_at(const void* gate, int dt, int64_t off) {
  return dt == DT_F32 ? reinterpret_cast<const float*>(gate)[off] : __bfloat162float(reinterpret_cast<const bf16*>(gate)[off]);
}

template <int TM>          // rows per thread: 8 -> 128-row tiles, 2 -> 32-row tiles
__global__ void __launch_bounds__(256) gemm_simt_kernel(SimtParams p) {
  pdl_enter();
  constexpr int BM = 16 * TM, AR = (BM + 63) / 64;      // AR: A rows fetched per loader thread
  __shared__ __align__(16) float As[2][BK][BM + PADM];
  __shared__ __align__(16) float Bs[2][BK][BN + PADM];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int g = blockIdx.z;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const float* A = p.A + g * p.a_gs;
  const float* W = p.W + g * p.w_gs;

  // loader assignment: 2 rows of A and 2 rows of W per thread, one float4 along K each
  const int lr = tid >> 2, lk = (tid & 3) * 4;
  int64_t a_base[2]; int a_t[2]; bool a_ok[2];
  int64_t w_base[2]; bool w_ok[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    int r = m0 + lr + 64 * i;
    a_ok[i] = i < AR && lr + 64 * i < BM && r < p.M;
    int rr = a_ok[i] ? r : 0;
    if (p.tap_w > 0) {
      int b = rr / p.a_map.rpb;
      a_t[i] = rr - b * p.a_map.rpb;
      a_base[i] = (int64_t)b * p.a_map.bs;
    } else {
      a_t[i] = 0;
      a_base[i] = p.a_map.off(rr);
    }
    int n = n0 + lr + 64 * i;
    w_ok[i] = n < p.N;
    w_base[i] = (int64_t)(w_ok[i] ? n : 0) * p.ldw;
  }

  float4 ra[2], rb[2];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (a_ok[i]) {
        if (p.tap_w > 0) {
          int j = k0 / p.tap_w, c0 = k0 - j * p.tap_w;
          int tt = a_t[i] + j - p.tap_pad;
          if (tt >= 0 && tt < p.a_map.rpb)
            v = *reinterpret_cast<const float4*>(A + a_base[i] + (int64_t)tt * p.a_map.rs + c0 + lk);
        } else {
          v = *reinterpret_cast<const float4*>(A + a_base[i] + k0 + lk);
        }
      }
      ra[i] = v;
      float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
      if (w_ok[i]) w = *reinterpret_cast<const float4*>(W + w_base[i] + k0 + lk);
      rb[i] = w;
    }
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      int r = lr + 64 * i;
      if (r < BM) { As[buf][lk + 0][r] = ra[i].x; As[buf][lk + 1][r] = ra[i].y; As[buf][lk + 2][r] = ra[i].z; As[buf][lk + 3][r] = ra[i].w; }
      Bs[buf][lk + 0][r] = rb[i].x; Bs[buf][lk + 1][r] = rb[i].y; Bs[buf][lk + 2][r] = rb[i].z; Bs[buf][lk + 3][r] = rb[i].w;
    }
  };

  float acc[TM][8];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  const int nk = p.K / BK;
  fetch(0);
  stash(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    int buf = kt & 1;
    if (kt + 1 < nk) fetch((kt + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM];
      if constexpr (TM == 8) {
        float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
        float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
        a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      } else {
        float2 a0 = *reinterpret_cast<const float2*>(&As[buf][k][ty * 2]);
        a[0] = a0.x; a[1] = a0.y;
      }
      float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
      float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      stash(buf ^ 1);
      __syncthreads();
    }
  }

  // ---- epilogue: out = resid + gate * act(acc + bias)
  const float* bias = p.bias ? p.bias + g * p.bias_gs : nullptr;
  const int64_t cg = g * p.c_gs;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int r = TM == 8 ? m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4)) : m0 + ty * 2 + i;
    if (r >= p.M) continue;
    int64_t c_off = p.c_map.off(r) + cg;
    int64_t g_off = p.gate ? p.gate_map.off(r) + cg : 0;
    int64_t r_off = p.resid ? p.resid_map.off(r) + cg : 0;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      int c = n0 + h * 64 + tx * 4;
      if (c >= p.N) continue;
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float t = acc[i][h * 4 + j];
        if (bias && c + j < p.N) t += bias[c + j];
        v[j] = apply_act(t, p.act);
      }
      if (p.vec_ok && c + 3 < p.N) {
        if (p.gate) {
          float gv[4];
          if (p.gate_dt == DT_F32) load4(reinterpret_cast<const float*>(p.gate) + g_off + c, gv);
          else load4(reinterpret_cast<const bf16*>(p.gate) + g_off + c, gv);
#pragma unroll
          for (int j = 0; j < 4; ++j) v[j] *= gv[j];
        }
        if (p.resid) {
          float rv[4];
          load4(p.resid + r_off + c, rv);
#pragma unroll
          for (int j = 0; j < 4; ++j) v[j] += rv[j];
        }
        if (p.out32) store4(p.out32 + c_off + c, v);
        if (p.out_act) {
          if (p.out_act_dt == DT_F32) store4(reinterpret_cast<float*>(p.out_act) + c_off + c, v);
          else store4(reinterpret_cast<bf16*>(p.out_act) + c_off + c, v);
        }
      } else {
        for (int j = 0; j < 4 && c + j < p.N; ++j) {
          float t = v[j];
          if (p.gate) t *= gate_at(p.gate, p.gate_dt, g_off + c + j);
          if (p.resid) t += p.resid[r_off + c + j];
          if (p.out32) p.out32[c_off + c + j] = t;
          if (p.out_act) {
            if (p.out_act_dt == DT_F32) reinterpret_cast<float*>(p.out_act)[c_off + c + j] = t;
            else reinterpret_cast<bf16*>(p.out_act)[c_off + c + j] = __float2bfloat16_rn(t);
          }
        }
      }
    }
  }
}
}  // namespace

int launch_gemm_simt(const GemmArgs& g, cudaStream_t st) {
  if (g.M <= 0 || g.N <= 0) return AT_OK;
  AT_REQUIRE(g.K > 0 && g.K % BK == 0, "gemm_simt: K=%d must be a positive multiple of %d", g.K, BK);
  AT_REQUIRE(g.A && g.W && (g.out32 || g.out_act), "gemm_simt: null operand");
  AT_REQUIRE(g.ldw % 4 == 0 && g.a_map.rs % 4 == 0 && g.a_map.bs % 4 == 0 && g.a_gs % 4 == 0 && g.w_gs % 4 == 0,
             "gemm_simt: operand strides must be multiples of 4 elements");
  AT_REQUIRE(g.tap_w == 0 || (g.tap_w % BK == 0 && g.a_map.rpb > 0 && g.K % g.tap_w == 0), "gemm_simt: bad tap mode");
  SimtParams p;
  p.A = (const float*)g.A; p.a_map = g.a_map; p.W = (const float*)g.W; p.ldw = g.ldw; p.M = g.M; p.N = g.N; p.K = g.K;
  p.tap_w = g.tap_w; p.tap_pad = g.tap_pad; p.a_gs = g.a_gs; p.w_gs = g.w_gs; p.c_gs = g.c_gs; p.bias_gs = g.bias_gs;
  p.bias = g.bias; p.act = g.act; p.gate = g.gate; p.gate_dt = g.gate_dt; p.gate_map = g.gate_map;
  p.resid = g.resid; p.resid_map = g.resid_map; p.out32 = g.out32; p.out_act = g.out_act; p.out_act_dt = g.out_act_dt;
  p.c_map = g.c_map;
  bool v = (g.N % 4 == 0) && (g.c_map.rs % 4 == 0) && (g.c_map.bs % 4 == 0) && (g.c_gs % 4 == 0);
  if (g.gate) v = v && (g.gate_map.rs % 4 == 0) && (g.gate_map.bs % 4 == 0) && (((uintptr_t)g.gate) % 16 == 0);
  if (g.resid) v = v && (g.resid_map.rs % 4 == 0) && (g.resid_map.bs % 4 == 0) && (((uintptr_t)g.resid) % 16 == 0);
  if (g.out32) v = v && (((uintptr_t)g.out32) % 16 == 0);
  if (g.out_act) v = v && (((uintptr_t)g.out_act) % 16 == 0);
  p.vec_ok = v ? 1 : 0;
  g_trace_dims[0] = g.M; g_trace_dims[1] = g.N; g_trace_dims[2] = g.K * g.groups;
  // less than one wave of 128-row tiles (148 SMs): 32-row tiles put 4x as many CTAs on the machine
  if ((long)ceil_div(g.N, BN) * ceil_div(g.M, 128) * g.groups < 148) {
    dim3 grid(ceil_div(g.N, BN), ceil_div(g.M, 32), g.groups);
    AT_CUDA(launch_k(gemm_simt_kernel<2>, dim3(grid), dim3(256), 0, st, p));
  } else {
    dim3 grid(ceil_div(g.N, BN), ceil_div(g.M, 128), g.groups);
    AT_CUDA(launch_k(gemm_simt_kernel<8>, dim3(grid), dim3(256), 0, st, p));
  }
  AT_LAUNCH_CHECK();
  return AT_OK;
}

}  // namespace artalk
