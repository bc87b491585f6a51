// Latency path of the recurrence: skinny-M bf16 GEMM.
//
// The first scale steps of every chunk (1 and 5 new tokens per clip: M = 64 / 320 rows at 64 clips) and every step of the
// batch-1 streaming mode are pure latency: 3 % of the AR flops took 22 % of the AR time on the tcgen05 kernels, whose fixed
// cost per launch (TMEM allocation, tensor-map fetch, 128-row tiles, barrier ring, persistent scheduler) is ~10 us and whose
// tile shape leaves most SMs idle (M = 64, N = 768: 6-24 CTAs walk K = 3072 serially). Here:
//
//  * skinny_gemm_kernel<TN, NST>: out = resid + gate * act(A W^T + b) for M <= ~512 rows. One CTA = 64 rows x TN columns
//    (TN = 8..64 chosen so that ~100-300 CTAs stream disjoint slices of W at once), K walked in 128-wide chunks through an
//    NST-deep cp.async ring (up to ~200 KB in flight per SM: the whole K = 768 panel of a tile is requested at once),
//    mma.sync.m16n8k16 (bf16, fp32 accumulate) from ldmatrix fragments, 8 warps = 4 row groups x 2 K halves reduced through
//    shared memory, and one generic epilogue pass over the fp32 tile (bias, activation, gate, residual, fp32 / bf16 outputs
//    through row maps, or the fused AR q/k/v head normalisation + KV-cache scatter of app/transformer.py:68-74).
//    Weights do not depend on the previous kernel: their first ring fill is issued before griddepcontrol.wait (PDL).
//
// The tensor pipe is irrelevant at these sizes (<= 1.5 GFLOP per launch); what matters is bytes in flight and CTA count.
#define ARTALK_PDL_CLASS 1
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "kernels.cuh"

namespace artalk {

namespace {

constexpr int SK_BM = 64, SK_KC = 128, SK_THREADS = 256;
constexpr int SK_PITCH = SK_KC * 2 + 16;            // bytes per shared-memory row: 17 x 16 B, conflict-free for ldmatrix
constexpr int SK_A_STAGE = SK_BM * SK_PITCH;

struct SkParams {
  const bf16* A; RowMap a_map;
  const bf16* W; int64_t ldw;
  int M, N, n_chunks;
  const float* bias; int act;
  const void* gate; int gate_dt; RowMap gate_map;
  const float* resid; RowMap resid_map;
  float* out32; void* out_act; int out_act_dt; RowMap c_map;
  int qkv_mode, qkv_C; const float* head_scale; bf16* qbuf; bf16* kcache; bf16* vcache; RowMap kv_map; int64_t kv_layer_stride;
};

__device__ __forceinline__ uint32_t sm_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2(uint32_t (&r)[2], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0, %1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int TN> struct SkCfg {
  static constexpr int STAGE = SK_A_STAGE + TN * SK_PITCH;
  // deep ring: one CTA per SM with ~200 KB in flight; shallow: two CTAs per SM
  static constexpr int NST_DEEP = TN == 64 ? 6 : TN == 32 ? 8 : 10;
  static constexpr int NST_SHALLOW = TN == 64 ? 3 : TN == 32 ? 4 : 5;
};

// One 64-row x TN-column output tile: ring-fed main loop, K-half reduction and epilogue. PDL: the tile is a whole kernel
// launched with programmatic serialization (weights are requested before the grid dependency resolves). All 256 threads
// of the CTA call it; the shared memory may be reused as soon as it returns and the CTA has synchronised.
// W_PRE: the caller has already issued this tile's first ring fill of W (sk_prefetch_w), e.g. before a grid barrier.
template <int TN, int NST>
__device__ __forceinline__ void sk_prefetch_w(const bf16* W, int64_t ldw, int nch, const int col0, uint8_t* sk_smem) {
  constexpr int STAGE = SkCfg<TN>::STAGE;
  const int tid = threadIdx.x, pc = tid & 15, r_lo = tid >> 4;
  const uint32_t sbase = sm_u32(sk_smem);
  const int pre = nch < NST ? nch : NST;
  for (int c = 0; c < pre; ++c) {
    const uint32_t dst = sbase + c * STAGE + SK_A_STAGE + pc * 16;
#pragma unroll
    for (int j = 0; j < (TN + 15) / 16; ++j) {
      const int r = r_lo + 16 * j;
      if (TN % 16 == 0 || r < TN) cp_async16(dst + r * SK_PITCH, W + (int64_t)(col0 + r) * ldw + c * SK_KC + pc * 8);
    }
  }
}

template <int TN, int NST, bool PDL, bool W_PRE = false>
__device__ __forceinline__ void sk_tile(const SkParams& p, const int col0, const int row0, uint8_t* sk_smem) {
  constexpr int STAGE = SkCfg<TN>::STAGE;
  static_assert(NST >= 3, "ring depth");
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rg = warp & 3, kh = warp >> 2;
  const uint32_t sbase = sm_u32(sk_smem);
  const int nch = p.n_chunks;

  // each thread copies the same 16-byte column piece of 4 A rows (and of up to 4 W rows) of every chunk
  const int pc = tid & 15, r_lo = tid >> 4;
  const bf16* a_src[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    // rows past M are not loaded (their accumulators see whatever the ring holds and are never stored). Clamping them to
    // a valid row instead made every CTA hammer the same few cache lines: 63 us for M = 1, K = 3072
    const int gr = row0 + r_lo + 16 * j;
    a_src[j] = gr < p.M ? p.A + p.a_map.off(gr) + pc * 8 : nullptr;
  }
  auto load_w = [&](int chunk, int slot) {
    const uint32_t dst = sbase + slot * STAGE + SK_A_STAGE + pc * 16;
#pragma unroll
    for (int j = 0; j < (TN + 15) / 16; ++j) {
      const int r = r_lo + 16 * j;
      if (TN % 16 == 0 || r < TN) cp_async16(dst + r * SK_PITCH, p.W + (int64_t)(col0 + r) * p.ldw + chunk * SK_KC + pc * 8);
    }
  };
  auto load_a = [&](int chunk, int slot) {
    const uint32_t dst = sbase + slot * STAGE + pc * 16;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (a_src[j]) cp_async16(dst + (r_lo + 16 * j) * SK_PITCH, a_src[j] + chunk * SK_KC);
  };

  if (PDL) pdl_launch_dependents();
  const int pre = nch < NST ? nch : NST;
  if (!W_PRE)
    for (int c = 0; c < pre; ++c) load_w(c, c);         // weights: independent of the previous kernel
  if (PDL) pdl_wait();
#pragma unroll 1
  for (int c = 0; c < NST; ++c) {                       // always NST groups so that chunk c <-> group c
    if (c < pre) load_a(c, c);
    cp_async_commit();
  }

  float acc[TN / 8][4];
#pragma unroll
  for (int i = 0; i < TN / 8; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;

  int slot = 0;
#pragma unroll 1
  for (int c = 0; c < nch; ++c) {
    cp_async_wait<NST - 2>();                            // groups <= c complete (conservative by one at c = 0)
    __syncthreads();                                     // chunk c visible to all; everyone is done with chunk c - 1
    if (c >= 1) {                                        // refill the slot chunk c - 1 used
      const int nc = c - 1 + NST;
      if (nc < nch) { const int ps = slot == 0 ? NST - 1 : slot - 1; load_w(nc, ps); load_a(nc, ps); }
      cp_async_commit();
    }
    const uint32_t sa = sbase + slot * STAGE, sw = sa + SK_A_STAGE;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const int k0 = kh * 64 + ks * 16;
      uint32_t a[4];
      ldsm_x4(a, sa + (rg * 16 + (lane & 15)) * SK_PITCH + (k0 + (lane >> 4) * 8) * 2);
      if constexpr (TN == 8) {
        uint32_t b[2];
        ldsm_x2(b, sw + (lane & 7) * SK_PITCH + (k0 + ((lane >> 3) & 1) * 8) * 2);
        mma_bf16(acc[0], a, b[0], b[1]);
      } else {
#pragma unroll
        for (int nt = 0; nt < TN / 8; nt += 2) {
          uint32_t b[4];
          ldsm_x4(b, sw + (nt * 8 + (lane & 7) + ((lane >> 4) & 1) * 8) * SK_PITCH + (k0 + ((lane >> 3) & 1) * 8) * 2);
          mma_bf16(acc[nt], a, b[0], b[1]);
          mma_bf16(acc[nt + 1], a, b[2], b[3]);
        }
      }
    }
    slot = slot + 1 == NST ? 0 : slot + 1;
  }
  cp_async_wait<0>();
  __syncthreads();

  // ---- the two K halves meet in a fp32 tile in shared memory
  constexpr int CP = TN + 4;                              // row pitch (floats)
  float* ct = reinterpret_cast<float*>(sk_smem);
  {
    const int r = rg * 16 + (lane >> 2), cc = (lane & 3) * 2;
    if (kh == 0) {
#pragma unroll
      for (int nt = 0; nt < TN / 8; ++nt) {
        *reinterpret_cast<float2*>(&ct[r * CP + nt * 8 + cc]) = make_float2(acc[nt][0], acc[nt][1]);
        *reinterpret_cast<float2*>(&ct[(r + 8) * CP + nt * 8 + cc]) = make_float2(acc[nt][2], acc[nt][3]);
      }
    }
    __syncthreads();
    if (kh == 1) {
#pragma unroll
      for (int nt = 0; nt < TN / 8; ++nt) {
        float2* d0 = reinterpret_cast<float2*>(&ct[r * CP + nt * 8 + cc]);
        float2* d1 = reinterpret_cast<float2*>(&ct[(r + 8) * CP + nt * 8 + cc]);
        float2 t0 = *d0, t1 = *d1;
        *d0 = make_float2(t0.x + acc[nt][0], t0.y + acc[nt][1]);
        *d1 = make_float2(t1.x + acc[nt][2], t1.y + acc[nt][3]);
      }
    }
    __syncthreads();
  }

  // ---- epilogue: 4 threads per row, TN / 4 consecutive columns each
  constexpr int CPT = TN / 4;
  const int row = tid >> 2, cg = tid & 3;
  const int gr = row0 + row;
  const bool ok = gr < p.M;
  const int cb = col0 + cg * CPT;                         // first global column of this thread
  float v[CPT];
#pragma unroll
  for (int j = 0; j < CPT; ++j) v[j] = ct[row * CP + cg * CPT + j];
  if (p.bias) {
#pragma unroll
    for (int j = 0; j < CPT; ++j) v[j] += __ldg(p.bias + cb + j);
  }
  if constexpr (TN == 64) {
    if (p.qkv_mode) {
      // fused AR q/k/v epilogue: the tile is one 64-wide head of q, k or v
      const int period = (p.qkv_mode == 1 ? 3 : 2) * p.qkv_C;
      const int layer = col0 / period, o = col0 - layer * period;
      const int sec = o / p.qkv_C, hc = o - sec * p.qkv_C;
      const bool is_q = (p.qkv_mode == 1 && sec == 0);
      const bool is_k = (p.qkv_mode == 1) ? (sec == 1) : (sec == 0);
      float scale = 1.0f;
      if (is_q || is_k) {
        float ss = 0.f;
#pragma unroll
        for (int j = 0; j < CPT; ++j) ss = fmaf(v[j], v[j], ss);
        ss += __shfl_xor_sync(0xffffffffu, ss, 1);
        ss += __shfl_xor_sync(0xffffffffu, ss, 2);
        scale = 1.0f / fmaxf(sqrtf(ss), 1e-12f);            // F.normalize eps
        if (is_q) scale *= p.head_scale[hc >> 6];
      }
      if (ok) {
        bf16* dst = is_q ? p.qbuf + (int64_t)gr * p.qkv_C + hc
                         : (is_k ? p.kcache : p.vcache) + (int64_t)layer * p.kv_layer_stride + p.kv_map.off(gr) + hc;
        dst += cg * CPT;
#pragma unroll
        for (int j = 0; j < CPT; j += 8) {
          __nv_bfloat162 h0 = __floats2bfloat162_rn(v[j] * scale, v[j + 1] * scale), h1 = __floats2bfloat162_rn(v[j + 2] * scale, v[j + 3] * scale);
          __nv_bfloat162 h2 = __floats2bfloat162_rn(v[j + 4] * scale, v[j + 5] * scale), h3 = __floats2bfloat162_rn(v[j + 6] * scale, v[j + 7] * scale);
          uint4 pk;
          pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
          pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
          *reinterpret_cast<uint4*>(dst + j) = pk;
        }
      }
      return;
    }
  }
  switch (p.act) {
    case ACT_GELU_ERF:
#pragma unroll
      for (int j = 0; j < CPT; ++j) v[j] = gelu_erf_fast(v[j]);
      break;
    case ACT_GELU_TANH:
#pragma unroll
      for (int j = 0; j < CPT; ++j) v[j] = gelu_tanh_fast(v[j]);
      break;
    case ACT_LEAKY02:
#pragma unroll
      for (int j = 0; j < CPT; ++j) v[j] = v[j] > 0.f ? v[j] : 0.2f * v[j];
      break;
    case ACT_SILU:
#pragma unroll
      for (int j = 0; j < CPT; ++j) v[j] = v[j] / (1.0f + __expf(-v[j]));
      break;
    default: break;
  }
  if (!ok) return;
  if (p.gate) {
    const int64_t g_off = p.gate_map.off(gr) + cb;
#pragma unroll
    for (int j = 0; j < CPT; j += 2) {
      if (p.gate_dt == DT_F32) {
        const float2 g2 = *reinterpret_cast<const float2*>(reinterpret_cast<const float*>(p.gate) + g_off + j);
        v[j] *= g2.x; v[j + 1] *= g2.y;
      } else {
        const __nv_bfloat162 g2 = *reinterpret_cast<const __nv_bfloat162*>(reinterpret_cast<const bf16*>(p.gate) + g_off + j);
        v[j] *= __low2float(g2); v[j + 1] *= __high2float(g2);
      }
    }
  }
  if (p.resid) {
    const float* rp = p.resid + p.resid_map.off(gr) + cb;
#pragma unroll
    for (int j = 0; j < CPT; j += 2) {
      const float2 r2 = *reinterpret_cast<const float2*>(rp + j);
      v[j] += r2.x; v[j + 1] += r2.y;
    }
  }
  const int64_t c_off = p.c_map.off(gr) + cb;
  if (p.out32) {
#pragma unroll
    for (int j = 0; j < CPT; j += 2) *reinterpret_cast<float2*>(p.out32 + c_off + j) = make_float2(v[j], v[j + 1]);
  }
  if (p.out_act) {
    if (p.out_act_dt == DT_F32) {
#pragma unroll
      for (int j = 0; j < CPT; j += 2) *reinterpret_cast<float2*>(reinterpret_cast<float*>(p.out_act) + c_off + j) = make_float2(v[j], v[j + 1]);
    } else {
#pragma unroll
      for (int j = 0; j < CPT; j += 2)
        *reinterpret_cast<__nv_bfloat162*>(reinterpret_cast<bf16*>(p.out_act) + c_off + j) = __floats2bfloat162_rn(v[j], v[j + 1]);
    }
  }
}

template <int TN, int NST>
__global__ void __launch_bounds__(SK_THREADS) skinny_gemm_kernel(const SkParams p) {
  extern __shared__ __align__(128) uint8_t sk_smem_dyn[];
  sk_tile<TN, NST, true>(p, (int)blockIdx.x * TN, (int)blockIdx.y * SK_BM, sk_smem_dyn);
}

int g_skinny_max_m = 512;       // option "skinny_max_m": row cap of the skinny kernel (0 = off); GemmArgs::skinny opts a call in

template <int TN, int NST>
int launch_sk(const SkParams& p, dim3 grid, cudaStream_t st) {
  constexpr int smem = NST * SkCfg<TN>::STAGE;
  AT_TRY(ensure_dyn_smem((const void*)skinny_gemm_kernel<TN, NST>, smem));
  AT_CUDA(launch_k(skinny_gemm_kernel<TN, NST>, grid, dim3(SK_THREADS), smem, st, p));
  AT_LAUNCH_CHECK();
  return AT_OK;
}
template <int TN>
int launch_sk_tn(const SkParams& p, dim3 grid, bool deep, cudaStream_t st) {
  return deep ? launch_sk<TN, SkCfg<TN>::NST_DEEP>(p, grid, st) : launch_sk<TN, SkCfg<TN>::NST_SHALLOW>(p, grid, st);
}

}  // namespace

void set_skinny_max_m(int v) { g_skinny_max_m = v; }

bool gemm_skinny_supported(const GemmArgs& g) {
  if (g_skinny_max_m <= 0 || g.M <= 0 || g.M > g_skinny_max_m) return false;
  if (g.tap_w || g.groups != 1) return false;
  if (g.K <= 0 || g.K % SK_KC != 0 || g.N % 8 != 0) return false;
  if (g.ldw % 8 || g.a_map.rs % 8 || g.a_map.bs % 8 || ((uintptr_t)g.A % 16) || ((uintptr_t)g.W % 16)) return false;
  if (g.qkv_mode) {
    if (g.N % 64 || g.qkv_C % 64 || g.kv_map.rs % 8 || g.kv_map.bs % 8 || g.kv_layer_stride % 8) return false;
    if (((uintptr_t)g.kcache % 16) || ((uintptr_t)g.vcache % 16) || (g.qbuf && ((uintptr_t)g.qbuf % 16))) return false;
    return true;
  }
  // pairwise (8-byte fp32 / 4-byte bf16) accesses in the epilogue
  if (g.c_map.rs % 2 || g.c_map.bs % 2) return false;
  if (g.out32 && ((uintptr_t)g.out32 % 8)) return false;
  if (g.out_act && ((uintptr_t)g.out_act % 8)) return false;
  if (g.gate && (g.gate_map.rs % 2 || g.gate_map.bs % 2 || ((uintptr_t)g.gate % 8))) return false;
  if (g.resid && (g.resid_map.rs % 2 || g.resid_map.bs % 2 || ((uintptr_t)g.resid % 8))) return false;
  return true;
}

int launch_gemm_skinny(const GemmArgs& g, cudaStream_t st) {
  AT_REQUIRE(gemm_skinny_supported(g), "gemm_skinny: unsupported shape");
  AT_REQUIRE(g.A && g.W && (g.out32 || g.out_act || g.qkv_mode), "gemm_skinny: null operand");
  if (g.qkv_mode) {
    AT_REQUIRE((g.qkv_mode == 1 || g.qkv_mode == 2) && g.qkv_C > 0 && g.N % ((g.qkv_mode == 1 ? 3 : 2) * g.qkv_C) == 0 && g.kcache &&
               g.vcache && (g.qkv_mode == 2 || (g.qbuf && g.head_scale)) && g.act == ACT_NONE && !g.gate && !g.resid,
               "gemm_skinny: bad fused q/k/v arguments");
  }
  const DevCtx* dc = nullptr;
  AT_TRY(dev_ctx(&dc));
  const int g_sk_num_sms = dc->num_sms;
  SkParams p;
  p.A = (const bf16*)g.A; p.a_map = g.a_map; p.W = (const bf16*)g.W; p.ldw = g.ldw;
  p.M = g.M; p.N = g.N; p.n_chunks = g.K / SK_KC;
  p.bias = g.bias; p.act = g.act; p.gate = g.gate; p.gate_dt = g.gate_dt; p.gate_map = g.gate_map;
  p.resid = g.resid; p.resid_map = g.resid_map; p.out32 = g.out32; p.out_act = g.out_act; p.out_act_dt = g.out_act_dt; p.c_map = g.c_map;
  p.qkv_mode = g.qkv_mode; p.qkv_C = g.qkv_C; p.head_scale = g.head_scale; p.qbuf = (bf16*)g.qbuf; p.kcache = (bf16*)g.kcache;
  p.vcache = (bf16*)g.vcache; p.kv_map = g.kv_map; p.kv_layer_stride = g.kv_layer_stride;
  const int slabs = ceil_div(g.M, SK_BM);
  // widest column tile that still gives ~100 CTAs (every CTA re-reads its 64-row A slab: wide tiles cut that traffic)
  int TN = 8;
  if (g.qkv_mode) TN = 64;
  else {
    const int cand[4] = {64, 32, 16, 8};
    for (int i = 0; i < 4; ++i)
      if (g.N % cand[i] == 0 && (long)slabs * (g.N / cand[i]) >= 90) { TN = cand[i]; break; }
  }
  const dim3 grid(g.N / TN, slabs);
  const bool deep = (long)grid.x * grid.y <= g_sk_num_sms;
  g_trace_dims[0] = g.M; g_trace_dims[1] = g.N; g_trace_dims[2] = g.K;
  switch (TN) {
    case 64: return launch_sk_tn<64>(p, grid, deep, st);
    case 32: return launch_sk_tn<32>(p, grid, deep, st);
    case 16: return launch_sk_tn<16>(p, grid, deep, st);
    default: return launch_sk_tn<8>(p, grid, deep, st);
  }
}

}  // namespace artalk
