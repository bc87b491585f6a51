// Latency path of the recurrence: skinny-M bf16 GEMM and few-query attention.
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
//  * attn_few_kernel: softmax(q k^T) v for <= 8 query rows per (clip, head) and <= 256 resident keys
//    (app/transformer.py:75-77 in the KV-cached schedule: no mask). One 256-thread CTA per (clip, head): a thread scores
//    whole key rows (8 x 16-byte loads in flight per thread), warps reduce the softmax, and the P V product reads V as
//    16-byte pieces with 4 keys per warp instruction. It is an HBM-streaming kernel (36 MB of K/V per launch at 64 clips).
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
int g_sk_num_sms = 0;

template <int TN, int NST>
int launch_sk(const SkParams& p, dim3 grid, cudaStream_t st) {
  constexpr int smem = NST * SkCfg<TN>::STAGE;
  static bool attr_set = false;
  if (!attr_set) {
    AT_CUDA(cudaFuncSetAttribute(skinny_gemm_kernel<TN, NST>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
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
  if (!g_sk_num_sms) {
    int dev = 0;
    AT_CUDA(cudaGetDevice(&dev));
    AT_CUDA(cudaDeviceGetAttribute(&g_sk_num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
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

// ---------------------------------------------------------------- few-query attention
namespace {
constexpr int AF_MAXQ = 8, AF_MAXK = 256, AF_THREADS = 256;

template <int LQ>
__global__ void __launch_bounds__(AF_THREADS) attn_few_kernel(const AttnArgs a) {
  // The kernel is one memory round trip: every K row (one per thread) and every V piece (4 key rows per warp instruction,
  // 8 instructions per warp) is requested into registers up front; scores, softmax and P V then run out of registers and
  // shared memory. (A first version loaded V after the softmax in dependent batches: 19.6 us against a 6 us DRAM floor.)
  __shared__ float qs[LQ][64];
  __shared__ float sc[LQ][AF_MAXK];
  __shared__ float part[AF_THREADS / 32][LQ][64];
  __shared__ float inv[LQ];
  pdl_enter();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int head = blockIdx.x, seq = blockIdx.y;
  const int lq = a.lq, lk = a.lk;
  const bf16* qb = reinterpret_cast<const bf16*>(a.q) + (int64_t)seq * a.q_ss + head * 64;
  const bf16* kb = reinterpret_cast<const bf16*>(a.k) + (int64_t)seq * a.k_ss + head * 64;
  const bf16* vb = reinterpret_cast<const bf16*>(a.v) + (int64_t)seq * a.v_ss + head * 64;
  const int ksub = lane >> 3, dg = lane & 7;
  constexpr int VIT = AF_MAXK / AF_THREADS * 8;             // V iterations: 32 key rows per iteration over the CTA
  uint4 vraw[VIT], kraw[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) kraw[c] = tid < lk ? __ldg(reinterpret_cast<const uint4*>(kb + (int64_t)tid * a.k_rs) + c) : make_uint4(0, 0, 0, 0);
#pragma unroll
  for (int it = 0; it < VIT; ++it) {
    const int j = it * (AF_THREADS / 8) + warp * 4 + ksub;
    vraw[it] = j < lk ? __ldg(reinterpret_cast<const uint4*>(vb + (int64_t)j * a.v_rs + dg * 8)) : make_uint4(0, 0, 0, 0);
  }
  for (int i = tid; i < lq * 64; i += AF_THREADS) {
    const int r = i >> 6, d = i & 63;
    qs[r][d] = __bfloat162float(qb[(int64_t)r * a.q_rs + d]) * a.scale;
  }
  __syncthreads();
  // ---- scores: one key row per thread
  if (tid < lk) {
    float s[LQ];
#pragma unroll
    for (int i = 0; i < LQ; ++i) s[i] = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const uint32_t w[4] = {kraw[c].x, kraw[c].y, kraw[c].z, kraw[c].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w[e]);
        const float k0 = __low2float(h), k1 = __high2float(h);
        const int d = c * 8 + e * 2;
#pragma unroll
        for (int i = 0; i < LQ; ++i)
          if (i < lq) s[i] = fmaf(qs[i][d + 1], k1, fmaf(qs[i][d], k0, s[i]));
      }
    }
#pragma unroll
    for (int i = 0; i < LQ; ++i)
      if (i < lq) sc[i][tid] = s[i];
  }
  __syncthreads();
  // ---- softmax: one query row per warp
  for (int i = warp; i < lq; i += AF_THREADS / 32) {
    float m = -INFINITY;
    for (int j = lane; j < lk; j += 32) m = fmaxf(m, sc[i][j]);
    m = warp_max(m);
    float sum = 0.f;
    for (int j = lane; j < lk; j += 32) {
      const float e = __expf(sc[i][j] - m);
      sc[i][j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    if (lane == 0) inv[i] = 1.0f / sum;
  }
  __syncthreads();
  // ---- P V out of the registers (lane -> key lane/8 of the warp's 4, dims (lane%8)*8 .. +8)
  float acc[LQ][8];
#pragma unroll
  for (int i = 0; i < LQ; ++i)
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[i][e] = 0.f;
#pragma unroll
  for (int it = 0; it < VIT; ++it) {
    const int j = it * (AF_THREADS / 8) + warp * 4 + ksub;
    if (j < lk) {
      const uint32_t w[4] = {vraw[it].x, vraw[it].y, vraw[it].z, vraw[it].w};
      float vv[8];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w[e]);
        vv[2 * e] = __low2float(h); vv[2 * e + 1] = __high2float(h);
      }
#pragma unroll
      for (int i = 0; i < LQ; ++i) {
        if (i < lq) {
          const float pij = sc[i][j];
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[i][e] = fmaf(pij, vv[e], acc[i][e]);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < LQ; ++i) {
    if (i < lq) {                                           // warp-uniform
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float t = acc[i][e];
        t += __shfl_xor_sync(0xffffffffu, t, 8);
        t += __shfl_xor_sync(0xffffffffu, t, 16);
        if (ksub == 0) part[warp][i][dg * 8 + e] = t;
      }
    }
  }
  __syncthreads();
  bf16* ob = reinterpret_cast<bf16*>(a.out) + (int64_t)seq * a.o_ss + head * 64;
  for (int i = tid; i < lq * 32; i += AF_THREADS) {
    const int r = i >> 5, d = (i & 31) * 2;
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int w = 0; w < AF_THREADS / 32; ++w) { s0 += part[w][r][d]; s1 += part[w][r][d + 1]; }
    *reinterpret_cast<__nv_bfloat162*>(ob + (int64_t)r * a.o_rs + d) = __floats2bfloat162_rn(s0 * inv[r], s1 * inv[r]);
  }
}

int g_attn_few_max_lq = 0;         // option "attn_few_max_lq" (1..8 = on). Off: measured 19.6 / 24.7 us against 19.1 us for the tcgen05 kernel
int g_ar_small_on = 0;              // option "ar_small". Off: measured 0.91 / 1.44 ms per 1- / 5-token step against ~1.0 ms as separate kernels

// ---------------------------------------------------------------- whole-stack kernel for the few-token scale steps
// One launch runs all AR blocks (+ the head) of a scale step with <= 8 new tokens per clip (app/transformer.py:30-79 x depth,
// app/models.py:103,145-148). 7 phases per block (AdaLN1 | q/k/v GEMM + head norm + cache scatter | attention | out-proj |
// gated residual + AdaLN2 | FFN1 + GELU | FFN2), separated by a grid-wide barrier (one atomic counter; every CTA is resident:
// grid <= #SMs, one CTA per SM). As separate kernels each phase cost a dependent launch (~5 us in the chunk graph) plus its
// own ramp; here a phase costs the barrier (~1.5 us) plus one memory round trip. The GEMM phases reuse sk_tile; attention
// stages each (clip, head) item's K and V (<= 192 keys) in shared memory with cp.async, 4 items per CTA at a time.
// Data produced inside the kernel is read through L2 (cp.async.cg / ld.global.cg): L1 is not coherent across SMs.
struct ArLayerW {
  const bf16 *wqkv, *wproj, *wff1, *wff2;
  const float *bqkv, *bproj, *bff1, *bff2, *head_scale;
};
struct ArSmallParams {
  int B, n_new, M, C, NL, lk, heads;
  float* x; float* y; bf16* u; bf16* qbuf; bf16* o; bf16* f;
  const bf16* ada; RowMap ada_map;             // AdaLN rows of this scale's tokens, block 0 (block l at + l * 6C, head at + NL * 6C)
  bf16* kcache; bf16* vcache; int64_t kv_layer_stride, kv_seq_stride, kv_new_off; RowMap kv_new_map;   // caches of block 0; rows of the new tokens
  const ArLayerW* layers;
  const bf16* whead; const float* bhead; float* logits; RowMap logits_map; int n_logits;
  unsigned int* sync;                           // [0] barrier arrivals, [1] exits (both 0 between launches)
  unsigned int* err_flag;
  int tn_proj, tn_ff1, tn_ff2;
  float eps;
  unsigned long long* dbg;                      // developer trace (ARTALK_MG_DEBUG): CTA 0 stamps globaltimer around every barrier
};

constexpr int MG_KROWS = 192;                   // keys per staged item
constexpr int MG_ITEM_BYTES = 2 * MG_KROWS * 128;
constexpr int mg_max(int a, int b) { return a > b ? a : b; }
constexpr int MG_SMEM = mg_max(mg_max(SkCfg<64>::NST_DEEP * SkCfg<64>::STAGE, SkCfg<32>::NST_DEEP * SkCfg<32>::STAGE),
                               mg_max(mg_max(SkCfg<16>::NST_DEEP * SkCfg<16>::STAGE, SkCfg<8>::NST_DEEP * SkCfg<8>::STAGE),
                                      4 * MG_ITEM_BYTES + 4 * 2048 + 256));
static_assert(MG_SMEM <= 227 * 1024, "shared memory of the whole-stack kernel");

__device__ __forceinline__ void grid_barrier(unsigned int* ctr, unsigned int target, unsigned int* err) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(ctr, 1u);
    unsigned int spins = 0, v;
    while (true) {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
      if (v >= target) break;
      if (++spins > 8000000u) {                 // a protocol bug traps (launch error) instead of hanging the GPU
        if (err) atomicExch(err, 0xBA220000u | (target & 0xffffu));
        __trap();
      }
    }
    __threadfence();
  }
  __syncthreads();
}

// rows r = warp, warp + #warps, ...: x[r] (+= gate * y[r]) -> LN -> * (1 + scale) + shift -> bf16 u[r]   (norms.cu::adaln_kernel)
__device__ __forceinline__ void mg_adaln(const ArSmallParams& p, const bf16* ada_l, int scale_off, int shift_off, bool pending, int gate_off) {
  const int lane = threadIdx.x & 31;
  const int gw = (int)blockIdx.x * (SK_THREADS / 32) + (threadIdx.x >> 5), nw = (int)gridDim.x * (SK_THREADS / 32);
  for (int row = gw; row < p.M; row += nw) {
    float* xr = p.x + (int64_t)row * 768;
    const bf16* ar = ada_l + p.ada_map.off(row);
    float v[6][4];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const int c = (i * 32 + lane) * 4;
      const float4 t = __ldcg(reinterpret_cast<const float4*>(xr + c));
      v[i][0] = t.x; v[i][1] = t.y; v[i][2] = t.z; v[i][3] = t.w;
      if (pending) {
        const float4 yv = __ldcg(reinterpret_cast<const float4*>(p.y + (int64_t)row * 768 + c));
        float gv[4];
        load4(ar + gate_off + c, gv);
        v[i][0] = fmaf(yv.x, gv[0], v[i][0]); v[i][1] = fmaf(yv.y, gv[1], v[i][1]);
        v[i][2] = fmaf(yv.z, gv[2], v[i][2]); v[i][3] = fmaf(yv.w, gv[3], v[i][3]);
        __stcg(reinterpret_cast<float4*>(xr + c), make_float4(v[i][0], v[i][1], v[i][2], v[i][3]));
      }
      s += (v[i][0] + v[i][1]) + (v[i][2] + v[i][3]);
    }
    const float mean = warp_sum(s) * (1.0f / 768);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) { const float d = v[i][j] - mean; q += d * d; }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / 768) + p.eps);
    bf16* orow = p.u + (int64_t)row * 768;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const int c = (i * 32 + lane) * 4;
      float sc[4], sh[4], o4[4];
      load4(ar + scale_off + c, sc);
      load4(ar + shift_off + c, sh);
#pragma unroll
      for (int j = 0; j < 4; ++j) o4[j] = (v[i][j] - mean) * rstd * (1.0f + sc[j]) + sh[j];
      store4(orow + c, o4);
    }
  }
}

template <int TN>
__device__ __forceinline__ void mg_gemm_tn(const SkParams& g, uint8_t* smem, bool w_pre) {
  const int nt_n = g.N / TN, slabs = (g.M + SK_BM - 1) / SK_BM;
  int t = blockIdx.x;
  if (t < nt_n * slabs && w_pre) {                         // first tile: its W ring fill was issued before the grid barrier
    sk_tile<TN, SkCfg<TN>::NST_DEEP, false, true>(g, (t % nt_n) * TN, (t / nt_n) * SK_BM, smem);
    t += gridDim.x;
  }
  for (; t < nt_n * slabs; t += gridDim.x) {
    __syncthreads();                                       // the previous tile's epilogue is done with the shared memory
    sk_tile<TN, SkCfg<TN>::NST_DEEP, false>(g, (t % nt_n) * TN, (t / nt_n) * SK_BM, smem);
  }
}
__device__ __forceinline__ void mg_gemm(const SkParams& g, int tn, uint8_t* smem, bool w_pre) {
  switch (tn) {
    case 64: mg_gemm_tn<64>(g, smem, w_pre); break;
    case 32: mg_gemm_tn<32>(g, smem, w_pre); break;
    case 16: mg_gemm_tn<16>(g, smem, w_pre); break;
    default: mg_gemm_tn<8>(g, smem, w_pre); break;
  }
}
// issue the W ring fill of this CTA's first tile of an upcoming GEMM phase (weights are constants: no dependency on the
// phases in between). The shared memory must be idle: call after the previous user has synchronised.
template <int TN>
__device__ __forceinline__ void mg_prefetch_tn(const bf16* W, int64_t ldw, int N, int nch, int M, uint8_t* smem) {
  const int nt_n = N / TN, slabs = (M + SK_BM - 1) / SK_BM;
  const int t = blockIdx.x;
  if (t < nt_n * slabs) sk_prefetch_w<TN, SkCfg<TN>::NST_DEEP>(W, ldw, nch, (t % nt_n) * TN, smem);
}
__device__ __forceinline__ void mg_prefetch(const bf16* W, int64_t ldw, int N, int nch, int M, int tn, uint8_t* smem) {
  __syncthreads();
  switch (tn) {
    case 64: mg_prefetch_tn<64>(W, ldw, N, nch, M, smem); break;
    case 32: mg_prefetch_tn<32>(W, ldw, N, nch, M, smem); break;
    case 16: mg_prefetch_tn<16>(W, ldw, N, nch, M, smem); break;
    default: mg_prefetch_tn<8>(W, ldw, N, nch, M, smem); break;
  }
}

__device__ __forceinline__ void group_bar(int grp) { asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "r"(64) : "memory"); }

// softmax(q k^T) v for the new tokens of every (clip, head): 4 items per CTA at a time, 64 threads per item
template <int LQ>
__device__ __forceinline__ void mg_attention(const ArSmallParams& p, int layer, uint8_t* smem) {
  const int tid = threadIdx.x, grp = tid >> 6, t = tid & 63, w2 = t >> 5, lane = tid & 31;
  const int lq = p.n_new, lk = p.lk, n_items = p.B * p.heads;
  uint8_t* Ks = smem + grp * MG_ITEM_BYTES;                // [lk][128 B], 16-byte pieces XOR-swizzled with the row
  uint8_t* Vs = Ks + MG_KROWS * 128;                       // [lk][128 B]
  float* P = reinterpret_cast<float*>(Ks);                 // [8][192] fp32 scores / probabilities (reuses K once it is consumed)
  float* qs = reinterpret_cast<float*>(smem + 4 * MG_ITEM_BYTES + grp * 2048);   // [8][64]; later the cross-warp partial
  float* inv = reinterpret_cast<float*>(smem + 4 * MG_ITEM_BYTES + 4 * 2048) + grp * 8;
  const uint32_t ks_u = sm_u32(Ks), vs_u = sm_u32(Vs);
  const bf16* kc = p.kcache + (int64_t)layer * p.kv_layer_stride;
  const bf16* vc = p.vcache + (int64_t)layer * p.kv_layer_stride;
  for (int it0 = (int)blockIdx.x * 4; it0 < n_items; it0 += (int)gridDim.x * 4) {
    const int item = it0 + grp;
    if (item >= n_items) continue;                         // group-uniform: only group barriers below
    const int clip = item / p.heads, head = item - clip * p.heads;
    const bf16* kb = kc + (int64_t)clip * p.kv_seq_stride + head * 64;
    const bf16* vb = vc + (int64_t)clip * p.kv_seq_stride + head * 64;
    for (int i = t; i < lk * 8; i += 64) {
      const int j = i >> 3, pc = i & 7;
      cp_async16(ks_u + j * 128 + ((pc ^ (j & 7)) << 4), kb + (int64_t)j * p.C + pc * 8);
    }
    cp_async_commit();
    for (int i = t; i < lk * 8; i += 64) {
      const int j = i >> 3, pc = i & 7;
      cp_async16(vs_u + j * 128 + (pc << 4), vb + (int64_t)j * p.C + pc * 8);
    }
    cp_async_commit();
    for (int i = 0; i < lq; ++i) {
      const unsigned short raw = __ldcg(reinterpret_cast<const unsigned short*>(p.qbuf) + ((int64_t)(clip * lq + i) * p.C + head * 64 + t));
      qs[i * 64 + t] = __bfloat162float(__ushort_as_bfloat16(raw));
    }
    cp_async_wait<1>();
    group_bar(grp);                                        // K and q visible to the item's 64 threads
    float s[3][LQ];
#pragma unroll
    for (int kk = 0; kk < 3; ++kk) {
      const int j = t + 64 * kk;
#pragma unroll
      for (int i = 0; i < LQ; ++i) s[kk][i] = 0.f;
      if (j < lk) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint4 raw = *reinterpret_cast<const uint4*>(Ks + j * 128 + ((c ^ (j & 7)) << 4));
          const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w[e]);
            const float k0 = __low2float(h), k1 = __high2float(h);
            const int d = c * 8 + e * 2;
#pragma unroll
            for (int i = 0; i < LQ; ++i)
              if (i < lq) s[kk][i] = fmaf(qs[i * 64 + d + 1], k1, fmaf(qs[i * 64 + d], k0, s[kk][i]));
          }
        }
      }
    }
    group_bar(grp);                                        // everyone is done reading K: its space becomes P
#pragma unroll
    for (int kk = 0; kk < 3; ++kk) {
      const int j = t + 64 * kk;
      if (j < lk) {
#pragma unroll
        for (int i = 0; i < LQ; ++i)
          if (i < lq) P[i * MG_KROWS + j] = s[kk][i];
      }
    }
    group_bar(grp);
    for (int i = w2; i < lq; i += 2) {
      float m = -INFINITY;
      for (int j = lane; j < lk; j += 32) m = fmaxf(m, P[i * MG_KROWS + j]);
      m = warp_max(m);
      float sum = 0.f;
      for (int j = lane; j < lk; j += 32) {
        const float e = __expf(P[i * MG_KROWS + j] - m);
        P[i * MG_KROWS + j] = e;
        sum += e;
      }
      sum = warp_sum(sum);
      if (lane == 0) inv[i] = 1.0f / sum;
    }
    cp_async_wait<0>();
    group_bar(grp);                                        // probabilities, 1/sum and V visible
    float acc[LQ][2];
#pragma unroll
    for (int i = 0; i < LQ; ++i) acc[i][0] = acc[i][1] = 0.f;
#pragma unroll 4
    for (int j = w2; j < lk; j += 2) {
      const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(Vs + j * 128 + lane * 4);
      const float v0 = __low2float(h), v1 = __high2float(h);
#pragma unroll
      for (int i = 0; i < LQ; ++i) {
        if (i < lq) {
          const float pij = P[i * MG_KROWS + j];
          acc[i][0] = fmaf(pij, v0, acc[i][0]);
          acc[i][1] = fmaf(pij, v1, acc[i][1]);
        }
      }
    }
    if (w2 == 1) {
#pragma unroll
      for (int i = 0; i < LQ; ++i)
        if (i < lq) *reinterpret_cast<float2*>(&qs[i * 64 + lane * 2]) = make_float2(acc[i][0], acc[i][1]);
    }
    group_bar(grp);
    if (w2 == 0) {
#pragma unroll
      for (int i = 0; i < LQ; ++i) {
        if (i < lq) {
          const float2 o2 = *reinterpret_cast<const float2*>(&qs[i * 64 + lane * 2]);
          const float iv = inv[i];
          *reinterpret_cast<__nv_bfloat162*>(p.o + (int64_t)(clip * lq + i) * p.C + head * 64 + lane * 2) =
              __floats2bfloat162_rn((acc[i][0] + o2.x) * iv, (acc[i][1] + o2.y) * iv);
        }
      }
    }
    group_bar(grp);                                        // the item's shared memory is free for the next round
  }
}

__global__ void __launch_bounds__(SK_THREADS, 1) ar_small_kernel(const ArSmallParams p) {
  extern __shared__ __align__(128) uint8_t mg_smem[];
  const unsigned int nb = gridDim.x;
  unsigned int bar = 0;
  auto stamp = [&](int slot) {
    if (p.dbg && blockIdx.x == 0 && threadIdx.x == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      p.dbg[slot] = t;
    }
  };
  auto sync_grid = [&]() {
    ++bar;
    if (p.dbg) { __syncthreads(); stamp(2 * bar); }
    grid_barrier(p.sync, bar * nb, p.err_flag);
    stamp(2 * bar + 1);
  };
  stamp(1);
  const int C = p.C;
  SkParams g;
  g.a_map = plain_rows(C); g.M = p.M; g.gate = nullptr; g.gate_dt = DT_BF16; g.gate_map = plain_rows(0);
  g.resid = nullptr; g.resid_map = plain_rows(0); g.head_scale = nullptr; g.qbuf = nullptr; g.kcache = nullptr; g.vcache = nullptr;
  g.kv_map = p.kv_new_map; g.kv_layer_stride = 0; g.qkv_C = C;
  if (p.NL > 0) mg_prefetch(p.layers[0].wqkv, C, 3 * C, C / SK_KC, p.M, 64, mg_smem);
  for (int l = 0; l < p.NL; ++l) {
    const ArLayerW& w = p.layers[l];
    const bf16* ada_l = p.ada + (int64_t)l * 6 * C;          // chunk order g1, g2, s1, s2, b1, b2 (app/transformer.py:32)
    mg_adaln(p, ada_l, 2 * C, 4 * C, l > 0, -5 * C);         // x += gamma2(l-1) * y ; u = AdaLN1(x)
    sync_grid();
    g.A = p.u; g.W = w.wqkv; g.ldw = C; g.N = 3 * C; g.n_chunks = C / SK_KC; g.bias = w.bqkv; g.act = ACT_NONE;
    g.out32 = nullptr; g.out_act = nullptr; g.out_act_dt = DT_BF16; g.c_map = plain_rows(0);
    g.qkv_mode = 1; g.head_scale = w.head_scale; g.qbuf = p.qbuf;
    g.kcache = p.kcache + (int64_t)l * p.kv_layer_stride + p.kv_new_off; g.vcache = p.vcache + (int64_t)l * p.kv_layer_stride + p.kv_new_off;
    mg_gemm(g, 64, mg_smem, true);
    sync_grid();
    if (p.n_new == 1) mg_attention<1>(p, l, mg_smem);
    else if (p.n_new == 2) mg_attention<2>(p, l, mg_smem);
    else if (p.n_new <= 5) mg_attention<5>(p, l, mg_smem);
    else mg_attention<8>(p, l, mg_smem);
    mg_prefetch(w.wproj, C, C, C / SK_KC, p.M, p.tn_proj, mg_smem);
    sync_grid();
    g.qkv_mode = 0;
    g.A = p.o; g.W = w.wproj; g.N = C; g.bias = w.bproj; g.out32 = p.y; g.c_map = plain_rows(C);
    mg_gemm(g, p.tn_proj, mg_smem, true);
    mg_prefetch(w.wff1, C, 4 * C, C / SK_KC, p.M, p.tn_ff1, mg_smem);      // stays in flight across the AdaLN2 phase
    sync_grid();
    mg_adaln(p, ada_l, 3 * C, 5 * C, true, 0);               // x += gamma1 * y ; u = AdaLN2(x)
    sync_grid();
    g.A = p.u; g.W = w.wff1; g.N = 4 * C; g.bias = w.bff1; g.act = ACT_GELU_TANH; g.out32 = nullptr; g.out_act = p.f;
    g.c_map = plain_rows(4 * C);
    mg_gemm(g, p.tn_ff1, mg_smem, true);
    mg_prefetch(w.wff2, 4 * C, C, 4 * C / SK_KC, p.M, p.tn_ff2, mg_smem);
    sync_grid();
    g.A = p.f; g.a_map = plain_rows(4 * C); g.W = w.wff2; g.ldw = 4 * C; g.N = C; g.n_chunks = 4 * C / SK_KC; g.bias = w.bff2;
    g.act = ACT_NONE; g.out32 = p.y; g.out_act = nullptr; g.c_map = plain_rows(C);
    mg_gemm(g, p.tn_ff2, mg_smem, true);
    g.a_map = plain_rows(C);
    if (l + 1 < p.NL) mg_prefetch(p.layers[l + 1].wqkv, C, 3 * C, C / SK_KC, p.M, 64, mg_smem);
    else mg_prefetch(p.whead, C, p.n_logits, C / SK_KC, p.M, 8, mg_smem);
    sync_grid();
  }
  // head: AdaLN (scale, shift order; app/models.py:147) -> Linear C -> 2 * code_dim
  const bf16* ada_h = p.ada + (int64_t)p.NL * 6 * C;
  mg_adaln(p, ada_h, 0, C, p.NL > 0, -5 * C);
  sync_grid();
  g.A = p.u; g.W = p.whead; g.ldw = C; g.N = p.n_logits; g.n_chunks = C / SK_KC; g.bias = p.bhead; g.act = ACT_NONE;
  g.out32 = p.logits; g.out_act = nullptr; g.c_map = p.logits_map; g.qkv_mode = 0;
  mg_gemm(g, 8, mg_smem, p.NL > 0);
  // leave the counters at zero for the next launch: the last CTA to get here resets them
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(p.sync + 1, 1u) == nb - 1) { p.sync[0] = 0u; p.sync[1] = 0u; __threadfence(); }
  }
}

int g_mg_num_sms = 0;
}  // namespace

void set_attn_few_max_lq(int v) { g_attn_few_max_lq = v < AF_MAXQ ? v : AF_MAXQ; }

bool attention_few_supported(const AttnArgs& a) {
  if (a.dt != DT_BF16 || a.head_dim != 64 || a.split != 0) return false;
  if (a.lq <= 0 || a.lq > g_attn_few_max_lq || a.lk <= 0 || a.lk > AF_MAXK) return false;
  if (a.k_rs % 8 || a.v_rs % 8 || a.k_ss % 8 || a.v_ss % 8 || a.o_rs % 2 || a.o_ss % 2) return false;
  if (((uintptr_t)a.k % 16) || ((uintptr_t)a.v % 16) || ((uintptr_t)a.out % 4)) return false;
  return true;
}

int launch_attention_few(const AttnArgs& a, cudaStream_t st) {
  AT_REQUIRE(attention_few_supported(a), "attention_few: unsupported shape");
  g_trace_dims[0] = a.n_seq * a.n_heads; g_trace_dims[1] = a.lq; g_trace_dims[2] = a.lk;
  const dim3 grid(a.n_heads, a.n_seq);
  if (a.lq == 1) AT_CUDA(launch_k(attn_few_kernel<1>, grid, dim3(AF_THREADS), 0, st, a));
  else if (a.lq == 2) AT_CUDA(launch_k(attn_few_kernel<2>, grid, dim3(AF_THREADS), 0, st, a));
  else if (a.lq <= 5) AT_CUDA(launch_k(attn_few_kernel<5>, grid, dim3(AF_THREADS), 0, st, a));
  else AT_CUDA(launch_k(attn_few_kernel<8>, grid, dim3(AF_THREADS), 0, st, a));
  AT_LAUNCH_CHECK();
  return AT_OK;
}


// ---- host side of the whole-stack kernel
bool ar_small_supported(int B, int n_new, int C, int lk, int n_logits) {
  return g_ar_small_on && n_new >= 1 && n_new <= AF_MAXQ && B >= 1 && (long)B * n_new <= 2048 && C == 768 && lk <= MG_KROWS && n_logits % 8 == 0;
}

int launch_ar_small(const ArSmallArgs& a, cudaStream_t st) {
  AT_REQUIRE(ar_small_supported(a.B, a.n_new, a.C, a.lk, a.n_logits), "ar_small: unsupported shape");
  if (!g_mg_num_sms) {
    int dev = 0;
    AT_CUDA(cudaGetDevice(&dev));
    AT_CUDA(cudaDeviceGetAttribute(&g_mg_num_sms, cudaDevAttrMultiProcessorCount, dev));
    AT_CUDA(cudaFuncSetAttribute(ar_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MG_SMEM));
    int per_sm = 0;
    AT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ar_small_kernel, SK_THREADS, MG_SMEM));
    AT_REQUIRE(per_sm >= 1, "ar_small: kernel does not fit an SM");
  }
  ArSmallParams p;
  p.B = a.B; p.n_new = a.n_new; p.M = a.B * a.n_new; p.C = a.C; p.NL = a.NL; p.lk = a.lk; p.heads = a.C / 64;
  p.x = a.x; p.y = a.y; p.u = (bf16*)a.u; p.qbuf = (bf16*)a.qbuf; p.o = (bf16*)a.o; p.f = (bf16*)a.f;
  p.ada = (const bf16*)a.ada; p.ada_map = a.ada_map;
  p.kcache = (bf16*)a.kcache; p.vcache = (bf16*)a.vcache; p.kv_layer_stride = a.kv_layer_stride; p.kv_seq_stride = a.kv_seq_stride;
  p.kv_new_map = a.kv_new_map; p.kv_new_off = a.kv_new_off;
  p.layers = (const ArLayerW*)a.layer_table; p.whead = (const bf16*)a.whead; p.bhead = a.bhead; p.logits = a.logits;
  p.logits_map = a.logits_map; p.n_logits = a.n_logits;
  p.sync = a.sync; p.err_flag = nullptr; p.eps = a.eps;
  static unsigned long long* dbg_buf = nullptr;
  static const bool dbg_on = getenv("ARTALK_MG_DEBUG") != nullptr;
  if (dbg_on && !dbg_buf) AT_CUDA(cudaMalloc((void**)&dbg_buf, 512 * sizeof(unsigned long long)));
  p.dbg = dbg_on ? dbg_buf : nullptr;
  const int slabs = ceil_div(p.M, SK_BM);
  auto pick = [&](int N) {
    const int cand[4] = {64, 32, 16, 8};
    for (int i = 0; i < 4; ++i)
      if (N % cand[i] == 0 && (long)slabs * (N / cand[i]) >= 90) return cand[i];
    return 8;
  };
  p.tn_proj = pick(a.C); p.tn_ff1 = pick(4 * a.C); p.tn_ff2 = pick(a.C);
  g_trace_dims[0] = p.M; g_trace_dims[1] = a.NL; g_trace_dims[2] = a.lk;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(g_mg_num_sms); cfg.blockDim = dim3(SK_THREADS); cfg.dynamicSmemBytes = MG_SMEM; cfg.stream = st;
  // cooperative launch: every CTA is resident at once (the grid barrier spins), also when another stream competes for SMs.
  // No PDL: the kernel needs its predecessor's data from the first phase on.
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeCooperative; at[0].val.cooperative = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  AT_CUDA(cudaLaunchKernelEx(&cfg, ar_small_kernel, p));
  AT_LAUNCH_CHECK();
  if (dbg_on) {                                             // eager launches only (ARTALK graphs off): print CTA 0's phase times
    std::vector<unsigned long long> h(512);
    AT_CUDA(cudaStreamSynchronize(st));
    AT_CUDA(cudaMemcpy(h.data(), dbg_buf, 512 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    const int nbar = a.NL * 7 + 1;
    fprintf(stderr, "[ar_small M=%d] work/wait ns per phase (CTA 0):", p.M);
    for (int b = 1; b <= nbar && 2 * b + 1 < 512; ++b)
      fprintf(stderr, "%s %llu/%llu", (b - 1) % 7 == 0 ? "\n  " : "", h[2 * b] - h[2 * b - 1], h[2 * b + 1] - h[2 * b]);
    fprintf(stderr, "\n");
  }
  return AT_OK;
}
size_t ar_layer_table_bytes(int n_layers) { return sizeof(ArLayerW) * (size_t)n_layers; }
void ar_layer_table_fill(void* host_dst, int l, const void* wqkv, const void* wproj, const void* wff1, const void* wff2, const float* bqkv,
                         const float* bproj, const float* bff1, const float* bff2, const float* head_scale) {
  ArLayerW& w = reinterpret_cast<ArLayerW*>(host_dst)[l];
  w.wqkv = (const bf16*)wqkv; w.wproj = (const bf16*)wproj; w.wff1 = (const bf16*)wff1; w.wff2 = (const bf16*)wff2;
  w.bqkv = bqkv; w.bproj = bproj; w.bff1 = bff1; w.bff2 = bff2; w.head_scale = head_scale;
}
void set_ar_small(int on) { g_ar_small_on = on; }

}  // namespace artalk
