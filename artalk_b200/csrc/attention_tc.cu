// Tensor-core attention for the path's short sequences (<= 384 keys, head_dim 64, bf16) on sm_100a.
//
// One work item = (sequence, head, 128-row query tile); the whole key range fits on chip, so softmax is single pass:
//   TMA      : Q tile [128 x 64], K [lk_pad x 64], V [lk_pad x 64] -> shared memory (128B swizzle), double buffered
//   MMA 1    : S = Q K^T        tcgen05.mma, A/B from shared memory (both K-major), accumulator S in TMEM (lk_pad columns)
//   softmax  : 4 warps, one query row per thread: tcgen05.ld S -> max, exp2, row sum -> P (bf16) written back with
//              tcgen05.st over the S columns it has already consumed
//   MMA 2    : O = P V          tcgen05.mma with A = P from TMEM, B = V from shared memory consumed MN-major
//   epilogue : tcgen05.ld O, scale by 1/rowsum, bf16 stores
// Persistent CTAs (grid = min(items, #SMs)); the TMA loads of item i+1 overlap the compute of item i.
// Masking: keys >= lk never contribute (TMA zero-fills them and P is forced to 0); `split` implements the VAE's
// 2-block mask (app/modules/bitwise_vae.py:67-76). The KV-cached AR schedule needs no mask.
// Warp roles: 0 = TMA producer, 1 = MMA issuer, 2 = TMEM allocator, 4..7 = softmax + epilogue.
#include <cuda.h>
#include <mutex>
#include "kernels.cuh"

namespace artalk {

namespace {

constexpr int QT = 128;             // query rows per tile
constexpr int KC = 128;             // key rows per TMA chunk
constexpr int CHUNK_BYTES = 128 * 128;   // 128 rows x 64 bf16
constexpr int O_COL = 384;          // TMEM column of the O accumulator (S/P live in [0, 384))
constexpr int MAX_LK = 384;

struct AttnTcParams {
  int n_heads, lq, lk, lk_pad, n_kchunks, q_tiles, total_items;
  int split;
  float scale_log2e;
  bf16* out; int64_t o_ss, o_rs;
  unsigned int* err_flag;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, unsigned int* err_flag, int who) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > 20000000u) {
      if (err_flag) atomicExch(err_flag, 0xA77E0000u | (uint32_t)who);
      __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int x, int y, int z) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(x), "r"(y), "r"(z) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mma_ss(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// K-major operand tile, 128B swizzle: 8-row groups of 1024 B
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
// MN-major operand tile (V: rows = keys (K dim), 64 contiguous head dims = one 128 B swizzle row):
// canonical ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units -> 8-key groups 1024 B apart (SBO); a single 64-wide N block
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)(1024 >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// D = F32, A = B = BF16, M = 128; b_mn: B operand is MN-major
__device__ __forceinline__ uint32_t idesc(int n, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)b_mn << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

__global__ void __launch_bounds__(256, 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
               const __grid_constant__ CUtensorMap tmV, const AttnTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t stage_bytes = CHUNK_BYTES * (1 + 2 * p.n_kchunks);
  const uint32_t bar_base = smem_base + 2 * stage_bytes;
  // barriers: kv_full[2], kv_empty[2], s_full, p_full, o_full, s_empty, then the TMEM base word
  auto kv_full = [&](int b) { return bar_base + 8u * b; };
  auto kv_empty = [&](int b) { return bar_base + 8u * (2 + b); };
  const uint32_t s_full = bar_base + 32, p_full = bar_base + 40, o_full = bar_base + 48, s_empty = bar_base + 56;
  const uint32_t tmem_slot = bar_base + 64;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmQ)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmK)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmV)) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int b = 0; b < 2; ++b) { mbar_init(kv_full(b), 1); mbar_init(kv_empty(b), 1); }
    mbar_init(s_full, 1); mbar_init(p_full, 4); mbar_init(o_full, 1); mbar_init(s_empty, 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int item = blockIdx.x; item < p.total_items; item += gridDim.x, ++it) {
        const int b = it & 1;
        const int qt = item % p.q_tiles, sh = item / p.q_tiles, h = sh % p.n_heads, seq = sh / p.n_heads;
        mbar_wait(kv_empty(b), (((uint32_t)it >> 1) & 1u) ^ 1u, p.err_flag, 1);
        const uint32_t sq = smem_base + b * stage_bytes, sk = sq + CHUNK_BYTES, sv = sk + p.n_kchunks * CHUNK_BYTES;
        mbar_arrive_expect_tx(kv_full(b), stage_bytes);
        tma_load_3d(sq, &tmQ, kv_full(b), h * 64, qt * QT, seq);
        for (int c = 0; c < p.n_kchunks; ++c) {
          tma_load_3d(sk + c * CHUNK_BYTES, &tmK, kv_full(b), h * 64, c * KC, seq);
          tma_load_3d(sv + c * CHUNK_BYTES, &tmV, kv_full(b), h * 64, c * KC, seq);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      int it = 0;
      const uint32_t idesc_pv = idesc(64, 1);
      for (int item = blockIdx.x; item < p.total_items; item += gridDim.x, ++it) {
        const int b = it & 1;
        const uint32_t sq = smem_base + b * stage_bytes, sk = sq + CHUNK_BYTES, sv = sk + p.n_kchunks * CHUNK_BYTES;
        mbar_wait(kv_full(b), ((uint32_t)it >> 1) & 1u, p.err_flag, 2);
        mbar_wait(s_empty, ((uint32_t)it & 1u) ^ 1u, p.err_flag, 3);      // previous item's O has been read
        tc_fence_after();
        // S = Q K^T, 128 keys per instruction group
        const uint64_t dq = desc_kmajor(sq);
        for (int c = 0; c < p.n_kchunks; ++c) {
          const int n = min(KC, p.lk_pad - c * KC);
          const uint64_t dk = desc_kmajor(sk + c * CHUNK_BYTES);
          const uint32_t id = idesc(n, 0);
#pragma unroll
          for (int k = 0; k < 4; ++k) mma_ss(tmem + (uint32_t)(c * KC), dq + (uint64_t)(2 * k), dk + (uint64_t)(2 * k), id, k ? 1u : 0u);
        }
        tc_commit(s_full);
        mbar_wait(p_full, (uint32_t)it & 1u, p.err_flag, 4);
        tc_fence_after();
        // O = P V : A = P (bf16 pairs in TMEM, 8 columns per 16 keys), B = V tile MN-major (16 keys = 2048 B)
        const uint64_t dv = desc_mnmajor(sv);
        const int ksteps = p.lk_pad >> 4;
        for (int ks = 0; ks < ksteps; ++ks)
          mma_ts(tmem + O_COL, tmem + (uint32_t)(ks * 8), dv + (uint64_t)(ks * 128), idesc_pv, ks ? 1u : 0u);
        tc_commit(o_full);
        tc_commit(kv_empty(b));
      }
    }
  } else if (warp >= 4) {
    const int q = warp - 4;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    int it = 0;
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x, ++it) {
      const int qt = item % p.q_tiles, sh = item / p.q_tiles, h = sh % p.n_heads, seq = sh / p.n_heads;
      const int qi = qt * QT + q * 32 + lane;
      const int lk_r = (p.split > 0 && qi < p.split) ? p.split : p.lk;
      const int n16 = p.lk_pad >> 4;
      mbar_wait(s_full, (uint32_t)it & 1u, p.err_flag, 5);
      tc_fence_after();
      float m = -INFINITY;
      for (int c = 0; c < n16; ++c) {
        float s[16];
        tmem_ld16(tmem + lane_addr + (uint32_t)(c * 16), s);
#pragma unroll
        for (int j = 0; j < 16; ++j) if (c * 16 + j < lk_r) m = fmaxf(m, s[j]);
      }
      const float ms = m * p.scale_log2e;
      float sum = 0.f;
      for (int c = 0; c < n16; ++c) {
        float s[16];
        tmem_ld16(tmem + lane_addr + (uint32_t)(c * 16), s);
        uint32_t pk[8];
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
          float p0 = (c * 16 + j < lk_r) ? ex2(fmaf(s[j], p.scale_log2e, -ms)) : 0.f;
          float p1 = (c * 16 + j + 1 < lk_r) ? ex2(fmaf(s[j + 1], p.scale_log2e, -ms)) : 0.f;
          sum += p0 + p1;
          __nv_bfloat162 hh = __floats2bfloat162_rn(p0, p1);
          pk[j >> 1] = *reinterpret_cast<uint32_t*>(&hh);
        }
        tmem_st8(tmem + lane_addr + (uint32_t)(c * 8), pk);     // P chunk c overwrites S columns already consumed
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      mbar_wait(o_full, (uint32_t)it & 1u, p.err_flag, 6);
      tc_fence_after();
      const float inv = 1.0f / sum;
      float o[4][16];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld16(tmem + lane_addr + (uint32_t)(O_COL + c * 16), o[c]);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_empty);                        // S/P/O may be overwritten by the next item
      if (qi < p.lq) {
        bf16* orow = p.out + (int64_t)seq * p.o_ss + (int64_t)qi * p.o_rs + h * 64;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
#pragma unroll
          for (int j = 0; j < 16; j += 8) {
            __nv_bfloat162 h0 = __floats2bfloat162_rn(o[c][j] * inv, o[c][j + 1] * inv);
            __nv_bfloat162 h1 = __floats2bfloat162_rn(o[c][j + 2] * inv, o[c][j + 3] * inv);
            __nv_bfloat162 h2 = __floats2bfloat162_rn(o[c][j + 4] * inv, o[c][j + 5] * inv);
            __nv_bfloat162 h3 = __floats2bfloat162_rn(o[c][j + 6] * inv, o[c][j + 7] * inv);
            uint4 v;
            v.x = *reinterpret_cast<uint32_t*>(&h0); v.y = *reinterpret_cast<uint32_t*>(&h1);
            v.z = *reinterpret_cast<uint32_t*>(&h2); v.w = *reinterpret_cast<uint32_t*>(&h3);
            *reinterpret_cast<uint4*>(orow + c * 16 + j) = v;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)f;
  });
  return fn;
}

int make_map(CUtensorMap* m, const void* base, uint64_t width, uint64_t rows, uint64_t seqs, uint64_t rs_bytes, uint64_t ss_bytes) {
  EncodeTiledFn enc = get_encode();
  AT_REQUIRE(enc, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[3] = {width, rows, seqs};
  cuuint64_t strides[2] = {rs_bytes, ss_bytes};
  cuuint32_t box[3] = {64, 128, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("attention: cuTensorMapEncodeTiled failed (%d): dims=(%llu,%llu,%llu) strides=(%llu,%llu)", (int)r,
                   (unsigned long long)width, (unsigned long long)rows, (unsigned long long)seqs,
                   (unsigned long long)rs_bytes, (unsigned long long)ss_bytes);
    return AT_ECUDA;
  }
  return AT_OK;
}

unsigned int* g_err_flag = nullptr;
int g_num_sms = 0;
}  // namespace

bool attention_tc_supported(const AttnArgs& a) {
  return a.dt == DT_BF16 && a.head_dim == 64 && a.lk <= MAX_LK && a.lk >= 1 && a.q_rs % 8 == 0 && a.k_rs % 8 == 0 &&
         a.v_rs % 8 == 0 && a.q_ss % 8 == 0 && a.k_ss % 8 == 0 && a.v_ss % 8 == 0 && a.o_rs % 8 == 0 && a.o_ss % 8 == 0 &&
         ((uintptr_t)a.q % 16 == 0) && ((uintptr_t)a.k % 16 == 0) && ((uintptr_t)a.v % 16 == 0) && ((uintptr_t)a.out % 16 == 0);
}

int launch_attention_tc(const AttnArgs& a, cudaStream_t st) {
  if (a.n_seq <= 0 || a.lq <= 0) return AT_OK;
  AT_REQUIRE(attention_tc_supported(a), "attention_tc: unsupported shape/stride (lk=%d head_dim=%d)", a.lk, a.head_dim);
  if (!g_err_flag) {
    AT_CUDA(cudaMalloc((void**)&g_err_flag, sizeof(unsigned int)));
    AT_CUDA(cudaMemset(g_err_flag, 0, sizeof(unsigned int)));
    int dev = 0;
    AT_CUDA(cudaGetDevice(&dev));
    AT_CUDA(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
    AT_CUDA(cudaFuncSetAttribute(attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * CHUNK_BYTES * 7 + 1024 + 256));
  }
  AttnTcParams p;
  p.n_heads = a.n_heads; p.lq = a.lq; p.lk = a.lk;
  p.lk_pad = (a.lk + 15) & ~15;
  p.n_kchunks = ceil_div(p.lk_pad, KC);
  p.q_tiles = ceil_div(a.lq, QT);
  p.total_items = a.n_seq * a.n_heads * p.q_tiles;
  p.split = a.split;
  p.scale_log2e = a.scale * 1.4426950408889634f;
  p.out = (bf16*)a.out; p.o_ss = a.o_ss; p.o_rs = a.o_rs;
  p.err_flag = g_err_flag;
  CUtensorMap tmQ, tmK, tmV;
  const uint64_t wq = (uint64_t)a.n_heads * 64;
  auto ss = [](int64_t s, int64_t rs, int rows) { return (uint64_t)(s > 0 ? s : rs * rows) * 2; };   // n_seq == 1: any stride
  AT_TRY(make_map(&tmQ, a.q, wq, (uint64_t)a.lq, (uint64_t)a.n_seq, (uint64_t)a.q_rs * 2, ss(a.q_ss, a.q_rs, a.lq)));
  AT_TRY(make_map(&tmK, a.k, wq, (uint64_t)a.lk, (uint64_t)a.n_seq, (uint64_t)a.k_rs * 2, ss(a.k_ss, a.k_rs, a.lk)));
  AT_TRY(make_map(&tmV, a.v, wq, (uint64_t)a.lk, (uint64_t)a.n_seq, (uint64_t)a.v_rs * 2, ss(a.v_ss, a.v_rs, a.lk)));
  const size_t smem = (size_t)2 * CHUNK_BYTES * (1 + 2 * p.n_kchunks) + 1024 + 256;
  const int grid = p.total_items < g_num_sms ? p.total_items : g_num_sms;
  g_trace_dims[0] = a.n_seq * a.n_heads; g_trace_dims[1] = a.lq; g_trace_dims[2] = a.lk;
  attn_tc_kernel<<<grid, 256, smem, st>>>(tmQ, tmK, tmV, p);
  AT_LAUNCH_CHECK();
  return AT_OK;
}

}  // namespace artalk
