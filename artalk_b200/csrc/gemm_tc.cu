// bf16 tensor-core GEMM for sm_100a:  out = resid + gate * act(A W^T + bias), fp32 accumulation in TMEM.
//
//  * operands arrive by TMA (cp.async.bulk.tensor.3d, 128B swizzle) into a multi-stage shared-memory ring;
//    A is described by a 3-D tensor map (K, rows-in-batch, batch) so the same kernel covers plain matrices,
//    batched row views, the overlapping-row implicit-GEMM view of the wav2vec conv layers and (tap mode) the
//    grouped positional conv, with zero fill outside a batch's rows doing the conv padding;
//  * one elected thread issues tcgen05.mma (cta_group::1, M=128, N=BN, K=16 per instruction) with the
//    accumulator in TMEM; two accumulator buffers let the epilogue of tile i overlap the MMAs of tile i+1;
//  * 8 epilogue warps read their TMEM lane quarter with tcgen05.ld (one output row per thread) and apply
//    bias / activation / gate / residual before 16-byte global stores;
//  * persistent: grid = min(tiles, #SMs), tiles are walked N-fastest so an A tile is shared through L2.
// Warp roles: 0 = TMA producer, 1 = MMA issuer, 2 = TMEM allocator, 4..11 = epilogue (lane quarter x column half).
#include <cuda.h>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <mutex>
#define ARTALK_PDL_CLASS 1
#include "kernels.cuh"

namespace artalk {

namespace {

constexpr int BM = 128, BK = 64, UMMA_K = 16;
constexpr int A_STAGE_BYTES = BM * BK * 2;        // 16 KB

struct TcParams {
  int N;
  int rpb;                 // rows per batch of the A view (tile rows never cross a batch)
  int n_batches, tiles_per_batch, n_tiles_n, groups, total_tiles;
  // tail split: the last (partial) wave's tiles [main_tiles, main_tiles + r) are cut into `tail_split` column slices of
  // `tail_bn` columns each, so the ragged wave costs tail_bn/BN of a full one (total_tiles counts the slices)
  int main_tiles, tail_split, tail_bn;
  int band_n;              // pair kernel: > 0 = N tiles per L2 band (see decode)
  int tma_out;             // pair kernel: fp32 output chunks leave by TMA store from the warp's scratch (plain [M, N] out32 only)
  int tma_resid;           // pair kernel, EPI 1: the fp32 residual tile arrives by TMA (plain [M, N] residual, N % 32 == 0)
  int num_kb;              // K blocks of 64
  int tap_mode, tap_pad;   // tap mode: k-block j reads rows shifted by (j / tap_slots - tap_pad), columns of group g (piece slot j % tap_slots)
  int tap_slots;
  int exact;               // libm-accurate activations (parity-grade mode)
  int split_slots;         // SPLIT kernels: piece slots per 64-wide K block (3 / 6); slot 0 (p0 x p0) accumulates in the main
                           // accumulator, every other slot in the correction accumulator (see gemm_tc_kernel)
  int a_group_cols;        // column offset per group in the A view (tap mode)
  int64_t c_gs; int bias_gs;
  const float* bias; int act;
  const void* gate; int gate_dt; RowMap gate_map;
  const float* resid; RowMap resid_map;
  float* out32; void* out_act; int out_act_dt; RowMap c_map;
  int vec_ok;
  unsigned int* err_flag;
  int qkv_mode, qkv_C; const float* head_scale; bf16* qbuf; bf16* kcache; bf16* vcache; RowMap kv_map; int64_t kv_layer_stride;
};

// 64 fp32 values (one head of one row) -> bf16, 8 x 16-byte stores
__device__ __forceinline__ void store_head_bf16(bf16* dst, const float (&a)[32], const float (&b)[32], float scale) {
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const float* v = half ? b : a;
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      __nv_bfloat162 h0 = __floats2bfloat162_rn(v[j] * scale, v[j + 1] * scale), h1 = __floats2bfloat162_rn(v[j + 2] * scale, v[j + 3] * scale);
      __nv_bfloat162 h2 = __floats2bfloat162_rn(v[j + 4] * scale, v[j + 5] * scale), h3 = __floats2bfloat162_rn(v[j + 6] * scale, v[j + 7] * scale);
      uint4 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
      pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
      *reinterpret_cast<uint4*>(dst + half * 32 + j) = pk;
    }
  }
}

// ---------------------------------------------------------------- PTX wrappers
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
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded wait: a protocol bug traps (launch error) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, unsigned int* err_flag, int who) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > 20000000u) {
      if (err_flag) atomicExch(err_flag, 0xDEAD0000u | (uint32_t)who);
      __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int x, int y, int z) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(x), "r"(y), "r"(z)
      : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int x, int y, int z) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(x), "r"(y), "r"(z) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// shared-memory matrix descriptor, K-major, 128B swizzle: 8-row groups of 1024 B (SBO), version 1 (cute/arch/mma_sm100_desc.hpp)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);          // start address >> 4
  d |= (uint64_t)1 << 16;                          // leading byte offset (unused for swizzled K-major) = 1
  d |= (uint64_t)(1024 >> 4) << 32;                // stride byte offset = 1024 B
  d |= (uint64_t)1 << 46;                          // descriptor version 1 (Blackwell)
  d |= (uint64_t)2 << 61;                          // SWIZZLE_128B
  return d;
}
// instruction descriptor: D = F32, A = B = BF16, both K-major, M = 128, N = BN
__host__ __device__ constexpr uint32_t make_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

template <int BN, bool SPLIT = false> struct TileCfg {
  static constexpr int B_STAGE_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int STAGES = (BN == 256) ? 4 : (BN == 128) ? 6 : 8;
  static constexpr int ACC_COLS = SPLIT ? 2 * BN : BN;          // columns per accumulator buffer (SPLIT: main | correction)
  static constexpr int TMEM_COLS = (2 * ACC_COLS < 32) ? 32 : 2 * ACC_COLS;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 1024 /*barriers*/ + 8 * 4096 /*epilogue scratch, 1024-aligned: TMA store source*/;
};

// One 32-column chunk of 32 output rows (one warp; thread = row, as tcgen05.ld delivers the accumulator):
//   out = resid + gate * act(acc + bias).
// Row-per-thread global accesses cost one L1 tag lookup per lane per instruction (32 different 128-byte lines for 16 bytes
// each): 20 000 tag cycles per 128x256 tile with an fp32 residual read-modify-write and a bf16 copy, against 4 096-8 192 cycles
// of MMA for K = 512-1024. So the fp32 residual comes in and the outputs go out through a per-warp 4 KB shared-memory
// transposition (16-byte chunks XOR-swizzled with the row: conflict-free both ways): every global instruction of the warp
// covers 4 full 128-byte lines (fp32) or 8 x 64 bytes (bf16). The loads are issued ahead of the tcgen05.ld of the accumulator.
// Shared by the 1-CTA and the CTA-pair kernels. All 32 lanes must call it (shuffles, __syncwarp).
constexpr int EPI_SCRATCH_BYTES = 8 * 4096;       // 8 epilogue warps x (32 rows x 128 B)

template <int EPI, int CORR_OFF = 0>
__device__ __forceinline__ void epi_chunk(const TcParams& p, uint32_t taddr, bool row_ok, int64_t c_off, int64_t g_off, int64_t r_off,
                                          const float* bias, int col0, bool gate_bf, float4* scr, int lane,
                                          const float4* rsm = nullptr, const CUtensorMap* tmo = nullptr, int row0 = 0, int bz = 0) {
    // CORR_OFF > 0 (SPLIT kernels): the correction accumulator sits CORR_OFF columns after the main one and is added here
    // (row0, bz): first row of the warp's 32 inside its batch, batch index (output map = (cols, rows per batch, batches))
    // tmo: fp32 output tensor map (box 32 x 32, 128-byte swizzle = the scratch layout); the warp's chunk then leaves as one
    // TMA store of its scratch (rows past M are clipped by the map) instead of 8 read-back + st.global rounds per thread
    // rsm: this thread's residual row of the chunk in shared memory (128 B, TMA 128-byte swizzle), or null
    const bool live = row_ok && col0 < p.N;
    const bool full = p.vec_ok && (col0 + 32 <= p.N);            // warp-uniform
    const int sw = lane & 7;
    // operand loads of this chunk go out before the accumulator read (independent of it)
    uint4 gpre[4];
    float4 rpre[8];
    const bool pre_g = EPI == 1 && live && full && gate_bf;
    const bool use_r = EPI == 1 && full && p.resid != nullptr;   // warp-uniform
    if (pre_g) {
#pragma unroll
      for (int j = 0; j < 4; ++j) gpre[j] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(p.gate) + g_off + col0) + j);
    }
    if (use_r && rsm) {
#pragma unroll
      for (int j = 0; j < 8; ++j) rpre[j] = rsm[j ^ sw];
    } else if (use_r && live) {
      // row-per-thread 16-byte loads, in flight while the accumulator is read. Two transposed, coalesced variants of this
      // load (landing in the scratch before the tcgen05.ld; requested one chunk ahead into registers) measured 25-50 %
      // slower on the residual GEMMs (out-proj 110 -> 140-180 us), unlike the stores, so the loads stay per thread
#pragma unroll
      for (int j = 0; j < 8; ++j) rpre[j] = *(reinterpret_cast<const float4*>(p.resid + r_off + col0) + j);
    }
    // plain epilogue (registers to spare): the bias row is requested ahead of the accumulator read as well
    float4 bpre[8];
    const bool pre_b = EPI == 0 && bias != nullptr && full;
    if (pre_b) {
#pragma unroll
      for (int j = 0; j < 8; ++j) bpre[j] = __ldg(reinterpret_cast<const float4*>(bias + col0) + j);
    }
    float v[32];
    tmem_ld32(taddr, v);
    if constexpr (CORR_OFF > 0) {
      float cr[32];
      tmem_ld32(taddr + (uint32_t)CORR_OFF, cr);
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] += cr[j];                // one fp32 round-to-nearest add of the small terms
    }
    if (!full && !live) return;                                  // ragged path below is per thread (no warp collectives)
    // ---- bias
    if (bias) {
      if (pre_b) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { v[4 * j] += bpre[j].x; v[4 * j + 1] += bpre[j].y; v[4 * j + 2] += bpre[j].z; v[4 * j + 3] += bpre[j].w; }
      } else if (full) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float4 bv = __ldg(reinterpret_cast<const float4*>(bias + col0 + j));
          v[j] += bv.x; v[j + 1] += bv.y; v[j + 2] += bv.z; v[j + 3] += bv.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) if (col0 + j < p.N) v[j] += __ldg(bias + col0 + j);
      }
    }
    // ---- activation (switch hoisted out of the element loop; fast-math variants: outputs are rounded to bf16)
    if (p.exact) {
      if (p.act != ACT_NONE) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = apply_act(v[j], p.act);
      }
    } else
    switch (p.act) {
      case ACT_GELU_ERF:
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = gelu_erf_fast(v[j]);
        break;
      case ACT_GELU_TANH:
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = gelu_tanh_fast(v[j]);
        break;
      case ACT_LEAKY02:
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = v[j] > 0.f ? v[j] : 0.2f * v[j];
        break;
      case ACT_SILU:
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = v[j] / (1.0f + __expf(-v[j]));
        break;
      default: break;
    }
    if (full) {
      if (pre_g) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t w[4] = {gpre[j].x, gpre[j].y, gpre[j].z, gpre[j].w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w[e]);
            v[j * 8 + e * 2] *= __low2float(h); v[j * 8 + e * 2 + 1] *= __high2float(h);
          }
        }
      } else if (EPI == 1 && p.gate && live) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float gv[4];
          if (p.gate_dt == DT_F32) load4(reinterpret_cast<const float*>(p.gate) + g_off + col0 + j, gv);
          else load4(reinterpret_cast<const bf16*>(p.gate) + g_off + col0 + j, gv);
          v[j] *= gv[0]; v[j + 1] *= gv[1]; v[j + 2] *= gv[2]; v[j + 3] *= gv[3];
        }
      }
      if (use_r && (live || rsm)) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          v[j * 4] += rpre[j].x; v[j * 4 + 1] += rpre[j].y; v[j * 4 + 2] += rpre[j].z; v[j * 4 + 3] += rpre[j].w;
        }
      }
      // ---- outputs through the transposition: own row in, row-major 16-byte pieces out
      if (tmo) {                                                   // the previous chunk's store has read the scratch
        if (lane == 0) tma_store_wait_read();
        __syncwarp();
      }
      if (tmo && p.tma_out == 2) {
        // bf16 output tile (32 rows x 64 B, 64-byte swizzle: 16-byte piece j of row r sits at j ^ ((r >> 1) & 3)): packed in
        // registers, 4 conflict-free 16-byte shared stores per thread, one TMA store per warp
        uint4* s16 = reinterpret_cast<uint4*>(scr);
        const int sw4 = (lane >> 1) & 3;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          __nv_bfloat162 h0 = __floats2bfloat162_rn(v[8 * j], v[8 * j + 1]), h1 = __floats2bfloat162_rn(v[8 * j + 2], v[8 * j + 3]);
          __nv_bfloat162 h2 = __floats2bfloat162_rn(v[8 * j + 4], v[8 * j + 5]), h3 = __floats2bfloat162_rn(v[8 * j + 6], v[8 * j + 7]);
          uint4 pk;
          pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
          pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
          s16[lane * 4 + (j ^ sw4)] = pk;
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) tma_store_3d(tmo, smem_u32(scr), col0, row0, bz);
        return;
      }
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) scr[lane * 8 + (ch ^ sw)] = make_float4(v[ch * 4], v[ch * 4 + 1], v[ch * 4 + 2], v[ch * 4 + 3]);
      if (tmo) {
        fence_proxy_async();                                       // generic-proxy writes -> visible to the TMA engine
        __syncwarp();
        if (lane == 0) tma_store_3d(tmo, smem_u32(scr), col0, row0, bz);
        return;
      }
      __syncwarp();
      const bool act_f32 = p.out_act && p.out_act_dt == DT_F32, act_bf = p.out_act && p.out_act_dt != DT_F32;
      if (p.out32 || act_f32) {
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int j = it * 4 + (lane >> 3), ch = lane & 7;
          const int64_t off_j = __shfl_sync(0xffffffffu, c_off, j);
          const bool ok_j = __shfl_sync(0xffffffffu, (int)row_ok, j) != 0;
          const float4 o = scr[j * 8 + (ch ^ (j & 7))];
          if (ok_j) {
            if (p.out32) *reinterpret_cast<float4*>(p.out32 + off_j + col0 + ch * 4) = o;
            if (act_f32) *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out_act) + off_j + col0 + ch * 4) = o;
          }
        }
      }
      if (act_bf) {
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int j = it * 8 + (lane >> 2), c2 = lane & 3;          // 4 lanes x 16 B (8 bf16) per row
          const int64_t off_j = __shfl_sync(0xffffffffu, c_off, j);
          const bool ok_j = __shfl_sync(0xffffffffu, (int)row_ok, j) != 0;
          const float4 a = scr[j * 8 + ((2 * c2) ^ (j & 7))], b = scr[j * 8 + ((2 * c2 + 1) ^ (j & 7))];
          __nv_bfloat162 h0 = __floats2bfloat162_rn(a.x, a.y), h1 = __floats2bfloat162_rn(a.z, a.w);
          __nv_bfloat162 h2 = __floats2bfloat162_rn(b.x, b.y), h3 = __floats2bfloat162_rn(b.z, b.w);
          uint4 pk;
          pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
          pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
          if (ok_j) *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(p.out_act) + off_j + col0 + c2 * 8) = pk;
        }
      }
      __syncwarp();                                              // the scratch may be overwritten by the next chunk
    } else {
      // ragged / unaligned tail: fully unrolled with compile-time indices so v[] stays in registers
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if (col0 + j < p.N) {
          float t = v[j];
          if (EPI == 1 && p.gate)
            t *= (p.gate_dt == DT_F32) ? reinterpret_cast<const float*>(p.gate)[g_off + col0 + j]
                                       : __bfloat162float(reinterpret_cast<const bf16*>(p.gate)[g_off + col0 + j]);
          if (EPI == 1 && p.resid) t += p.resid[r_off + col0 + j];
          if (p.out32) p.out32[c_off + col0 + j] = t;
          if (p.out_act) {
            if (p.out_act_dt == DT_F32) reinterpret_cast<float*>(p.out_act)[c_off + col0 + j] = t;
            else reinterpret_cast<bf16*>(p.out_act)[c_off + col0 + j] = __float2bfloat16_rn(t);
          }
        }
      }
    }
}

// Fused AR q/k/v epilogue (app/transformer.py:68-74) of one 32-row warp: a warp owns whole heads (64 columns = 2 TMEM chunks), heads
// alternate between the two warp halves; q and k heads are L2-normalised (q additionally scaled by head_scale), q -> qbuf[row],
// k / v -> the caches at kv_map(row) (+ layer stride). `trow` = TMEM address of the warp's lanes at the accumulator's column 0.
// Shared by the 1-CTA and the CTA-pair kernels. All 32 lanes must call it.
__device__ __forceinline__ void epi_qkv(const TcParams& p, uint32_t trow, int bn, int col_base, int half, bool row_ok, int r,
                                        const float* bias, float4* scr, int lane) {
  const int period = (p.qkv_mode == 1 ? 3 : 2) * p.qkv_C;
#pragma unroll 1
  for (int hd = half; hd < (bn >> 6); hd += 2) {
    float a[32], b2[32];
    tmem_ld32(trow + (uint32_t)(hd * 64), a);
    tmem_ld32(trow + (uint32_t)(hd * 64 + 32), b2);
    const int col0 = col_base + hd * 64;
    if (col0 >= p.N) continue;                                   // warp-uniform
    if (bias) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + col0 + j));
        float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + col0 + 32 + j));
        a[j] += b0.x; a[j + 1] += b0.y; a[j + 2] += b0.z; a[j + 3] += b0.w;
        b2[j] += b1.x; b2[j + 1] += b1.y; b2[j + 2] += b1.z; b2[j + 3] += b1.w;
      }
    }
    const int layer = col0 / period, o = col0 - layer * period;
    const int sec = o / p.qkv_C, hc = o - sec * p.qkv_C;
    const bool is_q = (p.qkv_mode == 1 && sec == 0);
    const bool is_k = (p.qkv_mode == 1) ? (sec == 1) : (sec == 0);
    float scale = 1.0f;
    if (is_q || is_k) {
      float ss = 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) ss = fmaf(a[j], a[j], fmaf(b2[j], b2[j], ss));
      scale = 1.0f / fmaxf(sqrtf(ss), 1e-12f);                     // F.normalize eps
      if (is_q) scale *= p.head_scale[hc >> 6];
    }
    // the head's 32 rows x 128 B leave through the warp's scratch so that every store instruction covers 4 full lines
    bf16* dst = is_q ? p.qbuf + hc : (is_k ? p.kcache : p.vcache) + (int64_t)layer * p.kv_layer_stride + hc;
    const int64_t row_off = !row_ok ? 0 : (is_q ? (int64_t)r * p.qkv_C : p.kv_map.off(r));
    uint4* scr4 = reinterpret_cast<uint4*>(scr);
    const int sw = lane & 7;
#pragma unroll
    for (int ch = 0; ch < 8; ++ch) {
      const float* src = ch < 4 ? &a[ch * 8] : &b2[(ch - 4) * 8];
      __nv_bfloat162 h0 = __floats2bfloat162_rn(src[0] * scale, src[1] * scale), h1 = __floats2bfloat162_rn(src[2] * scale, src[3] * scale);
      __nv_bfloat162 h2 = __floats2bfloat162_rn(src[4] * scale, src[5] * scale), h3 = __floats2bfloat162_rn(src[6] * scale, src[7] * scale);
      uint4 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
      pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
      scr4[lane * 8 + (ch ^ sw)] = pk;
    }
    __syncwarp();
#pragma unroll
    for (int it2 = 0; it2 < 8; ++it2) {
      const int j = it2 * 4 + (lane >> 3), ch = lane & 7;
      const int64_t off_j = __shfl_sync(0xffffffffu, row_off, j);
      const bool ok_j = __shfl_sync(0xffffffffu, (int)row_ok, j) != 0;
      const uint4 val = scr4[j * 8 + (ch ^ (j & 7))];
      if (ok_j) *reinterpret_cast<uint4*>(dst + off_j + ch * 8) = val;
    }
    __syncwarp();
  }
}

// EPI: 0 = bias/activation only, 1 = gate and/or residual operands, 2 = fused AR q/k/v epilogue
// SPLIT (parity-grade mode, operands are bf16 piece blocks, split.cu): the tensor core adds every MMA into its fp32 accumulator
// with truncation, a relative loss of ~2^-24 per instruction that grows with the chain length (measured: 1e-4 relative at
// K' = 6 x 4096). The p0 x p0 products (slot 0 of every K block) therefore get their own accumulator, whose chain is K / 16
// instructions as in a plain bf16 GEMM, and all the correction products (2^-8 and below) a second one whose truncation
// is negligible at that magnitude; the epilogue adds the two once. Two buffers x (main | correction) x BN columns: BN <= 128.
template <int BN, int EPI, bool SPLIT = false>
__global__ void __launch_bounds__(384, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
               const __grid_constant__ CUtensorMap tmWt, const __grid_constant__ CUtensorMap tmO, const TcParams p) {
  using Cfg = TileCfg<BN, SPLIT>;
  static_assert(!SPLIT || (BN <= 128 && EPI != 2), "SPLIT kernels: BN <= 128, plain or gate/residual epilogue");
  constexpr int STAGES = Cfg::STAGES;
  constexpr int ACC_COLS = Cfg::ACC_COLS;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + STAGES * Cfg::STAGE_BYTES;
  // barrier layout (8 B each): full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2], then the TMEM base word
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmW)) : "memory");
    if (p.tail_split > 1) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmWt)) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)Cfg::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  // tile -> (M tile, batch, group, first column, width); tiles past main_tiles are column slices of the ragged wave
  auto decode = [&](int tile, int& col_base, int& bn, int& mt, int& b, int& g) {
    int big = tile, sub = 0;
    bn = BN;
    if (tile >= p.main_tiles) {
      const int u = tile - p.main_tiles;
      big = p.main_tiles + u / p.tail_split;
      sub = u - (u / p.tail_split) * p.tail_split;
      bn = p.tail_bn;
    }
    const int n_idx = big % p.n_tiles_n;
    int rest = big / p.n_tiles_n;
    mt = rest % p.tiles_per_batch;
    rest /= p.tiles_per_batch;
    b = rest % p.n_batches;
    g = rest / p.n_batches;
    col_base = n_idx * BN + sub * p.tail_bn;
  };
  pdl_launch_dependents();       // the next kernel of the stream may start its own prologue

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      // weights do not depend on the previous kernel: the first ring fill of W is issued before the grid dependency
      // resolves (programmatic dependent launch), the A tiles (activations) after it
      int pre = 0;
      if (blockIdx.x < p.total_tiles) {
        int col_base, bn, mt, b, g;
        decode(blockIdx.x, col_base, bn, mt, b, g);
        pre = p.num_kb < STAGES ? p.num_kb : STAGES;
        for (int kb = 0; kb < pre; ++kb) {
          mbar_arrive_expect_tx(full_bar(kb), (uint32_t)(A_STAGE_BYTES + bn * BK * 2));
          tma_load_3d(smem_base + kb * Cfg::STAGE_BYTES + A_STAGE_BYTES, bn == BN ? &tmW : &tmWt, full_bar(kb), kb * BK, col_base, g);
        }
      }
      pdl_wait();
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        int col_base, bn, mt, b, g;
        decode(tile, col_base, bn, mt, b, g);
        for (int kb = 0; kb < p.num_kb; ++kb) {
          const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES, sb = sa + A_STAGE_BYTES;
          const bool w_done = pre > 0;
          if (w_done) --pre;
          else {
            mbar_wait(empty_bar(stage), phase ^ 1u, p.err_flag, 1);
            mbar_arrive_expect_tx(full_bar(stage), (uint32_t)(A_STAGE_BYTES + bn * BK * 2));
          }
          if (p.tap_mode) {
            const int tap = kb / p.tap_slots, slot = kb - tap * p.tap_slots;
            tma_load_3d(sa, &tmA, full_bar(stage), (g * p.a_group_cols) * p.tap_slots + slot * BK, mt * BM + tap - p.tap_pad, b);
          }
          else tma_load_3d(sa, &tmA, full_bar(stage), kb * BK, mt * BM, b);
          if (!w_done) tma_load_3d(sb, bn == BN ? &tmW : &tmWt, full_bar(stage), kb * BK, col_base, g);
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0; int it = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
        const uint32_t idesc = make_idesc(tile >= p.main_tiles ? p.tail_bn : BN);
        const int acc = it & 1;
        mbar_wait(tempty_bar(acc), (((uint32_t)it >> 1) & 1u) ^ 1u, p.err_flag, 2);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * ACC_COLS);
        int slot = 0;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase, p.err_flag, 3);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES, sb = sa + A_STAGE_BYTES;
          const uint64_t da = make_smem_desc(sa), db = make_smem_desc(sb);
          // SPLIT: slot 0 of a K block -> main accumulator, the others -> correction accumulator (first use of each overwrites)
          const bool corr = SPLIT && slot != 0;
          const uint32_t d_acc = corr ? d_tmem + (uint32_t)BN : d_tmem;
          const int first_kb = corr ? 1 : 0;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)
            tc_mma_bf16(d_acc, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb != first_kb || k != 0) ? 1u : 0u);
          if (SPLIT && ++slot == p.split_slots) slot = 0;
          tc_commit(empty_bar(stage));           // frees the smem slot once these MMAs have read it
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        tc_commit(tfull_bar(acc));               // accumulator ready for the epilogue
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: 8 warps, warp -> (TMEM lane quarter, column half) =====================
    // One output row per thread. The row's gate / residual operands are requested before the accumulator is waited for:
    // an L2 prefetch of every line the thread will touch when the tile starts (gate rows live in the 1.3 GB AdaLN table
    // and always miss), and the chunk's loads are issued ahead of the tcgen05.ld so both latencies overlap.
    const int q = (warp - 4) & 3, half = (warp - 4) >> 2;
    const bool gate_bf = EPI == 1 && p.gate && p.gate_dt == DT_BF16;
    float4* scr = reinterpret_cast<float4*>(smem_raw + (bar_base + 1024u - smem_u32(smem_raw)) + (warp - 4) * 4096);
    const CUtensorMap* tmo = (EPI != 2 && p.tma_out) ? &tmO : nullptr;
    int it = 0;
    pdl_wait();
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      int col_base, bn, mt, b, g;
      decode(tile, col_base, bn, mt, b, g);
      const int acc = it & 1;
      const int t_in_batch = mt * BM + q * 32 + lane;
      const bool row_ok = t_in_batch < p.rpb;
      const int r = b * p.rpb + t_in_batch;
      const int64_t cg = (int64_t)g * p.c_gs;
      const int64_t c_off = row_ok ? p.c_map.off(r) + cg : 0;
      const int64_t g_off = (row_ok && p.gate) ? p.gate_map.off(r) + cg : 0;
      const int64_t r_off = (row_ok && p.resid) ? p.resid_map.off(r) + cg : 0;
      const float* bias = p.bias ? p.bias + g * p.bias_gs : nullptr;
      const int n_chunks = bn >> 5;
      if (EPI == 1 && row_ok && p.vec_ok) {
        for (int c = half; c < n_chunks; c += 2) {
          const int col0 = col_base + c * 32;
          if (col0 + 32 > p.N) break;
          if (gate_bf) prefetch_l2(reinterpret_cast<const bf16*>(p.gate) + g_off + col0);
          if (p.resid) prefetch_l2(p.resid + r_off + col0);
        }
      }
      mbar_wait(tfull_bar(acc), ((uint32_t)it >> 1) & 1u, p.err_flag, 4);
      tc_fence_after();
      if constexpr (EPI == 2) {
        epi_qkv(p, tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * ACC_COLS), bn, col_base, half, row_ok, r, bias, scr, lane);
      } else {
#pragma unroll 1
      for (int c = half; c < n_chunks; c += 2)
        epi_chunk<EPI, SPLIT ? BN : 0>(p, tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * ACC_COLS + c * 32), row_ok, c_off,
                                       g_off, r_off, bias, col_base + c * 32, gate_bf, scr, lane, nullptr, tmo, mt * BM + q * 32, b);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
    }
    if (tmo && lane == 0) tma_store_wait_all();          // every bulk store of this warp has landed before the CTA exits
  }
  // ---- teardown
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------- CTA-pair kernel (cta_group::2)
// Large GEMMs: a 128x256 tile per SM needs 48 KB of operands per 64-wide k-block; at the MMA rate that is more shared
// memory traffic (TMA writes + MMA reads) than one SM sustains, which caps the 1-CTA kernel near 50-70 % tensor activity.
// Here two CTAs of a cluster (one TPC) share a 256x256 tile: each loads its own 128 rows of A and HALF of the W tile
// (32 KB per k-block per SM), the leader CTA issues tcgen05.mma.cta_group::2 (M = 256) which reads both halves of W from
// the two shared memories, and each CTA gets its 128 accumulator rows in its own TMEM for the epilogue.
// Barriers: full[s] lives in the leader (both producers arrive, both CTAs' TMA bytes complete on it); empty[s] / tmem_full
// are signalled in both CTAs by multicast commits; tmem_empty lives in the leader (16 epilogue warps arrive).
constexpr int STAGES2 = 6;                                     // EPI 1 runs 5 stages: 32 KB become two residual-chunk buffers
constexpr int RBUF_BYTES = 128 * 128;                          // 128 rows x 32 fp32
constexpr int STAGE2_BYTES = 2 * A_STAGE_BYTES;                 // per CTA: A 128x64 + W 128x64 (bf16)
constexpr int SMEM2_BYTES = STAGES2 * STAGE2_BYTES + 1024 + 1024 + 8 * 4096;     // + barriers, per-warp epilogue scratch (1024-aligned: TMA store source)

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  // default semantics (.release.cta): an explicit .release.cluster compiles to MEMBAR.ALL.GPU + error barriers, ~1 us per
  // arrive, which serialised the peer's producer once per k-block (ncu: 27 % tensor activity). Ordering of the TMEM reads
  // against the leader's next MMAs comes from tcgen05.fence::before_thread_sync, as in the 1-CTA kernel.
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_3d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t cluster_bar, int x, int y, int z) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(cluster_bar), "r"(x), "r"(y), "r"(z)
      : "memory");
}
__device__ __forceinline__ void tc_commit_2sm(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D = F32, A = B = BF16, both K-major, M = 256 (CTA pair), N = n
__host__ __device__ constexpr uint32_t make_idesc_2sm(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}

// RB (EPI 1): residual-chunk buffers per epilogue half. RB = 2 trades one operand stage for two more 16 KB buffers so that two
// residual TMA loads per half are in flight: for K <= 2048 (the wav2vec out-projection: 16 k-blocks per tile) the epilogue,
// not the MMA ring, is the critical path and each chunk exposed most of a DRAM round trip (~3 us per 32-column chunk).
// SPLIT2 = 256 (parity-grade modes, operands are bf16 piece blocks): the accumulator stage holds the main (p0 x p0) and the correction
// accumulator side by side (as gemm_tc_kernel<.., SPLIT>): 256 x 256 tiles, ONE stage of 2 x 256 columns = all of TMEM. The
// epilogue no longer overlaps the next tile's MMAs, but with K' = 3 K (>= 24 k-blocks of 512 clk) the main loop is 3-6x longer
// than the epilogue and runs at the 256-wide tile's operand reuse. (A 256 x 128 variant with two stages measured 10 % slower on
// the whole bf16x3 step: with N = 128 a k-block moves 24 KB into and 32 KB out of shared memory per 256 clk, tensor pipe 40-45 %,
// profiles/r2_ncu_new_kernels.md.)
// NEW = epilogue warps per CTA (8, or 12 for activation epilogues with a bf16-only TMA-store output: their tiles are 2 KB, so twelve
// warps share the 32 KB scratch): a 256-wide tile's eight 32-column chunks then go to three warps per lane quarter (3 / 3 / 2 chunks
// instead of 4 / 4) and three epilogue warps per scheduler hide each other's tcgen05.ld and MUFU latencies (FFN1's GELU epilogue
// held the tensor pipe at 72 %); 128 registers per thread, no spills.
template <int EPI, int RB = 1, int SPLIT2 = 0, int NEW = 8>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128 + 32 * NEW, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                const __grid_constant__ CUtensorMap tmWt, const __grid_constant__ CUtensorMap tmR,
                const __grid_constant__ CUtensorMap tmO, const TcParams p) {
  constexpr int BN = SPLIT2 ? SPLIT2 : 256;
  constexpr int ACC = SPLIT2 ? 2 * BN : BN;                      // TMEM columns per accumulator stage
  constexpr int NACC = 512 / ACC;                                // accumulator stages (1 or 2)
  constexpr int NST = EPI == 1 ? STAGES2 - RB : STAGES2;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t rbuf_base = smem_base + NST * STAGE2_BYTES;                       // EPI 1: 2 * RB x RBUF_BYTES (1024-aligned)
  const uint32_t bar_base = rbuf_base + (EPI == 1 ? 2 * RB * RBUF_BYTES : 0);
  auto rfull_bar = [&](int i) { return bar_base + 8u * (2 * STAGES2 + 5 + i); };   // buffer i = half + 2 * slot, i < 6
  auto rempty_bar = [&](int i) { return bar_base + 8u * (2 * STAGES2 + 11 + i); };
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES2 + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES2 + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES2 + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES2 + 4);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmW)) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES2; ++s) { mbar_init(full_bar(s), 2); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 2 * NEW); }
    for (int i = 0; i < 6; ++i) { mbar_init(rfull_bar(i), 1); mbar_init(rempty_bar(i), 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_launch_dependents();

  const int n_clusters = (int)(gridDim.x >> 1), cluster_id = (int)(blockIdx.x >> 1);
  // tiles past main_tiles are column slices (tail_bn wide) of the ragged last wave's tiles
  auto decode = [&](int tile, int& col_base, int& bn, int& mt, int& b) {
    int big = tile, sub = 0;
    bn = BN;
    if (tile >= p.main_tiles) {
      const int u = tile - p.main_tiles;
      big = p.main_tiles + u / p.tail_split;
      sub = u - (u / p.tail_split) * p.tail_split;
      bn = p.tail_bn;
    }
    int n_idx, rest;
    if (p.band_n > 0) {
      // W too large for L2 (the hoisted AdaLN GEMM: 116 MB): N is walked in bands of band_n tiles, all M tiles per band, so a
      // band of W (<= 32 MB) stays L2-resident while the (small) A matrix is re-read per band. Plain N-fastest order
      // streamed the whole W from HBM once per M tile: 5.3 GB per launch, 1100 TFLOP/s
      const int m_all = p.tiles_per_batch * p.n_batches, per_band = p.band_n * m_all;
      int band = big / per_band;
      const int n_bands = (p.n_tiles_n + p.band_n - 1) / p.band_n;
      if (band > n_bands - 1) band = n_bands - 1;
      const int r = big - band * per_band;
      const int width = (band == n_bands - 1) ? p.n_tiles_n - band * p.band_n : p.band_n;
      rest = r / width;
      n_idx = band * p.band_n + (r - rest * width);
    } else {
      n_idx = big % p.n_tiles_n;
      rest = big / p.n_tiles_n;
    }
    mt = rest % p.tiles_per_batch;
    b = rest / p.tiles_per_batch;
    col_base = n_idx * BN + sub * p.tail_bn;
  };

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      pdl_wait();
      for (int tile = cluster_id; tile < p.total_tiles; tile += n_clusters) {
        int col_base, bn, mt, b;
        decode(tile, col_base, bn, mt, b);
        const int w_rows = bn >> 1;                     // this CTA's half of the W tile
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u, p.err_flag, 0x21);
          const uint32_t sa = smem_base + stage * STAGE2_BYTES, sb = sa + A_STAGE_BYTES;
          const uint32_t lead_full = mapa_shared(full_bar(stage), 0);
          if (rank == 0) mbar_arrive_expect_tx(full_bar(stage), 2u * (uint32_t)(A_STAGE_BYTES + w_rows * BK * 2));
          else mbar_arrive_cluster(lead_full);
          tma_load_3d_2sm(sa, &tmA, lead_full, kb * BK, mt * 256 + (int)rank * 128, b);
          tma_load_3d_2sm(sb, bn == BN ? &tmW : &tmWt, lead_full, kb * BK, col_base + (int)rank * w_rows, 0);
          if (++stage == NST) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (lane == 0 && rank == 0) {
      int stage = 0; uint32_t phase = 0; int it = 0;
      for (int tile = cluster_id; tile < p.total_tiles; tile += n_clusters, ++it) {
        const uint32_t idesc = make_idesc_2sm(tile >= p.main_tiles ? p.tail_bn : BN);
        const int acc = it % NACC;
        mbar_wait(tempty_bar(acc), (((uint32_t)it / NACC) & 1u) ^ 1u, p.err_flag, 0x22);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * ACC);
        int slot = 0;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase, p.err_flag, 0x23);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * STAGE2_BYTES, sb = sa + A_STAGE_BYTES;
          const uint64_t da = make_smem_desc(sa), db = make_smem_desc(sb);
          // SPLIT2: slot 0 of a K block -> main accumulator, the others -> correction accumulator (first use of each overwrites)
          const bool corr = SPLIT2 && slot != 0;
          const uint32_t d_acc = corr ? d_tmem + (uint32_t)BN : d_tmem;
          const int first_kb = corr ? 1 : 0;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)
            tc_mma_bf16_2sm(d_acc, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb != first_kb || k != 0) ? 1u : 0u);
          if (SPLIT2 && ++slot == p.split_slots) slot = 0;
          tc_commit_2sm(empty_bar(stage));         // frees the slot in both CTAs once these MMAs have read it
          if (++stage == NST) { stage = 0; phase ^= 1u; }
        }
        tc_commit_2sm(tfull_bar(acc));             // accumulator ready for both CTAs' epilogues
      }
    }
  } else if (warp == 3) {
    // ===================== residual producer (EPI 1, both CTAs): one 128-row x 32-column fp32 chunk per TMA =====================
    // Chunks are requested in the order the two epilogue halves consume them (half h owns chunks h, h+2, ... and buffer h), so
    // a chunk lands while the previous one of that half is processed: no global-load latency on the epilogue's critical path.
    if (EPI == 1 && p.tma_resid && lane == 0) {
      uint32_t use[2] = {0u, 0u};
      pdl_wait();
      for (int tile = cluster_id; tile < p.total_tiles; tile += n_clusters) {
        int col_base, bn, mt, b;
        decode(tile, col_base, bn, mt, b);
        const int row0 = b * p.rpb + mt * 256 + (int)rank * 128;
        for (int c = 0; c < (bn >> 5); ++c) {
          const int h = c & 1;
          const int idx = h + 2 * (int)(use[h] % RB);
          const uint32_t round = use[h] / RB;
          mbar_wait(rempty_bar(idx), (round & 1u) ^ 1u, p.err_flag, 0x25);
          mbar_arrive_expect_tx(rfull_bar(idx), (uint32_t)RBUF_BYTES);
          tma_load_3d(rbuf_base + idx * RBUF_BYTES, &tmR, rfull_bar(idx), col_base + c * 32, row0, 0);
          ++use[h];
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (both CTAs, 128 rows each) =====================
    constexpr int NPART = NEW / 4;                       // warps per lane quarter: chunk c goes to part c % NPART
    const int q = (warp - 4) & 3, half = (warp - 4) >> 2;
    const bool gate_bf = EPI == 1 && p.gate && p.gate_dt == DT_BF16;
    float4* scr = reinterpret_cast<float4*>(smem_raw + (bar_base + 1024u - smem_u32(smem_raw)) + (warp - 4) * (NEW == 8 ? 4096 : 2048));
    const CUtensorMap* tmo = (EPI != 2 && p.tma_out) ? &tmO : nullptr;
    int it = 0;
    uint32_t r_use = 0;
    pdl_wait();
    for (int tile = cluster_id; tile < p.total_tiles; tile += n_clusters, ++it) {
      int col_base, bn, mt, b;
      decode(tile, col_base, bn, mt, b);
      const int n_chunks = bn >> 5;
      const int acc = it % NACC;
      const int t_in_batch = mt * 256 + (int)rank * 128 + q * 32 + lane;
      const bool row_ok = t_in_batch < p.rpb;
      const int r = b * p.rpb + t_in_batch;
      const int64_t c_off = row_ok ? p.c_map.off(r) : 0;
      const int64_t g_off = (row_ok && p.gate) ? p.gate_map.off(r) : 0;
      const int64_t r_off = (row_ok && p.resid) ? p.resid_map.off(r) : 0;
      if (EPI == 1 && row_ok && p.vec_ok) {
        for (int c = half; c < n_chunks; c += NPART) {
          const int col0 = col_base + c * 32;
          if (col0 + 32 > p.N) break;
          if (gate_bf) prefetch_l2(reinterpret_cast<const bf16*>(p.gate) + g_off + col0);
          if (p.resid && !p.tma_resid) prefetch_l2(p.resid + r_off + col0);
        }
      }
      mbar_wait(tfull_bar(acc), ((uint32_t)it / NACC) & 1u, p.err_flag, 0x24);
      tc_fence_after();
      if constexpr (EPI == 2) {
        epi_qkv(p, tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * ACC), bn, col_base, half, row_ok, r, p.bias, scr, lane);
      } else
#pragma unroll 1
      for (int c = half; c < n_chunks; c += NPART) {
        const float4* rsm = nullptr;
        const int ridx = half + 2 * (int)(r_use % RB);
        if (EPI == 1 && p.tma_resid) {
          mbar_wait(rfull_bar(ridx), (r_use / RB) & 1u, p.err_flag, 0x26);
          rsm = reinterpret_cast<const float4*>(smem_raw + (rbuf_base - smem_u32(smem_raw)) + ridx * RBUF_BYTES + (q * 32 + lane) * 128);
        }
        epi_chunk<EPI, SPLIT2 ? BN : 0>(p, tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * ACC + c * 32), row_ok, c_off, g_off, r_off,
                                        p.bias, col_base + c * 32, gate_bf, scr, lane, rsm, tmo, mt * 256 + (int)rank * 128 + q * 32, b);
        if (EPI == 1 && p.tma_resid) {             // the row was copied to registers at the top of epi_chunk
          __syncwarp();
          if (lane == 0) mbar_arrive(rempty_bar(ridx));
          ++r_use;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_shared(tempty_bar(acc), 0));
    }
    if (tmo && lane == 0) tma_store_wait_all();          // every bulk store of this warp has landed before the CTA exits
  }
  // ---- teardown: the peer's shared memory and the leader's barriers stay alive until both CTAs are done
  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ---------------------------------------------------------------- host side
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

int make_map_3d(CUtensorMap* m, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t s1_bytes, uint64_t s2_bytes,
                uint32_t b0, uint32_t b1, CUtensorMapDataType dtype = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
  EncodeTiledFn enc = get_encode();
  AT_REQUIRE(enc, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {s1_bytes, s2_bytes};
  cuuint32_t box[3] = {b0, b1, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(m, dtype, 3, const_cast<void*>(base), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed (%d): dims=(%llu,%llu,%llu) strides=(%llu,%llu) box=(%u,%u) base=%p", (int)r,
                   (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2, (unsigned long long)s1_bytes,
                   (unsigned long long)s2_bytes, b0, b1, base);
    return AT_ECUDA;
  }
  return AT_OK;
}

thread_local int g_num_sms = 0;          // SM count of the current device, refreshed from dev_ctx() at every launch
extern int g_tma_out;

// Output tensor map for the TMA-store epilogue: (cols, rows per batch, batches), 32 x 32 boxes, so rows a tile computes past
// its batch (or past M) are clipped. fp32-only outputs (128-byte swizzle = the scratch layout) or bf16-only outputs (64-byte
// swizzle); plain output rows, whole 32-column chunks, no groups. Returns 0 (st.global epilogue), 1 (fp32) or 2 (bf16).
int make_out_map(CUtensorMap* tmO, const GemmArgs& g, int rpb, int n_batches, bool vec_ok) {
  if (!g_tma_out || !vec_ok || g.c_map.rpb > 0 || g.N % 32 != 0 || g.groups != 1 || g.qkv_mode) return 0;
  if (g.out32 && !g.out_act && g.c_map.rs % 4 == 0 && ((uintptr_t)g.out32 % 16 == 0)) {
    if (make_map_3d(tmO, g.out32, (uint64_t)g.N, (uint64_t)rpb, (uint64_t)n_batches, (uint64_t)g.c_map.rs * 4,
                    (uint64_t)rpb * g.c_map.rs * 4, 32, 32, CU_TENSOR_MAP_DATA_TYPE_FLOAT32) != AT_OK) return -1;
    return 1;
  }
  if (g_tma_out >= 2 && !g.out32 && g.out_act && g.out_act_dt == DT_BF16 && g.c_map.rs % 8 == 0 && ((uintptr_t)g.out_act % 16 == 0)) {
    if (make_map_3d(tmO, g.out_act, (uint64_t)g.N, (uint64_t)rpb, (uint64_t)n_batches, (uint64_t)g.c_map.rs * 2,
                    (uint64_t)rpb * g.c_map.rs * 2, 32, 32, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, CU_TENSOR_MAP_SWIZZLE_64B) != AT_OK) return -1;
    return 2;
  }
  return 0;
}

template <int BN, int EPI, bool SPLIT = false>
int launch_bn_epi(const CUtensorMap& tmA, const CUtensorMap& tmW, const CUtensorMap& tmWt, const CUtensorMap& tmO, const TcParams& p,
                  cudaStream_t st) {
  using Cfg = TileCfg<BN, SPLIT>;
  AT_TRY(ensure_dyn_smem((const void*)gemm_tc_kernel<BN, EPI, SPLIT>, Cfg::SMEM_BYTES));
  int grid = p.total_tiles < g_num_sms ? p.total_tiles : g_num_sms;
  g_trace_dims[0] = p.rpb * p.n_batches; g_trace_dims[1] = p.N; g_trace_dims[2] = p.num_kb * BK * p.groups;
  AT_CUDA(launch_k(gemm_tc_kernel<BN, EPI, SPLIT>, dim3(grid), dim3(384), Cfg::SMEM_BYTES, st, tmA, tmW, tmWt, tmO, p));
  AT_LAUNCH_CHECK();
  return AT_OK;
}

template <int BN>
int launch_bn(const CUtensorMap& tmA, const CUtensorMap& tmW, const CUtensorMap& tmWt, const CUtensorMap& tmO, const TcParams& p,
              cudaStream_t st) {
  if constexpr (BN <= 128) {
    if (p.split_slots) {
      if (p.gate || p.resid) return launch_bn_epi<BN, 1, true>(tmA, tmW, tmWt, tmO, p, st);
      return launch_bn_epi<BN, 0, true>(tmA, tmW, tmWt, tmO, p, st);
    }
  }
  if constexpr (BN >= 128) {
    if (p.qkv_mode) return launch_bn_epi<BN, 2>(tmA, tmW, tmWt, tmO, p, st);
  }
  if (p.gate || p.resid) return launch_bn_epi<BN, 1>(tmA, tmW, tmWt, tmO, p, st);
  return launch_bn_epi<BN, 0>(tmA, tmW, tmWt, tmO, p, st);
}

template <int EPI, int RB = 1, int SPLIT2 = 0, int NEW = 8>
int launch_pair_epi(const CUtensorMap& tmA, const CUtensorMap& tmW, const CUtensorMap& tmWt, const CUtensorMap& tmR,
                    const CUtensorMap& tmO, const TcParams& p, cudaStream_t st) {
  static int max_clusters_dev[16] = {-1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1};
  int& max_clusters = per_device_slot(max_clusters_dev);
  if (max_clusters < 0) {
    AT_TRY(ensure_dyn_smem((const void*)gemm_tc2_kernel<EPI, RB, SPLIT2, NEW>, SMEM2_BYTES));
    cudaLaunchConfig_t qc = {};
    qc.gridDim = dim3(g_num_sms & ~1); qc.blockDim = dim3(128 + 32 * NEW); qc.dynamicSmemBytes = SMEM2_BYTES;
    cudaLaunchAttribute qa[1];
    qa[0].id = cudaLaunchAttributeClusterDimension; qa[0].val.clusterDim.x = 2; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
    qc.attrs = qa; qc.numAttrs = 1;
    int n = 0;
    AT_CUDA(cudaOccupancyMaxActiveClusters(&n, gemm_tc2_kernel<EPI, RB, SPLIT2, NEW>, &qc));
    max_clusters = n > 0 ? n : 1;
    if (getenv("ARTALK_DEBUG")) fprintf(stderr, "[artalk] gemm pair kernel EPI=%d: max active clusters %d (SMs %d)\n", EPI, n, g_num_sms);
    if (max_clusters > g_num_sms / 2) max_clusters = g_num_sms / 2;
  }
  const int clusters = p.total_tiles < max_clusters ? p.total_tiles : max_clusters;
  g_trace_dims[0] = p.rpb * p.n_batches; g_trace_dims[1] = p.N; g_trace_dims[2] = p.num_kb * BK;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * clusters); cfg.blockDim = dim3(128 + 32 * NEW); cfg.dynamicSmemBytes = SMEM2_BYTES; cfg.stream = st;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl_on() ? 2 : 1;
  AT_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc2_kernel<EPI, RB, SPLIT2, NEW>, tmA, tmW, tmWt, tmR, tmO, p));
  AT_LAUNCH_CHECK();
  return AT_OK;
}
int g_pair_mode = 1;      // 0: never use the CTA-pair kernel (developer switch, ARTALK_GEMM_PAIR=0)
int g_tma_resid = 1;      // developer switch (option "gemm_tma_resid")
int g_tma_out = 2;        // (declared above make_out_map) option "gemm_tma_out": 0 = st.global epilogue, 1 = TMA stores for fp32-only outputs, 2 = also bf16-only outputs
int g_resid_deep = 1;     // option "gemm_resid_deep": residual buffers per epilogue half beyond one: 1 -> two for K <= 2048, 2 -> also three for K <= 1024
int g_band_mb = 32;       // option "gemm_band_mb": W larger than twice this is walked in L2 bands of this size (0 = off)
int g_pair_split = 1;     // option "gemm_pair_split": parity-grade (piece-block) GEMMs take the CTA-pair kernel (0: the 1-CTA kernel)
int g_pair_min_waves10 = 18;   // option "gemm_pair_min_waves10": the pair kernel needs at least this many tenths of a wave of 256-row tiles (was 40: 256 x 30 s step 590.0 -> 579.8 ms, bit-identical)
int g_pair_qkv = 1;       // option "gemm_pair_qkv": the fused q/k/v epilogue GEMMs may take the CTA-pair kernel
int g_epi_warps = 12;     // option "gemm_epi_warps": epilogue warps of the pair kernel for plain bf16 TMA-store outputs (8 or 12)
int g_force_bn = 0;       // developer switch: force the 1-CTA kernel's N tile (option "gemm_force_bn")

}  // namespace

void set_gemm_pair_mode(int on) { g_pair_mode = on; }
void set_gemm_force_bn(int bn) { g_force_bn = bn; }
void set_gemm_tma_resid(int on) { g_tma_resid = on; }
void set_gemm_band_mb(int mb) { g_band_mb = mb; }
void set_gemm_tma_out(int on) { g_tma_out = on; }
void set_gemm_resid_deep(int on) { g_resid_deep = on; }
void set_gemm_pair_split(int on) { g_pair_split = on; }
void set_gemm_pair_min_waves10(int v) { g_pair_min_waves10 = v; }
void set_gemm_pair_qkv(int v) { g_pair_qkv = v; }
void set_gemm_epi_warps(int v) { g_epi_warps = v; }

int launch_gemm_tc(const GemmArgs& g_in, cudaStream_t st) {
  GemmArgs g = g_in;
  if (g.M <= 0 || g.N <= 0) return AT_OK;
  // an fp32 "activation" output alone is just the fp32 output (parity-grade mode keeps activations in fp32): TMA-store epilogue
  if (!g.out32 && g.out_act && g.out_act_dt == DT_F32) { g.out32 = (float*)g.out_act; g.out_act = nullptr; }
  AT_REQUIRE(g.A && g.W && (g.out32 || g.out_act || g.qkv_mode), "gemm_tc: null operand");
  AT_REQUIRE(g.K > 0 && g.K % 16 == 0, "gemm_tc: K=%d must be a positive multiple of 16", g.K);
  AT_REQUIRE(g.ldw % 8 == 0 && g.a_map.rs % 8 == 0 && g.a_map.bs % 8 == 0 && g.a_gs % 8 == 0 && g.w_gs % 8 == 0,
             "gemm_tc: operand strides must be multiples of 8 elements (16 bytes)");
  AT_REQUIRE(((uintptr_t)g.A) % 16 == 0 && ((uintptr_t)g.W) % 16 == 0, "gemm_tc: operands must be 16-byte aligned");
  AT_REQUIRE(g.tap_w == 0 || (g.tap_w == BK && g.a_map.rpb > 0 && g.K % BK == 0), "gemm_tc: tap mode needs tap_w == 64");
  AT_REQUIRE(g.groups == 1 || g.tap_w > 0, "gemm_tc: groups are only supported in tap mode");
  if (g.skinny && gemm_skinny_supported(g)) return launch_gemm_skinny(g, st);      // latency-bound shapes (few rows): skinny.cu
  const DevCtx* dc = nullptr;
  AT_TRY(dev_ctx(&dc));
  g_num_sms = dc->num_sms;
  unsigned int* const g_err_flag = dc->err_flag;
  TcParams p;
  p.tma_resid = 0; p.band_n = 0; p.tma_out = 0;
  p.N = g.N;
  const bool batched = g.a_map.rpb > 0;
  p.rpb = batched ? g.a_map.rpb : g.M;
  p.n_batches = batched ? ceil_div(g.M, g.a_map.rpb) : 1;
  AT_REQUIRE(!batched || g.M % g.a_map.rpb == 0, "gemm_tc: M must be a multiple of the A view's rows per batch");
  p.tiles_per_batch = ceil_div(p.rpb, BM);
  // CTA-pair kernel (256x256 tiles over two SMs) for the large GEMMs: >= 1.8 waves of pair tiles at >= 85 % wave efficiency
  const bool pair_split = g.split_acc != 0;
  if (g_pair_mode && (!pair_split || (g_pair_split && (g.K / BK) % g.split_acc == 0)) && !g.tap_w && g.groups == 1 && (!g.qkv_mode || (g_pair_qkv && !pair_split)) && g.N >= 256 &&
      g.N % 128 == 0) {
    const int BN2 = 256;
    const int tpb2 = ceil_div(p.rpb, 256), n_tiles_n2 = ceil_div(g.N, BN2), n_cl = g_num_sms / 2;
    const long tiles2 = (long)p.n_batches * tpb2 * n_tiles_n2;
    const double row_eff = (double)p.rpb / ((double)tpb2 * 256.0);
    // ragged last wave: its tiles are cut into column slices (>= 32 wide) that run side by side on the idle CTA pairs
    const int rem2 = (int)(tiles2 % n_cl);
    int split2 = 1;
    if (rem2 > 0) while (BN2 / (split2 * 2) >= (g.qkv_mode ? 64 : 32) && rem2 * split2 * 2 <= n_cl) split2 *= 2;     // q/k/v epilogue: whole heads
    const double waves_eff = (double)(tiles2 / n_cl) + (rem2 ? (split2 > 1 ? 1.3 / split2 : 1.0) : 0.0);
    if (10L * tiles2 >= (long)g_pair_min_waves10 * n_cl && (double)tiles2 / (waves_eff * n_cl) >= 0.85 && row_eff >= 0.85 && tiles2 < (1L << 30)) {
      p.tiles_per_batch = tpb2; p.n_tiles_n = n_tiles_n2; p.groups = 1;
      if (g_band_mb > 0 && (double)g.N * g.K * 2.0 > 2.0 * g_band_mb * 1048576.0) {
        const int bn_tiles = (int)((double)g_band_mb * 1048576.0 / ((double)BN2 * g.K * 2.0));
        p.band_n = bn_tiles < 1 ? 1 : bn_tiles;
      }
      p.main_tiles = (int)tiles2 - (split2 > 1 ? rem2 : 0); p.tail_split = split2; p.tail_bn = BN2 / split2;
      p.total_tiles = p.main_tiles + (split2 > 1 ? rem2 * split2 : 0);
      p.num_kb = ceil_div(g.K, BK);
      p.tap_mode = 0; p.tap_pad = 0; p.a_group_cols = 0; p.tap_slots = 1; p.exact = g.exact; p.split_slots = g.split_acc;
      p.c_gs = 0; p.bias_gs = 0; p.bias = g.bias; p.act = g.act;
      p.gate = g.gate; p.gate_dt = g.gate_dt; p.gate_map = g.gate_map; p.resid = g.resid; p.resid_map = g.resid_map;
      p.out32 = g.out32; p.out_act = g.out_act; p.out_act_dt = g.out_act_dt; p.c_map = g.c_map;
      p.err_flag = g_err_flag;
      p.qkv_mode = g.qkv_mode; p.qkv_C = g.qkv_C; p.head_scale = g.head_scale; p.qbuf = (bf16*)g.qbuf; p.kcache = (bf16*)g.kcache;
      p.vcache = (bf16*)g.vcache; p.kv_map = g.kv_map; p.kv_layer_stride = g.kv_layer_stride;
      if (g.qkv_mode) {
        AT_REQUIRE((g.qkv_mode == 1 || g.qkv_mode == 2) && g.qkv_C > 0 && g.qkv_C % 64 == 0 && g.N % 64 == 0 &&
                   g.N % ((g.qkv_mode == 1 ? 3 : 2) * g.qkv_C) == 0, "gemm_tc: bad fused q/k/v shape (N=%d C=%d)", g.N, g.qkv_C);
        AT_REQUIRE(g.kcache && g.vcache && (g.qkv_mode == 2 || (g.qbuf && g.head_scale)) && g.act == ACT_NONE && !g.gate && !g.resid,
                   "gemm_tc: bad fused q/k/v arguments");
        AT_REQUIRE(g.kv_map.rs % 8 == 0 && g.kv_map.bs % 8 == 0 && g.kv_layer_stride % 8 == 0 && ((uintptr_t)g.kcache % 16 == 0) &&
                   ((uintptr_t)g.vcache % 16 == 0) && (!g.qbuf || (uintptr_t)g.qbuf % 16 == 0) &&
                   (!g.bias || ((uintptr_t)g.bias % 16 == 0)), "gemm_tc: fused q/k/v outputs must be 16-byte aligned");
      }
      bool v = (g.c_map.rs % 8 == 0) && (g.c_map.bs % 8 == 0);
      if (g.bias) v = v && (((uintptr_t)g.bias) % 16 == 0);
      if (g.gate) v = v && (g.gate_map.rs % 8 == 0) && (g.gate_map.bs % 8 == 0) && (((uintptr_t)g.gate) % 16 == 0);
      if (g.resid) v = v && (g.resid_map.rs % 4 == 0) && (g.resid_map.bs % 4 == 0) && (((uintptr_t)g.resid) % 16 == 0);
      if (g.out32) v = v && (((uintptr_t)g.out32) % 16 == 0);
      if (g.out_act) v = v && (((uintptr_t)g.out_act) % 16 == 0);
      p.vec_ok = v ? 1 : 0;
      CUtensorMap tmA2, tmW2;
      const uint64_t a_s1 = (uint64_t)g.a_map.rs * 2;
      const uint64_t a_s2 = batched ? (uint64_t)g.a_map.bs * 2 : (uint64_t)p.rpb * g.a_map.rs * 2;
      AT_TRY(make_map_3d(&tmA2, g.A, (uint64_t)g.K, (uint64_t)p.rpb, (uint64_t)p.n_batches, a_s1, a_s2 ? a_s2 : 16, BK, 128));
      AT_TRY(make_map_3d(&tmW2, g.W, (uint64_t)g.K, (uint64_t)g.N, 1, (uint64_t)g.ldw * 2, (uint64_t)g.N * g.ldw * 2, BK, (uint32_t)(BN2 / 2)));
      CUtensorMap tmW2t = tmW2;
      if (p.tail_split > 1)
        AT_TRY(make_map_3d(&tmW2t, g.W, (uint64_t)g.K, (uint64_t)g.N, 1, (uint64_t)g.ldw * 2, (uint64_t)g.N * g.ldw * 2, BK,
                           (uint32_t)(p.tail_bn / 2)));
      // fp32 residual by TMA: plain contiguous-row residual, 16-byte aligned rows, whole 32-column chunks
      CUtensorMap tmR = tmW2;
      p.tma_resid = (g_tma_resid && !pair_split && g.resid && !g.gate && p.vec_ok && g.resid_map.rpb <= 0 && g.N % 32 == 0 && g.resid_map.rs % 4 == 0 &&
                     ((uintptr_t)g.resid % 16 == 0)) ? 1 : 0;
      if (p.tma_resid)
        AT_TRY(make_map_3d(&tmR, g.resid, (uint64_t)g.N, (uint64_t)g.M, 1, (uint64_t)g.resid_map.rs * 4, (uint64_t)g.M * g.resid_map.rs * 4,
                           32, 128, CU_TENSOR_MAP_DATA_TYPE_FLOAT32));
      CUtensorMap tmO = tmW2;
      {
        const int t = make_out_map(&tmO, g, p.rpb, p.n_batches, p.vec_ok != 0);
        if (t < 0) return AT_ECUDA;
        p.tma_out = t;
      }
      if (g.qkv_mode) return launch_pair_epi<2>(tmA2, tmW2, tmW2t, tmR, tmO, p, st);
      if (pair_split) {
        if (p.gate || p.resid) return launch_pair_epi<1, 1, 256>(tmA2, tmW2, tmW2t, tmR, tmO, p, st);
        return launch_pair_epi<0, 1, 256>(tmA2, tmW2, tmW2t, tmR, tmO, p, st);
      }
      if (p.tma_resid && g_resid_deep >= 2 && p.num_kb <= 16) return launch_pair_epi<1, 3>(tmA2, tmW2, tmW2t, tmR, tmO, p, st);
      if (p.tma_resid && g_resid_deep >= 1 && p.num_kb <= 32) return launch_pair_epi<1, 2>(tmA2, tmW2, tmW2t, tmR, tmO, p, st);
      if (p.gate || p.resid) return launch_pair_epi<1>(tmA2, tmW2, tmW2t, tmR, tmO, p, st);
      // activation epilogues with a bf16-only TMA-store output (FFN1 + GELU of wav2vec / AR / VAE): twelve epilogue warps. Op-level
      // (tools_opbench.py): FFN1 + GELU 250-255 -> 241-243 us; without an activation (QKV: equal, hoisted AdaLN GEMM: 5 % slower) eight
      if (g_epi_warps == 12 && p.tma_out == 2 && p.act != ACT_NONE) return launch_pair_epi<0, 1, 0, 12>(tmA2, tmW2, tmW2t, tmR, tmO, p, st);
      return launch_pair_epi<0>(tmA2, tmW2, tmW2t, tmR, tmO, p, st);
    }
  }
  // N-tile choice: persistent CTAs on 148 SMs quantise badly for the recurrence's GEMMs (e.g. 50 x 3 tiles of 128x256 =
  // 1.01 waves). Cost model per candidate BN: full waves x time of a BN-wide tile + the ragged wave, whose tiles are cut
  // into column slices (down to 32 columns, 64 for the fused q/k/v epilogue) when that lets them run side by side.
  int BN = 32, tail_split = 1;
  {
    const int cand[4] = {256, 128, 64, 32};
    const double tile_time[4] = {256.0, 139.0, 91.0, 71.0};      // BN / efficiency of a BN-wide tile (1, .92, .70, .45)
    auto time_of = [&](int bn) { for (int i = 0; i < 4; ++i) if (cand[i] == bn) return tile_time[i]; return 71.0; };
    const int min_slice = g.qkv_mode ? 64 : 32;
    double best = 1e30;
    const long m_tiles = (long)g.groups * p.n_batches * p.tiles_per_batch;
    for (int i = 0; i < 4; ++i) {
      if (g.qkv_mode && cand[i] < 128) continue;                 // fused q/k/v epilogue needs whole heads per warp
      if (g.split_acc && cand[i] > 128) continue;               // SPLIT kernels hold main + correction accumulators: BN <= 128
      if (cand[i] > 32 && cand[i] / 2 >= g.N) continue;          // tile mostly padding
      if (g_force_bn && cand[i] != g_force_bn && !(g.qkv_mode && g_force_bn < 128)) continue;
      const long tiles = m_tiles * ceil_div(g.N, cand[i]);
      const long waves = tiles / g_num_sms, rem = tiles % g_num_sms;
      double cost = (double)waves * tile_time[i];
      int split = 1;
      if (rem > 0) {
        if (waves > 0) while (cand[i] / (split * 2) >= min_slice && rem * split * 2 <= g_num_sms) split *= 2;
        cost += time_of(cand[i] / split);
      }
      cost += 12.0;                                              // fixed per-launch latency (pipeline fill, epilogue drain)
      if (cost < best) { best = cost; BN = cand[i]; tail_split = split; }
    }
  }
  p.n_tiles_n = ceil_div(g.N, BN);
  p.groups = g.groups;
  {
    const int tiles = p.groups * p.n_batches * p.tiles_per_batch * p.n_tiles_n;
    const int rem = tail_split > 1 ? tiles % g_num_sms : 0;
    p.main_tiles = tiles - rem;
    p.tail_split = tail_split;
    p.tail_bn = BN / tail_split;
    p.total_tiles = p.main_tiles + rem * tail_split;
  }
  p.num_kb = ceil_div(g.K, BK);
  p.tap_mode = g.tap_w > 0 ? 1 : 0; p.tap_pad = g.tap_pad; p.a_group_cols = (int)g.a_gs; p.tap_slots = g.tap_slots > 0 ? g.tap_slots : 1;
  p.exact = g.exact;
  p.split_slots = g.split_acc;
  AT_REQUIRE(!g.split_acc || ((g.split_acc == 3 || g.split_acc == 6) && !g.qkv_mode && (g.K / BK) % g.split_acc == 0),
             "gemm_tc: split accumulation needs K = slots x 64 x n and a plain / gated epilogue");
  p.c_gs = g.c_gs; p.bias_gs = g.bias_gs; p.bias = g.bias; p.act = g.act;
  p.gate = g.gate; p.gate_dt = g.gate_dt; p.gate_map = g.gate_map; p.resid = g.resid; p.resid_map = g.resid_map;
  p.out32 = g.out32; p.out_act = g.out_act; p.out_act_dt = g.out_act_dt; p.c_map = g.c_map;
  p.err_flag = g_err_flag;
  p.qkv_mode = g.qkv_mode; p.qkv_C = g.qkv_C; p.head_scale = g.head_scale; p.qbuf = (bf16*)g.qbuf; p.kcache = (bf16*)g.kcache;
  p.vcache = (bf16*)g.vcache; p.kv_map = g.kv_map; p.kv_layer_stride = g.kv_layer_stride;
  if (g.qkv_mode) {
    AT_REQUIRE((g.qkv_mode == 1 || g.qkv_mode == 2) && g.qkv_C > 0 && g.qkv_C % 64 == 0 && g.N % 64 == 0 &&
               g.N % ((g.qkv_mode == 1 ? 3 : 2) * g.qkv_C) == 0, "gemm_tc: bad fused q/k/v shape (N=%d C=%d)", g.N, g.qkv_C);
    AT_REQUIRE(g.kcache && g.vcache && (g.qkv_mode == 2 || (g.qbuf && g.head_scale)) && g.act == ACT_NONE && !g.gate && !g.resid,
               "gemm_tc: bad fused q/k/v arguments");
    AT_REQUIRE(g.kv_map.rs % 8 == 0 && g.kv_map.bs % 8 == 0 && g.kv_layer_stride % 8 == 0 && ((uintptr_t)g.kcache % 16 == 0) &&
               ((uintptr_t)g.vcache % 16 == 0) && (!g.qbuf || (uintptr_t)g.qbuf % 16 == 0) &&
               (!g.bias || ((uintptr_t)g.bias % 16 == 0)), "gemm_tc: fused q/k/v outputs must be 16-byte aligned");
  }
  bool v = (g.c_map.rs % 8 == 0) && (g.c_map.bs % 8 == 0) && (g.c_gs % 8 == 0);
  if (g.bias) v = v && (((uintptr_t)g.bias) % 16 == 0) && (g.bias_gs % 4 == 0);
  if (g.gate) v = v && (g.gate_map.rs % 8 == 0) && (g.gate_map.bs % 8 == 0) && (((uintptr_t)g.gate) % 16 == 0);
  if (g.resid) v = v && (g.resid_map.rs % 4 == 0) && (g.resid_map.bs % 4 == 0) && (((uintptr_t)g.resid) % 16 == 0);
  if (g.out32) v = v && (((uintptr_t)g.out32) % 16 == 0);
  if (g.out_act) v = v && (((uintptr_t)g.out_act) % 16 == 0);
  p.vec_ok = v ? 1 : 0;

  CUtensorMap tmA, tmW;
  // A view: (K or full row width in tap mode, rows per batch, batches)
  const uint64_t a_d0 = p.tap_mode ? (uint64_t)g.a_map.rs : (uint64_t)g.K;
  const uint64_t a_s1 = (uint64_t)g.a_map.rs * 2;
  const uint64_t a_s2 = batched ? (uint64_t)g.a_map.bs * 2 : (uint64_t)p.rpb * g.a_map.rs * 2;
  AT_TRY(make_map_3d(&tmA, g.A, a_d0, (uint64_t)p.rpb, (uint64_t)p.n_batches, a_s1, a_s2 ? a_s2 : 16, BK, BM));
  const uint64_t w_s2 = g.groups > 1 ? (uint64_t)g.w_gs * 2 : (uint64_t)g.N * g.ldw * 2;
  AT_TRY(make_map_3d(&tmW, g.W, (uint64_t)g.K, (uint64_t)g.N, (uint64_t)g.groups, (uint64_t)g.ldw * 2, w_s2, BK, (uint32_t)BN));
  CUtensorMap tmWt = tmW;
  if (p.tail_split > 1)
    AT_TRY(make_map_3d(&tmWt, g.W, (uint64_t)g.K, (uint64_t)g.N, (uint64_t)g.groups, (uint64_t)g.ldw * 2, w_s2, BK, (uint32_t)p.tail_bn));
  CUtensorMap tmO = tmW;
  if (!p.tap_mode) {
    const int t = make_out_map(&tmO, g, p.rpb, p.n_batches, p.vec_ok != 0);
    if (t < 0) return AT_ECUDA;
    p.tma_out = t;
  }
  switch (BN) {
    case 256: return launch_bn<256>(tmA, tmW, tmWt, tmO, p, st);
    case 128: return launch_bn<128>(tmA, tmW, tmWt, tmO, p, st);
    case 64: return launch_bn<64>(tmA, tmW, tmWt, tmO, p, st);
    default: return launch_bn<32>(tmA, tmW, tmWt, tmO, p, st);
  }
}

}  // namespace artalk
