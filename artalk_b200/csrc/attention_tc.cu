// Tensor-core attention for the path's short sequences (<= 384 keys, head_dim 64, bf16) on sm_100a.
//
// One work item = (sequence, head, 128-row query tile); the whole key range fits on chip, so softmax is single pass:
//   TMA      : Q tile [128 x 64], K [lk_pad x 64], V [lk_pad x 64] -> shared memory (128B swizzle), 2-4 stage ring
//   MMA 1    : S = Q K^T        tcgen05.mma, A/B from shared memory (both K-major), accumulator S in TMEM
//   softmax  : one query row per thread: tcgen05.ld S -> max, exp2, row sum -> P (bf16) written back with tcgen05.st
//              over the S columns it has already consumed
//   MMA 2    : O = P V          tcgen05.mma with A = P from TMEM, B = V from shared memory consumed MN-major
//   epilogue : tcgen05.ld O, scale by 1/rowsum, bf16 stores
// TMEM holds two S/P/O buffers (256 columns each: S in [0, n), P in [0, n/2), O in [128, 192)) and there are two softmax
// warpgroups, used in one of two ways (the per-item chain MMA1 -> softmax -> MMA2 -> epilogue is latency bound, so the
// point is to keep two chains in flight):
//   * ping-pong (lk_pad <= 256): warpgroup w owns the items with local index = w (mod 2) and buffer w; the MMA thread issues
//     MMA1 of item i+1 before MMA2 of item i, so one warpgroup's softmax overlaps the other's MMAs / epilogue;
//   * split keys (lk_pad > 256, the AR steps of the two finest scales): both warpgroups work on the same item, warpgroup w on
//     the key half held in buffer w; row maxima and sums are exchanged through shared memory, MMA2 accumulates both halves
//     into one O, and each warpgroup stores half of the 64 output columns.
// Persistent CTAs (grid = min(items, #SMs)). Masking: keys >= lk never contribute (TMA zero-fills them and P is forced to
// 0); `split` implements the VAE's 2-block mask (app/modules/bitwise_vae.py:67-76). The KV-cached AR schedule needs no mask.
// Warp roles: 0 = TMA producer, 1 = MMA issuer, 2 = TMEM allocator, 4..7 / 8..11 = softmax + epilogue warpgroups 0 / 1.
#define ARTALK_PDL_CLASS 2
#include "kernels.cuh"
#include "tc_ptx.cuh"

namespace artalk {

namespace {

using namespace ptx;

constexpr int QT = 128;             // query rows per tile
constexpr int Q_BYTES = 128 * 128;  // 128 rows x 64 bf16
constexpr int BUF_COLS = 256;       // TMEM columns per S/P/O buffer
constexpr int O_COL = 128;          // O accumulator inside a buffer (S columns there are consumed before MMA 2 runs)
constexpr int MAX_LK = 384;
constexpr int MAX_STAGES = 4;
constexpr int XCH_BYTES = 2 * 2 * 2 * 128 * 4;     // [max|sum][item parity][warpgroup][row]
constexpr int SMEM_LIMIT = 232448;

struct AttnTcParams {
  int n_heads, lq, lk, lk_pad, q_tiles, total_items;
  int split;                 // VAE mask: rows < split see keys < split
  int split_keys;            // 1: both warpgroups share an item (key halves), 0: ping-pong over items
  int h0;                    // keys held in buffer 0 (= lk_pad when ping-pong)
  int n_stages, stage_bytes, kv_bytes;
  float scale_log2e;
  const float* bound;        // per-head bound of the scaled scores (null: row maxima from a first pass over S)
  int lo_z;                  // SPLIT kernels: the lo planes of Q / K / V are the sequences [lo_z, 2 lo_z) of the same tensor maps
  bf16* out; float* out32; int64_t o_ss, o_rs;
  unsigned int* err_flag;
};

__device__ __forceinline__ void mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
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
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// SPLIT (parity-grade mode "bf16x3", fp32 data flow): Q, K, V arrive as two bf16 pieces each (x = hi + lo, planes of one tensor)
// and both contractions keep the three piece products hi.hi + lo.hi + hi.lo (relative error ~2^-16 per product, as the split
// GEMMs): S = Qh Kh^T + Ql Kh^T + Qh Kl^T in one accumulator; P is split in registers and written IN PLACE over the 16 S columns
// just read (hi pieces in the first 8 columns, lo pieces in the last 8), so the three MMA-2 products take their A operands from
// TMEM without extra columns; O = Ph Vh + Pl Vh + Ph Vl lives outside the S range (columns 192..255 of buffer 0) and leaves as
// fp32. One stage of operands fills shared memory, so SPLIT launches always run in split-key mode (one item per CTA at a time).
template <bool SPLIT = false>
__global__ void __launch_bounds__(384, 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
               const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmK2,
               const __grid_constant__ CUtensorMap tmV2, const AttnTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int NS = p.n_stages;
  const uint32_t bar_base = smem_base + NS * p.stage_bytes;
  // barriers (8 B each): kv_full[4], kv_empty[4], s_full[2], p_full[2], o_full[2], s_empty[2], then the TMEM base word
  auto kv_full = [&](int s) { return bar_base + 8u * s; };
  auto kv_empty = [&](int s) { return bar_base + 8u * (MAX_STAGES + s); };
  auto s_full = [&](int b) { return bar_base + 8u * (2 * MAX_STAGES + b); };
  auto p_full = [&](int b) { return bar_base + 8u * (2 * MAX_STAGES + 2 + b); };
  auto o_full = [&](int b) { return bar_base + 8u * (2 * MAX_STAGES + 4 + b); };
  auto s_empty = [&](int b) { return bar_base + 8u * (2 * MAX_STAGES + 6 + b); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * MAX_STAGES + 8);
  const uint32_t xch_off = (bar_base + 256u) - smem_u32(smem_raw);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  volatile float* xch = reinterpret_cast<volatile float*>(smem_raw + xch_off);       // [kind][parity][wg][128]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int SK = p.split_keys;
  constexpr int OC = SPLIT ? 192 : O_COL;           // O accumulator inside buffer 0 / the item's buffer
  // operand tiles inside a stage: Q [| Q lo] | K [| K lo] | V [| V lo]
  const uint32_t QB = SPLIT ? 2u * Q_BYTES : (uint32_t)Q_BYTES, KB = (SPLIT ? 2u : 1u) * (uint32_t)p.kv_bytes;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmQ); prefetch_tensormap(&tmK); prefetch_tensormap(&tmV);
    if (SK) { prefetch_tensormap(&tmK2); prefetch_tensormap(&tmV2); }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < MAX_STAGES; ++s) { mbar_init(kv_full(s), 1); mbar_init(kv_empty(s), 1); }
    const uint32_t wg_count = SK ? 8 : 4;
    for (int b = 0; b < 2; ++b) { mbar_init(s_full(b), 1); mbar_init(p_full(b), wg_count); mbar_init(o_full(b), 1); mbar_init(s_empty(b), wg_count); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;
  pdl_launch_dependents();

  const int n_local = (p.total_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int h1 = p.lk_pad - p.h0;          // keys in buffer 1 (split-keys mode)

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      pdl_wait();
      for (int it = 0; it < n_local; ++it) {
        const int item = blockIdx.x + it * gridDim.x;
        const int st = it % NS;
        const int qt = item % p.q_tiles, sh = item / p.q_tiles, h = sh % p.n_heads, seq = sh / p.n_heads;
        mbar_wait(kv_empty(st), (((uint32_t)(it / NS)) & 1u) ^ 1u, p.err_flag, 0xA77E0001u);
        const uint32_t sq = smem_base + st * p.stage_bytes, sk = sq + QB, sv = sk + KB;
        mbar_arrive_expect_tx(kv_full(st), (SPLIT ? 2u : 1u) * (uint32_t)(Q_BYTES + 2 * p.lk_pad * 128));
#pragma unroll
        for (int pl = 0; pl < (SPLIT ? 2 : 1); ++pl) {                   // piece planes: hi, lo
          const int z = seq + pl * p.lo_z;
          tma_load_3d(sq + pl * Q_BYTES, &tmQ, kv_full(st), h * 64, qt * QT, z);
          tma_load_3d(sk + pl * p.kv_bytes, &tmK, kv_full(st), h * 64, 0, z);
          tma_load_3d(sv + pl * p.kv_bytes, &tmV, kv_full(st), h * 64, 0, z);
          if (SK) {
            tma_load_3d(sk + pl * p.kv_bytes + p.h0 * 128, &tmK2, kv_full(st), h * 64, p.h0, z);
            tma_load_3d(sv + pl * p.kv_bytes + p.h0 * 128, &tmV2, kv_full(st), h * 64, p.h0, z);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc_pv = idesc(64, 1);
      auto stage_addr = [&](int it) { return smem_base + (uint32_t)((it % NS) * p.stage_bytes); };
      // S = Q K^T for the keys [k0, k0 + n) of item `it` into buffer b
      auto qk = [&](int it, int b, int k0, int n) {
        const uint32_t sq = stage_addr(it), sk = sq + QB + (uint32_t)(k0 * 128);
        const uint64_t dq = desc_kmajor_sw128(sq), dk = desc_kmajor_sw128(sk);
        const uint32_t id = idesc(n, 0);
#pragma unroll
        for (int k = 0; k < 4; ++k) mma_ss(tmem + (uint32_t)(b * BUF_COLS), dq + (uint64_t)(2 * k), dk + (uint64_t)(2 * k), id, k ? 1u : 0u);
        if (SPLIT) {                                                     // + Q_lo K_hi^T + Q_hi K_lo^T
          const uint64_t dql = desc_kmajor_sw128(sq + Q_BYTES), dkl = desc_kmajor_sw128(sk + (uint32_t)p.kv_bytes);
#pragma unroll
          for (int k = 0; k < 4; ++k) mma_ss(tmem + (uint32_t)(b * BUF_COLS), dql + (uint64_t)(2 * k), dk + (uint64_t)(2 * k), id, 1u);
#pragma unroll
          for (int k = 0; k < 4; ++k) mma_ss(tmem + (uint32_t)(b * BUF_COLS), dq + (uint64_t)(2 * k), dkl + (uint64_t)(2 * k), id, 1u);
        }
      };
      // O(buffer bo) (+)= P(buffer bp)[keys k0 .. k0+n) V
      auto pv = [&](int it, int bo, int bp, int k0, int n, bool first) {
        const uint32_t sv = stage_addr(it) + QB + KB;
        const uint64_t dv = desc_mnmajor(sv);
        if (!SPLIT) {
          for (int ks = 0; ks < (n >> 4); ++ks)
            mma_ts(tmem + (uint32_t)(bo * BUF_COLS + OC), tmem + (uint32_t)(bp * BUF_COLS + ks * 8),
                   dv + (uint64_t)(((k0 >> 4) + ks) * 128), idesc_pv, (first && ks == 0) ? 0u : 1u);
        } else {
          // P of 16 keys sits in the 16 S columns they came from: hi pieces in columns [16 ks, +8), lo pieces in [16 ks + 8, +8)
          const uint64_t dvl = desc_mnmajor(sv + (uint32_t)p.kv_bytes);
          for (int ks = 0; ks < (n >> 4); ++ks) {
            const uint32_t d = tmem + (uint32_t)(bo * BUF_COLS + OC), ah = tmem + (uint32_t)(bp * BUF_COLS + ks * 16);
            const uint64_t off = (uint64_t)(((k0 >> 4) + ks) * 128);
            mma_ts(d, ah, dv + off, idesc_pv, (first && ks == 0) ? 0u : 1u);
            mma_ts(d, ah + 8u, dv + off, idesc_pv, 1u);
            mma_ts(d, ah, dvl + off, idesc_pv, 1u);
          }
        }
      };
      if (!SK) {
        auto mma1 = [&](int it) {
          const int b = it & 1; const uint32_t j = (uint32_t)(it >> 1);
          mbar_wait(kv_full(it % NS), ((uint32_t)(it / NS)) & 1u, p.err_flag, 0xA77E0002u);
          mbar_wait(s_empty(b), (j & 1u) ^ 1u, p.err_flag, 0xA77E0003u);          // O of item it-2 has been read
          tc_fence_after();
          qk(it, b, 0, p.lk_pad);
          tc_commit(s_full(b));
        };
        if (n_local > 0) mma1(0);
        for (int it = 0; it < n_local; ++it) {
          if (it + 1 < n_local) mma1(it + 1);
          const int b = it & 1; const uint32_t j = (uint32_t)(it >> 1);
          mbar_wait(p_full(b), j & 1u, p.err_flag, 0xA77E0004u);
          tc_fence_after();
          pv(it, b, b, 0, p.lk_pad, true);
          tc_commit(o_full(b));
          tc_commit(kv_empty(it % NS));
        }
      } else {
        for (int it = 0; it < n_local; ++it) {
          mbar_wait(kv_full(it % NS), ((uint32_t)(it / NS)) & 1u, p.err_flag, 0xA77E0002u);
          mbar_wait(s_empty(0), ((uint32_t)it & 1u) ^ 1u, p.err_flag, 0xA77E0003u);
          tc_fence_after();
          qk(it, 0, 0, p.h0);
          qk(it, 1, p.h0, h1);
          tc_commit(s_full(0));
          mbar_wait(p_full(0), (uint32_t)it & 1u, p.err_flag, 0xA77E0004u);
          tc_fence_after();
          pv(it, 0, 0, 0, p.h0, true);
          pv(it, 0, 1, p.h0, h1, false);
          tc_commit(o_full(0));
          tc_commit(kv_empty(it % NS));
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== softmax + epilogue warpgroups =====================
    const int wg = (warp - 4) >> 2, q = (warp - 4) & 3;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const int row = q * 32 + lane;
    pdl_wait();
    for (int it = SK ? 0 : wg; it < n_local; it += SK ? 1 : 2) {
      const int item = blockIdx.x + it * gridDim.x;
      const int qt = item % p.q_tiles, sh = item / p.q_tiles, h = sh % p.n_heads, seq = sh / p.n_heads;
      const int qi = qt * QT + row;
      const int b = SK ? wg : (it & 1);                    // TMEM buffer this warpgroup reads S from / writes P to
      const int bar = SK ? 0 : b;
      const uint32_t par = SK ? ((uint32_t)it & 1u) : ((uint32_t)(it >> 1) & 1u);
      const int k0 = (SK && wg) ? p.h0 : 0;                // first key of this warpgroup's range
      const int n_keys = SK ? (wg ? h1 : p.h0) : p.lk_pad; // S columns of the buffer
      int lk_r = (p.split > 0 && qi < p.split) ? p.split : p.lk;
      lk_r = min(max(lk_r - k0, 0), n_keys);               // valid keys of this row within the buffer
      const bool warp_live = (qt * QT + q * 32) < p.lq;    // warp-uniform: some row of this warp is a real query
      // warp-uniform bound of the columns worth reading (rows of a warp may straddle `split`)
      int lk_w = lk_r;
      if (p.split > 0) lk_w = __reduce_max_sync(0xffffffffu, lk_r);
      const int n16 = warp_live ? ((lk_w + 15) >> 4) : 0;
      const uint32_t tb = tmem + lane_addr + (uint32_t)(b * BUF_COLS);
      mbar_wait(s_full(bar), par, p.err_flag, 0xA77E0005u);
      tc_fence_after();
      // ---- pass 1: row maximum. 16 columns per step; the next step's tcgen05.ld is in flight while this one is reduced
      // (4 independent running maxima: a single fmax chain would cost 4 cycles per column)
      // (skipped when the caller bounds the scores: one pass over S instead of two)
      // A head takes the bound only if exp2(-2 * bound * log2 e) stays a normal number (bound <= 32): warp- and item-uniform
      const float bnd = p.bound ? __ldg(p.bound + h) : 0.f;
      const bool bounded = bnd > 0.f && bnd <= 32.0f;
      float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
      if (!bounded) {
        uint32_t sa[16], sb[16];
        auto reduce = [&](const uint32_t (&s)[16], int c) {
          const int base = c * 16;
          if (base + 16 <= lk_r) {
#pragma unroll
            for (int j = 0; j < 16; ++j) m4[j & 3] = fmaxf(m4[j & 3], __uint_as_float(s[j]));
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) if (base + j < lk_r) m4[j & 3] = fmaxf(m4[j & 3], __uint_as_float(s[j]));
          }
        };
        if (n16 > 0) tmem_ld16_nowait(tb, sa);
        for (int c = 0; c < n16; c += 2) {
          tmem_ld_wait();
          if (c + 1 < n16) tmem_ld16_nowait(tb + (uint32_t)(c * 16 + 16), sb);
          reduce(sa, c);
          if (c + 1 < n16) {
            tmem_ld_wait();
            if (c + 2 < n16) tmem_ld16_nowait(tb + (uint32_t)(c * 16 + 32), sa);
            reduce(sb, c + 1);
          }
        }
      }
      float m = bounded ? bnd : fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
      if (SK && !bounded) {
        volatile float* xm = xch + ((0 * 2 + (it & 1)) * 2) * 128;
        xm[wg * 128 + row] = m;
        named_bar_sync(1, 256);
        m = fmaxf(xm[row], xm[128 + row]);
      }
      const float ms = (m == -INFINITY) ? 0.f : m * (bounded ? 1.4426950408889634f : p.scale_log2e);
      // ---- pass 2: P = exp2(S * scale - max) as bf16 pairs over the S columns already consumed; row sum.
      // Double-buffered 16-column loads; two independent partial sums
      float sum2[2] = {0.f, 0.f};
      const int n16_all = warp_live ? (n_keys >> 4) : 0;
      {
        uint32_t sa[16], sb[16];
        auto emit = [&](const uint32_t (&s0)[16], int c) {
          uint32_t pk[8], pl[8];
          const int base = c * 16;
          if (base + 16 <= lk_r) {
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
              const float x0 = fmaf(__uint_as_float(s0[j]), p.scale_log2e, -ms), x1 = fmaf(__uint_as_float(s0[j + 1]), p.scale_log2e, -ms);
              const float p0 = ex2(x0);
              const float p1 = ex2(x1);
              sum2[(j >> 1) & 1] += p0 + p1;
              __nv_bfloat162 hh = __floats2bfloat162_rn(p0, p1);
              pk[j >> 1] = *reinterpret_cast<uint32_t*>(&hh);
              if (SPLIT) {
                __nv_bfloat162 ll = __floats2bfloat162_rn(p0 - __low2float(hh), p1 - __high2float(hh));
                pl[j >> 1] = *reinterpret_cast<uint32_t*>(&ll);
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
              const float p0 = (base + j < lk_r) ? ex2(fmaf(__uint_as_float(s0[j]), p.scale_log2e, -ms)) : 0.f;
              const float p1 = (base + j + 1 < lk_r) ? ex2(fmaf(__uint_as_float(s0[j + 1]), p.scale_log2e, -ms)) : 0.f;
              sum2[(j >> 1) & 1] += p0 + p1;
              __nv_bfloat162 hh = __floats2bfloat162_rn(p0, p1);
              pk[j >> 1] = *reinterpret_cast<uint32_t*>(&hh);
              if (SPLIT) {
                __nv_bfloat162 ll = __floats2bfloat162_rn(p0 - __low2float(hh), p1 - __high2float(hh));
                pl[j >> 1] = *reinterpret_cast<uint32_t*>(&ll);
              }
            }
          }
          if (SPLIT) { tmem_st8(tb + (uint32_t)(c * 16), pk); tmem_st8(tb + (uint32_t)(c * 16 + 8), pl); }
          else tmem_st8(tb + (uint32_t)(c * 8), pk);
        };
        if (n16 > 0) tmem_ld16_nowait(tb, sa);
        for (int c = 0; c < n16; c += 2) {
          tmem_ld_wait();
          if (c + 1 < n16) tmem_ld16_nowait(tb + (uint32_t)(c * 16 + 16), sb);
          emit(sa, c);
          if (c + 1 < n16) {
            tmem_ld_wait();
            if (c + 2 < n16) tmem_ld16_nowait(tb + (uint32_t)(c * 16 + 32), sa);
            emit(sb, c + 1);
          }
        }
        const uint32_t zero[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
        for (int c = n16; c < n16_all; ++c) {                                                 // masked for every row of the warp
          if (SPLIT) { tmem_st8(tb + (uint32_t)(c * 16), zero); tmem_st8(tb + (uint32_t)(c * 16 + 8), zero); }
          else tmem_st8(tb + (uint32_t)(c * 8), zero);
        }
      }
      float sum = sum2[0] + sum2[1];
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      if (SK) xch[((1 * 2 + (it & 1)) * 2) * 128 + wg * 128 + row] = sum;
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full(bar));
      mbar_wait(o_full(bar), par, p.err_flag, 0xA77E0006u);
      tc_fence_after();
      if (SK) {
        volatile float* xs = xch + ((1 * 2 + (it & 1)) * 2) * 128;
        sum = xs[row] + xs[128 + row];
      }
      const float inv = 1.0f / sum;
      // ---- epilogue: ping-pong -> all 64 columns of the own buffer; split keys -> 32 columns of buffer 0's O
      const int c_first = SK ? wg * 32 : 0;
      const uint32_t ob = tmem + lane_addr + (uint32_t)((SK ? 0 : b) * BUF_COLS + OC + c_first);
      uint32_t o0[16], o1[16], o2[16], o3[16];
      if (warp_live) {
        tmem_ld16_nowait(ob, o0);
        tmem_ld16_nowait(ob + 16, o1);
        if (!SK) { tmem_ld16_nowait(ob + 32, o2); tmem_ld16_nowait(ob + 48, o3); }
        tmem_ld_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_empty(bar));                    // S/P/O of this buffer may be overwritten
      if (SPLIT && warp_live && qi < p.lq) {
        float* orow = p.out32 + (int64_t)seq * p.o_ss + (int64_t)qi * p.o_rs + h * 64 + c_first;
        auto store16f = [&](const uint32_t (&o)[16], int col) {
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            *reinterpret_cast<float4*>(orow + col + j) = make_float4(__uint_as_float(o[j]) * inv, __uint_as_float(o[j + 1]) * inv,
                                                                     __uint_as_float(o[j + 2]) * inv, __uint_as_float(o[j + 3]) * inv);
        };
        store16f(o0, 0);
        store16f(o1, 16);
        if (!SK) { store16f(o2, 32); store16f(o3, 48); }
      } else if (warp_live && qi < p.lq) {
        bf16* orow = p.out + (int64_t)seq * p.o_ss + (int64_t)qi * p.o_rs + h * 64 + c_first;
        auto store16 = [&](const uint32_t (&o)[16], int col) {
#pragma unroll
          for (int j = 0; j < 16; j += 8) {
            __nv_bfloat162 h0 = __floats2bfloat162_rn(__uint_as_float(o[j]) * inv, __uint_as_float(o[j + 1]) * inv);
            __nv_bfloat162 h1v = __floats2bfloat162_rn(__uint_as_float(o[j + 2]) * inv, __uint_as_float(o[j + 3]) * inv);
            __nv_bfloat162 h2 = __floats2bfloat162_rn(__uint_as_float(o[j + 4]) * inv, __uint_as_float(o[j + 5]) * inv);
            __nv_bfloat162 h3 = __floats2bfloat162_rn(__uint_as_float(o[j + 6]) * inv, __uint_as_float(o[j + 7]) * inv);
            uint4 v;
            v.x = *reinterpret_cast<uint32_t*>(&h0); v.y = *reinterpret_cast<uint32_t*>(&h1v);
            v.z = *reinterpret_cast<uint32_t*>(&h2); v.w = *reinterpret_cast<uint32_t*>(&h3);
            *reinterpret_cast<uint4*>(orow + col + j) = v;
          }
        };
        store16(o0, 0);
        store16(o1, 16);
        if (!SK) { store16(o2, 32); store16(o3, 48); }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem, 512u);
  }
}

// ---------------------------------------------------------------- block-wise kernel: 256 < keys <= 384 (the caller gives score bounds)
// The AR steps of the two finest scales see 262 / 362 keys: S no longer fits one 256-column buffer, and the split-key mode above
// puts BOTH warpgroups on one item, so the chain MMA1 -> softmax -> MMA2 -> epilogue of an item runs with nothing else in flight
// (6.5 us per item measured against 2.6 us of TMEM traffic). With bounded scores (|q.k| <= head_scale, no row maximum) the keys
// can be consumed in two independent blocks of <= 192: S_blk = Q K_blk^T -> P_blk = exp2(S_blk - bound) -> O += P_blk V_blk, the
// row sum simply accumulates (a head whose bound is too large to use falls back to an online softmax over the two blocks, with
// the usual rescaling of sum and O). So each warpgroup keeps its own item and 256-column buffer (S_blk in [0, 192), P_blk over the
// S columns already read, O in [192, 256)) as in ping-pong mode: two item chains in flight per SM.
//   * K / V arrive per block in a ring of four 48 KB stages (two per warpgroup), Q in one 16 KB slot per warpgroup, so the next
//     item's first block loads while the current item's second block is processed;
//   * one MMA issuer thread per buffer (warps 1 and 3): neither item's chain ever waits behind the other's barriers;
//   * P_blk0 is read by MMA 2 from the columns MMA 1 of block 1 overwrites: the issuer waits for MMA 2 (block 0) to complete
//     (tcgen05.commit on a private barrier) before it issues MMA 1 (block 1) - different accumulators are not ordered otherwise.
constexpr int BLK_MAX = 192;                      // keys per block
constexpr int KVB_BYTES = BLK_MAX * 128;          // K (or V) rows of one block stage
constexpr int BSTAGE_BYTES = 2 * KVB_BYTES;       // K block | V block
constexpr int BLK_O_COL = 192;
constexpr int SMEM_BLK = 4 * BSTAGE_BYTES + 2 * Q_BYTES + 1024 /*align*/ + 256 /*barriers*/;
static_assert(SMEM_BLK <= SMEM_LIMIT, "attn_blk: shared memory budget");

__global__ void __launch_bounds__(384, 1)
attn_blk_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmK2,
                const __grid_constant__ CUtensorMap tmV2, const AttnTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t q_base = smem_base + 4 * BSTAGE_BYTES;
  const uint32_t bar_base = q_base + 2 * Q_BYTES;
  // barriers (8 B each): kv_full[4], kv_empty[4], then per buffer q_full, q_empty, s_full, p_full, o_full, s_empty, pv_done
  auto kv_full = [&](int s) { return bar_base + 8u * s; };
  auto kv_empty = [&](int s) { return bar_base + 8u * (4 + s); };
  auto q_full = [&](int b) { return bar_base + 8u * (8 + b); };
  auto q_empty = [&](int b) { return bar_base + 8u * (10 + b); };
  auto s_full = [&](int b) { return bar_base + 8u * (12 + b); };
  auto p_full = [&](int b) { return bar_base + 8u * (14 + b); };
  auto o_full = [&](int b) { return bar_base + 8u * (16 + b); };
  auto s_empty = [&](int b) { return bar_base + 8u * (18 + b); };
  auto pv_done = [&](int b) { return bar_base + 8u * (20 + b); };
  const uint32_t tmem_slot = bar_base + 8u * 22;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmQ); prefetch_tensormap(&tmK); prefetch_tensormap(&tmV); prefetch_tensormap(&tmK2); prefetch_tensormap(&tmV2);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < 4; ++s) { mbar_init(kv_full(s), 1); mbar_init(kv_empty(s), 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(q_full(b), 1); mbar_init(q_empty(b), 1); mbar_init(s_full(b), 1); mbar_init(p_full(b), 4); mbar_init(o_full(b), 1);
      mbar_init(s_empty(b), 4); mbar_init(pv_done(b), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;
  pdl_launch_dependents();

  const int n_local = (p.total_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int h1 = p.lk_pad - p.h0;

  if (warp == 0 || warp == 2) {
    // ===================== TMA producer of buffer b (one thread per buffer: neither waits behind the other's stages) =====
    // Order per item: block 0 (its stage frees after MMA 2 of the previous item's block 0), Q (frees after MMA 1 of the previous
    // item's block 1), block 1: every load of the next item is requested as early as its shared memory allows
    if (lane == 0) {
      const int b = warp == 0 ? 0 : 1;
      pdl_wait();
      for (int it = b; it < n_local; it += 2) {
        const int item = blockIdx.x + it * gridDim.x;
        const int qt = item % p.q_tiles, sh = item / p.q_tiles, h = sh % p.n_heads, seq = sh / p.n_heads;
        const uint32_t par = (uint32_t)(it >> 1) & 1u;
        for (int blk = 0; blk < 2; ++blk) {
          const int st = 2 * b + blk, nk = blk ? h1 : p.h0, k0 = blk ? p.h0 : 0;
          mbar_wait(kv_empty(st), par ^ 1u, p.err_flag, 0xA77B0002u);
          mbar_arrive_expect_tx(kv_full(st), (uint32_t)(2 * nk * 128));
          const uint32_t sk = smem_base + (uint32_t)(st * BSTAGE_BYTES), sv = sk + KVB_BYTES;
          tma_load_3d(sk, blk ? &tmK2 : &tmK, kv_full(st), h * 64, k0, seq);
          tma_load_3d(sv, blk ? &tmV2 : &tmV, kv_full(st), h * 64, k0, seq);
          if (blk == 0) {
            mbar_wait(q_empty(b), par ^ 1u, p.err_flag, 0xA77B0001u);
            mbar_arrive_expect_tx(q_full(b), (uint32_t)Q_BYTES);
            tma_load_3d(q_base + (uint32_t)(b * Q_BYTES), &tmQ, q_full(b), h * 64, qt * QT, seq);
          }
        }
      }
    }
  } else if (warp == 1 || warp == 3) {
    // ===================== MMA issuer of buffer b =====================
    if (lane == 0) {
      const int b = warp == 1 ? 0 : 1;
      const uint32_t idesc_pv = idesc(64, 1);
      const uint32_t tb = tmem + (uint32_t)(b * BUF_COLS);
      const uint64_t dq = desc_kmajor_sw128(q_base + (uint32_t)(b * Q_BYTES));
      uint32_t c = 0;                                              // block uses of this buffer (s_full / p_full / pv_done phases)
      for (int it = b; it < n_local; it += 2) {
        const uint32_t par = (uint32_t)(it >> 1) & 1u;
        mbar_wait(q_full(b), par, p.err_flag, 0xA77B0003u);
        mbar_wait(s_empty(b), par ^ 1u, p.err_flag, 0xA77B0004u);      // O of this buffer's previous item has been read
        for (int blk = 0; blk < 2; ++blk, ++c) {
          const int st = 2 * b + blk, nk = blk ? h1 : p.h0;
          mbar_wait(kv_full(st), par, p.err_flag, 0xA77B0005u);
          tc_fence_after();
          const uint32_t sk = smem_base + (uint32_t)(st * BSTAGE_BYTES), sv = sk + KVB_BYTES;
          const uint64_t dk = desc_kmajor_sw128(sk), dv = desc_mnmajor(sv);
          const uint32_t id = idesc(nk, 0);
#pragma unroll
          for (int k = 0; k < 4; ++k) mma_ss(tb, dq + (uint64_t)(2 * k), dk + (uint64_t)(2 * k), id, k ? 1u : 0u);
          tc_commit(s_full(b));
          if (blk == 1) tc_commit(q_empty(b));                       // Q has been read once these MMAs complete
          mbar_wait(p_full(b), c & 1u, p.err_flag, 0xA77B0006u);
          tc_fence_after();
          for (int ks = 0; ks < (nk >> 4); ++ks)
            mma_ts(tb + (uint32_t)BLK_O_COL, tb + (uint32_t)(ks * 8), dv + (uint64_t)(ks * 128), idesc_pv, (blk == 0 && ks == 0) ? 0u : 1u);
          tc_commit(kv_empty(st));
          if (blk == 0) {
            // MMA 1 of block 1 overwrites the columns these MMAs read P from
            tc_commit(pv_done(b));
            mbar_wait(pv_done(b), (uint32_t)(it >> 1) & 1u, p.err_flag, 0xA77B0007u);
            tc_fence_after();
          }
        }
        tc_commit(o_full(b));
      }
    }
  } else if (warp >= 4) {
    // ===================== softmax + epilogue warpgroup b =====================
    const int b = (warp - 4) >> 2, q = (warp - 4) & 3;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const int row = q * 32 + lane;
    const uint32_t tb = tmem + lane_addr + (uint32_t)(b * BUF_COLS);
    uint32_t c = 0;
    pdl_wait();
    for (int it = b; it < n_local; it += 2) {
      const int item = blockIdx.x + it * gridDim.x;
      const int qt = item % p.q_tiles, sh = item / p.q_tiles, h = sh % p.n_heads, seq = sh / p.n_heads;
      const int qi = qt * QT + row;
      const uint32_t par = (uint32_t)(it >> 1) & 1u;
      const bool warp_live = (qt * QT + q * 32) < p.lq;
      // a head takes its bound as the softmax shift if exp2(-2 * bound * log2 e) stays a normal number (bound <= 32); any other
      // head (the clamp value 100, or no bound given) runs the two blocks as an online softmax: block maximum first, and the
      // second block rescales the row sum and the O accumulator of the first (item-uniform branch)
      const float bnd = p.bound ? __ldg(p.bound + h) : 0.f;
      const bool bounded = bnd > 0.f && bnd <= 32.0f;
      float ms = bnd * 1.4426950408889634f, m_run = -INFINITY;
      float sum2[2] = {0.f, 0.f};
      for (int blk = 0; blk < 2; ++blk, ++c) {
        const int nk = blk ? h1 : p.h0, k0 = blk ? p.h0 : 0;
        const int lk_r = min(max(p.lk - k0, 0), nk);
        const int n16 = warp_live ? ((lk_r + 15) >> 4) : 0, n16_all = warp_live ? (nk >> 4) : 0;
        mbar_wait(s_full(b), c & 1u, p.err_flag, 0xA77B0008u);
        tc_fence_after();
        uint32_t sa[16], sb[16];
        if (!bounded) {
          float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
          auto reduce = [&](const uint32_t (&s0)[16], int cc) {
            const int base = cc * 16;
            if (base + 16 <= lk_r) {
#pragma unroll
              for (int j = 0; j < 16; ++j) m4[j & 3] = fmaxf(m4[j & 3], __uint_as_float(s0[j]));
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) if (base + j < lk_r) m4[j & 3] = fmaxf(m4[j & 3], __uint_as_float(s0[j]));
            }
          };
          if (n16 > 0) tmem_ld16_nowait(tb, sa);
          for (int cc = 0; cc < n16; cc += 2) {
            tmem_ld_wait();
            if (cc + 1 < n16) tmem_ld16_nowait(tb + (uint32_t)(cc * 16 + 16), sb);
            reduce(sa, cc);
            if (cc + 1 < n16) {
              tmem_ld_wait();
              if (cc + 2 < n16) tmem_ld16_nowait(tb + (uint32_t)(cc * 16 + 32), sa);
              reduce(sb, cc + 1);
            }
          }
          const float m_new = fmaxf(m_run, fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])));
          if (blk == 1) {
            // MMA 2 of block 0 has completed (the issuer waited for it before MMA 1 of this block) and MMA 2 of this block waits
            // for p_full: O is ours to rescale
            const float alpha = (m_run == -INFINITY) ? 0.f : ex2((m_run - m_new) * p.scale_log2e);
            sum2[0] *= alpha; sum2[1] *= alpha;
            if (warp_live) {
#pragma unroll 1
              for (int g16 = 0; g16 < 4; ++g16) {
                uint32_t o[16];
                tmem_ld16_nowait(tb + (uint32_t)(BLK_O_COL + g16 * 16), o);
                tmem_ld_wait();
                uint32_t lo[8], hi[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) { lo[j] = __float_as_uint(__uint_as_float(o[j]) * alpha); hi[j] = __float_as_uint(__uint_as_float(o[8 + j]) * alpha); }
                tmem_st8(tb + (uint32_t)(BLK_O_COL + g16 * 16), lo);
                tmem_st8(tb + (uint32_t)(BLK_O_COL + g16 * 16 + 8), hi);
              }
            }
          }
          m_run = m_new;
          ms = (m_run == -INFINITY) ? 0.f : m_run * p.scale_log2e;
        }
        auto emit = [&](const uint32_t (&s0)[16], int cc) {
          uint32_t pk[8];
          const int base = cc * 16;
          if (base + 16 <= lk_r) {
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
              const float p0 = ex2(fmaf(__uint_as_float(s0[j]), p.scale_log2e, -ms)), p1 = ex2(fmaf(__uint_as_float(s0[j + 1]), p.scale_log2e, -ms));
              sum2[(j >> 1) & 1] += p0 + p1;
              __nv_bfloat162 hh = __floats2bfloat162_rn(p0, p1);
              pk[j >> 1] = *reinterpret_cast<uint32_t*>(&hh);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
              const float p0 = (base + j < lk_r) ? ex2(fmaf(__uint_as_float(s0[j]), p.scale_log2e, -ms)) : 0.f;
              const float p1 = (base + j + 1 < lk_r) ? ex2(fmaf(__uint_as_float(s0[j + 1]), p.scale_log2e, -ms)) : 0.f;
              sum2[(j >> 1) & 1] += p0 + p1;
              __nv_bfloat162 hh = __floats2bfloat162_rn(p0, p1);
              pk[j >> 1] = *reinterpret_cast<uint32_t*>(&hh);
            }
          }
          tmem_st8(tb + (uint32_t)(cc * 8), pk);
        };
        if (n16 > 0) tmem_ld16_nowait(tb, sa);
        for (int cc = 0; cc < n16; cc += 2) {
          tmem_ld_wait();
          if (cc + 1 < n16) tmem_ld16_nowait(tb + (uint32_t)(cc * 16 + 16), sb);
          emit(sa, cc);
          if (cc + 1 < n16) {
            tmem_ld_wait();
            if (cc + 2 < n16) tmem_ld16_nowait(tb + (uint32_t)(cc * 16 + 32), sa);
            emit(sb, cc + 1);
          }
        }
        const uint32_t zero[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
        for (int cc = n16; cc < n16_all; ++cc) tmem_st8(tb + (uint32_t)(cc * 8), zero);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full(b));
      }
      mbar_wait(o_full(b), par, p.err_flag, 0xA77B0009u);
      tc_fence_after();
      const float inv = 1.0f / (sum2[0] + sum2[1]);
      const uint32_t ob = tb + (uint32_t)BLK_O_COL;
      uint32_t o0[16], o1[16], o2[16], o3[16];
      if (warp_live) {
        tmem_ld16_nowait(ob, o0); tmem_ld16_nowait(ob + 16, o1); tmem_ld16_nowait(ob + 32, o2); tmem_ld16_nowait(ob + 48, o3);
        tmem_ld_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_empty(b));
      if (warp_live && qi < p.lq) {
        bf16* orow = p.out + (int64_t)seq * p.o_ss + (int64_t)qi * p.o_rs + h * 64;
        auto store16 = [&](const uint32_t (&o)[16], int col) {
#pragma unroll
          for (int j = 0; j < 16; j += 8) {
            __nv_bfloat162 e0 = __floats2bfloat162_rn(__uint_as_float(o[j]) * inv, __uint_as_float(o[j + 1]) * inv);
            __nv_bfloat162 e1 = __floats2bfloat162_rn(__uint_as_float(o[j + 2]) * inv, __uint_as_float(o[j + 3]) * inv);
            __nv_bfloat162 e2 = __floats2bfloat162_rn(__uint_as_float(o[j + 4]) * inv, __uint_as_float(o[j + 5]) * inv);
            __nv_bfloat162 e3 = __floats2bfloat162_rn(__uint_as_float(o[j + 6]) * inv, __uint_as_float(o[j + 7]) * inv);
            uint4 v;
            v.x = *reinterpret_cast<uint32_t*>(&e0); v.y = *reinterpret_cast<uint32_t*>(&e1);
            v.z = *reinterpret_cast<uint32_t*>(&e2); v.w = *reinterpret_cast<uint32_t*>(&e3);
            *reinterpret_cast<uint4*>(orow + col + j) = v;
          }
        };
        store16(o0, 0); store16(o1, 16); store16(o2, 32); store16(o3, 48);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem, 512u);
  }
}

int make_map(CUtensorMap* m, const void* base, uint64_t width, uint64_t rows, uint64_t seqs, uint64_t rs_bytes, uint64_t ss_bytes,
             uint32_t box_rows) {
  return make_map_bf16_3d(m, base, width, rows, seqs, rs_bytes, ss_bytes, 64, box_rows);
}

}  // namespace

int g_attn_blk = 1;        // option "attn_blk": bounded AR attention over 257..384 keys takes the block-wise kernel (one item per warpgroup)
void set_attn_blk(int v) { g_attn_blk = v; }

// parity-grade launches (AttnArgs::split_planes): two key halves of 16..192 keys, one stage of hi + lo tiles in shared memory
bool attention_tc_split_supported(int lq, int lk, int head_dim) {
  const int lk_pad = (lk + 15) & ~15;
  const int h0 = ((lk_pad / 2) + 15) & ~15;
  return head_dim == 64 && lq >= 1 && lk_pad >= 32 && lk_pad - h0 >= 16 && h0 <= 192 &&
         2 * Q_BYTES + 4 * lk_pad * 128 + 1024 + 256 + XCH_BYTES <= SMEM_LIMIT;
}

bool attention_tc_supported(const AttnArgs& a) {
  return a.dt == DT_BF16 && a.head_dim == 64 && a.lk <= MAX_LK && a.lk >= 1 && a.q_rs % 8 == 0 && a.k_rs % 8 == 0 &&
         a.v_rs % 8 == 0 && a.q_ss % 8 == 0 && a.k_ss % 8 == 0 && a.v_ss % 8 == 0 && a.o_rs % 8 == 0 && a.o_ss % 8 == 0 &&
         ((uintptr_t)a.q % 16 == 0) && ((uintptr_t)a.k % 16 == 0) && ((uintptr_t)a.v % 16 == 0) && ((uintptr_t)a.out % 16 == 0);
}

int launch_attention_tc(const AttnArgs& a, cudaStream_t st) {
  if (a.n_seq <= 0 || a.lq <= 0) return AT_OK;
  AT_REQUIRE(attention_tc_supported(a), "attention_tc: unsupported shape/stride (lk=%d head_dim=%d)", a.lk, a.head_dim);
  const DevCtx* dc = nullptr;
  AT_TRY(dev_ctx(&dc));
  const int g_num_sms = dc->num_sms;
  unsigned int* const g_err_flag = dc->err_flag;
  AT_TRY(ensure_dyn_smem((const void*)attn_tc_kernel<false>, SMEM_LIMIT));
  AT_TRY(ensure_dyn_smem((const void*)attn_tc_kernel<true>, SMEM_LIMIT));
  const bool split = a.split_planes != 0;
  AT_REQUIRE(!split || (attention_tc_split_supported(a.lq, a.lk, a.head_dim) && a.q_ss > 0 && a.k_ss > 0 && a.v_ss > 0 && a.o_rs % 4 == 0 &&
                        a.o_ss % 4 == 0), "attention_tc: unsupported split launch (lq=%d lk=%d)", a.lq, a.lk);
  AttnTcParams p;
  p.n_heads = a.n_heads; p.lq = a.lq; p.lk = a.lk;
  p.lk_pad = (a.lk + 15) & ~15;
  p.split_keys = (split || p.lk_pad > BUF_COLS) ? 1 : 0;
  p.h0 = p.split_keys ? (((p.lk_pad / 2) + 15) & ~15) : p.lk_pad;
  p.kv_bytes = p.lk_pad * 128;                      // lk_pad is a multiple of 16 rows -> a multiple of 2048 B (1024 B swizzle atoms)
  p.stage_bytes = split ? 2 * Q_BYTES + 4 * p.kv_bytes : Q_BYTES + 2 * p.kv_bytes;
  p.q_tiles = ceil_div(a.lq, QT);
  p.total_items = a.n_seq * a.n_heads * p.q_tiles;
  const int fixed = 1024 /*align*/ + 256 /*barriers*/ + XCH_BYTES;
  p.n_stages = (SMEM_LIMIT - fixed) / p.stage_bytes;
  if (p.n_stages > MAX_STAGES) p.n_stages = MAX_STAGES;
  AT_REQUIRE(p.n_stages >= 1, "attention_tc: %d keys do not fit in shared memory", a.lk);
  p.split = a.split;
  p.scale_log2e = a.scale * 1.4426950408889634f;
  p.bound = a.key_bound;
  p.out = (bf16*)a.out; p.out32 = (float*)a.out; p.o_ss = a.o_ss; p.o_rs = a.o_rs;
  p.lo_z = split ? a.n_seq : 0;
  const uint64_t n_z = (uint64_t)a.n_seq * (split ? 2 : 1);          // the lo planes are the sequences [n_seq, 2 n_seq) of each tensor
  p.err_flag = g_err_flag;
  CUtensorMap tmQ, tmK, tmV, tmK2, tmV2;
  const uint64_t wq = (uint64_t)a.n_heads * 64;
  auto ss = [](int64_t s, int64_t rs, int rows) { return (uint64_t)(s > 0 ? s : rs * rows) * 2; };   // n_seq == 1: any stride
  AT_TRY(make_map(&tmQ, a.q, wq, (uint64_t)a.lq, n_z, (uint64_t)a.q_rs * 2, ss(a.q_ss, a.q_rs, a.lq), QT));
  AT_TRY(make_map(&tmK, a.k, wq, (uint64_t)a.lk, n_z, (uint64_t)a.k_rs * 2, ss(a.k_ss, a.k_rs, a.lk), (uint32_t)p.h0));
  AT_TRY(make_map(&tmV, a.v, wq, (uint64_t)a.lk, n_z, (uint64_t)a.v_rs * 2, ss(a.v_ss, a.v_rs, a.lk), (uint32_t)p.h0));
  tmK2 = tmK; tmV2 = tmV;
  if (p.split_keys) {
    const uint32_t r2 = (uint32_t)(p.lk_pad - p.h0);
    AT_TRY(make_map(&tmK2, a.k, wq, (uint64_t)a.lk, n_z, (uint64_t)a.k_rs * 2, ss(a.k_ss, a.k_rs, a.lk), r2));
    AT_TRY(make_map(&tmV2, a.v, wq, (uint64_t)a.lk, n_z, (uint64_t)a.v_rs * 2, ss(a.v_ss, a.v_rs, a.lk), r2));
  }
  const size_t smem = (size_t)p.n_stages * p.stage_bytes + fixed;
  const int grid = p.total_items < g_num_sms ? p.total_items : g_num_sms;
  g_trace_dims[0] = a.n_seq * a.n_heads; g_trace_dims[1] = a.lq; g_trace_dims[2] = a.lk;
  if (split) {
    AT_CUDA(launch_k(attn_tc_kernel<true>, dim3(grid), dim3(384), smem, st, tmQ, tmK, tmV, tmK2, tmV2, p));
    AT_LAUNCH_CHECK();
    return AT_OK;
  }
  if (g_attn_blk && p.split_keys && a.key_bound && a.split == 0) {
    // two key blocks per item, one item per warpgroup (see attn_blk_kernel); heads without a usable bound run online
    AT_TRY(ensure_dyn_smem((const void*)attn_blk_kernel, SMEM_BLK));
    AT_CUDA(launch_k(attn_blk_kernel, dim3(grid), dim3(384), (size_t)SMEM_BLK, st, tmQ, tmK, tmV, tmK2, tmV2, p));
    AT_LAUNCH_CHECK();
    return AT_OK;
  }
  AT_CUDA(launch_k(attn_tc_kernel<false>, dim3(grid), dim3(384), smem, st, tmQ, tmK, tmV, tmK2, tmV2, p));
  AT_LAUNCH_CHECK();
  return AT_OK;
}

}  // namespace artalk
