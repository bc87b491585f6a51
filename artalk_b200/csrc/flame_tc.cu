// FLAME blendshapes + LBS on the tensor cores (app/flame_model/lbs.py:142-232 behind FLAMEModel.forward).
//
//   v_posed[f][e] = base[e] + sum_l coef[f][l] * dirs[l][e]      (e = 3*vertex + xyz;  GEMM frames x bases x 15069)
//   out[f][v]     = scale * (sum_j w[v][j] A[f][j]) (v_posed[f][v], 1)                  (skinning epilogue)
//
// The blend is a tcgen05 GEMM (M = 128 frames, N = 192 = 64 vertices, TMA-fed 3-stage ring, 2 TMEM accumulators).
// To keep ~fp32 accuracy on bf16 tensor cores both operands are split hi + lo (bf16 each) and the K dimension is
// laid out [hi | lo | hi] x [hi | hi | lo], i.e. hi*hi + lo*hi + hi*lo with fp32 accumulation (rel. error ~2^-16).
// 8 epilogue warps (one frame per thread) read the accumulator, add the template, skin with the frame's 5 relative
// transforms held in registers, and stage the tile in shared memory so the (N, 5023, 3) rows — which are only 4-byte
// aligned (60 276 B pitch) — are written with fully coalesced 128-byte warp stores.
#include "kernels.cuh"
#include "tc_ptx.cuh"

namespace artalk {

namespace {
using namespace ptx;

constexpr int BM = 128, BN = 192, BK = 64, STAGES = 3, VT = BN / 3;     // 64 vertices per tile
constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int OUT_PITCH = BN + 1;
constexpr int OUT_BYTES = BM * OUT_PITCH * 4;
constexpr int CONST_BYTES = 8 * 256 * 4;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + OUT_BYTES + CONST_BYTES + 1024 + 256;
static_assert(SMEM_BYTES <= 232448, "flame_tc: shared memory budget");
constexpr int COEF_A = 60;

struct FlameTcParams {
  int V, n_frames, m_tiles, n_tiles, total_tiles, num_kb;
  const float* base;        // [V*3] template (or the per-call static-shape template)
  const float* lbs_w;       // [V][5]
  const float* coef;        // [frame][coef_stride] fp32; relative transforms at coef_A_off
  int coef_stride, coef_A_off;
  float scale;
  float* verts;
  unsigned int* err_flag;
};

__global__ void __launch_bounds__(384, 1)
flame_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const FlameTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  float* out_s = reinterpret_cast<float*>(smem_raw + (smem_base - smem_u32(smem_raw)) + STAGES * STAGE_BYTES);
  float* const_s = out_s + BM * OUT_PITCH;                    // 8 warps x 256 floats of per-vertex constants
  const uint32_t bar_base = smem_base + STAGES * STAGE_BYTES + OUT_BYTES + CONST_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) { prefetch_tensormap(&tmA); prefetch_tensormap(&tmB); }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int n_idx = tile / p.m_tiles, mt = tile - n_idx * p.m_tiles;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u, p.err_flag, 0xF1A00001u);
          const uint32_t sa = smem_base + stage * STAGE_BYTES, sb = sa + A_BYTES;
          mbar_arrive_expect_tx(full_bar(stage), STAGE_BYTES);
          tma_load_3d(sa, &tmA, full_bar(stage), kb * BK, mt * BM, 0);
          tma_load_3d(sb, &tmB, full_bar(stage), kb * BK, n_idx * BN, 0);
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = idesc_bf16(BN);
      int stage = 0; uint32_t phase = 0; int it = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        mbar_wait(tempty_bar(acc), (((uint32_t)it >> 1) & 1u) ^ 1u, p.err_flag, 0xF1A00002u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase, p.err_flag, 0xF1A00003u);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * STAGE_BYTES, sb = sa + A_BYTES;
          const uint64_t da = desc_kmajor_sw128(sa), db = desc_kmajor_sw128(sb);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) mma_ss(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) ? 1u : 0u);
          tc_commit(empty_bar(stage));
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        tc_commit(tfull_bar(acc));
      }
    }
  } else if (warp >= 4) {
    const int we = warp - 4, q = we & 3, half = we >> 2;
    float* cw = const_s + we * 256;                            // [160 weights | 96 template coords] of 32 vertices
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const int row_local = q * 32 + lane;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const int n_idx = tile / p.m_tiles, mt = tile - n_idx * p.m_tiles;
      const int acc = it & 1;
      const int f = mt * BM + row_local;
      const bool row_ok = f < p.n_frames;
      float A[COEF_A];
      {
        const float* ap = p.coef + (int64_t)(row_ok ? f : 0) * p.coef_stride + p.coef_A_off;
#pragma unroll
        for (int i = 0; i < COEF_A; i += 4) {
          float4 t = *reinterpret_cast<const float4*>(ap + i);
          A[i] = t.x; A[i + 1] = t.y; A[i + 2] = t.z; A[i + 3] = t.w;
        }
      }
      // stage this warp's 32 vertices' constants (5 skinning weights + 3 template coords each) in its private smem
      // slice with coalesced loads; the per-vertex reads below are then conflict-free broadcasts
      {
        const int vbase = n_idx * VT + half * 32;
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 5; ++i) {
          const int idx = lane + 32 * i, v = vbase + idx / 5;
          cw[idx] = (v < p.V) ? __ldg(p.lbs_w + (int64_t)vbase * 5 + idx) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const int idx = lane + 32 * i, v = vbase + idx / 3;
          cw[160 + idx] = (v < p.V) ? __ldg(p.base + (int64_t)vbase * 3 + idx) : 0.f;
        }
        __syncwarp();
      }
      mbar_wait(tfull_bar(acc), ((uint32_t)it >> 1) & 1u, p.err_flag, 0xF1A00004u);
      tc_fence_after();
#pragma unroll 1
      for (int g8 = 0; g8 < 4; ++g8) {                       // 8 vertices = 24 accumulator columns per iteration
        const int vloc0 = half * 32 + g8 * 8;
        float x[24];
        {
          uint32_t r[24];
          const uint32_t ta = tmem_base + lane_addr + (uint32_t)(acc * BN + vloc0 * 3);
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                       : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(ta));
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                       : "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(ta + 8));
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                       : "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]) : "r"(ta + 16));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");      // one wait for the three loads
#pragma unroll
          for (int i = 0; i < 24; ++i) x[i] = __uint_as_float(r[i]);
        }
#pragma unroll
        for (int vv = 0; vv < 8; ++vv) {
          // branch free: vertices past V have zero weights / template (staged above) and are never stored
          const int vl = g8 * 8 + vv;                          // vertex index inside this warp's slice
          const float px = x[vv * 3] + cw[160 + vl * 3], py = x[vv * 3 + 1] + cw[160 + vl * 3 + 1],
                      pz = x[vv * 3 + 2] + cw[160 + vl * 3 + 2];
          float T[12];
          {
            const float w0 = cw[vl * 5];
#pragma unroll
            for (int k = 0; k < 12; ++k) T[k] = w0 * A[k];
          }
#pragma unroll
          for (int j = 1; j < 5; ++j) {
            const float wj = cw[vl * 5 + j];
#pragma unroll
            for (int k = 0; k < 12; ++k) T[k] = fmaf(wj, A[j * 12 + k], T[k]);
          }
          float* os = out_s + row_local * OUT_PITCH + (vloc0 + vv) * 3;
          os[0] = fmaf(T[0], px, fmaf(T[1], py, fmaf(T[2], pz, T[3]))) * p.scale;
          os[1] = fmaf(T[4], px, fmaf(T[5], py, fmaf(T[6], pz, T[7]))) * p.scale;
          os[2] = fmaf(T[8], px, fmaf(T[9], py, fmaf(T[10], pz, T[11]))) * p.scale;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));           // accumulator free for the MMA of tile it + 2
      asm volatile("bar.sync 1, 256;" ::: "memory");         // all 8 epilogue warps: tile staged
      const int e0 = n_idx * BN, e_max = p.V * 3;
      for (int rr = we; rr < BM; rr += 8) {
        const int fr = mt * BM + rr;
        if (fr >= p.n_frames) break;
        float* dst = p.verts + (int64_t)fr * e_max + e0;
        const float* src = out_s + rr * OUT_PITCH;
#pragma unroll
        for (int i = 0; i < BN / 32; ++i) {
          const int c = lane + 32 * i;
          if (e0 + c < e_max) dst[c] = src[c];
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");         // staging buffer reusable
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// coef fp32 [frame][stride] -> A' bf16 [frame][3*KS] = [hi | lo | hi] of bases [l_begin, l_begin + n_l), zero padded
__global__ void __launch_bounds__(256) flame_pack_kernel(const float* __restrict__ coef, int coef_stride, int l_begin, int n_l,
                                                         int KS, bf16* __restrict__ a_split, int64_t total) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % KS);
    const int64_t f = i / KS;
    const float v = (c < n_l) ? coef[f * coef_stride + l_begin + c] : 0.f;
    const bf16 hi = __float2bfloat16_rn(v);
    const bf16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
    bf16* row = a_split + f * 3 * KS;
    row[c] = hi; row[KS + c] = lo; row[2 * KS + c] = hi;
  }
}

}  // namespace

// coef: output of flame_coef_kernel; b_split: [V*3][3*KS] bf16 = [hi | hi | lo] of dirs[l_begin : l_begin + n_l]^T
int launch_flame_tc(const FlameModel& m, const float* base, const float* coef, int coef_stride, int l_begin, int n_l,
                    const void* b_split, int KS, void* a_split_ws, float* verts, int n_frames, cudaStream_t st) {
  if (n_frames <= 0) return AT_OK;
  AT_REQUIRE(KS % 64 == 0 && KS >= n_l && b_split && a_split_ws, "flame_tc: bad split operands");
  const DevCtx* dc = nullptr;
  AT_TRY(dev_ctx(&dc));
  const int g_num_sms = dc->num_sms;
  unsigned int* const g_err_flag = dc->err_flag;
  AT_TRY(ensure_dyn_smem((const void*)flame_tc_kernel, SMEM_BYTES));
  const int64_t total = (int64_t)n_frames * KS;
  int pgrid = (int)((total + 255) / 256);
  if (pgrid > 148 * 16) pgrid = 148 * 16;
  flame_pack_kernel<<<pgrid, 256, 0, st>>>(coef, coef_stride, l_begin, n_l, KS, (bf16*)a_split_ws, total);
  AT_LAUNCH_CHECK();
  FlameTcParams p;
  p.V = m.V; p.n_frames = n_frames; p.m_tiles = ceil_div(n_frames, BM); p.n_tiles = ceil_div(m.V * 3, BN);
  p.total_tiles = p.m_tiles * p.n_tiles; p.num_kb = 3 * KS / BK;
  p.base = base; p.lbs_w = m.lbs_weights; p.coef = coef; p.coef_stride = coef_stride;
  p.coef_A_off = m.n_shape + m.n_exp + 36; p.scale = m.scale; p.verts = verts; p.err_flag = g_err_flag;
  AT_REQUIRE(p.coef_A_off % 4 == 0 && coef_stride % 4 == 0, "flame_tc: coefficient rows must be 16-byte aligned");
  CUtensorMap tmA, tmB;
  AT_TRY(make_map_bf16_3d(&tmA, a_split_ws, (uint64_t)3 * KS, (uint64_t)n_frames, 1, (uint64_t)3 * KS * 2,
                          (uint64_t)n_frames * 3 * KS * 2, BK, BM));
  AT_TRY(make_map_bf16_3d(&tmB, b_split, (uint64_t)3 * KS, (uint64_t)m.V * 3, 1, (uint64_t)3 * KS * 2,
                          (uint64_t)m.V * 3 * 3 * KS * 2, BK, BN));
  const int grid = p.total_tiles < g_num_sms ? p.total_tiles : g_num_sms;
  flame_tc_kernel<<<grid, 384, SMEM_BYTES, st>>>(tmA, tmB, p);
  AT_LAUNCH_CHECK();
  return AT_OK;
}

}  // namespace artalk
