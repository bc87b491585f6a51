// wav2vec2 positional convolution (modeling_wav2vec2.py:360-368,764-765: grouped Conv1d, 16 groups x 64 channels, kernel 128,
// padding 64, last output dropped, + bias, GELU, + residual) as a tcgen05 CTA-pair GEMM with N = 256.
//
// As a plain per-group GEMM the conv has N = 64 output channels: a 128x64 tile re-reads its A operand once per 64 output
// columns and the tensor pipe sits at 22 % (profiles/r2_ncu_kernel_table.md, gemm_tc_kernel<64, 1>). Here FOUR consecutive
// output frames share one A row ("shift-4" form): with t = 4 t' + s,
//   out[4t'+s][co] = sum_{j'} W[co][j' - s] . x[4t' + j' - 64]        (j' = j + s in [0, 130], W zero outside [0, 128))
// so A'[t'][j'] = x[4t' + j' - 64] (one row per FOUR frames) and B'[s*64 + co][j'] = W[co][j' - s] (four shifted copies of the
// group's filter, built once at load: weights.repack "w2v.pos.w4"). M shrinks 4x, N grows to 256, K grows 128 -> 131 taps:
// the same flops (+2 %) in full-rate 256x256 pair tiles, and A is read once per 256 output values instead of once per 64.
// Products are accumulated in the same tap order as before (the extra products are exact zeros), so results are bit-identical
// to the N = 64 kernel.
//
// A operand: x is [chunk][F = 199][1024] bf16. Tap offset d = j' - 64 = 4q + r0 selects input frames 4(t' + q) + r0: one
// tensor map per r0 (base x + r0 rows; dims (channel, t' (frames r0, r0+4, ...), chunk); box 64 channels x 50 t' x 2 chunks),
// so one TMA load per k-block brings the 100 live rows of a CTA's tile (two chunks) and the conv padding is TMA zero fill.
// Tile = 4 chunks x 256 (shift, channel) columns of one group over a CTA pair; tiles are walked group-major so the clusters
// share one group's 4.3 MB of shifted filters in L2.
// Warp roles: 0 = TMA producer, 1 = MMA issuer (leader CTA), 2 = TMEM allocator, 4..11 = epilogue.
#define ARTALK_PDL_CLASS 1
#include "kernels.cuh"
#include "tc_ptx.cuh"

namespace artalk {

namespace {

using namespace ptx;

constexpr int BK = 64;
constexpr int A_BYTES = 128 * BK * 2;            // 16 KB slot (100 rows written per load)
constexpr int W_BYTES = 128 * BK * 2;            // this CTA's half of the 256 (shift, channel) filter rows
constexpr int STAGE_BYTES = A_BYTES + W_BYTES;
constexpr int STAGES = 6;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 1024 /*barriers*/;

struct PosParams {
  int n_chunks, F, H, tpc;          // tpc = ceil(F / 4): rows (groups of four frames) per chunk
  int n_mt, groups, num_kb, total_tiles;
  const float* bias; const float* resid; float* out;
  unsigned int* err_flag;
};

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
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {      // .release.cta: see gemm_tc.cu
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_3d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t cluster_bar, int x, int y, int z) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(cluster_bar), "r"(x), "r"(y), "r"(z) : "memory");
}
__device__ __forceinline__ void tc_commit_2sm(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void mma_2sm(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
// D = F32, A = B = BF16, both K-major, M = 256 (CTA pair), N = 256
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);

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

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(384, 1)
posconv4_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmA3,
                const __grid_constant__ CUtensorMap tmW, const PosParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + STAGES * STAGE_BYTES;
  // barriers (8 B each): full[STAGES] (leader), empty[STAGES], tmem_full[2], tmem_empty[2] (leader), then the TMEM base word
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmA0); prefetch_tensormap(&tmA1); prefetch_tensormap(&tmA2); prefetch_tensormap(&tmA3);
    prefetch_tensormap(&tmW);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 2); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 16); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  // rows 100..127 of every A slot are never written by the loads (a box is 2 chunks x 50 rows): zero them once so the MMA's
  // unused accumulator rows stay finite
  {
    const int live_bytes = 2 * p.tpc * 128;
    for (int s = 0; s < STAGES; ++s) {
      uint8_t* slot = smem_raw + (smem_base - smem_u32(smem_raw)) + s * STAGE_BYTES;
      for (int o = live_bytes + (int)threadIdx.x * 16; o < A_BYTES; o += 384 * 16) *reinterpret_cast<uint4*>(slot + o) = make_uint4(0u, 0u, 0u, 0u);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_launch_dependents();

  const int n_clusters = (int)(gridDim.x >> 1), cluster_id = (int)(blockIdx.x >> 1);

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      const CUtensorMap* mapA[4] = {&tmA0, &tmA1, &tmA2, &tmA3};
      int stage = 0; uint32_t phase = 0;
      pdl_wait();
      for (int tile = cluster_id; tile < p.total_tiles; tile += n_clusters) {
        const int g = tile / p.n_mt, mt = tile - g * p.n_mt;
        const int b0 = 4 * mt + 2 * (int)rank;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u, p.err_flag, 0x31u);
          const uint32_t sa = smem_base + stage * STAGE_BYTES, sb = sa + A_BYTES;
          const uint32_t lead_full = mapa_shared(full_bar(stage), 0);
          const uint32_t a_bytes = (uint32_t)(2 * p.tpc * 128);
          if (rank == 0) mbar_arrive_expect_tx(full_bar(stage), 2u * (a_bytes + (uint32_t)W_BYTES));
          else mbar_arrive_cluster(lead_full);
          const int d = kb - 64, q = d >> 2, r0 = d & 3;           // input frame = 4 (t' + q) + r0
          tma_load_3d_2sm(sa, mapA[r0], lead_full, g * 64, q, b0);
          tma_load_3d_2sm(sb, &tmW, lead_full, kb * BK, (int)rank * 128, g);
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (lane == 0 && rank == 0) {
      int stage = 0; uint32_t phase = 0; int it = 0;
      for (int tile = cluster_id; tile < p.total_tiles; tile += n_clusters, ++it) {
        const int acc = it & 1;
        mbar_wait(tempty_bar(acc), (((uint32_t)it >> 1) & 1u) ^ 1u, p.err_flag, 0x32u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 256);
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase, p.err_flag, 0x33u);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * STAGE_BYTES, sb = sa + A_BYTES;
          const uint64_t da = desc_kmajor_sw128(sa), db = desc_kmajor_sw128(sb);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) mma_2sm(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), IDESC, (kb | k) != 0 ? 1u : 0u);
          tc_commit_2sm(empty_bar(stage));
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        tc_commit_2sm(tfull_bar(acc));
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (both CTAs): thread = one A' row = four output frames of one chunk =====================
    // column chunk c (32 wide) of the accumulator = shift c / 2, channels (c % 2) * 32 .. +32 of the group; the two warp
    // halves take the even / odd chunks, i.e. the lower / upper 32 channels of the group for all four shifts.
    // A tile's MMAs take ~35 us (131 k-blocks), so the row-per-thread fp32 accesses below are far off the critical path.
    const int q = (warp - 4) & 3, half = (warp - 4) >> 2;
    const int m = q * 32 + lane;
    const int cb = m / p.tpc, tp = m - cb * p.tpc;
    int it = 0;
    pdl_wait();
    for (int tile = cluster_id; tile < p.total_tiles; tile += n_clusters, ++it) {
      const int g = tile / p.n_mt, mt = tile - g * p.n_mt;
      const int acc = it & 1;
      const int b = 4 * mt + 2 * (int)rank + cb;
      const bool row_live = cb < 2 && b < p.n_chunks;
      const int col = g * 64 + half * 32;
      const int64_t row0 = ((int64_t)b * p.F + 4 * tp) * p.H + col;            // frame 4 t' of the chunk
      if (row_live) {
#pragma unroll
        for (int s = 0; s < 4; ++s)
          if (4 * tp + s < p.F) prefetch_l2(p.resid + row0 + (int64_t)s * p.H);
      }
      float4 bv[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) bv[j] = __ldg(reinterpret_cast<const float4*>(p.bias + col) + j);
      mbar_wait(tfull_bar(acc), ((uint32_t)it >> 1) & 1u, p.err_flag, 0x34u);
      tc_fence_after();
#pragma unroll 1
      for (int s = 0; s < 4; ++s) {
        const bool ok = row_live && (4 * tp + s < p.F);
        const int64_t off = row0 + (int64_t)s * p.H;
        float4 rv[8];
        if (ok) {
#pragma unroll
          for (int j = 0; j < 8; ++j) rv[j] = *(reinterpret_cast<const float4*>(p.resid + off) + j);
        }
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 256 + (2 * s + half) * 32), v);
        if (ok) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 o;
            o.x = rv[j].x + gelu_erf_fast(v[4 * j] + bv[j].x);
            o.y = rv[j].y + gelu_erf_fast(v[4 * j + 1] + bv[j].y);
            o.z = rv[j].z + gelu_erf_fast(v[4 * j + 2] + bv[j].z);
            o.w = rv[j].w + gelu_erf_fast(v[4 * j + 3] + bv[j].w);
            *(reinterpret_cast<float4*>(p.out + off) + j) = o;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_shared(tempty_bar(acc), 0));
    }
  }
  // ---- teardown: the peer's shared memory and the leader's barriers stay alive until both CTAs are done
  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

int make_map(CUtensorMap* m, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t s1_bytes, uint64_t s2_bytes,
             uint32_t b0, uint32_t b1, uint32_t b2) {
  EncodeTiledFn enc = get_encode();
  AT_REQUIRE(enc, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {s1_bytes, s2_bytes};
  cuuint32_t box[3] = {b0, b1, b2};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("posconv4: cuTensorMapEncodeTiled failed (%d): dims=(%llu,%llu,%llu) strides=(%llu,%llu) box=(%u,%u,%u)", (int)r,
                   (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2, (unsigned long long)s1_bytes,
                   (unsigned long long)s2_bytes, b0, b1, b2);
    return AT_ECUDA;
  }
  return AT_OK;
}

}  // namespace

bool posconv4_supported(int F, int H, int groups, int taps) {
  return H == groups * 64 && taps % 4 == 0 && taps / 2 == 64 && F >= 4 && 2 * ((F + 3) / 4) <= 128;
}

// x [n_chunks][F][H] bf16 (contiguous), w4 [groups][256][(taps + 3) * 64] bf16, bias [H], resid / out [n_chunks * F][H] fp32
int launch_posconv4(const void* x, const void* w4, const float* bias, const float* resid, float* out, int n_chunks, int F, int H,
                    int groups, int taps, cudaStream_t st) {
  if (n_chunks <= 0) return AT_OK;
  AT_REQUIRE(posconv4_supported(F, H, groups, taps), "posconv4: unsupported shape (F=%d H=%d groups=%d taps=%d)", F, H, groups, taps);
  AT_REQUIRE(((uintptr_t)x % 16 == 0) && ((uintptr_t)w4 % 16 == 0) && ((uintptr_t)bias % 16 == 0) && ((uintptr_t)resid % 16 == 0) &&
             ((uintptr_t)out % 16 == 0), "posconv4: operands must be 16-byte aligned");
  const DevCtx* dc = nullptr;
  AT_TRY(dev_ctx(&dc));
  PosParams p;
  p.n_chunks = n_chunks; p.F = F; p.H = H; p.tpc = (F + 3) / 4;
  p.n_mt = (n_chunks + 3) / 4; p.groups = groups; p.num_kb = taps + 3; p.total_tiles = p.n_mt * groups;
  p.bias = bias; p.resid = resid; p.out = out; p.err_flag = dc->err_flag;
  CUtensorMap tmA[4], tmW;
  for (int r0 = 0; r0 < 4; ++r0) {
    const uint64_t cnt = (uint64_t)((F - r0 + 3) / 4);           // frames r0, r0 + 4, ... < F
    AT_TRY(make_map(&tmA[r0], (const bf16*)x + (size_t)r0 * H, (uint64_t)H, cnt, (uint64_t)n_chunks, (uint64_t)4 * H * 2, (uint64_t)F * H * 2,
                    64, (uint32_t)p.tpc, 2));
  }
  const uint64_t Kp = (uint64_t)p.num_kb * 64;
  AT_TRY(make_map(&tmW, w4, Kp, 256, (uint64_t)groups, Kp * 2, 256 * Kp * 2, 64, 128, 1));
  static int max_clusters_dev[16] = {-1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1};
  int& max_clusters = per_device_slot(max_clusters_dev);
  if (max_clusters < 0) {
    AT_TRY(ensure_dyn_smem((const void*)posconv4_kernel, SMEM_BYTES));
    cudaLaunchConfig_t qc = {};
    qc.gridDim = dim3(dc->num_sms & ~1); qc.blockDim = dim3(384); qc.dynamicSmemBytes = SMEM_BYTES;
    cudaLaunchAttribute qa[1];
    qa[0].id = cudaLaunchAttributeClusterDimension; qa[0].val.clusterDim.x = 2; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
    qc.attrs = qa; qc.numAttrs = 1;
    int n = 0;
    AT_CUDA(cudaOccupancyMaxActiveClusters(&n, posconv4_kernel, &qc));
    max_clusters = n > 0 ? n : 1;
    if (max_clusters > dc->num_sms / 2) max_clusters = dc->num_sms / 2;
  }
  const int clusters = p.total_tiles < max_clusters ? p.total_tiles : max_clusters;
  g_trace_dims[0] = n_chunks * F; g_trace_dims[1] = H; g_trace_dims[2] = taps * 64;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * clusters); cfg.blockDim = dim3(384); cfg.dynamicSmemBytes = SMEM_BYTES; cfg.stream = st;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl_on() ? 2 : 1;
  AT_CUDA(cudaLaunchKernelEx(&cfg, posconv4_kernel, tmA[0], tmA[1], tmA[2], tmA[3], tmW, p));
  AT_LAUNCH_CHECK();
  return AT_OK;
}

}  // namespace artalk
