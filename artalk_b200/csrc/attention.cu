// Short-sequence multi-head attention (<= a few hundred keys, head_dim 32/64) in fp32 arithmetic with
// flash-style online softmax, plus the AR block's q/k L2-normalise + KV-cache scatter.
// Covers wav2vec2 self-attention (transformers modeling_wav2vec2.py:466-549), AR ModifiedSelfAttention
// (app/transformer.py:65-79; mask-free in the KV-cached schedule), VAE SimpleSelfAttention with the 2-block mask
// (app/modules/bitwise_vae.py:67-76,194-215) and the style encoder's nn.MultiheadAttention.
#include "kernels.cuh"

namespace artalk {

namespace {
constexpr int KT = 64;        // keys per shared-memory tile

// RPW query rows per warp (register blocking: every K / V element read from shared memory feeds RPW FMAs), WARPS warps per
// block. <4, 4> (16 rows per block, 41 KB static-size footprint) serves the few-row AR steps; <8, 8> (64 rows per block) the
// long sequences of the fp32 data-flow modes (wav2vec 199 rows, VAE 100 / 200, AR 50 / 100): twice the FMAs per shared-memory
// load and a quarter of the K / V tile loads.
template <int D, int RPW, int WARPS> struct AttnSmem {
  static constexpr int KS = KT * (D + 1), VS = KT * D, QS = WARPS * D * RPW, PS = WARPS * KT * RPW;
  static constexpr int BYTES = (KS + VS + QS + PS) * 4;
};

template <typename T, int D, int RPW, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) attn_kernel(AttnArgs a) {
  pdl_enter();
  static_assert(RPW % 4 == 0, "rows per warp are read as float4 groups");
  constexpr int DPL = D / 32;                 // output dims per lane
  constexpr int R4 = RPW / 4;
  using SM = AttnSmem<D, RPW, WARPS>;
  extern __shared__ __align__(16) float attn_smem[];
  float (*Vs)[D] = reinterpret_cast<float (*)[D]>(attn_smem);                                   // [KT][D]
  float (*Qs)[D][RPW] = reinterpret_cast<float (*)[D][RPW]>(attn_smem + SM::VS);                // [WARPS][D][RPW]
  float (*Ps)[KT][RPW] = reinterpret_cast<float (*)[KT][RPW]>(attn_smem + SM::VS + SM::QS);     // [WARPS][KT][RPW]
  float (*Ks)[D + 1] = reinterpret_cast<float (*)[D + 1]>(attn_smem + SM::VS + SM::QS + SM::PS); // [KT][D + 1]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int head = blockIdx.y, seq = blockIdx.z;
  const int q0 = blockIdx.x * (WARPS * RPW) + warp * RPW;
  const T* qb = reinterpret_cast<const T*>(a.q) + (int64_t)seq * a.q_ss + head * D;
  const T* kb = reinterpret_cast<const T*>(a.k) + (int64_t)seq * a.k_ss + head * D;
  const T* vb = reinterpret_cast<const T*>(a.v) + (int64_t)seq * a.v_ss + head * D;

  // stage this warp's RPW query rows (pre-scaled) as Qs[d][r]
  int lk_r[RPW];
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    int qi = q0 + r;
    lk_r[r] = (a.split > 0 && qi < a.split) ? a.split : a.lk;
    for (int d = lane; d < D; d += 32)
      Qs[warp][d][r] = (qi < a.lq) ? to_f32(qb[(int64_t)qi * a.q_rs + d]) * a.scale : 0.f;
  }
  // keys needed by any row of the block
  int blk_q_last = min(a.lq, (int)(blockIdx.x + 1) * (WARPS * RPW)) - 1;
  int lk_blk = (a.split > 0 && blk_q_last < a.split) ? a.split : a.lk;

  float m[RPW], l[RPW], acc[RPW][DPL];
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    m[r] = -INFINITY; l[r] = 0.f;
#pragma unroll
    for (int j = 0; j < DPL; ++j) acc[r][j] = 0.f;
  }

  for (int k0 = 0; k0 < lk_blk; k0 += KT) {
    __syncthreads();
    // cooperative tile load, 4 elements per access
    for (int i = tid; i < KT * (D / 4); i += WARPS * 32) {
      int key = i / (D / 4), d4 = (i - key * (D / 4)) * 4;
      float kv[4] = {0.f, 0.f, 0.f, 0.f}, vv[4] = {0.f, 0.f, 0.f, 0.f};
      if (k0 + key < lk_blk) {
        load4(kb + (int64_t)(k0 + key) * a.k_rs + d4, kv);
        load4(vb + (int64_t)(k0 + key) * a.v_rs + d4, vv);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) { Ks[key][d4 + j] = kv[j]; Vs[key][d4 + j] = vv[j]; }
    }
    __syncthreads();
    // scores for keys (lane, lane + 32) x RPW rows
    float s[RPW][2];
#pragma unroll
    for (int r = 0; r < RPW; ++r) s[r][0] = s[r][1] = 0.f;
#pragma unroll 8
    for (int d = 0; d < D; ++d) {
      const float ka = Ks[lane][d], kc = Ks[lane + 32][d];
#pragma unroll
      for (int g = 0; g < R4; ++g) {
        const float4 q4 = *reinterpret_cast<const float4*>(&Qs[warp][d][4 * g]);
        s[4 * g][0] = fmaf(q4.x, ka, s[4 * g][0]); s[4 * g][1] = fmaf(q4.x, kc, s[4 * g][1]);
        s[4 * g + 1][0] = fmaf(q4.y, ka, s[4 * g + 1][0]); s[4 * g + 1][1] = fmaf(q4.y, kc, s[4 * g + 1][1]);
        s[4 * g + 2][0] = fmaf(q4.z, ka, s[4 * g + 2][0]); s[4 * g + 2][1] = fmaf(q4.z, kc, s[4 * g + 2][1]);
        s[4 * g + 3][0] = fmaf(q4.w, ka, s[4 * g + 3][0]); s[4 * g + 3][1] = fmaf(q4.w, kc, s[4 * g + 3][1]);
      }
    }
    float p[RPW][2];
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
      if (k0 + lane >= lk_r[r]) s[r][0] = -INFINITY;
      if (k0 + lane + 32 >= lk_r[r]) s[r][1] = -INFINITY;
      float tmax = warp_max(fmaxf(s[r][0], s[r][1]));
      float m_new = fmaxf(m[r], tmax);
      float corr, p0, p1;
      if (m_new == -INFINITY) { corr = 1.f; p0 = p1 = 0.f; }      // row sees no key in this tile yet
      else { corr = __expf(m[r] - m_new); p0 = __expf(s[r][0] - m_new); p1 = __expf(s[r][1] - m_new); }
      l[r] = l[r] * corr + warp_sum(p0 + p1);
      m[r] = m_new;
#pragma unroll
      for (int j = 0; j < DPL; ++j) acc[r][j] *= corr;
      p[r][0] = p0; p[r][1] = p1;
    }
#pragma unroll
    for (int g = 0; g < R4; ++g) {
      *reinterpret_cast<float4*>(&Ps[warp][lane][4 * g]) = make_float4(p[4 * g][0], p[4 * g + 1][0], p[4 * g + 2][0], p[4 * g + 3][0]);
      *reinterpret_cast<float4*>(&Ps[warp][lane + 32][4 * g]) = make_float4(p[4 * g][1], p[4 * g + 1][1], p[4 * g + 2][1], p[4 * g + 3][1]);
    }
    __syncwarp();
    int kmax = min(KT, lk_blk - k0);
    for (int key = 0; key < kmax; ++key) {
      float vv[DPL];
#pragma unroll
      for (int j = 0; j < DPL; ++j) vv[j] = Vs[key][lane + 32 * j];
#pragma unroll
      for (int g = 0; g < R4; ++g) {
        const float4 p4 = *reinterpret_cast<const float4*>(&Ps[warp][key][4 * g]);
#pragma unroll
        for (int j = 0; j < DPL; ++j) {
          acc[4 * g][j] = fmaf(p4.x, vv[j], acc[4 * g][j]);
          acc[4 * g + 1][j] = fmaf(p4.y, vv[j], acc[4 * g + 1][j]);
          acc[4 * g + 2][j] = fmaf(p4.z, vv[j], acc[4 * g + 2][j]);
          acc[4 * g + 3][j] = fmaf(p4.w, vv[j], acc[4 * g + 3][j]);
        }
      }
    }
    __syncwarp();
  }
  T* ob = reinterpret_cast<T*>(a.out) + (int64_t)seq * a.o_ss + head * D;
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    int qi = q0 + r;
    if (qi >= a.lq) continue;
    float inv = 1.0f / l[r];
#pragma unroll
    for (int j = 0; j < DPL; ++j) ob[(int64_t)qi * a.o_rs + lane + 32 * j] = from_f32<T>(acc[r][j] * inv);
  }
}

template <typename T, int D, int RPW, int WARPS>
int launch_simt(const AttnArgs& a, cudaStream_t st) {
  constexpr int smem = AttnSmem<D, RPW, WARPS>::BYTES;
  AT_TRY(ensure_dyn_smem((const void*)attn_kernel<T, D, RPW, WARPS>, smem));
  dim3 grid(ceil_div(a.lq, WARPS * RPW), a.n_heads, a.n_seq);
  AT_CUDA(launch_k(attn_kernel<T, D, RPW, WARPS>, grid, dim3(WARPS * 32), (size_t)smem, st, a));
  return AT_OK;
}
template <typename T, int D>
int launch_simt_rows(const AttnArgs& a, cudaStream_t st) {
  return a.lq > 32 ? launch_simt<T, D, 8, 8>(a, st) : launch_simt<T, D, 4, 4>(a, st);
}
}  // namespace

bool attention_split_supported(const AttnArgs& a) {
  return a.dt == DT_F32 && attention_tc_split_supported(a.lq, a.lk, a.head_dim) && a.q_rs % 4 == 0 && a.k_rs % 4 == 0 && a.v_rs % 4 == 0 &&
         a.q_ss % 4 == 0 && a.k_ss % 4 == 0 && a.v_ss % 4 == 0 && a.o_rs % 4 == 0 && a.o_ss % 4 == 0 && ((uintptr_t)a.q % 16 == 0) &&
         ((uintptr_t)a.k % 16 == 0) && ((uintptr_t)a.v % 16 == 0) && ((uintptr_t)a.out % 16 == 0);
}
static size_t pad128(size_t n) { return (n + 127) & ~(size_t)127; }
size_t attention_split_scratch_bytes(const AttnArgs& a) {
  const size_t W = (size_t)a.n_heads * a.head_dim;
  const size_t nq = (size_t)a.n_seq * a.lq * W, nk = (size_t)a.n_seq * a.lk * W;
  return (pad128(2 * nq) + 2 * pad128(2 * nk)) * sizeof(bf16);
}
int launch_attention_split(const AttnArgs& a, void* scratch, cudaStream_t st) {
  if (a.n_seq <= 0 || a.lq <= 0) return AT_OK;
  AT_REQUIRE(attention_split_supported(a) && scratch && ((uintptr_t)scratch % 256 == 0), "attention_split: unsupported launch (lq=%d lk=%d)", a.lq, a.lk);
  const int W = a.n_heads * a.head_dim;
  const size_t nq = (size_t)a.n_seq * a.lq * W, nk = (size_t)a.n_seq * a.lk * W;
  bf16* qs = (bf16*)scratch;
  bf16* ks = qs + pad128(2 * nq);
  bf16* vs = ks + pad128(2 * nk);
  // with a single sequence the views may carry a zero sequence stride
  AT_TRY(launch_split2_rows((const float*)a.q, a.q_ss, a.q_rs, a.n_seq, a.lq, W, qs, st));
  AT_TRY(launch_split2_rows((const float*)a.k, a.k_ss, a.k_rs, a.n_seq, a.lk, W, ks, st));
  AT_TRY(launch_split2_rows((const float*)a.v, a.v_ss, a.v_rs, a.n_seq, a.lk, W, vs, st));
  AttnArgs b = a;
  b.dt = DT_BF16; b.q = qs; b.k = ks; b.v = vs; b.split_planes = 1;
  b.q_ss = (int64_t)a.lq * W; b.q_rs = W; b.k_ss = b.v_ss = (int64_t)a.lk * W; b.k_rs = b.v_rs = W;
  return launch_attention_tc(b, st);
}

int g_attn_simt_max_lq = 0;     // bf16 launches with at most this many query rows take the SIMT kernel (option "attn_simt_max_lq")
void set_attn_simt_max_lq(int v) { g_attn_simt_max_lq = v; }

int launch_attention(const AttnArgs& a, cudaStream_t st) {
  if (a.n_seq <= 0 || a.lq <= 0) return AT_OK;
  if (attention_tc_supported(a) && a.lq > g_attn_simt_max_lq) return launch_attention_tc(a, st);
  AT_REQUIRE(a.head_dim == 64 || a.head_dim == 32, "attention: head_dim %d", a.head_dim);
  AT_REQUIRE(a.lk > 0 && a.k_rs % 4 == 0 && a.v_rs % 4 == 0 && a.k_ss % 4 == 0 && a.v_ss % 4 == 0,
             "attention: key/value strides must be multiples of 4");
  g_trace_dims[0] = a.n_seq * a.n_heads; g_trace_dims[1] = a.lq; g_trace_dims[2] = a.lk;
  if (a.dt == DT_F32) {
    if (a.head_dim == 64) AT_TRY((launch_simt_rows<float, 64>(a, st)));
    else AT_TRY((launch_simt_rows<float, 32>(a, st)));
  } else {
    if (a.head_dim == 64) AT_TRY((launch_simt_rows<bf16, 64>(a, st)));
    else AT_TRY((launch_simt_rows<bf16, 32>(a, st)));
  }
  AT_LAUNCH_CHECK();
  return AT_OK;
}

// ---------------------------------------------------------------- q/k L2-normalise + KV-cache scatter
// One warp per (row, head): lane owns 2 of the 64 head dims. F.normalize: x / max(||x||, 1e-12).
template <typename T>
__global__ void __launch_bounds__(256) qkv_norm_scatter_kernel(const T* __restrict__ qkv, int64_t qkv_rs, int has_q,
                                                               const float* __restrict__ head_scale, T* __restrict__ qbuf,
                                                               T* __restrict__ kcache, T* __restrict__ vcache, RowMap kv_map,
                                                               int rows, int n_heads) {
  pdl_enter();
  int C = n_heads * 64;
  int64_t gw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (gw >= (int64_t)rows * n_heads) return;
  int r = (int)(gw / n_heads), h = (int)(gw - (int64_t)r * n_heads);
  const T* src = qkv + (int64_t)r * qkv_rs + h * 64 + lane * 2;
  int64_t dst = kv_map.off(r) + h * 64 + lane * 2;
  int koff = has_q ? C : 0;
  if (has_q) {
    float q0 = to_f32(src[0]), q1 = to_f32(src[1]);
    float n = sqrtf(warp_sum(q0 * q0 + q1 * q1));
    float sc = head_scale[h] / fmaxf(n, 1e-12f);
    T* qd = qbuf + (int64_t)r * C + h * 64 + lane * 2;
    qd[0] = from_f32<T>(q0 * sc); qd[1] = from_f32<T>(q1 * sc);
  }
  float k0 = to_f32(src[koff]), k1 = to_f32(src[koff + 1]);
  float n = sqrtf(warp_sum(k0 * k0 + k1 * k1));
  float sc = 1.0f / fmaxf(n, 1e-12f);
  kcache[dst] = from_f32<T>(k0 * sc); kcache[dst + 1] = from_f32<T>(k1 * sc);
  vcache[dst] = src[koff + C]; vcache[dst + 1] = src[koff + C + 1];
}

int launch_qkv_norm_scatter(const void* qkv, int dt, int64_t qkv_rs, int has_q, const float* head_scale, void* qbuf,
                            void* kcache, void* vcache, RowMap kv_map, int rows, int n_heads, cudaStream_t st) {
  if (rows <= 0) return AT_OK;
  int64_t warps = (int64_t)rows * n_heads;
  int grid = (int)((warps + 7) / 8);
  if (dt == DT_F32)
    AT_CUDA(launch_k(qkv_norm_scatter_kernel<float>, dim3(grid), dim3(256), 0, st, (const float*)qkv, qkv_rs, has_q, head_scale, (float*)qbuf,
                                                         (float*)kcache, (float*)vcache, kv_map, rows, n_heads));
  else
    AT_CUDA(launch_k(qkv_norm_scatter_kernel<bf16>, dim3(grid), dim3(256), 0, st, (const bf16*)qkv, qkv_rs, has_q, head_scale, (bf16*)qbuf,
                                                        (bf16*)kcache, (bf16*)vcache, kv_map, rows, n_heads));
  AT_LAUNCH_CHECK();
  return AT_OK;
}

}  // namespace artalk
