// Bit-plane kernels: argmax head -> packed bit words, bit words -> next-scale token embeddings,
// bit words -> VAE decoder latent, residual multi-scale BSQ re-quantisation.
// A token's 32 code bits are one uint32 (bit j = code dim j). The linear-upsample / area-pool operators of
// app/modules/bitwise_vae.py:227-305 are fixed tables (BitsTables) built on the host with ATen's index formula.
#include "kernels.cuh"

namespace artalk {

namespace {
constexpr int CD = 32;                  // code dim
constexpr int MAXT = 128;               // max frames per chunk supported by the shared-memory tiles
__device__ __forceinline__ float bit_val(uint32_t w, int c) { return ((w >> c) & 1u) ? 0.17677669529663687f : -0.17677669529663687f; }
// (bit*2-1)/sqrt(32) evaluated like the reference: (b*2 - 1.0) / 32**0.5 ; 1/sqrt(32) rounds to the same float
}

// ---------------------------------------------------------------- argmax over bit pairs
__global__ void __launch_bounds__(256) argmax_bits_kernel(const float* __restrict__ logits, RowMap l_map,
                                                          uint32_t* __restrict__ words, RowMap w_map, int rows) {
  pdl_enter();
  int row = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  float2 l = *reinterpret_cast<const float2*>(logits + l_map.off(row) + lane * 2);
  uint32_t w = __ballot_sync(0xffffffffu, l.y > l.x);         // argmax tie -> index 0
  if (lane == 0) words[w_map.off(row)] = w;
}

int launch_argmax_bits(const float* logits, RowMap l_map, uint32_t* words, RowMap w_map, int rows, cudaStream_t st) {
  if (rows <= 0) return AT_OK;
  AT_CUDA(launch_k(argmax_bits_kernel, dim3(ceil_div(rows, 8)), dim3(256), 0, st, logits, l_map, words, w_map, rows));
  AT_LAUNCH_CHECK();
  return AT_OK;
}

// ---------------------------------------------------------------- tokens from bits
template <typename TO>
__global__ void __launch_bounds__(256) bits_tokens_kernel(BitsTables tb, const uint32_t* __restrict__ words, int64_t words_cs,
                                                          const float* __restrict__ style, const float* __restrict__ embed_w,
                                                          const float* __restrict__ embed_b, const float* __restrict__ pos,
                                                          TO* __restrict__ out, int q_lo, int q_hi, int C) {
  pdl_enter();
  __shared__ uint32_t sw[256];
  __shared__ float f_hat[MAXT][CD];
  __shared__ float feat[MAXT][CD];
  const int clip = blockIdx.x, tid = threadIdx.x;
  const int n = blockIdx.y * blockDim.x + tid;            // output column
  const int T = tb.T;
  const int base_row = q_lo ? tb.cum[q_lo - 1] : 0;
  const int rows_per_clip = tb.cum[q_hi] - base_row;
  TO* oc = out + (int64_t)clip * rows_per_clip * C;
  // operator tables -> shared memory in one round trip (read per element from global they were a chain of dependent L2
  // loads in every level's loops: 72 us for the 181-token launch)
  __shared__ int s_i0[8 * MAXT / 2], s_i1[8 * MAXT / 2], s_ps[8 * MAXT / 2], s_pe[8 * MAXT / 2];
  __shared__ float s_w1[8 * MAXT / 2];
  for (int i = tid; i < (q_hi + 1) * T; i += blockDim.x) {
    s_i0[i] = tb.up_i0[i]; s_i1[i] = tb.up_i1[i]; s_w1[i] = tb.up_w1[i]; s_ps[i] = tb.pool_start[i]; s_pe[i] = tb.pool_end[i];
  }
  for (int i = tid; i < tb.L; i += blockDim.x) sw[i] = words[(int64_t)clip * words_cs + i];
  for (int i = tid; i < T * CD; i += blockDim.x) (&f_hat[0][0])[i] = 0.f;
  float w[CD];
  float be = 0.f;
  if (n < C) {
#pragma unroll
    for (int c = 0; c < CD; c += 4) {
      float4 t = *reinterpret_cast<const float4*>(embed_w + (int64_t)n * CD + c);
      w[c] = t.x; w[c + 1] = t.y; w[c + 2] = t.z; w[c + 3] = t.w;
    }
    be = embed_b[n];
  }
  __syncthreads();
  if (q_lo == 0 && n < C) oc[n] = from_f32<TO>(style[(int64_t)clip * C + n] + pos[n]);
  for (int q = 1; q <= q_hi; ++q) {
    // f_hat += U_{pn[q-1]} h^{(q-1)}
    const int lvl = q - 1, src0 = lvl ? tb.cum[lvl - 1] : 0;
    for (int i = tid; i < T * CD; i += blockDim.x) {
      int t = i >> 5, c = i & 31;
      int i0 = s_i0[lvl * T + t], i1 = s_i1[lvl * T + t];
      float w1 = s_w1[lvl * T + t], w0 = 1.0f - w1;
      f_hat[t][c] += w0 * bit_val(sw[src0 + i0], c) + w1 * bit_val(sw[src0 + i1], c);
    }
    __syncthreads();
    if (q < q_lo) continue;
    const int pq = tb.pn[q];
    for (int i = tid; i < pq * CD; i += blockDim.x) {
      int r = i >> 5, c = i & 31;
      int s = s_ps[q * T + r], e = s_pe[q * T + r];
      float a = 0.f;
      for (int t = s; t < e; ++t) a += f_hat[t][c];
      feat[r][c] = a / (float)(e - s);
    }
    __syncthreads();
    if (n < C) {
      const int row0 = tb.cum[q - 1];
      // the position-embedding loads (one L2 round trip each, independent of everything else) are the latency of this loop:
      // groups of 16 rows, the next group's loads in flight while the current group's 16 x 32 FMAs run
      constexpr int G = 16;
      float pv[G];
#pragma unroll
      for (int u = 0; u < G; ++u) pv[u] = (u < pq) ? __ldg(pos + (int64_t)(row0 + u) * C + n) : 0.f;
      for (int r = 0; r < pq; r += G) {
        float nx[G];
#pragma unroll
        for (int u = 0; u < G; ++u) nx[u] = (r + G + u < pq) ? __ldg(pos + (int64_t)(row0 + r + G + u) * C + n) : 0.f;
#pragma unroll
        for (int u = 0; u < G; ++u) {
          if (r + u < pq) {
            float a = be;
#pragma unroll
            for (int c = 0; c < CD; ++c) a = fmaf(feat[r + u][c], w[c], a);
            a += pv[u];
            oc[(int64_t)(row0 - base_row + r + u) * C + n] = from_f32<TO>(a);
          }
        }
#pragma unroll
        for (int u = 0; u < G; ++u) pv[u] = nx[u];
      }
    }
    __syncthreads();
  }
}

int launch_bits_tokens(const BitsTables& tb, const uint32_t* words, int64_t words_cs, const float* style,
                       const float* embed_w, const float* embed_b, const float* pos, void* out, int out_dt,
                       int n_clips, int q_lo, int q_hi, int C, cudaStream_t st) {
  if (n_clips <= 0) return AT_OK;
  AT_REQUIRE(tb.T <= MAXT && tb.L <= 256 && q_lo >= 0 && q_hi < tb.n_levels && q_lo <= q_hi && tb.n_levels * tb.T <= 8 * MAXT / 2,
             "bits_tokens: bad levels");
  dim3 grid(n_clips, ceil_div(C, 256));
  if (out_dt == DT_F32)
    AT_CUDA(launch_k(bits_tokens_kernel<float>, dim3(grid), dim3(256), 0, st, tb, words, words_cs, style, embed_w, embed_b, pos, (float*)out, q_lo, q_hi, C));
  else
    AT_CUDA(launch_k(bits_tokens_kernel<bf16>, dim3(grid), dim3(256), 0, st, tb, words, words_cs, style, embed_w, embed_b, pos, (bf16*)out, q_lo, q_hi, C));
  AT_LAUNCH_CHECK();
  return AT_OK;
}

// ---------------------------------------------------------------- decoder latent
template <typename TO>
__global__ void __launch_bounds__(256) bits_latent_kernel(BitsTables tb, const uint32_t* __restrict__ words, int64_t words_cs,
                                                          const float* __restrict__ dec_pos, TO* __restrict__ out, int half) {
  pdl_enter();
  __shared__ uint32_t sw[256];
  __shared__ int s_i0[8 * MAXT / 2], s_i1[8 * MAXT / 2];
  __shared__ float s_w1[8 * MAXT / 2];
  const int clip = blockIdx.x, tid = threadIdx.x, T = tb.T;
  for (int i = tid; i < tb.L; i += blockDim.x) sw[i] = words[(int64_t)clip * words_cs + i];
  for (int i = tid; i < (tb.n_levels - 1) * T; i += blockDim.x) { s_i0[i] = tb.up_i0[i]; s_i1[i] = tb.up_i1[i]; s_w1[i] = tb.up_w1[i]; }
  __syncthreads();
  for (int i = tid; i < T * CD; i += blockDim.x) {
    int t = i >> 5, c = i & 31;
    float f = 0.f;
    for (int lvl = 0; lvl + 1 < tb.n_levels; ++lvl) {
      int src0 = lvl ? tb.cum[lvl - 1] : 0;
      int i0 = s_i0[lvl * T + t], i1 = s_i1[lvl * T + t];
      float w1 = s_w1[lvl * T + t], w0 = 1.0f - w1;
      f += w0 * bit_val(sw[src0 + i0], c) + w1 * bit_val(sw[src0 + i1], c);
    }
    f += bit_val(sw[tb.cum[tb.n_levels - 2] + t], c);
    f += dec_pos[(int64_t)(half * T + t) * CD + c];
    out[((int64_t)clip * 2 * T + half * T + t) * CD + c] = from_f32<TO>(f);
  }
}

int launch_bits_latent(const BitsTables& tb, const uint32_t* words, int64_t words_cs, const float* dec_pos, void* out,
                       int out_dt, int n_clips, int half, cudaStream_t st) {
  if (n_clips <= 0) return AT_OK;
  AT_REQUIRE(tb.L <= 256 && tb.n_levels >= 2 && tb.n_levels * tb.T <= 8 * MAXT / 2, "bits_latent: bad tables");
  if (out_dt == DT_F32) AT_CUDA(launch_k(bits_latent_kernel<float>, dim3(n_clips), dim3(256), 0, st, tb, words, words_cs, dec_pos, (float*)out, half));
  else AT_CUDA(launch_k(bits_latent_kernel<bf16>, dim3(n_clips), dim3(256), 0, st, tb, words, words_cs, dec_pos, (bf16*)out, half));
  AT_LAUNCH_CHECK();
  return AT_OK;
}

// ---------------------------------------------------------------- residual multi-scale BSQ
__global__ void __launch_bounds__(256) bsq_kernel(BitsTables tb, const float* __restrict__ enc_out, uint32_t* __restrict__ words,
                                                  int64_t words_cs) {
  pdl_enter();
  __shared__ float r[MAXT][CD];
  __shared__ float qs[MAXT][CD];
  const int clip = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, T = tb.T;
  const float q_scale = 0.17677669529663687f;
  __shared__ int s_i0[8 * MAXT / 2], s_i1[8 * MAXT / 2], s_ps[8 * MAXT / 2], s_pe[8 * MAXT / 2];
  __shared__ float s_w1[8 * MAXT / 2];
  for (int i = tid; i < tb.n_levels * T; i += blockDim.x) {   // operator tables: one round trip instead of one per element
    s_i0[i] = tb.up_i0[i]; s_i1[i] = tb.up_i1[i]; s_w1[i] = tb.up_w1[i]; s_ps[i] = tb.pool_start[i]; s_pe[i] = tb.pool_end[i];
  }
  for (int i = tid; i < T * CD; i += blockDim.x) (&r[0][0])[i] = enc_out[(int64_t)clip * T * CD + i];
  __syncthreads();
  for (int k = 0; k < tb.n_levels; ++k) {
    const int pt = tb.pn[k], dst0 = k ? tb.cum[k - 1] : 0;
    for (int i = warp; i < pt; i += 8) {
      float a;
      if (pt == T) a = r[i][lane];
      else {
        int s = s_ps[k * T + i], e = s_pe[k * T + i];
        a = 0.f;
        for (int t = s; t < e; ++t) a += r[t][lane];
        a = a / (float)(e - s);
      }
      float nrm = sqrtf(warp_sum(a * a));
      float z = a / fmaxf(nrm, 1e-12f);
      float zhat = (z > 0.f ? 1.0f : -1.0f) * q_scale;
      float qz = z + (zhat - z);                      // bitwise_vae.py:334
      uint32_t wv = __ballot_sync(0xffffffffu, qz > 0.f);
      if (lane == 0) words[(int64_t)clip * words_cs + dst0 + i] = wv;
      qs[i][lane] = qz;
    }
    __syncthreads();
    if (k + 1 < tb.n_levels) {
      for (int i = tid; i < T * CD; i += blockDim.x) {
        int t = i >> 5, c = i & 31;
        float up;
        if (pt == T) up = qs[t][c];
        else {
          int i0 = s_i0[k * T + t], i1 = s_i1[k * T + t];
          float w1 = s_w1[k * T + t], w0 = 1.0f - w1;
          up = w0 * qs[i0][c] + w1 * qs[i1][c];
        }
        r[t][c] -= up;
      }
      __syncthreads();
    }
  }
}

int launch_bsq_quantize(const BitsTables& tb, const float* enc_out, uint32_t* words, int64_t words_cs, int n_clips,
                        cudaStream_t st) {
  if (n_clips <= 0) return AT_OK;
  AT_REQUIRE(tb.T <= MAXT && tb.n_levels * tb.T <= 8 * MAXT / 2, "bsq: T too large");
  AT_CUDA(launch_k(bsq_kernel, dim3(n_clips), dim3(256), 0, st, tb, enc_out, words, words_cs));
  AT_LAUNCH_CHECK();
  return AT_OK;
}

// ---------------------------------------------------------------- encoder input
template <typename TO>
__global__ void __launch_bounds__(256) motion_norm_pos_kernel(const float* __restrict__ motion, const float* __restrict__ mean,
                                                              const float* __restrict__ stdv, const float* __restrict__ pos,
                                                              TO* __restrict__ out, int64_t total, int T, int dim, int k_pad) {
  pdl_enter();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % k_pad);
    int64_t row = i / k_pad;
    int t = (int)(row % T);
    float v = 0.f;
    if (c < dim) v = (motion[row * dim + c] - mean[c]) / stdv[c] + pos[(int64_t)t * dim + c];
    out[i] = from_f32<TO>(v);
  }
}

int launch_motion_norm_pos(const float* motion, const float* mean, const float* stdv, const float* pos, void* out,
                           int out_dt, int n_clips, int T, int dim, int k_pad, cudaStream_t st) {
  if (n_clips <= 0) return AT_OK;
  int64_t total = (int64_t)n_clips * T * k_pad;
  int grid = (int)((total + 255) / 256);
  if (grid > 148 * 16) grid = 148 * 16;
  if (out_dt == DT_F32) AT_CUDA(launch_k(motion_norm_pos_kernel<float>, dim3(grid), dim3(256), 0, st, motion, mean, stdv, pos, (float*)out, total, T, dim, k_pad));
  else AT_CUDA(launch_k(motion_norm_pos_kernel<bf16>, dim3(grid), dim3(256), 0, st, motion, mean, stdv, pos, (bf16*)out, total, T, dim, k_pad));
  AT_LAUNCH_CHECK();
  return AT_OK;
}

}  // namespace artalk
