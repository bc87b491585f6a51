// Host-side launchers of every kernel in the library. All launches are asynchronous on `st`.
// DType: 0 = f32, 1 = bf16 ("act" tensors feeding GEMM A operands are bf16 in bf16 mode).
#pragma once
#include "common.cuh"

namespace artalk {

enum DType : int { DT_F32 = 0, DT_BF16 = 1 };
static inline size_t dt_size(int dt) { return dt == DT_F32 ? 4 : 2; }

// ---------------- norms.cu ----------------
// out[r] = act(LN_eps(x[r]) * gamma + beta); gamma/beta may be null (no affine). cols in {128,512,768,1024}.
int launch_layernorm(const float* x, int64_t x_rs, void* out, int out_dt, int64_t out_rs, const float* gamma,
                     const float* beta, int rows, int cols, float eps, int act, cudaStream_t st);
// out[r] = LN_eps(x[r]) * (1 + ada[map(r)][scale_off + c]) + ada[map(r)][shift_off + c]   (app/transformer.py:35,40)
// With y != null the row is first updated in place: x[r] <- fma(ada[map(r)][gate_off + c], y[r][c], x[r][c]) (the deferred
// gated residual of the previous projection / FFN GEMM, y = [rows, cols] fp32).
int launch_adaln_modulate(float* x, const void* ada, int ada_dt, RowMap ada_map, int scale_off, int shift_off,
                          void* out, int out_dt, int rows, int cols, float eps, cudaStream_t st, const float* y = nullptr,
                          int gate_off = 0);
// stats[chunk] = (mean, 1/(std_unbiased + 1e-6))   (app/modules/wav2vec.py:23-27)
int launch_audio_stats(const float* audio, int n_chunks, int n_samples, float2* stats, cudaStream_t st);
// conv layer 0 (Cin=1,k=10,s=5) on normalised audio + LN(512) + GELU_erf, channels-last output [n, L_out, 512]
int launch_conv0_ln_gelu(const float* audio, const float2* stats, const float* w_kc, const float* bias,
                         const float* ln_g, const float* ln_b, void* out, int out_dt, int n_chunks, int n_samples,
                         int l_out, int kernel, int stride, float eps, cudaStream_t st);
// adaptive average pooling over time to each patch size, concatenated: [n, Lin, C] -> [n, sum(pn), C] (app/models.py:94-95)
int launch_audio_pool(const float* x, float* cond, int n, int l_in, int cols, const int* patch_nums, int n_levels,
                      cudaStream_t st);
// out[r][c] = act(x[map(r)][c]) cast to out_dt
int launch_act_cast(const float* x, RowMap x_map, void* out, int out_dt, int rows, int cols, int act, cudaStream_t st);

// ---------------- gemm ----------------
struct GemmArgs {
  // C[M,N] = A[M,K] * W[N,K]^T ; A rows addressed through a_map, K contiguous
  const void* A; RowMap a_map;
  const void* W; int64_t ldw;          // W row stride (elements)
  int M, N, K;
  // tap mode (grouped temporal conv): K = taps*tap_w ; element (r,k): j=k/tap_w, c=k%tap_w ->
  // A[batch(r)][t(r)+j-tap_pad][c], zero outside [0, a_map.rpb)
  int tap_w, tap_pad;
  // groups: blockIdx.z = g ; per-group element offsets
  int groups; int64_t a_gs, w_gs, c_gs; int bias_gs;
  const float* bias;                    // [N] or null
  int act;                              // applied to acc + bias
  const void* gate; int gate_dt; RowMap gate_map;   // optional per-element multiplier gate[map(r)][c]
  const float* resid; RowMap resid_map;             // optional fp32 addend (may alias out32)
  float* out32; void* out_act; int out_act_dt; RowMap c_map;   // either/both outputs, same row map
  // fused AR q/k/v epilogue (tcgen05 kernel only; app/transformer.py:68-74): output columns are blocks of qkv_C,
  // mode 1 = [q | k | v], mode 2 = per layer [k | v] (period 2*qkv_C); q and k heads (64 wide) are L2-normalised
  // (q additionally scaled by head_scale[head]); q -> qbuf[row][C], k/v -> caches at kv_map(row) (+ layer stride)
  int qkv_mode; int qkv_C; const float* head_scale; void* qbuf; void* kcache; void* vcache; RowMap kv_map;
  int64_t kv_layer_stride;
  // 1: the caller allows the latency kernel of skinny.cu (taken when the shape qualifies). Callers decide by PER-CLIP
  // shapes (tokens per clip of the scale step), never by the batch, so that a clip's arithmetic does not depend on how
  // many clips run with it (batched == per-clip loop stays bit-exact in bf16 mode)
  int skinny;
  // parity-grade mode (split.cu): operands are bf16 piece blocks, K = slots * K_fp32. tap_slots: in tap mode every 64-column
  // block of an A row holds `tap_slots` consecutive piece blocks and k-block kb reads slot kb % tap_slots of tap kb / tap_slots.
  // exact: activations use the libm-accurate functions (erff / tanhf / expf) instead of the bf16-grade fast forms
  int tap_slots; int exact;
  // split_acc = slots (3 / 6): operands are piece blocks and the p0 x p0 products are accumulated apart from the correction
  // products (the tensor core's accumulate truncates: see gemm_tc_kernel<.., SPLIT>); 0 = one accumulator
  int split_acc;
};
static inline GemmArgs gemm_args() {
  GemmArgs g;
  g.A = nullptr; g.a_map = plain_rows(0); g.W = nullptr; g.ldw = 0; g.M = g.N = g.K = 0;
  g.tap_w = 0; g.tap_pad = 0; g.groups = 1; g.a_gs = g.w_gs = g.c_gs = 0; g.bias_gs = 0;
  g.bias = nullptr; g.act = ACT_NONE; g.gate = nullptr; g.gate_dt = DT_F32; g.gate_map = plain_rows(0);
  g.resid = nullptr; g.resid_map = plain_rows(0); g.out32 = nullptr; g.out_act = nullptr; g.out_act_dt = DT_F32;
  g.c_map = plain_rows(0);
  g.qkv_mode = 0; g.qkv_C = 0; g.head_scale = nullptr; g.qbuf = nullptr; g.kcache = nullptr; g.vcache = nullptr;
  g.kv_map = plain_rows(0); g.kv_layer_stride = 0; g.skinny = 0; g.tap_slots = 1; g.exact = 0; g.split_acc = 0;
  return g;
}
// fp32 CUDA-core GEMM (A and W fp32). gemm_simt.cu
int launch_gemm_simt(const GemmArgs& g, cudaStream_t st);
// bf16 tcgen05/TMEM GEMM fed by TMA (A and W bf16). gemm_tc.cu
int launch_gemm_tc(const GemmArgs& g, cudaStream_t st);
// developer switch: 0 = never use the CTA-pair (cta_group::2) kernel for large GEMMs
void set_gemm_pair_mode(int on);
void set_gemm_force_bn(int bn);
void set_gemm_tma_resid(int on);
void set_gemm_band_mb(int mb);
void set_gemm_tma_out(int on);
void set_gemm_resid_deep(int on);
void set_gemm_pair_split(int on);
void set_gemm_pair_min_waves10(int v);
void set_gemm_pair_qkv(int v);
void set_gemm_epi_warps(int v);
// skinny-M bf16 GEMM (mma.sync, cp.async ring; M <= option "skinny_max_m") for the latency-bound steps. skinny.cu
bool gemm_skinny_supported(const GemmArgs& g);
int launch_gemm_skinny(const GemmArgs& g, cudaStream_t st);
void set_skinny_max_m(int v);

// ---------------- conv0_fold.cu ----------------
// bf16 path of conv layer 0 with the LayerNorm folded through the convolution (statistics from the frame's 10 samples, packed
// FFMA2 arithmetic). wq [10][512], bq [512], beta [512], qf [11][12] from weights.conv0_fold; out bf16 [n_chunks][l_out][512]
int launch_conv0_fold(const float* audio, const float2* stats, const float* wq, const float* bq, const float* beta, const float* qf,
                      void* out, int n_chunks, int n_samples, int l_out, float eps, cudaStream_t st);

// ---------------- posconv_tc.cu ----------------
// wav2vec2 positional conv (16 groups x 64 channels, 128 taps, pad 64) + bias + GELU + residual as a CTA-pair tcgen05 GEMM with
// four output frames per A row (N = 256). x [n_chunks][F][H] bf16, w4 [groups][256][(taps + 3) * 64] bf16 (weights.repack
// "w2v.pos.w4"), resid / out [n_chunks * F][H] fp32 (may not alias).
bool posconv4_supported(int F, int H, int groups, int taps);
int launch_posconv4(const void* x, const void* w4, const float* bias, const float* resid, float* out, int n_chunks, int F, int H,
                    int groups, int taps, cudaStream_t st);

// ---------------- split.cu ----------------
// fp32 [n_elems] -> bf16 piece blocks [n_elems / 64][slots][64] (slots 3: 2 pieces / 3 MMA passes, 6: 3 pieces / 6 passes);
// is_w selects the weight-side slot order so that slot s of A times slot s of W enumerates the kept piece products
int launch_split_bf16(const float* x, void* out, int64_t n_elems, int slots, int is_w, cudaStream_t st);
// fp32 rows x[seq][row][0 .. cols) (sequence stride ss, row stride rs) -> two bf16 piece planes out[2][n_seq][rows][cols]
// (hi = bf16(x), lo = bf16(x - hi)): the operands of the parity-grade attention (attention_tc.cu, SPLIT). cols % 8 == 0
int launch_split2_rows(const float* x, int64_t ss, int64_t rs, int n_seq, int rows, int cols, void* out, cudaStream_t st);

// ---------------- attention.cu ----------------
struct AttnArgs {
  const void* q; const void* k; const void* v; void* out;   // dtype dt
  int dt;
  int n_seq, n_heads, head_dim;      // head_dim in {32, 64}
  int lq, lk;
  int64_t q_ss, q_rs;                // sequence stride, row stride (elements); head h at column h*head_dim
  int64_t k_ss, k_rs, v_ss, v_rs;
  int64_t o_ss, o_rs;
  float scale;
  int split;                          // >0: query rows < split only see keys < split (bitwise_vae.py:67-76)
  // optional [n_heads] device array: an upper bound of |q.k| * scale for every row of head h (the AR attention's q and k are
  // L2-normalised per head, app/transformer.py:72-74, so |q.k| <= head_scale[h]). The tcgen05 kernel then subtracts the bound
  // instead of the row maximum (softmax is shift invariant) and skips its max pass over S; ignored by the SIMT kernel
  const float* key_bound = nullptr;
  // parity-grade tensor-core launch (precision "bf16x3"): q / k / v point to the bf16 HI planes of [2][n_seq][rows][width] piece
  // tensors (x = hi + lo, launch_split2_rows; the lo plane of a tensor starts n_seq * its sequence stride later) and `out` is fp32
  int split_planes = 0;
};
// dispatch: bf16 / head_dim 64 / <= 384 keys -> tcgen05 kernel (attention_tc.cu), otherwise the fp32-arithmetic SIMT kernel
int launch_attention(const AttnArgs& a, cudaStream_t st);
bool attention_tc_supported(const AttnArgs& a);
bool attention_tc_split_supported(int lq, int lk, int head_dim);
// fp32-grade attention on the tensor cores (precision "bf16x3"): a.q / a.k / a.v / a.out are fp32 views; `scratch`
// (attention_split_scratch_bytes(a) bytes, 256-byte aligned) receives the two bf16 piece planes of each operand
bool attention_split_supported(const AttnArgs& a);
size_t attention_split_scratch_bytes(const AttnArgs& a);
int launch_attention_split(const AttnArgs& a, void* scratch, cudaStream_t st);
void set_attn_simt_max_lq(int v);
void set_attn_blk(int v);
int launch_attention_tc(const AttnArgs& a, cudaStream_t st);
// AR q/k/v post-processing (app/transformer.py:71-74): per-head L2 normalise q (x exp(min(scale_mul, ln100))) and k,
// q -> qbuf [M, C]; k,v -> cache rows given by kv_map. qkv: [M, 3C] (q | k | v), or [M, 2C] (k | v) when has_q = 0.
int launch_qkv_norm_scatter(const void* qkv, int dt, int64_t qkv_rs, int has_q, const float* head_scale, void* qbuf,
                            void* kcache, void* vcache, RowMap kv_map, int rows, int n_heads, cudaStream_t st);

// ---------------- bits.cu ----------------
struct ScaleOps;   // device tables of the up/down-sampling operators, built by bits_build_tables
struct BitsTables {
  int n_levels; int pn[8]; int cum[8]; int T; int L;
  // linear upsample pn[k] -> T (align_corners = False): for level k and target t: i0, i1, w1
  const int* up_i0; const int* up_i1; const float* up_w1;     // [n_levels][T]
  // area pool T -> pn[k]: window [start, end)
  const int* pool_start; const int* pool_end;                  // [n_levels][T]
};
// logits [rows,64] f32 (row map) -> bit j = (l[2j+1] > l[2j]), one packed word per token (app/models.py:104)
int launch_argmax_bits(const float* logits, RowMap l_map, uint32_t* words, RowMap w_map, int rows, cudaStream_t st);
// tokens from bits (bitwise_vae.py:264-305 + app/models.py:89,100,107,113): for levels q in [q_lo, q_hi]:
//   q = 0 : style ; q >= 1 : embed(A_{pn[q]} f_{q-1})      (+ pos[cum_{q-1} + i])
// written to out rows (clip, row_off + i) with row_off = (cum_{q-1} - cum_{q_lo - 1}).
int launch_bits_tokens(const BitsTables& tb, const uint32_t* words, int64_t words_cs, const float* style,
                       const float* embed_w, const float* embed_b, const float* pos, void* out, int out_dt,
                       int n_clips, int q_lo, int q_hi, int C, cudaStream_t st);
// decoder latent (bitwise_vae.py:280-288) + dec_pos_embed rows [half*T, half*T+T): out[clip][half*T + t][32]
int launch_bits_latent(const BitsTables& tb, const uint32_t* words, int64_t words_cs, const float* dec_pos, void* out,
                       int out_dt, int n_clips, int half, cudaStream_t st);
// residual multi-scale BSQ (bitwise_vae.py:227-242,316-334): enc_out [n,T,32] f32 -> words [n, L]
int launch_bsq_quantize(const BitsTables& tb, const float* enc_out, uint32_t* words, int64_t words_cs, int n_clips,
                        cudaStream_t st);
// (motion - mean)/std + enc_pos, zero padded to k_pad columns (bitwise_vae.py:88-89)
int launch_motion_norm_pos(const float* motion, const float* mean, const float* stdv, const float* pos, void* out,
                           int out_dt, int n_clips, int T, int dim, int k_pad, cudaStream_t st);

// ---------------- flame.cu ----------------
struct FlameModel {
  int V, n_shape, n_exp;                  // 5023, 300, 100  (+36 pose-corrective bases)
  const float* v_template;                // [V*3]
  const float* dirs;                      // [n_shape + n_exp + 36][V*3]  blend bases, basis-major
  const float* j_template;                // [5*3]   J_regressor @ v_template
  const float* j_dirs;                    // [n_shape + n_exp][15]  J_regressor @ shapedirs
  const float* lbs_weights;               // [V][5]
  int parents[5];
  float scale;
  // optional tensor-core operands (flame_tc.cu): dirs^T split into bf16 hi/lo, [V*3][3*KS] = [hi | hi | lo];
  // *_full covers all n_shape+n_exp+36 bases, *_expr the last n_exp+36 (shared shape row). null -> fp32 CUDA-core kernel
  const void* bsplit_full; int ks_full;
  const void* bsplit_expr; int ks_expr;
};
int launch_flame_tc(const FlameModel& m, const float* base, const float* coef, int coef_stride, int l_begin, int n_l,
                    const void* b_split, int KS, void* a_split_ws, float* verts, int n_frames, cudaStream_t st);
// shape (N,300) [stride 0 allowed for a shared shape row], expr (N,100), pose6 (N,6) -> verts (N,V,3)
int launch_flame(const FlameModel& fm, const float* shape, int64_t shape_rs, const float* expr, int64_t expr_rs,
                 const float* pose, int64_t pose_rs, int zero_global, float* coef_ws, float* verts, int n_frames,
                 cudaStream_t st);
size_t flame_workspace_floats(const FlameModel& fm, int n_frames);

// ---------------- postproc.cu ----------------
// Savitzky-Golay (win 5/poly 2; dims 100:103 win 9/poly 3, mode 'interp') + clip + pose/eye zeroing (inference.py:52-56,89-95)
int launch_savgol_post(const float* motion, float* out, int n_clips, int T, int T_out, int dim, int fix_pose,
                       int zero_tail, cudaStream_t st);

// forehead EMA scan of the GAGAvatar point builder (app/GAGAvatar/models.py:120-125): points [n_frames][V][3] updated in place
int launch_ema_scan(float* points, int64_t frame_stride, const int* idx, int n_idx, int n_frames, float* state, int has_state,
                    float keep, cudaStream_t st);

// ---------------- mesh.cu ----------------
// area-weighted per-vertex normals (pytorch3d Meshes.verts_normals semantics) from a host-built CSR vertex adjacency
int launch_vertex_normals(const float* verts, int64_t frame_stride, int V, const int* adj_off, const int* adj_pair, float* normals,
                          int n_frames, cudaStream_t st);

// ---------------- frontend.cu ----------------
// torchaudio-style polyphase sinc resampling + channel mean (inference.py:112-113,230-231): in [channels][length] (channel
// stride ch_stride), bank [new][taps] with taps = 2*width + orig, out [out_len <= ceil(new*length/orig)]
int launch_resample_mix(const float* in, int channels, int64_t ch_stride, int64_t length, const float* bank, int orig, int new_f,
                        int taps, int width, float* out, int64_t out_len, cudaStream_t st);

}  // namespace artalk
