/* artalk_b200 — C ABI of the B200-native ARTalk audio->motion path (libartalk_b200.so).
 *
 * The reference (zsc/ARTalk) has no native code, plugin registry or FFI: its boundary for this path is the Python
 * surface ARTAvatarInferEngine / BitwiseARModel.inference / BITWISE_VAE.get_flame_verts / FLAMEModel.forward. The Python
 * mirror of that surface lives in artalk_b200/{engine,model,flame}.py and binds exactly the entry points declared here
 * with ctypes; each entry point cites the reference code it replaces.
 *
 * Conventions: plain pointers and sizes, no C++/torch types; every pointer is a DEVICE pointer unless named host_*;
 * tensors are dense row-major; `stream` is a cudaStream_t passed as void* (0 = legacy default stream); calls are
 * asynchronous on that stream; return value 0 = ok, otherwise an ARTALK_E* code with text in artalk_last_error().
 * An engine is not thread-safe; different engines (also on different devices of one process) may be used concurrently:
 * per-device state is kept per device ordinal and the launch counter is atomic. Every call must be made with the engine's
 * device current (cudaSetDevice / torch.cuda.device), as for any CUDA library that takes a stream. artalk_set_option /
 * artalk_enable_pdl are process-wide developer switches: set them while no other thread is launching. Tensor memory handed
 * to artalk_set_tensor stays owned by the caller and must outlive the engine.
 */
#ifndef ARTALK_B200_H
#define ARTALK_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ARTALK_OK 0
#define ARTALK_EINVAL 1    /* bad argument / unsupported shape */
#define ARTALK_ECUDA 2     /* CUDA runtime error */
#define ARTALK_ENOMEM 3    /* workspace allocation failed */
#define ARTALK_EMISSING 4  /* artalk_finalize: a required tensor was not provided (strict load) */
#define ARTALK_ESTATE 5

#define ARTALK_F32 0
#define ARTALK_BF16 1
#define ARTALK_I32 2

#define ARTALK_PRECISION_FP32 0 /* CUDA-core fp32 GEMMs; parity tolerance 1e-3 */
#define ARTALK_PRECISION_BF16 1 /* tcgen05 bf16 GEMMs, fp32 accumulate; parity tolerance 2e-2 */
/* parity-grade tensor-core modes: fp32 data flow (weights handed over in fp32), every GEMM operand split into 2 / 3 bf16
 * pieces and multiplied on the tcgen05 pipe in 3 / 6 passes with fp32 accumulation; attention, norms and activations in fp32.
 * Same tolerance contract as FP32 (bits exact where the logit margin > 1e-3, motion / vertices within 1e-3). */
#define ARTALK_PRECISION_BF16X3 2
#define ARTALK_PRECISION_BF16X6 3

typedef struct artalk_engine artalk_engine_t;

/* assets/config.json:1-15 + the constants hard-coded at app/models.py:19,22,27,37,41-42 + the XLS-R-300m Wav2Vec2Config */
typedef struct artalk_config {
  int precision;
  int ar_depth, ar_heads, embed_dim, cond_dim;
  int vae_depth, vae_heads, vae_hidden, code_dim, motion_dim;
  int n_levels;
  int patch_nums[8];
  int w2v_layers, w2v_heads, w2v_hidden, w2v_ffn, w2v_conv_dim, w2v_n_conv;
  int w2v_conv_kernel[8];
  int w2v_conv_stride[8];
  int w2v_pos_kernel, w2v_pos_groups;
  int style_dim, style_layers, style_heads, style_ffn, style_len;
  int chunk_samples;
  float w2v_ln_eps;
} artalk_config_t;

const char* artalk_last_error(void);
int artalk_abi_version(void);

/* --- engine life cycle: replaces BitwiseARModel.__init__ + load_state_dict(strict=True) (inference.py:24-28) --- */
int artalk_create(const artalk_config_t* cfg, artalk_engine_t** out);
int artalk_destroy(artalk_engine_t* e);
/* register one repacked weight / table by canonical name (see artalk_b200/weights.py for the name list) */
int artalk_set_tensor(artalk_engine_t* e, const char* name, void* ptr, int dtype, int64_t numel);
/* strict check that every tensor the path needs is present with the right dtype and size */
int artalk_finalize(artalk_engine_t* e);
/* soft budget (bytes) used to size wav2vec sub-batches; the library grows its device workspace on demand */
int artalk_set_workspace_limit(artalk_engine_t* e, size_t bytes);
size_t artalk_workspace_bytes(const artalk_engine_t* e);
/* the body of artalk_ar_chunk is replayed from a CUDA graph after one eager warm-up per (n_clips, teacher forcing);
 * enable = 0 drops the graphs and launches eagerly (default: enabled) */
int artalk_enable_graphs(artalk_engine_t* e, int enable);
/* graphs currently instantiated / replays so far (either pointer may be NULL). Returns ARTALK_ESTATE, with the reason in
 * artalk_last_error(), if a capture or instantiation failed and the engine fell back to eager launches (it also says so once
 * on stderr); artalk_enable_graphs(e, 1) re-arms capture. */
int artalk_graph_status(const artalk_engine_t* e, int* n_graphs, int* n_replays);
/* latency mode for batch-1 / few-clip streaming (BASELINE configs[4]; the reference's per-chunk loop, app/models.py:76-115,
 * run one chunk at a time): max_rows > 0 sends every bf16 GEMM of this engine with at most max_rows rows (and a small output)
 * to the latency kernel (skinny.cu: 64-row tiles, ~100 CTAs streaming disjoint weight slices) instead of the 128-row tcgen05
 * tiles; 0 = throughput mode (default), where kernel choice depends only on per-clip shapes so a clip's arithmetic does not
 * depend on the batch. The two modes differ at bf16 rounding level (fp32 accumulation order). Drops the captured graphs. */
int artalk_set_latency_mode(artalk_engine_t* e, int max_rows);

/* --- Wav2Vec2Model.forward + multi-scale area pooling (app/modules/wav2vec.py:11-27, app/models.py:93-95) ---
 * audio [n_chunks, chunk_samples] f32 (each row normalised on its own) -> cond [n_chunks, 181, 1024] f32 */
int artalk_audio_encode(artalk_engine_t* e, const float* audio, int n_chunks, float* cond, void* stream);

/* --- StyleEncoder + style_cond_embed + 1.1/-0.1 mix (app/modules/style_encoder.py:26-42, app/models.py:67-70) ---
 * style_motion [n_clips, 50, 106] f32 -> style [n_clips, 768] f32 */
int artalk_style_encode(artalk_engine_t* e, const float* style_motion, int n_clips, float* style, void* stream);

/* --- BITWISE_VAE.quant_to_vqidx(prev, None) (app/modules/bitwise_vae.py:78-93) ---
 * motion [n_clips, 100, 106] f32 -> words [n_clips, 181] u32 (bit j of a word = code dim j); enc_out (nullable)
 * receives the encoder output [n_clips, 100, 32] f32 */
int artalk_motion_to_bits(artalk_engine_t* e, const float* motion, int n_clips, uint32_t* words, float* enc_out, void* stream);

/* --- BITWISE_VAE.vqidx_to_motion (app/modules/bitwise_vae.py:105-113): returns the new half [n_clips, 100, 106] --- */
int artalk_bits_to_motion(artalk_engine_t* e, const uint32_t* prev_words, const uint32_t* words, int n_clips, float* motion,
                          void* stream);

/* --- one chunk of BitwiseARModel.inference's loop body (app/models.py:93-114) for n_clips clips at once ---
 * cond: clip b's [181,1024] block at cond + b*cond_clip_stride ; style [n_clips,768] ;
 * prev_words [n_clips,181] in: re-encoded bits of the previous chunk, out: those of this chunk ;
 * motion_out [n_clips,100,106] ; words_out / logits_out [n_clips,181,64] / enc_out nullable ;
 * forced_words (nullable) [n_clips,181]: teacher forcing — next-scale inputs and the decode use these bits. */
int artalk_ar_chunk(artalk_engine_t* e, int n_clips, const float* cond, int64_t cond_clip_stride, const float* style,
                    uint32_t* prev_words, float* motion_out, uint32_t* words_out, float* logits_out,
                    const uint32_t* forced_words, float* enc_out, void* stream);

/* --- FLAMEModel.forward / lbs (app/flame_model/FLAME.py:117-149, app/flame_model/lbs.py:142-383) --- */
typedef struct artalk_flame_model {
  int n_verts, n_shape, n_exp;
  const float* v_template;   /* [V*3] */
  const float* dirs;         /* [n_shape+n_exp+36][V*3]: shape, expression then pose-corrective bases */
  const float* j_template;   /* [15]  J_regressor @ v_template */
  const float* j_dirs;       /* [n_shape+n_exp][15] */
  const float* lbs_weights;  /* [V][5] */
  int parents[5];
  float scale;
  /* optional tensor-core operands: dirs^T split into bf16 hi/lo, [V*3][3*ks] = [hi | hi | lo] (zero padded to ks, a
   * multiple of 64). bsplit_full: all n_shape+n_exp+36 bases; bsplit_expr: the last n_exp+36 (used when all frames share
   * one shape row). NULL selects the fp32 CUDA-core kernel. */
  const void* bsplit_full; int ks_full;
  const void* bsplit_expr; int ks_expr;
} artalk_flame_model_t;
size_t artalk_flame_workspace_floats(const artalk_flame_model_t* fm, int n_frames);
/* shape [n, n_shape] (row stride shape_stride, 0 = one shared row), expr [n, n_exp], pose [n, 6] = (global rot, jaw);
 * zero_global != 0 drops the global rotation (get_flame_verts with_global=False, bitwise_vae.py:45-46) */
int artalk_flame_vertices(const artalk_flame_model_t* fm, const float* shape, int64_t shape_stride, const float* expr,
                          int64_t expr_stride, const float* pose, int64_t pose_stride, int zero_global, float* workspace,
                          float* verts, int n_frames, void* stream);

/* --- smooth_motion_savgol + [:clip_length] + pose / eye zeroing (inference.py:52-56,89-95) ---
 * host_h5 [5*5], host_h9 [9*9]: hat matrices of the window-5/poly-2 and window-9/poly-3 fits (host pointers) */
int artalk_set_savgol_tables(const float* host_h5, const float* host_h9);
/* fix_pose: zero dims 100:103 (inference.py:53-54); zero_tail: zero dims 104:106 (inference.py:56) */
int artalk_smooth_motion(const float* motion, float* out, int n_clips, int n_frames, int n_frames_out, int fix_pose,
                         int zero_tail, void* stream);

/* Audio front-end of the callers (inference.py:112-113,230-231: torchaudio.transforms.Resample(sr, 16000)(audio).mean(dim=0)):
 * polyphase windowed-sinc resampling fused with the channel mean. `in` is [channels][length] fp32 on the device (channel
 * stride in elements), `bank` the [new][taps] filter bank of torchaudio's _get_sinc_resample_kernel for the gcd-reduced
 * rates orig/new (taps = 2*width + orig; see artalk_b200/audio.py::sinc_resample_bank), `out` [out_len] with
 * out_len <= ceil(new * length / orig). */
int artalk_resample_mono(const float* in, int channels, int64_t ch_stride, int64_t length, const float* bank, int orig, int new_f,
                         int taps, int width, float* out, int64_t out_len, void* stream);

/* Second FLAME consumer (app/GAGAvatar/models.py:98-128, build_forward_batch): after FLAME decode with scale 5.0 and the
 * avatar's shape code, the forehead vertices follow an exponential moving average across frames (models.py:120-125). `points`
 * is [n_frames][V][3] (frame stride in floats), `idx` the forehead vertex indices on the device, `state` [n_idx][3] carried
 * across calls (has_state = 0 on the first call: the first frame initialises it unblended), keep = 0.98. In place. */
int artalk_ema_scan(float* points, int64_t frame_stride, const int* idx, int n_idx, int n_frames, float* state, int has_state,
                    float keep, void* stream);

/* Vertex normals for a mesh rasteriser fed with the decoded vertices (the reference builds pytorch3d Meshes(verts, faces),
 * app/flame_model/renderer_utils.py; FLAMEModel.get_faces, app/flame_model/FLAME.py:68-69): area-weighted sum of the incident
 * faces' normals cross(P_next - P_v, P_prev - P_v), normalised with eps 1e-6 (pytorch3d verts_normals semantics). The
 * vertex -> incident-face adjacency is CSR: adj_offsets [n_verts + 1], adj_pairs [2 * adj_offsets[n_verts]] = the (next, prev)
 * vertex of every incidence in the face's cyclic order (artalk_b200/flame.py::vertex_adjacency builds it from the faces).
 * verts / normals [n_frames][n_verts][3] fp32 (frame_stride of verts in floats). */
int artalk_vertex_normals(const float* verts, int64_t frame_stride, int n_verts, const int* adj_offsets, const int* adj_pairs,
                          float* normals, int n_frames, void* stream);

/* --- measurement hooks (bench.py) ---
 * artalk_launch_count: kernels launched by this library in this process so far.
 * artalk_profile_enable(e, 1): bracket every GEMM / attention launch of the engine with CUDA events on the launching
 * stream; artalk_profile_read synchronises and fills host_out16[16] (16 doubles) = {gemm launches, gemm ms, gemm flops,
 * attention launches, attention ms, attention flops, then the GEMM shape with the largest summed duration: launches, total
 * ms, flops per launch, M, N, K, 0...} accumulated since the last enable call. */
unsigned long long artalk_launch_count(void);
/* programmatic dependent launch (default on): kernels are launched with the programmatic-stream-serialization attribute so
 * a kernel's prologue (barrier init, TMEM allocation, weight prefetch) overlaps the tail of its predecessor; 0 = plain
 * stream order. Process-wide; takes effect for launches (and CUDA-graph captures) made after the call. */
int artalk_enable_pdl(int enable);
/* process-wide tuning switches (developer / A-B measurement): "pdl" (0/1), "gemm_pair" (0 = never use the CTA-pair
 * cta_group::2 GEMM kernel). Unknown names return an error. */
int artalk_set_option(const char* name, int value);
/* launch trace: between begin and end every kernel launch of the library records a CUDA event on its stream;
 * artalk_trace_end synchronises and writes "launcher,d0,d1,d2,microseconds" lines (time since the previous launch's
 * completion) into host_buf; returns the byte count (a value >= cap means the buffer was too small), -1 on error */
int artalk_trace_begin(void* stream);
long artalk_trace_end(char* host_buf, long cap, void* stream);
int artalk_profile_enable(artalk_engine_t* e, int enable);
int artalk_profile_read(artalk_engine_t* e, double* host_out16, void* stream);

/* ---------------- operator-level entry points (unit parity tests of single kernels) ---------------- */
typedef struct artalk_rowmap { int rpb; int64_t bs, rs; } artalk_rowmap_t;
typedef struct artalk_gemm {
  const void* A; artalk_rowmap_t a_map;
  const void* W; int64_t ldw;
  int M, N, K;
  int tap_w, tap_pad;
  int groups; int64_t a_gs, w_gs, c_gs; int bias_gs;
  const float* bias;
  int act;                       /* 0 none, 1 gelu(erf), 2 gelu(tanh), 3 leaky_relu(0.2), 4 silu */
  const void* gate; int gate_dt; artalk_rowmap_t gate_map;
  const float* resid; artalk_rowmap_t resid_map;
  float* out32; void* out_act; int out_act_dt; artalk_rowmap_t c_map;
  int tap_slots;                 /* bf16 kernel, tap mode on piece blocks: slots per 64-column block (0 / 1 = plain operands) */
  int exact;                     /* bf16 kernel: 1 = libm-accurate activations (parity-grade mode) */
  int split_acc;                 /* bf16 kernel on piece blocks: slots (3 / 6) -> the p0 x p0 products get their own accumulator */
} artalk_gemm_t;
/* out = resid + gate * act(A W^T + bias); precision selects the fp32 CUDA-core or the bf16 tcgen05 kernel */
int artalk_op_gemm(const artalk_gemm_t* g, int precision, void* stream);
/* operand splitting of the parity-grade modes: x [n] f32 -> out [n / 64][slots][64] bf16 piece blocks (slots = 3 or 6; n a
 * multiple of 64; is_w selects the weight-side slot order). A bf16 artalk_op_gemm on split A and W with K' = slots * K (row
 * strides multiplied alike) then computes A W^T to ~2^-17 (3) / ~2^-24 (6) relative per product. */
int artalk_op_split_bf16(const float* x, void* out, int64_t n, int slots, int is_w, void* stream);

typedef struct artalk_attn {
  const void* q; const void* k; const void* v; void* out;
  int dt, n_seq, n_heads, head_dim, lq, lk;
  int64_t q_ss, q_rs, k_ss, k_rs, v_ss, v_rs, o_ss, o_rs;
  float scale;
  int split;
  const float* key_bound;        /* optional [n_heads] device array: |q.k| * scale <= key_bound[h] (bf16 tensor-core kernel: one pass) */
} artalk_attn_t;
int artalk_op_attention(const artalk_attn_t* a, void* stream);
/* fp32-grade attention on the tensor cores (precision "bf16x3"): q / k / v / out fp32 (dt = ARTALK_F32), head_dim 64, 17..368 keys.
 * Operands are split into two bf16 pieces each in `scratch` (256-byte aligned device memory,
 * 4 * n_seq * (lq + 2 * lk) * n_heads * 64 bytes + 1 KiB is enough) and both contractions keep the three piece products. */
int artalk_op_attention_split(const artalk_attn_t* a, void* scratch, size_t scratch_bytes, void* stream);
int artalk_op_layernorm(const float* x, void* out, int out_dt, const float* gamma, const float* beta, int rows, int cols,
                        float eps, int act, void* stream);
/* wav2vec2 feature-extractor layer 0 (modeling_wav2vec2.py:291-299: Conv1d(1, 512, k=10, s=5) + LayerNorm + GELU) on per-chunk
 * normalised audio (app/modules/wav2vec.py:23-27): audio [n_chunks][n_samples] f32 -> out [n_chunks][(n_samples-10)/5+1][512].
 * wq == NULL: direct form (w_kc [10][512], bias, ln_g, ln_b; out f32 or bf16). wq != NULL: bf16 path with the LayerNorm folded
 * through the conv (wq [10][512], bq [512], qf [11][12] from artalk_b200.weights.conv0_fold; ln_b; bf16 out).
 * stats_ws: 2 * n_chunks floats of scratch. */
int artalk_op_conv0(const float* audio, int n_chunks, int n_samples, const float* w_kc, const float* bias, const float* ln_g,
                    const float* ln_b, const float* wq, const float* bq, const float* qf, float* stats_ws, void* out, int out_dt,
                    float eps, void* stream);
/* wav2vec2 positional conv of the bf16 path (transformers modeling_wav2vec2.py:360-368,764-765: grouped Conv1d k=128 pad 64,
 * last output dropped, + bias, GELU, + residual) in its four-frames-per-row tensor-core form: x [n_chunks][frames][hidden]
 * bf16, w4 [groups][256][(taps + 3) * 64] bf16 (artalk_b200.weights.posconv_shift4), bias [hidden], resid / out
 * [n_chunks * frames][hidden] f32 (distinct buffers). hidden = groups * 64, taps = 128. */
int artalk_op_posconv4(const void* x, const void* w4, const float* bias, const float* resid, float* out, int n_chunks, int frames,
                       int hidden, int groups, int taps, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ARTALK_B200_H */
