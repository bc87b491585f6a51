// extern "C" surface of libartalk_b200.so (include/artalk_b200.h).
#include <cstring>
#include <new>
#include "../../include/artalk_b200.h"
#include "engine.cuh"

using namespace artalk;

namespace artalk { const char* last_error(); int trace_begin(cudaStream_t); long trace_end(char*, long, cudaStream_t); }

struct artalk_engine { Engine eng; };

static_assert(sizeof(artalk_config_t) == sizeof(EngineConfig), "artalk_config_t / EngineConfig layout mismatch");

static inline RowMap rm(const artalk_rowmap_t& m) { RowMap r; r.rpb = m.rpb; r.bs = m.bs; r.rs = m.rs; return r; }

extern "C" {

const char* artalk_last_error(void) { return last_error(); }
int artalk_abi_version(void) { return 2; }

int artalk_create(const artalk_config_t* cfg, artalk_engine_t** out) {
  AT_REQUIRE(cfg && out, "artalk_create: null argument");
  artalk_engine* e = new (std::nothrow) artalk_engine();
  if (!e) return AT_ENOMEM;
  std::memcpy(&e->eng.cfg, cfg, sizeof(EngineConfig));
  // wav2vec sub-batch budget: 45 % of the device's memory, at most 80 GiB (a B200 then encodes the 2048 chunks of 256 x 30 s in one
  // sub-batch of 66 GB instead of three: 577 -> 571 ms per step); the arena itself only grows to what a call needs
  // (the piece-block modes keep the 24 GiB default: their operand-split buffers grow with the sub-batch, 49 GB for conv layer 1 at
  // 1250 chunks)
  size_t free_b = 0, total_b = 0;
  if (e->eng.cfg.precision >= 2) {
  } else if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && total_b > 0) {
    const size_t cap = (size_t)80 << 30, share = (size_t)((double)total_b * 0.45);
    e->eng.ws_limit = share < cap ? share : cap;
  } else {
    cudaGetLastError();
  }
  *out = e;
  return AT_OK;
}

int artalk_destroy(artalk_engine_t* e) {
  if (!e) return AT_OK;
  e->eng.drop_graphs();
  e->eng.free_split();
  if (e->eng.split_buf) cudaFree(e->eng.split_buf);
  if (e->eng.asplit_buf) cudaFree(e->eng.asplit_buf);
  if (e->eng.ws) cudaFree(e->eng.ws);
  if (e->eng.gstream) cudaStreamDestroy(e->eng.gstream);
  if (e->eng.gev_in) cudaEventDestroy(e->eng.gev_in);
  if (e->eng.gev_out) cudaEventDestroy(e->eng.gev_out);
  for (auto& ev : e->eng.prof_ev) cudaEventDestroy(ev);
  delete e;
  return AT_OK;
}

int artalk_set_tensor(artalk_engine_t* e, const char* name, void* ptr, int dtype, int64_t numel) {
  AT_REQUIRE(e, "null engine");
  return e->eng.set_tensor(name, ptr, dtype, numel);
}
int artalk_finalize(artalk_engine_t* e) {
  AT_REQUIRE(e, "null engine");
  return e->eng.finalize();
}
int artalk_set_workspace_limit(artalk_engine_t* e, size_t bytes) {
  AT_REQUIRE(e && bytes >= ((size_t)64 << 20), "workspace limit must be >= 64 MiB");
  e->eng.ws_limit = bytes;
  return AT_OK;
}
size_t artalk_workspace_bytes(const artalk_engine_t* e) { return e ? e->eng.ws_cap : 0; }
int artalk_enable_graphs(artalk_engine_t* e, int enable) {
  AT_REQUIRE(e, "null engine");
  if (!enable) e->eng.drop_graphs();
  else e->eng.graph_failure.clear();
  e->eng.use_graphs = enable != 0;
  return AT_OK;
}

int artalk_graph_status(const artalk_engine_t* e, int* n_graphs, int* n_replays) {
  AT_REQUIRE(e, "null engine");
  int n = 0;
  for (auto& kv : e->eng.graphs) n += kv.second.exec != nullptr;
  if (n_graphs) *n_graphs = n;
  if (n_replays) *n_replays = e->eng.graph_replays;
  if (!e->eng.graph_failure.empty()) {
    set_last_error("%s", e->eng.graph_failure.c_str());
    return AT_ESTATE;
  }
  return AT_OK;
}

int artalk_set_latency_mode(artalk_engine_t* e, int max_rows) {
  AT_REQUIRE(e && max_rows >= 0, "artalk_set_latency_mode: bad argument");
  if (e->eng.latency_rows != max_rows) e->eng.drop_graphs();
  e->eng.latency_rows = max_rows;
  return AT_OK;
}

int artalk_audio_encode(artalk_engine_t* e, const float* audio, int n_chunks, float* cond, void* stream) {
  AT_REQUIRE(e && audio && cond && n_chunks >= 0, "artalk_audio_encode: bad argument");
  return e->eng.audio_encode(audio, n_chunks, cond, (cudaStream_t)stream);
}

int artalk_style_encode(artalk_engine_t* e, const float* style_motion, int n_clips, float* style, void* stream) {
  AT_REQUIRE(e && style_motion && style && n_clips >= 0, "artalk_style_encode: bad argument");
  return e->eng.style_encode(style_motion, n_clips, style, (cudaStream_t)stream);
}

int artalk_motion_to_bits(artalk_engine_t* e, const float* motion, int n_clips, uint32_t* words, float* enc_out, void* stream) {
  AT_REQUIRE(e && motion && words && n_clips >= 0, "artalk_motion_to_bits: bad argument");
  AT_REQUIRE(e->eng.finalized, "engine not finalized");
  Engine& g = e->eng;
  const size_t s = dt_size(g.act_dt());
  size_t need = (size_t)n_clips * g.T * (128 * s + g.cfg.vae_hidden * (4 + 7 * s) + g.cfg.code_dim * 4) + (2 << 20);
  AT_TRY(g.ws_reserve(need, (cudaStream_t)stream));
  return g.vae_encode_bits(motion, n_clips, words, enc_out, (cudaStream_t)stream);
}

int artalk_bits_to_motion(artalk_engine_t* e, const uint32_t* prev_words, const uint32_t* words, int n_clips, float* motion,
                          void* stream) {
  AT_REQUIRE(e && prev_words && words && motion && n_clips >= 0, "artalk_bits_to_motion: bad argument");
  AT_REQUIRE(e->eng.finalized, "engine not finalized");
  Engine& g = e->eng;
  const size_t s = dt_size(g.act_dt());
  size_t need = (size_t)n_clips * 2 * g.T * (g.cfg.code_dim * s + g.cfg.vae_hidden * (4 + 7 * s)) + (2 << 20);
  AT_TRY(g.ws_reserve(need, (cudaStream_t)stream));
  return g.vae_decode(prev_words, words, n_clips, motion, (cudaStream_t)stream);
}

int artalk_ar_chunk(artalk_engine_t* e, int n_clips, const float* cond, int64_t cond_clip_stride, const float* style,
                    uint32_t* prev_words, float* motion_out, uint32_t* words_out, float* logits_out,
                    const uint32_t* forced_words, float* enc_out, void* stream) {
  AT_REQUIRE(e && cond && style && prev_words && motion_out && n_clips >= 0, "artalk_ar_chunk: bad argument");
  return e->eng.ar_chunk(n_clips, cond, cond_clip_stride, style, prev_words, motion_out, words_out, logits_out, forced_words,
                         enc_out, (cudaStream_t)stream);
}

static FlameModel to_flame(const artalk_flame_model_t* f) {
  FlameModel m;
  m.V = f->n_verts; m.n_shape = f->n_shape; m.n_exp = f->n_exp; m.v_template = f->v_template; m.dirs = f->dirs;
  m.j_template = f->j_template; m.j_dirs = f->j_dirs; m.lbs_weights = f->lbs_weights;
  for (int i = 0; i < 5; ++i) m.parents[i] = f->parents[i];
  m.scale = f->scale;
  m.bsplit_full = f->bsplit_full; m.ks_full = f->ks_full; m.bsplit_expr = f->bsplit_expr; m.ks_expr = f->ks_expr;
  return m;
}

size_t artalk_flame_workspace_floats(const artalk_flame_model_t* fm, int n_frames) {
  if (!fm || n_frames < 0) return 0;
  return flame_workspace_floats(to_flame(fm), n_frames);
}

int artalk_flame_vertices(const artalk_flame_model_t* fm, const float* shape, int64_t shape_stride, const float* expr,
                          int64_t expr_stride, const float* pose, int64_t pose_stride, int zero_global, float* workspace,
                          float* verts, int n_frames, void* stream) {
  AT_REQUIRE(fm && shape && expr && pose && workspace && verts && n_frames >= 0, "artalk_flame_vertices: bad argument");
  AT_REQUIRE(fm->parents[0] < 0, "flame: joint 0 must be the root");
  for (int j = 1; j < 5; ++j) AT_REQUIRE(fm->parents[j] >= 0 && fm->parents[j] < j, "flame: parents must precede children");
  return launch_flame(to_flame(fm), shape, shape_stride, expr, expr_stride, pose, pose_stride, zero_global, workspace, verts,
                      n_frames, (cudaStream_t)stream);
}

int artalk_set_savgol_tables(const float* host_h5, const float* host_h9) {
  AT_REQUIRE(host_h5 && host_h9, "null tables");
  set_savgol_tables(host_h5, host_h9);
  return AT_OK;
}

int artalk_smooth_motion(const float* motion, float* out, int n_clips, int n_frames, int n_frames_out, int fix_pose,
                         int zero_tail, void* stream) {
  AT_REQUIRE(motion && out, "artalk_smooth_motion: null argument");
  return launch_savgol_post(motion, out, n_clips, n_frames, n_frames_out, 106, fix_pose, zero_tail, (cudaStream_t)stream);
}

int artalk_resample_mono(const float* in, int channels, int64_t ch_stride, int64_t length, const float* bank, int orig, int new_f,
                         int taps, int width, float* out, int64_t out_len, void* stream) {
  return launch_resample_mix(in, channels, ch_stride, length, bank, orig, new_f, taps, width, out, out_len, (cudaStream_t)stream);
}

int artalk_ema_scan(float* points, int64_t frame_stride, const int* idx, int n_idx, int n_frames, float* state, int has_state,
                    float keep, void* stream) {
  return launch_ema_scan(points, frame_stride, idx, n_idx, n_frames, state, has_state, keep, (cudaStream_t)stream);
}

int artalk_vertex_normals(const float* verts, int64_t frame_stride, int n_verts, const int* adj_offsets, const int* adj_pairs,
                          float* normals, int n_frames, void* stream) {
  return launch_vertex_normals(verts, frame_stride, n_verts, adj_offsets, adj_pairs, normals, n_frames, (cudaStream_t)stream);
}

unsigned long long artalk_launch_count(void) { return g_launch_count.load(); }
int artalk_enable_pdl(int enable) { g_pdl = enable != 0; ++g_option_epoch; return AT_OK; }
int artalk_set_option(const char* name, int value) {
  AT_REQUIRE(name, "artalk_set_option: null name");
  ++g_option_epoch;                          // kernel selection may change: chunk graphs captured so far are re-captured
  if (!std::strcmp(name, "pdl")) { g_pdl = value != 0; return AT_OK; }
  if (!std::strcmp(name, "pdl_mask")) { g_pdl_mask = value; return AT_OK; }
  if (!std::strcmp(name, "pdl_w2v_max_chunks")) { g_pdl_w2v_max_chunks = value; return AT_OK; }
  if (!std::strcmp(name, "gemm_pair")) { set_gemm_pair_mode(value); return AT_OK; }
  if (!std::strcmp(name, "gemm_force_bn")) { set_gemm_force_bn(value); return AT_OK; }
  if (!std::strcmp(name, "gemm_resid_deep")) { set_gemm_resid_deep(value); return AT_OK; }
  if (!std::strcmp(name, "gemm_pair_split")) { set_gemm_pair_split(value); return AT_OK; }
  if (!std::strcmp(name, "gemm_pair_min_waves10")) { set_gemm_pair_min_waves10(value); return AT_OK; }
  if (!std::strcmp(name, "gemm_pair_qkv")) { set_gemm_pair_qkv(value); return AT_OK; }
  if (!std::strcmp(name, "gemm_epi_warps")) { set_gemm_epi_warps(value); return AT_OK; }
  if (!std::strcmp(name, "gemm_tma_out")) { set_gemm_tma_out(value); return AT_OK; }
  if (!std::strcmp(name, "gemm_band_mb")) { set_gemm_band_mb(value); return AT_OK; }
  if (!std::strcmp(name, "gemm_tma_resid")) { set_gemm_tma_resid(value); return AT_OK; }
  if (!std::strcmp(name, "attn_simt_max_lq")) { set_attn_simt_max_lq(value); return AT_OK; }
  if (!std::strcmp(name, "skinny_tokens")) { g_skinny_tokens = value; return AT_OK; }
  if (!std::strcmp(name, "attn_bound")) { g_attn_bound = value; return AT_OK; }
  if (!std::strcmp(name, "posconv4")) { g_posconv4 = value; return AT_OK; }
  if (!std::strcmp(name, "conv0_fold")) { g_conv0_fold = value; return AT_OK; }
  if (!std::strcmp(name, "attn_split")) { g_attn_split = value; return AT_OK; }
  if (!std::strcmp(name, "attn_blk")) { set_attn_blk(value); return AT_OK; }
  if (!std::strcmp(name, "w2v_graph_chunks")) { g_w2v_graph_chunks = value; return AT_OK; }
  if (!std::strcmp(name, "skinny_max_m")) { set_skinny_max_m(value); return AT_OK; }
  set_last_error("artalk_set_option: unknown option '%s'", name);
  return AT_EINVAL;
}
int artalk_trace_begin(void* stream) { return trace_begin((cudaStream_t)stream); }
long artalk_trace_end(char* host_buf, long cap, void* stream) { return trace_end(host_buf, cap, (cudaStream_t)stream); }
int artalk_profile_enable(artalk_engine_t* e, int enable) {
  AT_REQUIRE(e, "null engine");
  return e->eng.prof_begin(enable);
}
int artalk_profile_read(artalk_engine_t* e, double* host_out8, void* stream) {
  AT_REQUIRE(e && host_out8, "null argument");
  return e->eng.prof_read(host_out8, (cudaStream_t)stream);
}

int artalk_op_gemm(const artalk_gemm_t* a, int precision, void* stream) {
  AT_REQUIRE(a, "null gemm");
  GemmArgs g = gemm_args();
  g.A = a->A; g.a_map = rm(a->a_map); g.W = a->W; g.ldw = a->ldw; g.M = a->M; g.N = a->N; g.K = a->K;
  g.tap_w = a->tap_w; g.tap_pad = a->tap_pad; g.groups = a->groups > 0 ? a->groups : 1;
  g.a_gs = a->a_gs; g.w_gs = a->w_gs; g.c_gs = a->c_gs; g.bias_gs = a->bias_gs;
  g.bias = a->bias; g.act = a->act; g.gate = a->gate; g.gate_dt = a->gate_dt; g.gate_map = rm(a->gate_map);
  g.resid = a->resid; g.resid_map = rm(a->resid_map); g.out32 = a->out32; g.out_act = a->out_act; g.out_act_dt = a->out_act_dt;
  g.c_map = rm(a->c_map);
  g.tap_slots = a->tap_slots > 0 ? a->tap_slots : 1; g.exact = a->exact; g.split_acc = a->split_acc;
  g.skinny = g.exact ? 0 : 1;   // op-level calls: any shape within option "skinny_max_m" may take the latency kernel (bf16-grade epilogue)
  return precision == ARTALK_PRECISION_FP32 ? launch_gemm_simt(g, (cudaStream_t)stream) : launch_gemm_tc(g, (cudaStream_t)stream);
}

int artalk_op_split_bf16(const float* x, void* out, int64_t n, int slots, int is_w, void* stream) {
  return launch_split_bf16(x, out, n, slots, is_w, (cudaStream_t)stream);
}

int artalk_op_attention(const artalk_attn_t* a, void* stream) {
  AT_REQUIRE(a, "null attn");
  AttnArgs x;
  x.q = a->q; x.k = a->k; x.v = a->v; x.out = a->out; x.dt = a->dt; x.n_seq = a->n_seq; x.n_heads = a->n_heads;
  x.head_dim = a->head_dim; x.lq = a->lq; x.lk = a->lk; x.q_ss = a->q_ss; x.q_rs = a->q_rs; x.k_ss = a->k_ss; x.k_rs = a->k_rs;
  x.v_ss = a->v_ss; x.v_rs = a->v_rs; x.o_ss = a->o_ss; x.o_rs = a->o_rs; x.scale = a->scale; x.split = a->split;
  x.key_bound = a->key_bound;
  return launch_attention(x, (cudaStream_t)stream);
}

int artalk_op_attention_split(const artalk_attn_t* a, void* scratch, size_t scratch_bytes, void* stream) {
  AT_REQUIRE(a && scratch, "null argument");
  AttnArgs x;
  x.q = a->q; x.k = a->k; x.v = a->v; x.out = a->out; x.dt = a->dt; x.n_seq = a->n_seq; x.n_heads = a->n_heads;
  x.head_dim = a->head_dim; x.lq = a->lq; x.lk = a->lk; x.q_ss = a->q_ss; x.q_rs = a->q_rs; x.k_ss = a->k_ss; x.k_rs = a->k_rs;
  x.v_ss = a->v_ss; x.v_rs = a->v_rs; x.o_ss = a->o_ss; x.o_rs = a->o_rs; x.scale = a->scale; x.split = a->split;
  x.key_bound = a->key_bound;
  AT_REQUIRE(attention_split_supported(x), "attention_split: unsupported shape (lq=%d lk=%d head_dim=%d)", x.lq, x.lk, x.head_dim);
  AT_REQUIRE(scratch_bytes >= attention_split_scratch_bytes(x), "attention_split: scratch too small (%zu < %zu bytes)", scratch_bytes,
             attention_split_scratch_bytes(x));
  return launch_attention_split(x, scratch, (cudaStream_t)stream);
}

int artalk_op_posconv4(const void* x, const void* w4, const float* bias, const float* resid, float* out, int n_chunks, int frames,
                       int hidden, int groups, int taps, void* stream) {
  AT_REQUIRE(x && w4 && bias && resid && out && resid != out, "null or aliased argument");
  return launch_posconv4(x, w4, bias, resid, out, n_chunks, frames, hidden, groups, taps, (cudaStream_t)stream);
}

int artalk_op_conv0(const float* audio, int n_chunks, int n_samples, const float* w_kc, const float* bias, const float* ln_g,
                    const float* ln_b, const float* wq, const float* bq, const float* qf, float* stats_ws, void* out, int out_dt,
                    float eps, void* stream) {
  AT_REQUIRE(audio && ln_b && stats_ws && out && n_samples >= 10, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int l_out = (n_samples - 10) / 5 + 1;
  AT_TRY(launch_audio_stats(audio, n_chunks, n_samples, (float2*)stats_ws, st));
  if (wq) {
    AT_REQUIRE(bq && qf && out_dt == DT_BF16, "conv0: the folded form needs wq, bq, qf and a bf16 output");
    return launch_conv0_fold(audio, (const float2*)stats_ws, wq, bq, ln_b, qf, out, n_chunks, n_samples, l_out, eps, st);
  }
  AT_REQUIRE(w_kc && bias && ln_g, "null argument");
  return launch_conv0_ln_gelu(audio, (const float2*)stats_ws, w_kc, bias, ln_g, ln_b, out, out_dt, n_chunks, n_samples, l_out, 10, 5, eps, st);
}

int artalk_op_layernorm(const float* x, void* out, int out_dt, const float* gamma, const float* beta, int rows, int cols,
                        float eps, int act, void* stream) {
  AT_REQUIRE(x && out, "null argument");
  return launch_layernorm(x, cols, out, out_dt, cols, gamma, beta, rows, cols, eps, act, (cudaStream_t)stream);
}

}  // extern "C"
