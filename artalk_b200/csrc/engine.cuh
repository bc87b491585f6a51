// Engine: named device tensors (repacked weights, owned by the Python host), a growable device workspace and the
// forward passes of the path. One Engine per (device, weight set); not thread-safe; all work is enqueued on the
// caller's stream.
#pragma once
#include <array>
#include <functional>
#include <map>
#include <string>
#include <vector>
#include "kernels.cuh"

namespace artalk {

struct EngineConfig {        // mirrors artalk_config_t in include/artalk_b200.h
  int precision;             // 0 = fp32 (CUDA-core GEMM), 1 = bf16 (tcgen05 GEMM, fp32 accumulate),
                             // 2 / 3 = parity-grade tensor-core mode: fp32 data flow, tcgen05 GEMMs on 2 / 3 bf16 pieces per
                             // operand (3 / 6 MMA passes, split.cu)
  int ar_depth, ar_heads, embed_dim, cond_dim;
  int vae_depth, vae_heads, vae_hidden, code_dim, motion_dim;
  int n_levels; int patch_nums[8];
  int w2v_layers, w2v_heads, w2v_hidden, w2v_ffn, w2v_conv_dim, w2v_n_conv;
  int w2v_conv_kernel[8]; int w2v_conv_stride[8];
  int w2v_pos_kernel, w2v_pos_groups;
  int style_dim, style_layers, style_heads, style_ffn, style_len;
  int chunk_samples;
  float w2v_ln_eps;
};

struct Tensor { void* ptr; int dt; int64_t numel; };

struct Engine {
  EngineConfig cfg;
  std::map<std::string, Tensor> tensors;
  bool finalized = false;
  // workspace
  char* ws = nullptr; size_t ws_cap = 0; size_t ws_off = 0;
  size_t ws_limit = (size_t)24 << 30;     // soft budget used to size wav2vec sub-batches (artalk_create: 45 % of device memory, <= 80 GiB)
  BitsTables tb;
  // derived
  int L = 0, T = 0, n_audio_frames = 0; int conv_len[8];
  int act_dt() const { return cfg.precision == 1 ? DT_BF16 : DT_F32; }
  int split_slots() const { return cfg.precision == 2 ? 3 : (cfg.precision == 3 ? 6 : 0); }
  // parity-grade mode: bf16 piece blocks of every tensor-core weight (owned, built by finalize) keyed by the fp32 tensor's
  // device pointer, and a grow-on-demand buffer for the piece blocks of the current GEMM's A operand
  std::map<const void*, void*> wsplit;
  char* split_buf = nullptr; size_t split_cap = 0;
  char* asplit_buf = nullptr; size_t asplit_cap = 0;      // q / k / v piece planes of the parity-grade attention
  int grow_piece_buffer(char*& buf, size_t& cap, size_t bytes, cudaStream_t st);
  int attention_split(const AttnArgs& a, cudaStream_t st);
  int gemm_split(const GemmArgs& g, cudaStream_t st);
  void free_split();

  int set_tensor(const char* name, void* ptr, int dt, int64_t numel);
  int finalize();
  const Tensor* find(const std::string& name) const;
  template <typename T> const T* get(const std::string& name) const { const Tensor* t = find(name); return t ? (const T*)t->ptr : nullptr; }
  const void* getw(const std::string& name) const { const Tensor* t = find(name); return t ? t->ptr : nullptr; }

  int ws_reserve(size_t bytes, cudaStream_t st);
  void* ws_alloc(size_t bytes);
  void ws_reset() { ws_off = 0; }

  int gemm(const GemmArgs& g, cudaStream_t st);
  int attention(const AttnArgs& a, cudaStream_t st);
  int posconv4(const void* x, const void* w4, const float* bias, const float* resid, float* out, int n, cudaStream_t st);
  // optional CUDA-event instrumentation of the GEMM / attention launches (bench.py roofline pass)
  bool prof = false;
  std::vector<cudaEvent_t> prof_ev; std::vector<double> prof_flops; std::vector<int> prof_cls; std::vector<std::array<int, 3>> prof_dims;
  int prof_begin(int enable);
  int prof_read(double* out8, cudaStream_t st);

  size_t audio_ws_per_chunk() const;
  int audio_encode(const float* audio, int n_chunks, float* cond, cudaStream_t st);
  int audio_encode_sub(const float* audio, int n, float* cond, cudaStream_t st);
  int style_encode(const float* style_motion, int n_clips, float* style_out, cudaStream_t st);
  int vae_stack(const char* side, int n_clips, int rows_per_clip, int split, float* x, void* xa, cudaStream_t st);
  int vae_decode(const uint32_t* prev_words, const uint32_t* words, int n_clips, float* motion, cudaStream_t st);
  int vae_encode_bits(const float* motion, int n_clips, uint32_t* words_out, float* enc_out_opt, cudaStream_t st);
  int ar_chunk(int n_clips, const float* cond, int64_t cond_cs, const float* style, uint32_t* prev_words,
               float* motion_out, uint32_t* words_out, float* logits_out, const uint32_t* forced_words,
               float* enc_out, cudaStream_t st);
  int ar_chunk_body(int n_clips, const char* scond, const float* style, uint32_t* prev_words, float* motion_out,
                    uint32_t* words, float* logits, const uint32_t* forced_words, float* enc_out, cudaStream_t st);
  // CUDA graphs of the chunk body, keyed by (clips, teacher forcing); valid while the workspace base is unchanged
  struct GraphEntry { cudaGraphExec_t exec; char* ws_base; int warm; unsigned long long n_launches; unsigned int epoch; };
  // set when a capture / instantiation failed: the engine then launches eagerly (correct, ~10-15 % slower at batch 64, several
  // times slower at batch 1) and says so: once on stderr, and to the host through artalk_graph_status
  std::string graph_failure;
  int graph_replays = 0;
  std::map<int, GraphEntry> graphs;
  int run_graphed(int key, cudaStream_t st, size_t body_mark, const std::function<int(cudaStream_t)>& body);
  bool use_graphs = true;
  cudaStream_t gstream = nullptr; cudaEvent_t gev_in = nullptr, gev_out = nullptr;
  int latency_rows = 0;      // artalk_set_latency_mode: GEMMs with at most this many rows take the latency kernel (0 = off)
  void drop_graphs() {
    for (auto& kv : graphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    graphs.clear();
  }
};

void set_savgol_tables(const float* h5, const float* h9);

}  // namespace artalk
