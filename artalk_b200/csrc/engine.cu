// Forward passes of the audio->motion path, scheduled B200-first:
//  * wav2vec2 over all chunks of all clips up-front (independent of the AR state), in sub-batches;
//  * the AR recurrence with a KV cache across the 5 scale steps (only the new tokens of a scale are computed; the
//    cached schedule needs no attention mask), AdaLN parameters for all 12 blocks + head hoisted to one GEMM per chunk,
//    previous-chunk K/V for all 12 blocks in one GEMM per chunk;
//  * VAE decode + re-encode + residual BSQ per chunk.
// Reference semantics: app/models.py:62-121, app/transformer.py:30-79, app/modules/bitwise_vae.py, HF wav2vec2.
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <array>
#include <mutex>
#include <set>
#include "engine.cuh"

namespace artalk {

static thread_local char g_err[1024] = "";
void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }

static std::string S(const char* fmt, ...) {
  char buf[256];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  return std::string(buf);
}

// ------------------------------------------------------------------ registry
int Engine::set_tensor(const char* name, void* ptr, int dt, int64_t numel) {
  AT_REQUIRE(name && ptr && numel > 0, "set_tensor: bad argument for '%s'", name ? name : "?");
  AT_REQUIRE(((uintptr_t)ptr) % 16 == 0, "set_tensor: '%s' must be 16-byte aligned", name);
  Tensor t; t.ptr = ptr; t.dt = dt; t.numel = numel;
  tensors[name] = t;
  finalized = false;
  return AT_OK;
}

const Tensor* Engine::find(const std::string& name) const {
  auto it = tensors.find(name);
  return it == tensors.end() ? nullptr : &it->second;
}

int Engine::finalize() {
  const EngineConfig& c = cfg;
  AT_REQUIRE(c.precision >= 0 && c.precision <= 3, "precision must be 0 (fp32), 1 (bf16), 2 (bf16x3) or 3 (bf16x6)");
  AT_REQUIRE(c.embed_dim == 768 && c.embed_dim / c.ar_heads == 64, "AR width must be 768 with 64-d heads");
  AT_REQUIRE(c.vae_hidden == 512 && c.vae_hidden / c.vae_heads == 64, "VAE width must be 512 with 64-d heads");
  AT_REQUIRE(c.code_dim == 32 && c.motion_dim == 106, "code_dim 32 / motion_dim 106 expected");
  AT_REQUIRE(c.w2v_hidden == 1024 && c.w2v_conv_dim == 512 && c.w2v_hidden / c.w2v_heads == 64, "wav2vec dims");
  AT_REQUIRE(c.n_levels >= 2 && c.n_levels <= 8, "n_levels");
  AT_REQUIRE(c.style_dim == 128 && c.style_dim / c.style_heads == 32, "style encoder dims");
  const int TC_W = 0x100;                  // marks the operands of tensor-core GEMMs (split into bf16 pieces in modes 2 / 3)
  const int wt = act_dt() | TC_W;
  struct Need { std::string name; int dt; int64_t numel; bool tc; };
  std::vector<Need> need;
  auto req = [&](const std::string& n, int dt, int64_t numel) { need.push_back({n, dt & 0xff, numel, (dt & TC_W) != 0}); };
  L = 0;
  for (int i = 0; i < c.n_levels; ++i) L += c.patch_nums[i];
  T = c.patch_nums[c.n_levels - 1];
  int len = c.chunk_samples;
  for (int i = 0; i < c.w2v_n_conv; ++i) { len = (len - c.w2v_conv_kernel[i]) / c.w2v_conv_stride[i] + 1; conv_len[i] = len; }
  n_audio_frames = len;
  const int CD = c.w2v_conv_dim, H = c.w2v_hidden, C = c.embed_dim, VH = c.vae_hidden;
  // wav2vec
  req("w2v.conv0.w", DT_F32, (int64_t)c.w2v_conv_kernel[0] * CD);
  for (int i = 0; i < c.w2v_n_conv; ++i) {
    if (i) req(S("w2v.conv%d.w", i), wt, (int64_t)CD * c.w2v_conv_kernel[i] * CD);
    req(S("w2v.conv%d.b", i), DT_F32, CD); req(S("w2v.conv%d.ln_g", i), DT_F32, CD); req(S("w2v.conv%d.ln_b", i), DT_F32, CD);
  }
  req("w2v.proj.ln_g", DT_F32, CD); req("w2v.proj.ln_b", DT_F32, CD);
  req("w2v.proj.w", wt, (int64_t)H * CD); req("w2v.proj.b", DT_F32, H);
  req("w2v.pos.w", wt, (int64_t)H * (H / c.w2v_pos_groups) * c.w2v_pos_kernel); req("w2v.pos.b", DT_F32, H);
  req("w2v.enc_ln_g", DT_F32, H); req("w2v.enc_ln_b", DT_F32, H);
  for (int l = 0; l < c.w2v_layers; ++l) {
    req(S("w2v.l%d.ln1_g", l), DT_F32, H); req(S("w2v.l%d.ln1_b", l), DT_F32, H);
    req(S("w2v.l%d.qkv.w", l), wt, (int64_t)3 * H * H); req(S("w2v.l%d.qkv.b", l), DT_F32, 3 * H);
    req(S("w2v.l%d.out.w", l), wt, (int64_t)H * H); req(S("w2v.l%d.out.b", l), DT_F32, H);
    req(S("w2v.l%d.ln2_g", l), DT_F32, H); req(S("w2v.l%d.ln2_b", l), DT_F32, H);
    req(S("w2v.l%d.ff1.w", l), wt, (int64_t)c.w2v_ffn * H); req(S("w2v.l%d.ff1.b", l), DT_F32, c.w2v_ffn);
    req(S("w2v.l%d.ff2.w", l), wt, (int64_t)H * c.w2v_ffn); req(S("w2v.l%d.ff2.b", l), DT_F32, H);
  }
  // AR
  const int n_ada = c.ar_depth * 6 * C + 2 * C;
  req("ar.ada.w", wt, (int64_t)n_ada * c.cond_dim); req("ar.ada.b", DT_F32, n_ada);
  req("ar.prevkv.w", wt, (int64_t)c.ar_depth * 2 * C * C); req("ar.prevkv.b", DT_F32, c.ar_depth * 2 * C);
  for (int l = 0; l < c.ar_depth; ++l) {
    req(S("ar.l%d.qkv.w", l), wt, (int64_t)3 * C * C); req(S("ar.l%d.qkv.b", l), DT_F32, 3 * C);
    req(S("ar.l%d.head_scale", l), DT_F32, c.ar_heads);
    req(S("ar.l%d.proj.w", l), wt, (int64_t)C * C); req(S("ar.l%d.proj.b", l), DT_F32, C);
    req(S("ar.l%d.ff1.w", l), wt, (int64_t)4 * C * C); req(S("ar.l%d.ff1.b", l), DT_F32, 4 * C);
    req(S("ar.l%d.ff2.w", l), wt, (int64_t)4 * C * C); req(S("ar.l%d.ff2.b", l), DT_F32, C);
  }
  req("ar.head.w", wt, (int64_t)2 * c.code_dim * C); req("ar.head.b", DT_F32, 2 * c.code_dim);
  req("ar.embed.w", DT_F32, (int64_t)C * c.code_dim); req("ar.embed.b", DT_F32, C);
  req("ar.lvl_pos", DT_F32, (int64_t)L * C); req("ar.prev_lvl_pos", DT_F32, (int64_t)L * C);
  // VAE
  for (const char* side : {"dec", "enc"}) {
    bool dec = side[0] == 'd';
    req(S("vae.%s.in.w", side), wt, (int64_t)VH * (dec ? c.code_dim : 128)); req(S("vae.%s.in.b", side), DT_F32, VH);
    for (int l = 0; l < c.vae_depth; ++l) {
      req(S("vae.%s.l%d.ln_g", side, l), DT_F32, VH); req(S("vae.%s.l%d.ln_b", side, l), DT_F32, VH);
      req(S("vae.%s.l%d.qkv.w", side, l), wt, (int64_t)3 * VH * VH);
      req(S("vae.%s.l%d.out.w", side, l), wt, (int64_t)VH * VH); req(S("vae.%s.l%d.out.b", side, l), DT_F32, VH);
      req(S("vae.%s.l%d.ff1.w", side, l), wt, (int64_t)(VH * 3 / 2) * VH); req(S("vae.%s.l%d.ff1.b", side, l), DT_F32, VH * 3 / 2);
      req(S("vae.%s.l%d.ff2.w", side, l), wt, (int64_t)VH * (VH * 3 / 2)); req(S("vae.%s.l%d.ff2.b", side, l), DT_F32, VH);
    }
    req(S("vae.%s.out.w", side), wt, (int64_t)(dec ? c.motion_dim : c.code_dim) * VH);
    req(S("vae.%s.out.b", side), DT_F32, dec ? c.motion_dim : c.code_dim);
  }
  req("vae.dec_pos", DT_F32, (int64_t)2 * T * c.code_dim); req("vae.enc_pos", DT_F32, (int64_t)T * c.motion_dim);
  req("vae.mean", DT_F32, c.motion_dim); req("vae.std", DT_F32, c.motion_dim);
  // style encoder (always fp32)
  const int SD = c.style_dim;
  req("style.mean", DT_F32, c.motion_dim); req("style.std", DT_F32, c.motion_dim);
  req("style.zero_pos", DT_F32, (int64_t)c.style_len * c.motion_dim);
  req("style.proj.w", DT_F32, (int64_t)SD * 112); req("style.proj.b", DT_F32, SD);
  for (int l = 0; l < c.style_layers; ++l) {
    req(S("style.l%d.qkv.w", l), DT_F32, (int64_t)3 * SD * SD); req(S("style.l%d.qkv.b", l), DT_F32, 3 * SD);
    req(S("style.l%d.out.w", l), DT_F32, (int64_t)SD * SD); req(S("style.l%d.out.b", l), DT_F32, SD);
    req(S("style.l%d.ln1_g", l), DT_F32, SD); req(S("style.l%d.ln1_b", l), DT_F32, SD);
    req(S("style.l%d.ff1.w", l), DT_F32, (int64_t)c.style_ffn * SD); req(S("style.l%d.ff1.b", l), DT_F32, c.style_ffn);
    req(S("style.l%d.ff2.w", l), DT_F32, (int64_t)SD * c.style_ffn); req(S("style.l%d.ff2.b", l), DT_F32, SD);
    req(S("style.l%d.ln2_g", l), DT_F32, SD); req(S("style.l%d.ln2_b", l), DT_F32, SD);
  }
  req("style.embed.w", DT_F32, (int64_t)C * SD); req("style.embed.b", DT_F32, C);
  req("style.null", DT_F32, C);
  // operator tables
  for (const char* n : {"tb.up_i0", "tb.up_i1", "tb.pool_start", "tb.pool_end"}) req(n, 2, (int64_t)c.n_levels * T);
  req("tb.up_w1", DT_F32, (int64_t)c.n_levels * T);
  for (const Need& n : need) {
    const Tensor* t = find(n.name);
    if (!t) { set_last_error("finalize: missing tensor '%s'", n.name.c_str()); return AT_EMISSING; }
    if (t->dt != n.dt || t->numel != n.numel) {
      set_last_error("finalize: tensor '%s' has dtype %d numel %lld, expected dtype %d numel %lld", n.name.c_str(), t->dt,
                     (long long)t->numel, n.dt, (long long)n.numel);
      return AT_EINVAL;
    }
  }
  {                                                 // optional: folded conv layer 0 operands (bf16 mode)
    const Tensor* qf = find("w2v.conv0.qf"); const Tensor* wq = find("w2v.conv0.wq"); const Tensor* bq = find("w2v.conv0.bq");
    if (qf || wq || bq) {
      if (!(qf && wq && bq) || qf->dt != DT_F32 || wq->dt != DT_F32 || bq->dt != DT_F32 || qf->numel != 11 * 12 || wq->numel != 10 * (int64_t)CD ||
          bq->numel != CD || CD != 512) {
        set_last_error("finalize: 'w2v.conv0.wq' [10][512] / 'w2v.conv0.bq' [512] / 'w2v.conv0.qf' [11][12] (f32) must be given together");
        return AT_EINVAL;
      }
    }
  }
  if (const Tensor* w4 = find("w2v.pos.w4")) {      // optional: shifted filter copies of the positional conv (bf16 mode)
    const int64_t want = (int64_t)c.w2v_pos_groups * 256 * (c.w2v_pos_kernel + 3) * 64;
    if (w4->dt != DT_BF16 || w4->numel != want) {
      set_last_error("finalize: tensor 'w2v.pos.w4' has dtype %d numel %lld, expected bf16 numel %lld", w4->dt, (long long)w4->numel, (long long)want);
      return AT_EINVAL;
    }
  }
  tb.n_levels = c.n_levels; tb.T = T; tb.L = L;
  int cum = 0;
  for (int i = 0; i < 8; ++i) { tb.pn[i] = 0; tb.cum[i] = 0; }
  for (int i = 0; i < c.n_levels; ++i) { tb.pn[i] = c.patch_nums[i]; cum += c.patch_nums[i]; tb.cum[i] = cum; }
  tb.up_i0 = get<int>("tb.up_i0"); tb.up_i1 = get<int>("tb.up_i1"); tb.up_w1 = get<float>("tb.up_w1");
  tb.pool_start = get<int>("tb.pool_start"); tb.pool_end = get<int>("tb.pool_end");
  // parity-grade mode: bf16 piece blocks of every tensor-core weight, once per weight set
  free_split();
  if (split_slots()) {
    const int S = split_slots();
    for (const Need& n : need) {
      if (!n.tc || n.numel % 64 != 0) continue;
      const Tensor* t = find(n.name);
      void* d = nullptr;
      cudaError_t e = cudaMalloc(&d, (size_t)n.numel * S * sizeof(bf16));
      if (e != cudaSuccess) {
        cudaGetLastError();
        set_last_error("finalize: cudaMalloc of the %d piece blocks of '%s' failed: %s", S, n.name.c_str(), cudaGetErrorString(e));
        return AT_ENOMEM;
      }
      wsplit[t->ptr] = d;
      AT_TRY(launch_split_bf16((const float*)t->ptr, d, n.numel, S, 1, (cudaStream_t)0));
    }
    AT_CUDA(cudaDeviceSynchronize());
  }
  drop_graphs();                            // a re-finalized engine may have new weights behind old addresses
  finalized = true;
  return AT_OK;
}

void Engine::free_split() {
  for (auto& kv : wsplit) cudaFree(kv.second);
  wsplit.clear();
}

// Piece buffers of the parity-grade modes grow on demand (outside graph capture: the eager warm-up call of a chunk sizes them)
int Engine::grow_piece_buffer(char*& buf, size_t& cap, size_t bytes, cudaStream_t st) {
  if (bytes <= cap) return AT_OK;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  AT_CUDA(cudaStreamIsCapturing(st, &cs));
  AT_REQUIRE(cs == cudaStreamCaptureStatusNone, "piece buffer would have to grow during graph capture");
  AT_CUDA(cudaStreamSynchronize(st));
  if (gstream) AT_CUDA(cudaStreamSynchronize(gstream));
  drop_graphs();                          // captured launches point into the old buffer
  if (buf) { AT_CUDA(cudaFree(buf)); buf = nullptr; cap = 0; }
  const size_t want = bytes + bytes / 4 + (1 << 20);
  cudaError_t e = cudaMalloc((void**)&buf, want);
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_last_error("piece buffer: cudaMalloc(%zu MiB) failed: %s", want >> 20, cudaGetErrorString(e));
    return AT_ENOMEM;
  }
  cap = want;
  return AT_OK;
}

// Attention at fp32 grade on the tensor cores (precision "bf16x3"): q, k, v (fp32, strided views) are split into two bf16 piece
// planes each and the SPLIT variant of the tcgen05 kernel keeps the three piece products of both contractions
// (attention_tc.cu). Shapes it cannot take (head_dim 32, fewer than 17 or more than 368 keys) stay on the fp32 SIMT kernel.
int Engine::attention_split(const AttnArgs& a, cudaStream_t st) {
  AT_TRY(grow_piece_buffer(asplit_buf, asplit_cap, attention_split_scratch_bytes(a), st));
  return launch_attention_split(a, asplit_buf, st);
}

// C = A W^T on the tensor cores at fp32 grade: A is split into bf16 piece blocks on the fly, W was split by finalize, and the
// bf16 GEMM kernel runs the `slots` piece products as one GEMM with K' = slots * K (split.cu). Shapes the block layout cannot
// express (K or a stride not a multiple of 64: the VAE decoder's 32-wide input mapping) take the fp32 CUDA-core kernel.
int Engine::gemm_split(const GemmArgs& g, cudaStream_t st) {
  const int S = split_slots();
  const bool tap = g.tap_w > 0;
  const bool shape_ok = g.K % 64 == 0 && g.ldw % 64 == 0 && g.a_map.rs % 64 == 0 && g.a_map.bs % 64 == 0 && g.w_gs % 64 == 0 &&
                        (!tap || (g.a_gs == 64 && g.tap_w == 64)) && (tap || g.groups == 1) && !g.qkv_mode;
  if (!shape_ok) return launch_gemm_simt(g, st);
  auto it = wsplit.find(g.W);
  AT_REQUIRE(it != wsplit.end(), "gemm_split: the weight operand %p is not a registered tensor-core weight", g.W);
  int64_t extent;
  if (g.a_map.rpb > 0) {
    const int64_t nb = g.M / g.a_map.rpb;
    extent = (nb - 1) * g.a_map.bs + (int64_t)(g.a_map.rpb - 1) * g.a_map.rs + (tap ? g.a_map.rs : (int64_t)g.K);
  } else {
    extent = (int64_t)(g.M - 1) * g.a_map.rs + g.K;
  }
  const size_t bytes = (size_t)extent * S * sizeof(bf16);
  AT_TRY(grow_piece_buffer(split_buf, split_cap, bytes, st));
  AT_TRY(launch_split_bf16((const float*)g.A, split_buf, extent, S, 0, st));
  GemmArgs h = g;
  h.A = split_buf; h.a_map.rs *= S; h.a_map.bs *= S;
  h.W = it->second; h.ldw *= S; h.w_gs *= S; h.K *= S;
  h.tap_slots = tap ? S : 1; h.exact = 1; h.skinny = 0; h.split_acc = S;
  return launch_gemm_tc(h, st);
}

// ------------------------------------------------------------------ workspace
int Engine::ws_reserve(size_t bytes, cudaStream_t st) {
  ws_off = 0;
  if (bytes <= ws_cap) return AT_OK;
  AT_CUDA(cudaStreamSynchronize(st));
  if (ws) { AT_CUDA(cudaFree(ws)); ws = nullptr; ws_cap = 0; }
  cudaError_t e = cudaMalloc((void**)&ws, bytes);
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_last_error("workspace: cudaMalloc(%zu MiB) failed: %s", bytes >> 20, cudaGetErrorString(e));
    return AT_ENOMEM;
  }
  ws_cap = bytes;
  return AT_OK;
}

void* Engine::ws_alloc(size_t bytes) {
  size_t off = (ws_off + 255) & ~(size_t)255;
  if (off + bytes > ws_cap) { set_last_error("workspace overflow: need %zu at %zu of %zu", bytes, off, ws_cap); return nullptr; }
  ws_off = off + bytes;
  return ws + off;
}
#define WS(var, type, bytes)                              \
  type var = (type)ws_alloc(bytes);                       \
  if (!var) return AT_ENOMEM

std::atomic<unsigned long long> g_launch_count{0};
std::atomic<unsigned int> g_option_epoch{0};

// ------------------------------------------------------------------ per-device state (common.cuh)
namespace {
std::mutex g_dev_mu;
DevCtx g_dev[16];
bool g_dev_init[16] = {};
std::set<std::pair<int, const void*>> g_smem_done;
int dev_slot(int dev) { return dev < 0 ? 0 : (dev > 15 ? 15 : dev); }
}  // namespace

int dev_ctx(const DevCtx** out) {
  int dev = 0;
  AT_CUDA(cudaGetDevice(&dev));
  AT_REQUIRE(dev >= 0 && dev < 16, "device ordinal %d is not supported (0..15)", dev);
  std::lock_guard<std::mutex> lk(g_dev_mu);
  DevCtx& c = g_dev[dev];
  if (!g_dev_init[dev]) {
    c.dev = dev;
    AT_CUDA(cudaDeviceGetAttribute(&c.num_sms, cudaDevAttrMultiProcessorCount, dev));
    AT_CUDA(cudaMalloc((void**)&c.err_flag, sizeof(unsigned int)));
    AT_CUDA(cudaMemset(c.err_flag, 0, sizeof(unsigned int)));
    g_dev_init[dev] = true;
  }
  *out = &c;
  return AT_OK;
}

int ensure_dyn_smem(const void* kernel, int bytes) {
  int dev = 0;
  AT_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(g_dev_mu);
  if (g_smem_done.count(std::make_pair(dev, kernel))) return AT_OK;
  AT_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  g_smem_done.insert(std::make_pair(dev, kernel));
  return AT_OK;
}

int& per_device_slot(int (&slots)[16]) {
  int dev = 0;
  cudaGetDevice(&dev);
  return slots[dev_slot(dev)];
}
bool g_pdl = true;
int g_skinny_tokens = 1;        // AR scale steps with at most this many new tokens per clip take the latency kernels (option "skinny_tokens")
int g_pdl_mask = 3;            // measured in the chunk graph: GEMM/attention edges help (-3.4 %), elementwise edges cancel it
int g_pdl_w2v_max_chunks = 1 << 30;
int g_w2v_graph_chunks = 4;     // option "w2v_graph_chunks": wav2vec calls with at most this many chunks replay a CUDA graph
int g_attn_split = 1;          // option "attn_split": precision bf16x3 runs attention on the tensor cores from two bf16 pieces per operand
int g_conv0_fold = 1;          // option "conv0_fold": bf16 mode runs conv layer 0 with the LayerNorm folded through the conv (conv0_fold.cu)
int g_posconv4 = 1;            // option "posconv4": bf16 mode runs the positional conv in its four-frames-per-row form (posconv_tc.cu)
int g_attn_bound = 1;          // option "attn_bound": AR attention subtracts the per-head score bound instead of the row maximum

// ------------------------------------------------------------------ launch trace
bool g_trace_on = false;
int g_trace_dims[3] = {0, 0, 0};
struct TraceRec { const char* func; int d[3]; cudaEvent_t ev; };
static std::vector<TraceRec> g_trace;
static std::vector<cudaEvent_t> g_trace_pool;

void trace_event(const char* func, cudaStream_t st) {
  if (g_trace.size() >= g_trace_pool.size()) return;
  TraceRec r;
  r.func = func; r.d[0] = g_trace_dims[0]; r.d[1] = g_trace_dims[1]; r.d[2] = g_trace_dims[2];
  r.ev = g_trace_pool[g_trace.size()];
  g_trace_dims[0] = g_trace_dims[1] = g_trace_dims[2] = 0;
  if (cudaEventRecord(r.ev, st) == cudaSuccess) g_trace.push_back(r);
}

int trace_begin(cudaStream_t st) {
  if (g_trace_pool.empty()) {
    g_trace_pool.resize(16384);
    for (auto& e : g_trace_pool) AT_CUDA(cudaEventCreate(&e));
  }
  g_trace.clear();
  g_trace_on = true;
  trace_event("begin", st);
  return AT_OK;
}

// writes "func,d0,d1,d2,microseconds\n" lines; returns the number of bytes written (or needed if larger than cap)
long trace_end(char* buf, long cap, cudaStream_t st) {
  g_trace_on = false;
  if (cudaStreamSynchronize(st) != cudaSuccess) return -1;
  long off = 0;
  for (size_t i = 1; i < g_trace.size(); ++i) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, g_trace[i - 1].ev, g_trace[i].ev);
    char line[160];
    int n = snprintf(line, sizeof(line), "%s,%d,%d,%d,%.2f\n", g_trace[i].func, g_trace[i].d[0], g_trace[i].d[1], g_trace[i].d[2],
                     ms * 1000.0f);
    if (off + n < cap) memcpy(buf + off, line, n);
    off += n;
  }
  if (off < cap) buf[off] = 0;
  return off;
}

int Engine::prof_begin(int enable) {
  prof = enable != 0;
  prof_flops.clear(); prof_cls.clear(); prof_dims.clear();
  if (prof && prof_ev.empty()) {
    prof_ev.resize(2 * 8192);
    for (auto& e : prof_ev) AT_CUDA(cudaEventCreate(&e));
  }
  return AT_OK;
}

// out[16] = {gemm launches, gemm ms, gemm flops, attn launches, attn ms, attn flops, dominant GEMM shape: launches, total ms,
//            flops per launch, M, N, K, 0...}
int Engine::prof_read(double* out, cudaStream_t st) {
  AT_CUDA(cudaStreamSynchronize(st));
  for (int i = 0; i < 16; ++i) out[i] = 0.0;
  std::map<std::array<int, 3>, std::pair<double, double>> by_shape;        // GEMM (M, N, K) -> (launches, total ms)
  for (size_t i = 0; i < prof_cls.size(); ++i) {
    float ms = 0.f;
    AT_CUDA(cudaEventElapsedTime(&ms, prof_ev[2 * i], prof_ev[2 * i + 1]));
    int c = prof_cls[i] * 3;
    out[c] += 1.0; out[c + 1] += ms; out[c + 2] += prof_flops[i];
    if (prof_cls[i] == 0) { auto& e = by_shape[prof_dims[i]]; e.first += 1.0; e.second += ms; }
  }
  // out[6..8]: the dominant GEMM shape (largest summed duration): launches, total ms, flops per launch
  for (auto& kv : by_shape)
    if (kv.second.second > out[7]) {
      out[6] = kv.second.first; out[7] = kv.second.second;
      out[9] = kv.first[0]; out[10] = kv.first[1]; out[11] = kv.first[2];
      out[8] = 2.0 * out[9] * out[10] * out[11];
    }
  return AT_OK;
}

#define PROF_WRAP(cls, flops, d0, d1, d2, call)                                           \
  do {                                                                          \
    size_t _i = prof_cls.size();                                                \
    bool _on = prof && 2 * _i + 1 < prof_ev.size();                             \
    if (_on) AT_CUDA(cudaEventRecord(prof_ev[2 * _i], st));                     \
    AT_TRY(call);                                                               \
    if (_on) {                                                                  \
      AT_CUDA(cudaEventRecord(prof_ev[2 * _i + 1], st));                        \
      prof_cls.push_back(cls); prof_flops.push_back(flops);                     \
      prof_dims.push_back(std::array<int, 3>{{(int)(d0), (int)(d1), (int)(d2)}}); \
    }                                                                           \
    return AT_OK;                                                               \
  } while (0)

int Engine::gemm(const GemmArgs& g0, cudaStream_t st) {
  GemmArgs g = g0;
  // latency mode: few rows and a small output (the hoisted AdaLN / previous-chunk K/V GEMMs stay on the tensor-core kernel)
  if (latency_rows > 0 && g.M <= latency_rows && (int64_t)g.M * g.N <= ((int64_t)1 << 21)) g.skinny = 1;
  PROF_WRAP(0, 2.0 * g.M * g.N * g.K * g.groups, g.M, g.N * g.groups, g.K,
            cfg.precision == 0 ? launch_gemm_simt(g, st) : (cfg.precision == 1 ? launch_gemm_tc(g, st) : gemm_split(g, st)));
}

int Engine::attention(const AttnArgs& a, cudaStream_t st) {
  double keys = (a.split > 0) ? 0.5 * (a.split + a.lk) : a.lk;     // rows < split see `split` keys, the rest see lk
  const bool tc_split = cfg.precision == 2 && g_attn_split && attention_split_supported(a);
  PROF_WRAP(1, 4.0 * a.n_seq * a.n_heads * a.head_dim * a.lq * keys, a.n_seq * a.n_heads, a.lq, a.lk,
            tc_split ? attention_split(a, st) : launch_attention(a, st));
}

int Engine::posconv4(const void* x, const void* w4, const float* bias, const float* resid, float* out, int n, cudaStream_t st) {
  const EngineConfig& c = cfg;
  const int G = c.w2v_pos_groups, H = c.w2v_hidden, F = n_audio_frames;
  PROF_WRAP(0, 2.0 * n * F * H * (double)(H / G) * c.w2v_pos_kernel, n * F, H, c.w2v_pos_kernel * (H / G),
            launch_posconv4(x, w4, bias, resid, out, n, F, H, G, c.w2v_pos_kernel, st));
}

// ------------------------------------------------------------------ wav2vec2
size_t Engine::audio_ws_per_chunk() const {
  const size_t s = dt_size(act_dt());
  const size_t CD = cfg.w2v_conv_dim, H = cfg.w2v_hidden;
  size_t bufA = std::max((size_t)conv_len[0] * CD, (size_t)n_audio_frames * std::max((size_t)cfg.w2v_ffn, 3 * H)) * s;
  size_t bufB = std::max((size_t)conv_len[1] * CD, (size_t)n_audio_frames * H) * s;
  size_t pre = (size_t)conv_len[1] * CD * 4;
  size_t h = (size_t)n_audio_frames * H * 4;
  return bufA + bufB + pre + 2 * h + 4096;
}

int Engine::audio_encode(const float* audio, int n_chunks, float* cond, cudaStream_t st) {
  AT_REQUIRE(finalized, "engine not finalized");
  if (n_chunks <= 0) return AT_OK;
  size_t per = audio_ws_per_chunk();
  if (n_chunks <= g_w2v_graph_chunks) {
    // streaming / few-chunk calls are launch bound on the host (~190 launches of 5-8 us kernels per chunk): the encoder body
    // runs from a CUDA graph like the AR chunk, with the audio staged in and the conditioning staged out of the workspace
    const size_t in_b = (size_t)n_chunks * cfg.chunk_samples * 4, out_b = (size_t)n_chunks * L * cfg.w2v_hidden * 4;
    AT_TRY(ws_reserve((size_t)n_chunks * per + in_b + out_b + (2 << 20), st));
    WS(a_in, float*, in_b);
    WS(c_out, float*, out_b);
    const size_t body_mark = ws_off;
    AT_CUDA(cudaMemcpyAsync(a_in, audio, in_b, cudaMemcpyDeviceToDevice, st));
    AT_TRY(run_graphed(0x40000000 | n_chunks, st, body_mark, [&](cudaStream_t s) { return audio_encode_sub(a_in, n_chunks, c_out, s); }));
    AT_CUDA(cudaMemcpyAsync(cond, c_out, out_b, cudaMemcpyDeviceToDevice, st));
    return AT_OK;
  }
  int sub = (int)std::max<size_t>(1, std::min<size_t>((size_t)n_chunks, ws_limit / per));
  AT_TRY(ws_reserve((size_t)sub * per + (1 << 20), st));
  for (int c0 = 0; c0 < n_chunks; c0 += sub) {
    int n = std::min(sub, n_chunks - c0);
    ws_reset();
    AT_TRY(audio_encode_sub(audio + (int64_t)c0 * cfg.chunk_samples, n, cond + (int64_t)c0 * L * cfg.w2v_hidden, st));
  }
  return AT_OK;
}

int Engine::audio_encode_sub(const float* audio, int n, float* cond, cudaStream_t st) {
  const EngineConfig& c = cfg;
  const int adt = act_dt();
  const size_t s = dt_size(adt);
  const int CD = c.w2v_conv_dim, H = c.w2v_hidden, F = n_audio_frames, M = n * F;
  // programmatic dependent launch on every kernel class cost 4 % on a 192-chunk batch (elementwise kernels with 10^4 CTAs as
  // early-resident dependents); with the class mask restricted to the tcgen05 kernels (g_pdl_mask = 3) it gains ~1.5 %, so the
  // chunk-count cut-off is off by default and kept as an option
  struct PdlScope { bool saved; PdlScope(bool on) : saved(g_pdl) { g_pdl = g_pdl && on; } ~PdlScope() { g_pdl = saved; } } pdl_scope(n <= g_pdl_w2v_max_chunks);
  WS(stats, float2*, (size_t)n * sizeof(float2));
  WS(bufA, char*, (size_t)n * std::max((size_t)conv_len[0] * CD, (size_t)F * std::max((size_t)c.w2v_ffn, (size_t)3 * H)) * s);
  WS(bufB, char*, (size_t)n * std::max((size_t)conv_len[1] * CD, (size_t)F * H) * s);
  WS(pre, float*, (size_t)n * conv_len[1] * CD * 4);
  WS(h, float*, (size_t)M * H * 4);
  WS(h2, float*, (size_t)M * H * 4);

  AT_TRY(launch_audio_stats(audio, n, c.chunk_samples, stats, st));
  const Tensor* c0q = (c.precision == 1 && g_conv0_fold && c.w2v_conv_kernel[0] == 10 && c.w2v_conv_stride[0] == 5) ? find("w2v.conv0.qf") : nullptr;
  if (c0q && find("w2v.conv0.wq") && find("w2v.conv0.bq"))
    // bf16 mode: LayerNorm statistics from the frame's samples (quadratic form), filters in registers, packed FMAs (conv0_fold.cu)
    AT_TRY(launch_conv0_fold(audio, stats, get<float>("w2v.conv0.wq"), get<float>("w2v.conv0.bq"), get<float>("w2v.conv0.ln_b"),
                             (const float*)c0q->ptr, bufA, n, c.chunk_samples, conv_len[0], c.w2v_ln_eps, st));
  else
    AT_TRY(launch_conv0_ln_gelu(audio, stats, get<float>("w2v.conv0.w"), get<float>("w2v.conv0.b"), get<float>("w2v.conv0.ln_g"),
                                get<float>("w2v.conv0.ln_b"), bufA, adt, n, c.chunk_samples, conv_len[0], c.w2v_conv_kernel[0],
                                c.w2v_conv_stride[0], c.w2v_ln_eps, st));
  char* in = bufA; char* out = bufB;
  for (int i = 1; i < c.w2v_n_conv; ++i) {
    // implicit GEMM: row t of the A operand is the contiguous window in[t*stride : t*stride + k][:] of the
    // channels-last input (overlapping rows, row stride = stride*512)
    GemmArgs g = gemm_args();
    g.A = in; g.a_map = batched_rows(conv_len[i], (int64_t)conv_len[i - 1] * CD, (int64_t)c.w2v_conv_stride[i] * CD);
    g.W = getw(S("w2v.conv%d.w", i)); g.ldw = (int64_t)c.w2v_conv_kernel[i] * CD;
    g.M = n * conv_len[i]; g.N = CD; g.K = c.w2v_conv_kernel[i] * CD;
    g.bias = get<float>(S("w2v.conv%d.b", i));
    g.out32 = pre; g.c_map = plain_rows(CD);
    AT_TRY(gemm(g, st));
    bool last = (i == c.w2v_n_conv - 1);
    // the last conv layer's LN+GELU output feeds another LayerNorm -> keep it fp32 (staged in h)
    AT_TRY(launch_layernorm(pre, CD, last ? (void*)h : (void*)out, last ? DT_F32 : adt, CD, get<float>(S("w2v.conv%d.ln_g", i)),
                            get<float>(S("w2v.conv%d.ln_b", i)), g.M, CD, c.w2v_ln_eps, ACT_GELU_ERF, st));
    std::swap(in, out);
  }
  // feature projection: LN(512) -> Linear 512 -> 1024   (modeling_wav2vec2.py:422-434)
  AT_TRY(launch_layernorm(h, CD, bufB, adt, CD, get<float>("w2v.proj.ln_g"), get<float>("w2v.proj.ln_b"), M, CD, c.w2v_ln_eps,
                          ACT_NONE, st));
  {
    GemmArgs g = gemm_args();
    g.A = bufB; g.a_map = plain_rows(CD); g.W = getw("w2v.proj.w"); g.ldw = CD; g.M = M; g.N = H; g.K = CD;
    g.bias = get<float>("w2v.proj.b"); g.out32 = h; g.c_map = plain_rows(H);
    if (adt != DT_F32) { g.out_act = bufA; g.out_act_dt = adt; }
    AT_TRY(gemm(g, st));
  }
  // positional conv embedding: grouped conv k=128 pad 64 (last output dropped) + GELU + residual (:360-368,764-765)
  const Tensor* w4 = (c.precision == 1 && g_posconv4) ? find("w2v.pos.w4") : nullptr;
  if (w4 && posconv4_supported(F, H, c.w2v_pos_groups, c.w2v_pos_kernel)) {
    // bf16 mode: four output frames per A row, N = 256 pair tiles (posconv_tc.cu); bit-identical to the tap-mode GEMM below
    AT_TRY(posconv4(bufA, w4->ptr, get<float>("w2v.pos.b"), h, h2, n, st));
    std::swap(h, h2);
  } else {
    const int G = c.w2v_pos_groups, gw = H / G;
    GemmArgs g = gemm_args();
    g.A = (adt == DT_F32) ? (const void*)h : (const void*)bufA;
    g.a_map = batched_rows(F, (int64_t)F * H, H);
    g.tap_w = gw; g.tap_pad = c.w2v_pos_kernel / 2;
    g.W = getw("w2v.pos.w"); g.ldw = (int64_t)c.w2v_pos_kernel * gw;
    g.M = M; g.N = gw; g.K = c.w2v_pos_kernel * gw;
    g.groups = G; g.a_gs = gw; g.w_gs = (int64_t)gw * c.w2v_pos_kernel * gw; g.c_gs = gw; g.bias_gs = gw;
    g.bias = get<float>("w2v.pos.b"); g.act = ACT_GELU_ERF;
    g.resid = h; g.resid_map = plain_rows(H); g.out32 = h2; g.c_map = plain_rows(H);
    AT_TRY(gemm(g, st));
    std::swap(h, h2);
  }
  const float att_scale = 1.0f / sqrtf((float)(H / c.w2v_heads));
  for (int l = 0; l < c.w2v_layers; ++l) {
    AT_TRY(launch_layernorm(h, H, bufB, adt, H, get<float>(S("w2v.l%d.ln1_g", l)), get<float>(S("w2v.l%d.ln1_b", l)), M, H,
                            c.w2v_ln_eps, ACT_NONE, st));
    GemmArgs g = gemm_args();
    g.A = bufB; g.a_map = plain_rows(H); g.W = getw(S("w2v.l%d.qkv.w", l)); g.ldw = H; g.M = M; g.N = 3 * H; g.K = H;
    g.bias = get<float>(S("w2v.l%d.qkv.b", l)); g.out_act = bufA; g.out_act_dt = adt; g.c_map = plain_rows(3 * H);
    AT_TRY(gemm(g, st));
    AttnArgs a;
    a.q = bufA; a.k = bufA + (size_t)H * s; a.v = bufA + (size_t)2 * H * s; a.out = bufB; a.dt = adt;
    a.n_seq = n; a.n_heads = c.w2v_heads; a.head_dim = H / c.w2v_heads; a.lq = F; a.lk = F;
    a.q_ss = a.k_ss = a.v_ss = (int64_t)F * 3 * H; a.q_rs = a.k_rs = a.v_rs = 3 * H;
    a.o_ss = (int64_t)F * H; a.o_rs = H; a.scale = att_scale; a.split = 0;
    AT_TRY(attention(a, st));
    g = gemm_args();
    g.A = bufB; g.a_map = plain_rows(H); g.W = getw(S("w2v.l%d.out.w", l)); g.ldw = H; g.M = M; g.N = H; g.K = H;
    g.bias = get<float>(S("w2v.l%d.out.b", l)); g.resid = h; g.resid_map = plain_rows(H); g.out32 = h; g.c_map = plain_rows(H);
    AT_TRY(gemm(g, st));
    AT_TRY(launch_layernorm(h, H, bufB, adt, H, get<float>(S("w2v.l%d.ln2_g", l)), get<float>(S("w2v.l%d.ln2_b", l)), M, H,
                            c.w2v_ln_eps, ACT_NONE, st));
    g = gemm_args();
    g.A = bufB; g.a_map = plain_rows(H); g.W = getw(S("w2v.l%d.ff1.w", l)); g.ldw = H; g.M = M; g.N = c.w2v_ffn; g.K = H;
    g.bias = get<float>(S("w2v.l%d.ff1.b", l)); g.act = ACT_GELU_ERF; g.out_act = bufA; g.out_act_dt = adt;
    g.c_map = plain_rows(c.w2v_ffn);
    AT_TRY(gemm(g, st));
    g = gemm_args();
    g.A = bufA; g.a_map = plain_rows(c.w2v_ffn); g.W = getw(S("w2v.l%d.ff2.w", l)); g.ldw = c.w2v_ffn; g.M = M; g.N = H;
    g.K = c.w2v_ffn; g.bias = get<float>(S("w2v.l%d.ff2.b", l)); g.resid = h; g.resid_map = plain_rows(H); g.out32 = h;
    g.c_map = plain_rows(H);
    AT_TRY(gemm(g, st));
  }
  AT_TRY(launch_layernorm(h, H, h2, DT_F32, H, get<float>("w2v.enc_ln_g"), get<float>("w2v.enc_ln_b"), M, H, c.w2v_ln_eps,
                          ACT_NONE, st));
  AT_TRY(launch_audio_pool(h2, cond, n, F, H, c.patch_nums, c.n_levels, st));
  return AT_OK;
}

// ------------------------------------------------------------------ style encoder (fp32, once per clip)
int Engine::style_encode(const float* style_motion, int n, float* style_out, cudaStream_t st) {
  AT_REQUIRE(finalized, "engine not finalized");
  if (n <= 0) return AT_OK;
  const EngineConfig& c = cfg;
  const int SD = c.style_dim, SL = c.style_len, M = n * SL, KP = 112;
  AT_TRY(ws_reserve((size_t)M * (KP + SD * 2 + 3 * SD + SD + c.style_ffn) * 4 + (size_t)n * SD * 4 + (1 << 16), st));
  WS(xin, float*, (size_t)M * KP * 4);
  WS(x, float*, (size_t)M * SD * 4);
  WS(t, float*, (size_t)M * SD * 4);
  WS(qkv, float*, (size_t)M * 3 * SD * 4);
  WS(o, float*, (size_t)M * SD * 4);
  WS(f, float*, (size_t)M * c.style_ffn * 4);
  WS(pooled, float*, (size_t)n * SD * 4);
  AT_TRY(launch_motion_norm_pos(style_motion, get<float>("style.mean"), get<float>("style.std"), get<float>("style.zero_pos"), xin,
                                DT_F32, n, SL, c.motion_dim, KP, st));
  GemmArgs g = gemm_args();
  g.A = xin; g.a_map = plain_rows(KP); g.W = getw("style.proj.w"); g.ldw = KP; g.M = M; g.N = SD; g.K = KP;
  g.bias = get<float>("style.proj.b"); g.out32 = x; g.c_map = plain_rows(SD);       // bias already holds + pe[style_len]
  AT_TRY(launch_gemm_simt(g, st));
  for (int l = 0; l < c.style_layers; ++l) {
    g = gemm_args();
    g.A = x; g.a_map = plain_rows(SD); g.W = getw(S("style.l%d.qkv.w", l)); g.ldw = SD; g.M = M; g.N = 3 * SD; g.K = SD;
    g.bias = get<float>(S("style.l%d.qkv.b", l)); g.out32 = qkv; g.c_map = plain_rows(3 * SD);
    AT_TRY(launch_gemm_simt(g, st));
    AttnArgs a;
    a.q = qkv; a.k = qkv + SD; a.v = qkv + 2 * SD; a.out = o; a.dt = DT_F32; a.n_seq = n; a.n_heads = c.style_heads;
    a.head_dim = SD / c.style_heads; a.lq = SL; a.lk = SL; a.q_ss = a.k_ss = a.v_ss = (int64_t)SL * 3 * SD;
    a.q_rs = a.k_rs = a.v_rs = 3 * SD; a.o_ss = (int64_t)SL * SD; a.o_rs = SD;
    a.scale = 1.0f / sqrtf((float)a.head_dim); a.split = 0;
    AT_TRY(attention(a, st));
    g = gemm_args();
    g.A = o; g.a_map = plain_rows(SD); g.W = getw(S("style.l%d.out.w", l)); g.ldw = SD; g.M = M; g.N = SD; g.K = SD;
    g.bias = get<float>(S("style.l%d.out.b", l)); g.resid = x; g.resid_map = plain_rows(SD); g.out32 = t; g.c_map = plain_rows(SD);
    AT_TRY(launch_gemm_simt(g, st));
    AT_TRY(launch_layernorm(t, SD, x, DT_F32, SD, get<float>(S("style.l%d.ln1_g", l)), get<float>(S("style.l%d.ln1_b", l)), M, SD,
                            1e-5f, ACT_NONE, st));                                        // post-norm
    g = gemm_args();
    g.A = x; g.a_map = plain_rows(SD); g.W = getw(S("style.l%d.ff1.w", l)); g.ldw = SD; g.M = M; g.N = c.style_ffn; g.K = SD;
    g.bias = get<float>(S("style.l%d.ff1.b", l)); g.act = ACT_GELU_ERF; g.out32 = f; g.c_map = plain_rows(c.style_ffn);
    AT_TRY(launch_gemm_simt(g, st));
    g = gemm_args();
    g.A = f; g.a_map = plain_rows(c.style_ffn); g.W = getw(S("style.l%d.ff2.w", l)); g.ldw = c.style_ffn; g.M = M; g.N = SD;
    g.K = c.style_ffn; g.bias = get<float>(S("style.l%d.ff2.b", l)); g.resid = x; g.resid_map = plain_rows(SD); g.out32 = t;
    g.c_map = plain_rows(SD);
    AT_TRY(launch_gemm_simt(g, st));
    AT_TRY(launch_layernorm(t, SD, x, DT_F32, SD, get<float>(S("style.l%d.ln2_g", l)), get<float>(S("style.l%d.ln2_b", l)), M, SD,
                            1e-5f, ACT_NONE, st));
  }
  int one = 1;
  AT_TRY(launch_audio_pool(x, pooled, n, SL, SD, &one, 1, st));                        // mean over time
  g = gemm_args();
  g.A = pooled; g.a_map = plain_rows(SD); g.W = getw("style.embed.w"); g.ldw = SD; g.M = n; g.N = c.embed_dim; g.K = SD;
  g.bias = get<float>("style.embed.b"); g.out32 = style_out; g.c_map = plain_rows(c.embed_dim);   // 1.1*W, 1.1*b - 0.1*null folded
  AT_TRY(launch_gemm_simt(g, st));
  return AT_OK;
}

// ------------------------------------------------------------------ VAE transformer stack (bitwise_vae.py:128-215)
// x: fp32 residual stream [n*rows, 512]; xa: act copy of x (bf16 mode) or null (fp32 mode: x itself is the operand)
int Engine::vae_stack(const char* side, int n, int rows, int split, float* x, void* xa, cudaStream_t st) {
  const EngineConfig& c = cfg;
  const int adt = act_dt(), VH = c.vae_hidden, M = n * rows, FF = VH * 3 / 2;
  const size_t s = dt_size(adt);
  WS(nv, char*, (size_t)M * VH * s);
  WS(qkv, char*, (size_t)M * 3 * VH * s);
  WS(f, char*, (size_t)M * FF * s);
  const void* x_op = (adt == DT_F32) ? (const void*)x : (const void*)xa;
  for (int l = 0; l < c.vae_depth; ++l) {
    AT_TRY(launch_layernorm(x, VH, nv, adt, VH, get<float>(S("vae.%s.l%d.ln_g", side, l)), get<float>(S("vae.%s.l%d.ln_b", side, l)),
                            M, VH, 1e-5f, ACT_NONE, st));
    GemmArgs g = gemm_args();
    g.A = nv; g.a_map = plain_rows(VH); g.W = getw(S("vae.%s.l%d.qkv.w", side, l)); g.ldw = VH; g.M = M; g.N = 3 * VH; g.K = VH;
    g.out_act = qkv; g.out_act_dt = adt; g.c_map = plain_rows(3 * VH);
    AT_TRY(gemm(g, st));
    AttnArgs a;     // to_qkv layout (qkv, head, d): q | k | v blocks of 512 columns (quirk 3)
    a.q = qkv; a.k = qkv + (size_t)VH * s; a.v = qkv + (size_t)2 * VH * s; a.out = nv; a.dt = adt;
    a.n_seq = n; a.n_heads = c.vae_heads; a.head_dim = 64; a.lq = rows; a.lk = rows;
    a.q_ss = a.k_ss = a.v_ss = (int64_t)rows * 3 * VH; a.q_rs = a.k_rs = a.v_rs = 3 * VH; a.o_ss = (int64_t)rows * VH; a.o_rs = VH;
    a.scale = 1.0f / sqrtf((float)VH);            // quirk 2: hidden_dim ** -0.5
    a.split = split;
    AT_TRY(attention(a, st));
    g = gemm_args();
    g.A = nv; g.a_map = plain_rows(VH); g.W = getw(S("vae.%s.l%d.out.w", side, l)); g.ldw = VH; g.M = M; g.N = VH; g.K = VH;
    g.bias = get<float>(S("vae.%s.l%d.out.b", side, l)); g.resid = x; g.resid_map = plain_rows(VH); g.out32 = x; g.c_map = plain_rows(VH);
    if (adt != DT_F32) { g.out_act = xa; g.out_act_dt = adt; }
    AT_TRY(gemm(g, st));
    g = gemm_args();      // MLP on the un-normalised stream
    g.A = x_op; g.a_map = plain_rows(VH); g.W = getw(S("vae.%s.l%d.ff1.w", side, l)); g.ldw = VH; g.M = M; g.N = FF; g.K = VH;
    g.bias = get<float>(S("vae.%s.l%d.ff1.b", side, l)); g.act = ACT_GELU_TANH; g.out_act = f; g.out_act_dt = adt; g.c_map = plain_rows(FF);
    AT_TRY(gemm(g, st));
    g = gemm_args();
    g.A = f; g.a_map = plain_rows(FF); g.W = getw(S("vae.%s.l%d.ff2.w", side, l)); g.ldw = FF; g.M = M; g.N = VH; g.K = FF;
    g.bias = get<float>(S("vae.%s.l%d.ff2.b", side, l)); g.resid = x; g.resid_map = plain_rows(VH); g.out32 = x; g.c_map = plain_rows(VH);
    // the bf16 copy of x is read by this layer's FFN1 (written by the out-projection above) and, after the last layer, by the
    // caller's output mapping: only the last FFN2 has to write it (single-output GEMMs take the TMA-store epilogue)
    if (adt != DT_F32 && l == c.vae_depth - 1) { g.out_act = xa; g.out_act_dt = adt; }
    AT_TRY(gemm(g, st));
  }
  return AT_OK;
}

int Engine::vae_decode(const uint32_t* prev_words, const uint32_t* words, int n, float* motion, cudaStream_t st) {
  const EngineConfig& c = cfg;
  const int adt = act_dt(), VH = c.vae_hidden, rows = 2 * T, M = n * rows;
  const size_t s = dt_size(adt);
  size_t mark = ws_off;
  WS(z, char*, (size_t)M * c.code_dim * s);
  WS(x, float*, (size_t)M * VH * 4);
  char* xa = nullptr;
  if (adt != DT_F32) { xa = (char*)ws_alloc((size_t)M * VH * s); if (!xa) return AT_ENOMEM; }
  AT_TRY(launch_bits_latent(tb, prev_words, L, get<float>("vae.dec_pos"), z, adt, n, 0, st));
  AT_TRY(launch_bits_latent(tb, words, L, get<float>("vae.dec_pos"), z, adt, n, 1, st));
  GemmArgs g = gemm_args();
  g.A = z; g.a_map = plain_rows(c.code_dim); g.W = getw("vae.dec.in.w"); g.ldw = c.code_dim; g.M = M; g.N = VH; g.K = c.code_dim;
  g.bias = get<float>("vae.dec.in.b"); g.act = ACT_LEAKY02; g.out32 = x; g.c_map = plain_rows(VH);
  if (adt != DT_F32 && c.vae_depth == 0) { g.out_act = xa; g.out_act_dt = adt; }   // layer 0 rewrites xa before anyone reads it
  AT_TRY(gemm(g, st));
  AT_TRY(vae_stack("dec", n, rows, T, x, xa, st));
  // out_mapping on the new half only (rows T..2T of each clip); motion_std / motion_mean folded into W, b
  g = gemm_args();
  const char* src = (adt == DT_F32) ? (const char*)x : (const char*)xa;
  g.A = src + (size_t)T * VH * s; g.a_map = batched_rows(T, (int64_t)rows * VH, VH);
  g.W = getw("vae.dec.out.w"); g.ldw = VH; g.M = n * T; g.N = c.motion_dim; g.K = VH;
  g.bias = get<float>("vae.dec.out.b"); g.out32 = motion; g.c_map = plain_rows(c.motion_dim);
  AT_TRY(gemm(g, st));
  ws_off = mark;
  return AT_OK;
}

int Engine::vae_encode_bits(const float* motion, int n, uint32_t* words_out, float* enc_out_opt, cudaStream_t st) {
  const EngineConfig& c = cfg;
  const int adt = act_dt(), VH = c.vae_hidden, M = n * T, KP = 128;
  const size_t s = dt_size(adt);
  size_t mark = ws_off;
  WS(xin, char*, (size_t)M * KP * s);
  WS(x, float*, (size_t)M * VH * 4);
  char* xa = nullptr;
  if (adt != DT_F32) { xa = (char*)ws_alloc((size_t)M * VH * s); if (!xa) return AT_ENOMEM; }
  float* enc = enc_out_opt;
  if (!enc) { enc = (float*)ws_alloc((size_t)M * c.code_dim * 4); if (!enc) return AT_ENOMEM; }
  AT_TRY(launch_motion_norm_pos(motion, get<float>("vae.mean"), get<float>("vae.std"), get<float>("vae.enc_pos"), xin, adt, n, T,
                                c.motion_dim, KP, st));
  GemmArgs g = gemm_args();
  g.A = xin; g.a_map = plain_rows(KP); g.W = getw("vae.enc.in.w"); g.ldw = KP; g.M = M; g.N = VH; g.K = KP;
  g.bias = get<float>("vae.enc.in.b"); g.act = ACT_LEAKY02; g.out32 = x; g.c_map = plain_rows(VH);
  if (adt != DT_F32 && c.vae_depth == 0) { g.out_act = xa; g.out_act_dt = adt; }
  AT_TRY(gemm(g, st));
  AT_TRY(vae_stack("enc", n, T, 0, x, xa, st));
  g = gemm_args();
  g.A = (adt == DT_F32) ? (const void*)x : (const void*)xa; g.a_map = plain_rows(VH);
  g.W = getw("vae.enc.out.w"); g.ldw = VH; g.M = M; g.N = c.code_dim; g.K = VH;
  g.bias = get<float>("vae.enc.out.b"); g.out32 = enc; g.c_map = plain_rows(c.code_dim);
  AT_TRY(gemm(g, st));
  AT_TRY(launch_bsq_quantize(tb, enc, words_out, L, n, st));
  ws_off = mark;
  return AT_OK;
}

// ------------------------------------------------------------------ CUDA-graph replay of a fixed-address body
// `body` only touches fixed workspace addresses (everything it allocates comes from the workspace after `body_mark`, a
// deterministic function of `key`), so after one eager warm-up per key it is captured and replayed. The caller's stream may be
// the legacy default stream, which cannot be captured: capture and replay run on an internal stream fenced against the
// caller's stream with events. A graph is stale when the workspace was re-allocated (addresses) or a process-wide option
// changed which kernels a launch selects. A failed capture / instantiation switches the engine to eager launches and says so.
int Engine::run_graphed(int key, cudaStream_t st, size_t body_mark, const std::function<int(cudaStream_t)>& body) {
  if (use_graphs && !prof && !g_trace_on) {
    auto it = graphs.find(key);
    if (it != graphs.end() && (it->second.ws_base != ws || it->second.epoch != g_option_epoch.load())) {
      if (it->second.exec) cudaGraphExecDestroy(it->second.exec);
      graphs.erase(it);
      it = graphs.end();
    }
    if (it == graphs.end()) {
      GraphEntry ge; ge.exec = nullptr; ge.ws_base = ws; ge.warm = 0; ge.n_launches = 0; ge.epoch = g_option_epoch.load();
      it = graphs.insert(std::make_pair(key, ge)).first;
    }
    GraphEntry& ge = it->second;
    if (!gstream) {
      AT_CUDA(cudaStreamCreateWithFlags(&gstream, cudaStreamNonBlocking));
      AT_CUDA(cudaEventCreateWithFlags(&gev_in, cudaEventDisableTiming));
      AT_CUDA(cudaEventCreateWithFlags(&gev_out, cudaEventDisableTiming));
    }
    if (!ge.exec && ge.warm >= 1) {
      // second call for this key: capture the body
      cudaGraph_t graph = nullptr;
      AT_CUDA(cudaStreamBeginCapture(gstream, cudaStreamCaptureModeThreadLocal));
      const int rc = body(gstream);
      const cudaError_t ce = cudaStreamEndCapture(gstream, &graph);
      ws_off = body_mark;
      if (rc != AT_OK || ce != cudaSuccess || !graph) {
        const cudaError_t sticky = cudaGetLastError();
        if (graph) cudaGraphDestroy(graph);
        use_graphs = false;                                          // eager launches from here on, and say so
        graph_failure = S("capture of graph %d failed (body status %d, cudaStreamEndCapture: %s / %s)", key, rc,
                          cudaGetErrorString(ce), cudaGetErrorString(sticky));
        fprintf(stderr, "[artalk_b200] WARNING: %s; this engine now launches eagerly (slower)\n", graph_failure.c_str());
        if (rc != AT_OK) return rc;
      } else {
        if (getenv("ARTALK_DEBUG")) {
          size_t ne = 0;
          if (cudaGraphGetEdges_v2(graph, nullptr, nullptr, nullptr, &ne) == cudaSuccess && ne > 0) {
            std::vector<cudaGraphNode_t> from(ne), to(ne);
            std::vector<cudaGraphEdgeData> ed(ne);
            size_t prog = 0;
            if (cudaGraphGetEdges_v2(graph, from.data(), to.data(), ed.data(), &ne) == cudaSuccess)
              for (size_t i = 0; i < ne; ++i) prog += ed[i].type == cudaGraphDependencyTypeProgrammatic;
            fprintf(stderr, "[artalk] captured graph %d: %zu edges, %zu programmatic (PDL)\n", key, ne, prog);
          }
          cudaGetLastError();
        }
        const cudaError_t ie = cudaGraphInstantiate(&ge.exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ie != cudaSuccess) {
          cudaGetLastError();
          ge.exec = nullptr; use_graphs = false;
          graph_failure = S("cudaGraphInstantiate of graph %d failed: %s", key, cudaGetErrorString(ie));
          fprintf(stderr, "[artalk_b200] WARNING: %s; this engine now launches eagerly (slower)\n", graph_failure.c_str());
        }
      }
    }
    if (ge.exec) {
      AT_CUDA(cudaEventRecord(gev_in, st));
      AT_CUDA(cudaStreamWaitEvent(gstream, gev_in, 0));
      AT_CUDA(cudaGraphLaunch(ge.exec, gstream));
      ++graph_replays;
      AT_CUDA(cudaEventRecord(gev_out, gstream));
      AT_CUDA(cudaStreamWaitEvent(st, gev_out, 0));
      g_launch_count.fetch_add(ge.n_launches, std::memory_order_relaxed);
      return AT_OK;
    }
    if (use_graphs) {
      const unsigned long long l0 = g_launch_count.load();
      AT_TRY(body(st));
      ge.n_launches = g_launch_count.load() - l0;
      ge.warm++;
      return AT_OK;
    }
  }
  return body(st);
}

// ------------------------------------------------------------------ one 100-frame chunk of the AR recurrence
// The chunk body only reads/writes fixed workspace addresses, so after one eager warm-up per (clips, teacher forcing)
// it is captured into a CUDA graph and replayed: ~700 launches per chunk become one graph launch (launch gaps dominate
// the small GEMMs of the recurrence, and batch-1 latency). Inputs/outputs are staged with small D2D copies.
int Engine::ar_chunk(int B, const float* cond, int64_t cond_cs, const float* style, uint32_t* prev_words, float* motion_out,
                     uint32_t* words_out, float* logits_out, const uint32_t* forced_words, float* enc_out, cudaStream_t st) {
  AT_REQUIRE(finalized, "engine not finalized");
  if (B <= 0) return AT_OK;
  const EngineConfig& c = cfg;
  const int adt = act_dt(), C = c.embed_dim, D = c.cond_dim, NL = c.ar_depth;
  const size_t s = dt_size(adt);
  const int n_ada = NL * 6 * C + 2 * C, KV = 2 * L, Tm = T;
  const int64_t BL = (int64_t)B * L;
  size_t need = (size_t)BL * D * s + (size_t)BL * n_ada * s + (size_t)BL * C * s + (size_t)BL * NL * 2 * C * s +
                2 * (size_t)NL * B * KV * C * s + (size_t)B * Tm * C * (8 + 3 * s) + (size_t)B * Tm * 4 * C * s +
                (size_t)BL * 2 * c.code_dim * 4 + (size_t)BL * 4 * 3 + (size_t)B * C * 4 +
                (size_t)B * Tm * (c.motion_dim + c.code_dim) * 4 +
                (size_t)B * 2 * Tm * (c.code_dim * s + c.vae_hidden * (4 + s) + c.vae_hidden * s * 4 + c.vae_hidden * s * 3 / 2) +
                (size_t)B * Tm * (128 * s + c.code_dim * 4) + (4 << 20);
  AT_TRY(ws_reserve(need, st));
  // staging buffers (fixed offsets at the head of the workspace)
  WS(scond, char*, (size_t)BL * D * s);
  WS(style_ws, float*, (size_t)B * C * 4);
  WS(prev_ws, uint32_t*, (size_t)BL * 4);
  WS(forced_ws, uint32_t*, (size_t)BL * 4);
  WS(words_ws, uint32_t*, (size_t)BL * 4);
  WS(logits_ws, float*, (size_t)BL * 2 * c.code_dim * 4);
  WS(motion_ws, float*, (size_t)B * Tm * c.motion_dim * 4);
  WS(enc_ws, float*, (size_t)B * Tm * c.code_dim * 4);
  const size_t body_mark = ws_off;
  AT_TRY(launch_act_cast(cond, batched_rows(L, cond_cs, D), scond, adt, (int)BL, D, ACT_SILU, st));
  AT_CUDA(cudaMemcpyAsync(style_ws, style, (size_t)B * C * 4, cudaMemcpyDeviceToDevice, st));
  AT_CUDA(cudaMemcpyAsync(prev_ws, prev_words, (size_t)BL * 4, cudaMemcpyDeviceToDevice, st));
  if (forced_words) AT_CUDA(cudaMemcpyAsync(forced_ws, forced_words, (size_t)BL * 4, cudaMemcpyDeviceToDevice, st));

  AT_TRY(run_graphed(B * 2 + (forced_words ? 1 : 0), st, body_mark, [&](cudaStream_t s) {
    return ar_chunk_body(B, scond, style_ws, prev_ws, motion_ws, words_ws, logits_ws, forced_words ? forced_ws : nullptr, enc_ws, s);
  }));
  AT_CUDA(cudaMemcpyAsync(motion_out, motion_ws, (size_t)B * Tm * c.motion_dim * 4, cudaMemcpyDeviceToDevice, st));
  AT_CUDA(cudaMemcpyAsync(prev_words, prev_ws, (size_t)BL * 4, cudaMemcpyDeviceToDevice, st));
  if (words_out) AT_CUDA(cudaMemcpyAsync(words_out, words_ws, (size_t)BL * 4, cudaMemcpyDeviceToDevice, st));
  if (logits_out) AT_CUDA(cudaMemcpyAsync(logits_out, logits_ws, (size_t)BL * 2 * c.code_dim * 4, cudaMemcpyDeviceToDevice, st));
  if (enc_out) AT_CUDA(cudaMemcpyAsync(enc_out, enc_ws, (size_t)B * Tm * c.code_dim * 4, cudaMemcpyDeviceToDevice, st));
  return AT_OK;
}

// all pointer arguments live in the workspace (see ar_chunk); allocations below are a deterministic function of B
int Engine::ar_chunk_body(int B, const char* scond, const float* style, uint32_t* prev_words, float* motion_out, uint32_t* words,
                          float* logits, const uint32_t* forced_words, float* enc_out, cudaStream_t st) {
  const EngineConfig& c = cfg;
  const int adt = act_dt(), C = c.embed_dim, D = c.cond_dim, NL = c.ar_depth;
  const size_t s = dt_size(adt);
  const int n_ada = NL * 6 * C + 2 * C, P = L, KV = P + L, Tm = T;
  const int64_t BL = (int64_t)B * L;
  WS(ada, char*, (size_t)BL * n_ada * s);
  WS(prev_tok, char*, (size_t)BL * C * s);
  WS(kvtmp, char*, (size_t)BL * NL * 2 * C * s);            // prev K|V of all layers; later reused as per-step qkv
  WS(kcache, char*, (size_t)NL * B * KV * C * s);
  WS(vcache, char*, (size_t)NL * B * KV * C * s);
  WS(x, float*, (size_t)B * Tm * C * 4);
  WS(u, char*, (size_t)B * Tm * C * s);
  WS(qbuf, char*, (size_t)B * Tm * C * s);
  WS(o, char*, (size_t)B * Tm * C * s);
  WS(f, char*, (size_t)B * Tm * 4 * C * s);
  // bf16 mode: the projection / FFN2 GEMMs write y = A W^T + b (fp32) with the plain epilogue and the gated residual update
  // x += gamma * y is folded into the next AdaLN kernel (bit-identical, see norms.cu); fp32 mode keeps the fused epilogue
  const bool defer = (adt == DT_BF16);
  float* ybuf = nullptr;
  if (defer) { ybuf = (float*)ws_alloc((size_t)B * Tm * C * 4); if (!ybuf) return AT_ENOMEM; }

  // AdaLN parameters of every block + head for all 181 tokens, once per chunk (audio-only)
  GemmArgs g = gemm_args();
  g.A = scond; g.a_map = plain_rows(D); g.W = getw("ar.ada.w"); g.ldw = D; g.M = (int)BL; g.N = n_ada; g.K = D;
  g.bias = get<float>("ar.ada.b"); g.out_act = ada; g.out_act_dt = adt; g.c_map = plain_rows(n_ada);
  AT_TRY(gemm(g, st));
  // previous-chunk tokens (+ prev_lvl_pos) and their K/V for all blocks (same raw input for every block, quirk 8)
  AT_TRY(launch_bits_tokens(tb, prev_words, L, style, get<float>("ar.embed.w"), get<float>("ar.embed.b"), get<float>("ar.prev_lvl_pos"),
                            prev_tok, adt, B, 0, c.n_levels - 1, C, st));
  g = gemm_args();
  g.A = prev_tok; g.a_map = plain_rows(C); g.W = getw("ar.prevkv.w"); g.ldw = C; g.M = (int)BL; g.N = NL * 2 * C; g.K = C;
  g.bias = get<float>("ar.prevkv.b");
  const bool fused_qkv = (adt == DT_BF16);      // tcgen05 epilogue normalises k heads and writes the caches directly
  if (fused_qkv) {
    g.qkv_mode = 2; g.qkv_C = C; g.kcache = kcache; g.vcache = vcache; g.kv_map = batched_rows(P, (int64_t)KV * C, C);
    g.kv_layer_stride = (int64_t)B * KV * C;
    AT_TRY(gemm(g, st));
  } else {
    g.out_act = kvtmp; g.out_act_dt = adt; g.c_map = plain_rows(NL * 2 * C);
    AT_TRY(gemm(g, st));
    for (int l = 0; l < NL; ++l)
      AT_TRY(launch_qkv_norm_scatter(kvtmp + (size_t)l * 2 * C * s, adt, (int64_t)NL * 2 * C, 0, nullptr, nullptr,
                                     kcache + (size_t)l * B * KV * C * s, vcache + (size_t)l * B * KV * C * s,
                                     batched_rows(P, (int64_t)KV * C, C), (int)BL, c.ar_heads, st));
  }
  for (int p = 0; p < c.n_levels; ++p) {
    const int n_new = c.patch_nums[p], off = p ? tb.cum[p - 1] : 0, M = B * n_new;
    const int sk = n_new <= g_skinny_tokens ? 1 : 0;           // per-clip criterion: batch-size independent arithmetic
    const RowMap ada_map = batched_rows(n_new, (int64_t)L * n_ada, n_ada);
    const char* ada_p = ada + (size_t)off * n_ada * s;
    const uint32_t* src_words = forced_words ? forced_words : words;
    AT_TRY(launch_bits_tokens(tb, src_words, L, style, get<float>("ar.embed.w"), get<float>("ar.embed.b"), get<float>("ar.lvl_pos"), x,
                              DT_F32, B, p, p, C, st));
    bool pending = false;                                    // ybuf holds the previous layer's FFN2 output, gate gamma2 of that layer
    for (int l = 0; l < NL; ++l) {
      const char* ada_l = ada_p + (size_t)l * 6 * C * s;     // chunk order: g1, g2, s1, s2, b1, b2 (quirk 9)
      AT_TRY(launch_adaln_modulate(x, ada_l, adt, ada_map, 2 * C, 4 * C, u, adt, M, C, 1e-6f, st, pending ? ybuf : nullptr, -5 * C));
      pending = false;
      g = gemm_args();
      g.A = u; g.a_map = plain_rows(C); g.W = getw(S("ar.l%d.qkv.w", l)); g.ldw = C; g.M = M; g.N = 3 * C; g.K = C;
      g.bias = get<float>(S("ar.l%d.qkv.b", l)); g.skinny = sk;
      char* kc = kcache + (size_t)l * B * KV * C * s; char* vc = vcache + (size_t)l * B * KV * C * s;
      if (fused_qkv) {
        g.qkv_mode = 1; g.qkv_C = C; g.head_scale = get<float>(S("ar.l%d.head_scale", l)); g.qbuf = qbuf;
        g.kcache = kc + (size_t)(P + off) * C * s; g.vcache = vc + (size_t)(P + off) * C * s;
        g.kv_map = batched_rows(n_new, (int64_t)KV * C, C);
        AT_TRY(gemm(g, st));
      } else {
        g.out_act = kvtmp; g.out_act_dt = adt; g.c_map = plain_rows(3 * C);
        AT_TRY(gemm(g, st));
        AT_TRY(launch_qkv_norm_scatter(kvtmp, adt, 3 * C, 1, get<float>(S("ar.l%d.head_scale", l)), qbuf, kc + (size_t)(P + off) * C * s,
                                       vc + (size_t)(P + off) * C * s, batched_rows(n_new, (int64_t)KV * C, C), M, c.ar_heads, st));
      }
      AttnArgs a;
      a.q = qbuf; a.k = kc; a.v = vc; a.out = o; a.dt = adt; a.n_seq = B; a.n_heads = c.ar_heads; a.head_dim = 64;
      a.lq = n_new; a.lk = P + off + n_new;          // prev chunk + every current token of scale <= p
      a.q_ss = (int64_t)n_new * C; a.q_rs = C; a.k_ss = a.v_ss = (int64_t)KV * C; a.k_rs = a.v_rs = C;
      a.o_ss = (int64_t)n_new * C; a.o_rs = C; a.scale = 1.0f; a.split = 0;
      // |q.k| <= head_scale[h] (q and k are L2-normalised per head): the tcgen05 kernel subtracts that bound instead of the row
      // maximum for every head whose scale is small enough (<= 32) and skips its max pass over S
      if (g_attn_bound) a.key_bound = get<float>(S("ar.l%d.head_scale", l));
      AT_TRY(attention(a, st));
      g = gemm_args();
      g.A = o; g.a_map = plain_rows(C); g.W = getw(S("ar.l%d.proj.w", l)); g.ldw = C; g.M = M; g.N = C; g.K = C;
      g.bias = get<float>(S("ar.l%d.proj.b", l)); g.c_map = plain_rows(C); g.skinny = sk;
      if (defer) g.out32 = ybuf;
      else { g.gate = ada_l; g.gate_dt = adt; g.gate_map = ada_map; g.resid = x; g.resid_map = plain_rows(C); g.out32 = x; }   // gamma1
      AT_TRY(gemm(g, st));
      AT_TRY(launch_adaln_modulate(x, ada_l, adt, ada_map, 3 * C, 5 * C, u, adt, M, C, 1e-6f, st, defer ? ybuf : nullptr, 0));
      g = gemm_args();
      g.A = u; g.a_map = plain_rows(C); g.W = getw(S("ar.l%d.ff1.w", l)); g.ldw = C; g.M = M; g.N = 4 * C; g.K = C;
      g.bias = get<float>(S("ar.l%d.ff1.b", l)); g.act = ACT_GELU_TANH; g.out_act = f; g.out_act_dt = adt; g.c_map = plain_rows(4 * C);
      g.skinny = sk;
      AT_TRY(gemm(g, st));
      g = gemm_args();
      g.A = f; g.a_map = plain_rows(4 * C); g.W = getw(S("ar.l%d.ff2.w", l)); g.ldw = 4 * C; g.M = M; g.N = C; g.K = 4 * C;
      g.bias = get<float>(S("ar.l%d.ff2.b", l)); g.c_map = plain_rows(C); g.skinny = sk;
      if (defer) { g.out32 = ybuf; pending = true; }
      else { g.gate = ada_l + (size_t)C * s; g.gate_dt = adt; g.gate_map = ada_map; g.resid = x; g.resid_map = plain_rows(C); g.out32 = x; }   // gamma2
      AT_TRY(gemm(g, st));
    }
    // head: AdaLN (scale, shift order) -> Linear 768 -> 64 -> pairwise argmax (app/models.py:103-104,145-148)
    const char* ada_h = ada_p + (size_t)NL * 6 * C * s;
    AT_TRY(launch_adaln_modulate(x, ada_h, adt, ada_map, 0, C, u, adt, M, C, 1e-6f, st, pending ? ybuf : nullptr, -5 * C));
    g = gemm_args();
    g.A = u; g.a_map = plain_rows(C); g.W = getw("ar.head.w"); g.ldw = C; g.M = M; g.N = 2 * c.code_dim; g.K = C;
    g.bias = get<float>("ar.head.b"); g.out32 = logits + (size_t)off * 2 * c.code_dim;
    g.c_map = batched_rows(n_new, (int64_t)L * 2 * c.code_dim, 2 * c.code_dim); g.skinny = sk;
    AT_TRY(gemm(g, st));
    AT_TRY(launch_argmax_bits(logits + (size_t)off * 2 * c.code_dim, g.c_map, words + off, batched_rows(n_new, L, 1), M, st));
  }
  // VAE decode with the *re-encoded* prev bits (quirk 7), then re-encode the prediction for the next chunk
  AT_TRY(vae_decode(prev_words, forced_words ? forced_words : words, B, motion_out, st));
  AT_TRY(vae_encode_bits(motion_out, B, prev_words, enc_out, st));
  return AT_OK;
}

}  // namespace artalk
