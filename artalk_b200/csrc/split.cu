// Operand splitting for the parity-grade tensor-core mode (precision "bf16x3" / "bf16x6").
//
// The reference computes in fp32 (app/models.py:62-121 under torch defaults); its sampled bits are decisions on logit
// differences as small as 1e-4, which plain bf16 operands (2^-9 relative rounding) cannot reproduce. In this mode every fp32
// operand x of a tensor-core GEMM is written as a sum of bf16 pieces
//     p0 = bf16(x),  p1 = bf16(x - p0),  p2 = bf16(x - p0 - p1)            (the subtractions are exact in fp32)
// and the product A W^T is accumulated in fp32 (TMEM) over the piece pairs whose magnitude matters:
//     2 pieces / 3 MMA passes: a0 w0 + a1 w0 + a0 w1                       (relative error ~2^-17 per product)
//     3 pieces / 6 MMA passes: ... + a1 w1 + a2 w0 + a0 w2                 (relative error ~2^-24: fp32 grade)
// The passes are laid out along K so that the UNCHANGED tcgen05 GEMM kernels (gemm_tc.cu) run them as one GEMM with
// K' = slots * K: every 64-element K block of the operand becomes `slots` consecutive 64-element blocks holding the piece that
// slot multiplies (A: [p0 p1 p0 | p1 p2 p0], W: [p0 p0 p1 | p1 p0 p2]). Because the transformation is elementwise on 64-element
// blocks of the contiguous buffer, batched row views and the overlapping-row implicit-GEMM windows of the wav2vec conv layers
// keep working with every stride multiplied by `slots`.
#include "kernels.cuh"

namespace artalk {

namespace {

template <int S> __device__ __forceinline__ int slot_piece(bool is_w, int s) {
  // A: 0 1 0 | 1 2 0     W: 0 0 1 | 1 0 2
  if (S == 3) return is_w ? (s == 2 ? 1 : 0) : (s == 1 ? 1 : 0);
  if (is_w) return s == 2 ? 1 : (s == 3 ? 1 : (s == 5 ? 2 : 0));
  return s == 1 ? 1 : (s == 3 ? 1 : (s == 4 ? 2 : 0));
}

__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// one thread = 8 consecutive elements (32 B in, S x 16 B out)
template <int S>
__global__ void __launch_bounds__(256) split_bf16_kernel(const float* __restrict__ x, bf16* __restrict__ out, int64_t n8, int is_w) {
  pdl_enter();                       // the output buffer may still be read by the previous GEMM (write-after-read)
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(x) + 2 * i), b = __ldg(reinterpret_cast<const float4*>(x) + 2 * i + 1);
    const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    float p[3][8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float p0 = __bfloat162float(__float2bfloat16_rn(v[e]));
      const float r1 = v[e] - p0;
      const float p1 = __bfloat162float(__float2bfloat16_rn(r1));
      const float r2 = r1 - p1;
      p[0][e] = p0; p[1][e] = p1; p[2][e] = r2;                 // the last piece is rounded when packed
    }
    const int64_t blk = i >> 3;
    const int j8 = (int)(i & 7);
    uint4* dst = reinterpret_cast<uint4*>(out + (blk * S) * 64 + j8 * 8);
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const int pc = slot_piece<S>(is_w != 0, s);
      uint4 pk;
      pk.x = pack2(p[pc][0], p[pc][1]); pk.y = pack2(p[pc][2], p[pc][3]);
      pk.z = pack2(p[pc][4], p[pc][5]); pk.w = pack2(p[pc][6], p[pc][7]);
      dst[s * 8] = pk;                                            // slot s of this 64-block: + s * 64 elements = 8 x 16 B
    }
  }
}

// one thread = 8 consecutive elements of a row; planes are compact [n_seq][rows][cols]
__global__ void __launch_bounds__(256) split2_rows_kernel(const float* __restrict__ x, int64_t ss, int64_t rs, int rows, int cols8,
                                                          bf16* __restrict__ out, int64_t plane, int64_t total8) {
  pdl_enter();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total8; i += (int64_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % cols8);
    const int64_t r = i / cols8;
    const int row = (int)(r % rows);
    const int64_t seq = r / rows;
    const float* src = x + seq * ss + (int64_t)row * rs + c8 * 8;
    const float4 a = __ldg(reinterpret_cast<const float4*>(src)), b = __ldg(reinterpret_cast<const float4*>(src) + 1);
    const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    float hi[8], lo[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { hi[e] = __bfloat162float(__float2bfloat16_rn(v[e])); lo[e] = v[e] - hi[e]; }
    uint4 ph, pl;
    ph.x = pack2(hi[0], hi[1]); ph.y = pack2(hi[2], hi[3]); ph.z = pack2(hi[4], hi[5]); ph.w = pack2(hi[6], hi[7]);
    pl.x = pack2(lo[0], lo[1]); pl.y = pack2(lo[2], lo[3]); pl.z = pack2(lo[4], lo[5]); pl.w = pack2(lo[6], lo[7]);
    *reinterpret_cast<uint4*>(out + i * 8) = ph;
    *reinterpret_cast<uint4*>(out + plane + i * 8) = pl;
  }
}

}  // namespace

int launch_split2_rows(const float* x, int64_t ss, int64_t rs, int n_seq, int rows, int cols, void* out, cudaStream_t st) {
  if (n_seq <= 0 || rows <= 0) return AT_OK;
  AT_REQUIRE(x && out && cols % 8 == 0 && ss % 4 == 0 && rs % 4 == 0 && ((uintptr_t)x % 16 == 0) && ((uintptr_t)out % 16 == 0),
             "split2_rows: bad argument (cols=%d)", cols);
  const int64_t plane = (int64_t)n_seq * rows * cols, total8 = plane / 8;
  int64_t blocks = (total8 + 255) / 256;
  const DevCtx* dc = nullptr;
  AT_TRY(dev_ctx(&dc));
  if (blocks > (int64_t)dc->num_sms * 16) blocks = (int64_t)dc->num_sms * 16;
  AT_CUDA(launch_k(split2_rows_kernel, dim3((unsigned)blocks), dim3(256), 0, st, x, ss, rs, rows, cols / 8, (bf16*)out, plane, total8));
  AT_LAUNCH_CHECK();
  return AT_OK;
}

// x [n_elems] fp32 (n_elems % 64 == 0, 16-byte aligned) -> out [n_elems / 64][slots][64] bf16; slots = 3 or 6
int launch_split_bf16(const float* x, void* out, int64_t n_elems, int slots, int is_w, cudaStream_t st) {
  if (n_elems <= 0) return AT_OK;
  AT_REQUIRE(x && out && n_elems % 64 == 0 && (slots == 3 || slots == 6), "split_bf16: bad argument (n=%lld slots=%d)",
             (long long)n_elems, slots);
  AT_REQUIRE(((uintptr_t)x % 16 == 0) && ((uintptr_t)out % 16 == 0), "split_bf16: operands must be 16-byte aligned");
  const int64_t n8 = n_elems / 8;
  int64_t blocks = (n8 + 255) / 256;
  const DevCtx* dc = nullptr;
  AT_TRY(dev_ctx(&dc));
  if (blocks > (int64_t)dc->num_sms * 16) blocks = (int64_t)dc->num_sms * 16;
  if (slots == 3) AT_CUDA(launch_k(split_bf16_kernel<3>, dim3((unsigned)blocks), dim3(256), 0, st, x, (bf16*)out, n8, is_w));
  else AT_CUDA(launch_k(split_bf16_kernel<6>, dim3((unsigned)blocks), dim3(256), 0, st, x, (bf16*)out, n8, is_w));
  AT_LAUNCH_CHECK();
  return AT_OK;
}

}  // namespace artalk
