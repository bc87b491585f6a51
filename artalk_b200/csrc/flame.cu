// FLAME blendshapes + linear blend skinning (app/flame_model/FLAME.py:117-149, app/flame_model/lbs.py:142-383).
// Two kernels:
//   flame_coef_kernel  : one warp per frame — betas, joints (from precomputed J_template / J_dirs), Rodrigues,
//                        pose feature (R[1:] - I), kinematic chain, relative transforms A (5 x 3x4).
//   flame_verts_kernel : vertex-parallel blend (v_template + sum_l coef_l * dirs_l) and skinning; a block owns
//                        128 vertices x FB frames so each basis value loaded from L2 is reused FB times.
// When every frame shares one shape row (the mesh path, inference.py:64) the 300 shape bases are folded into a
// per-call static template first, leaving 136 bases (100 expression + 36 pose) per frame.
#include "kernels.cuh"

namespace artalk {

namespace {
constexpr int FB = 16;            // frames per block
constexpr int NJ = 5;
constexpr int COEF_A = 60;        // 5 joints x 12

struct FlameDev {
  int V, n_shape, n_exp, n_bases;           // n_bases = n_shape + n_exp + 36
  const float* v_template; const float* dirs; const float* j_template; const float* j_dirs; const float* lbs_w;
  int parents[NJ];
  float scale;
};

__device__ void rodrigues(const float r[3], float R[9]) {
  float ax = r[0] + 1e-8f, ay = r[1] + 1e-8f, az = r[2] + 1e-8f;      // lbs.py:294
  float angle = sqrtf(ax * ax + ay * ay + az * az);
  float dx = r[0] / angle, dy = r[1] / angle, dz = r[2] / angle;
  float s = sinf(angle), c = cosf(angle);
  float K[9] = {0.f, -dz, dy, dz, 0.f, -dx, -dy, dx, 0.f};
  float KK[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) KK[i * 3 + j] = K[i * 3] * K[j] + K[i * 3 + 1] * K[3 + j] + K[i * 3 + 2] * K[6 + j];
#pragma unroll
  for (int i = 0; i < 9; ++i) R[i] = ((i % 4 == 0) ? 1.f : 0.f) + s * K[i] + (1.f - c) * KK[i];
}

// coef row layout: [0, n_shape+n_exp) betas | 36 pose feature | 60 A   (stride n_bases + 60)
__global__ void __launch_bounds__(128) flame_coef_kernel(FlameDev fm, const float* __restrict__ shape, int64_t shape_rs,
                                                         const float* __restrict__ expr, int64_t expr_rs,
                                                         const float* __restrict__ pose, int64_t pose_rs, int zero_global,
                                                         float* __restrict__ coef, int n_frames) {
  int f = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (f >= n_frames) return;
  const int nb = fm.n_shape + fm.n_exp, stride = fm.n_bases + COEF_A;
  float* cf = coef + (int64_t)f * stride;
  float jacc[NJ * 3];
#pragma unroll
  for (int i = 0; i < NJ * 3; ++i) jacc[i] = 0.f;
  for (int l = lane; l < nb; l += 32) {
    float b = (l < fm.n_shape) ? shape[(int64_t)f * shape_rs + l] : expr[(int64_t)f * expr_rs + (l - fm.n_shape)];
    cf[l] = b;
#pragma unroll
    for (int i = 0; i < NJ * 3; ++i) jacc[i] = fmaf(b, __ldg(fm.j_dirs + i * nb + l), jacc[i]);   // [15][nb]: coalesced over l
  }
#pragma unroll
  for (int i = 0; i < NJ * 3; ++i) jacc[i] = warp_sum(jacc[i]) + fm.j_template[i];
  // hand the joints to flame_pose_kernel (one thread per frame) through the first 15 slots of the A region
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NJ * 3; ++i) cf[fm.n_bases + i] = jacc[i];
  }
}

// one thread per frame: Rodrigues x5, pose feature (R[1:] - I), kinematic chain, relative transforms A
__global__ void __launch_bounds__(128) flame_pose_kernel(FlameDev fm, const float* __restrict__ pose, int64_t pose_rs,
                                                         int zero_global, float* __restrict__ coef, int n_frames) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= n_frames) return;
  const int nb = fm.n_shape + fm.n_exp, stride = fm.n_bases + COEF_A;
  float* cf = coef + (int64_t)f * stride;
  float jacc[NJ * 3];
#pragma unroll
  for (int i = 0; i < NJ * 3; ++i) jacc[i] = cf[fm.n_bases + i];
  const float* pp = pose + (int64_t)f * pose_rs;
  float rv[NJ][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
  if (!zero_global) { rv[0][0] = pp[0]; rv[0][1] = pp[1]; rv[0][2] = pp[2]; }
  rv[2][0] = pp[3]; rv[2][1] = pp[4]; rv[2][2] = pp[5];                 // FLAME.py:137-141: global, neck=0, jaw, eyes=0
  float R[NJ][9];
#pragma unroll
  for (int j = 0; j < NJ; ++j) rodrigues(rv[j], R[j]);
#pragma unroll
  for (int j = 1; j < NJ; ++j)
#pragma unroll
    for (int i = 0; i < 9; ++i) cf[nb + (j - 1) * 9 + i] = R[j][i] - ((i % 4 == 0) ? 1.f : 0.f);
  // kinematic chain: G_j = G_parent * [R_j | J_j - J_parent]   (lbs.py:326-383)
  float G[NJ][12];
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    int par = fm.parents[j];
    float t[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) t[k] = jacc[j * 3 + k] - (par >= 0 ? jacc[par * 3 + k] : 0.f);
    if (par < 0) {
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        G[j][r * 4 + 0] = R[j][r * 3 + 0]; G[j][r * 4 + 1] = R[j][r * 3 + 1]; G[j][r * 4 + 2] = R[j][r * 3 + 2];
        G[j][r * 4 + 3] = t[r];
      }
    } else {
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const float* gp = &G[par][r * 4];
#pragma unroll
        for (int c = 0; c < 3; ++c) G[j][r * 4 + c] = gp[0] * R[j][c] + gp[1] * R[j][3 + c] + gp[2] * R[j][6 + c];
        G[j][r * 4 + 3] = gp[0] * t[0] + gp[1] * t[1] + gp[2] * t[2] + gp[3];
      }
    }
  }
  // A_j = G_j - [0 | G_j (J_j, 0)]
#pragma unroll
  for (int j = 0; j < NJ; ++j)
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      float* a = cf + fm.n_bases + j * 12 + r * 4;
      const float* g = &G[j][r * 4];
      a[0] = g[0]; a[1] = g[1]; a[2] = g[2];
      a[3] = g[3] - (g[0] * jacc[j * 3] + g[1] * jacc[j * 3 + 1] + g[2] * jacc[j * 3 + 2]);
    }
}

// static template for a shared shape row: out[e] = v_template[e] + sum_{l < n_l} beta[l] * dirs[l][e].
// 64 elements x 4 basis groups per block (coalesced over e, 4-way split of the serial basis loop, smem reduce).
__global__ void __launch_bounds__(256) flame_static_kernel(const float* __restrict__ v_template, const float* __restrict__ dirs,
                                                           const float* __restrict__ beta, int n_l, int n_e, float* __restrict__ out) {
  __shared__ float part[4][64];
  const int el = threadIdx.x & 63, grp = threadIdx.x >> 6, e = blockIdx.x * 64 + el;
  float acc = 0.f;
  if (e < n_e) {
    const int per = (n_l + 3) / 4, l0 = grp * per, l1 = min(n_l, l0 + per);
#pragma unroll 5
    for (int l = l0; l < l1; ++l) acc = fmaf(__ldg(beta + l), __ldg(dirs + (int64_t)l * n_e + e), acc);
  }
  part[grp][el] = acc;
  __syncthreads();
  if (grp == 0 && e < n_e) out[e] = v_template[e] + ((part[0][el] + part[1][el]) + (part[2][el] + part[3][el]));
}

// out[f][v][:] : SKIN ? scale * (sum_j w[v][j] A[f][j]) (base + blend, 1) : base + blend
template <bool SKIN>
__global__ void __launch_bounds__(128) flame_verts_kernel(FlameDev fm, const float* __restrict__ base, int l_begin, int l_end,
                                                          const float* __restrict__ coef, int coef_stride,
                                                          float* __restrict__ verts, int n_frames) {
  extern __shared__ __align__(16) float sm[];
  float* cs = sm;                                   // [n_l][FB]
  float* As = sm + (size_t)(l_end - l_begin) * FB;  // [FB][60]
  const int tid = threadIdx.x, v = blockIdx.y * 128 + tid, f0 = blockIdx.x * FB;
  const int n_l = l_end - l_begin;
  for (int i = tid; i < n_l * FB; i += 128) {
    int l = i / FB, f = i - l * FB;
    cs[i] = (f0 + f < n_frames) ? coef[(int64_t)(f0 + f) * coef_stride + l_begin + l] : 0.f;
  }
  if (SKIN)
    for (int i = tid; i < FB * COEF_A; i += 128) {
      int f = i / COEF_A, k = i - f * COEF_A;
      As[i] = (f0 + f < n_frames) ? coef[(int64_t)(f0 + f) * coef_stride + fm.n_bases + k] : 0.f;
    }
  __syncthreads();
  if (v >= fm.V) return;
  float acc[FB][3];
  float b0 = base[v * 3], b1 = base[v * 3 + 1], b2 = base[v * 3 + 2];
#pragma unroll
  for (int f = 0; f < FB; ++f) { acc[f][0] = b0; acc[f][1] = b1; acc[f][2] = b2; }
  const float* d = fm.dirs + (int64_t)l_begin * fm.V * 3 + v * 3;
  const int64_t ds = (int64_t)fm.V * 3;
#pragma unroll 2
  for (int l = 0; l < n_l; ++l) {
    float d0 = d[0], d1 = d[1], d2 = d[2];
    d += ds;
    const float4* c4 = reinterpret_cast<const float4*>(cs + l * FB);
#pragma unroll
    for (int q = 0; q < FB / 4; ++q) {
      float4 c = c4[q];
      float cc[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc[q * 4 + j][0] = fmaf(cc[j], d0, acc[q * 4 + j][0]);
        acc[q * 4 + j][1] = fmaf(cc[j], d1, acc[q * 4 + j][1]);
        acc[q * 4 + j][2] = fmaf(cc[j], d2, acc[q * 4 + j][2]);
      }
    }
  }
  if (!SKIN) {
    verts[v * 3] = acc[0][0]; verts[v * 3 + 1] = acc[0][1]; verts[v * 3 + 2] = acc[0][2];
    return;
  }
  float w[NJ];
#pragma unroll
  for (int j = 0; j < NJ; ++j) w[j] = fm.lbs_w[v * NJ + j];
#pragma unroll
  for (int f = 0; f < FB; ++f) {
    if (f0 + f >= n_frames) break;
    float T[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) {
      float t = 0.f;
#pragma unroll
      for (int j = 0; j < NJ; ++j) t = fmaf(w[j], As[f * COEF_A + j * 12 + k], t);
      T[k] = t;
    }
    float* o = verts + ((int64_t)(f0 + f) * fm.V + v) * 3;
#pragma unroll
    for (int r = 0; r < 3; ++r)
      o[r] = (T[r * 4] * acc[f][0] + T[r * 4 + 1] * acc[f][1] + T[r * 4 + 2] * acc[f][2] + T[r * 4 + 3]) * fm.scale;
  }
}
}  // namespace

size_t flame_workspace_floats(const FlameModel& fm, int n_frames) {
  size_t ks = fm.ks_full > fm.ks_expr ? fm.ks_full : fm.ks_expr;                  // split-bf16 A operand: 3*KS bf16 per frame
  return (size_t)n_frames * (fm.n_shape + fm.n_exp + 36 + COEF_A) + (size_t)fm.V * 3 + 128 + ((size_t)n_frames * 3 * ks + 1) / 2;
}

int launch_flame(const FlameModel& m, const float* shape, int64_t shape_rs, const float* expr, int64_t expr_rs,
                 const float* pose, int64_t pose_rs, int zero_global, float* ws, float* verts, int n_frames,
                 cudaStream_t st) {
  if (n_frames <= 0) return AT_OK;
  FlameDev fm;
  fm.V = m.V; fm.n_shape = m.n_shape; fm.n_exp = m.n_exp; fm.n_bases = m.n_shape + m.n_exp + 36;
  fm.v_template = m.v_template; fm.dirs = m.dirs; fm.j_template = m.j_template; fm.j_dirs = m.j_dirs; fm.lbs_w = m.lbs_weights;
  for (int j = 0; j < NJ; ++j) fm.parents[j] = m.parents[j];
  fm.scale = m.scale;
  const int stride = fm.n_bases + COEF_A;
  float* coef = ws;
  float* v_static = ws + (size_t)n_frames * stride;
  flame_coef_kernel<<<ceil_div(n_frames, 4), 128, 0, st>>>(fm, shape, shape_rs, expr, expr_rs, pose, pose_rs, zero_global,
                                                           coef, n_frames);
  AT_LAUNCH_CHECK();
  flame_pose_kernel<<<ceil_div(n_frames, 128), 128, 0, st>>>(fm, pose, pose_rs, zero_global, coef, n_frames);
  AT_LAUNCH_CHECK();
  int l_begin = 0;
  const float* base = fm.v_template;
  dim3 grid_v(1, ceil_div(fm.V, 128));
  if (shape_rs == 0 && n_frames > 1 && fm.n_shape > 0) {
    // shared shape row: fold the shape bases into a static template once
    flame_static_kernel<<<ceil_div(fm.V * 3, 64), 256, 0, st>>>(fm.v_template, fm.dirs, coef, fm.n_shape, fm.V * 3, v_static);
    AT_LAUNCH_CHECK();
    base = v_static;
    l_begin = fm.n_shape;
  }
  // tensor-core path (flame_tc.cu) when the split-bf16 basis operands were provided
  const void* bsplit = l_begin ? m.bsplit_expr : m.bsplit_full;
  const int KS = l_begin ? m.ks_expr : m.ks_full;
  if (bsplit && KS > 0) {
    size_t a_off = ((size_t)n_frames * stride + (size_t)fm.V * 3 + 63) & ~(size_t)63;      // 256-byte aligned
    return launch_flame_tc(m, base, coef, stride, l_begin, fm.n_bases - l_begin, bsplit, KS, ws + a_off, verts, n_frames, st);
  }
  size_t smem = ((size_t)(fm.n_bases - l_begin) * FB + FB * COEF_A) * sizeof(float);
  AT_CUDA(cudaFuncSetAttribute(flame_verts_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  AT_REQUIRE(smem <= 64 * 1024, "flame: too many bases (%d)", fm.n_bases);
  dim3 grid(ceil_div(n_frames, FB), ceil_div(fm.V, 128));
  flame_verts_kernel<true><<<grid, 128, smem, st>>>(fm, base, l_begin, fm.n_bases, coef, stride, verts, n_frames);
  AT_LAUNCH_CHECK();
  return AT_OK;
}

}  // namespace artalk
