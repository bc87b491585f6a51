// Per-vertex normals of the decoded FLAME meshes for a mesh rasteriser (the reference hands
// (verts, faces) to pytorch3d's Meshes, app/flame_model/renderer_utils.py, whose verts_normals are the area-weighted sum of
// the incident faces' normals, normalised with eps 1e-6):
//     n_v = normalize( sum over faces (i0,i1,i2) containing v of cross(P_next - P_v, P_prev - P_v) )      (cyclic order)
//
// Memory-bound row work (60 KB in + 60 KB out per frame, ~35 flops per byte-free incidence): one CTA per frame stages the
// frame's 5023 x 3 floats in shared memory with coalesced loads, then a thread per vertex GATHERS its incident faces from a
// CSR adjacency built once on the host (no atomics, deterministic summation order) and writes its normal. The adjacency
// (vertex -> ordered (next, prev) pairs, 240 KB) is shared by every frame and stays L2-resident.
#include "kernels.cuh"

namespace artalk {

namespace {

__global__ void __launch_bounds__(256) vertex_normals_kernel(const float* __restrict__ verts, int64_t frame_stride, int V,
                                                             const int* __restrict__ adj_off, const int2* __restrict__ adj_pair,
                                                             float* __restrict__ normals, int n_frames) {
  extern __shared__ float sv[];                    // V * 3 floats
  pdl_enter();
  for (int f = blockIdx.x; f < n_frames; f += gridDim.x) {
    const float* src = verts + (int64_t)f * frame_stride;
    __syncthreads();                               // previous frame's readers are done
    for (int i = threadIdx.x; i < V * 3; i += blockDim.x) sv[i] = src[i];
    __syncthreads();
    float* dst = normals + (int64_t)f * V * 3;
    for (int v = threadIdx.x; v < V; v += blockDim.x) {
      const float px = sv[v * 3], py = sv[v * 3 + 1], pz = sv[v * 3 + 2];
      float nx = 0.f, ny = 0.f, nz = 0.f;
      const int e0 = __ldg(adj_off + v), e1 = __ldg(adj_off + v + 1);
      // four incidences per step: their adjacency loads (L2) and the twelve shared-memory gathers are independent, so the
      // chain per vertex is ~2 round trips instead of one per incident face (FLAME vertices have ~6)
      for (int e = e0; e < e1; e += 4) {
        int2 pr[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) pr[i] = (e + i < e1) ? __ldg(adj_pair + e + i) : make_int2(v, v);      // (v, v): zero cross product
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float ax = sv[pr[i].x * 3] - px, ay = sv[pr[i].x * 3 + 1] - py, az = sv[pr[i].x * 3 + 2] - pz;
          const float bx = sv[pr[i].y * 3] - px, by = sv[pr[i].y * 3 + 1] - py, bz = sv[pr[i].y * 3 + 2] - pz;
          nx += ay * bz - az * by;
          ny += az * bx - ax * bz;
          nz += ax * by - ay * bx;
        }
      }
      const float inv = 1.0f / fmaxf(sqrtf(nx * nx + ny * ny + nz * nz), 1e-6f);      // F.normalize(eps = 1e-6)
      dst[v * 3] = nx * inv; dst[v * 3 + 1] = ny * inv; dst[v * 3 + 2] = nz * inv;
    }
  }
}

}  // namespace

// verts [n_frames][V][3] (frame stride in floats), adj_off [V + 1], adj_pair [adj_off[V]] = (next, prev) vertex of every
// incidence in the face's cyclic order; normals [n_frames][V][3]
int launch_vertex_normals(const float* verts, int64_t frame_stride, int V, const int* adj_off, const int* adj_pair, float* normals,
                          int n_frames, cudaStream_t st) {
  if (n_frames <= 0) return AT_OK;
  AT_REQUIRE(verts && adj_off && adj_pair && normals && V > 0, "vertex_normals: bad argument");
  AT_REQUIRE(((uintptr_t)adj_pair % 8) == 0, "vertex_normals: adjacency pairs must be 8-byte aligned");
  const int smem = V * 3 * (int)sizeof(float);
  AT_REQUIRE(smem <= 200 * 1024, "vertex_normals: %d vertices do not fit in shared memory", V);
  AT_TRY(ensure_dyn_smem((const void*)vertex_normals_kernel, smem));
  const DevCtx* dc = nullptr;
  AT_TRY(dev_ctx(&dc));
  const int per_sm = (227 * 1024) / (smem + 1024) > 0 ? (227 * 1024) / (smem + 1024) : 1;
  int grid = dc->num_sms * per_sm;
  if (grid > n_frames) grid = n_frames;
  AT_CUDA(launch_k(vertex_normals_kernel, dim3(grid), dim3(256), (size_t)smem, st, verts, frame_stride, V, adj_off,
                   reinterpret_cast<const int2*>(adj_pair), normals, n_frames));
  AT_LAUNCH_CHECK();
  return AT_OK;
}

}  // namespace artalk
