// Stand-alone contact scoring entry points (BASELINE config 3: "MANO LBS + hand-object penetration/contact scoring").
//   vpho_anchor_contact : the scoring part of HandAggregator.select_by_physics (lib/model/aggregation.py:553-590) for n posed
//                         hands against one object point cloud per group: nearest distance of the 32 force anchors
//                         (exact-mode cdist + min, :1145-1158) and the 5 per-finger physics scores.
//   vpho_vertex_contact : dense variant, every MANO vertex against the cloud (nearest distance per vertex) -- a superset
//                         stress case, NOT something the reference computes (it scores 32 anchors, aggregation.py:972).
// Both scan shared-memory tiles of object points; the anchor kernel maps one anchor per lane with a cross-warp
// (distance, index) min-reduction, the vertex kernel one vertex per thread.
#include "agg_device.cuh"
#include "vpho_b200.h"

namespace vpho {

constexpr int kCsThreads = 256;

__global__ void __launch_bounds__(kCsThreads) k_anchor_contact(const float* __restrict__ fpoint, const float* __restrict__ fglobal,
                                                                const float* __restrict__ obj, int n_pts, int group,
                                                                float* __restrict__ dist, float* __restrict__ score) {
  __shared__ float4 tile[kCsThreads];
  __shared__ float red_d2[8 * 32];
  __shared__ float s_fn[32], s_dir[32][3], s_score[32];
  const int i = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* ov = obj + (size_t)(i / group) * n_pts * 3;
  const float ax = fpoint[((size_t)i * kAnchors + lane) * 3 + 0], ay = fpoint[((size_t)i * kAnchors + lane) * 3 + 1],
              az = fpoint[((size_t)i * kAnchors + lane) * 3 + 2];
  float best = INFINITY;
  for (int p0 = 0; p0 < n_pts; p0 += kCsThreads) {
    __syncthreads();
    if (p0 + tid < n_pts) tile[tid] = make_float4(ov[(size_t)(p0 + tid) * 3], ov[(size_t)(p0 + tid) * 3 + 1], ov[(size_t)(p0 + tid) * 3 + 2], 0.f);
    __syncthreads();
    const int base = warp * 32, cnt = min(32, n_pts - p0 - base);
    for (int q = 0; q < cnt; ++q) {
      const float4 p = tile[base + q];
      const float dx = ax - p.x, dy = ay - p.y, dz = az - p.z;
      best = fminf(best, (dx * dx + dy * dy) + dz * dz);
    }
  }
  red_d2[warp * 32 + lane] = best;
  __syncthreads();
  if (warp == 0) {
    float bd = red_d2[lane];
    for (int w = 1; w < kCsThreads / 32; ++w) bd = fminf(bd, red_d2[w * 32 + lane]);
    const int j = lane;
    float fg[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) fg[d] = fglobal[((size_t)i * kAnchors + j) * 3 + d];
    const float fn = sqrtf((fg[0] * fg[0] + fg[1] * fg[1]) + fg[2] * fg[2]);
    s_fn[j] = fn;
#pragma unroll
    for (int d = 0; d < 3; ++d) s_dir[j][d] = fg[d] / fn;
    __syncwarp();
    float fsum = 0.f, I[3] = {0.f, 0.f, 0.f};
    for (int k = 0; k < kAnchors; ++k) { fsum += s_fn[k]; I[0] += s_dir[k][0]; I[1] += s_dir[k][1]; I[2] += s_dir[k][2]; }
    const float In = sqrtf((I[0] * I[0] + I[1] * I[1]) + I[2] * I[2]);
    const float d = sqrtf(bd);
    if (dist) dist[(size_t)i * kAnchors + j] = d;
    s_score[j] = -(((fn / fsum) * d) * In);
    __syncwarp();
    if (j < 5 && score) {
      float acc = 0.f;
      for (int q = 0; q < 4; ++q) acc += s_score[finger_anchor(j, q)];
      score[(size_t)i * 5 + j] = acc;
    }
  }
}

// grid (ceil(778/256), n): nearest object point of every vertex of candidate blockIdx.y
__global__ void __launch_bounds__(kCsThreads) k_vertex_contact(const float* __restrict__ verts, const float* __restrict__ obj,
                                                                int n_pts, int group, float* __restrict__ dist) {
  __shared__ float4 tile[kCsThreads];
  const int i = blockIdx.y, v = blockIdx.x * kCsThreads + threadIdx.x;
  const float* ov = obj + (size_t)(i / group) * n_pts * 3;
  float x = 0.f, y = 0.f, z = 0.f;
  if (v < kVerts) { x = verts[((size_t)i * kVerts + v) * 3]; y = verts[((size_t)i * kVerts + v) * 3 + 1]; z = verts[((size_t)i * kVerts + v) * 3 + 2]; }
  float best = INFINITY;
  for (int p0 = 0; p0 < n_pts; p0 += kCsThreads) {
    __syncthreads();
    const int p = p0 + threadIdx.x;
    if (p < n_pts) tile[threadIdx.x] = make_float4(ov[(size_t)p * 3], ov[(size_t)p * 3 + 1], ov[(size_t)p * 3 + 2], 0.f);
    __syncthreads();
    const int cnt = min(kCsThreads, n_pts - p0);
#pragma unroll 8
    for (int q = 0; q < cnt; ++q) {
      const float4 pt = tile[q];
      const float dx = x - pt.x, dy = y - pt.y, dz = z - pt.z;
      best = fminf(best, (dx * dx + dy * dy) + dz * dz);
    }
  }
  if (v < kVerts) dist[(size_t)i * kVerts + v] = sqrtf(best);
}

}  // namespace vpho

using namespace vpho;

extern "C" int vpho_anchor_contact(const float* force_point, const float* force_global, const float* obj_points, int n, int group,
                                   int n_pts, float* dist, float* finger_score, void* stream) {
  if (n < 0 || group <= 0 || n_pts <= 0) return VPHO_ERR_INVALID;
  if (n == 0) return VPHO_OK;
  if (!force_point || !force_global || !obj_points || (!dist && !finger_score)) return VPHO_ERR_INVALID;
  VPHO_LAUNCH(k_anchor_contact, dim3(n), dim3(kCsThreads), 0, (cudaStream_t)stream, force_point, force_global, obj_points, n_pts, group,
              dist, finger_score);
  VPHO_CHECK_LAUNCH();
  return VPHO_OK;
}

extern "C" int vpho_vertex_contact(const float* verts, const float* obj_points, int n, int group, int n_pts, float* dist, void* stream) {
  if (n < 0 || group <= 0 || n_pts <= 0) return VPHO_ERR_INVALID;
  if (n == 0) return VPHO_OK;
  if (!verts || !obj_points || !dist) return VPHO_ERR_INVALID;
  VPHO_LAUNCH(k_vertex_contact, dim3((kVerts + kCsThreads - 1) / kCsThreads, n), dim3(kCsThreads), 0, (cudaStream_t)stream, verts,
              obj_points, n_pts, group, dist);
  VPHO_CHECK_LAUNCH();
  return VPHO_OK;
}
