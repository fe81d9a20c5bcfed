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

// grid (n, ceil(778/256)): nearest object point of every vertex of candidate blockIdx.x (x: up to 2^31 - 1 candidates)
__global__ void __launch_bounds__(kCsThreads) k_vertex_contact(const float* __restrict__ verts, const float* __restrict__ obj,
                                                                int n_pts, int group, float* __restrict__ dist) {
  __shared__ float4 tile[kCsThreads];
  const int i = blockIdx.x, v = blockIdx.y * kCsThreads + threadIdx.x;
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

// Pseudo-force evaluation of one posed hand per CTA (BASELINE config 5): the forward math of one iteration of
// ForceOptimizer.optimize_batch (lib/engine/force_optimization.py:141-171) -- get_local_force (lib/model/physics.py:546-557:
// softmax over the 8 friction-cone anchors, normalised direction, |scale|), VERT2ANCHOR frames, from_local_to_global, then
// per hand: |sum f + g|, (sum f).(-g), |sum (p - CoM) x f| and the contact-distribution term mean_j (log|c_j / s_j| * mask_j)^2.
__global__ void __launch_bounds__(128) k_force_eval(AssetsDev as, const float* __restrict__ verts, const float* __restrict__ scale,
                                                    const float* __restrict__ weight, const unsigned char* __restrict__ mask,
                                                    const float* __restrict__ force_contact, const float* __restrict__ cone,
                                                    const float* __restrict__ gravity, const float* __restrict__ com, int group,
                                                    float* __restrict__ terms, float* __restrict__ force_local_out,
                                                    float* __restrict__ point_out, float* __restrict__ force_out) {
  __shared__ float j21[21 * 3];
  __shared__ float s_fp[32][3], s_fg[32][3], s_s[32], s_c[32];
  const int i = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* v = verts + (size_t)i * kVerts * 3;
  for (int o = warp; o < 63; o += 4) {
    const int k = o / 3, d = o % 3;
    float acc = 0.f;
    for (int vv = lane; vv < kVerts; vv += 32) acc = fmaf(as.v2j[k * kVerts + vv], v[vv * 3 + d], acc);
#pragma unroll
    for (int sft = 16; sft > 0; sft >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, sft);
    if (lane == 0) j21[o] = acc;
  }
  __syncthreads();
  if (tid < kAnchors) {
    const int j = tid;
    const float m = mask ? (mask[(size_t)i * kAnchors + j] ? 1.f : 0.f) : 1.f;
    const float sc = scale[(size_t)i * kAnchors + j] * m;                      // scale * contact_mask  (:141)
    const float* w = weight + ((size_t)i * kAnchors + j) * 8;
    float mx = w[0];
#pragma unroll
    for (int k = 1; k < 8; ++k) mx = fmaxf(mx, w[k]);
    float e[8], es = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) { e[k] = expf(w[k] - mx); es += e[k]; }
    float dir[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float wk = e[k] / es;
#pragma unroll
      for (int d = 0; d < 3; ++d) dir[d] += wk * cone[k * 3 + d];
    }
    const float dn = sqrtf((dir[0] * dir[0] + dir[1] * dir[1]) + dir[2] * dir[2]) + 1e-8f;
    float fl[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) fl[d] = (dir[d] / dn) * fabsf(sc);
    float pt[3], fg[3];
    anchor_point_and_force(
        as, j, [&](int vid, float* out) { out[0] = v[vid * 3 + 0]; out[1] = v[vid * 3 + 1]; out[2] = v[vid * 3 + 2]; }, j21, fl, pt,
        fg);
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      s_fp[j][d] = pt[d]; s_fg[j][d] = fg[d];
      if (force_local_out) force_local_out[((size_t)i * kAnchors + j) * 3 + d] = fl[d];
      if (point_out) point_out[((size_t)i * kAnchors + j) * 3 + d] = pt[d];
      if (force_out) force_out[((size_t)i * kAnchors + j) * 3 + d] = fg[d];
    }
    s_s[j] = sc;
    s_c[j] = force_contact ? force_contact[(size_t)i * kAnchors + j] : 0.f;
    __syncwarp();
    if (j == 0) {
      const float* g = gravity + (size_t)(i / group) * 3;
      const float* cm = com + (size_t)(i / group) * 3;
      float F[3] = {0.f, 0.f, 0.f}, M[3] = {0.f, 0.f, 0.f}, ss = 0.f, cs = 0.f;
      for (int k = 0; k < kAnchors; ++k) {
        const float arm[3] = {s_fp[k][0] - cm[0], s_fp[k][1] - cm[1], s_fp[k][2] - cm[2]};
        F[0] += s_fg[k][0]; F[1] += s_fg[k][1]; F[2] += s_fg[k][2];
        M[0] += arm[1] * s_fg[k][2] - arm[2] * s_fg[k][1];
        M[1] += arm[2] * s_fg[k][0] - arm[0] * s_fg[k][2];
        M[2] += arm[0] * s_fg[k][1] - arm[1] * s_fg[k][0];
        ss += s_s[k] * s_s[k];
        cs += s_c[k] * s_c[k];
      }
      const float r[3] = {F[0] + g[0], F[1] + g[1], F[2] + g[2]};
      float* o = terms + (size_t)i * 4;
      o[0] = sqrtf((r[0] * r[0] + r[1] * r[1]) + r[2] * r[2]);
      o[1] = (F[0] * (-1.f * g[0]) + F[1] * (-1.f * g[1])) + F[2] * (-1.f * g[2]);
      o[2] = sqrtf((M[0] * M[0] + M[1] * M[1]) + M[2] * M[2]);
      float dsum = 0.f;
      if (force_contact) {
        const float sn = sqrtf(ss) + 1e-8f, cn = sqrtf(cs) + 1e-8f;
        for (int k = 0; k < kAnchors; ++k) {
          const float mk = mask ? (mask[(size_t)i * kAnchors + k] ? 1.f : 0.f) : 1.f;
          const float dd = logf(fabsf((s_c[k] / cn) / (s_s[k] / sn + 1e-8f)) + 1e-8f) * mk;
          dsum += dd * dd;
        }
      }
      o[3] = dsum / (float)kAnchors;
    }
  }
}

// Final pose-error metrics of one image per CTA: MJE / MVE of TesterHand (lib/engine/test.py:657-679,
// mean Euclidean distance, metres -> mm) and ADD / ADD-S of TesterObject.criterion_ADD_REP (test.py:413-442) on the
// object's sampled vertices.  Deterministic block reductions; ADD-S scans shared-memory tiles of ground-truth points.
__device__ __forceinline__ float block_sum_256(float v, float* red) {
  const int tid = threadIdx.x;
  red[tid] = v;
  __syncthreads();
  for (int st = 128; st > 0; st >>= 1) {
    if (tid < st) red[tid] += red[tid + st];
    __syncthreads();
  }
  const float r = red[0];
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(256) k_pose_metrics(AssetsDev as, const float* __restrict__ pd_joint, const float* __restrict__ gt_joint,
                                                      const float* __restrict__ pd_vert, const float* __restrict__ gt_vert,
                                                      const double* __restrict__ pd_obj, const double* __restrict__ gt_obj,
                                                      const int* __restrict__ obj_id, float* __restrict__ out) {
  __shared__ float red[256];
  __shared__ float4 tile[256];
  const int b = blockIdx.x, tid = threadIdx.x;
  float acc = 0.f;
  if (tid < 21) {
    const float* p = pd_joint + ((size_t)b * 21 + tid) * 3;
    const float* g = gt_joint + ((size_t)b * 21 + tid) * 3;
    const float dx = p[0] - g[0], dy = p[1] - g[1], dz = p[2] - g[2];
    acc = sqrtf((dx * dx + dy * dy) + dz * dz);
  }
  const float mje = block_sum_256(acc, red) / 21.f * 1000.f;
  acc = 0.f;
  for (int v = tid; v < kVerts; v += 256) {
    const float* p = pd_vert + ((size_t)b * kVerts + v) * 3;
    const float* g = gt_vert + ((size_t)b * kVerts + v) * 3;
    const float dx = p[0] - g[0], dy = p[1] - g[1], dz = p[2] - g[2];
    acc += sqrtf((dx * dx + dy * dy) + dz * dz);
  }
  const float mve = block_sum_256(acc, red) / (float)kVerts * 1000.f;
  // object: pose both clouds on the fly (HeadObject semantics, no flip, no root offset)
  ObjPose pp, pg;
  const float zero[3] = {0.f, 0.f, 0.f};
  make_obj_pose(pd_obj + (size_t)b * 9, nullptr, zero, true, pp);
  make_obj_pose(gt_obj + (size_t)b * 9, nullptr, zero, true, pg);
  const float* base = as.verts + (size_t)obj_index(as, obj_id[b]) * as.n_pts * 3;
  float add = 0.f, adds = 0.f;
  for (int p0 = 0; p0 < as.n_pts; p0 += 256) {          // this thread's predicted point of the current slab
    const int i = p0 + tid;
    float a[3] = {0.f, 0.f, 0.f};
    if (i < as.n_pts) {
      float g[3];
      obj_point(pp, base + (size_t)i * 3, a);
      obj_point(pg, base + (size_t)i * 3, g);
      const float dx = a[0] - g[0], dy = a[1] - g[1], dz = a[2] - g[2];
      add += sqrtf((dx * dx + dy * dy) + dz * dz);
    }
    float best = INFINITY;
    for (int q0 = 0; q0 < as.n_pts; q0 += 256) {
      __syncthreads();
      if (q0 + tid < as.n_pts) {
        float g[3];
        obj_point(pg, base + (size_t)(q0 + tid) * 3, g);
        tile[tid] = make_float4(g[0], g[1], g[2], 0.f);
      }
      __syncthreads();
      const int cnt = min(256, as.n_pts - q0);
#pragma unroll 8
      for (int q = 0; q < cnt; ++q) {
        const float4 t = tile[q];
        const float dx = a[0] - t.x, dy = a[1] - t.y, dz = a[2] - t.z;
        best = fminf(best, (dx * dx + dy * dy) + dz * dz);
      }
    }
    if (i < as.n_pts) adds += sqrtf(best);
  }
  const float add_mm = block_sum_256(add, red) / (float)as.n_pts * 1000.f;
  const float adds_mm = block_sum_256(adds, red) / (float)as.n_pts * 1000.f;
  if (tid == 0) {
    out[(size_t)b * 4 + 0] = mje; out[(size_t)b * 4 + 1] = mve; out[(size_t)b * 4 + 2] = add_mm; out[(size_t)b * 4 + 3] = adds_mm;
  }
}


// ------------------------------------------------------------------------------------------------------------
// Procrustes-aligned hand errors: PA-MJE / PA-MVE and the per-joint errors of TesterHand.criterion_MJE_PAMJE
// (lib/engine/test.py:657-679) with rigid_align_AtoB (lib/utils/transform_fn.py:43-66): similarity transform
// (c, R, t) = argmin |c R A + t - B| from the SVD of the 3x3 cross-covariance.  One CTA per image; float64 moments with
// fixed-order block reductions; the SVD is taken from the Jacobi eigen-decomposition of H^T H by one thread.
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double block_sum_256d(double v, double* red) {
  const int tid = threadIdx.x;
  red[tid] = v;
  __syncthreads();
  for (int st = 128; st > 0; st >>= 1) {
    if (tid < st) red[tid] += red[tid + st];
    __syncthreads();
  }
  const double r = red[0];
  __syncthreads();
  return r;
}

// eigen-decomposition of a symmetric 3x3 (row-major M): eigenvalues w descending, eigenvectors in the COLUMNS of V
__device__ void sym3_eig(const double* M, double* w, double* V) {
  double a[9];
  for (int i = 0; i < 9; ++i) { a[i] = M[i]; V[i] = (i % 4 == 0) ? 1.0 : 0.0; }
  for (int sweep = 0; sweep < 30; ++sweep) {
    const double off = a[1] * a[1] + a[2] * a[2] + a[5] * a[5];
    if (off <= 1e-300) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        const double apq = a[p * 3 + q];
        if (fabs(apq) <= 1e-300) continue;
        const double theta = (a[q * 3 + q] - a[p * 3 + p]) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
        for (int k = 0; k < 3; ++k) {          // A <- A J
          const double akp = a[k * 3 + p], akq = a[k * 3 + q];
          a[k * 3 + p] = c * akp - sn * akq;
          a[k * 3 + q] = sn * akp + c * akq;
        }
        for (int k = 0; k < 3; ++k) {          // A <- J^T A
          const double apk = a[p * 3 + k], aqk = a[q * 3 + k];
          a[p * 3 + k] = c * apk - sn * aqk;
          a[q * 3 + k] = sn * apk + c * aqk;
        }
        for (int k = 0; k < 3; ++k) {          // V <- V J
          const double vkp = V[k * 3 + p], vkq = V[k * 3 + q];
          V[k * 3 + p] = c * vkp - sn * vkq;
          V[k * 3 + q] = sn * vkp + c * vkq;
        }
      }
  }
  w[0] = a[0]; w[1] = a[4]; w[2] = a[8];
  for (int i = 0; i < 2; ++i)                   // sort descending (columns follow)
    for (int j = 0; j < 2 - i; ++j)
      if (w[j] < w[j + 1]) {
        const double tw = w[j]; w[j] = w[j + 1]; w[j + 1] = tw;
        for (int k = 0; k < 3; ++k) { const double tv = V[k * 3 + j]; V[k * 3 + j] = V[k * 3 + j + 1]; V[k * 3 + j + 1] = tv; }
      }
}

// similarity transform of the point set A [n][3] onto B: fills T[12] = {c R (row-major 3x3), t}
__device__ void procrustes_256(const float* __restrict__ A, const float* __restrict__ B, int n, double* red, double* T /*shared [12]*/) {
  const int tid = threadIdx.x;
  double ca[3], cb[3];
  for (int d = 0; d < 3; ++d) {
    double sa = 0.0, sb = 0.0;
    for (int i = tid; i < n; i += 256) { sa += (double)A[i * 3 + d]; sb += (double)B[i * 3 + d]; }
    ca[d] = block_sum_256d(sa, red) / n;
    cb[d] = block_sum_256d(sb, red) / n;
  }
  double H[9], var = 0.0;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) {
      double acc = 0.0;
      for (int i = tid; i < n; i += 256) acc += ((double)A[i * 3 + r] - ca[r]) * ((double)B[i * 3 + c] - cb[c]);
      H[r * 3 + c] = block_sum_256d(acc, red) / n;          // H = (A - cA)^T (B - cB) / n
    }
  for (int d = 0; d < 3; ++d) {
    double acc = 0.0;
    for (int i = tid; i < n; i += 256) { const double x = (double)A[i * 3 + d] - ca[d]; acc += x * x; }
    var += block_sum_256d(acc, red) / n;                    // np.var(A, axis=0).sum()
  }
  if (tid == 0) {
    // H = U S V^T.  Eigen-decompose H^T H = V S^2 V^T, make V right-handed, u_i = H v_i / s_i, u_3 = u_1 x u_2; the
    // reflection case of the reference (det(V U^T) < 0 -> negate the last singular value) is the sign of det(H).
    double M[9], w[3], V[9];
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) M[r * 3 + c] = (H[0 * 3 + r] * H[0 * 3 + c] + H[1 * 3 + r] * H[1 * 3 + c]) + H[2 * 3 + r] * H[2 * 3 + c];
    sym3_eig(M, w, V);
    double v[3][3], u[3][3], sv[3];
    for (int i = 0; i < 3; ++i)
      for (int k = 0; k < 3; ++k) v[i][k] = V[k * 3 + i];
    v[2][0] = v[0][1] * v[1][2] - v[0][2] * v[1][1];         // v3 = v1 x v2
    v[2][1] = v[0][2] * v[1][0] - v[0][0] * v[1][2];
    v[2][2] = v[0][0] * v[1][1] - v[0][1] * v[1][0];
    for (int i = 0; i < 3; ++i) sv[i] = sqrt(w[i] > 0.0 ? w[i] : 0.0);
    for (int i = 0; i < 2; ++i) {
      double hv[3], nrm = 0.0;
      for (int r = 0; r < 3; ++r) { hv[r] = (H[r * 3 + 0] * v[i][0] + H[r * 3 + 1] * v[i][1]) + H[r * 3 + 2] * v[i][2]; nrm += hv[r] * hv[r]; }
      nrm = sqrt(nrm);
      for (int r = 0; r < 3; ++r) u[i][r] = nrm > 0.0 ? hv[r] / nrm : (r == i ? 1.0 : 0.0);
    }
    u[2][0] = u[0][1] * u[1][2] - u[0][2] * u[1][1];         // u3 = u1 x u2
    u[2][1] = u[0][2] * u[1][0] - u[0][0] * u[1][2];
    u[2][2] = u[0][0] * u[1][1] - u[0][1] * u[1][0];
    const double det = H[0] * (H[4] * H[8] - H[5] * H[7]) - H[1] * (H[3] * H[8] - H[5] * H[6]) + H[2] * (H[3] * H[7] - H[4] * H[6]);
    const double dsg = det < 0.0 ? -1.0 : 1.0;
    const double scale = ((sv[0] + sv[1]) + dsg * sv[2]) / var;          // c = sum(s) / var(A)
    // R = V U^T with both bases right-handed: for det(H) < 0 this IS the reference's "flip the last singular pair"
    // rotation (its own u_3 is then -(u_1 x u_2)); only the scale sees the sign
    double R[9];
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) R[r * 3 + c] = (v[0][r] * u[0][c] + v[1][r] * u[1][c]) + v[2][r] * u[2][c];
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 3; ++c) T[r * 3 + c] = scale * R[r * 3 + c];
      T[9 + r] = cb[r] - ((T[r * 3 + 0] * ca[0] + T[r * 3 + 1] * ca[1]) + T[r * 3 + 2] * ca[2]);      // t = cB - c R cA
    }
  }
  __syncthreads();
}

__device__ __forceinline__ double aligned_error(const float* a, const float* b, const double* T) {
  double e2 = 0.0;
  for (int r = 0; r < 3; ++r) {
    const double x = ((T[r * 3 + 0] * (double)a[0] + T[r * 3 + 1] * (double)a[1]) + T[r * 3 + 2] * (double)a[2]) + T[9 + r];
    const double d = (double)b[r] - x;
    e2 += d * d;
  }
  return sqrt(e2);
}

// full = 0: out [n][23] = {PA-MJE, PA-MVE, JE[21]} in mm;  full = 1: out [n][25] = {MJE, PA-MJE, MVE, PA-MVE, JE[21]}, the whole
// per-image row of TesterHand.__call__ (lib/engine/test.py:589-654)
__global__ void __launch_bounds__(256) k_hand_pa_metrics(const float* __restrict__ pd_joint, const float* __restrict__ gt_joint,
                                                         const float* __restrict__ pd_vert, const float* __restrict__ gt_vert,
                                                         float* __restrict__ out, int full, int gt_rows) {
  __shared__ double red[256];
  __shared__ double T[12];
  const int b = blockIdx.x, tid = threadIdx.x, g = b % gt_rows;      // several prediction sets may share one ground truth
  const int W = full ? 25 : 23, o_je = full ? 4 : 2, o_pamje = full ? 1 : 0, o_pamve = full ? 3 : 1;
  const float* pj = pd_joint + (size_t)b * 21 * 3;
  const float* gj = gt_joint + (size_t)g * 21 * 3;
  const float* pv = pd_vert + (size_t)b * kVerts * 3;
  const float* gv = gt_vert + (size_t)g * kVerts * 3;
  double je = 0.0;
  if (tid < 21) {
    const float dx = gj[tid * 3 + 0] - pj[tid * 3 + 0], dy = gj[tid * 3 + 1] - pj[tid * 3 + 1], dz = gj[tid * 3 + 2] - pj[tid * 3 + 2];
    const float d = sqrtf((dx * dx + dy * dy) + dz * dz);
    out[(size_t)b * W + o_je + tid] = d * 1000.f;
    je = (double)d;
  }
  if (full) {
    const double mje = block_sum_256d(je, red) / 21.0;
    double ve = 0.0;
    for (int i = tid; i < kVerts; i += 256) {
      const float dx = gv[i * 3 + 0] - pv[i * 3 + 0], dy = gv[i * 3 + 1] - pv[i * 3 + 1], dz = gv[i * 3 + 2] - pv[i * 3 + 2];
      ve += (double)sqrtf((dx * dx + dy * dy) + dz * dz);
    }
    const double mve = block_sum_256d(ve, red) / (double)kVerts;
    if (tid == 0) {
      out[(size_t)b * W + 0] = (float)(mje * 1000.0);
      out[(size_t)b * W + 2] = (float)(mve * 1000.0);
    }
  }
  procrustes_256(pj, gj, 21, red, T);
  double e = tid < 21 ? aligned_error(pj + tid * 3, gj + tid * 3, T) : 0.0;
  const double pa_mje = block_sum_256d(e, red) / 21.0;
  procrustes_256(pv, gv, kVerts, red, T);
  e = 0.0;
  for (int i = tid; i < kVerts; i += 256) e += aligned_error(pv + i * 3, gv + i * 3, T);
  const double pa_mve = block_sum_256d(e, red) / (double)kVerts;
  if (tid == 0) {
    out[(size_t)b * W + o_pamje] = (float)(pa_mje * 1000.0);
    out[(size_t)b * W + o_pamve] = (float)(pa_mve * 1000.0);
  }
}

// used by vpho_eval_record (metrics.cu): `rows` predictions against ground truth rows [rows % gt_rows]
int launch_hand_metrics_full(const float* pd_joint, const float* gt_joint, const float* pd_vert, const float* gt_vert, int rows, int gt_rows,
                             float* metrics, cudaStream_t st) {
  VPHO_LAUNCH(k_hand_pa_metrics, dim3(rows), dim3(256), 0, st, pd_joint, gt_joint, pd_vert, gt_vert, metrics, 1, gt_rows);
  return VPHO_OK;
}

}  // namespace vpho

using namespace vpho;

extern "C" int vpho_pose_metrics(vpho_assets_t h, const float* pd_joint, const float* gt_joint, const float* pd_vert,
                                 const float* gt_vert, const double* pd_obj6d, const double* gt_obj6d, const int32_t* obj_id, int n,
                                 float* metrics, void* stream) {
  if (!h || n < 0) return VPHO_ERR_INVALID;
  if (n == 0) return VPHO_OK;
  if (!pd_joint || !gt_joint || !pd_vert || !gt_vert || !pd_obj6d || !gt_obj6d || !obj_id || !metrics) return VPHO_ERR_INVALID;
  VPHO_LAUNCH(k_pose_metrics, dim3(n), dim3(256), 0, (cudaStream_t)stream, static_cast<AssetsHost*>(h)->dev, pd_joint, gt_joint, pd_vert,
              gt_vert, pd_obj6d, gt_obj6d, obj_id, metrics);
  VPHO_CHECK_LAUNCH();
  return VPHO_OK;
}

extern "C" int vpho_hand_pa_metrics(const float* pd_joint, const float* gt_joint, const float* pd_vert, const float* gt_vert, int n,
                                    float* metrics, void* stream) {
  if (n < 0) return VPHO_ERR_INVALID;
  if (n == 0) return VPHO_OK;
  if (!pd_joint || !gt_joint || !pd_vert || !gt_vert || !metrics) return VPHO_ERR_INVALID;
  VPHO_LAUNCH(k_hand_pa_metrics, dim3(n), dim3(256), 0, (cudaStream_t)stream, pd_joint, gt_joint, pd_vert, gt_vert, metrics, 0, n);
  VPHO_CHECK_LAUNCH();
  return VPHO_OK;
}

extern "C" int vpho_hand_metrics(const float* pd_joint, const float* gt_joint, const float* pd_vert, const float* gt_vert, int n,
                                 float* metrics, void* stream) {
  if (n < 0) return VPHO_ERR_INVALID;
  if (n == 0) return VPHO_OK;
  if (!pd_joint || !gt_joint || !pd_vert || !gt_vert || !metrics) return VPHO_ERR_INVALID;
  VPHO_LAUNCH(k_hand_pa_metrics, dim3(n), dim3(256), 0, (cudaStream_t)stream, pd_joint, gt_joint, pd_vert, gt_vert, metrics, 1, n);
  VPHO_CHECK_LAUNCH();
  return VPHO_OK;
}

extern "C" int vpho_force_eval(vpho_assets_t h, const float* verts, const float* scale, const float* weight, const uint8_t* contact_mask,
                               const float* force_contact, const float* cone_anchor, const float* gravity, const float* com, int n,
                               int group, float* terms, float* force_local, float* force_point, float* force_global, void* stream) {
  if (!h || n < 0 || group <= 0) return VPHO_ERR_INVALID;
  if (n == 0) return VPHO_OK;
  if (!verts || !scale || !weight || !cone_anchor || !gravity || !com || !terms) return VPHO_ERR_INVALID;
  VPHO_LAUNCH(k_force_eval, dim3(n), dim3(128), 0, (cudaStream_t)stream, static_cast<AssetsHost*>(h)->dev, verts, scale, weight,
              contact_mask, force_contact, cone_anchor, gravity, com, group, terms, force_local, force_point, force_global);
  VPHO_CHECK_LAUNCH();
  return VPHO_OK;
}

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
  VPHO_LAUNCH(k_vertex_contact, dim3(n, (kVerts + kCsThreads - 1) / kCsThreads), dim3(kCsThreads), 0, (cudaStream_t)stream, verts,
              obj_points, n_pts, group, dist);
  VPHO_CHECK_LAUNCH();
  return VPHO_OK;
}
