// Visual + physical scoring, top-k selection and weighted aggregation for sm_100a.
// Replaces HOI_Aggregator.__call__ (lib/model/aggregation.py:1167-1353) and everything it reaches:
//   HandAggregator.select_by_heatmap_cascade (:115-178) / select_topk_hand_by_observed_heatmap_and_fuse_by_index
//   (:180-284), ObjectAggregator.select_topk_object_by_heatmap (:742-780), fuse_topk (:729-740), average_rot6d
//   (:50-56), select_topk_object_by_physics3 (:947-997) with cdist_memory_save / nn_for_r_memory_save (:1115-1142),
//   HandAggregator.select_by_physics (:537-626), average_quaternion (lib/utils/transform_fn.py:101-125),
//   VERT2ANCHOR (lib/utils/physics_fn.py:224-257), from_local_to_global (lib/model/physics.py:362-371),
//   HeadObject.forward / flip_pt3d (lib/model/head_object.py:36-67).
//
// Design: no posed point cloud of a candidate is ever written to HBM.  Candidate joints come from a joints-only
// MANO evaluation (16 kinematic joints + 5 fingertip vertices), object key-points / surface points are posed on the
// fly, anchor->surface distances are scanned with one anchor per lane over shared-memory tiles of posed points, and
// every top-k is a warp-level bitonic sort on (value desc, index asc) keys fused with the weighting / quaternion
// averaging that consumes it.
#include "agg_device.cuh"
#include "vpho_b200.h"

#include <vector>

namespace vpho {

int mano_forward_dev(const ManoModelDev& m, const float* pose, const float* shape, int pose_stride, int shape_stride,
                     int n, float* verts, float* joints, cudaStream_t stream, bool simt = false, int debug_blend = 0);
const ManoModelDev& mano_model_dev(const void* handle);

// everything the aggregation kernels need, passed by value
struct HoiDev {
  vpho_hoi_args a;
  int nc;            // hand physics candidates per image = topk_hand + 1
  int kk;            // recombined object candidates per image = topk_obj^2
  int n_pts;
  // workspace
  float* hscore;     // [bs][2S][5]
  float* fused;      // [bs][48]   cascade-fused pose (levels written as they are fused)
  float* l4;         // [bs][topk_hand][5][3]  level-3 axis-angles of the per-finger top-k (aggregation.py:1310)
  float* cverts;     // [bs][778][3] cascade-fused hand (wrist-centred)
  float* cjoints;    // [bs][21][3]
  float* fpoint;     // [bs][32][3]
  float* fglobal;    // [bs][32][3]
  float* oscore;     // [bs][max(S,kk)]
  float* pscore;     // [bs][kk]
  int* t_topk;       // [bs][topk_obj]
  double* t_fused;   // [bs][3]
  float* ppose;      // [bs][nc][48]
  float* pverts;     // [bs*nc][778][3]
  float* pjoints;    // [bs*nc][21][3]
  float* ppoint;     // [bs*nc][32][3]
  float* pforce;     // [bs*nc][32][3]
  float* fscore;     // [bs][5][nc]
  float* ocom;       // [bs][3]
  float* pshape;     // [bs*nc][10]
};

// ------------------------------------------------------------------------------------------------------------
// hand cascade: candidate parameter source
// ------------------------------------------------------------------------------------------------------------
// parameter p of candidate c of image b as seen by cascade level `level` (aggregation.py:120-143, 235-236, 268-269)
__device__ __forceinline__ float cascade_param(const HoiDev& h, int b, int c, int p, int level) {
  const int S = h.a.S;
  if (param_level(p) < level) return h.fused[b * 48 + p];
  if (c < S) return h.a.hand_pose_diff[((size_t)b * S + c) * 48 + p];
  if (p < 3) return h.a.hand_pose_diff[((size_t)b * S + (c - S)) * 48 + p];   // :141-143 wrist copied from candidate c-S
  return h.a.hand_pose_reg[(size_t)b * 48 + p];
}

template <int TC>
struct HandScoreSmem {
  ManoSmem<TC> mano;
  float tipv[TC][5][3];
  float j21[TC][21][3];
  float heat[TC][21];
};

// one CTA = TC candidates of one image: joints-only MANO -> projection -> bicubic heat sampling -> finger scores
template <int TC>
__global__ void __launch_bounds__(128) k_hand_level_score(ManoModelDev m, HoiDev h, int level) {
  pdl_wait();          // launched with VPHO_LAUNCH_PDL
  pdl_trigger();
  VPHO_DYN_SMEM(HandScoreSmem<TC>, sp);
  HandScoreSmem<TC>& s = *sp;
  const int b = blockIdx.y, c0 = blockIdx.x * TC, tid = threadIdx.x, nt = blockDim.x;
  const int S = h.a.S;
  const int ncand = (level == 0) ? 2 * S : S + 1;     // levels 1-3: candidates S..2S-1 are exact duplicates of S
  mano_pose_setup<TC>(
      m,
      [&](int c, int j, float* a) {
        if (c0 + c >= ncand) return false;
        a[0] = cascade_param(h, b, c0 + c, 3 * j + 0, level);
        a[1] = cascade_param(h, b, c0 + c, 3 * j + 1, level);
        a[2] = cascade_param(h, b, c0 + c, 3 * j + 2, level);
        return true;
      },
      [&](int c, float* beta) {
        if (c0 + c >= ncand) return false;
        const int cs = (c0 + c) % S;                     // shape.repeat(1, 2, 1)  (aggregation.py:126)
        const float* sp2 = h.a.hand_shape + ((size_t)b * S + cs) * 10;
#pragma unroll
        for (int k = 0; k < 10; ++k) beta[k] = sp2[k];
        return true;
      },
      s.mano);
  // fingertip rest positions: same summation order as the full skinning kernel (template, then k = 0..144)
  for (int it = tid; it < TC * 15; it += nt) {
    const int c = it / 15, td = it % 15;
    float acc = m.tip_template[td];
    const float* dir = m.tip_dirs + (size_t)td * kBlendK;
    for (int k = 0; k < kBlendK; ++k) acc = fmaf(dir[k], s.mano.coefT[k][c], acc);
    s.tipv[c][td / 3][td % 3] = acc;
  }
  __syncthreads();
  for (int it = tid; it < TC * 21; it += nt) {
    const int c = it / 21, q = it % 21;
    float o[3];
    if (q < 16) {
#pragma unroll
      for (int d = 0; d < 3; ++d) o[d] = mano_center_scale(s.mano.G[c][q][d * 4 + 3], s.mano.G[c][0][d * 4 + 3]);
#pragma unroll
      for (int d = 0; d < 3; ++d) s.j21[c][joint16_to_21(q)][d] = o[d];
    } else {
      const int t = q - 16;
      float w[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) w[j] = m.tip_weights[t * 16 + j];
      mano_skin_point<TC>(s.mano, c, w, s.tipv[c][t], o);
#pragma unroll
      for (int d = 0; d < 3; ++d) s.j21[c][tip_to_21(t)][d] = mano_center_scale(o[d], s.mano.G[c][0][d * 4 + 3]);
    }
  }
  __syncthreads();
  const float* K = h.a.cam_intrinsic + (size_t)b * 9;
  const float* bbox = h.a.hand_bbox + (size_t)b * 4;
  const float* root = h.a.root_joint_flip + (size_t)b * 3;
  for (int it = tid; it < TC * 20; it += nt) {
    const int c = it / 20, q = 1 + it % 20;           // joints 1..20 (the wrist is never observed)
    const int jl = ((q - 1) & 3) + 1;                 // level of this joint
    float hv = 0.f;
    if (jl > level && c0 + c < ncand) {
      float gx, gy;
      project_to_grid(K, bbox, s.j21[c][q][0] + root[0], s.j21[c][q][1] + root[1], s.j21[c][q][2] + root[2], gx, gy);
      hv = bicubic_sample64(h.a.hand_heatmap + ((size_t)b * 21 + q) * kHm * kHm, gx, gy);
    }
    s.heat[c][q] = hv;
  }
  __syncthreads();
  for (int it = tid; it < TC * 5; it += nt) {
    const int c = it / 5, f = it % 5;
    if (c0 + c >= ncand) continue;
    float* dst = h.hscore + ((size_t)b * 2 * S + c0 + c) * 5;
    if (level == 0) {
      if (f != 0) continue;
      float acc = 0.f;                                 // heat_val.sum(-1) in observe_index order (level-major)
      for (int l = 1; l <= 4; ++l)
        for (int ff = 0; ff < 5; ++ff) acc += s.heat[c][finger_joint21(ff, l)];
      dst[0] = acc;
    } else {
      float acc = 0.f;                                 // mean over the observed levels of this finger (:245)
      for (int l = level + 1; l <= 4; ++l) acc += s.heat[c][finger_joint21(f, l)];
      dst[f] = acc / (float)(4 - level);
    }
  }
}

// one CTA per image, one warp per finger list: top-k -> weights -> weighted quaternion average -> fused axis-angle
template <int EL>
__global__ void __launch_bounds__(160) k_hand_level_fuse(HoiDev h, int level) {
  pdl_wait();          // launched with VPHO_LAUNCH_PDL
  pdl_trigger();
  __shared__ float s_val[5][64];
  __shared__ int s_idx[5][64];
  __shared__ float s_q[5][64][4];
  const int b = blockIdx.x, f = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = h.a.S, K = h.a.topk_hand, n = 2 * S;
  if (level == 0 && f != 0) return;
  const float* sc = h.hscore + (size_t)b * n * 5;
  auto value_of = [&](int i) { return sc[(size_t)((level > 0 && i > S) ? S : i) * 5 + (level == 0 ? 0 : f)]; };
  warp_topk<EL>(n, K, value_of, s_val[f], s_idx[f], lane);
  if (h.a.dbg_hand_score) {
    for (int i = lane; i < n; i += 32)
      h.a.dbg_hand_score[(((size_t)level * h.a.bs + b) * n + i) * 5 + f] = value_of(i);
  }
  if (h.a.dbg_hand_topk)
    for (int r = lane; r < K; r += 32) h.a.dbg_hand_topk[(((size_t)level * h.a.bs + b) * 5 + f) * K + r] = s_idx[f][r];
  const int jm = (level == 0) ? 0 : finger_mano_joint(f, level);
  if (level == 3 && h.l4) {
    for (int r = lane; r < K; r += 32)
      for (int d = 0; d < 3; ++d)
        h.l4[(((size_t)b * K + r) * 5 + f) * 3 + d] = cascade_param(h, b, s_idx[f][r], 3 * jm + d, level);
  }
  // signed unit quaternions of the K winners, one candidate per lane (the parameter gathers are independent loads)
  for (int r = lane; r < K; r += 32) {
    float aa[3], q[4];
#pragma unroll
    for (int d = 0; d < 3; ++d) aa[d] = cascade_param(h, b, s_idx[f][r], 3 * jm + d, level);
    axis_angle_to_quaternion(aa, q);
    const float sg = q[0] > 0.f ? 1.f : -1.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) s_q[f][r][i] = sg * q[i];
  }
  __syncwarp();
  // lane 4 i + j accumulates entry (i, j) of the weighted outer-product sum over the winners in rank order
  float vsum = 0.f;
  for (int r = 0; r < K; ++r) vsum += s_val[f][r];
  float a_ij = 0.f, wsum = 0.f;
  {
    const int i = (lane >> 2) & 3, j = lane & 3;
    for (int r = 0; r < K; ++r) {
      const float w = (s_val[f][r] + 1e-8f) / (vsum + 1e-8f);     // aggregation.py:218, 247
      wsum += w;
      a_ij += (s_q[f][r][i] * s_q[f][r][j]) * w;
    }
    a_ij /= wsum;
  }
  float A[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) A[i] = __shfl_sync(0xffffffffu, a_ij, i);
  if (lane == 0) {
    float qm[4], faa[3];
    sym4_top_eigvec(A, qm);
    quaternion_to_axis_angle(qm, faa);
#pragma unroll
    for (int d = 0; d < 3; ++d) h.fused[b * 48 + 3 * jm + d] = faa[d];
  }
}

__global__ void k_copy_cascade_pose(HoiDev h) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (h.a.dbg_cascade_pose && i < h.a.bs * 48) h.a.dbg_cascade_pose[i] = h.fused[i];
}

// ------------------------------------------------------------------------------------------------------------
// force anchors: one CTA per posed hand
// ------------------------------------------------------------------------------------------------------------
// verts [n][778][3] wrist-centred; root [n/group][3] is added first (vert_cam = vert + root_joint_flip);
// force_local [n/group][32][3]; outputs point/force [n][32][3]
__global__ void __launch_bounds__(256) k_force_anchors(AssetsDev as, const float* __restrict__ verts,
                                                       const float* __restrict__ root, const float* __restrict__ force_local,
                                                       int n, int group, float* __restrict__ point, float* __restrict__ force) {
  pdl_wait();          // launched with VPHO_LAUNCH_PDL
  pdl_trigger();
  __shared__ float j21[21 * 3];
  const int i = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* v = verts + (size_t)i * kVerts * 3;
  float r[3] = {0.f, 0.f, 0.f};
  if (root) { r[0] = root[(i / group) * 3 + 0]; r[1] = root[(i / group) * 3 + 1]; r[2] = root[(i / group) * 3 + 2]; }
  // joints = vert2joint . vert_cam   (hand_fn.py:436-448)
  for (int o = warp; o < 63; o += (int)(blockDim.x >> 5)) {
    const int k = o / 3, d = o % 3;
    float acc = 0.f;
    for (int vv = lane; vv < kVerts; vv += 32) acc = fmaf(as.v2j[k * kVerts + vv], v[vv * 3 + d] + r[d], acc);
#pragma unroll
    for (int sft = 16; sft > 0; sft >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, sft);
    if (lane == 0) j21[o] = acc;
  }
  __syncthreads();
  if (tid < kAnchors) {
    float pt[3], fg[3];
    const float* fl = force_local + ((size_t)(i / group) * kAnchors + tid) * 3;
    anchor_point_and_force(
        as, tid, [&](int vid, float* out) { out[0] = v[vid * 3 + 0] + r[0]; out[1] = v[vid * 3 + 1] + r[1]; out[2] = v[vid * 3 + 2] + r[2]; },
        j21, fl, pt, fg);
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      point[((size_t)i * kAnchors + tid) * 3 + d] = pt[d];
      force[((size_t)i * kAnchors + tid) * 3 + d] = fg[d];
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// object: key-point heat score of every candidate pose (one warp per candidate, one lane per key-point)
// ------------------------------------------------------------------------------------------------------------
// pose [bs][C][9] f64; transl (optional) [bs][3] f64 replaces every candidate's translation (aggregation.py:1213-1216)
__global__ void __launch_bounds__(256) k_obj_heat_score(AssetsDev as, HoiDev h, const double* __restrict__ pose,
                                                        const double* __restrict__ transl, int C, float* __restrict__ score) {
  pdl_wait();          // launched with VPHO_LAUNCH_PDL
  pdl_trigger();
  __shared__ float hv[8][32];
  const int b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x * 8 + warp;
  if (c >= C) return;
  ObjPose op;
  make_obj_pose(pose + ((size_t)b * C + c) * 9, transl ? transl + (size_t)b * 3 : nullptr, h.a.root_joint + (size_t)b * 3,
                h.a.is_right[b] != 0, op);
  float v = 0.f;
  if (lane < kKpts) {
    float pt[3], gx, gy;
    obj_point(op, as.kpt + ((size_t)obj_index(as, h.a.obj_id[b]) * kKpts + lane) * 3, pt);
    project_to_grid(h.a.cam_intrinsic + (size_t)b * 9, h.a.obj_bbox + (size_t)b * 4, pt[0], pt[1], pt[2], gx, gy);
    v = bicubic_sample64(h.a.obj_heatmap + ((size_t)b * kKpts + lane) * kHm * kHm, gx, gy);
  }
  hv[warp][lane] = v;
  __syncwarp();
  if (lane == 0) {
    float acc = 0.f;
    for (int k = 0; k < kKpts; ++k) acc += hv[warp][k];
    score[(size_t)b * C + c] = acc;
  }
}

// step 1 (aggregation.py:1200-1211): top-k on the heat score -> weights -> fused translation (float64)
template <int EL>
__global__ void __launch_bounds__(32) k_obj_transl_fuse(HoiDev h) {
  pdl_wait();          // launched with VPHO_LAUNCH_PDL
  pdl_trigger();
  __shared__ float s_val[64];
  __shared__ int s_idx[64];
  const int b = blockIdx.x, lane = threadIdx.x;
  const int S = h.a.S, K = h.a.topk_obj;
  const float* sc = h.oscore + (size_t)b * S;
  warp_topk<EL>(S, K, [&](int i) { return sc[i]; }, s_val, s_idx, lane);
  for (int r = lane; r < K; r += 32) h.t_topk[b * K + r] = s_idx[r];
  if (h.a.dbg_obj_topk) for (int r = lane; r < K; r += 32) h.a.dbg_obj_topk[((size_t)0 * h.a.bs + b) * max(K, h.a.phy_topk) + r] = s_idx[r];
  if (lane < 3) {
    float vsum = 0.f;
    for (int r = 0; r < K; ++r) vsum += s_val[r];
    double acc = 0.0;
    for (int r = 0; r < K; ++r) {
      const float w = (s_val[r] + 1e-8f) / (vsum + 1e-8f);
      acc += h.a.obj_pose6d[((size_t)b * S + s_idx[r]) * 9 + 6 + lane] * (double)w;
    }
    h.t_fused[b * 3 + lane] = acc;
  }
}

// step 2 (aggregation.py:1218-1242): top-k rotations under the fused translation -> K x K recombination
template <int EL>
__global__ void __launch_bounds__(32) k_obj_recombine(HoiDev h) {
  pdl_wait();          // launched with VPHO_LAUNCH_PDL
  pdl_trigger();
  __shared__ float s_val[64];
  __shared__ int s_idx[64];
  const int b = blockIdx.x, lane = threadIdx.x;
  const int S = h.a.S, K = h.a.topk_obj;
  const float* sc = h.oscore + (size_t)b * S;
  warp_topk<EL>(S, K, [&](int i) { return sc[i]; }, s_val, s_idx, lane);
  if (h.a.dbg_obj_topk) for (int r = lane; r < K; r += 32) h.a.dbg_obj_topk[((size_t)1 * h.a.bs + b) * max(K, h.a.phy_topk) + r] = s_idx[r];
  for (int it = lane; it < K * K * 9; it += 32) {
    const int cidx = it / 9, e = it % 9, i = cidx / K, j = cidx % K;
    const int src = (e < 6) ? s_idx[j] : h.t_topk[b * K + i];
    h.a.pose6d_candidate[((size_t)b * K * K + cidx) * 9 + e] = h.a.obj_pose6d[((size_t)b * S + src) * 9 + e];
  }
}

// ------------------------------------------------------------------------------------------------------------
// anchor -> posed surface nearest point scan: lane j owns anchor j, warps split the points of a tile
// ------------------------------------------------------------------------------------------------------------
constexpr int kScanThreads = 256;

#ifndef VPHO_EMU
// packed FP32 pairs (sm_100 FADD2 / FMUL2 / FFMA2): each half is an ordinary IEEE single-precision operation
__device__ __forceinline__ unsigned long long pack_f32x2(float a, float b) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void unpack_f32x2(unsigned long long v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ unsigned long long sub_f32x2(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ unsigned long long mul_f32x2(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ unsigned long long fma_f32x2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
#endif

// pts(i, out[3]) yields posed point i; results (min squared distance, first arg-min) of the 32 anchors land in
// s_d2[0..32), s_arg[0..32) (shared) after the call.  All kScanThreads threads of the CTA must call it.
template <typename PtFn>
__device__ __forceinline__ void anchor_nearest_scan(int n_pts, const float* anchor /*[3] of this lane's anchor*/, PtFn pts,
                                                    float4* tile /*[kScanThreads]*/, float* red_d2 /*[8][32]*/,
                                                    int* red_arg /*[8][32]*/, float* s_d2, int* s_arg) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float best = INFINITY;
  int arg = 0;
  const float ax = anchor[0], ay = anchor[1], az = anchor[2];
#ifndef VPHO_EMU
  // Packed-FP32 scan: the tile is kept as three coordinate planes and two points are handled per instruction
  // (FADD2 / FMUL2 / FFMA2).  Every lane of a pair performs exactly the scalar sequence of the loop below,
  // d2 = fma(dz, dz, fma(dx, dx, dy * dy)), so distances, ties and the first arg-min are unchanged.  Slots past the last
  // point hold a far-away coordinate: they can never win against the real points that precede them in a warp's slice.
  float* tx = reinterpret_cast<float*>(tile);
  float* ty = tx + kScanThreads;
  float* tz = ty + kScanThreads;
  const unsigned long long A = pack_f32x2(ax, ax), B = pack_f32x2(ay, ay), C = pack_f32x2(az, az);
  for (int p0 = 0; p0 < n_pts; p0 += kScanThreads) {
    __syncthreads();
    {
      float o[3] = {1e18f, 1e18f, 1e18f};
      if (p0 + tid < n_pts) pts(p0 + tid, o);
      tx[tid] = o[0]; ty[tid] = o[1]; tz[tid] = o[2];
    }
    __syncthreads();
    const int base = warp * 32;
    if (n_pts - p0 - base <= 0) continue;
#pragma unroll 4
    for (int q = 0; q < 32; q += 2) {
      const unsigned long long X = *reinterpret_cast<const unsigned long long*>(tx + base + q);
      const unsigned long long Y = *reinterpret_cast<const unsigned long long*>(ty + base + q);
      const unsigned long long Z = *reinterpret_cast<const unsigned long long*>(tz + base + q);
      const unsigned long long dx = sub_f32x2(A, X), dy = sub_f32x2(B, Y), dz = sub_f32x2(C, Z);
      const unsigned long long d = fma_f32x2(dz, dz, fma_f32x2(dx, dx, mul_f32x2(dy, dy)));
      float d0, d1;
      unpack_f32x2(d, d0, d1);
      if (d0 < best) { best = d0; arg = p0 + base + q; }
      if (d1 < best) { best = d1; arg = p0 + base + q + 1; }
    }
  }
#else
  for (int p0 = 0; p0 < n_pts; p0 += kScanThreads) {
    __syncthreads();
    {
      float o[3] = {0.f, 0.f, 0.f};
      if (p0 + tid < n_pts) pts(p0 + tid, o);
      tile[tid] = make_float4(o[0], o[1], o[2], 0.f);
    }
    __syncthreads();
    const int base = warp * 32;
    const int cnt = min(32, n_pts - p0 - base);
    for (int q = 0; q < cnt; ++q) {
      const float4 p = tile[base + q];
      const float dx = ax - p.x, dy = ay - p.y, dz = az - p.z;
      const float d2 = (dx * dx + dy * dy) + dz * dz;
      if (d2 < best) { best = d2; arg = p0 + base + q; }
    }
  }
#endif
  red_d2[warp * 32 + lane] = best;
  red_arg[warp * 32 + lane] = arg;
  __syncthreads();
  if (warp == 0) {
    float bd = red_d2[lane];
    int ba = red_arg[lane];
    for (int w = 1; w < kScanThreads / 32; ++w) {
      const float d = red_d2[w * 32 + lane];
      const int a = red_arg[w * 32 + lane];
      if (d < bd || (d == bd && a < ba)) { bd = d; ba = a; }
    }
    s_d2[lane] = bd;
    s_arg[lane] = ba;
  }
  __syncthreads();
}

// physics3 score of one recombined object candidate (aggregation.py:947-997): grid (kk, bs)
__global__ void __launch_bounds__(kScanThreads) k_obj_physics3(AssetsDev as, HoiDev h) {
  pdl_wait();          // launched with VPHO_LAUNCH_PDL
  pdl_trigger();
  __shared__ float4 tile[kScanThreads];
  __shared__ float red_d2[8 * 32];
  __shared__ int red_arg[8 * 32];
  __shared__ float s_d2[32];
  __shared__ int s_arg[32];
  __shared__ float s_term[32], s_cross[32][3];
  const int c = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, lane = tid & 31;
  ObjPose op;
  make_obj_pose(h.a.pose6d_candidate + ((size_t)b * h.kk + c) * 9, nullptr, h.a.root_joint + (size_t)b * 3,
                h.a.is_right[b] != 0, op);
  const float* base_pts = as.verts + (size_t)obj_index(as, h.a.obj_id[b]) * as.n_pts * 3;
  float anchor[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) anchor[d] = h.fpoint[((size_t)b * kAnchors + lane) * 3 + d];
  auto pts = [&](int i, float* o) { obj_point(op, base_pts + (size_t)i * 3, o); };
  anchor_nearest_scan(as.n_pts, anchor, pts, tile, red_d2, red_arg, s_d2, s_arg);
  if (tid < kAnchors) {
    const int j = tid;
    float fg[3], fn, fsum = 0.f;
#pragma unroll
    for (int d = 0; d < 3; ++d) fg[d] = h.fglobal[((size_t)b * kAnchors + j) * 3 + d];
    fn = sqrtf((fg[0] * fg[0] + fg[1] * fg[1]) + fg[2] * fg[2]);
    s_term[j] = fn;
    __syncwarp();
    for (int k = 0; k < kAnchors; ++k) fsum += s_term[k];     // force_norm.sum(-1)
    __syncwarp();
    const float fw = fn / fsum;
    const float dist = sqrtf(s_d2[j]);
    float vstar[3], com[3];
    pts(s_arg[j], vstar);
    obj_point(op, as.com + (size_t)obj_index(as, h.a.obj_id[b]) * 3, com);
    float r[3], fdir[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) { r[d] = (anchor[d] - vstar[d]) - com[d]; fdir[d] = fg[d] / fn; }
    s_term[j] = dist * fw;
    s_cross[j][0] = fdir[1] * r[2] - fdir[2] * r[1];
    s_cross[j][1] = fdir[2] * r[0] - fdir[0] * r[2];
    s_cross[j][2] = fdir[0] * r[1] - fdir[1] * r[0];
    __syncwarp();
    if (j == 0) {
      float sd = 0.f, L[3] = {0.f, 0.f, 0.f};
      for (int k = 0; k < kAnchors; ++k) {
        sd += s_term[k];
        L[0] += s_cross[k][0]; L[1] += s_cross[k][1]; L[2] += s_cross[k][2];
      }
      const float Ln = sqrtf((L[0] * L[0] + L[1] * L[1]) + L[2] * L[2]);
      h.pscore[(size_t)b * h.kk + c] = -(sd * Ln);
    }
  }
}

// final object selection + fusion (aggregation.py:1247-1287): one CTA per image
template <int EL>
__global__ void __launch_bounds__(256) k_obj_final(AssetsDev as, HoiDev h) {
  pdl_wait();          // launched with VPHO_LAUNCH_PDL
  pdl_trigger();
  __shared__ float p_val[64], q_val[64];
  __shared__ int p_idx[64], q_idx[64];
  __shared__ double s_pose[9];
  __shared__ double s_q[64][4], s_t[64][3];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int kk = h.kk, K5 = h.a.phy_topk;
  if (warp == 0) warp_topk<EL>(kk, K5, [&](int i) { return h.pscore[(size_t)b * kk + i]; }, p_val, p_idx, lane);
  if (warp == 1) warp_topk<EL>(kk, K5, [&](int i) { return h.oscore[(size_t)b * kk + i]; }, q_val, q_idx, lane);
  __syncthreads();
  if (h.a.dbg_obj_topk && tid < K5) {
    h.a.dbg_obj_topk[((size_t)2 * h.a.bs + b) * max(h.a.topk_obj, K5) + tid] = p_idx[tid];
    h.a.dbg_obj_topk[((size_t)3 * h.a.bs + b) * max(h.a.topk_obj, K5) + tid] = q_idx[tid];
  }
  // signed quaternions of the K5 winners, one candidate per thread (float64: rot6d -> matrix -> quaternion)
  const bool grasped = h.a.is_grasped[b] != 0;
  const double* cand = h.a.pose6d_candidate + (size_t)b * kk * 9;
  if (tid < K5) {
    const int* idx = grasped ? p_idx : q_idx;
    const double* p = cand + (size_t)idx[tid] * 9;
    double R[9], q[4];
    rot6d_to_matrix(p, R);
    matrix_to_quaternion(R, q);
    const double sg = q[0] > 0.0 ? 1.0 : -1.0;
    for (int i = 0; i < 4; ++i) s_q[tid][i] = sg * q[i];
    for (int d = 0; d < 3; ++d) s_t[tid][d] = p[6 + d];
  }
  __syncthreads();
  if (tid == 0) {
    float w[64];
    if (grasped) {
      // weight = ones / ones.sum()   (aggregation.py:990-991)
      float ws = 0.f;
      for (int r = 0; r < K5; ++r) ws += 1.f;
      for (int r = 0; r < K5; ++r) w[r] = 1.f / ws;
    } else {
      float vs = 0.f;
      for (int r = 0; r < K5; ++r) vs += q_val[r];
      for (int r = 0; r < K5; ++r) w[r] = (q_val[r] + 1e-8f) / (vs + 1e-8f);
    }
    // fuse_topk (aggregation.py:729-740): float64 poses, float32 weights
    double tr[3] = {0, 0, 0};
    double A[16];
    for (int i = 0; i < 16; ++i) A[i] = 0.0;
    float wsum = 0.f;
    for (int r = 0; r < K5; ++r) {
      for (int d = 0; d < 3; ++d) tr[d] += s_t[r][d] * (double)w[r];
      wsum += w[r];
      for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) A[i * 4 + j] += (s_q[r][i] * s_q[r][j]) * (double)w[r];
    }
    for (int i = 0; i < 16; ++i) A[i] /= (double)wsum;
    double qm[4], Rm[9];
    sym4_top_eigvec(A, qm);
    quaternion_to_matrix(qm, Rm);
    for (int e = 0; e < 6; ++e) s_pose[e] = Rm[e];          // matrix_to_rotation_6d: first two rows
    for (int d = 0; d < 3; ++d) s_pose[6 + d] = tr[d];
    for (int e = 0; e < 9; ++e) h.a.obj_agg_6d[(size_t)b * 9 + e] = s_pose[e];
  }
  __syncthreads();
  // posed surface + centre of mass of the fused object (aggregation.py:1282-1287)
  ObjPose op;
  make_obj_pose(s_pose, nullptr, h.a.root_joint + (size_t)b * 3, h.a.is_right[b] != 0, op);
  const float* base_pts = as.verts + (size_t)obj_index(as, h.a.obj_id[b]) * as.n_pts * 3;
  for (int i = tid; i < as.n_pts; i += blockDim.x) {
    float o[3];
    obj_point(op, base_pts + (size_t)i * 3, o);
    float* dst = h.a.agg_obj_vert + ((size_t)b * as.n_pts + i) * 3;
    dst[0] = o[0]; dst[1] = o[1]; dst[2] = o[2];
  }
  if (tid == 0) {
    float o[3];
    obj_point(op, as.com + (size_t)obj_index(as, h.a.obj_id[b]) * 3, o);
    h.ocom[b * 3 + 0] = o[0]; h.ocom[b * 3 + 1] = o[1]; h.ocom[b * 3 + 2] = o[2];
  }
}

// ------------------------------------------------------------------------------------------------------------
// hand physics refinement (aggregation.py:1306-1337, 537-626)
// ------------------------------------------------------------------------------------------------------------
// candidate poses: cascade-fused pose with the DIP (level-3) parameters of rank r's per-finger winners; last = fused
__global__ void k_build_phys_pose(HoiDev h) {
  pdl_wait();          // launched with VPHO_LAUNCH_PDL
  pdl_trigger();
  const int b = blockIdx.x, K = h.a.topk_hand, nc = h.nc;
  for (int it = threadIdx.x; it < nc * 48; it += blockDim.x) {
    const int r = it / 48, p = it % 48;
    float v = h.fused[b * 48 + p];
    if (r < K && param_level(p) == 3) {
      const int jm = p / 3;
      int f = 0;
      for (int ff = 0; ff < 5; ++ff)
        if (finger_mano_joint(ff, 3) == jm) f = ff;
      v = h.l4[(((size_t)b * K + r) * 5 + f) * 3 + (p % 3)];
    }
    h.ppose[((size_t)b * nc + r) * 48 + p] = v;
  }
  for (int it = threadIdx.x; it < nc * 10; it += blockDim.x)
    h.pshape[(size_t)b * nc * 10 + it] = h.a.hand_shape[((size_t)b * h.a.S) * 10 + it % 10];
}

// grid (nc, bs): anchors of candidate (b, r) against the fused object's posed surface -> per-finger scores
__global__ void __launch_bounds__(kScanThreads) k_hand_phys_score(HoiDev h) {
  pdl_wait();          // launched with VPHO_LAUNCH_PDL
  pdl_trigger();
  __shared__ float4 tile[kScanThreads];
  __shared__ float red_d2[8 * 32];
  __shared__ int red_arg[8 * 32];
  __shared__ float s_d2[32];
  __shared__ int s_arg[32];
  __shared__ float s_fn[32], s_dir[32][3], s_score[32];
  const int r = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, lane = tid & 31;
  const size_t ci = (size_t)b * h.nc + r;
  float anchor[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) anchor[d] = h.ppoint[(ci * kAnchors + lane) * 3 + d];
  const float* ov = h.a.agg_obj_vert + (size_t)b * h.n_pts * 3;
  auto pts = [&](int i, float* o) { o[0] = ov[(size_t)i * 3 + 0]; o[1] = ov[(size_t)i * 3 + 1]; o[2] = ov[(size_t)i * 3 + 2]; };
  anchor_nearest_scan(h.n_pts, anchor, pts, tile, red_d2, red_arg, s_d2, s_arg);
  if (tid < kAnchors) {
    const int j = tid;
    float fg[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) fg[d] = h.pforce[(ci * kAnchors + j) * 3 + d];
    const float fn = sqrtf((fg[0] * fg[0] + fg[1] * fg[1]) + fg[2] * fg[2]);
    s_fn[j] = fn;
#pragma unroll
    for (int d = 0; d < 3; ++d) s_dir[j][d] = fg[d] / fn;
    __syncwarp();
    float fsum = 0.f, I[3] = {0.f, 0.f, 0.f};
    for (int k = 0; k < kAnchors; ++k) {
      fsum += s_fn[k];
      I[0] += s_dir[k][0]; I[1] += s_dir[k][1]; I[2] += s_dir[k][2];
    }
    const float In = sqrtf((I[0] * I[0] + I[1] * I[1]) + I[2] * I[2]);
    const float fw = fn / fsum;
    s_score[j] = -((fw * sqrtf(s_d2[j])) * In);                  // aggregation.py:566, 577-579
    __syncwarp();
    if (j < 5) {
      float acc = 0.f;
      for (int i = 0; i < 4; ++i) acc += s_score[finger_anchor(j, i)];
      h.fscore[((size_t)b * 5 + j) * h.nc + r] = acc;
    }
  }
}

// one CTA per image, one warp per finger: top-5 candidates -> unweighted quaternion average of PIP and DIP
template <int EL>
__global__ void __launch_bounds__(160) k_hand_phys_fuse(HoiDev h) {
  pdl_wait();          // launched with VPHO_LAUNCH_PDL
  pdl_trigger();
  __shared__ float s_val[5][64];
  __shared__ int s_idx[5][64];
  const int b = blockIdx.x, f = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nc = h.nc, K5 = h.a.phy_topk;
  const float* sc = h.fscore + ((size_t)b * 5 + f) * nc;
  warp_topk<EL>(nc, K5, [&](int i) { return sc[i]; }, s_val[f], s_idx[f], lane);
  if (h.a.dbg_finger_score) for (int i = lane; i < nc; i += 32) h.a.dbg_finger_score[((size_t)b * 5 + f) * nc + i] = sc[i];
  if (h.a.dbg_finger_topk) for (int r = lane; r < K5; r += 32) h.a.dbg_finger_topk[((size_t)b * 5 + f) * K5 + r] = s_idx[f][r];
  // fuse_pose = pose[:, 0].clone()   (aggregation.py:601): every non-PIP/DIP parameter comes from candidate 0
  if (f == 0) {
    for (int p = lane; p < 48; p += 32)
      if (param_level(p) < 2) h.a.hand_agg_mano[(size_t)b * 58 + p] = h.ppose[((size_t)b * nc) * 48 + p];
    for (int k = lane; k < 10; k += 32) h.a.hand_agg_mano[(size_t)b * 58 + 48 + k] = h.a.hand_shape[((size_t)b * h.a.S) * 10 + k];
  }
  if (lane < 2) {
    const int jm = finger_mano_joint(f, 2 + lane);
    float A[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) A[i] = 0.f;
    float wsum = 0.f;
    for (int r = 0; r < K5; ++r) {
      const float* pp = h.ppose + ((size_t)b * nc + s_idx[f][r]) * 48 + 3 * jm;
      float aa[3] = {pp[0], pp[1], pp[2]}, q[4];
      axis_angle_to_quaternion(aa, q);
      const float sg = q[0] > 0.f ? 1.f : -1.f;
      wsum += 1.f;
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) A[i * 4 + j] += ((sg * q[i]) * (sg * q[j])) * 1.f;
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) A[i] /= wsum;
    float qm[4], faa[3];
    sym4_top_eigvec(A, qm);
    quaternion_to_axis_angle(qm, faa);
#pragma unroll
    for (int d = 0; d < 3; ++d) h.a.hand_agg_mano[(size_t)b * 58 + 3 * jm + d] = faa[d];
  }
}

__global__ void k_copy_f32(const float* __restrict__ src, float* __restrict__ dst, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i];
}

// stand-alone HeadObject.forward + flip_pt3d: out[b][c][v][3]
__global__ void k_object_points(AssetsDev as, const float* __restrict__ pose6d, const int* __restrict__ obj_id,
                                const unsigned char* __restrict__ is_right, int C, int which, int flip, int V,
                                float* __restrict__ out) {
  const int b = blockIdx.z, c = blockIdx.y;
  const float* p = pose6d + ((size_t)b * C + c) * 9;
  ObjPose op;
  float d6[6] = {p[0], p[1], p[2], p[3], p[4], p[5]};
  rot6d_to_matrix(d6, op.R);
  op.t[0] = p[6]; op.t[1] = p[7]; op.t[2] = p[8];
  op.flip = flip && is_right && !is_right[b];
  const float* tab = which == 0 ? as.kpt + (size_t)obj_index(as, obj_id[b]) * kKpts * 3
                                : (which == 1 ? as.verts + (size_t)obj_index(as, obj_id[b]) * as.n_pts * 3 : as.com + (size_t)obj_index(as, obj_id[b]) * 3);
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < V; v += gridDim.x * blockDim.x) {
    float o[3];
    obj_point(op, tab + (size_t)v * 3, o);
    float* dst = out + (((size_t)b * C + c) * V + v) * 3;
    dst[0] = o[0]; dst[1] = o[1]; dst[2] = o[2];
  }
}

// ------------------------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------------------------
static size_t up256(size_t x) { return (x + 255) / 256 * 256; }

static size_t hoi_carve(void* base, int bs, int S, int Kh, int Ko, int n_pts, HoiDev* h) {
  const int nc = Kh + 1, kk = Ko * Ko;
  const int omax = S > kk ? S : kk;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = up256(off + bytes); return o; };
  const size_t o_hscore = take((size_t)bs * 2 * S * 5 * 4), o_fused = take((size_t)bs * 48 * 4),
               o_l4 = take((size_t)bs * Kh * 15 * 4), o_cverts = take((size_t)bs * kVerts * 3 * 4),
               o_cjoints = take((size_t)bs * 63 * 4), o_fpoint = take((size_t)bs * 96 * 4),
               o_fglobal = take((size_t)bs * 96 * 4), o_oscore = take((size_t)bs * omax * 4),
               o_pscore = take((size_t)bs * kk * 4), o_ttopk = take((size_t)bs * Ko * 4),
               o_tfused = take((size_t)bs * 3 * 8), o_ppose = take((size_t)bs * nc * 48 * 4),
               o_pverts = take((size_t)bs * nc * kVerts * 3 * 4), o_pjoints = take((size_t)bs * nc * 63 * 4),
               o_ppoint = take((size_t)bs * nc * 96 * 4), o_pforce = take((size_t)bs * nc * 96 * 4),
               o_fscore = take((size_t)bs * 5 * nc * 4), o_ocom = take((size_t)bs * 3 * 4),
               o_pshape = take((size_t)bs * nc * 10 * 4);
  if (h) {
    char* p = static_cast<char*>(base);
    h->nc = nc; h->kk = kk; h->n_pts = n_pts;
    h->hscore = (float*)(p + o_hscore); h->fused = (float*)(p + o_fused); h->l4 = (float*)(p + o_l4);
    h->cverts = (float*)(p + o_cverts); h->cjoints = (float*)(p + o_cjoints); h->fpoint = (float*)(p + o_fpoint);
    h->fglobal = (float*)(p + o_fglobal); h->oscore = (float*)(p + o_oscore); h->pscore = (float*)(p + o_pscore);
    h->t_topk = (int*)(p + o_ttopk); h->t_fused = (double*)(p + o_tfused); h->ppose = (float*)(p + o_ppose);
    h->pverts = (float*)(p + o_pverts); h->pjoints = (float*)(p + o_pjoints); h->ppoint = (float*)(p + o_ppoint);
    h->pforce = (float*)(p + o_pforce); h->fscore = (float*)(p + o_fscore); h->ocom = (float*)(p + o_ocom); h->pshape = (float*)(p + o_pshape);
  }
  return off;
}

template <int EL>
static int run_hoi_enqueue(const ManoModelDev& m, AssetsHost& ah, const HoiDev& h, cudaStream_t st, bool* forked);

// Error paths between the fork and the join must not leave side-stream work un-joined while the caller frees or reuses
// the workspace: whatever happens, the caller's stream waits for the side stream before this returns.
template <int EL>
static int run_hoi(const ManoModelDev& m, AssetsHost& ah, const HoiDev& h, cudaStream_t st) {
  bool forked = false;
  const int rc = run_hoi_enqueue<EL>(m, ah, h, st, &forked);
#ifndef VPHO_EMU
  if (rc != VPHO_OK && forked) {
    if (cudaEventRecord(ah.ev_join, ah.side) == cudaSuccess) cudaStreamWaitEvent(st, ah.ev_join, 0);
  }
#endif
  return rc;
}

template <int EL>
static int run_hoi_enqueue(const ManoModelDev& m, AssetsHost& ah, const HoiDev& h, cudaStream_t st, bool* forked) {
  const AssetsDev& as = ah.dev;
  const vpho_hoi_args& a = h.a;
  const int bs = a.bs, S = a.S;
  constexpr int TC = 4;
  const size_t smem = sizeof(HandScoreSmem<TC>);
#ifndef VPHO_EMU
  {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev != ah.device) return VPHO_ERR_INVALID;      // handle belongs to another device
    static bool attr_set[64] = {};
    if (dev < 64 && !attr_set[dev]) {
      if (cudaFuncSetAttribute(k_hand_level_score<TC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return VPHO_ERR_LAUNCH;
      attr_set[dev] = true;
    }
  }
#endif
  // The object's translation / rotation selection and the K x K recombination do not depend on the hand: they run on a
  // library-owned side stream next to the hand cascade and are joined before the physics score of the recombined poses.
  cudaStream_t so = st;
#ifndef VPHO_EMU
  // vpho_set_pdl(0) (bench.py's serialised per-kernel pass) also keeps the object branch on the caller's stream
  const bool fork = pdl_enabled();
  cudaStream_t side = fork ? ah.side : st;
  cudaEvent_t ev_fork = ah.ev_fork, ev_join = ah.ev_join;
  if (fork) {
    if (cudaEventRecord(ev_fork, st) != cudaSuccess || cudaStreamWaitEvent(side, ev_fork, 0) != cudaSuccess) return VPHO_ERR_LAUNCH;
    *forked = true;
  }
  so = side;
#endif
  // ---- object: translation, rotation, recombination (aggregation.py:1200-1242)
  const int omax = S > h.kk ? S : h.kk;
  VPHO_LAUNCH_PDL(k_obj_heat_score, dim3((S + 7) / 8, bs), dim3(256), 0, so, as, h, a.obj_pose6d, (const double*)nullptr, S, h.oscore);
  if (a.dbg_obj_score) VPHO_LAUNCH(k_copy_f32, dim3((bs * S + 255) / 256), dim3(256), 0, so, h.oscore, a.dbg_obj_score + (size_t)0 * bs * omax, bs * S);
  VPHO_LAUNCH_PDL(k_obj_transl_fuse<EL>, dim3(bs), dim3(32), 0, so, h);
  VPHO_LAUNCH_PDL(k_obj_heat_score, dim3((S + 7) / 8, bs), dim3(256), 0, so, as, h, a.obj_pose6d, (const double*)h.t_fused, S, h.oscore);
  if (a.dbg_obj_score) VPHO_LAUNCH(k_copy_f32, dim3((bs * S + 255) / 256), dim3(256), 0, so, h.oscore, a.dbg_obj_score + (size_t)1 * bs * omax, bs * S);
  VPHO_LAUNCH_PDL(k_obj_recombine<EL>, dim3(bs), dim3(32), 0, so, h);
#ifndef VPHO_EMU
  if (fork && cudaEventRecord(ev_join, side) != cudaSuccess) return VPHO_ERR_LAUNCH;
#endif
  // ---- hand heat-map cascade (aggregation.py:115-178)
  for (int level = 0; level < 4; ++level) {
    const int ncand = level == 0 ? 2 * S : S + 1;
    profile_begin(VPHO_TAG_HAND_SCORE, st);
    VPHO_LAUNCH_PDL(k_hand_level_score<TC>, dim3((ncand + TC - 1) / TC, bs), dim3(128), smem, st, m, h, level);
    profile_end(VPHO_TAG_HAND_SCORE, st);
    VPHO_LAUNCH_PDL(k_hand_level_fuse<EL>, dim3(bs), dim3(160), 0, st, h, level);
  }
  VPHO_CHECK_LAUNCH();
  if (a.dbg_cascade_pose) VPHO_LAUNCH(k_copy_cascade_pose, dim3((bs * 48 + 255) / 256), dim3(256), 0, st, h);
  int rc = mano_forward_dev(m, h.fused, a.hand_shape, 48, 10 * S, bs, h.cverts, h.cjoints, st);
  if (rc) return rc;
  // ---- force anchors of the fused hand (aggregation.py:1195-1196)
  VPHO_LAUNCH_PDL(k_force_anchors, dim3(bs), dim3(256), 0, st, as, h.cverts, a.root_joint_flip, a.force_local, bs, 1, h.fpoint,
              h.fglobal);
  if (a.dbg_force_point) VPHO_LAUNCH(k_copy_f32, dim3((bs * 96 + 255) / 256), dim3(256), 0, st, h.fpoint, a.dbg_force_point, bs * 96);
  if (a.dbg_force_global) VPHO_LAUNCH(k_copy_f32, dim3((bs * 96 + 255) / 256), dim3(256), 0, st, h.fglobal, a.dbg_force_global, bs * 96);
#ifndef VPHO_EMU
  if (fork && cudaStreamWaitEvent(st, ev_join, 0) != cudaSuccess) return VPHO_ERR_LAUNCH;
  *forked = false;           // joined
#endif
  // ---- object: physics / heat-map selection of the recombined candidates, fusion (aggregation.py:1247-1287)
  profile_begin(VPHO_TAG_PHYSICS3, st);
  VPHO_LAUNCH_PDL(k_obj_physics3, dim3(h.kk, bs), dim3(kScanThreads), 0, st, as, h);
  profile_end(VPHO_TAG_PHYSICS3, st);
  VPHO_LAUNCH_PDL(k_obj_heat_score, dim3((h.kk + 7) / 8, bs), dim3(256), 0, st, as, h, (const double*)a.pose6d_candidate, (const double*)nullptr, h.kk, h.oscore);
  if (a.dbg_obj_score) {
    VPHO_LAUNCH(k_copy_f32, dim3((bs * h.kk + 255) / 256), dim3(256), 0, st, h.pscore, a.dbg_obj_score + (size_t)2 * bs * omax, bs * h.kk);
    VPHO_LAUNCH(k_copy_f32, dim3((bs * h.kk + 255) / 256), dim3(256), 0, st, h.oscore, a.dbg_obj_score + (size_t)3 * bs * omax, bs * h.kk);
  }
  VPHO_LAUNCH_PDL(k_obj_final<EL>, dim3(bs), dim3(256), 0, st, as, h);
  VPHO_CHECK_LAUNCH();
  // ---- hand physics refinement (aggregation.py:1306-1337)
  VPHO_LAUNCH_PDL(k_build_phys_pose, dim3(bs), dim3(256), 0, st, h);
  rc = mano_forward_dev(m, h.ppose, h.pshape, 48, 10, bs * h.nc, h.pverts, h.pjoints, st);
  if (rc) return rc;
  VPHO_LAUNCH_PDL(k_force_anchors, dim3(bs * h.nc), dim3(256), 0, st, as, h.pverts, a.root_joint_flip, a.force_local, bs * h.nc,
              h.nc, h.ppoint, h.pforce);
  profile_begin(VPHO_TAG_HAND_PHYS, st);
  VPHO_LAUNCH_PDL(k_hand_phys_score, dim3(h.nc, bs), dim3(kScanThreads), 0, st, h);
  profile_end(VPHO_TAG_HAND_PHYS, st);
  VPHO_LAUNCH_PDL(k_hand_phys_fuse<EL>, dim3(bs), dim3(160), 0, st, h);
  VPHO_CHECK_LAUNCH();
  return mano_forward_dev(m, a.hand_agg_mano, a.hand_agg_mano + 48, 58, 58, bs, a.hand_agg_vert, a.hand_agg_joint, st);
}

}  // namespace vpho

using namespace vpho;

extern "C" int vpho_assets_create(const int32_t* face_vertex_idx, const float* anchor_weight, const float* vert2joint,
                                  int n_obj, int n_pts, const float* kpt3d, const float* verts, const float* com,
                                  vpho_assets_t* out) {
  if (!face_vertex_idx || !anchor_weight || !vert2joint || n_obj <= 0 || n_pts <= 0 || !kpt3d || !verts || !com || !out)
    return VPHO_ERR_INVALID;
  for (int i = 0; i < 96; ++i)
    if (face_vertex_idx[i] < 0 || face_vertex_idx[i] >= kVerts) return VPHO_ERR_INVALID;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = up256(off + bytes); return o; };
  const size_t o_face = take(96 * 4), o_aw = take(64 * 4), o_v2j = take((size_t)21 * kVerts * 4),
               o_kpt = take((size_t)n_obj * kKpts * 3 * 4), o_verts = take((size_t)n_obj * n_pts * 3 * 4),
               o_com = take((size_t)n_obj * 3 * 4);
  std::vector<char> hbuf(off, 0);
  memcpy(hbuf.data() + o_face, face_vertex_idx, 96 * 4);
  memcpy(hbuf.data() + o_aw, anchor_weight, 64 * 4);
  memcpy(hbuf.data() + o_v2j, vert2joint, (size_t)21 * kVerts * 4);
  memcpy(hbuf.data() + o_kpt, kpt3d, (size_t)n_obj * kKpts * 3 * 4);
  memcpy(hbuf.data() + o_verts, verts, (size_t)n_obj * n_pts * 3 * 4);
  memcpy(hbuf.data() + o_com, com, (size_t)n_obj * 3 * 4);
  AssetsHost* ah = new AssetsHost();
  if (cudaMalloc(&ah->blob, off) != cudaSuccess) { delete ah; return VPHO_ERR_ALLOC; }
  if (cudaMemcpy(ah->blob, hbuf.data(), off, cudaMemcpyHostToDevice) != cudaSuccess) { cudaFree(ah->blob); delete ah; return VPHO_ERR_ALLOC; }
#ifndef VPHO_EMU
  if (cudaGetDevice(&ah->device) != cudaSuccess ||
      cudaStreamCreateWithFlags(&ah->side, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&ah->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&ah->ev_join, cudaEventDisableTiming) != cudaSuccess) {
    if (ah->ev_fork) cudaEventDestroy(ah->ev_fork);
    if (ah->side) cudaStreamDestroy(ah->side);
    cudaFree(ah->blob); delete ah; return VPHO_ERR_ALLOC;
  }
#endif
  char* b = static_cast<char*>(ah->blob);
  ah->dev.face = (const int*)(b + o_face); ah->dev.aw = (const float*)(b + o_aw); ah->dev.v2j = (const float*)(b + o_v2j);
  ah->dev.n_obj = n_obj; ah->dev.n_pts = n_pts;
  ah->dev.kpt = (const float*)(b + o_kpt); ah->dev.verts = (const float*)(b + o_verts); ah->dev.com = (const float*)(b + o_com);
  *out = ah;
  return VPHO_OK;
}

extern "C" int vpho_assets_destroy(vpho_assets_t h) {
  if (!h) return VPHO_ERR_INVALID;
  AssetsHost* ah = static_cast<AssetsHost*>(h);
#ifndef VPHO_EMU
  if (ah->side) { cudaStreamSynchronize(ah->side); cudaStreamDestroy(ah->side); }
  if (ah->ev_fork) cudaEventDestroy(ah->ev_fork);
  if (ah->ev_join) cudaEventDestroy(ah->ev_join);
#endif
  cudaFree(ah->blob);
  delete ah;
  return VPHO_OK;
}

extern "C" int vpho_object_points(vpho_assets_t h, const float* pose6d, const int32_t* obj_id, const uint8_t* is_right,
                                  int bs, int C, int which, int flip, float* out, void* stream) {
  if (!h || bs < 0 || C < 0 || which < 0 || which > 2) return VPHO_ERR_INVALID;
  if (bs == 0 || C == 0) return VPHO_OK;
  if (!pose6d || !obj_id || !out || (flip && !is_right)) return VPHO_ERR_INVALID;
  const AssetsDev& as = static_cast<AssetsHost*>(h)->dev;
  const int V = which == 0 ? kKpts : (which == 1 ? as.n_pts : 1);
  VPHO_LAUNCH(k_object_points, dim3((V + 255) / 256, C, bs), dim3(256), 0, (cudaStream_t)stream, as, pose6d, obj_id, is_right, C,
              which, flip, V, out);
  VPHO_CHECK_LAUNCH();
  return VPHO_OK;
}

extern "C" int vpho_force_anchors(vpho_assets_t h, const float* verts, const float* force_local, int n, int group,
                                  float* force_point, float* force_global, void* stream) {
  if (!h || n < 0 || group <= 0) return VPHO_ERR_INVALID;
  if (n == 0) return VPHO_OK;
  if (!verts || !force_local || !force_point || !force_global) return VPHO_ERR_INVALID;
  VPHO_LAUNCH(k_force_anchors, dim3(n), dim3(256), 0, (cudaStream_t)stream, static_cast<AssetsHost*>(h)->dev, verts,
              (const float*)nullptr, force_local, n, group, force_point, force_global);
  VPHO_CHECK_LAUNCH();
  return VPHO_OK;
}

extern "C" size_t vpho_hoi_workspace_bytes(int bs, int S, int topk_hand, int topk_obj, int n_pts) {
  if (bs < 0 || S <= 0 || topk_hand <= 0 || topk_obj <= 0) return 0;
  return hoi_carve(nullptr, bs, S, topk_hand, topk_obj, n_pts, nullptr);
}

extern "C" int vpho_hoi_aggregate(vpho_mano_t mano, vpho_assets_t assets, const vpho_hoi_args* args, void* workspace,
                                  size_t workspace_bytes, void* stream) {
  if (!mano || !assets || !args || !workspace) return VPHO_ERR_INVALID;
  const vpho_hoi_args& a = *args;
  AssetsHost& ah = *static_cast<AssetsHost*>(assets);
  const AssetsDev& as = ah.dev;
  if (a.bs < 0 || a.S <= 0) return VPHO_ERR_INVALID;
  if (a.bs == 0) return VPHO_OK;
  const int nc = a.topk_hand + 1, kk = a.topk_obj * a.topk_obj;
  if (a.topk_hand < 1 || a.topk_hand > 64 || a.topk_hand > 2 * a.S) return VPHO_ERR_INVALID;
  if (a.topk_obj < 1 || a.topk_obj > 16 || a.topk_obj > a.S) return VPHO_ERR_INVALID;
  if (a.phy_topk < 1 || a.phy_topk > 64 || a.phy_topk > nc || a.phy_topk > kk) return VPHO_ERR_INVALID;
  if (!a.cam_intrinsic || !a.root_joint_flip || !a.root_joint || !a.is_right || !a.is_grasped || !a.force_local ||
      !a.hand_pose_diff || !a.hand_pose_reg || !a.hand_shape || !a.hand_heatmap || !a.hand_bbox || !a.obj_pose6d ||
      !a.obj_heatmap || !a.obj_bbox || !a.obj_id || !a.obj_agg_6d || !a.pose6d_candidate || !a.agg_obj_vert ||
      !a.hand_agg_mano || !a.hand_agg_vert || !a.hand_agg_joint)
    return VPHO_ERR_INVALID;
  HoiDev h;
  h.a = a;
  if (hoi_carve(workspace, a.bs, a.S, a.topk_hand, a.topk_obj, as.n_pts, &h) > workspace_bytes) return VPHO_ERR_INVALID;
  const ManoModelDev& m = mano_model_dev(mano);
  int nmax = 2 * a.S;
  if (kk > nmax) nmax = kk;
  if (nc > nmax) nmax = nc;
  cudaStream_t st = (cudaStream_t)stream;
  if (nmax > 1024) return VPHO_ERR_INVALID;
  profile_begin(VPHO_TAG_AGGREGATE, st);
  const int rc = nmax <= 256 ? run_hoi<8>(m, ah, h, st) : (nmax <= 512 ? run_hoi<16>(m, ah, h, st) : run_hoi<32>(m, ah, h, st));
  profile_end(VPHO_TAG_AGGREGATE, st);
  return rc;
}
