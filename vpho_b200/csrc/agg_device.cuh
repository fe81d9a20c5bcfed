// Device building blocks of the visual / physical scoring and top-k aggregation stage.
//
// Reference: lib/model/aggregation.py (HOI_Aggregator.__call__ :1167-1353 and the helpers it reaches),
// lib/utils/physics_fn.py:224-257 (VERT2ANCHOR), lib/utils/hand_fn.py:240-274,427-448, lib/model/physics.py:362-371,
// lib/model/head_object.py:36-67.  torch.nn.functional.grid_sample(bicubic, align_corners=False, zeros padding)
// and torch.topk are restated.
#pragma once
#include "mano_device.cuh"

namespace vpho {

constexpr int kAnchors = 32;
constexpr int kKpts = 27;
constexpr int kHm = 64;   // heat-map side

struct AssetsDev {
  const int* face;      // [32][3] vertex ids of each force anchor's triangle
  const float* aw;      // [32][2] barycentric-like weights (w1, w2)
  const float* v2j;     // [21][778] dense vertex->joint regressor (asset/ours/vert2joint.pkl)
  int n_obj, n_pts;
  const float* kpt;     // [n_obj][27][3]
  const float* verts;   // [n_obj][n_pts][3]
  const float* com;     // [n_obj][3]
};

// An object id outside [0, n_obj) (a caller bug the host mirror rejects for host-side ids) must never index the tables out
// of bounds: device-resident ids are clamped here.
__device__ __forceinline__ int obj_index(const AssetsDev& as, int id) { return id < 0 ? 0 : (id >= as.n_obj ? as.n_obj - 1 : id); }

// host-side handle behind vpho_assets_t
struct AssetsHost {
  AssetsDev dev;
  void* blob = nullptr;
#ifndef VPHO_EMU
  // Side stream + fork/join events of the aggregation (object branch beside the hand cascade): owned by the handle, created
  // on the device that was current in vpho_assets_create.  One aggregation in flight per handle.
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  int device = -1;
#endif
};

// anchor -> (joint a, joint b) bone used for the frame's y axis (physics_fn.py:127-169 after argsort(label))
__device__ __forceinline__ void anchor_bone(int j, int& ja, int& jb) {
  const unsigned char t[32][2] = {{0, 1},   {2, 3},   {3, 4},   {3, 4},   {3, 4},   {0, 5},   {0, 1},   {5, 6},
                                  {6, 7},   {7, 8},   {7, 8},   {7, 8},   {0, 9},   {9, 10},  {10, 11}, {11, 12},
                                  {11, 12}, {11, 12}, {0, 13},  {0, 13},  {13, 14}, {14, 15}, {15, 16}, {15, 16},
                                  {15, 16}, {0, 17},  {0, 17},  {17, 18}, {18, 19}, {19, 20}, {19, 20}, {19, 20}};
  ja = t[j][0];
  jb = t[j][1];
}

// anchors summed per finger in the hand physics score (aggregation.py:584-590): thumb,index,middle,ring,pinky
__device__ __forceinline__ int finger_anchor(int f, int i) {
  const unsigned char t[5][4] = {{1, 2, 3, 4}, {8, 9, 10, 11}, {14, 15, 16, 17}, {21, 22, 23, 24}, {28, 29, 30, 31}};
  return t[f][i];
}

// MANO_PARAMS_LEVEL (hand_fn.py:240-247): finger order thumb,index,middle,ring,pinky -> manopth kinematic joint
__device__ __forceinline__ int finger_mano_joint(int f, int level /*1..3*/) {
  const int base[5] = {13, 1, 4, 10, 7};
  return base[f] + (level - 1);
}
// cascade level that owns MANO parameter p (0..47)
__device__ __forceinline__ int param_level(int p) {
  const int jm = p / 3;
  return jm == 0 ? 0 : ((jm - 1) % 3) + 1;
}
// MANO_JOINT_LEVEL (hand_fn.py:250-263): joint of finger f at level l (1..4) in the 21-joint order
__device__ __forceinline__ int finger_joint21(int f, int l) { return 1 + 4 * f + (l - 1); }

// ---- projection + box normalisation (aggregation.py:24-32, 201-204, 758-762) ----
__device__ __forceinline__ void project_to_grid(const float* K, const float* bbox, float x, float y, float z, float& gx,
                                                float& gy) {
  const float u = (x * K[0] + y * K[1]) + z * K[2];
  const float v = (x * K[3] + y * K[4]) + z * K[5];
  const float w = (x * K[6] + y * K[7]) + z * K[8];
  float px = u / w, py = v / w;
  px = px - bbox[0];
  py = py - bbox[1];
  gx = 2.f * px / (bbox[2] - bbox[0]) - 1.f;
  gy = 2.f * py / (bbox[3] - bbox[1]) - 1.f;
}

// ---- bicubic grid_sample of one 64x64 map at one point (Appendix A.3) ----
__device__ __forceinline__ float cubic_cc1(float x) { return ((1.25f * x - 2.25f) * x) * x + 1.f; }                     // A = -0.75
__device__ __forceinline__ float cubic_cc2(float x) { return ((-0.75f * x + 3.75f) * x - 6.f) * x + 3.f; }

__device__ __forceinline__ float bicubic_sample64(const float* __restrict__ hm, float gx, float gy) {
  const float ix = ((gx + 1.f) * (float)kHm - 1.f) / 2.f;
  const float iy = ((gy + 1.f) * (float)kHm - 1.f) / 2.f;
  if (!(ix > -3.f && ix < (float)kHm + 2.f && iy > -3.f && iy < (float)kHm + 2.f)) return 0.f;   // every tap is padding
  const float fx = floorf(ix), fy = floorf(iy);
  const float tx = ix - fx, ty = iy - fy;
  const int x0 = (int)fx, y0 = (int)fy;
  float cx[4], cy[4];
  cx[0] = cubic_cc2(tx + 1.f); cx[1] = cubic_cc1(tx); cx[2] = cubic_cc1(1.f - tx); cx[3] = cubic_cc2((1.f - tx) + 1.f);
  cy[0] = cubic_cc2(ty + 1.f); cy[1] = cubic_cc1(ty); cy[2] = cubic_cc1(1.f - ty); cy[3] = cubic_cc2((1.f - ty) + 1.f);
  float rows[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int yy = y0 - 1 + i;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int xx = x0 - 1 + j;
      v[j] = (xx >= 0 && xx < kHm && yy >= 0 && yy < kHm) ? __ldg(hm + yy * kHm + xx) : 0.f;
    }
    rows[i] = ((v[0] * cx[0] + v[1] * cx[1]) + v[2] * cx[2]) + v[3] * cx[3];
  }
  return ((rows[0] * cy[0] + rows[1] * cy[1]) + rows[2] * cy[2]) + rows[3] * cy[3];
}

// ---- object points (head_object.py:36-67): R(rot6d) p + t, x negated for left hands ----
struct ObjPose {
  float R[9];
  float t[3];
  bool flip;
};
__device__ __forceinline__ void make_obj_pose(const double* pose6d, const double* transl_override, const float* root,
                                              bool is_right, ObjPose& o) {
  float d[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) d[k] = (float)pose6d[k];
  rot6d_to_matrix(d, o.R);
#pragma unroll
  for (int k = 0; k < 3; ++k) o.t[k] = (float)(transl_override ? transl_override[k] : pose6d[6 + k]) + root[k];
  o.flip = !is_right;
}
__device__ __forceinline__ void obj_point(const ObjPose& o, const float* p, float* out) {
#pragma unroll
  for (int j = 0; j < 3; ++j) out[j] = ((p[0] * o.R[j * 3 + 0] + p[1] * o.R[j * 3 + 1]) + p[2] * o.R[j * 3 + 2]) + o.t[j];
  if (o.flip) out[0] = -out[0];
}

// ---- warp-level bitonic sort, descending by (value, then lower index first) ----
__device__ __forceinline__ unsigned long long topk_key(float v, int idx) {
  v = v + 0.f;                                   // -0 -> +0
  unsigned u = __float_as_uint(v);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  if (v != v) u = 0xFFFFFFFFu;                   // torch.topk ranks NaN as the largest value
  return ((unsigned long long)u << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)idx);
}
__device__ __forceinline__ int topk_key_index(unsigned long long k) { return (int)(0xFFFFFFFFu - (unsigned)(k & 0xFFFFFFFFull)); }

// EL keys per lane, element e = lane*EL + r; after the call element e holds the e-th largest key
template <int EL>
__device__ __forceinline__ void warp_bitonic_sort_desc(unsigned long long (&key)[EL], int lane) {
#pragma unroll
  for (int k = 2; k <= 32 * EL; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j >= EL) {
#pragma unroll
        for (int r = 0; r < EL; ++r) {
          const unsigned long long other = __shfl_xor_sync(0xffffffffu, key[r], j / EL);
          const int e = lane * EL + r;
          const bool up = ((e & k) == 0), lower = ((e & j) == 0);
          const bool keep_max = (lower == up);
          const unsigned long long mx = key[r] > other ? key[r] : other, mn = key[r] > other ? other : key[r];
          key[r] = keep_max ? mx : mn;
        }
      } else {
#pragma unroll
        for (int r = 0; r < EL; ++r) {
          if ((r & j) == 0) {
            const int r2 = r | j;
            const int e = lane * EL + r;
            const bool up = ((e & k) == 0);
            const unsigned long long a = key[r], b = key[r2];
            const unsigned long long mx = a > b ? a : b, mn = a > b ? b : a;
            key[r] = up ? mx : mn;
            key[r2] = up ? mn : mx;
          }
        }
      }
    }
  }
}

// One warp selects the K largest of n values (value desc, index asc).  `value_of(i)` is evaluated by the lane that
// owns element i.  Results go to out_val[0..K), out_idx[0..K) (shared or global memory); n <= 32*EL.
template <int EL, typename ValFn>
__device__ __forceinline__ void warp_topk(int n, int K, ValFn value_of, float* out_val, int* out_idx, int lane) {
  unsigned long long key[EL];
  float val[EL];
#pragma unroll
  for (int r = 0; r < EL; ++r) {
    const int i = lane * EL + r;
    val[r] = 0.f;
    if (i < n) { val[r] = value_of(i); key[r] = topk_key(val[r], i); }
    else key[r] = 0ull;
  }
  warp_bitonic_sort_desc<EL>(key, lane);
#pragma unroll
  for (int r = 0; r < EL; ++r) {
    const int e = lane * EL + r;
    if (e < K && e < n) {
      const int idx = topk_key_index(key[r]);
      out_idx[e] = idx;
    }
  }
  __syncwarp();
  // values are re-read through value_of so that NaN / -0 survive unchanged
  for (int e = lane; e < K && e < n; e += 32) out_val[e] = value_of(out_idx[e]);
  __syncwarp();
}

// ---- force anchors of one posed hand (VERT2ANCHOR + Vert2Joint + from_local_to_global) ----
// vert(v, out[3]) returns the camera-frame vertex; j21 [21][3] are the vert2joint-regressed joints of the same hand.
template <typename VertFn>
__device__ __forceinline__ void anchor_point_and_force(const AssetsDev& as, int j, VertFn vert, const float* j21,
                                                       const float* force_local, float* point, float* force_global) {
  float v0[3], v1[3], v2[3];
  vert(as.face[j * 3 + 0], v0);
  vert(as.face[j * 3 + 1], v1);
  vert(as.face[j * 3 + 2], v2);
  float b1[3], b2[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) { b1[d] = v1[d] - v0[d]; b2[d] = v2[d] - v0[d]; }
  int ja, jb;
  anchor_bone(j, ja, jb);
  float dy[3] = {j21[jb * 3 + 0] - j21[ja * 3 + 0], j21[jb * 3 + 1] - j21[ja * 3 + 1], j21[jb * 3 + 2] - j21[ja * 3 + 2]};
  float dz[3] = {b1[1] * b2[2] - b1[2] * b2[1], b1[2] * b2[0] - b1[0] * b2[2], b1[0] * b2[1] - b1[1] * b2[0]};
  float nz = sqrtf((dz[0] * dz[0] + dz[1] * dz[1]) + dz[2] * dz[2]) + 1e-8f;
  float ny = sqrtf((dy[0] * dy[0] + dy[1] * dy[1]) + dy[2] * dy[2]) + 1e-8f;
#pragma unroll
  for (int d = 0; d < 3; ++d) { dz[d] = dz[d] / nz; dy[d] = dy[d] / ny; }
  float dx[3] = {dy[1] * dz[2] - dy[2] * dz[1], dy[2] * dz[0] - dy[0] * dz[2], dy[0] * dz[1] - dy[1] * dz[0]};
  float dy2[3] = {dz[1] * dx[2] - dz[2] * dx[1], dz[2] * dx[0] - dz[0] * dx[2], dz[0] * dx[1] - dz[1] * dx[0]};
  ny = sqrtf((dy2[0] * dy2[0] + dy2[1] * dy2[1]) + dy2[2] * dy2[2]) + 1e-8f;
#pragma unroll
  for (int d = 0; d < 3; ++d) dy2[d] = dy2[d] / ny;
  const float w1 = as.aw[j * 2 + 0], w2 = as.aw[j * 2 + 1];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    point[d] = (w1 * b1[d] + w2 * b2[d]) + v0[d];
    // force_global = frame . force_local, frame columns (dx, dy, dz)   (physics.py:368)
    force_global[d] = (force_local[0] * dx[d] + force_local[1] * dy2[d]) + force_local[2] * dz[d];
  }
}

}  // namespace vpho
