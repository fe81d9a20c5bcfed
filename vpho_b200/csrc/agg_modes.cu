// N4: the aggregation modes the predict branch does not use, as device primitives behind the C ABI.
//   HandAggregator.select_topk_hand_by_observed_heatmap_and_fuse_by_index   lib/model/aggregation.py:180-284  (any index sets)
//   HandAggregator.select_by_heatmap / select_by_heatmap_cascade_n_level    :82-113, 469-535  (compositions of the above)
//   HandAggregator.select_by_2D_pt ('2D_pt_pose' / '2D_pt_joint')           :286-377
//   HandAggregator.average_all / random                                     :379-467
//   ObjectAggregator.select_topk_object_by_heatmap / fuse_topk              :729-781  (select_by_heatmap :646-659 and the
//                                                                            non-physics branch of select_by_heatmap_cascade)
// The host mirror (vpho_b200/aggregation_modes.py) strings these together exactly as the reference's methods do; the MANO
// forward of the candidates is vpho_mano_forward.  Heat-maps are 64 x 64 (kHm) as everywhere in this library.
// Top-k is the library's (value descending, index ascending) order; sums run in index order like the cascade kernels'.
#include "vpho_common.cuh"
#include "vpho_b200.h"
#include "agg_device.cuh"
#include "rot_math.cuh"

namespace vpho {

// ---- per-joint scores of every candidate -------------------------------------------------------------------------------
// heat[b][c][j]   = grid_sample(heatmap[b][j], normalised projection of joint j of candidate c)      (:196-210)
// dist2d[b][c][j] = -|| projection - argmax position of heatmap[b][j] ||                             (:313-326)
__global__ void __launch_bounds__(256) k_heat_argmax(const float* __restrict__ heatmap, int n_maps, float* __restrict__ peak) {
  // torch.argmax over the flattened map (first maximum); the reference reads X from index // H and Y from index % H of a
  // meshgrid built with the default 'ij' indexing (:313-323), i.e. x comes from the ROW of the maximum
  const int mi = blockIdx.x;
  if (mi >= n_maps) return;
  const float* hm = heatmap + (size_t)mi * kHm * kHm;
  __shared__ float s_v[256];
  __shared__ int s_i[256];
  float best = hm[threadIdx.x];
  int bi = threadIdx.x;
  for (int i = threadIdx.x + 256; i < kHm * kHm; i += 256) {
    const float v = hm[i];
    if (v > best || (v != v && best == best)) { best = v; bi = i; }      // strictly greater keeps the first; NaN ranks highest
  }
  s_v[threadIdx.x] = best;
  s_i[threadIdx.x] = bi;
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) {
    if ((int)threadIdx.x < off) {
      const float a = s_v[threadIdx.x], b = s_v[threadIdx.x + off];
      const int ia = s_i[threadIdx.x], ib = s_i[threadIdx.x + off];
      const bool a_nan = a != a, b_nan = b != b;
      bool take_b;
      if (a_nan || b_nan) take_b = b_nan && (!a_nan || ib < ia);
      else take_b = b > a || (b == a && ib < ia);
      if (take_b) { s_v[threadIdx.x] = b; s_i[threadIdx.x] = ib; }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const int ind = s_i[0];
    peak[mi * 2 + 0] = (float)(ind / kHm) / (float)(kHm - 1) * 2.f - 1.f;
    peak[mi * 2 + 1] = (float)(ind % kHm) / (float)(kHm - 1) * 2.f - 1.f;
  }
}

__global__ void __launch_bounds__(256) k_joint_scores(vpho_joint_scores_args a, const float* __restrict__ peak) {
  const int total = a.bs * a.n * a.n_joints;
  for (int it = blockIdx.x * blockDim.x + threadIdx.x; it < total; it += gridDim.x * blockDim.x) {
    const int j = it % a.n_joints, bc = it / a.n_joints, b = bc / a.n;
    const float* p = a.joint + (size_t)it * 3;
    const float* root = a.root_joint + (size_t)b * 3;
    float gx, gy;
    project_to_grid(a.cam_intrinsic + (size_t)b * 9, a.bbox + (size_t)b * 4, p[0] + root[0], p[1] + root[1], p[2] + root[2], gx, gy);
    if (a.heat) a.heat[it] = bicubic_sample64(a.heatmap + ((size_t)b * a.n_joints + j) * kHm * kHm, gx, gy);
    if (a.dist2d) {
      const float dx = gx - peak[((size_t)b * a.n_joints + j) * 2 + 0], dy = gy - peak[((size_t)b * a.n_joints + j) * 2 + 1];
      a.dist2d[it] = -sqrtf(dx * dx + dy * dy);
    }
  }
}

// ---- one level: lists -> top-k -> weights -> fuse ------------------------------------------------------------------------
struct LevelDev {
  vpho_hand_level_args a;
  unsigned char obs[64];
  unsigned char fuse[192];     // pose parameters (fuse_kind 0: < 48) or joint coordinates (fuse_kind 1: < 3 * n_joints)
};

// One CTA per image, one warp per list (a single list when the level is not independent, one per fused joint otherwise).
template <int EL>
__global__ void __launch_bounds__(512) k_hand_level(LevelDev d) {
  const vpho_hand_level_args& a = d.a;
  VPHO_DYN_SMEM(float, s_dyn);                           // [n][lists] list scores, then [lists][64] values / indices
  const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int n = a.n, K = a.K, J = a.n_joints;
  const int nj = a.n_fuse / 3;                           // fused joints
  const int lists = a.independent ? nj : 1;
  const int n_obs = a.independent ? a.n_observe / nj : a.n_observe;
  // list scores: sum over the observed joints in index order (:212), or the mean over each fused joint's observations of
  // heat_val.reshape(bs, -1, n_obs, N/3) (:245)
  for (int it = threadIdx.x; it < n * lists; it += blockDim.x) {
    const int c = it / lists, l = it % lists;
    const float* sc = a.score + ((size_t)b * n + c) * J;
    float acc = 0.f;
    if (!a.independent) {
      for (int m = 0; m < n_obs; ++m) acc += sc[d.obs[m]];
    } else {
      for (int o = 0; o < n_obs; ++o) acc += sc[d.obs[o * nj + l]];
      acc = acc / (float)n_obs;
    }
    s_dyn[(size_t)c * lists + l] = acc;
  }
  __syncthreads();
  float* s_val = s_dyn + (size_t)n * lists;              // [lists][64]
  int* s_idx = reinterpret_cast<int*>(s_val + lists * 64);   // [lists][64]
  for (int l = warp; l < lists; l += nwarps) {
    auto value_of = [&](int i) { return s_dyn[(size_t)i * lists + l]; };
    warp_topk<EL>(n, K, value_of, s_val + l * 64, s_idx + l * 64, lane);
    for (int r = lane; r < K; r += 32) {
      if (a.val) a.val[((size_t)b * K + r) * lists + l] = s_val[l * 64 + r];
      if (a.topk) a.topk[((size_t)b * K + r) * lists + l] = s_idx[l * 64 + r];
    }
  }
  __syncthreads();
  if (a.fuse_kind == 1) {
    // 2D_pt_joint (:357-362): every joint is its own list; the fused joint is the mean of the winners' positions
    for (int it = threadIdx.x; it < lists * 3; it += blockDim.x) {
      const int l = it / 3, dd = it % 3;
      float acc = 0.f;
      for (int r = 0; r < K; ++r) acc += a.joint[(((size_t)b * n + s_idx[l * 64 + r]) * J + d.fuse[3 * l] / 3) * 3 + dd];
      a.fused[((size_t)b * lists + l) * 3 + dd] = acc / (float)K;
    }
    return;
  }
  // weighted quaternion average of every fused joint over its list's winners, one thread per joint (:218-231, 247-266)
  for (int g = threadIdx.x; g < nj; g += blockDim.x) {
    const int l = a.independent ? g : 0;
    const int jm = d.fuse[3 * g] / 3;                    // MANO joint whose 3 parameters are fused
    float vsum = 0.f;
    for (int r = 0; r < K; ++r) vsum += s_val[l * 64 + r];
    float A[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) A[i] = 0.f;
    float wsum = 0.f;
    for (int r = 0; r < K; ++r) {
      const float* pp = a.pose + ((size_t)b * n + s_idx[l * 64 + r]) * 48 + 3 * jm;
      float aa[3] = {pp[0], pp[1], pp[2]}, q[4];
      axis_angle_to_quaternion(aa, q);
      const float sg = q[0] > 0.f ? 1.f : -1.f;
      const float w = a.is_weight ? (s_val[l * 64 + r] + 1e-8f) / (vsum + 1e-8f) : 1.f;
      wsum += w;
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) A[i * 4 + j] += ((sg * q[i]) * (sg * q[j])) * w;
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) A[i] /= wsum;
    float qm[4], faa[3];
    sym4_top_eigvec(A, qm);
    quaternion_to_axis_angle(qm, faa);
#pragma unroll
    for (int dd = 0; dd < 3; ++dd) a.fused[(size_t)b * a.n_fuse + 3 * g + dd] = faa[dd];
  }
  if (!a.write_back) return;
  __syncthreads();
  // fused_pose[:, :, fuse_index] = fused_pose[:, :, fuse_index] * 0 + fused (:235-236): every candidate takes the fused value
  for (int it = threadIdx.x; it < n * a.n_fuse; it += blockDim.x) {
    const int c = it / a.n_fuse, f = it % a.n_fuse;
    float* pp = a.pose + ((size_t)b * n + c) * 48 + d.fuse[f];
    *pp = *pp * 0.f + a.fused[(size_t)b * a.n_fuse + f];
  }
}

// average_all (:400-404): unweighted quaternion average of every joint over ALL candidates, in candidate order
__global__ void __launch_bounds__(64) k_quat_average_all(const float* __restrict__ pose, int n, int n_joints, float* __restrict__ out) {
  const int b = blockIdx.x, j = threadIdx.x;
  if (j >= n_joints) return;
  float A[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) A[i] = 0.f;
  float wsum = 0.f;
  for (int c = 0; c < n; ++c) {
    const float* pp = pose + ((size_t)b * n + c) * 3 * n_joints + 3 * j;
    float aa[3] = {pp[0], pp[1], pp[2]}, q[4];
    axis_angle_to_quaternion(aa, q);
    const float sg = q[0] > 0.f ? 1.f : -1.f;
    wsum += 1.f;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int k = 0; k < 4; ++k) A[i * 4 + k] += ((sg * q[i]) * (sg * q[k])) * 1.f;
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) A[i] /= wsum;
  float qm[4], faa[3];
  sym4_top_eigvec(A, qm);
  quaternion_to_axis_angle(qm, faa);
#pragma unroll
  for (int dd = 0; dd < 3; ++dd) out[((size_t)b * n_joints + j) * 3 + dd] = faa[dd];
}

// ---- object: heat-map score of every candidate pose, top-k, fuse ---------------------------------------------------------
// score[b][c] = sum over the key-points of the bicubic heat value at the projected key-point (:752-776)
__global__ void __launch_bounds__(256) k_obj_mode_score(AssetsDev as, vpho_obj_select_args a, float* __restrict__ score,
                                                        const float* __restrict__ peak) {
  const int b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x * 8 + warp;
  if (c >= a.n) return;
  const int oid = obj_index(as, a.obj_id[b]);
  ObjPose o;
  make_obj_pose(a.pose6d + ((size_t)b * a.n + c) * 9, nullptr, a.root_joint + (size_t)b * 3, a.is_right[b] != 0, o);
  float hv = 0.f;
  if (lane < kKpts) {
    float p[3], gx, gy;
    obj_point(o, as.kpt + ((size_t)oid * kKpts + lane) * 3, p);
    project_to_grid(a.cam_intrinsic + (size_t)b * 9, a.bbox + (size_t)b * 4, p[0], p[1], p[2], gx, gy);
    if (a.score_kind == 0) {
      hv = bicubic_sample64(a.heatmap + ((size_t)b * kKpts + lane) * kHm * kHm, gx, gy);
    } else {
      const float dx = gx - peak[((size_t)b * kKpts + lane) * 2 + 0], dy = gy - peak[((size_t)b * kKpts + lane) * 2 + 1];
      hv = -sqrtf(dx * dx + dy * dy);
    }
  }
  // heatval.sum(dim=-1) in key-point order
  float acc = 0.f;
  for (int k = 0; k < kKpts; ++k) acc += __shfl_sync(0xffffffffu, hv, k);
  if (lane == 0) score[(size_t)b * a.n + c] = acc;
}

// top-k -> weights -> fuse_topk (:729-740, average_rot6d :50-56); float64 like the predict branch's object pose
template <int EL>
__global__ void __launch_bounds__(32) k_obj_mode_fuse(vpho_obj_select_args a, const float* __restrict__ score) {
  __shared__ float s_val[64];
  __shared__ int s_idx[64];
  const int b = blockIdx.x, lane = threadIdx.x;
  const int n = a.n, K = a.K;
  if (a.topk_in) {
    for (int r = lane; r < K; r += 32) { s_idx[r] = a.topk_in[(size_t)b * K + r]; s_val[r] = 0.f; }
    __syncwarp();
  } else {
    auto value_of = [&](int i) { return score[(size_t)b * n + i]; };
    warp_topk<EL>(n, K, value_of, s_val, s_idx, lane);
  }
  float vsum = 0.f;
  for (int r = 0; r < K; ++r) vsum += s_val[r];
  for (int r = lane; r < K; r += 32) {
    if (a.topk) a.topk[(size_t)b * K + r] = s_idx[r];
    if (a.weight) a.weight[(size_t)b * K + r] = (s_val[r] + 1e-8f) / (vsum + 1e-8f);
  }
  if (lane != 0 || !a.fused) return;
  // weights: the float32 heat weights when asked for, else ones / K (average_rot6d) and a plain mean of the translations
  double t[3] = {0.0, 0.0, 0.0};
  double A[16];
  for (int i = 0; i < 16; ++i) A[i] = 0.0;
  double wsum = 0.0;
  for (int r = 0; r < K; ++r) {
    const double* p = a.pose6d + ((size_t)b * n + s_idx[r]) * 9;
    const double w = a.is_weight ? (double)((s_val[r] + 1e-8f) / (vsum + 1e-8f)) : 1.0 / (double)K;
    for (int dd = 0; dd < 3; ++dd) t[dd] += a.is_weight ? p[6 + dd] * w : p[6 + dd];
    double R[9], q[4];
    rot6d_to_matrix(p, R);
    matrix_to_quaternion(R, q);
    const double sg = q[0] > 0.0 ? 1.0 : -1.0;
    wsum += w;
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) A[i * 4 + j] += ((sg * q[i]) * (sg * q[j])) * w;
  }
  for (int i = 0; i < 16; ++i) A[i] /= wsum;
  double qm[4], R[9];
  sym4_top_eigvec(A, qm);
  const double sg = qm[0] > 0.0 ? 1.0 : -1.0;
  for (int i = 0; i < 4; ++i) qm[i] *= sg;
  quaternion_to_matrix(qm, R);
  double* out = a.fused + (size_t)b * 9;
  // matrix_to_rotation_6d: the first two ROWS
  for (int i = 0; i < 6; ++i) out[i] = R[i];
  for (int dd = 0; dd < 3; ++dd) out[6 + dd] = a.is_weight ? t[dd] : t[dd] / (double)K;
}

}  // namespace vpho

using namespace vpho;

extern "C" int vpho_joint_scores(const vpho_joint_scores_args* args, void* workspace, size_t workspace_bytes, void* stream) {
  if (!args) return VPHO_ERR_INVALID;
  const vpho_joint_scores_args& a = *args;
  if (a.bs < 0 || a.n <= 0 || a.n_joints <= 0 || a.n_joints > 64) return VPHO_ERR_INVALID;
  if (a.bs == 0) return VPHO_OK;
  if (!a.joint || !a.root_joint || !a.cam_intrinsic || !a.bbox || !a.heatmap || (!a.heat && !a.dist2d)) return VPHO_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  float* peak = nullptr;
  if (a.dist2d) {
    if (!workspace || workspace_bytes < (size_t)a.bs * a.n_joints * 2 * sizeof(float)) return VPHO_ERR_INVALID;
    peak = static_cast<float*>(workspace);
    VPHO_LAUNCH(k_heat_argmax, dim3(a.bs * a.n_joints), dim3(256), 0, st, a.heatmap, a.bs * a.n_joints, peak);
  }
  const int total = a.bs * a.n * a.n_joints;
  int blocks = (total + 255) / 256;
  if (blocks > 4096) blocks = 4096;
  VPHO_LAUNCH(k_joint_scores, dim3(blocks), dim3(256), 0, st, a, peak);
  VPHO_CHECK_LAUNCH();
  return VPHO_OK;
}

extern "C" int vpho_hand_level(const vpho_hand_level_args* args, void* stream) {
  if (!args) return VPHO_ERR_INVALID;
  const vpho_hand_level_args& a = *args;
  if (a.bs < 0 || a.n <= 0 || a.n > 1024 || a.K < 1 || a.K > 64 || a.K > a.n || a.n_joints <= 0 || a.n_joints > 64) return VPHO_ERR_INVALID;
  if (a.bs == 0) return VPHO_OK;
  if (!a.score || !a.fused || !a.observe_index || !a.fuse_index) return VPHO_ERR_INVALID;
  if (a.n_observe < 1 || a.n_observe > 64 || a.n_fuse < 3 || a.n_fuse > (a.fuse_kind == 0 ? 48 : 3 * a.n_joints) || a.n_fuse % 3 != 0)
    return VPHO_ERR_INVALID;
  const int nj = a.n_fuse / 3;
  if (a.independent && a.n_observe % nj != 0) return VPHO_ERR_INVALID;           // the reference asserts M % (N // 3) == 0
  if (a.fuse_kind != 0 && a.fuse_kind != 1) return VPHO_ERR_INVALID;
  if (a.fuse_kind == 0 && !a.pose) return VPHO_ERR_INVALID;
  if (a.fuse_kind == 1 && (!a.joint || !a.independent)) return VPHO_ERR_INVALID;
  LevelDev d;
  d.a = a;
  for (int i = 0; i < a.n_observe; ++i) {
    if (a.observe_index[i] < 0 || a.observe_index[i] >= a.n_joints) return VPHO_ERR_INVALID;
    d.obs[i] = (unsigned char)a.observe_index[i];
  }
  for (int i = 0; i < a.n_fuse; ++i) {
    if (a.fuse_index[i] < 0 || a.fuse_index[i] >= (a.fuse_kind == 0 ? 48 : 3 * a.n_joints)) return VPHO_ERR_INVALID;
    // whole joints only: (3j, 3j+1, 3j+2)
    if ((i % 3 == 0 && a.fuse_index[i] % 3 != 0) || (i % 3 != 0 && a.fuse_index[i] != a.fuse_index[i - 1] + 1)) return VPHO_ERR_INVALID;
    d.fuse[i] = (unsigned char)a.fuse_index[i];
  }
  d.a.observe_index = nullptr;
  d.a.fuse_index = nullptr;
  const int lists = a.independent ? nj : 1;
  const size_t smem = ((size_t)a.n * lists + (size_t)lists * 64 * 2) * sizeof(float);
  if (smem > 200 * 1024) return VPHO_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
#ifndef VPHO_EMU
#define VPHO_LEVEL_ATTR(EL)                                                                                              \
  if (smem > 48 * 1024 &&                                                                                                \
      cudaFuncSetAttribute(k_hand_level<EL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)     \
    return VPHO_ERR_LAUNCH;
#else
#define VPHO_LEVEL_ATTR(EL)
#endif
#define VPHO_LEVEL_LAUNCH(EL)                                                                                            \
  do {                                                                                                                   \
    VPHO_LEVEL_ATTR(EL)                                                                                                  \
    VPHO_LAUNCH(k_hand_level<EL>, dim3(a.bs), dim3(512), smem, st, d);                                                   \
  } while (0)
  if (a.n <= 256) VPHO_LEVEL_LAUNCH(8);
  else if (a.n <= 512) VPHO_LEVEL_LAUNCH(16);
  else VPHO_LEVEL_LAUNCH(32);
#undef VPHO_LEVEL_LAUNCH
#undef VPHO_LEVEL_ATTR
  VPHO_CHECK_LAUNCH();
  return VPHO_OK;
}

extern "C" int vpho_quat_average_all(const float* pose, int bs, int n, int n_joints, float* out, void* stream) {
  if (bs < 0 || n <= 0 || n_joints <= 0 || n_joints > 64) return VPHO_ERR_INVALID;
  if (bs == 0) return VPHO_OK;
  if (!pose || !out) return VPHO_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  VPHO_LAUNCH(k_quat_average_all, dim3(bs), dim3(64), 0, st, pose, n, n_joints, out);
  VPHO_CHECK_LAUNCH();
  return VPHO_OK;
}

extern "C" int vpho_obj_select(vpho_assets_t assets, const vpho_obj_select_args* args, void* workspace, size_t workspace_bytes,
                               void* stream) {
  if (!assets || !args) return VPHO_ERR_INVALID;
  const vpho_obj_select_args& a = *args;
  if (a.bs < 0 || a.n <= 0 || a.n > 1024 || a.K < 1 || a.K > 64 || a.K > a.n) return VPHO_ERR_INVALID;
  if (a.bs == 0) return VPHO_OK;
  if (!a.pose6d || !a.root_joint || !a.cam_intrinsic || !a.bbox || !a.heatmap || !a.is_right || !a.obj_id) return VPHO_ERR_INVALID;
  if (a.score_kind != 0 && a.score_kind != 1) return VPHO_ERR_INVALID;
  const size_t need = ((size_t)a.bs * a.n + (a.score_kind == 1 ? (size_t)a.bs * kKpts * 2 : 0)) * sizeof(float);
  if (!workspace || workspace_bytes < need) return VPHO_ERR_INVALID;
  const AssetsDev& as = static_cast<AssetsHost*>(assets)->dev;
  float* score = static_cast<float*>(workspace);
  float* peak = score + (size_t)a.bs * a.n;
  cudaStream_t st = (cudaStream_t)stream;
  if (a.is_weight && (a.topk_in || a.score_kind == 1)) return VPHO_ERR_INVALID;       // no heat values to weight with
  if (!a.topk_in) {
    if (a.score_kind == 1) VPHO_LAUNCH(k_heat_argmax, dim3(a.bs * kKpts), dim3(256), 0, st, a.heatmap, a.bs * kKpts, peak);
    VPHO_LAUNCH(k_obj_mode_score, dim3((a.n + 7) / 8, a.bs), dim3(256), 0, st, as, a, score, peak);
  }
  if (a.n <= 256) VPHO_LAUNCH(k_obj_mode_fuse<8>, dim3(a.bs), dim3(32), 0, st, a, score);
  else if (a.n <= 512) VPHO_LAUNCH(k_obj_mode_fuse<16>, dim3(a.bs), dim3(32), 0, st, a, score);
  else VPHO_LAUNCH(k_obj_mode_fuse<32>, dim3(a.bs), dim3(32), 0, st, a, score);
  VPHO_CHECK_LAUNCH();
  return VPHO_OK;
}
