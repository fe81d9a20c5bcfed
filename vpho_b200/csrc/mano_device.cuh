// Device-side MANO building blocks shared by the full skinning kernel (mano.cu) and the joints-only
// scorer of the heat-map cascade (aggregate.cu).
//
// Arithmetic restated from manopth `ManoLayer.forward` (un-vendored third party behind
// lib/model/head_mano.py:78-87; constants in SURVEY.md §3.2 / A.9):
//   Rodrigues(16) -> pose_map = R[1:] - I (135) -> v_posed = template + shapedirs.beta + posedirs.pose_map
//   J = J_regressor.(template + shapedirs.beta) -> root + 3 chained levels of 5 joints -> A_j = G_j - [0 | G_j J_j]
//   v = (sum_j w_vj A_j) [v_posed; 1] -> tips [745,317,444,556,673] -> 21 joints -> wrist-centred -> *1000 (then /1000
//   in HeadMano.get_hand_verts).
#pragma once
#include "rot_math.cuh"

namespace vpho {

struct ManoModelDev {
  const float* dirs;          // [145][3][kVPad]   k<10: shapedirs, k>=10: posedirs; chunk-padded vertex slots
  const float* v_template;    // [3][kVPad]
  const float* weights;       // [16][kVPad]       skinning weights, joint-major
  const float* J_template;    // [16][3]           J_regressor . v_template      (packed in f64, rounded once)
  const float* J_shapedirs;   // [16][3][10]       J_regressor . shapedirs
  const float* tip_dirs;      // [5][3][145]       blend rows of the 5 fingertip vertices
  const float* tip_template;  // [5][3]
  const float* tip_weights;   // [5][16]
  const void* tc_host;        // HOST pointer (opaque on the device): tables + TMA descriptors of the tensor-core kernel (mano_tc.cu)
};

// kinematic joint j (manopth order) -> slot in the 21-joint output; fingertip t -> slot
__device__ __forceinline__ int joint16_to_21(int j) {
  const int map[16] = {0, 5, 6, 7, 9, 10, 11, 17, 18, 19, 13, 14, 15, 1, 2, 3};
  return map[j];
}
__device__ __forceinline__ int tip_to_21(int t) { return 4 + 4 * t; }          // thumb,index,middle,ring,pinky
__device__ __forceinline__ int tip_vertex(int t) {
  const int tv[5] = {745, 317, 444, 556, 673};
  return tv[t];
}

template <int TC>
struct ManoSmem {
  float coefT[kBlendK][TC];   // blend coefficients, transposed so one LDS.128 feeds 4 candidates
  float J[TC][16][3];
  float R[TC][16][9];
  float G[TC][16][12];        // global 3x4 transform of each kinematic joint
  float A[TC][16][12];        // skinning transform  G - [0 | G.J]
};

// 3x4 (row-major [r][4]) composed with a relative [R | d] exactly in the order of the 4x4 matmul the reference
// performs: out[r][c] = sum_k P[r][k] rel[k][c], translation column gets + P[r][3]*1 last.
__device__ __forceinline__ void compose34(const float* P, const float* Rrel, const float* d, float* out) {
#pragma unroll
  for (int r = 0; r < 3; ++r) {
#pragma unroll
    for (int c = 0; c < 3; ++c)
      out[r * 4 + c] = (P[r * 4 + 0] * Rrel[0 * 3 + c] + P[r * 4 + 1] * Rrel[1 * 3 + c]) + P[r * 4 + 2] * Rrel[2 * 3 + c];
    out[r * 4 + 3] = ((P[r * 4 + 0] * d[0] + P[r * 4 + 1] * d[1]) + P[r * 4 + 2] * d[2]) + P[r * 4 + 3];
  }
}

// Block-cooperative pose set-up for TC candidates.  `pose_of(c, j, a)` writes the axis-angle of kinematic joint j of
// local candidate c into a[3]; `shape_of(c, beta)` writes its 10 betas; both return false for a padding slot.
template <int TC, typename PoseFn, typename ShapeFn>
__device__ __forceinline__ void mano_pose_setup(const ManoModelDev& m, PoseFn pose_of, ShapeFn shape_of,
                                                ManoSmem<TC>& s) {
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int it = tid; it < TC * 16; it += nt) {
    const int c = it >> 4, j = it & 15;
    float R[9], a[3], beta[10];
    const bool valid = pose_of(c, j, a);
    if (valid) {
      manopth_rodrigues(a, R);
    } else {
#pragma unroll
      for (int e = 0; e < 9; ++e) R[e] = (e % 4 == 0) ? 1.f : 0.f;
    }
#pragma unroll
    for (int e = 0; e < 9; ++e) s.R[c][j][e] = R[e];
    if (j >= 1) {
#pragma unroll
      for (int e = 0; e < 9; ++e) s.coefT[10 + (j - 1) * 9 + e][c] = R[e] - ((e % 4 == 0) ? 1.f : 0.f);
    }
    if (!shape_of(c, beta)) {
#pragma unroll
      for (int k = 0; k < 10; ++k) beta[k] = 0.f;
    }
    if (j == 0) {
#pragma unroll
      for (int k = 0; k < 10; ++k) s.coefT[k][c] = beta[k];
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < 10; ++k) acc = fmaf(m.J_shapedirs[(j * 3 + d) * 10 + k], beta[k], acc);
      s.J[c][j][d] = acc + m.J_template[j * 3 + d];
    }
  }
  __syncthreads();
  // kinematic chain: one thread per (candidate, finger); root handled by finger 0
  for (int it = tid; it < TC * 5; it += nt) {
    const int c = it / 5, f = it % 5;
    float G0[12];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) G0[r * 4 + cc] = s.R[c][0][r * 3 + cc];
      G0[r * 4 + 3] = s.J[c][0][r];
    }
    if (f == 0) {
#pragma unroll
      for (int e = 0; e < 12; ++e) s.G[c][0][e] = G0[e];
    }
    float Gp[12];
#pragma unroll
    for (int e = 0; e < 12; ++e) Gp[e] = G0[e];
    int jp = 0;
#pragma unroll
    for (int lv = 0; lv < 3; ++lv) {
      const int j = 1 + 3 * f + lv;
      float d[3] = {s.J[c][j][0] - s.J[c][jp][0], s.J[c][j][1] - s.J[c][jp][1], s.J[c][j][2] - s.J[c][jp][2]};
      float Gn[12];
      compose34(Gp, s.R[c][j], d, Gn);
#pragma unroll
      for (int e = 0; e < 12; ++e) { s.G[c][j][e] = Gn[e]; Gp[e] = Gn[e]; }
      jp = j;
    }
  }
  __syncthreads();
  for (int it = tid; it < TC * 16; it += nt) {
    const int c = it >> 4, j = it & 15;
    const float* G = s.G[c][j];
    const float* Jj = s.J[c][j];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      float gj = (G[r * 4 + 0] * Jj[0] + G[r * 4 + 1] * Jj[1]) + G[r * 4 + 2] * Jj[2];
      s.A[c][j][r * 4 + 0] = G[r * 4 + 0];
      s.A[c][j][r * 4 + 1] = G[r * 4 + 1];
      s.A[c][j][r * 4 + 2] = G[r * 4 + 2];
      s.A[c][j][r * 4 + 3] = G[r * 4 + 3] - gj;
    }
  }
  __syncthreads();
}

// skin one rest-pose point with its 16 weights against candidate c's transforms
template <int TC>
__device__ __forceinline__ void mano_skin_point(const ManoSmem<TC>& s, int c, const float* w, const float* vp,
                                                float* out) {
  float T[12];
#pragma unroll
  for (int e = 0; e < 12; ++e) T[e] = 0.f;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float4* a4 = reinterpret_cast<const float4*>(s.A[c][j]);
    float4 a0 = a4[0], a1 = a4[1], a2 = a4[2];
    const float wj = w[j];
    T[0] = fmaf(wj, a0.x, T[0]); T[1] = fmaf(wj, a0.y, T[1]); T[2] = fmaf(wj, a0.z, T[2]); T[3] = fmaf(wj, a0.w, T[3]);
    T[4] = fmaf(wj, a1.x, T[4]); T[5] = fmaf(wj, a1.y, T[5]); T[6] = fmaf(wj, a1.z, T[6]); T[7] = fmaf(wj, a1.w, T[7]);
    T[8] = fmaf(wj, a2.x, T[8]); T[9] = fmaf(wj, a2.y, T[9]); T[10] = fmaf(wj, a2.z, T[10]); T[11] = fmaf(wj, a2.w, T[11]);
  }
#pragma unroll
  for (int r = 0; r < 3; ++r)
    out[r] = ((T[r * 4 + 0] * vp[0] + T[r * 4 + 1] * vp[1]) + T[r * 4 + 2] * vp[2]) + T[r * 4 + 3];
}

// wrist-centre, then the reference's mm round trip (manopth *1000, head_mano.py:85-86 /1000)
__device__ __forceinline__ float mano_center_scale(float x, float center) {
  return __fdiv_rn(__fmul_rn(__fsub_rn(x, center), 1000.f), 1000.f);
}

}  // namespace vpho
