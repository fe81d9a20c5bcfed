// Batched MANO layer (blend shapes + pose correctives + linear-blend skinning) for sm_100a.
// Replaces `HeadMano.get_hand_verts` -> manopth `ManoLayer.forward` (lib/model/head_mano.py:78-87).
//
// One CTA = TC candidates x one 195-vertex chunk.  The 145-coefficient blend is a register-tiled FP32
// contraction (each blend direction is loaded once per CTA, coalesced, and reused for TC candidates; the
// per-candidate coefficients are broadcast from shared memory four at a time), followed by the 16-joint skinning
// sum and the wrist-centring.  Roofline: HBM-bound on the 9 588 B/candidate it writes when verts are materialised.
#include "mano_device.cuh"
#include "vpho_b200.h"

#include <vector>

namespace vpho {

struct ManoModelHost {
  ManoModelDev dev;
  float* blob = nullptr;
};

#ifndef VPHO_EMU
// tensor-core layer (mano_tc.cu)
int mano_tc_create(const float* v_template, const float* shapedirs, const float* posedirs, const float* weights, void** out);
void mano_tc_destroy(void* h);
int mano_tc_forward(const void* h, const ManoModelDev& m, const float* pose, const float* shape, int pose_stride, int shape_stride,
                    int n, float* verts, float* joints, cudaStream_t stream, int debug_blend);
#endif

template <int TC>
__global__ void __launch_bounds__(kVChunkPad) mano_forward_kernel(ManoModelDev m, const float* __restrict__ pose,
                                                                  const float* __restrict__ shape, int pose_stride,
                                                                  int shape_stride, int n, float* __restrict__ verts,
                                                                  float* __restrict__ joints) {
  pdl_wait();          // launched with VPHO_LAUNCH_PDL
  pdl_trigger();
  VPHO_DYN_SMEM(ManoSmem<TC>, sp);
  ManoSmem<TC>& s = *sp;
  const int c0 = blockIdx.x * TC;
  const int chunk = blockIdx.y;
  const int tid = threadIdx.x;
  mano_pose_setup<TC>(
      m,
      [&](int c, int j, float* a) {
        if (c0 + c >= n) return false;
        const float* pp = pose + (size_t)(c0 + c) * pose_stride + 3 * j;
        a[0] = pp[0]; a[1] = pp[1]; a[2] = pp[2];
        return true;
      },
      [&](int c, float* beta) {
        if (c0 + c >= n) return false;
        const float* sp = shape + (size_t)(c0 + c) * shape_stride;
#pragma unroll
        for (int k = 0; k < 10; ++k) beta[k] = sp[k];
        return true;
      },
      s);

  if (chunk == 0) {
    for (int it = tid; it < TC * 16; it += blockDim.x) {
      const int c = it >> 4, j = it & 15;
      if (c0 + c >= n) continue;
      float* o = joints + ((size_t)(c0 + c) * 21 + joint16_to_21(j)) * 3;
#pragma unroll
      for (int d = 0; d < 3; ++d) o[d] = mano_center_scale(s.G[c][j][d * 4 + 3], s.G[c][0][d * 4 + 3]);
    }
  }
  if (tid >= kVChunk) return;
  const int v = chunk * kVChunk + tid;
  if (v >= kVerts) return;
  const int pv = chunk * kVChunkPad + tid;

  float acc[TC][3];
  {
    const float t0 = m.v_template[0 * kVPad + pv], t1 = m.v_template[1 * kVPad + pv], t2 = m.v_template[2 * kVPad + pv];
#pragma unroll
    for (int c = 0; c < TC; ++c) { acc[c][0] = t0; acc[c][1] = t1; acc[c][2] = t2; }
  }
#pragma unroll 5
  for (int k = 0; k < kBlendK; ++k) {
    const float d0 = __ldg(m.dirs + (size_t)(k * 3 + 0) * kVPad + pv);
    const float d1 = __ldg(m.dirs + (size_t)(k * 3 + 1) * kVPad + pv);
    const float d2 = __ldg(m.dirs + (size_t)(k * 3 + 2) * kVPad + pv);
    const float4* cf4 = reinterpret_cast<const float4*>(s.coefT[k]);
#pragma unroll
    for (int q = 0; q < TC / 4; ++q) {
      const float4 cf = cf4[q];
      acc[4 * q + 0][0] = fmaf(d0, cf.x, acc[4 * q + 0][0]); acc[4 * q + 0][1] = fmaf(d1, cf.x, acc[4 * q + 0][1]); acc[4 * q + 0][2] = fmaf(d2, cf.x, acc[4 * q + 0][2]);
      acc[4 * q + 1][0] = fmaf(d0, cf.y, acc[4 * q + 1][0]); acc[4 * q + 1][1] = fmaf(d1, cf.y, acc[4 * q + 1][1]); acc[4 * q + 1][2] = fmaf(d2, cf.y, acc[4 * q + 1][2]);
      acc[4 * q + 2][0] = fmaf(d0, cf.z, acc[4 * q + 2][0]); acc[4 * q + 2][1] = fmaf(d1, cf.z, acc[4 * q + 2][1]); acc[4 * q + 2][2] = fmaf(d2, cf.z, acc[4 * q + 2][2]);
      acc[4 * q + 3][0] = fmaf(d0, cf.w, acc[4 * q + 3][0]); acc[4 * q + 3][1] = fmaf(d1, cf.w, acc[4 * q + 3][1]); acc[4 * q + 3][2] = fmaf(d2, cf.w, acc[4 * q + 3][2]);
    }
  }
  float w[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) w[j] = __ldg(m.weights + (size_t)j * kVPad + pv);
  int tip = -1;
#pragma unroll
  for (int t = 0; t < 5; ++t)
    if (v == tip_vertex(t)) tip = t;
#pragma unroll
  for (int c = 0; c < TC; ++c) {
    if (c0 + c >= n) break;
    float o[3];
    mano_skin_point<TC>(s, c, w, acc[c], o);
#pragma unroll
    for (int d = 0; d < 3; ++d) o[d] = mano_center_scale(o[d], s.G[c][0][d * 4 + 3]);
    if (verts) {
      float* dst = verts + ((size_t)(c0 + c) * kVerts + v) * 3;
      dst[0] = o[0]; dst[1] = o[1]; dst[2] = o[2];
    }
    if (tip >= 0) {
      float* dst = joints + ((size_t)(c0 + c) * 21 + tip_to_21(tip)) * 3;
      dst[0] = o[0]; dst[1] = o[1]; dst[2] = o[2];
    }
  }
}

template <int TC>
static int launch_mano_forward(const ManoModelDev& m, const float* pose, const float* shape, int pose_stride,
                               int shape_stride, int n, float* verts, float* joints, cudaStream_t stream) {
  dim3 grid((n + TC - 1) / TC, kNumVChunks);
  const size_t smem = sizeof(ManoSmem<TC>);
#ifndef VPHO_EMU
  {
    int dev = 0;
    static bool attr_set[64] = {};                   // function attributes are per device
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return VPHO_ERR_LAUNCH;
    if (!attr_set[dev]) {
      if (cudaFuncSetAttribute(mano_forward_kernel<TC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return VPHO_ERR_LAUNCH;
      attr_set[dev] = true;
    }
  }
#endif
  profile_begin(VPHO_TAG_MANO_FULL, stream);
  VPHO_LAUNCH_PDL(mano_forward_kernel<TC>, grid, dim3(kVChunkPad), smem, stream, m, pose, shape, pose_stride, shape_stride, n, verts,
              joints);
  profile_end(VPHO_TAG_MANO_FULL, stream);
  VPHO_CHECK_LAUNCH();
  return VPHO_OK;
}

// pose/shape rows may be strided (elements) so callers can pick e.g. candidate 0 of every image without a gather
// Vertices materialised: the tcgen05 kernel (mano_tc.cu).  Joints only (verts == nullptr), the explicit FP32 cross-check
// (simt = true) and the CPU emulator build: the SIMT kernel above.
int mano_forward_dev(const ManoModelDev& m, const float* pose, const float* shape, int pose_stride, int shape_stride,
                     int n, float* verts, float* joints, cudaStream_t stream, bool simt, int debug_blend) {
  if (n <= 0) return VPHO_OK;
#ifndef VPHO_EMU
  if (verts && !simt) return mano_tc_forward(m.tc_host, m, pose, shape, pose_stride, shape_stride, n, verts, joints, stream, debug_blend);
#endif
  if (n >= 148 * 32) return launch_mano_forward<16>(m, pose, shape, pose_stride, shape_stride, n, verts, joints, stream);
  if (n >= 148 * 4) return launch_mano_forward<8>(m, pose, shape, pose_stride, shape_stride, n, verts, joints, stream);
  return launch_mano_forward<4>(m, pose, shape, pose_stride, shape_stride, n, verts, joints, stream);
}

const ManoModelDev& mano_model_dev(const void* handle) { return static_cast<const ManoModelHost*>(handle)->dev; }

}  // namespace vpho

using namespace vpho;

extern "C" int vpho_mano_create(const float* v_template, const float* shapedirs, const float* posedirs,
                                const float* J_regressor, const float* weights, vpho_mano_t* out) {
  if (!v_template || !shapedirs || !posedirs || !J_regressor || !weights || !out) return VPHO_ERR_INVALID;
  const size_t n_dirs = (size_t)kBlendK * 3 * kVPad, n_tmpl = 3 * kVPad, n_w = 16 * kVPad;
  const size_t n_jt = 16 * 3, n_js = 16 * 3 * 10, n_td = 5 * 3 * kBlendK, n_tt = 5 * 3, n_tw = 5 * 16;
  const size_t total = n_dirs + n_tmpl + n_w + n_jt + n_js + n_td + n_tt + n_tw;
  std::vector<float> h(total, 0.f);
  float* dirs = h.data();
  float* tmpl = dirs + n_dirs;
  float* wgt = tmpl + n_tmpl;
  float* jt = wgt + n_w;
  float* js = jt + n_jt;
  float* td = js + n_js;
  float* tt = td + n_td;
  float* tw = tt + n_tt;
  const int tipv[5] = {745, 317, 444, 556, 673};
  for (int v = 0; v < kVerts; ++v) {
    const int pv = (v / kVChunk) * kVChunkPad + (v % kVChunk);
    for (int d = 0; d < 3; ++d) {
      tmpl[d * kVPad + pv] = v_template[v * 3 + d];
      for (int k = 0; k < 10; ++k) dirs[(size_t)(k * 3 + d) * kVPad + pv] = shapedirs[(v * 3 + d) * 10 + k];
      for (int k = 0; k < 135; ++k) dirs[(size_t)((10 + k) * 3 + d) * kVPad + pv] = posedirs[(v * 3 + d) * 135 + k];
    }
    for (int j = 0; j < 16; ++j) wgt[(size_t)j * kVPad + pv] = weights[v * 16 + j];
  }
  for (int j = 0; j < 16; ++j)
    for (int d = 0; d < 3; ++d) {
      double a = 0.0;
      for (int v = 0; v < kVerts; ++v) a += (double)J_regressor[j * kVerts + v] * (double)v_template[v * 3 + d];
      jt[j * 3 + d] = (float)a;
      for (int k = 0; k < 10; ++k) {
        double b = 0.0;
        for (int v = 0; v < kVerts; ++v) b += (double)J_regressor[j * kVerts + v] * (double)shapedirs[(v * 3 + d) * 10 + k];
        js[(j * 3 + d) * 10 + k] = (float)b;
      }
    }
  for (int t = 0; t < 5; ++t) {
    const int v = tipv[t];
    for (int d = 0; d < 3; ++d) {
      tt[t * 3 + d] = v_template[v * 3 + d];
      for (int k = 0; k < 10; ++k) td[(t * 3 + d) * kBlendK + k] = shapedirs[(v * 3 + d) * 10 + k];
      for (int k = 0; k < 135; ++k) td[(t * 3 + d) * kBlendK + 10 + k] = posedirs[(v * 3 + d) * 135 + k];
    }
    for (int j = 0; j < 16; ++j) tw[t * 16 + j] = weights[v * 16 + j];
  }
  ManoModelHost* mh = new ManoModelHost();
  if (cudaMalloc((void**)&mh->blob, total * sizeof(float)) != cudaSuccess) { delete mh; return VPHO_ERR_ALLOC; }
  if (cudaMemcpy(mh->blob, h.data(), total * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) {
    cudaFree(mh->blob); delete mh; return VPHO_ERR_ALLOC;
  }
  float* b = mh->blob;
  mh->dev.dirs = b;
  mh->dev.v_template = b + n_dirs;
  mh->dev.weights = mh->dev.v_template + n_tmpl;
  mh->dev.J_template = mh->dev.weights + n_w;
  mh->dev.J_shapedirs = mh->dev.J_template + n_jt;
  mh->dev.tip_dirs = mh->dev.J_shapedirs + n_js;
  mh->dev.tip_template = mh->dev.tip_dirs + n_td;
  mh->dev.tip_weights = mh->dev.tip_template + n_tt;
  mh->dev.tc_host = nullptr;
#ifndef VPHO_EMU
  {
    // the tensor-core layer is the product path: if its planes or TMA descriptors cannot be built the creation FAILS
    void* tc = nullptr;
    const int rc = mano_tc_create(v_template, shapedirs, posedirs, weights, &tc);
    if (rc != VPHO_OK) { cudaFree(mh->blob); delete mh; return rc; }
    mh->dev.tc_host = tc;
  }
#endif
  *out = mh;
  return VPHO_OK;
}

extern "C" int vpho_mano_destroy(vpho_mano_t h) {
  if (!h) return VPHO_ERR_INVALID;
  ManoModelHost* mh = static_cast<ManoModelHost*>(h);
#ifndef VPHO_EMU
  mano_tc_destroy(const_cast<void*>(mh->dev.tc_host));
#endif
  cudaFree(mh->blob);
  delete mh;
  return VPHO_OK;
}

extern "C" int vpho_mano_forward_ex(vpho_mano_t h, const float* pose, const float* shape, int n, float* verts,
                                    float* joints, int flags, void* stream) {
  if (!h || n < 0 || (flags & ~(VPHO_MANO_STRICT_FP32 | VPHO_MANO_DEBUG_BLEND))) return VPHO_ERR_INVALID;
  if (n == 0) return VPHO_OK;
  if (!pose || !shape || !joints) return VPHO_ERR_INVALID;
  return mano_forward_dev(static_cast<ManoModelHost*>(h)->dev, pose, shape, 48, 10, n, verts, joints,
                          (cudaStream_t)stream, (flags & VPHO_MANO_STRICT_FP32) != 0, (flags & VPHO_MANO_DEBUG_BLEND) ? 1 : 0);
}

extern "C" int vpho_mano_forward(vpho_mano_t h, const float* pose, const float* shape, int n, float* verts,
                                 float* joints, void* stream) {
  return vpho_mano_forward_ex(h, pose, shape, n, verts, joints, 0, stream);
}
