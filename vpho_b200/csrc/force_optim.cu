// ForceOptimizer.optimize_batch (lib/engine/force_optimization.py:110-207) as ONE persistent kernel: the reference runs
// 3000 autograd iterations of two torch.optim.AdamW optimisers over scale [bs][32] and weight [bs][32][8]; here the backward
// pass is written out analytically and the whole loop stays on the device (no launch, no host round trip per iteration).
//
// One CTA of 32 warps; a warp owns hand b = warp, warp + 32, ... and a lane owns force anchor j of that hand, so every
// per-hand sum (resultant force, moment about the centre of mass, |scale|) is a fixed-order warp butterfly and the one
// batch-wide quantity of an iteration (sum_weight = force_loss.detach(), the mean over hands of |sum f + g|, :153-154) is one
// __syncthreads.  The hand geometry (anchor points, local frames: VERT2ANCHOR, lib/utils/physics_fn.py:224-257) does not
// change during the optimisation and is computed once.
//
// Forward (per anchor):  s' = s * mask;  p = softmax(w);  a = sum_k p_k A_k (friction-cone anchors, physics.py:546-557);
//   dir = a / (|a| + 1e-8);  f_local = dir |s'|;  f_global = Frame f_local.
// Losses (:150-176):  force = mean_b |sum_j f_j + g|;  gravity = mean_b ((sum_j f_j).(-g) - 1)^2;
//   moment = 30 mean_b |sum_j (p_j - CoM) x f_j| / (100 sw^2 + 1e-8);  dist = 0.1 mean_{b,j} (log(|c_j / (s'_j / (|s'| + 1e-8) +
//   1e-8)| + 1e-8) mask_j)^2 / (1000 sw^2 + 1e-8), c = force_contact / (|force_contact| + 1e-8); |s'| and sw detached.
//   i < switch_iter: gravity loss, optimiser 1 (weight only); afterwards force + moment + dist, optimiser 2 (scale, weight).
// AdamW step (torch.optim.AdamW defaults, weight_decay 0.01):  p *= 1 - lr wd;  m = b1 m + (1 - b1) g;  v = b2 v + (1 - b2) g^2;
//   p -= (lr / (1 - b1^t)) m / (sqrt(v) / sqrt(1 - b2^t) + eps).
#include "agg_device.cuh"
#include "vpho_b200.h"

namespace vpho {

constexpr int kFoWarps = 32;

struct FoArgs {
  const float* verts;           // [n][778][3] camera frame (flipped frame for left hands, :134-137)
  const float* force_contact;   // [n][32]
  const float* gravity;         // [n][3]
  const float* com;             // [n][3]
  const unsigned char* is_grasped;   // [n] or nullptr
  const float* cone;            // [8][3] friction-cone anchors (xy already scaled by the friction coefficient)
  int n, n_iter, switch_iter;
  float lr, beta1, beta2, eps, weight_decay;
  float* scale;                 // [n][32]     out: parameters after the last step
  float* weight;                // [n][32][8]
  float* force_local;           // [n][32][3]  out: forward pass of the LAST iteration (what the reference saves, :185-206)
  float* force_global;          // [n][32][3]
  float* losses;                // [n_iter][5] or nullptr: loss, force, gravity, moment, dist of every iteration
  float* ws;                    // workspace: geometry [n][32][12] + Adam moments [n][32][36]
};

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
  return v;
}

__global__ void __launch_bounds__(kFoWarps * 32, 1) k_force_optimize(AssetsDev as, FoArgs a) {
  __shared__ float j21[kFoWarps][63];
  __shared__ float s_red[kFoWarps][4];
  __shared__ float s_cone[8][3];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid < 24) s_cone[tid / 3][tid % 3] = a.cone[tid];
  float* geo = a.ws;                                   // [n][32][12]: frame columns (dx, dy, dz) then arm = point - CoM
  float* adam = a.ws + (size_t)a.n * 32 * 12;          // [n][32][36]: opt1 (m_w, v_w) 16, opt2 (m_s, v_s, m_w, v_w) 18, pad
  // ---- geometry, once: joints = vert2joint . verts (hand_fn.py:436-448), anchor frames
  for (int b = warp; b < a.n; b += kFoWarps) {
    const float* v = a.verts + (size_t)b * kVerts * 3;
    for (int o = 0; o < 63; ++o) {
      const int k = o / 3, d = o % 3;
      float acc = 0.f;
      for (int vv = lane; vv < kVerts; vv += 32) acc = fmaf(as.v2j[k * kVerts + vv], v[vv * 3 + d], acc);
#pragma unroll
      for (int sft = 16; sft > 0; sft >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, sft);
      if (lane == 0) j21[warp][o] = acc;
    }
    __syncwarp();
    float pt[3], col[3][3];
    auto vert = [&](int vid, float* out) { out[0] = v[vid * 3 + 0]; out[1] = v[vid * 3 + 1]; out[2] = v[vid * 3 + 2]; };
    for (int e = 0; e < 3; ++e) {
      const float unit[3] = {e == 0 ? 1.f : 0.f, e == 1 ? 1.f : 0.f, e == 2 ? 1.f : 0.f};
      anchor_point_and_force(as, lane, vert, j21[warp], unit, pt, col[e]);       // column e of the frame
    }
    float* gg = geo + ((size_t)b * 32 + lane) * 12;
    for (int e = 0; e < 3; ++e)
      for (int d = 0; d < 3; ++d) gg[e * 3 + d] = col[e][d];
    for (int d = 0; d < 3; ++d) gg[9 + d] = pt[d] - a.com[(size_t)b * 3 + d];
    float* ad = adam + ((size_t)b * 32 + lane) * 36;
    for (int i = 0; i < 36; ++i) ad[i] = 0.f;
    a.scale[(size_t)b * 32 + lane] = 0.05f;                                      // init_param (:42-44)
    for (int k = 0; k < 8; ++k) a.weight[((size_t)b * 32 + lane) * 8 + k] = 0.f;
    __syncwarp();
  }
  __syncthreads();
  const float inv_bs = 1.f / (float)a.n;
  for (int it = 0; it < a.n_iter; ++it) {
    const bool phase2 = it >= a.switch_iter;
    const int t_step = phase2 ? it - a.switch_iter + 1 : it + 1;                 // step count of the optimiser in use
    // ---------------- pass 1: forward and the per-hand sums; batch sum of |R_b| through shared memory
    float nR_mine = 0.f, grav_mine = 0.f, nM_mine = 0.f, dist_mine = 0.f;
    for (int b = warp; b < a.n; b += kFoWarps) {
      const size_t bj = (size_t)b * 32 + lane;
      const float mask = a.force_contact[bj] > 0.1f ? 1.f : 0.f;
      const float sm = a.scale[bj] * mask;
      const float* w = a.weight + bj * 8;
      float mx = w[0];
#pragma unroll
      for (int k = 1; k < 8; ++k) mx = fmaxf(mx, w[k]);
      float p[8], es = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) { p[k] = expf(w[k] - mx); es += p[k]; }
      float av[3] = {0.f, 0.f, 0.f};
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        p[k] = p[k] / es;
#pragma unroll
        for (int d = 0; d < 3; ++d) av[d] += p[k] * s_cone[k][d];
      }
      const float an = sqrtf((av[0] * av[0] + av[1] * av[1]) + av[2] * av[2]);
      const float mag = fabsf(sm);
      float fl[3], fg[3];
#pragma unroll
      for (int d = 0; d < 3; ++d) fl[d] = (av[d] / (an + 1e-8f)) * mag;
      const float* gg = geo + bj * 12;
#pragma unroll
      for (int d = 0; d < 3; ++d) fg[d] = (fl[0] * gg[0 * 3 + d] + fl[1] * gg[1 * 3 + d]) + fl[2] * gg[2 * 3 + d];
      const float arm[3] = {gg[9], gg[10], gg[11]};
      const float Fs[3] = {warp_sum_f(fg[0]), warp_sum_f(fg[1]), warp_sum_f(fg[2])};
      const float Ms[3] = {warp_sum_f(arm[1] * fg[2] - arm[2] * fg[1]), warp_sum_f(arm[2] * fg[0] - arm[0] * fg[2]),
                           warp_sum_f(arm[0] * fg[1] - arm[1] * fg[0])};
      const float* g = a.gravity + (size_t)b * 3;
      const float R[3] = {Fs[0] + g[0], Fs[1] + g[1], Fs[2] + g[2]};
      nR_mine += sqrtf((R[0] * R[0] + R[1] * R[1]) + R[2] * R[2]);
      const float cosp = (Fs[0] * (-g[0]) + Fs[1] * (-g[1])) + Fs[2] * (-g[2]);
      grav_mine += (cosp - 1.f) * (cosp - 1.f);
      nM_mine += sqrtf((Ms[0] * Ms[0] + Ms[1] * Ms[1]) + Ms[2] * Ms[2]);
      if (a.losses) {
        const float sn = sqrtf(warp_sum_f(sm * sm)), cn = sqrtf(warp_sum_f(a.force_contact[bj] * a.force_contact[bj]));
        const float u = sm / (sn + 1e-8f) + 1e-8f, r = (a.force_contact[bj] / (cn + 1e-8f)) / u;
        const float dd = logf(fabsf(r) + 1e-8f) * mask;
        dist_mine += warp_sum_f(dd * dd);
      }
      if (it == a.n_iter - 1) {
        const float keep = (a.is_grasped && !a.is_grasped[b]) ? 0.f : 1.f;         // :191-194
#pragma unroll
        for (int d = 0; d < 3; ++d) { a.force_local[bj * 3 + d] = fl[d] * keep; a.force_global[bj * 3 + d] = fg[d] * keep; }
      }
    }
    if (lane == 0) { s_red[warp][0] = nR_mine; s_red[warp][1] = grav_mine; s_red[warp][2] = nM_mine; s_red[warp][3] = dist_mine; }
    __syncthreads();
    float tot[4] = {0.f, 0.f, 0.f, 0.f};
    for (int wv = 0; wv < kFoWarps; ++wv)
      for (int k = 0; k < 4; ++k) tot[k] += s_red[wv][k];
    const float sw = tot[0] * inv_bs;                                              // force_loss, detached as sum_weight
    const float cm = 30.f / (100.f * sw * sw + 1e-8f), cd = 0.1f / (1000.f * sw * sw + 1e-8f);
    if (a.losses && tid == 0) {
      const float fl_ = sw, gl_ = tot[1] * inv_bs, ml_ = tot[2] * inv_bs * cm, dl_ = tot[3] * inv_bs / 32.f * cd;
      float* lo = a.losses + (size_t)it * 5;
      lo[0] = phase2 ? fl_ + ml_ + dl_ : gl_; lo[1] = fl_; lo[2] = gl_; lo[3] = ml_; lo[4] = dl_;
    }
    __syncthreads();                                                               // s_red is rewritten next iteration
    // ---------------- pass 2: analytic backward and the AdamW step, per anchor (forward recomputed: cheaper than storing it)
    const float b1 = a.beta1, b2 = a.beta2;
    const float bc1 = 1.f - powf(b1, (float)t_step), bc2s = sqrtf(1.f - powf(b2, (float)t_step));
    const float step_size = a.lr / bc1, decay = 1.f - a.lr * a.weight_decay;
    for (int b = warp; b < a.n; b += kFoWarps) {
      const size_t bj = (size_t)b * 32 + lane;
      const float fc = a.force_contact[bj];
      const float mask = fc > 0.1f ? 1.f : 0.f;
      const float s_raw = a.scale[bj], sm = s_raw * mask;
      float* w = a.weight + bj * 8;
      float mx = w[0];
#pragma unroll
      for (int k = 1; k < 8; ++k) mx = fmaxf(mx, w[k]);
      float p[8], es = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) { p[k] = expf(w[k] - mx); es += p[k]; }
      float av[3] = {0.f, 0.f, 0.f};
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        p[k] = p[k] / es;
#pragma unroll
        for (int d = 0; d < 3; ++d) av[d] += p[k] * s_cone[k][d];
      }
      const float an = sqrtf((av[0] * av[0] + av[1] * av[1]) + av[2] * av[2]), ane = an + 1e-8f;
      const float mag = fabsf(sm);
      float dir[3], fg[3];
#pragma unroll
      for (int d = 0; d < 3; ++d) dir[d] = av[d] / ane;
      const float* gg = geo + bj * 12;
#pragma unroll
      for (int d = 0; d < 3; ++d) fg[d] = ((dir[0] * mag) * gg[0 * 3 + d] + (dir[1] * mag) * gg[1 * 3 + d]) + (dir[2] * mag) * gg[2 * 3 + d];
      const float arm[3] = {gg[9], gg[10], gg[11]};
      const float Fs[3] = {warp_sum_f(fg[0]), warp_sum_f(fg[1]), warp_sum_f(fg[2])};
      const float* g = a.gravity + (size_t)b * 3;
      // dL / d f_global of this anchor
      float gf[3];
      if (!phase2) {
        const float cosp = (Fs[0] * (-g[0]) + Fs[1] * (-g[1])) + Fs[2] * (-g[2]);
        const float c2 = 2.f * (cosp - 1.f) * inv_bs;
#pragma unroll
        for (int d = 0; d < 3; ++d) gf[d] = c2 * (-g[d]);
      } else {
        const float Ms[3] = {warp_sum_f(arm[1] * fg[2] - arm[2] * fg[1]), warp_sum_f(arm[2] * fg[0] - arm[0] * fg[2]),
                             warp_sum_f(arm[0] * fg[1] - arm[1] * fg[0])};
        const float R[3] = {Fs[0] + g[0], Fs[1] + g[1], Fs[2] + g[2]};
        const float nR = sqrtf((R[0] * R[0] + R[1] * R[1]) + R[2] * R[2]), nM = sqrtf((Ms[0] * Ms[0] + Ms[1] * Ms[1]) + Ms[2] * Ms[2]);
        const float iR = nR > 0.f ? inv_bs / nR : 0.f, iM = nM > 0.f ? cm * inv_bs / nM : 0.f;      // norm backward: 0 at 0
        // d|M| / d f_j = Mhat x arm_j
        gf[0] = R[0] * iR + (Ms[1] * arm[2] - Ms[2] * arm[1]) * iM;
        gf[1] = R[1] * iR + (Ms[2] * arm[0] - Ms[0] * arm[2]) * iM;
        gf[2] = R[2] * iR + (Ms[0] * arm[1] - Ms[1] * arm[0]) * iM;
      }
      // f_global = Frame f_local  ->  g_fl = Frame^T g_f ;  f_local = dir |s'|
      float gfl[3];
#pragma unroll
      for (int e = 0; e < 3; ++e) gfl[e] = (gg[e * 3 + 0] * gf[0] + gg[e * 3 + 1] * gf[1]) + gg[e * 3 + 2] * gf[2];
      const float g_mag = (dir[0] * gfl[0] + dir[1] * gfl[1]) + dir[2] * gfl[2];
      // dir = a / (|a| + eps):  g_a = g_dir / (n + eps) - a (a . g_dir) / (n (n + eps)^2)
      const float adg = ((av[0] * gfl[0] + av[1] * gfl[1]) + av[2] * gfl[2]) * mag;
      float ga[3];
#pragma unroll
      for (int d = 0; d < 3; ++d) ga[d] = (gfl[d] * mag) / ane - (an > 0.f ? av[d] * adg / (an * ane * ane) : 0.f);
      float gp[8], pg = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        gp[k] = (s_cone[k][0] * ga[0] + s_cone[k][1] * ga[1]) + s_cone[k][2] * ga[2];
        pg += p[k] * gp[k];
      }
      float* ad = adam + bj * 36;
      float* mw = ad + (phase2 ? 18 : 0);
      float* vw = mw + 8;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float gk = p[k] * (gp[k] - pg);                                       // softmax backward
        float wk = w[k] * decay;
        const float m = b1 * mw[k] + (1.f - b1) * gk, v = b2 * vw[k] + (1.f - b2) * gk * gk;
        mw[k] = m; vw[k] = v;
        wk -= step_size * (m / (sqrtf(v) / bc2s + a.eps));
        w[k] = wk;
      }
      if (phase2) {
        float gs = (sm > 0.f ? 1.f : (sm < 0.f ? -1.f : 0.f)) * g_mag;             // |s'| backward
        // contact-distribution term: d/ds' of cd/(32 bs) (log(|c / u| + 1e-8) mask)^2,  u = s' / (|s'|_2 + 1e-8) + 1e-8
        const float sn = sqrtf(warp_sum_f(sm * sm)), cn = sqrtf(warp_sum_f(fc * fc));
        const float cnrm = fc / (cn + 1e-8f), u = sm / (sn + 1e-8f) + 1e-8f, r = cnrm / u;
        const float dd = logf(fabsf(r) + 1e-8f) * mask;
        const float sgn = r > 0.f ? 1.f : (r < 0.f ? -1.f : 0.f);
        const float dr = sgn * (-cnrm / (u * u)) / (fabsf(r) + 1e-8f);
        gs += (cd * inv_bs / 32.f) * 2.f * dd * mask * dr / (sn + 1e-8f);
        gs *= mask;                                                                 // s' = s * mask
        float sv = s_raw * decay;
        const float m = b1 * ad[16] + (1.f - b1) * gs, v = b2 * ad[17] + (1.f - b2) * gs * gs;
        ad[16] = m; ad[17] = v;
        sv -= step_size * (m / (sqrtf(v) / bc2s + a.eps));
        a.scale[bj] = sv;
      }
    }
    __syncthreads();
  }
}

}  // namespace vpho

using namespace vpho;

extern "C" size_t vpho_force_optimize_workspace_bytes(int n) { return n > 0 ? (size_t)n * 32 * (12 + 36) * sizeof(float) : 0; }

extern "C" int vpho_force_optimize(vpho_assets_t h, const float* verts, const float* force_contact, const float* gravity,
                                   const float* com, const uint8_t* is_grasped, const float* cone_anchor, int n, int n_iter,
                                   int switch_iter, float lr, float* scale, float* weight, float* force_local, float* force_global,
                                   float* losses, void* workspace, size_t workspace_bytes, void* stream) {
  if (!h || n < 0 || n_iter < 1 || switch_iter < 0) return VPHO_ERR_INVALID;
  if (n == 0) return VPHO_OK;
  if (!verts || !force_contact || !gravity || !com || !cone_anchor || !scale || !weight || !force_local || !force_global || !workspace ||
      workspace_bytes < vpho_force_optimize_workspace_bytes(n))
    return VPHO_ERR_INVALID;
  FoArgs a{verts, force_contact, gravity, com, is_grasped, cone_anchor, n, n_iter, switch_iter, lr, 0.9f, 0.999f, 1e-8f, 0.01f,
           scale, weight, force_local, force_global, losses, static_cast<float*>(workspace)};
  VPHO_LAUNCH(k_force_optimize, dim3(1), dim3(kFoWarps * 32), 0, (cudaStream_t)stream, static_cast<AssetsHost*>(h)->dev, a);
  VPHO_CHECK_LAUNCH();
  return VPHO_OK;
}
