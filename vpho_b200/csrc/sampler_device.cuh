// Device-side pieces of the sampler shared by the FP32-SIMT head GEMM (sampler.cu) and the tcgen05 head GEMM
// (scorenet_tc.cu): Dormand-Prince coefficients, the time / SDE scalars of one network evaluation, and what happens
// to one network output element in each evaluation mode.
#pragma once
#include "sampler.cuh"

#include <cstring>

namespace vpho {

// ------------------------------------------------------------------------------------------------------------
// Dormand-Prince coefficients exactly as scipy/integrate/_ivp/rk.py (class RK45) writes them
// ------------------------------------------------------------------------------------------------------------
VPHO_CONSTANT double kC[7] = {0.0, 1.0 / 5, 3.0 / 10, 4.0 / 5, 8.0 / 9, 1.0, 1.0};
VPHO_CONSTANT double kA[7][6] = {
    {0, 0, 0, 0, 0, 0},
    {1.0 / 5, 0, 0, 0, 0, 0},
    {3.0 / 40, 9.0 / 40, 0, 0, 0, 0},
    {44.0 / 45, -56.0 / 15, 32.0 / 9, 0, 0, 0},
    {19372.0 / 6561, -25360.0 / 2187, 64448.0 / 6561, -212.0 / 729, 0, 0},
    {9017.0 / 3168, -355.0 / 33, 46732.0 / 5247, 49.0 / 176, -5103.0 / 18656, 0},
    {35.0 / 384, 0, 500.0 / 1113, 125.0 / 192, -2187.0 / 6784, 11.0 / 84}};   // row 6 = B (y_new)
VPHO_CONSTANT double kE[7] = {-71.0 / 57600, 0, 71.0 / 16695, -71.0 / 1920, 17253.0 / 339200, -22.0 / 525, 1.0 / 40};
VPHO_CONSTANT double kP[7][4] = {
    {1, -8048581381.0 / 2820520608, 8663915743.0 / 2820520608, -12715105075.0 / 11282082432},
    {0, 0, 0, 0},
    {0, 131558114200.0 / 32700410799, -68118460800.0 / 10900136933, 87487479700.0 / 32700410799},
    {0, -1754552775.0 / 470086768, 14199869525.0 / 1410260304, -10690763975.0 / 1880347072},
    {0, 127303824393.0 / 49829197408, -318862633887.0 / 49829197408, 701980252875.0 / 199316789632},
    {0, -282668133.0 / 205662961, 2019193451.0 / 616988883, -1453857185.0 / 822651844},
    {0, 40617522.0 / 29380423, -110615467.0 / 29380423, 69997945.0 / 29380423}};

constexpr double kSafety = 0.9, kMinFactor = 0.2, kMaxFactor = 10.0, kErrExponent = -0.2;
constexpr double kSigmaMin = 0.01, kSigmaRatio = 5000.0;   // sigma_max / sigma_min = 50 / 0.01 (sde.py:92-93)
// sqrt(2 * (log(50) - log(0.01))) evaluated in float64 (torch.tensor(np.float64), sde.py:23)
constexpr double kSqrt2LogRatio = 4.1272734804992597;

// ------------------------------------------------------------------------------------------------------------
// time of an evaluation and the SDE scalars that go with it
// ------------------------------------------------------------------------------------------------------------

__device__ __forceinline__ float sigma_f32(float t32) {
  // torch: 0.01 * (5000.0 ** t) on a float32 tensor; pow evaluated in double and rounded once
  float p = (float)pow(5000.0, (double)t32);
  return __fmul_rn(0.01f, p);
}

// float64 time of an evaluation (cheap); eval_time adds the SDE scalars (double pow: expensive on this chip's FP64
// pipe, so it is evaluated by ONE thread per network call -- the first kernel of the call stores it in ctrl->et)
__device__ __forceinline__ double eval_t64(const RkCtrl& c, int mode, int s) {
  if (mode == kModeInit0) return c.T0;
  if (mode == kModeInit1) return c.T0 + c.h0 * c.direction;
  if (mode == kModeStage) return (s == 6) ? (c.t + c.h) : (c.t + kC[s] * c.h);
  if (mode == kModeFinal) return c.eps;
  return (double)c.eval_t32;
}

__device__ __forceinline__ EvalTime eval_time(const RkCtrl& c, int mode, int s) {
  EvalTime e;
  const double t64 = eval_t64(c, mode, s);
  e.t32 = (float)t64;
  e.std32 = __fadd_rn(sigma_f32(e.t32), 1e-7f);
  double sigma;
  if (mode == kModeInit0) sigma = (double)sigma_f32((float)c.T0);          // torch.tensor(python float) is float32
  else sigma = kSigmaMin * pow(kSigmaRatio, t64);
  double g = sigma * kSqrt2LogRatio;
  e.coef = 0.5 * (g * g);
  e.g32 = __fmul_rn(sigma_f32(e.t32), (float)kSqrt2LogRatio);
  return e;
}

// K slot value with the reference's `nan_to_num(score, 0, 0, 0)` applied when that evaluation produced a NaN
// The slot stores the float32 score; the drift  -0.5 g(t)^2 * score  is formed here in float64 exactly as
// `drift - 0.5 * diffusion**2 * score` is under numpy >= 2 promotion, so keeping the narrow value in
// memory halves the K traffic of every RK kernel without changing a bit of the result.
// float -> double, exact, on the integer pipe.  The FP64 pipe of this part issues only a few lanes per clock per SM (measured:
// the stage-input phase of the pose kernel is bound by it), and F2F.F64.F32 runs there too: widening by bit manipulation
// takes one of the three FP64-pipe instructions per (element, K slot) off it.  Zeros, subnormals, infinities and NaNs take
// the conversion instruction.
__device__ __forceinline__ double widen_f32(float f) {
#if defined(__CUDA_ARCH__)
  const unsigned u = __float_as_uint(f), e = (u >> 23) & 0xFFu;
  if (e == 0u || e == 255u) return (double)f;
  return __hiloint2double((int)((u & 0x80000000u) | (((u & 0x7FFFFFFFu) >> 3) + 0x38000000u)), (int)(u << 29));
#else
  return (double)f;
#endif
}
__device__ __forceinline__ double kval(const float* K, const RkCtrl& c, int slot, int n, int i) {
  double v = -(c.kcoef[slot] * widen_f32(K[(size_t)slot * n + i]));
  if (c.nan_stage[slot] && !isfinite(v)) v = 0.0;
  return v;
}
// time-term / EvalTime slot of a network call: the stage index inside an RK attempt, 0 for the other modes
__device__ __forceinline__ int tt_slot(int mode, int s) { return mode == kModeStage ? s : 0; }
__device__ __forceinline__ int k_slot_of(int mode, int s) { return (mode == kModeInit0) ? 0 : (mode == kModeInit1 ? 1 : s); }

__device__ __forceinline__ bool eval_active(const RkCtrl& c, int mode) {
  if (mode == kModeEval) return true;
  if (mode == kModeFinal) return c.status == 1;
  return c.status == 0;
}

// ------------------------------------------------------------------------------------------------------------
// what happens to one network output element, per evaluation mode
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void emit_score(const SamplerWs& ws, RkCtrl& c, const EvalTime& et, int mode, int s, int i,
                                           float score) {
  if (mode == kModeEval) { ws.eval_out[i] = score; return; }
  if (mode == kModeFinal) {
    // drift = 0 - diffusion**2 * grad (float32); x = x + drift * ((1-eps)/num_steps)   (score_based_model.py:95-104)
    const float g2 = __fmul_rn(et.g32, et.g32);
    const float drift = __fsub_rn(0.f, __fmul_rn(g2, score));
    const float scale = (float)((1.0 - c.eps) / (double)c.num_steps);
    c.x_out[i] = __dadd_rn(ws.y[i], (double)__fmul_rn(drift, scale));
    return;
  }
  const int slot = k_slot_of(mode, s);
  VPHO_BOUNDS(slot >= 0 && slot < 7 && i >= 0 && i < c.n);
  if (isnan(score)) { c.nan_stage[slot] = 1; c.nan_seen = 1; }
  ws.K[(size_t)slot * c.n + i] = score;          // read back through kval with kcoef[slot] = et.coef of this call
}


// Input of one network evaluation for element (row, d): the float64 RK combination that scipy hands to `fun`
// (rk.py rk_step / common.py select_initial_step), or the raw state / stand-alone input.  Returns 0 for padding rows.
// Side effect: the 7th stage (s == 6) stores y_new.
__device__ __forceinline__ double stage_input(const SamplerWs& ws, const RkCtrl& c, int mode, int s, int row, int d, int n_rows,
                                              int D) {
  if (row >= n_rows) return 0.0;
  const int i = row * D + d, n = c.n;
  if (mode == kModeEval) return (double)ws.eval_x[i];
  const double y = ws.y[i];
  if (mode == kModeInit0 || mode == kModeFinal) return y;
  if (mode == kModeInit1) return __dadd_rn(y, __dmul_rn(c.h0 * c.direction, kval(ws.K, c, 0, n, i)));
  // dy = np.dot(K[:s].T, a[:s]) * h ; y + dy          (rk.py rk_step)
  double acc = 0.0;
  const int ns = (s == 6) ? 6 : s;
  for (int j = 0; j < ns; ++j) acc += kval(ws.K, c, j, n, i) * kA[s][j];
  const double v = __dadd_rn(y, __dmul_rn(acc, c.h));
  if (s == 6) ws.ynew[i] = v;
  return v;
}

// The same stage input for FOUR consecutive elements i..i+3 (i a multiple of 4, D a multiple of 4), with the controller
// scalars read once (`StageScalars`) and the K slots / state read as 16-byte vectors.  Operation for operation the arithmetic
// of stage_input / kval above -- the fused pose-encoder kernel runs this 6 times per thread per network call, where the
// scalar form spends its time re-reading the controller words.
struct StageScalars {
  double kc[6];     // kcoef of the K slots that enter the combination
  double a[6];      // their Dormand-Prince weights kA[s][j]
  double h;         // step (h0 * direction for the second initial-step evaluation)
  int nanf[6];
  int ns;           // slots that enter (0: the input is the state itself)
};
__device__ __forceinline__ StageScalars stage_scalars(const RkCtrl& c, int mode, int s) {
  StageScalars q;
  q.ns = 0;
  q.h = 0.0;
  if (mode == kModeInit1) { q.ns = 1; q.h = c.h0 * c.direction; }
  if (mode == kModeStage) { q.ns = (s == 6) ? 6 : s; q.h = c.h; }
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    const double kcj = c.kcoef[j];            // unconditional: all the controller words in flight at once
    const int nfj = c.nan_stage[j];
    q.kc[j] = j < q.ns ? kcj : 0.0;
    q.nanf[j] = j < q.ns ? nfj : 0;
    q.a[j] = mode == kModeStage ? kA[s][j] : 1.0;
  }
  return q;
}
// The vector form comes in two halves (loads / arithmetic) so that the caller can keep several units' loads in flight before
// the float64 arithmetic starts (the load latency of one unit is otherwise paid once per unit).  `ns` is the number of K
// slots that enter (host-known from mode / s: stage_ns).
struct StageRaw {
  double2 y01, y23;
  float4 kv[6];
};
__host__ __device__ __forceinline__ int stage_ns(int mode, int s) {
  return mode == kModeInit1 ? 1 : (mode == kModeStage ? ((s == 6) ? 6 : s) : 0);
}
__device__ __forceinline__ void stage_load4(const SamplerWs& ws, int mode, int ns, int i, int n, StageRaw& r) {
  if (mode == kModeEval) {
    r.kv[0] = *reinterpret_cast<const float4*>(ws.eval_x + i);
    return;
  }
#if defined(__CUDA_ARCH__)
  // one 32-byte load per unit (sm_100 256-bit LDG): with two 16-byte loads the lanes of a warp touch every 32-byte sector twice
  asm volatile("ld.global.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(r.y01.x), "=d"(r.y01.y), "=d"(r.y23.x), "=d"(r.y23.y) : "l"(ws.y + i) : "memory");
#else
  r.y01 = *reinterpret_cast<const double2*>(ws.y + i);
  r.y23 = *reinterpret_cast<const double2*>(ws.y + i + 2);
#endif
#pragma unroll
  for (int j = 0; j < 6; ++j)
    if (j < ns) r.kv[j] = *reinterpret_cast<const float4*>(ws.K + (size_t)j * n + i);
}
__device__ __forceinline__ void stage_math4(const SamplerWs& ws, const StageScalars& q, int mode, int s, int i, const StageRaw& r,
                                            float* out) {
  if (mode == kModeEval) {
    out[0] = r.kv[0].x; out[1] = r.kv[0].y; out[2] = r.kv[0].z; out[3] = r.kv[0].w;
    return;
  }
  const double y[4] = {r.y01.x, r.y01.y, r.y23.x, r.y23.y};
  if (q.ns == 0) {
#pragma unroll
    for (int e = 0; e < 4; ++e) out[e] = (float)y[e];
    return;
  }
  double v[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    if (mode == kModeInit1) {
      const float k0 = e == 0 ? r.kv[0].x : e == 1 ? r.kv[0].y : e == 2 ? r.kv[0].z : r.kv[0].w;
      double kvv = -(q.kc[0] * widen_f32(k0));
      if (q.nanf[0] && !isfinite(kvv)) kvv = 0.0;
      v[e] = __dadd_rn(y[e], __dmul_rn(q.h, kvv));
    } else {
      double acc = 0.0;
#pragma unroll
      for (int j = 0; j < 6; ++j)
        if (j < q.ns) {
          const float kj = e == 0 ? r.kv[j].x : e == 1 ? r.kv[j].y : e == 2 ? r.kv[j].z : r.kv[j].w;
          double kvv = -(q.kc[j] * widen_f32(kj));
          if (q.nanf[j] && !isfinite(kvv)) kvv = 0.0;
          acc += kvv * q.a[j];
        }
      v[e] = __dadd_rn(y[e], __dmul_rn(acc, q.h));
    }
    out[e] = (float)v[e];
  }
  if (mode == kModeStage && s == 6) {
#if defined(__CUDA_ARCH__)
    asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(ws.ynew + i), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]) : "memory");
#else
    *reinterpret_cast<double2*>(ws.ynew + i) = make_double2(v[0], v[1]);
    *reinterpret_cast<double2*>(ws.ynew + i + 2) = make_double2(v[2], v[3]);
#endif
  }
}

// round-to-nearest (ties to even) onto the 10-bit TF32 mantissa, result kept in a float container
__host__ __device__ __forceinline__ float tf32_round(float x) {
  unsigned u;
  memcpy(&u, &x, 4);
  if ((u & 0x7F800000u) == 0x7F800000u) return x;
  u += 0xFFFu + ((u >> 13) & 1u);
  u &= 0xFFFFE000u;
  float r;
  memcpy(&r, &u, 4);
  return r;
}


// ------------------------------------------------------------------------------------------------------------
// time-term: Fourier embedding -> t_encoder -> Tt[col] = sum_k t_feat[k] Wa_t[k][col]   (one per network call)
// ------------------------------------------------------------------------------------------------------------
// One block = 64 output columns of Tt.  The 128-long contraction over t_feat is split over 4 thread groups (32 k each,
// all loads in flight at once) and summed in a fixed order, so the latency is one L2 round trip instead of 128.
constexpr int kTtCols = 64;
__device__ __forceinline__ void time_term_block(const DenoiserDev& dn, const SamplerWs& ws, const RkCtrl& c, int mode, int s,
                                                int block) {
  __shared__ float four[kTDim];
  __shared__ float tfeat[kTDim];
  __shared__ float part[4][kTtCols];
  const int tid = threadIdx.x;
  const float t32 = (float)eval_t64(c, mode, s);
  if (block == 0 && tid == 255) {
    const EvalTime et = eval_time(c, mode, s);
    ws.ctrl->et[tt_slot(mode, s)] = et;                                    // read by the head GEMM of that call
    if (mode == kModeInit0 || mode == kModeInit1 || mode == kModeStage) ws.ctrl->kcoef[k_slot_of(mode, s)] = et.coef;
  }
  if (tid < 64) {
    // x_proj = t * W * 2 * np.pi in float32, left to right (denoiser.py:29-31)
    float xp = __fmul_rn(__fmul_rn(__fmul_rn(t32, dn.fourier_W[tid]), 2.0f), 3.14159265358979323846f);
    four[tid] = (float)sin((double)xp);
    four[64 + tid] = (float)cos((double)xp);
  }
  __syncthreads();
  if (tid < kTDim) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 8
    for (int k = 0; k < kTDim; k += 4) {
      a0 = fmaf(four[k + 0], __ldg(dn.Wt + (k + 0) * kTDim + tid), a0);
      a1 = fmaf(four[k + 1], __ldg(dn.Wt + (k + 1) * kTDim + tid), a1);
      a2 = fmaf(four[k + 2], __ldg(dn.Wt + (k + 2) * kTDim + tid), a2);
      a3 = fmaf(four[k + 3], __ldg(dn.Wt + (k + 3) * kTDim + tid), a3);
    }
    const float a = ((a0 + a1) + (a2 + a3)) + dn.bt[tid];
    tfeat[tid] = a > 0.f ? a : 0.f;
  }
  __syncthreads();
  const int g = tid >> 6, cl = tid & 63, col = block * kTtCols + cl;
  float a = 0.f;
  if (col < dn.hid) {
#pragma unroll
    for (int k = 0; k < 32; ++k) a = fmaf(tfeat[g * 32 + k], __ldg(dn.Wa_t + (size_t)(g * 32 + k) * dn.hid + col), a);
  }
  part[g][cl] = a;
  __syncthreads();
  if (tid < kTtCols && col < dn.hid)
    ws.Tt[(size_t)tt_slot(mode, s) * dn.hid + col] = (part[0][tid] + part[1][tid]) + (part[2][tid] + part[3][tid]);
}


}  // namespace vpho
