// Device-side pieces of the sampler shared by the FP32-SIMT head GEMM (sampler.cu) and the tcgen05 head GEMM
// (scorenet_tc.cu): Dormand-Prince coefficients, the time / SDE scalars of one network evaluation, and what happens
// to one network output element in each evaluation mode.
#pragma once
#include "sampler.cuh"

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
__device__ __forceinline__ double kval(const float* K, const RkCtrl& c, int slot, int n, int i) {
  double v = -(c.kcoef[slot] * (double)K[(size_t)slot * n + i]);
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

}  // namespace vpho
