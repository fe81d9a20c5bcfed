// Shared declarations of the score-based sampler (score network + on-device RK45 controller).
//
// Reference path: ScoreBasedModelAgent.sample -> cond_ode_sampler (lib/model/score_based_model.py:45-105,130-146)
// driving BaseDenoiser.forward (lib/model/denoiser.py:68-82) through scipy.integrate.solve_ivp(method='RK45')
// (scipy/integrate/_ivp/rk.py, common.py, ivp.py -- un-vendored third party, algorithm restated in sampler.cu).
#pragma once
#include "rot_math.cuh"

namespace vpho {

constexpr int kTDim = 128;      // time-embedding width
constexpr int kPDim = 256;      // pose-feature width
constexpr int kFDim = 1024;     // conditioning feature width
constexpr int kHeadHid = 256;   // hidden width of each ParallelLinear head
constexpr int kRowTile = 128;   // candidate rows per CTA of the head GEMM; Npad is a multiple of it

struct DenoiserDev {
  int n_heads;   // 32 hand / 3 object
  int D;         // 3 * n_heads
  int hid;       // n_heads * 256
  const float* fourier_W;  // [64]
  const float* Wt;         // [128 in][128 out]   t_encoder.1.weight^T
  const float* bt;         // [128]
  const float* W1;         // [D][256]            pose_encoder.0.weight^T
  const float* b1;         // [256]
  const float* W2;         // [256][256]          pose_encoder.2.weight^T
  const float* b2;         // [256]
  const float* Wa_t;       // [128][hid]          head.0.weight[n][0:128][c]     -> [k][n*256+c]
  const float* Wa_p;       // [256][hid]          head.0.weight[n][128:384][c]
  const float* Wa_f;       // [1024][hid]         head.0.weight[n][384:1408][c]
  const float* ba;         // [hid]
  const float* Wb;         // [hid][4]            head.2.weight[n][c][0..2], padded to float4
  const float* bb;         // [D]
  const float* Wscale_inv; // [n_heads] exact power-of-two un-scaling of the FP16 weight planes (3xFP16 head GEMM)
  float W2scale_inv;       // the second pose-encoder GEMM runs on FP16 planes of W2 scaled by 1 / W2scale_inv (power of two)
};

enum StageMode : int {
  kModeInit0 = 0,   // f0 = fun(T0, y0): t is a python float -> float32 sde coefficients
  kModeInit1 = 1,   // f1 = fun(T0 + h0*dir, y0 + h0*dir*f0)            (select_initial_step)
  kModeStage = 2,   // RK stage s in 1..6 (6 = f_new at y_new)
  kModeFinal = 3,   // denoise predictor step at t = eps (score_based_model.py:95-104)
  kModeEval = 4,    // stand-alone score evaluation (vpho_score_eval)
};

// time of one network evaluation and the SDE scalars that go with it
struct EvalTime {
  float t32;      // time fed to the network: torch.ones(N,1) * t  -> float32
  float std32;    // sigma(t32) + 1e-7 in float32 (denoiser.py:78-81)
  double coef;    // 0.5 * g(t)^2 in float64 (score_based_model.py:84, numpy >= 2 promotion)
  float g32;      // float32 diffusion for the predictor step (sde_coeff(vec_eps))
};

struct RkCtrl {
  // configuration
  double T0, eps, rtol, atol, max_step;
  int n;            // N * D state size
  int n_rows, D, n_eval, num_steps;
  int rows_per_feat;
  double direction; // -1 (T0 > eps)
  // controller state (scipy RungeKutta attributes of the same names)
  double t, h_abs, h, t_new, h0;
  double d0, d1;
  int status;       // 0 running, 1 finished, -1 step too small
  int step_rejected;
  int accepted_now; // the attempt that just ended was accepted: post-step kernel must emit dense output + roll state
  double t_old, h_done;   // interval of the accepted step (for the dense output)
  int te_next;      // next t_eval index (descending times) still to be emitted
  int te_lo, te_hi; // window [lo,hi) emitted by the post-step kernel of this attempt
  int nfev, n_acc, n_rej, attempts, finished_final;
  int nan_seen;     // any NaN ever produced by the network (score_based_model.py:69-71)
  int nan_stage[8]; // K slot s holds non-finite values that must read as 0 (nan_to_num)
  double kcoef[8];  // K slot s holds raw float32 scores; its float64 value is -(kcoef[s] * score)  (kval)
  unsigned int block_counter;
  float eval_t32;   // vpho_score_eval: the time of a stand-alone evaluation
  EvalTime et[8];   // scalars of a network call, by time-term slot (tt_slot): RK stage s uses slot s, every other mode slot 0;
                    // an attempt's six slots are filled together by the first stage's kernel
  // caller-owned outputs of the running sample()
  double* xs;       // [n_eval][N][D] or nullptr
  float* xs32;      // the same dense output rounded to float32 (`.float()`), or nullptr
  double* x_out;    // [N][D]
  int32_t* counters;
};

struct SamplerWs {
  RkCtrl* ctrl;
  float* F;        // [R][hid]  feat-term + bias, once per sample()
  float* Fpart;    // [8][R][hid] split-K partial sums of the feat-term
  float* Tt;       // [7][hid]  time-term by slot (tt_slot): the six stages of an RK attempt are computed together
  float* P2T;      // [256][Npad] pose features, k-major (FP32-SIMT head GEMM)
  float* P2hi;     // [Npad][256] __half (hi, lo) planes of the pose features scaled per row, row-major = K-major (tcgen05
  float* P2lo;     // head GEMM); nullptr on the strict-FP32 SIMT path
  float* P2scale;  // [Npad] exact power-of-two un-scaling of each row of the FP16 planes; nullptr on the SIMT path
  double* y;       // [n]
  double* ynew;    // [n]
  float* K;        // [7][n] raw network outputs (float32) of the RK stages; see kval for the float64 drift they stand for
  double* partial; // [3][kMaxRedBlocks]
  double* t_eval;  // [n_eval]
  float* eval_out; // vpho_score_eval: output / input / geometry of a stand-alone evaluation
  const float* eval_x;
  int eval_rows, eval_rpf;
  int Npad;
  int R;
};

constexpr int kMaxRedBlocks = 512;

// one sampler's share of a tensor-core launch (scorenet_tc.cu); two of them = two samplers advancing in lock-step
struct TcHeadJob { const void *mapA_hi, *mapA_lo, *mapB_hi, *mapB_lo; const DenoiserDev* dn; const SamplerWs* ws; };
struct TcPoseJob { const void *mapW1_hi, *mapW1_lo, *mapW2_hi, *mapW2_lo; const DenoiserDev* dn; const SamplerWs* ws; };

}  // namespace vpho
