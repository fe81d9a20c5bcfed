// Score-based sampler for sm_100a: VE-SDE probability-flow ODE integrated by an adaptive Dormand-Prince RK45
// controller that lives entirely on the device (no host round trip per stage), around a factored score network.
//
// Replaces ScoreBasedModelAgent.sample -> cond_ode_sampler (lib/model/score_based_model.py:45-105,130-146),
// BaseDenoiser.forward (lib/model/denoiser.py:68-82), ParallelLinear (lib/model/parallel_linear.py:27-35) and
// the VE-SDE coefficients (lib/model/sde.py:15-24).  scipy.integrate.solve_ivp(RK45) is an un-vendored third
// party (reference pins scipy==1.12.0): its step controller (rk.py RungeKutta._step_impl, rk_step), initial step
// heuristic (common.py select_initial_step), RMS norm (common.py norm), t_eval handling (ivp.py) and quartic
// dense output (rk.py RkDenseOutput, RK45.P) are restated here operation by operation.
//
// Factoring of the first ParallelLinear (K = 1408 = 128 time + 256 pose + 1024 conditioning features):
//   feat-term  F[img][hid]   once per sample()        (k_feat_term)
//   time-term  Tt[hid]       once per network call    (k_time_term; all rows share t on the sampling path)
//   pose-term  P2 . Wa_p     per candidate per call   (k_head_simt: FP32 register-tiled GEMM, K = 256,
//                                                      fused bias/ReLU/second ParallelLinear/sigma division)
// State, stage combinations, error norm and dense output are float64 as in scipy; the network is float32.
#include "sampler_device.cuh"
#include "vpho_b200.h"

#include <cstdlib>
#include <cstring>
#include <vector>
#ifndef VPHO_EMU
#include <cuda_fp16.h>
#endif

namespace vpho {

// tcgen05 path (scorenet_tc.cu)
bool tc_available();
bool tc_make_map(void* map, const void* base, int rows, int box_rows, int kdim, bool half = false);
int tc_launch_pose(const TcPoseJob* jobs, int n_jobs, int mode, int s, cudaStream_t st);
int tc_launch_head(const TcHeadJob* jobs, int n_jobs, int mode, int s, int ctas, cudaStream_t st);

struct alignas(64) TensorMapBlob { unsigned char b[128]; };

struct DenoiserHost {
  DenoiserDev dev;
  float* blob = nullptr;
  // Exactly one numerical path per handle, fixed at creation (vpho_denoiser_create_ex):
  //   tensor-core (default): tcgen05 pose encoder (3xTF32 + 3xFP16) feeding the 3xFP16 head GEMM;
  //   strict FP32 (VPHO_DENOISER_STRICT_FP32, and the only path of the CPU emulator build): SIMT kernels.
  bool use_tc = false;
  void* w_half = nullptr;            // [2][hid][256] __half (hi, lo) head weights scaled per head + [n_heads] float un-scale factors
  TensorMapBlob mapBh_hi, mapBh_lo;   // 256-row boxes: single-CTA head GEMM (a stand-alone sampler with few heads)
  TensorMapBlob mapBp_hi, mapBp_lo;   // same planes, 128-row boxes: the CTA-pair kernel loads half a head per CTA
  int pair_min_heads = 8;            // CTA-pair head GEMM from this many heads up (and always for two samplers in lock-step)
  TensorMapBlob mapA_hi, mapA_lo;
  // tcgen05 pose encoder: K-major TF32 hi/lo planes of pose_encoder.0 [256][Kpad1], FP16 hi/lo planes of pose_encoder.2
  float* pe_planes = nullptr;
  TensorMapBlob mapW1_hi, mapW1_lo;
  void* w2_half = nullptr;           // [2][256][256] __half (hi, lo) planes of pose_encoder.2 scaled by a power of two
  TensorMapBlob mapW2h_hi, mapW2h_lo;
  const float* mapA_for = nullptr;   // P2hi pointer the cached A maps were built for
  int mapA_rows = 0;
};

// ------------------------------------------------------------------------------------------------------------
// controller set-up
// ------------------------------------------------------------------------------------------------------------
struct SampleCfg {
  double T0, eps, rtol, atol, max_step;
  int n_rows, D, n_eval, num_steps, rows_per_feat;
  const double* t_eval_in;   // device pointer or nullptr (= numpy.linspace(T0, eps, n_eval))
  double* xs;
  float* xs32;
  double* x_out;
  int32_t* counters;
};

__global__ void k_init_ctrl(SampleCfg cfg, SamplerWs ws) {
  RkCtrl& c = *ws.ctrl;
  const int tid = threadIdx.x;
  // numpy.linspace: step = (stop-start)/div ; y = arange(num)*step + start ; y[-1] = stop
  const int div = cfg.n_eval - 1;
  for (int i = tid; i < cfg.n_eval; i += blockDim.x) {
    double v;
    if (cfg.t_eval_in) v = cfg.t_eval_in[i];
    else if (div > 0) {
      double step = (cfg.eps - cfg.T0) / (double)div;
      v = __dadd_rn(__dmul_rn((double)i, step), cfg.T0);
      if (i == cfg.n_eval - 1) v = cfg.eps;
    } else v = cfg.T0;
    ws.t_eval[i] = v;
  }
  if (tid == 0) {
    c.T0 = cfg.T0; c.eps = cfg.eps; c.rtol = cfg.rtol; c.atol = cfg.atol; c.max_step = cfg.max_step;
    c.n = cfg.n_rows * cfg.D; c.n_rows = cfg.n_rows; c.D = cfg.D; c.n_eval = cfg.n_eval; c.num_steps = cfg.num_steps;
    c.rows_per_feat = cfg.rows_per_feat;
    c.direction = cfg.eps >= cfg.T0 ? 1.0 : -1.0;
    c.t = cfg.T0; c.h_abs = 0; c.h = 0; c.t_new = cfg.T0; c.h0 = 0; c.d0 = 0; c.d1 = 0;
    c.status = (c.n == 0 || cfg.T0 == cfg.eps) ? 1 : 0;
    c.step_rejected = 0; c.accepted_now = 0; c.t_old = cfg.T0; c.h_done = 0;
    c.te_next = 0; c.te_lo = 0; c.te_hi = 0;
    c.nfev = 0; c.n_acc = 0; c.n_rej = 0; c.attempts = 0; c.finished_final = 0; c.nan_seen = 0;
    for (int k = 0; k < 8; ++k) c.nan_stage[k] = 0;
    c.block_counter = 0u;
    c.eval_t32 = 0.f;
    c.xs = cfg.xs; c.xs32 = cfg.xs32; c.x_out = cfg.x_out; c.counters = cfg.counters;
    if (c.counters) for (int k = 0; k < 8; ++k) c.counters[k] = 0;
  }
}

__global__ void k_init_state(SamplerWs ws, const float* __restrict__ init_x, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) ws.y[i] = (double)init_x[i];
}

__global__ void k_set_eval_time(SamplerWs ws, float t32) {
  RkCtrl& c = *ws.ctrl;
  c.eval_t32 = t32;
  c.status = 0;
  c.nan_seen = 0;
  for (int k = 0; k < 8; ++k) c.nan_stage[k] = 0;
}

// ------------------------------------------------------------------------------------------------------------
// feat-term: F[r][col] = sum_k feat[r][k] Wa_f[k][col] + ba[col]       (once per sample(); R = images)
// ------------------------------------------------------------------------------------------------------------
constexpr int kFtRows = 32, kFtCols = 128, kFtK = 32, kFtSplit = 8;

// split-K: blockIdx.z owns K slice [z*128, z*128+128) and writes its partial sums to Fpart[z]; k_feat_sum adds the eight
// partials in a fixed order plus the bias (8 slices: the kernel is a chain of dependent global-load rounds, and 4 rounds
// per CTA instead of 8 nearly halve its latency).  FP32 FMA accumulation throughout (the tensor-core variant accumulates with
// truncation over a 384-instruction chain, which costs ~1 decimal digit on this K = 1024 contraction).
__device__ __forceinline__ void feat_term_block(const DenoiserDev& dn, const float* __restrict__ feat, int R, float* __restrict__ Fpart,
                                                int bx) {
  __shared__ __align__(16) float fs[kFtK][kFtRows];
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  const int c0 = bx * kFtCols, r0 = blockIdx.y * kFtRows;
  if (r0 >= R) return;
  const int hid = dn.hid;
  const int kbeg = blockIdx.z * (kFDim / kFtSplit), kend = kbeg + kFDim / kFtSplit;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = kbeg; k0 < kend; k0 += kFtK) {
    __syncthreads();
    for (int it = tid; it < kFtK * kFtRows; it += 256) {
      const int r = it / kFtK, k = it % kFtK;
      fs[k][r] = (r0 + r < R) ? feat[(size_t)(r0 + r) * kFDim + k0 + k] : 0.f;
    }
    __syncthreads();
#pragma unroll 16
    for (int k = 0; k < kFtK; ++k) {
      const float4 f = *reinterpret_cast<const float4*>(&fs[k][4 * ty]);
      const float4 w = __ldg(reinterpret_cast<const float4*>(dn.Wa_f + (size_t)(k0 + k) * hid + c0 + 4 * tx));
      const float fr[4] = {f.x, f.y, f.z, f.w}, wc[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(fr[i], wc[j], acc[i][j]);
    }
  }
  float* dst = Fpart + (size_t)blockIdx.z * R * hid;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + 4 * ty + i;
    if (r >= R) continue;
    *reinterpret_cast<float4*>(dst + (size_t)r * hid + c0 + 4 * tx) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
  }
}

// grid.x: column blocks of job 0 followed by those of job 1 (two samplers begun together; xb0 = gridDim.x for one job);
// grid.y covers the larger row count
__global__ void __launch_bounds__(256) k_feat_term(DenoiserDev dn0, const float* __restrict__ feat0, int R0, float* __restrict__ Fpart0,
                                                  DenoiserDev dn1, const float* __restrict__ feat1, int R1, float* __restrict__ Fpart1,
                                                  int xb0) {
  if ((int)blockIdx.x < xb0) feat_term_block(dn0, feat0, R0, Fpart0, blockIdx.x);
  else feat_term_block(dn1, feat1, R1, Fpart1, (int)blockIdx.x - xb0);
}

__global__ void k_feat_sum(DenoiserDev dn, const float* __restrict__ Fpart, int R, float* __restrict__ F) {
  static_assert(kFtSplit == 8, "the fixed summation tree below is written for 8 partials");
  const size_t n4 = (size_t)R * dn.hid / 4, stride = (size_t)R * dn.hid / 4;
  const float4* p = reinterpret_cast<const float4*>(Fpart);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 q[kFtSplit];
#pragma unroll
    for (int z = 0; z < kFtSplit; ++z) q[z] = p[i + z * stride];
    const float4 bias = __ldg(reinterpret_cast<const float4*>(dn.ba) + (i % (dn.hid / 4)));
    float4 o;          // fixed tree: ((p0 + p1) + (p2 + p3)) + ((p4 + p5) + (p6 + p7)), then the bias
    o.x = (((q[0].x + q[1].x) + (q[2].x + q[3].x)) + ((q[4].x + q[5].x) + (q[6].x + q[7].x))) + bias.x;
    o.y = (((q[0].y + q[1].y) + (q[2].y + q[3].y)) + ((q[4].y + q[5].y) + (q[6].y + q[7].y))) + bias.y;
    o.z = (((q[0].z + q[1].z) + (q[2].z + q[3].z)) + ((q[4].z + q[5].z) + (q[6].z + q[7].z))) + bias.z;
    o.w = (((q[0].w + q[1].w) + (q[2].w + q[3].w)) + ((q[4].w + q[5].w) + (q[6].w + q[7].w))) + bias.w;
    reinterpret_cast<float4*>(F)[i] = o;
  }
}

__global__ void __launch_bounds__(256) k_time_term(DenoiserDev dn, SamplerWs ws, int mode, int s) {
  const RkCtrl& c = *ws.ctrl;
  if (!eval_active(c, mode)) return;
  time_term_block(dn, ws, c, mode, s, blockIdx.x);
}

// tcgen05 path: the time terms of a network call.  The times of all six stages of an RK attempt are known once the attempt
// has begun (t + c_s h), so the first stage's launch computes all six (six groups of column blocks) and stages 2..6 launch
// nothing: their critical path is the fused stage-input + pose-encoder kernel alone (k_pose_tc).
__device__ __forceinline__ void time_terms_block(const DenoiserDev& dn, const SamplerWs& ws, int mode, int s, int bid) {
  const RkCtrl& c = *ws.ctrl;
  if (!eval_active(c, mode)) return;
  const int per = (dn.hid + kTtCols - 1) / kTtCols;
  if (mode == kModeStage) time_term_block(dn, ws, c, mode, 1 + bid / per, bid % per);
  else time_term_block(dn, ws, c, mode, s, bid);
}

// blocks [0, blocks0) work for job 0, the rest for job 1 (two samplers in lock-step; blocks0 = gridDim.x for one)
__global__ void __launch_bounds__(256) k_time_terms(DenoiserDev dn0, SamplerWs ws0, DenoiserDev dn1, SamplerWs ws1, int blocks0, int mode,
                                                   int s) {
  pdl_wait();                 // launched with launch_pdl: nothing of the previous kernel may be read above this line
  pdl_trigger();
  if ((int)blockIdx.x < blocks0) time_terms_block(dn0, ws0, mode, s, blockIdx.x);
  else time_terms_block(dn1, ws1, mode, s, (int)blockIdx.x - blocks0);
}

// ------------------------------------------------------------------------------------------------------------
// stage input (float64 RK combination -> float32) + pose encoder (D -> 256 -> 256, ReLU) -> P2T (k-major)
// ------------------------------------------------------------------------------------------------------------
constexpr int kPeRows = 32;
constexpr int kMaxD = 96;

__global__ void __launch_bounds__(256) k_pose_encoder(DenoiserDev dn, SamplerWs ws, int mode, int s) {
  const RkCtrl& c = *ws.ctrl;
  if (!eval_active(c, mode)) return;
  __shared__ __align__(16) float xs[kMaxD][kPeRows];
  __shared__ __align__(16) float h1s[kPDim][kPeRows];
  const int tid = threadIdx.x, tx = tid & 7, ty = tid >> 3;
  const int D = dn.D, n_rows = (mode == kModeEval) ? ws.eval_rows : c.n_rows;
  const int r0 = blockIdx.x * kPeRows;
  for (int it = tid; it < kPeRows * D; it += 256) {
    const int r = it / D, d = it - r * D;
    const double v = stage_input(ws, c, mode, s, r0 + r, d, n_rows, D);
    xs[d][r] = (float)v;
  }
  __syncthreads();
  float acc[4][8];
  {
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(dn.b1 + 8 * ty));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(dn.b1 + 8 * ty + 4));
    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = bb[j];
  }
  for (int k = 0; k < D; ++k) {
    const float4 x4 = *reinterpret_cast<const float4*>(&xs[k][4 * tx]);
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(dn.W1 + (size_t)k * kPDim + 8 * ty));
    const float4 w1 = __ldg(reinterpret_cast<const float4*>(dn.W1 + (size_t)k * kPDim + 8 * ty + 4));
    const float xr[4] = {x4.x, x4.y, x4.z, x4.w};
    const float wc[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(xr[i], wc[j], acc[i][j]);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float4 o = make_float4(fmaxf(acc[0][j], 0.f), fmaxf(acc[1][j], 0.f), fmaxf(acc[2][j], 0.f), fmaxf(acc[3][j], 0.f));
    *reinterpret_cast<float4*>(&h1s[8 * ty + j][4 * tx]) = o;
  }
  __syncthreads();
  {
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(dn.b2 + 8 * ty));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(dn.b2 + 8 * ty + 4));
    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = bb[j];
  }
#pragma unroll 4
  for (int k = 0; k < kPDim; ++k) {
    const float4 x4 = *reinterpret_cast<const float4*>(&h1s[k][4 * tx]);
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(dn.W2 + (size_t)k * kPDim + 8 * ty));
    const float4 w1 = __ldg(reinterpret_cast<const float4*>(dn.W2 + (size_t)k * kPDim + 8 * ty + 4));
    const float xr[4] = {x4.x, x4.y, x4.z, x4.w};
    const float wc[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(xr[i], wc[j], acc[i][j]);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float4 o = make_float4(fmaxf(acc[0][j], 0.f), fmaxf(acc[1][j], 0.f), fmaxf(acc[2][j], 0.f), fmaxf(acc[3][j], 0.f));
    *reinterpret_cast<float4*>(ws.P2T + (size_t)(8 * ty + j) * ws.Npad + r0 + 4 * tx) = o;
  }
}

// ------------------------------------------------------------------------------------------------------------
// head GEMM, FP32 SIMT: one CTA = 128 candidate rows x one ParallelLinear head (256 hidden units, two 128-wide
// passes), K = 256 pose features streamed with cp.async double buffering; epilogue fuses + F[img] + Tt, ReLU,
// the 256 -> 3 second ParallelLinear, its bias and the division by sigma(t).
// ------------------------------------------------------------------------------------------------------------
constexpr int kHgBK = 16;

__global__ void __launch_bounds__(256) k_head_simt(DenoiserDev dn, SamplerWs ws, int mode, int s) {
  RkCtrl& c = *ws.ctrl;
  if (!eval_active(c, mode)) return;
  __shared__ __align__(16) float As[2][kHgBK][128];
  __shared__ __align__(16) float Bs[2][kHgBK][128];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int head = blockIdx.y, r0 = blockIdx.x * kRowTile;
  const int hid = dn.hid, Npad = ws.Npad;
  const int n_rows = (mode == kModeEval) ? ws.eval_rows : c.n_rows;
  const int rpf = (mode == kModeEval) ? ws.eval_rpf : c.rows_per_feat;
  const EvalTime et = c.et[tt_slot(mode, s)];

  int rows[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) rows[i] = r0 + (i < 4 ? 4 * ty + i : 64 + 4 * ty + (i - 4));
  float o[8][3];
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i][0] = o[i][1] = o[i][2] = 0.f;

  for (int pass = 0; pass < 2; ++pass) {
    const int c0 = head * kHeadHid + pass * 128;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    auto load_tiles = [&](int buf, int k0) {
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int it = tid + q * 256;          // 512 float4 per operand tile
        const int k = it >> 5, x4 = (it & 31) * 4;
        cp_async16(&As[buf][k][x4], ws.P2T + (size_t)(k0 + k) * Npad + r0 + x4);
        cp_async16(&Bs[buf][k][x4], dn.Wa_p + (size_t)(k0 + k) * hid + c0 + x4);
      }
      cp_async_commit();
    };
    __syncthreads();
    load_tiles(0, 0);
    for (int kt = 0; kt < kPDim / kHgBK; ++kt) {
      const int buf = kt & 1;
      if (kt + 1 < kPDim / kHgBK) { load_tiles(buf ^ 1, (kt + 1) * kHgBK); cp_async_wait<1>(); }
      else cp_async_wait<0>();
      __syncthreads();
#pragma unroll
      for (int k = 0; k < kHgBK; ++k) {
        const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][4 * ty]);
        const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + 4 * ty]);
        const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][4 * tx]);
        const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + 4 * tx]);
        const float ar[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float bc[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(ar[i], bc[j], acc[i][j]);
      }
      __syncthreads();
    }
    // epilogue of this pass: hidden = relu(acc + F[img] + Tt); o += hidden * Wb
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) {
      const int cb = c0 + jj * 64 + 4 * tx;
      const float4 tt = *reinterpret_cast<const float4*>(ws.Tt + (size_t)tt_slot(mode, s) * dn.hid + cb);
      const float tta[4] = {tt.x, tt.y, tt.z, tt.w};
      float4 wb[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) wb[j] = __ldg(reinterpret_cast<const float4*>(dn.Wb + (size_t)(cb + j) * 4));
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (rows[i] >= n_rows) continue;
        const float4 f4 = __ldg(reinterpret_cast<const float4*>(ws.F + (size_t)(rows[i] / rpf) * hid + cb));
        const float fa[4] = {f4.x, f4.y, f4.z, f4.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float hval = (acc[i][jj * 4 + j] + fa[j]) + tta[j];
          hval = hval > 0.f ? hval : 0.f;
          o[i][0] = fmaf(hval, wb[j].x, o[i][0]);
          o[i][1] = fmaf(hval, wb[j].y, o[i][1]);
          o[i][2] = fmaf(hval, wb[j].z, o[i][2]);
        }
      }
    }
  }
  // reduce the 16 column-threads of each row (lanes tx of a half-warp), fixed butterfly order
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      float v = o[i][d];
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      o[i][d] = v;
    }
  if (tx == 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (rows[i] >= n_rows) continue;
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const float out = o[i][d] + dn.bb[head * 3 + d];
        const float score = __fdiv_rn(out, et.std32);
        emit_score(ws, c, et, mode, s, rows[i] * dn.D + head * 3 + d, score);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// reductions + the step controller (one thread of the last block to finish)
// ------------------------------------------------------------------------------------------------------------
enum RedKind : int { kRedInit0 = 0, kRedInit1 = 1, kRedErr = 2 };

__device__ void export_counters(const SamplerWs& ws, const RkCtrl& c) {
  if (!c.counters) return;
  c.counters[0] = c.status; c.counters[1] = c.nfev; c.counters[2] = c.n_acc; c.counters[3] = c.n_rej;
  c.counters[4] = c.nan_seen; c.counters[5] = c.attempts;
}

// scipy RungeKutta._step_impl: head of one `while not step_accepted` iteration
__device__ void begin_attempt(RkCtrl& c, bool new_step) {
  const double min_step = 10.0 * fabs(nextafter(c.t, c.direction * (double)INFINITY) - c.t);
  if (new_step) {
    if (c.h_abs > c.max_step) c.h_abs = c.max_step;
    else if (c.h_abs < min_step) c.h_abs = min_step;
    c.step_rejected = 0;
  }
  if (c.h_abs < min_step) { c.status = -1; return; }
  double h = c.h_abs * c.direction;
  double t_new = c.t + h;
  if (c.direction * (t_new - c.eps) > 0) t_new = c.eps;
  h = t_new - c.t;
  c.h_abs = fabs(h);
  c.h = h;
  c.t_new = t_new;
}

__device__ void controller(const SamplerWs& ws, RkCtrl& c, int kind, double s0, double s1, int te_count) {
  const double sqrt_n = sqrt((double)c.n);     // x.size ** 0.5
  if (kind == kRedInit0) {
    c.nfev = 1;
    const double d0 = sqrt(s0) / sqrt_n, d1 = sqrt(s1) / sqrt_n;
    double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
    const double interval = fabs(c.eps - c.T0);
    h0 = fmin(h0, interval);
    c.h0 = h0; c.d0 = d0; c.d1 = d1;
  } else if (kind == kRedInit1) {
    c.nfev = 2;
    const double d2 = (sqrt(s0) / sqrt_n) / c.h0;
    double h1;
    if (c.d1 <= 1e-15 && d2 <= 1e-15) h1 = fmax(1e-6, c.h0 * 1e-3);
    else h1 = pow(0.01 / fmax(c.d1, d2), 1.0 / 5.0);           // order + 1 = 5
    const double interval = fabs(c.eps - c.T0);
    c.h_abs = fmin(fmin(100.0 * c.h0, h1), fmin(interval, c.max_step));
    begin_attempt(c, true);
  } else {
    const double err = sqrt(s0) / sqrt_n;
    c.attempts += 1;
    c.nfev += 6;
    if (err < 1.0) {
      double factor = (err == 0.0) ? kMaxFactor : fmin(kMaxFactor, kSafety * pow(err, kErrExponent));
      if (c.step_rejected) factor = fmin(1.0, factor);
      c.h_abs *= factor;
      c.t_old = c.t; c.h_done = c.h;
      c.t = c.t_new;
      c.n_acc += 1;
      c.accepted_now = 1;
      // ivp.py: searchsorted(side='left') on reversed t_eval.  `te_count` = number of output times from te_next on that are
      // >= t_new, counted by the whole block beforehand (t_eval is sorted): a serial scan here is a chain of up to 17
      // dependent global loads per accepted step
      int hi = c.te_next + te_count;
      c.te_lo = c.te_next; c.te_hi = hi; c.te_next = hi;
      if (c.direction * (c.t - c.eps) >= 0) c.status = 1;
      else begin_attempt(c, true);
    } else {
      const double f = kSafety * pow(err, kErrExponent);
      c.h_abs *= (f > kMinFactor ? f : kMinFactor);            // python max(MIN_FACTOR, x): NaN -> MIN_FACTOR
      c.step_rejected = 1;
      c.n_rej += 1;
      c.accepted_now = 0;
      for (int k = 1; k < 7; ++k) c.nan_stage[k] = 0;
      begin_attempt(c, false);
    }
  }
  export_counters(ws, c);
}

__device__ __forceinline__ void reduce_block(const SamplerWs& ws, int kind, int bid, int n_blocks) {
  RkCtrl& c = *ws.ctrl;
  if (c.status != 0) return;
  __shared__ double sh0[256], sh1[256];
  __shared__ bool is_last;
  const int tid = threadIdx.x, n = c.n;
  double a0 = 0.0, a1 = 0.0;
  for (int i = bid * 256 + tid; i < n; i += n_blocks * 256) {
    const double y = ws.y[i];
    if (kind == kRedInit0) {
      const double scale = c.atol + fabs(y) * c.rtol;
      const double u = y / scale, v = kval(ws.K, c, 0, n, i) / scale;
      a0 += u * u; a1 += v * v;
    } else if (kind == kRedInit1) {
      const double scale = c.atol + fabs(y) * c.rtol;
      const double u = (kval(ws.K, c, 1, n, i) - kval(ws.K, c, 0, n, i)) / scale;
      a0 += u * u;
    } else {
      const double scale = c.atol + fmax(fabs(y), fabs(ws.ynew[i])) * c.rtol;
      double e = 0.0;
#pragma unroll
      for (int j = 0; j < 7; ++j) e += kval(ws.K, c, j, n, i) * kE[j];
      const double u = (e * c.h) / scale;
      a0 += u * u;
    }
  }
  sh0[tid] = a0; sh1[tid] = a1;
  __syncthreads();
  for (int st = 128; st > 0; st >>= 1) {
    if (tid < st) { sh0[tid] += sh0[tid + st]; sh1[tid] += sh1[tid + st]; }
    __syncthreads();
  }
  if (tid == 0) {
    ws.partial[bid] = sh0[0];
    ws.partial[kMaxRedBlocks + bid] = sh1[0];
    __threadfence();
    const unsigned prev = atomicAdd(&c.block_counter, 1u);
    is_last = (prev == (unsigned)n_blocks - 1u);
  }
  __syncthreads();
  if (is_last) {
    // the last block to finish sums the per-block partials in a fixed order (strided per thread, then the same tree)
    __threadfence();
    double p0 = 0.0, p1 = 0.0;
    for (int b = tid; b < n_blocks; b += 256) { p0 += ws.partial[b]; p1 += ws.partial[kMaxRedBlocks + b]; }
    sh0[tid] = p0; sh1[tid] = p1;
    __syncthreads();
    for (int st = 128; st > 0; st >>= 1) {
      if (tid < st) { sh0[tid] += sh0[tid + st]; sh1[tid] += sh1[tid + st]; }
      __syncthreads();
    }
    // output times inside the step that is about to be judged (used only if it is accepted)
    __shared__ int s_cnt;
    if (tid == 0) s_cnt = 0;
    __syncthreads();
    if (kind == kRedErr) {
      int mine = 0;
      for (int e = c.te_next + tid; e < c.n_eval; e += 256) mine += (ws.t_eval[e] >= c.t_new) ? 1 : 0;
      if (mine) atomicAdd(&s_cnt, mine);
    }
    __syncthreads();
    if (tid == 0) {
      c.block_counter = 0u;
      controller(ws, c, kind, sh0[0], sh1[0], s_cnt);
    }
  }
}

__global__ void __launch_bounds__(256) k_reduce(SamplerWs ws0, SamplerWs ws1, int blocks0, int kind) {
  pdl_wait();
  pdl_trigger();
  if ((int)blockIdx.x < blocks0) reduce_block(ws0, kind, blockIdx.x, blocks0);
  else reduce_block(ws1, kind, (int)blockIdx.x - blocks0, (int)gridDim.x - blocks0);
}

// after an accepted step: quartic dense output at the t_eval points inside the step, then roll y <- y_new, f <- f_new
__device__ __forceinline__ void post_step_block(const SamplerWs& ws, int bid, int n_blocks) {
  RkCtrl& c = *ws.ctrl;
  if (c.status < 0 || !c.accepted_now) return;
  const int n = c.n;
  const int lo = c.te_lo, hi = c.te_hi;
  const double h = c.h_done, t_old = c.t_old;
  // The powers of x = (t - t_old) / h depend on the output point only: one thread per point forms them (np.cumprod order)
  // in shared memory, so that the element loop does no float64 division -- the FP64 pipe is what bounds this kernel.
  constexpr int kPwTile = 256;
  __shared__ double s_pw[kPwTile][4];
  const bool emit = (c.xs || c.xs32) && hi > lo;
  for (int e0 = lo; emit && e0 < hi; e0 += kPwTile) {
    const int ne = hi - e0 < kPwTile ? hi - e0 : kPwTile;
    __syncthreads();
    if ((int)threadIdx.x < ne) {
      const double x = (ws.t_eval[e0 + threadIdx.x] - t_old) / h;
      const double p1 = x, p2 = p1 * x, p3 = p2 * x, p4 = p3 * x;      // np.cumprod
      s_pw[threadIdx.x][0] = p1; s_pw[threadIdx.x][1] = p2; s_pw[threadIdx.x][2] = p3; s_pw[threadIdx.x][3] = p4;
    }
    __syncthreads();
    for (int i = bid * 256 + threadIdx.x; i < n; i += n_blocks * 256) {
      const double y_old = ws.y[i];
      double Q[4] = {0, 0, 0, 0};
#pragma unroll
      for (int j = 0; j < 7; ++j) {
        const double k = kval(ws.K, c, j, n, i);
#pragma unroll
        for (int q = 0; q < 4; ++q) Q[q] += k * kP[j][q];
      }
      for (int e = 0; e < ne; ++e) {
        const double p1 = s_pw[e][0], p2 = s_pw[e][1], p3 = s_pw[e][2], p4 = s_pw[e][3];
        const double dot = ((Q[0] * p1 + Q[1] * p2) + Q[2] * p3) + Q[3] * p4;
        const double v = __dadd_rn(__dmul_rn(h, dot), y_old);
        if (c.xs) c.xs[(size_t)(e0 + e) * n + i] = v;
        if (c.xs32) c.xs32[(size_t)(e0 + e) * n + i] = (float)v;          // the `.float()` the predict branch applies (VPHO.py:243)
      }
    }
  }
  for (int i = bid * 256 + threadIdx.x; i < n; i += n_blocks * 256) {
    ws.y[i] = ws.ynew[i];
    // f <- f_new (first-same-as-last): slot 0 takes slot 6's scores (its coefficient follows once every block is done);
    // a value nan_to_num would have zeroed is stored as -0 so that it reads back as the +0 kval returns for it
    float r = ws.K[(size_t)6 * n + i];
    if (c.nan_stage[6] && !isfinite(r)) r = -0.f;
    ws.K[i] = r;
  }
  __shared__ bool is_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned prev = atomicAdd(&c.block_counter, 1u);
    is_last = (prev == (unsigned)n_blocks - 1u);
    if (is_last) {
      c.block_counter = 0u;
      c.accepted_now = 0;
      c.kcoef[0] = c.kcoef[6];
      for (int k = 0; k < 7; ++k) c.nan_stage[k] = 0;
    }
  }
}

__global__ void __launch_bounds__(256) k_post_step(SamplerWs ws0, SamplerWs ws1, int blocks0) {
  pdl_wait();
  pdl_trigger();
  if ((int)blockIdx.x < blocks0) post_step_block(ws0, blockIdx.x, blocks0);
  else post_step_block(ws1, (int)blockIdx.x - blocks0, (int)gridDim.x - blocks0);
}

__global__ void k_export(SamplerWs ws) { export_counters(ws, *ws.ctrl); }

__global__ void k_rot6d_to_aa(const float* __restrict__ x6d, int n, float* __restrict__ aa) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float d[6], R[9], a[3];
#pragma unroll
  for (int k = 0; k < 6; ++k) d[k] = x6d[(size_t)i * 6 + k];
  rot6d_to_matrix(d, R);
  matrix_to_axis_angle(R, a);
  aa[(size_t)i * 3 + 0] = a[0]; aa[(size_t)i * 3 + 1] = a[1]; aa[(size_t)i * 3 + 2] = a[2];
}

// Fused `postprocess_diffusion_hand` (VPHO.py:306-331): float64 sampler output in storage order [n_steps][n_rows][96]
// -> float32 MANO vectors [n_rows][n_steps][58] = axis-angle of the 16 joints (the float64 -> float32 rounding of
// `.float()`, then the same 6D -> matrix -> axis-angle arithmetic as k_rot6d_to_aa) followed by the image's 10 shape
// coefficients.  One thread per (step, row, slot): slots 0..15 are joints, slot 16 copies the shape.
template <typename TIn>
__global__ void __launch_bounds__(256) k_postprocess_hand(const TIn* __restrict__ xs, int n_steps, int n_rows, int rows_per_shape,
                                                         const float* __restrict__ shape, float* __restrict__ out) {
  const size_t total = (size_t)n_steps * n_rows * 17;
  for (size_t it = (size_t)blockIdx.x * blockDim.x + threadIdx.x; it < total; it += (size_t)gridDim.x * blockDim.x) {
    const int slot = (int)(it % 17);
    const size_t sr = it / 17;
    const int row = (int)(sr % n_rows), step = (int)(sr / n_rows);
    float* dst = out + ((size_t)row * n_steps + step) * 58;
    if (slot < 16) {
      const TIn* src = xs + sr * 96 + slot * 6;
      float d[6], R[9], a[3];
#pragma unroll
      for (int k = 0; k < 6; ++k) d[k] = (float)src[k];
      rot6d_to_matrix(d, R);
      matrix_to_axis_angle(R, a);
      dst[slot * 3 + 0] = a[0]; dst[slot * 3 + 1] = a[1]; dst[slot * 3 + 2] = a[2];
    } else {
      const float* sp = shape + (size_t)(row / rows_per_shape) * 10;
#pragma unroll
      for (int k = 0; k < 10; ++k) dst[48 + k] = sp[k];
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------
static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static size_t carve(void* base, int n_heads, int n_rows, int rows_per_feat, int n_eval, SamplerWs* ws, bool use_tc = true) {
  const int D = 3 * n_heads, hid = n_heads * kHeadHid;
  const int R = (n_rows + rows_per_feat - 1) / rows_per_feat;
  const int Npad = (int)align_up((size_t)(n_rows > 0 ? n_rows : 1), kRowTile);
  const size_t n = (size_t)n_rows * D;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
  const size_t o_ctrl = take(sizeof(RkCtrl));
  const size_t o_F = take((size_t)R * hid * 4);
  const size_t o_Fpart = take((size_t)8 * R * hid * 4);      // kFtSplit partial sums
  const size_t o_Tt = take((size_t)7 * hid * 4);
  // pose features: [256][Npad] float (strict-FP32 SIMT path) or two [Npad][256] __half planes + per-row scales (tensor cores);
  // the workspace size does not depend on the path (same bytes either way)
  const size_t o_P2T = take((size_t)kPDim * Npad * 4);
  const size_t o_P2sc = take((size_t)Npad * 4);
  const size_t o_y = take(n * 8);
  const size_t o_yn = take(n * 8);
  const size_t o_K = take(7 * n * 4);
  const size_t o_part = take((size_t)2 * kMaxRedBlocks * 8);
  const size_t o_te = take((size_t)(n_eval > 0 ? n_eval : 1) * 8);
  if (ws) {
    char* b = static_cast<char*>(base);
    ws->ctrl = reinterpret_cast<RkCtrl*>(b + o_ctrl);
    ws->F = reinterpret_cast<float*>(b + o_F);
    ws->Fpart = reinterpret_cast<float*>(b + o_Fpart);
    ws->Tt = reinterpret_cast<float*>(b + o_Tt);
    ws->P2T = use_tc ? nullptr : reinterpret_cast<float*>(b + o_P2T);
    ws->P2hi = use_tc ? reinterpret_cast<float*>(b + o_P2T) : nullptr;
    ws->P2lo = use_tc ? reinterpret_cast<float*>(b + o_P2T + (size_t)kPDim * Npad * 2) : nullptr;
    ws->P2scale = use_tc ? reinterpret_cast<float*>(b + o_P2sc) : nullptr;
    ws->y = reinterpret_cast<double*>(b + o_y);
    ws->ynew = reinterpret_cast<double*>(b + o_yn);
    ws->K = reinterpret_cast<float*>(b + o_K);
    ws->partial = reinterpret_cast<double*>(b + o_part);
    ws->t_eval = reinterpret_cast<double*>(b + o_te);
    ws->eval_out = nullptr; ws->eval_x = nullptr;
    ws->eval_rows = n_rows; ws->eval_rpf = rows_per_feat;
    ws->Npad = Npad; ws->R = R;
  }
  return off;
}

static int red_blocks(int n) { int b = (n + 255) / 256; return b < 1 ? 1 : (b > kMaxRedBlocks ? kMaxRedBlocks : b); }

static int launch_feat_term(DenoiserHost& dh, const SamplerWs& ws, const float* feat, cudaStream_t st) {
  const DenoiserDev& dn = dh.dev;
  profile_begin(VPHO_TAG_FEAT_TERM, st);
  VPHO_LAUNCH(k_feat_term, dim3(dn.hid / kFtCols, (ws.R + kFtRows - 1) / kFtRows, kFtSplit), dim3(256), 0, st, dn, feat, ws.R,
              ws.Fpart, dn, feat, ws.R, ws.Fpart, dn.hid / kFtCols);
  const int n4 = ws.R * dn.hid / 4;
  VPHO_LAUNCH(k_feat_sum, dim3((n4 + 255) / 256 < 1184 ? (n4 + 255) / 256 : 1184), dim3(256), 0, st, dn, ws.Fpart, ws.R, ws.F);
  profile_end(VPHO_TAG_FEAT_TERM, st);
  VPHO_CHECK_LAUNCH();
  return VPHO_OK;
}

// One sampler's state on the host side of a launch; two of them advance in lock-step through the same kernel launches.
struct SamplerJob { DenoiserHost* dh; SamplerWs ws; int ws_n; /* n_rows * D */ };

// column blocks of the time-term kernel for this call (0: nothing to launch)
static int time_blocks(const DenoiserDev& dn, int mode, int s) {
  const int per = (dn.hid + kTtCols - 1) / kTtCols;
  return mode == kModeStage ? (s == 1 ? 6 * per : 0) : per;      // see k_time_terms
}

// One network evaluation of every job: stage input + time term, pose encoder, head GEMM.
static int launch_eval(SamplerJob* jobs, int n_jobs, int mode, int s, cudaStream_t st) {
#ifndef VPHO_EMU
  bool any_tc = false, all_tc = true;
  for (int j = 0; j < n_jobs; ++j) { any_tc = any_tc || jobs[j].dh->use_tc; all_tc = all_tc && jobs[j].dh->use_tc; }
  if (any_tc && !all_tc) {         // a tensor-core and a strict-FP32 handle do not share launches
    for (int j = 0; j < n_jobs; ++j) {
      int rc = launch_eval(jobs + j, 1, mode, s, st);
      if (rc) return rc;
    }
    return VPHO_OK;
  }
  if (all_tc) {
    SamplerJob& j0 = jobs[0];
    SamplerJob& j1 = jobs[n_jobs - 1];
    profile_begin(VPHO_TAG_POSE_ENCODER, st);
    const int b0 = time_blocks(j0.dh->dev, mode, s), b1 = n_jobs > 1 ? time_blocks(j1.dh->dev, mode, s) : 0;
    if (b0 + b1 > 0) {
      profile_begin(VPHO_TAG_STAGE_X, st);
      VPHO_LAUNCH_PDL(k_time_terms, dim3(b0 + b1), dim3(256), 0, st, j0.dh->dev, j0.ws, j1.dh->dev, j1.ws, b0, mode, s);
      profile_end(VPHO_TAG_STAGE_X, st);
    }
    TcPoseJob pj[2];
    TcHeadJob hj[2];
    bool pair = n_jobs > 1;
    for (int j = 0; j < n_jobs; ++j) {
      DenoiserHost& dh = *jobs[j].dh;
      const SamplerWs& ws = jobs[j].ws;
      if (dh.mapA_for != ws.P2hi || dh.mapA_rows != ws.Npad) {
        if (!tc_make_map(&dh.mapA_hi, ws.P2hi, ws.Npad, 128, kPDim, true) || !tc_make_map(&dh.mapA_lo, ws.P2lo, ws.Npad, 128, kPDim, true))
          return VPHO_ERR_LAUNCH;
        dh.mapA_for = ws.P2hi;
        dh.mapA_rows = ws.Npad;
      }
      pair = pair || dh.dev.n_heads >= dh.pair_min_heads;
      pj[j] = TcPoseJob{&dh.mapW1_hi, &dh.mapW1_lo, &dh.mapW2h_hi, &dh.mapW2h_lo, &dh.dev, &jobs[j].ws};
    }
    for (int j = 0; j < n_jobs; ++j) {
      DenoiserHost& dh = *jobs[j].dh;
      hj[j] = pair ? TcHeadJob{&dh.mapA_hi, &dh.mapA_lo, &dh.mapBp_hi, &dh.mapBp_lo, &dh.dev, &jobs[j].ws}
                   : TcHeadJob{&dh.mapA_hi, &dh.mapA_lo, &dh.mapBh_hi, &dh.mapBh_lo, &dh.dev, &jobs[j].ws};
    }
    int rc = tc_launch_pose(pj, n_jobs, mode, s, st);
    if (rc) return rc;
    profile_end(VPHO_TAG_POSE_ENCODER, st);
    const int tag = j0.dh->dev.n_heads >= 16 ? VPHO_TAG_HEAD_GEMM_HAND : VPHO_TAG_HEAD_GEMM_OBJ;
    profile_begin(tag, st);
    rc = tc_launch_head(hj, n_jobs, mode, s, pair ? 2 : 1, st);
    if (rc) return rc;
    profile_end(tag, st);
    return VPHO_OK;
  }
#endif
  // strict-FP32 SIMT path (cross-check of the tensor-core kernels; the CPU emulator build)
  for (int j = 0; j < n_jobs; ++j) {
    DenoiserHost& dh = *jobs[j].dh;
    const DenoiserDev& dn = dh.dev;
    const SamplerWs& ws = jobs[j].ws;
    profile_begin(VPHO_TAG_POSE_ENCODER, st);
    VPHO_LAUNCH(k_time_term, dim3((dn.hid + kTtCols - 1) / kTtCols), dim3(256), 0, st, dn, ws, mode, s);
    VPHO_LAUNCH(k_pose_encoder, dim3(ws.Npad / kPeRows), dim3(256), 0, st, dn, ws, mode, s);
    profile_end(VPHO_TAG_POSE_ENCODER, st);
    const int tag = dn.n_heads >= 16 ? VPHO_TAG_HEAD_GEMM_HAND : VPHO_TAG_HEAD_GEMM_OBJ;
    profile_begin(tag, st);
    VPHO_LAUNCH(k_head_simt, dim3(ws.Npad / kRowTile, dn.n_heads), dim3(256), 0, st, dn, ws, mode, s);
    profile_end(tag, st);
    VPHO_CHECK_LAUNCH();
  }
  return VPHO_OK;
}

static int launch_reduce(SamplerJob* jobs, int n_jobs, int kind, cudaStream_t st) {
  const SamplerWs& w0 = jobs[0].ws;
  const SamplerWs& w1 = jobs[n_jobs - 1].ws;
  const int b0 = red_blocks(jobs[0].ws_n), b1 = n_jobs > 1 ? red_blocks(jobs[1].ws_n) : 0;
  VPHO_LAUNCH_PDL(k_reduce, dim3(b0 + b1), dim3(256), 0, st, w0, w1, b0, kind);
  return VPHO_OK;
}

static int launch_attempts(SamplerJob* jobs, int n_jobs, int max_attempts, cudaStream_t st) {
  const SamplerWs& w0 = jobs[0].ws;
  const SamplerWs& w1 = jobs[n_jobs - 1].ws;
  const int b0 = red_blocks(jobs[0].ws_n), b1 = n_jobs > 1 ? red_blocks(jobs[1].ws_n) : 0;
  for (int a = 0; a < max_attempts; ++a) {
    for (int s = 1; s <= 6; ++s) {
      int rc = launch_eval(jobs, n_jobs, kModeStage, s, st);
      if (rc) return rc;
    }
    profile_begin(VPHO_TAG_RK_CONTROL, st);
    VPHO_LAUNCH_PDL(k_reduce, dim3(b0 + b1), dim3(256), 0, st, w0, w1, b0, (int)kRedErr);
    VPHO_LAUNCH_PDL(k_post_step, dim3(b0 + b1), dim3(256), 0, st, w0, w1, b0);
    profile_end(VPHO_TAG_RK_CONTROL, st);
    VPHO_CHECK_LAUNCH();
  }
  return VPHO_OK;
}

}  // namespace vpho

using namespace vpho;

static void free_denoiser(DenoiserHost* dh) {
  if (!dh) return;
  if (dh->blob) cudaFree(dh->blob);
  if (dh->pe_planes) cudaFree(dh->pe_planes);
  if (dh->w2_half) cudaFree(dh->w2_half);
  if (dh->w_half) cudaFree(dh->w_half);
  delete dh;
}

#ifndef VPHO_EMU
// power-of-two scale that brings `mx` into [2^13, 2^14) (exact in FP16 hi/lo splitting) and its inverse
static void pow2_scale(float mx, float* sc, float* inv) {
  *sc = 1.f; *inv = 1.f;
  if (mx > 0.f && mx < 3.0e38f) {
    int e = 0;
    frexpf(mx, &e);
    *sc = ldexpf(1.f, 14 - e);
    *inv = ldexpf(1.f, e - 14);
  }
}

// Operand planes + TMA descriptors of the tensor-core path.  Every resource is REQUIRED: a failure is reported to the
// caller (no silent change of numerical path).
static int build_tc_planes(DenoiserHost* dh, int n_heads, const float* p1_w, const float* p2_w, const float* ha_w) {
  DenoiserDev& d = dh->dev;
  const int D = d.D, hid = d.hid;
  if (!tc_available()) return VPHO_ERR_LAUNCH;       // driver without cuTensorMapEncodeTiled
  // pose_encoder.0: TF32 (hi, lo) planes [256][Kpad1]
  const int kp1 = (D + 31) / 32 * 32;
  const size_t n1 = (size_t)256 * kp1;
  std::vector<float> pl(2 * n1, 0.f);
  for (int o = 0; o < 256; ++o)
    for (int k = 0; k < D; ++k) {
      const float w = p1_w[(size_t)o * D + k], hi = tf32_round(w);
      pl[(size_t)o * kp1 + k] = hi;
      pl[n1 + (size_t)o * kp1 + k] = tf32_round(w - hi);
    }
  if (cudaMalloc((void**)&dh->pe_planes, pl.size() * sizeof(float)) != cudaSuccess) return VPHO_ERR_ALLOC;
  if (cudaMemcpy(dh->pe_planes, pl.data(), pl.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) return VPHO_ERR_ALLOC;
  if (!tc_make_map(&dh->mapW1_hi, dh->pe_planes, 256, 256, kp1) || !tc_make_map(&dh->mapW1_lo, dh->pe_planes + n1, 256, 256, kp1))
    return VPHO_ERR_LAUNCH;
  // pose_encoder.2: FP16 (hi, lo) planes [256][256], one power-of-two scale for the matrix
  {
    float mx = 0.f, sc, inv2;
    for (size_t i = 0; i < (size_t)256 * 256; ++i) mx = fmaxf(mx, fabsf(p2_w[i]));
    pow2_scale(mx, &sc, &inv2);
    const size_t n2h = (size_t)256 * 256;
    std::vector<unsigned short> w2(2 * n2h);
    for (int o = 0; o < 256; ++o)
      for (int k = 0; k < 256; ++k) {
        const float w = p2_w[(size_t)o * 256 + k] * sc;
        const __half h = __float2half_rn(w);
        w2[(size_t)o * 256 + k] = __half_as_ushort(h);
        w2[n2h + (size_t)o * 256 + k] = __half_as_ushort(__float2half_rn(w - __half2float(h)));
      }
    if (cudaMalloc(&dh->w2_half, w2.size() * sizeof(unsigned short)) != cudaSuccess) return VPHO_ERR_ALLOC;
    if (cudaMemcpy(dh->w2_half, w2.data(), w2.size() * sizeof(unsigned short), cudaMemcpyHostToDevice) != cudaSuccess) return VPHO_ERR_ALLOC;
    if (!tc_make_map(&dh->mapW2h_hi, dh->w2_half, 256, 256, 256, true) ||
        !tc_make_map(&dh->mapW2h_lo, static_cast<unsigned short*>(dh->w2_half) + n2h, 256, 256, 256, true))
      return VPHO_ERR_LAUNCH;
    d.W2scale_inv = inv2;
  }
  // head.0 pose slice: FP16 (hi, lo) planes [hid][256] K-major, one power-of-two scale per head
  {
    const size_t nw = (size_t)hid * kPDim;
    std::vector<unsigned short> hp(2 * nw);
    std::vector<float> inv(n_heads, 1.f);
    for (int nn = 0; nn < n_heads; ++nn) {
      float mx = 0.f, sc;
      for (int k = 0; k < kPDim; ++k) {
        const float* src = ha_w + ((size_t)nn * 1408 + 128 + k) * 256;
        for (int cc = 0; cc < 256; ++cc) mx = fmaxf(mx, fabsf(src[cc]));
      }
      pow2_scale(mx, &sc, &inv[nn]);
      for (int k = 0; k < kPDim; ++k) {
        const float* src = ha_w + ((size_t)nn * 1408 + 128 + k) * 256;
        for (int cc = 0; cc < 256; ++cc) {
          const float w = src[cc] * sc;
          const __half h = __float2half_rn(w);
          const __half l = __float2half_rn(w - __half2float(h));
          hp[((size_t)nn * 256 + cc) * kPDim + k] = __half_as_ushort(h);
          hp[nw + ((size_t)nn * 256 + cc) * kPDim + k] = __half_as_ushort(l);
        }
      }
    }
    const size_t bytes = 2 * nw * sizeof(unsigned short) + (size_t)n_heads * sizeof(float);
    if (cudaMalloc(&dh->w_half, bytes) != cudaSuccess) return VPHO_ERR_ALLOC;
    if (cudaMemcpy(dh->w_half, hp.data(), 2 * nw * sizeof(unsigned short), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(static_cast<char*>(dh->w_half) + 2 * nw * sizeof(unsigned short), inv.data(), n_heads * sizeof(float),
                   cudaMemcpyHostToDevice) != cudaSuccess)
      return VPHO_ERR_ALLOC;
    if (!tc_make_map(&dh->mapBh_hi, dh->w_half, hid, 256, kPDim, true) ||
        !tc_make_map(&dh->mapBh_lo, static_cast<unsigned short*>(dh->w_half) + nw, hid, 256, kPDim, true) ||
        !tc_make_map(&dh->mapBp_hi, dh->w_half, hid, 128, kPDim, true) ||
        !tc_make_map(&dh->mapBp_lo, static_cast<unsigned short*>(dh->w_half) + nw, hid, 128, kPDim, true))
      return VPHO_ERR_LAUNCH;
    d.Wscale_inv = reinterpret_cast<const float*>(static_cast<char*>(dh->w_half) + 2 * nw * sizeof(unsigned short));
  }
  dh->use_tc = true;
  return VPHO_OK;
}
#endif

extern "C" int vpho_denoiser_create_ex(int n_heads, const float* fourier_W, const float* t_w, const float* t_b,
                                       const float* p1_w, const float* p1_b, const float* p2_w, const float* p2_b,
                                       const float* ha_w, const float* ha_b, const float* hb_w, const float* hb_b,
                                       int flags, vpho_denoiser_t* out) {
  if (n_heads <= 0 || 3 * n_heads > kMaxD || !fourier_W || !t_w || !t_b || !p1_w || !p1_b || !p2_w || !p2_b || !ha_w ||
      !ha_b || !hb_w || !hb_b || !out || (flags & ~VPHO_DENOISER_STRICT_FP32))
    return VPHO_ERR_INVALID;
  const int D = 3 * n_heads, hid = n_heads * kHeadHid;
  const size_t n_four = 64, n_wt = 128 * 128, n_bt = 128, n_w1 = (size_t)D * 256, n_b1 = 256, n_w2 = 256 * 256,
               n_b2 = 256, n_wat = (size_t)128 * hid, n_wap = (size_t)256 * hid, n_waf = (size_t)1024 * hid,
               n_ba = hid, n_wb = (size_t)hid * 4, n_bb = D;
  size_t off = 0;
  auto take = [&](size_t cnt) { size_t o = off; off += (cnt + 63) / 64 * 64; return o; };
  const size_t o_four = take(n_four), o_wt = take(n_wt), o_bt = take(n_bt), o_w1 = take(n_w1), o_b1 = take(n_b1),
               o_w2 = take(n_w2), o_b2 = take(n_b2), o_wat = take(n_wat), o_wap = take(n_wap), o_waf = take(n_waf),
               o_ba = take(n_ba), o_wb = take(n_wb), o_bb = take(n_bb);
  std::vector<float> h(off, 0.f);
  for (size_t i = 0; i < n_four; ++i) h[o_four + i] = fourier_W[i];
  for (int o = 0; o < 128; ++o)
    for (int k = 0; k < 128; ++k) h[o_wt + (size_t)k * 128 + o] = t_w[(size_t)o * 128 + k];   // nn.Linear weight [out][in]
  for (int i = 0; i < 128; ++i) h[o_bt + i] = t_b[i];
  for (int o = 0; o < 256; ++o)
    for (int k = 0; k < D; ++k) h[o_w1 + (size_t)k * 256 + o] = p1_w[(size_t)o * D + k];
  for (int i = 0; i < 256; ++i) h[o_b1 + i] = p1_b[i];
  for (int o = 0; o < 256; ++o)
    for (int k = 0; k < 256; ++k) h[o_w2 + (size_t)k * 256 + o] = p2_w[(size_t)o * 256 + k];
  for (int i = 0; i < 256; ++i) h[o_b2 + i] = p2_b[i];
  // ParallelLinear weight [n][1408][256]: rows 0..127 time, 128..383 pose, 384..1407 conditioning (denoiser.py:75)
  for (int nn = 0; nn < n_heads; ++nn)
    for (int k = 0; k < 1408; ++k) {
      const float* src = ha_w + ((size_t)nn * 1408 + k) * 256;
      float* dst;
      if (k < 128) dst = &h[o_wat + (size_t)k * hid];
      else if (k < 384) dst = &h[o_wap + (size_t)(k - 128) * hid];
      else dst = &h[o_waf + (size_t)(k - 384) * hid];
      for (int cc = 0; cc < 256; ++cc) dst[nn * 256 + cc] = src[cc];
    }
  for (int i = 0; i < hid; ++i) h[o_ba + i] = ha_b[i];
  for (int nn = 0; nn < n_heads; ++nn)
    for (int cc = 0; cc < 256; ++cc)
      for (int d = 0; d < 3; ++d) h[o_wb + ((size_t)nn * 256 + cc) * 4 + d] = hb_w[((size_t)nn * 256 + cc) * 3 + d];
  for (int i = 0; i < D; ++i) h[o_bb + i] = hb_b[i];

  DenoiserHost* dh = new DenoiserHost();
  if (cudaMalloc((void**)&dh->blob, off * sizeof(float)) != cudaSuccess) { dh->blob = nullptr; free_denoiser(dh); return VPHO_ERR_ALLOC; }
  if (cudaMemcpy(dh->blob, h.data(), off * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) { free_denoiser(dh); return VPHO_ERR_ALLOC; }
  const float* b = dh->blob;
  DenoiserDev& d = dh->dev;
  d.n_heads = n_heads; d.D = D; d.hid = hid;
  d.fourier_W = b + o_four; d.Wt = b + o_wt; d.bt = b + o_bt; d.W1 = b + o_w1; d.b1 = b + o_b1; d.W2 = b + o_w2;
  d.b2 = b + o_b2; d.Wa_t = b + o_wat; d.Wa_p = b + o_wap; d.Wa_f = b + o_waf; d.ba = b + o_ba; d.Wb = b + o_wb;
  d.bb = b + o_bb; d.Wscale_inv = nullptr; d.W2scale_inv = 0.f;
#ifndef VPHO_EMU
  if (!(flags & VPHO_DENOISER_STRICT_FP32)) {
    const int rc = build_tc_planes(dh, n_heads, p1_w, p2_w, ha_w);
    if (rc != VPHO_OK) { free_denoiser(dh); return rc; }
  }
#endif
  *out = dh;
  return VPHO_OK;
}

extern "C" int vpho_denoiser_create(int n_heads, const float* fourier_W, const float* t_w, const float* t_b,
                                    const float* p1_w, const float* p1_b, const float* p2_w, const float* p2_b,
                                    const float* ha_w, const float* ha_b, const float* hb_w, const float* hb_b,
                                    vpho_denoiser_t* out) {
  return vpho_denoiser_create_ex(n_heads, fourier_W, t_w, t_b, p1_w, p1_b, p2_w, p2_b, ha_w, ha_b, hb_w, hb_b, 0, out);
}

extern "C" int vpho_denoiser_destroy(vpho_denoiser_t h) {
  if (!h) return VPHO_ERR_INVALID;
  free_denoiser(static_cast<DenoiserHost*>(h));
  return VPHO_OK;
}

#ifndef VPHO_EMU
namespace vpho { int tc_debug_clocks(int enable, unsigned long long* out, int n); }
#endif
// Debug only (not declared in the public header): timeline stamps of CTA 0 of the last tensor-core head GEMM launch.
extern "C" int vpho_debug_tc_clocks(int enable, unsigned long long* out, int n) {
#ifndef VPHO_EMU
  return vpho::tc_debug_clocks(enable, out, n);
#else
  return VPHO_ERR_INVALID;
#endif
}

extern "C" size_t vpho_sample_workspace_bytes(int n_heads, int n_rows, int rows_per_feat, int n_eval) {
  if (n_heads <= 0 || n_rows < 0 || rows_per_feat <= 0) return 0;
  return carve(nullptr, n_heads, n_rows, rows_per_feat, n_eval, nullptr);   // same size for both head-GEMM paths
}

extern "C" int vpho_score_eval(vpho_denoiser_t h, const float* x, float t, const float* feat, int n_rows,
                               int rows_per_feat, float* out, void* workspace, size_t workspace_bytes, void* stream) {
  if (!h || n_rows < 0 || rows_per_feat <= 0 || !workspace) return VPHO_ERR_INVALID;
  if (n_rows == 0) return VPHO_OK;
  if (!x || !feat || !out) return VPHO_ERR_INVALID;
  DenoiserHost& dh = *static_cast<DenoiserHost*>(h);
  const DenoiserDev& dn = dh.dev;
  SamplerJob job{&dh, {}, n_rows * dn.D};
  if (carve(workspace, dn.n_heads, n_rows, rows_per_feat, 1, &job.ws, dh.use_tc) > workspace_bytes) return VPHO_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  job.ws.eval_x = x; job.ws.eval_out = out;
  VPHO_LAUNCH(k_set_eval_time, dim3(1), dim3(1), 0, st, job.ws, t);
  int rcf = launch_feat_term(dh, job.ws, feat, st);
  if (rcf) return rcf;
  return launch_eval(&job, 1, kModeEval, 0, st);
}

// validates one sampler's arguments, carves its workspace and (begin only) initialises controller, state and feat-term
static int prepare_job(const vpho_sample_args* a, bool begin, SamplerJob* job, cudaStream_t st, bool defer_feat = false) {
  if (!a || !a->denoiser || a->n_rows < 0 || a->rows_per_feat <= 0 || !a->workspace) return VPHO_ERR_INVALID;
  if (begin && (a->n_eval < 1 || a->num_steps < 1)) return VPHO_ERR_INVALID;
  if (begin && a->n_rows > 0 && (!a->feat || !a->init_x || !a->x)) return VPHO_ERR_INVALID;
  DenoiserHost& dh = *static_cast<DenoiserHost*>(a->denoiser);
  const DenoiserDev& dn = dh.dev;
  job->dh = &dh;
  job->ws_n = a->n_rows * dn.D;
  if (carve(a->workspace, dn.n_heads, a->n_rows, a->rows_per_feat, a->n_eval, &job->ws, dh.use_tc) > a->workspace_bytes)
    return VPHO_ERR_INVALID;
  if (!begin) return VPHO_OK;
  SampleCfg cfg{a->T0, a->eps, a->rtol, a->atol, a->max_step, a->n_rows, dn.D, a->n_eval, a->num_steps, a->rows_per_feat, a->t_eval,
                a->xs, a->xs_f32, a->x, a->counters};
  VPHO_LAUNCH(k_init_ctrl, dim3(1), dim3(64), 0, st, cfg, job->ws);
  VPHO_CHECK_LAUNCH();
  if (a->n_rows == 0) return VPHO_OK;
  VPHO_LAUNCH(k_init_state, dim3(red_blocks(job->ws_n)), dim3(256), 0, st, job->ws, a->init_x, job->ws_n);
  if (defer_feat) return VPHO_OK;
  return launch_feat_term(dh, job->ws, a->feat, st);
}

// jobs with no rows drop out; returns the number left (their order is kept)
static int live_jobs(const vpho_sample_args* const* args, int n_args, bool begin, SamplerJob* jobs, cudaStream_t st, int* rc) {
  int n = 0;
  *rc = VPHO_OK;
  // two samplers begun together share one split-K feat-term launch (each alone is a latency-bound ~60 us kernel)
  bool both = begin && n_args == 2;
  for (int i = 0; both && i < 2; ++i)
    both = args[i] && args[i]->denoiser && args[i]->n_rows > 0;
  const float* feats[2] = {nullptr, nullptr};
  for (int i = 0; i < n_args; ++i) {
    SamplerJob j{};
    *rc = prepare_job(args[i], begin, &j, st, both);
    if (*rc) return 0;
    if (args[i]->n_rows > 0) { feats[n] = args[i]->feat; jobs[n++] = j; }
  }
  if (both && n == 2) {
    const DenoiserDev &d0 = jobs[0].dh->dev, &d1 = jobs[1].dh->dev;
    const SamplerWs &w0 = jobs[0].ws, &w1 = jobs[1].ws;
    const int rmax = w0.R > w1.R ? w0.R : w1.R;
    profile_begin(VPHO_TAG_FEAT_TERM, st);
    VPHO_LAUNCH(k_feat_term, dim3((d0.hid + d1.hid) / kFtCols, (rmax + kFtRows - 1) / kFtRows, kFtSplit), dim3(256), 0, st, d0, feats[0],
                w0.R, w0.Fpart, d1, feats[1], w1.R, w1.Fpart, d0.hid / kFtCols);
    for (int j = 0; j < 2; ++j) {
      const int n4 = jobs[j].ws.R * jobs[j].dh->dev.hid / 4;
      VPHO_LAUNCH(k_feat_sum, dim3((n4 + 255) / 256 < 1184 ? (n4 + 255) / 256 : 1184), dim3(256), 0, st, jobs[j].dh->dev, jobs[j].ws.Fpart,
                  jobs[j].ws.R, jobs[j].ws.F);
    }
    profile_end(VPHO_TAG_FEAT_TERM, st);
    if (cudaGetLastError() != cudaSuccess) { *rc = VPHO_ERR_LAUNCH; return 0; }
  }
  return n;
}

static int sample_begin(const vpho_sample_args* const* args, int n_args, int max_attempts, cudaStream_t st) {
  if (max_attempts < 0) return VPHO_ERR_INVALID;
  SamplerJob jobs[2];
  int rc;
  const int n = live_jobs(args, n_args, true, jobs, st, &rc);
  if (rc || n == 0) return rc;
  rc = launch_eval(jobs, n, kModeInit0, 0, st);
  if (rc) return rc;
  rc = launch_reduce(jobs, n, (int)kRedInit0, st);
  if (rc) return rc;
  rc = launch_eval(jobs, n, kModeInit1, 0, st);
  if (rc) return rc;
  rc = launch_reduce(jobs, n, (int)kRedInit1, st);
  if (rc) return rc;
  VPHO_CHECK_LAUNCH();
  return launch_attempts(jobs, n, max_attempts, st);
}

static int sample_continue(const vpho_sample_args* const* args, int n_args, int max_attempts, cudaStream_t st) {
  if (max_attempts < 0) return VPHO_ERR_INVALID;
  SamplerJob jobs[2];
  int rc;
  const int n = live_jobs(args, n_args, false, jobs, st, &rc);
  if (rc || n == 0) return rc;
  return launch_attempts(jobs, n, max_attempts, st);
}

static int sample_finish(const vpho_sample_args* const* args, int n_args, cudaStream_t st) {
  SamplerJob jobs[2];
  int rc;
  const int n = live_jobs(args, n_args, false, jobs, st, &rc);
  if (rc || n == 0) return rc;
  rc = launch_eval(jobs, n, kModeFinal, 0, st);
  if (rc) return rc;
  for (int j = 0; j < n; ++j) VPHO_LAUNCH(k_export, dim3(1), dim3(1), 0, st, jobs[j].ws);
  VPHO_CHECK_LAUNCH();
  return VPHO_OK;
}

extern "C" int vpho_sample_begin(vpho_denoiser_t h, const float* feat, int n_rows, int rows_per_feat,
                                 const float* init_x, double T0, double eps, const double* t_eval, int n_eval,
                                 double rtol, double atol, double max_step, int num_steps, int max_attempts,
                                 double* xs, double* x, int32_t* counters, void* workspace, size_t workspace_bytes,
                                 void* stream) {
  const vpho_sample_args a{h, feat, n_rows, rows_per_feat, init_x, T0, eps, t_eval, n_eval, rtol, atol, max_step, num_steps,
                           xs, x, counters, workspace, workspace_bytes, nullptr};
  const vpho_sample_args* p = &a;
  return sample_begin(&p, 1, max_attempts, (cudaStream_t)stream);
}

extern "C" int vpho_sample_continue(vpho_denoiser_t h, int n_rows, int rows_per_feat, int n_eval, int max_attempts,
                                    void* workspace, size_t workspace_bytes, void* stream) {
  vpho_sample_args a{};
  a.denoiser = h; a.n_rows = n_rows; a.rows_per_feat = rows_per_feat; a.n_eval = n_eval;
  a.workspace = workspace; a.workspace_bytes = workspace_bytes;
  const vpho_sample_args* p = &a;
  return sample_continue(&p, 1, max_attempts, (cudaStream_t)stream);
}

extern "C" int vpho_sample_finish(vpho_denoiser_t h, int n_rows, int rows_per_feat, int n_eval, void* workspace,
                                  size_t workspace_bytes, void* stream) {
  vpho_sample_args a{};
  a.denoiser = h; a.n_rows = n_rows; a.rows_per_feat = rows_per_feat; a.n_eval = n_eval;
  a.workspace = workspace; a.workspace_bytes = workspace_bytes;
  const vpho_sample_args* p = &a;
  return sample_finish(&p, 1, (cudaStream_t)stream);
}

extern "C" int vpho_sample_pair_begin(const vpho_sample_args* a, const vpho_sample_args* b, int max_attempts, void* stream) {
  if (!a || !b || a->workspace == b->workspace) return VPHO_ERR_INVALID;
  const vpho_sample_args* p[2] = {a, b};
  return sample_begin(p, 2, max_attempts, (cudaStream_t)stream);
}

extern "C" int vpho_sample_pair_continue(const vpho_sample_args* a, const vpho_sample_args* b, int max_attempts, void* stream) {
  if (!a || !b || a->workspace == b->workspace) return VPHO_ERR_INVALID;
  const vpho_sample_args* p[2] = {a, b};
  return sample_continue(p, 2, max_attempts, (cudaStream_t)stream);
}

extern "C" int vpho_sample_pair_finish(const vpho_sample_args* a, const vpho_sample_args* b, void* stream) {
  if (!a || !b || a->workspace == b->workspace) return VPHO_ERR_INVALID;
  const vpho_sample_args* p[2] = {a, b};
  return sample_finish(p, 2, (cudaStream_t)stream);
}

template <typename TIn>
static int postprocess_hand(const TIn* xs, int n_steps, int n_rows, int rows_per_shape, const float* shape, float* out, void* stream) {
  if (n_steps < 0 || n_rows < 0 || rows_per_shape <= 0) return VPHO_ERR_INVALID;
  if (n_steps == 0 || n_rows == 0) return VPHO_OK;
  if (!xs || !shape || !out) return VPHO_ERR_INVALID;
  const size_t total = (size_t)n_steps * n_rows * 17;
  size_t blocks = (total + 255) / 256;
  if (blocks > 148 * 64) blocks = 148 * 64;
  profile_begin(VPHO_TAG_POSTPROCESS, (cudaStream_t)stream);
  VPHO_LAUNCH(k_postprocess_hand<TIn>, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, xs, n_steps, n_rows, rows_per_shape,
              shape, out);
  profile_end(VPHO_TAG_POSTPROCESS, (cudaStream_t)stream);
  VPHO_CHECK_LAUNCH();
  return VPHO_OK;
}

extern "C" int vpho_postprocess_hand(const double* xs, int n_steps, int n_rows, int rows_per_shape, const float* shape,
                                     float* out, void* stream) {
  return postprocess_hand(xs, n_steps, n_rows, rows_per_shape, shape, out, stream);
}

extern "C" int vpho_postprocess_hand_f32(const float* xs, int n_steps, int n_rows, int rows_per_shape, const float* shape,
                                         float* out, void* stream) {
  return postprocess_hand(xs, n_steps, n_rows, rows_per_shape, shape, out, stream);
}

extern "C" int vpho_rot6d_to_axis_angle(const float* x6d, int n_rot, float* aa, void* stream) {
  if (n_rot < 0) return VPHO_ERR_INVALID;
  if (n_rot == 0) return VPHO_OK;
  if (!x6d || !aa) return VPHO_ERR_INVALID;
  VPHO_LAUNCH(k_rot6d_to_aa, dim3((n_rot + 255) / 256), dim3(256), 0, (cudaStream_t)stream, x6d, n_rot, aa);
  VPHO_CHECK_LAUNCH();
  return VPHO_OK;
}
