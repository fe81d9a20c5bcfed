// Interface between producers.cu (module wiring) and producers_tc.cu (tcgen05 implicit-GEMM kernel).  sm_100a only.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <vector>

namespace vpho {

struct TcGemm {
  int mode;                 // 0: rows (A = row-major planes [M][K]);  1: convolution over NHWC planes
  long long M;              // mode 0: rows
  int N;
  int chunks;               // K chunks of 64 per tap (= padded input channels / 64)
  int ntap;
  int n_img, H, W, bh, bn;  // mode 1: input geometry; tile = bn images x bh rows x W columns
  signed char dy[9], dx[9];
  float unscale;            // exact power of two that undoes the weight scaling
  const float* bias;
  const float* post_scale;  // BatchNorm (running statistics) folded to an affine map; nullptr = none
  const float* post_shift;
  float slope;              // activation: v >= 0 ? v : v * slope  (1 identity, 0 ReLU, 0.01 LeakyReLU)
  const __half* res_hi;     // residual as planes with the output's geometry (mode 1) ...
  const __half* res_lo;
  const float* res_f32;     // ... or float32 rows [M][ldc] (mode 0)
  __half* out_hi;           // planes output, channel / row stride out_cp (a multiple of 64)
  __half* out_lo;
  __half* out2_hi;          // second planes output: leaky(v * pre_scale[n] + pre_shift[n], pre_slope)
  __half* out2_lo;
  const float* pre_scale;
  const float* pre_shift;
  float pre_slope;
  int out_cp;
  float* out_f32;           // mode 1: NCHW [img][n][OH][OW] with image stride f32_img_stride;  mode 0: rows [M][ldc]
  long long f32_img_stride;
  int ldc;
  int os, py, px;           // mode 1: output pixel (y * os + py, x * os + px)
  int* overflow_flag;       // set when a value leaves the FP16 range of the planes (|v| >= 60000)
};

struct TcWeights {
  void* planes = nullptr;   // [2][Npad][Kp] __half (hi, lo)
  int N = 0, Npad = 0, Kp = 0, BN = 0;
  float unscale = 1.f;
  alignas(64) unsigned char map_hi[128];
  alignas(64) unsigned char map_lo[128];
};

bool pt_available();
bool pt_make_weights(TcWeights& w, const std::vector<float>& dense, int N, int Kp);
int pt_gemm(const TcWeights& w, TcGemm p, const __half* a_hi, const __half* a_lo, cudaStream_t st);
int pt_to_planes(const float* feat, const float* hm, const float* bbox, const float* bbox_rect, const unsigned char* is_right, int flip_feat,
                 int flip_hm, int bs, int C, int J, int roi, int Cp, __half* hi, __half* lo, cudaStream_t st);
int pt_pool(const __half* ihi, const __half* ilo, int n_img, int H, int W, int C, __half* ohi, __half* olo, __half* phi, __half* plo,
            const float* pre_scale, const float* pre_shift, float pre_slope, float* f32_nchw, cudaStream_t st);
int pt_split_rows(const float* x, long long n, __half* hi, __half* lo, cudaStream_t st);

}  // namespace vpho
