// N1: the modules that produce the hot path's inputs, `vpho_net.forward` lib/model/VPHO.py:129-178, eval mode.
//   HeadHeatmap2.forward   lib/model/head_inplane.py:99-104      Encoder / Residual   lib/model/encoding.py:21-73
//   HeadMano.forward       lib/model/head_mano.py:61-76           CrossModule.forward  lib/model/cross_module.py:119-137
//   HeadPhysics.forward    lib/model/physics.py:700-721           get_local_force      lib/model/physics.py:546-557
//   align_hm_to_bbox_rectangle / flip_tensor_by_mask_index / F.interpolate   lib/model/VPHO.py:136-148,333-357
//
// One FP32 SIMT GEMM core (128 x BN x 16 tiles, 8 x BN/16 outputs per thread, register-prefetched double buffer) runs every
// dense layer: 1x1 / 3x3 convolutions and the four output phases of the stride-2 transposed convolution as IMPLICIT GEMMs over
// the NCHW input (m = pixel, k = tap * Cin + ci; the halo is a predicated zero), linear layers over row-major activations.
// BatchNorm (running statistics) is an affine epilogue, the pre-activation BatchNorm + LeakyReLU of `Residual` is applied while
// the A tile is loaded, bias / activation / residual add are fused into the store.  Weights are re-laid out once at create
// time as [K][N] rows (zero padded to the tile) from the reference's state dict.  The same source compiles for the CPU SIMT
// emulator of tests/emu (toy dimensions): every dimension is read from the state dict's shapes.
#include "vpho_common.cuh"
#include "vpho_b200.h"
#include "rot_math.cuh"
#ifndef VPHO_EMU
#include "producers_tc.cuh"
#endif

#include <algorithm>
#include <map>
#include <string>
#include <vector>

namespace vpho {

constexpr int kGM = 128, kGK = 16;

struct GemmOp {
  // A operand
  const float* A;
  int mode;                 // 0: rows  A[m * lda + k];  1: implicit convolution over NCHW
  int M, N, K;
  int lda;
  int Cin, H, W;            // mode 1: input channels / size; M = images * H * W
  long long in_img_stride;  // elements between images of the input
  int ntap;
  signed char dy[9], dx[9];
  const float* pre_scale;   // mode 1, optional: a <- leaky(a * pre_scale[ci] + pre_shift[ci], pre_slope) for in-range taps
  const float* pre_shift;
  float pre_slope;
  // B operand: [Kpad][ldb], zero padded
  const float* B;
  int ldb;
  // epilogue: v = acc + bias[n]; v = v * post_scale[n] + post_shift[n]; v = v >= 0 ? v : v * slope; v += residual
  const float* bias;
  const float* post_scale;
  const float* post_shift;
  float slope;              // 1: identity, 0: ReLU, 0.01: LeakyReLU
  const float* residual;    // same layout as the output (mode 1: same image stride as the output)
  float* C;
  int ldc;                  // mode 0
  int os, py, px;           // mode 1: output pixel (y * os + py, x * os + px) of an (H * os, W * os) map
  long long out_img_stride;
};

// One 128 x BN output tile per CTA, 256 threads as 16 x 16: thread (ty, tx) owns rows ty*8..+8 and columns tx*TN..+TN.
template <int BN>
__global__ void __launch_bounds__(256) k_gemm_f32(const GemmOp op) {
  constexpr int TN = BN / 16;
  __shared__ __align__(16) float As[2][kGK][kGM + 4];
  __shared__ __align__(16) float Bs[2][kGK][BN];
  const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
  const int m0 = blockIdx.x * kGM, n0 = blockIdx.y * BN;

  // ---- A loader state
  // mode 1: thread -> pixel (t & 127), k offsets (t >> 7) + 2 i;  mode 0: thread -> row (t >> 1), k offsets (t & 1) * 8 + i
  const int am = op.mode ? (t & 127) : (t >> 1);
  const int m = m0 + am;
  const bool m_ok = m < op.M;
  const float* a_base = op.A;
  int py_ = 0, px_ = 0;
  if (op.mode) {
    const int HW = op.H * op.W;
    const int img = m_ok ? m / HW : 0, p = m_ok ? m - img * HW : 0;
    py_ = p / op.W;
    px_ = p - py_ * op.W;
    a_base += (long long)img * op.in_img_stride;
  } else {
    a_base += (long long)(m_ok ? m : 0) * op.lda;
  }
  float ra[8];
  constexpr int NB = (kGK * BN / 4 + 255) / 256;   // float4 loads of the B tile per thread per chunk
  float4 rb[NB];

  auto load_a = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float v = 0.f;
      if (op.mode) {
        const int k = k0 + (t >> 7) + 2 * i;
        if (m_ok && k < op.K) {
          const int tap = k / op.Cin, ci = k - tap * op.Cin;
          const int y = py_ + op.dy[tap], x = px_ + op.dx[tap];
          if (y >= 0 && y < op.H && x >= 0 && x < op.W) {
            v = a_base[((long long)ci * op.H + y) * op.W + x];
            if (op.pre_scale) {
              v = fmaf(v, op.pre_scale[ci], op.pre_shift[ci]);
              v = v >= 0.f ? v : v * op.pre_slope;
            }
          }
        }
      } else {
        const int k = k0 + (t & 1) * 8 + i;
        if (m_ok && k < op.K) v = a_base[k];
      }
      ra[i] = v;
    }
  };
  auto store_a = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int kl = op.mode ? (t >> 7) + 2 * i : (t & 1) * 8 + i;
      As[buf][kl][am] = ra[i];
    }
  };
  // B tile: kGK x BN floats = kGK * BN / 4 float4; thread f -> row f / (BN/4), column 4 * (f % (BN/4))
  auto load_b = [&](int k0) {
#pragma unroll
    for (int i = 0; i < NB; ++i) {
      const int f = t + i * 256;
      if (f < kGK * BN / 4) {
        const int r = f / (BN / 4), c = (f % (BN / 4)) * 4;
        rb[i] = *reinterpret_cast<const float4*>(op.B + (long long)(k0 + r) * op.ldb + n0 + c);
      }
    }
  };
  auto store_b = [&](int buf) {
#pragma unroll
    for (int i = 0; i < NB; ++i) {
      const int f = t + i * 256;
      if (f < kGK * BN / 4) {
        const int r = f / (BN / 4), c = (f % (BN / 4)) * 4;
        *reinterpret_cast<float4*>(&Bs[buf][r][c]) = rb[i];
      }
    }
  };

  // thread columns: TN = 8 -> two groups of four, 64 apart (a quarter-warp's LDS.128 then covers 128 contiguous bytes)
  auto col_of = [](int tx_, int j) { return TN == 8 ? (j >> 2) * 64 + tx_ * 4 + (j & 3) : tx_ * TN + j; };
  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const int nchunk = (op.K + kGK - 1) / kGK;
  load_a(0);
  load_b(0);
  store_a(0);
  store_b(0);
  __syncthreads();
  for (int c = 0; c < nchunk; ++c) {
    const int buf = c & 1;
    if (c + 1 < nchunk) {
      load_a((c + 1) * kGK);
      load_b((c + 1) * kGK);
    }
#pragma unroll
    for (int k = 0; k < kGK; ++k) {
      float a[8], b[TN];
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8 + 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Bs[buf][k][col_of(tx, j)];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (c + 1 < nchunk) {
      store_a(buf ^ 1);
      store_b(buf ^ 1);
    }
    __syncthreads();
  }

  // ---- epilogue
#pragma unroll
  for (int j = 0; j < TN; ++j) {
    const int n = n0 + col_of(tx, j);
    if (n >= op.N) continue;
    const float bias = op.bias ? op.bias[n] : 0.f;
    const float ps = op.post_scale ? op.post_scale[n] : 1.f, pb = op.post_scale ? op.post_shift[n] : 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int mm = m0 + ty * 8 + i;
      if (mm >= op.M) continue;
      float v = acc[i][j] + bias;
      if (op.post_scale) v = fmaf(v, ps, pb);
      v = v >= 0.f ? v : v * op.slope;
      long long o;
      if (op.mode) {
        const int HW = op.H * op.W;
        const int img = mm / HW, p = mm - img * HW, y = p / op.W, x = p - y * op.W;
        const int OW = op.W * op.os, OH = op.H * op.os;
        o = (long long)img * op.out_img_stride + ((long long)n * OH + (y * op.os + op.py)) * OW + (x * op.os + op.px);
      } else {
        o = (long long)mm * op.ldc + n;
      }
      if (op.residual) v += op.residual[o];
      op.C[o] = v;
    }
  }
}

// Encoder input (VPHO.py:136-151): channels [0, C) = the RoI feature map (mirrored along x for flagged images: the object
// branch's flip_tensor_by_mask_index), channels [C, C + J) = F.interpolate(flip(align_hm_to_bbox_rectangle(hm)), roi, bilinear).
// align: grid_sample(hm, grid[i][j] = (x = lin(i) * rel_w, y = lin(j) * rel_h)), bilinear, zeros, align_corners = False, with
// lin(i) = i / (n - 1) * 2 - 1 -- the (i, j) -> (x, y) order is the reference's ('ij' meshgrid stacked as (xx, yy)).
__device__ __forceinline__ float aligned_heat(const float* __restrict__ hm, int n, float relw, float relh, int i, int j) {
  const float gx = ((float)i / (float)(n - 1) * 2.f - 1.f) * relw;
  const float gy = ((float)j / (float)(n - 1) * 2.f - 1.f) * relh;
  const float ix = ((gx + 1.f) * (float)n - 1.f) * 0.5f, iy = ((gy + 1.f) * (float)n - 1.f) * 0.5f;
  const float fx = floorf(ix), fy = floorf(iy);
  const int x0 = (int)fx, y0 = (int)fy;
  const float tx = ix - fx, ty = iy - fy;
  auto at = [&](int y, int x) { return (y >= 0 && y < n && x >= 0 && x < n) ? hm[y * n + x] : 0.f; };
  // grid_sample's bilinear: nw * (1-tx)(1-ty) + ne * tx(1-ty) + sw * (1-tx) ty + se * tx ty
  return at(y0, x0) * ((1.f - tx) * (1.f - ty)) + at(y0, x0 + 1) * (tx * (1.f - ty)) + at(y0 + 1, x0) * ((1.f - tx) * ty) +
         at(y0 + 1, x0 + 1) * (tx * ty);
}

__global__ void __launch_bounds__(256) k_encoder_input(const float* __restrict__ feat, const float* __restrict__ hm,
                                                       const float* __restrict__ bbox, const float* __restrict__ bbox_rect,
                                                       const unsigned char* __restrict__ is_right, int flip_feat, int flip_hm,
                                                       int bs, int C, int J, int roi, float* __restrict__ out) {
  const int n = 2 * roi;
  const long long total = (long long)bs * (C + J) * roi * roi;
  for (long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x; it < total; it += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(it % roi), y = (int)(it / roi % roi);
    const int c = (int)(it / ((long long)roi * roi) % (C + J)), img = (int)(it / ((long long)roi * roi * (C + J)));
    const bool flip = !is_right[img];
    float v;
    if (c < C) {
      const int xs = (flip && flip_feat) ? roi - 1 - x : x;
      v = feat[(((long long)img * C + c) * roi + y) * roi + xs];
    } else {
      const float* h = hm + ((long long)img * J + (c - C)) * n * n;
      const float bw = bbox[img * 4 + 2] - bbox[img * 4 + 0], bh = bbox[img * 4 + 3] - bbox[img * 4 + 1];
      const float rw = (bbox_rect[img * 4 + 2] - bbox_rect[img * 4 + 0]) / bw, rh = (bbox_rect[img * 4 + 3] - bbox_rect[img * 4 + 1]) / bh;
      // F.interpolate 2n -> n (scale 2, align_corners False): source index 2 d + 0.5 -> taps 2d, 2d+1 with weights .5/.5
      float s[2][2];
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          const int i = 2 * y + a, j0 = 2 * x + b;
          const int j = (flip && flip_hm) ? n - 1 - j0 : j0;
          s[a][b] = aligned_heat(h, n, rw, rh, i, j);
        }
      v = 0.5f * (0.5f * s[0][0] + 0.5f * s[0][1]) + 0.5f * (0.5f * s[1][0] + 0.5f * s[1][1]);
    }
    out[it] = v;
  }
}

__global__ void __launch_bounds__(256) k_maxpool2(const float* __restrict__ in, long long planes, int H, int W, float* __restrict__ out) {
  const int OH = H / 2, OW = W / 2;
  const long long total = planes * OH * OW;
  for (long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x; it < total; it += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(it % OW), y = (int)(it / OW % OH);
    const long long pl = it / ((long long)OW * OH);
    const float* p = in + (pl * H + 2 * y) * W + 2 * x;
    out[it] = fmaxf(fmaxf(p[0], p[1]), fmaxf(p[W], p[W + 1]));
  }
}

// Gravity token + positional table (cross_module.py:125-131): row 2F of every image = gravity_proj(PosEmbedder(flipped gravity));
// then x[img][r][:] += pe[img][:] for all 2F + 1 rows.  Grid (images, slices): slice s adds the table to its share of the rows,
// the last slice also forms the gravity row (written with its table entry already added).
__global__ void __launch_bounds__(256) k_gravity_pe(const float* __restrict__ gravity, const unsigned char* __restrict__ is_right,
                                                    const float* __restrict__ Wg, const float* __restrict__ bg,
                                                    const float* __restrict__ pe, int n_tok, int d, float* __restrict__ x) {
  __shared__ float emb[64];
  const int img = blockIdx.x, ns = gridDim.y, sl = blockIdx.y;
  const int rows = n_tok - 1, r0 = (int)((long long)rows * sl / ns), r1 = (int)((long long)rows * (sl + 1) / ns);
  float* xi = x + (long long)img * n_tok * d;
  const float* pi = pe + (long long)img * d;
  for (int i = r0 * d + threadIdx.x; i < r1 * d; i += blockDim.x) xi[i] += pi[i % d];
  if (sl != ns - 1) return;
  if (threadIdx.x < 3) {
    float g = gravity[img * 3 + threadIdx.x];
    if (threadIdx.x == 0 && !is_right[img]) g = -g;          // flip_point3d_by_mask_index (VPHO.py:359-364)
    emb[threadIdx.x] = g;
    float f = 1.f;
    for (int q = 0; q < 10; ++q, f *= 2.f) {
      emb[3 + q * 6 + threadIdx.x] = sinf(g * f);
      emb[3 + q * 6 + 3 + threadIdx.x] = cosf(g * f);
    }
  }
  if (threadIdx.x == 63) emb[63] = 0.f;
  __syncthreads();
  float* row = xi + (long long)rows * d;
  // one warp per output: lanes over the 63 inputs (coalesced rows of Wg), fixed-order shuffle reduction
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int o = warp; o < d; o += nw) {
    float acc = emb[lane] * Wg[o * 63 + lane];
    if (lane + 32 < 63) acc = fmaf(emb[lane + 32], Wg[o * 63 + lane + 32], acc);
#pragma unroll
    for (int s2 = 16; s2; s2 >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s2);
    if (lane == 0) row[o] = acc + bg[o] + pi[o];
  }
}

// Self-attention of nn.TransformerEncoderLayer with batch_first = False on x (L = images, N = tokens, E): for token n and head
// h, image lq attends over all L images.  qkv [L][N][3E].  One CTA per (lq, h, n); scores live in shared memory.
__global__ void __launch_bounds__(256) k_attention(const float* __restrict__ qkv, int L, int N, int E, int nhead, float* __restrict__ out) {
  VPHO_DYN_SMEM(float, sc);       // [L] scores + [hd] query
  const int lq = blockIdx.x, h = blockIdx.y, n = blockIdx.z;
  const int hd = E / nhead;
  float* qs = sc + L;
  const float scale = 1.0f / sqrtf((float)hd);
  const float* q = qkv + ((long long)lq * N + n) * 3 * E + h * hd;
  for (int i = threadIdx.x; i < hd; i += blockDim.x) qs[i] = q[i] * scale;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int l = warp; l < L; l += nw) {
    const float* k = qkv + ((long long)l * N + n) * 3 * E + E + h * hd;
    float acc = 0.f;
    for (int i = lane; i < hd; i += 32) acc = fmaf(qs[i], k[i], acc);
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) sc[l] = acc;
  }
  __syncthreads();
  // softmax over the L scores by warp 0 (probabilities written back to shared memory once)
  if (warp == 0) {
    float mx = -INFINITY;
    for (int l = lane; l < L; l += 32) mx = fmaxf(mx, sc[l]);
#pragma unroll
    for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
    for (int l = lane; l < L; l += 32) {
      const float e = expf(sc[l] - mx);
      sc[l] = e;
      sum += e;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv = 1.f / sum;
    for (int l = lane; l < L; l += 32) sc[l] *= inv;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < hd; i += blockDim.x) {
    float acc = 0.f;
    for (int l = 0; l < L; ++l) acc = fmaf(sc[l], qkv[((long long)l * N + n) * 3 * E + 2 * E + h * hd + i], acc);
    out[((long long)lq * N + n) * E + h * hd + i] = acc;
  }
}

// The same attention with one CTA per (token n, head h): K and V of the L images are staged once in shared memory (row stride
// hd + 1: lanes walk different rows conflict-free) and every warp serves queries lq = warp, warp + 8, ...  Used when
// 2 L (hd + 1) floats fit in shared memory (the per-query kernel re-reads K and V from L2 for every query).
__global__ void __launch_bounds__(256) k_attention_tile(const float* __restrict__ qkv, int L, int N, int E, int nhead, float* __restrict__ out) {
  VPHO_DYN_SMEM(float, sm);       // K [L][hd + 1], V [L][hd + 1], per warp: q [hd], p [L]
  const int n = blockIdx.x, h = blockIdx.y, hd = E / nhead, ld = hd + 1;
  float* Ks = sm;
  float* Vs = sm + (size_t)L * ld;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  float* qs = Vs + (size_t)L * ld + (size_t)warp * (hd + L);
  float* ps = qs + hd;
  if (hd % 4 == 0) {
    for (int i = threadIdx.x; i < L * hd / 4; i += blockDim.x) {
      const int l = (i * 4) / hd, c = i * 4 - l * hd;
      const float* src = qkv + ((long long)l * N + n) * 3 * E + h * hd + c;
      const float4 k4 = *reinterpret_cast<const float4*>(src + E), v4 = *reinterpret_cast<const float4*>(src + 2 * E);
      float* kd = Ks + l * ld + c;
      float* vd = Vs + l * ld + c;
      kd[0] = k4.x; kd[1] = k4.y; kd[2] = k4.z; kd[3] = k4.w;
      vd[0] = v4.x; vd[1] = v4.y; vd[2] = v4.z; vd[3] = v4.w;
    }
  } else {
    for (int i = threadIdx.x; i < L * hd; i += blockDim.x) {
      const int l = i / hd, c = i - l * hd;
      const float* src = qkv + ((long long)l * N + n) * 3 * E + h * hd + c;
      Ks[l * ld + c] = src[E];
      Vs[l * ld + c] = src[2 * E];
    }
  }
  __syncthreads();
  const float scale = 1.0f / sqrtf((float)hd);
  for (int lq = blockIdx.z * nw + warp; lq < L; lq += nw * gridDim.z) {       // grid.z slices of the queries
    const float* q = qkv + ((long long)lq * N + n) * 3 * E + h * hd;
    for (int i = lane; i < hd; i += 32) qs[i] = q[i] * scale;
    __syncwarp();
    float mx = -INFINITY;
    for (int l = lane; l < L; l += 32) {
      float acc = 0.f;
      for (int i = 0; i < hd; ++i) acc = fmaf(qs[i], Ks[l * ld + i], acc);
      ps[l] = acc;
      mx = fmaxf(mx, acc);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
    for (int l = lane; l < L; l += 32) {
      const float e = expf(ps[l] - mx);
      ps[l] = e;
      sum += e;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv = 1.f / sum;
    __syncwarp();
    for (int i = lane; i < hd; i += 32) {
      float acc = 0.f;
      for (int l = 0; l < L; ++l) acc = fmaf(ps[l] * inv, Vs[l * ld + i], acc);
      out[((long long)lq * N + n) * E + h * hd + i] = acc;
    }
    __syncwarp();
  }
}

// y = LayerNorm(x) * w + b over rows of d (x already holds input + sub-layer output; eps 1e-5, biased variance).
__global__ void __launch_bounds__(128) k_layernorm(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                                                   int d, float* __restrict__ y) {
  __shared__ float red[4];
  const float* r = x + (long long)blockIdx.x * d;
  auto block_sum = [&](float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    return red[0] + red[1] + red[2] + red[3];
  };
  float s = 0.f;
  for (int i = threadIdx.x; i < d; i += 128) s += r[i];
  const float mean = block_sum(s) / (float)d;
  float q = 0.f;
  for (int i = threadIdx.x; i < d; i += 128) {
    const float c = r[i] - mean;
    q = fmaf(c, c, q);
  }
  const float rstd = 1.f / sqrtf(block_sum(q) / (float)d + 1e-5f);
  for (int i = threadIdx.x; i < d; i += 128) y[(long long)blockIdx.x * d + i] = (r[i] - mean) * rstd * w[i] + b[i];
}

// HeadPhysics tail (physics.py:700-721, 546-557): softmax over the 8 cone weights (fc_weight's Softmax), then get_local_force:
// softmax AGAIN, friction-scaled anchors, normalise, times |scale|.  head[row][12] = scale(1) | weight logits(8) | CoM(3).
__global__ void __launch_bounds__(128) k_physics_tail(const float* __restrict__ head, const float* __restrict__ anchor, int rows,
                                                      float* __restrict__ force_local, float* __restrict__ scale_out,
                                                      float* __restrict__ weight_out, float* __restrict__ com_out) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const float* h = head + (long long)r * 12;
  float w[8], mx = -INFINITY, sum = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) mx = fmaxf(mx, h[1 + i]);
#pragma unroll
  for (int i = 0; i < 8; ++i) { w[i] = expf(h[1 + i] - mx); sum += w[i]; }
#pragma unroll
  for (int i = 0; i < 8; ++i) w[i] /= sum;
  if (weight_out)
#pragma unroll
    for (int i = 0; i < 8; ++i) weight_out[(long long)r * 8 + i] = w[i];
  float w2[8];
  mx = -INFINITY; sum = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) mx = fmaxf(mx, w[i]);
#pragma unroll
  for (int i = 0; i < 8; ++i) { w2[i] = expf(w[i] - mx); sum += w2[i]; }
  float d[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float wi = w2[i] / sum;
    d[0] = fmaf(wi, anchor[i * 3 + 0] * 0.8f, d[0]);
    d[1] = fmaf(wi, anchor[i * 3 + 1] * 0.8f, d[1]);
    d[2] = fmaf(wi, anchor[i * 3 + 2], d[2]);
  }
  const float nrm = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]) + 1e-8f;
  const float sc = fabsf(h[0]);
#pragma unroll
  for (int k = 0; k < 3; ++k) force_local[(long long)r * 3 + k] = d[k] / nrm * sc;
  if (scale_out) scale_out[r] = h[0];
  if (com_out)
#pragma unroll
    for (int k = 0; k < 3; ++k) com_out[(long long)r * 3 + k] = h[9 + k];
}

// 6D -> matrix -> axis-angle of the 16 joints of every row (head_mano.py:66-69); rows of `ld` floats, the first 96 are the 6D pose
__global__ void __launch_bounds__(256) k_rot6d_rows(const float* __restrict__ x6d, int n, int ld, float* __restrict__ aa) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float d[6], R[9], a[3];
#pragma unroll
  for (int k = 0; k < 6; ++k) d[k] = x6d[(long long)(i / 16) * ld + (i % 16) * 6 + k];
  rot6d_to_matrix(d, R);
  matrix_to_axis_angle(R, a);
  aa[(long long)i * 3 + 0] = a[0]; aa[(long long)i * 3 + 1] = a[1]; aa[(long long)i * 3 + 2] = a[2];
}

__global__ void __launch_bounds__(256) k_copy_cols(const float* __restrict__ src, int rows, int lds, int c0, int nc, float* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < rows * nc) dst[i] = src[(long long)(i / nc) * lds + c0 + i % nc];
}

// ---------------------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------------------
struct DevMat {          // a [Kpad][ldb] weight matrix + per-output vectors
  float* B = nullptr;
  int K = 0, N = 0, ldb = 0;
  float* bias = nullptr;
  float* post_scale = nullptr;
  float* post_shift = nullptr;
#ifndef VPHO_EMU
  TcWeights tc;            // the same matrix as [N][ntap * Cp] FP16 hi/lo planes (tensor-core path); Cp = cin padded to 64
  int tc_cp = 0, tc_ntap = 0;
#endif
};
struct DevVec2 { float* scale = nullptr; float* shift = nullptr; };

struct HeatHead { DevMat c0, c1, dc[4], fin; int hid = 0, J = 0; };
struct ResBlock { DevVec2 pre; DevMat c1, c2, c3; };
struct EncoderW { DevMat project; std::vector<ResBlock> reg; int cin = 0, hid = 0; };
struct CrossW { DevMat proj_hand, proj_obj, in_proj, out_proj, lin1, lin2; float *Wg = nullptr, *bg = nullptr, *pe = nullptr, *n1w = nullptr,
                *n1b = nullptr, *n2w = nullptr, *n2b = nullptr; int pe_rows = 0, d = 0, ff = 0, proj_dim = 0; };

}  // namespace vpho

using namespace vpho;

struct vpho_heads {
  int C = 0, Jh = 0, Jo = 0, enc_dim = 0, d_model = 0, n_force = 32, heat_hid = 0, enc_hid = 0, mano_h1 = 0, mano_h2 = 0, phys_hid = 0;
  HeatHead hm_hand, hm_obj;
  EncoderW enc_hand, enc_obj;
  DevMat mano0, mano1, mano_out;      // mano_out: [fc_pose (96) | fc_shape (10)] stacked
  CrossW cross_hand, cross_obj;
  DevMat phys_scale0, phys_weight0, phys_com0, phys_scale2, phys_weight2, phys_com2;
  float* anchor = nullptr;
  bool tc_ok = false;       // tensor-core planes built (sm_100a build with cuTensorMapEncodeTiled; widths multiples of 64)
  int* overflow = nullptr;  // device flag: an activation left the FP16 range of the operand planes
  std::vector<void*> allocs;
};

namespace {

struct Table {
  std::map<std::string, const vpho_named_tensor*> m;
  bool ok = true;
  const vpho_named_tensor* get(const std::string& k, int ndim) {
    auto it = m.find(k);
    if (it == m.end() || it->second->ndim != ndim || !it->second->data) {
      ok = false;
      return nullptr;
    }
    return it->second;
  }
  bool has(const std::string& k) const { return m.count(k) != 0; }
};

float* upload(vpho_heads* h, const std::vector<float>& v) {
  float* d = nullptr;
  if (cudaMalloc((void**)&d, std::max<size_t>(v.size(), 1) * sizeof(float)) != cudaSuccess) return nullptr;
  h->allocs.push_back(d);
  cudaMemcpy(d, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice);
  return d;
}

int round_up(int a, int b) { return (a + b - 1) / b * b; }

// [K][N] matrix from a generator, zero padded to (multiple of 16) x (multiple of 128)
template <typename F>
bool make_mat(vpho_heads* h, DevMat& dm, int K, int N, F at, const float* bias, int ntap = 1, bool want_tc = false) {
  dm.K = K;
  dm.N = N;
  dm.ldb = round_up(N, 128);
  const int Kp = round_up(K, kGK);
  std::vector<float> B((size_t)Kp * dm.ldb, 0.f);
  for (int k = 0; k < K; ++k)
    for (int n = 0; n < N; ++n) B[(size_t)k * dm.ldb + n] = at(k, n);
  dm.B = upload(h, B);
  if (bias) dm.bias = upload(h, std::vector<float>(bias, bias + N));
#ifndef VPHO_EMU
  if (want_tc && h->tc_ok) {
    const int cin = K / ntap, Cp = round_up(cin, 64);
    std::vector<float> dense((size_t)N * ntap * Cp, 0.f);
    for (int n = 0; n < N; ++n)
      for (int t = 0; t < ntap; ++t)
        for (int c = 0; c < cin; ++c) dense[((size_t)n * ntap + t) * Cp + c] = at(t * cin + c, n);
    dm.tc_cp = Cp;
    dm.tc_ntap = ntap;
    if (!pt_make_weights(dm.tc, dense, N, ntap * Cp)) h->tc_ok = false;
    else h->allocs.push_back(dm.tc.planes);
  }
#endif
  return dm.B && (!bias || dm.bias);
}

// Conv2d weight [Co][Ci][k][k] -> rows k = tap * Ci + ci (tap = ky * ks + kx)
bool conv_mat(vpho_heads* h, Table& T, const std::string& p, DevMat& dm, int& Co, int& Ci, int& ks) {
  const vpho_named_tensor* w = T.get(p + ".weight", 4);
  if (!w) return false;
  Co = (int)w->shape[0]; Ci = (int)w->shape[1]; ks = (int)w->shape[2];
  if (w->shape[3] != ks || (ks != 1 && ks != 3)) return T.ok = false;
  const vpho_named_tensor* b = T.has(p + ".bias") ? T.get(p + ".bias", 1) : nullptr;
  if (b && b->shape[0] != Co) return T.ok = false;
  const float* W = w->data;
  const int kk = ks * ks, ci_n = Ci;
  return make_mat(h, dm, kk * Ci, Co, [=](int k, int n) { return W[((size_t)n * ci_n + k % ci_n) * kk + k / ci_n]; }, b ? b->data : nullptr, kk, true);
}

bool linear_mat(vpho_heads* h, Table& T, const std::string& p, DevMat& dm, int& out, int& in, bool want_tc = false) {
  const vpho_named_tensor* w = T.get(p + ".weight", 2);
  const vpho_named_tensor* b = T.get(p + ".bias", 1);
  if (!w || !b) return false;
  out = (int)w->shape[0]; in = (int)w->shape[1];
  if (b->shape[0] != out) return T.ok = false;
  const float* W = w->data;
  const int in_n = in;
  return make_mat(h, dm, in, out, [=](int k, int n) { return W[(size_t)n * in_n + k]; }, b->data, 1, want_tc);
}

// BatchNorm2d (eval) as y = x * scale + shift
bool bn_affine(vpho_heads* h, Table& T, const std::string& p, int C, float** scale, float** shift) {
  const vpho_named_tensor *w = T.get(p + ".weight", 1), *b = T.get(p + ".bias", 1), *m = T.get(p + ".running_mean", 1),
                          *v = T.get(p + ".running_var", 1);
  if (!w || !b || !m || !v) return false;
  if (w->shape[0] != C || b->shape[0] != C || m->shape[0] != C || v->shape[0] != C) return T.ok = false;
  std::vector<float> s(C), t(C);
  for (int c = 0; c < C; ++c) {
    s[c] = w->data[c] / sqrtf(v->data[c] + 1e-5f);
    t[c] = b->data[c] - m->data[c] * s[c];
  }
  *scale = upload(h, s);
  *shift = upload(h, t);
  return *scale && *shift;
}

bool vec(vpho_heads* h, Table& T, const std::string& k, int n, float** out) {
  const vpho_named_tensor* t = T.get(k, 1);
  if (!t) return false;
  if (t->shape[0] != n) return T.ok = false;
  *out = upload(h, std::vector<float>(t->data, t->data + n));
  return *out != nullptr;
}

bool build_heat(vpho_heads* h, Table& T, const std::string& p, HeatHead& hh, int& C) {
  int co, ci, ks;
  if (!conv_mat(h, T, p + ".conv_layers.0", hh.c0, co, ci, ks) || ks != 3) return false;
  C = ci;
  hh.hid = co;
  if (!conv_mat(h, T, p + ".conv_layers.1", hh.c1, co, ci, ks) || ks != 3 || co != hh.hid || ci != hh.hid) return false;
  if (!bn_affine(h, T, p + ".conv_layers.2", hh.hid, &hh.c1.post_scale, &hh.c1.post_shift)) return false;
  // ConvTranspose2d(hid, hid/2, 4, stride 2, padding 1, no bias): weight [Ci][Co][4][4]; phase (py, px) of the output uses
  // taps ky in {1 (dy 0), 3 (dy -1)} for even rows, {0 (dy +1), 2 (dy 0)} for odd rows; same along x.
  const vpho_named_tensor* w = T.get(p + ".deconv_layers.0.weight", 4);
  if (!w) return false;
  const int Ci = (int)w->shape[0], Co = (int)w->shape[1];
  if (Ci != hh.hid || w->shape[2] != 4 || w->shape[3] != 4) return T.ok = false;
  float *ds = nullptr, *dt = nullptr;
  if (!bn_affine(h, T, p + ".deconv_layers.1", Co, &ds, &dt)) return false;
  const float* W = w->data;
  for (int ph = 0; ph < 4; ++ph) {
    const int py = ph >> 1, px = ph & 1;
    const int kys[2] = {py ? 0 : 1, py ? 2 : 3}, kxs[2] = {px ? 0 : 1, px ? 2 : 3};
    if (!make_mat(h, hh.dc[ph], 4 * Ci, Co, [=](int k, int n) {
          const int tap = k / Ci, ci2 = k % Ci;
          return W[(((size_t)ci2 * Co + n) * 4 + kys[tap >> 1]) * 4 + kxs[tap & 1]];
        }, nullptr, 4, true)) return false;
    hh.dc[ph].post_scale = ds;
    hh.dc[ph].post_shift = dt;
  }
  if (!conv_mat(h, T, p + ".final_layer", hh.fin, co, ci, ks) || ks != 1 || ci != Co) return false;
  hh.J = co;
  return true;
}

bool build_encoder(vpho_heads* h, Table& T, const std::string& p, EncoderW& e) {
  int co, ci, ks;
  if (!conv_mat(h, T, p + ".project", e.project, co, ci, ks) || ks != 1) return false;
  e.cin = ci;
  e.hid = co;
  for (int r = 0; T.has(p + ".reg." + std::to_string(r) + ".conv1.weight"); ++r) {
    const std::string q = p + ".reg." + std::to_string(r);
    ResBlock rb;
    if (T.has(q + ".conv4.weight")) return T.ok = false;      // Encoder only builds Residual(hid, hid): no projection shortcut
    if (!bn_affine(h, T, q + ".bn", e.hid, &rb.pre.scale, &rb.pre.shift)) return false;
    if (!conv_mat(h, T, q + ".conv1", rb.c1, co, ci, ks) || ks != 1 || ci != e.hid) return false;
    const int mid = co;
    if (!bn_affine(h, T, q + ".bn1", mid, &rb.c1.post_scale, &rb.c1.post_shift)) return false;
    if (!conv_mat(h, T, q + ".conv2", rb.c2, co, ci, ks) || ks != 3 || ci != mid || co != mid) return false;
    if (!bn_affine(h, T, q + ".bn2", mid, &rb.c2.post_scale, &rb.c2.post_shift)) return false;
    if (!conv_mat(h, T, q + ".conv3", rb.c3, co, ci, ks) || ks != 1 || ci != mid || co != e.hid) return false;
    e.reg.push_back(rb);
  }
  return e.reg.size() == 8;      // Encoder(nRegBlock = 4, nRegModules = 2)
}

bool build_cross(vpho_heads* h, Table& T, const std::string& p, CrossW& c, int enc_hid) {
  int co, ci, ks, out, in;
  if (!conv_mat(h, T, p + ".proj_hand", c.proj_hand, co, ci, ks) || ks != 3 || ci != enc_hid) return false;
  c.proj_dim = co;
  if (!conv_mat(h, T, p + ".proj_obj", c.proj_obj, co, ci, ks) || ks != 3 || ci != enc_hid || co != c.proj_dim) return false;
  const vpho_named_tensor *wg = T.get(p + ".gravity_proj.weight", 2), *bg = T.get(p + ".gravity_proj.bias", 1);
  if (!wg || !bg) return false;
  if (wg->shape[1] != 63) return T.ok = false;
  c.d = (int)wg->shape[0];
  c.Wg = upload(h, std::vector<float>(wg->data, wg->data + (size_t)c.d * 63));
  c.bg = upload(h, std::vector<float>(bg->data, bg->data + c.d));
  const vpho_named_tensor* pe = T.get(p + ".pose_embedder.pe", 3);
  if (!pe) return false;
  if (pe->shape[2] != c.d || pe->shape[1] != 1) return T.ok = false;
  c.pe_rows = (int)std::min<int64_t>(pe->shape[0], 4096);
  c.pe = upload(h, std::vector<float>(pe->data, pe->data + (size_t)c.pe_rows * c.d));
  const std::string a = p + ".attn.layers.0";
  const vpho_named_tensor *ipw = T.get(a + ".self_attn.in_proj_weight", 2), *ipb = T.get(a + ".self_attn.in_proj_bias", 1);
  if (!ipw || !ipb) return false;
  if (ipw->shape[0] != 3 * c.d || ipw->shape[1] != c.d || ipb->shape[0] != 3 * c.d) return T.ok = false;
  {
    const float* W = ipw->data;
    const int d = c.d;
    if (!make_mat(h, c.in_proj, d, 3 * d, [=](int k, int n) { return W[(size_t)n * d + k]; }, ipb->data, 1, true)) return false;
  }
  if (!linear_mat(h, T, a + ".self_attn.out_proj", c.out_proj, out, in, true) || out != c.d || in != c.d) return false;
  if (!linear_mat(h, T, a + ".linear1", c.lin1, out, in, true) || in != c.d) return false;
  c.ff = out;
  if (!linear_mat(h, T, a + ".linear2", c.lin2, out, in, true) || in != c.ff || out != c.d) return false;
  return vec(h, T, a + ".norm1.weight", c.d, &c.n1w) && vec(h, T, a + ".norm1.bias", c.d, &c.n1b) &&
         vec(h, T, a + ".norm2.weight", c.d, &c.n2w) && vec(h, T, a + ".norm2.bias", c.d, &c.n2b) && c.Wg && c.bg && c.pe;
}

// ---- launches
int run_gemm(const GemmOp& op, cudaStream_t st) {
  if (op.M <= 0 || op.N <= 0) return VPHO_OK;
  if (op.N > 32) {
    VPHO_LAUNCH(k_gemm_f32<128>, dim3((op.M + kGM - 1) / kGM, (op.N + 127) / 128), dim3(256), 0, st, op);
  } else {
    VPHO_LAUNCH(k_gemm_f32<32>, dim3((op.M + kGM - 1) / kGM, 1), dim3(256), 0, st, op);
  }
  VPHO_CHECK_LAUNCH();
  return VPHO_OK;
}

GemmOp base_op(const DevMat& w, float slope) {
  GemmOp op = {};
  op.B = w.B; op.ldb = w.ldb; op.N = w.N; op.K = w.K;
  op.bias = w.bias; op.post_scale = w.post_scale; op.post_shift = w.post_shift;
  op.slope = slope;
  op.os = 1;
  return op;
}

// convolution (ks x ks, stride 1, same padding) of in [bs][Cin][H][W] (image stride in_stride) -> out [bs][N][H][W]
int conv(const DevMat& w, int ks, const float* in, long long in_stride, int bs, int Cin, int H, int W, float slope,
         const DevVec2* pre, const float* residual, float* out, long long out_stride, cudaStream_t st) {
  GemmOp op = base_op(w, slope);
  op.mode = 1; op.A = in; op.M = bs * H * W; op.Cin = Cin; op.H = H; op.W = W; op.in_img_stride = in_stride;
  op.ntap = ks * ks;
  for (int t = 0; t < op.ntap; ++t) { op.dy[t] = (signed char)(t / ks - ks / 2); op.dx[t] = (signed char)(t % ks - ks / 2); }
  if (pre) { op.pre_scale = pre->scale; op.pre_shift = pre->shift; op.pre_slope = 0.01f; }
  op.residual = residual; op.C = out; op.out_img_stride = out_stride;
  return run_gemm(op, st);
}

int linear(const DevMat& w, const float* in, int rows, int lda, float slope, const float* residual, float* out, int ldc, cudaStream_t st) {
  GemmOp op = base_op(w, slope);
  op.mode = 0; op.A = in; op.M = rows; op.lda = lda; op.residual = residual; op.C = out; op.ldc = ldc;
  return run_gemm(op, st);
}

int run_attention(const float* qkv, int L, int N, int E, int nhead, float* out, cudaStream_t st) {
  const int hd = E / nhead;
  const size_t tile = ((size_t)2 * L * (hd + 1) + (size_t)8 * (hd + L)) * sizeof(float);
  if (tile <= 200 * 1024) {
#ifndef VPHO_EMU
    static bool attr_done[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !attr_done[dev]) {
      if (cudaFuncSetAttribute(k_attention_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) return VPHO_ERR_LAUNCH;
      attr_done[dev] = true;
    }
#endif
    VPHO_LAUNCH(k_attention_tile, dim3(N, nhead, L >= 32 ? 4 : 1), dim3(256), tile, st, qkv, L, N, E, nhead, out);
  } else {
    VPHO_LAUNCH(k_attention, dim3(L, nhead, N), dim3(256), (size_t)(L + hd) * sizeof(float), st, qkv, L, N, E, nhead, out);
  }
  VPHO_CHECK_LAUNCH();
  return VPHO_OK;
}

int grid_for(long long total) { return (int)std::min<long long>((total + 255) / 256, 148 * 16); }

struct Ws {
  float *cat, *xa, *xb, *mid1, *mid2, *dec, *pool1_hand, *pool1_obj, *tok, *qkv, *att, *tmp, *ffn, *mano_a, *mano_b, *mano_o, *ph_tok, *ph_mid, *ph_head;
  // tensor-core path: FP16 (hi, lo) plane pairs; `n_*` = elements of ONE plane (lo follows hi)
  unsigned short *pF, *pX0, *pX1, *pXA, *pM1, *pM2, *pDEC, *pP1h, *pP1o, *pTOK, *pATT, *pFFN;
  size_t nF, nX, nM, nDEC, nP1, nTOK, nFFN;
  size_t bytes;
};

Ws carve(const vpho_heads* h, int bs, int roi, void* base) {
  Ws w = {};
  size_t off = 0;
  auto take = [&](size_t n) {
    float* p = base ? (float*)((char*)base + off) : nullptr;
    off += (n * sizeof(float) + 255) / 256 * 256;
    return p;
  };
  const size_t px = (size_t)roi * roi, b = (size_t)bs;
  const int Jmax = std::max(h->Jh, h->Jo), ntok = 2 * h->n_force + 1;
  w.cat = take(b * (h->C + Jmax) * px);
  w.xa = take(b * std::max(h->enc_hid, h->heat_hid) * px);
  w.xb = take(b * std::max(h->enc_hid, h->heat_hid) * px);
  w.mid1 = take(b * std::max(h->enc_hid / 2, h->heat_hid) * px);
  w.mid2 = take(b * (h->enc_hid / 2) * px);
  w.dec = take(b * (h->heat_hid / 2) * px * 4);
  w.pool1_hand = take(b * h->enc_hid * px / 16);
  w.pool1_obj = take(b * h->enc_hid * px / 16);
  w.tok = take(b * ntok * h->d_model);
  w.qkv = take(b * ntok * 3 * h->d_model);
  w.att = take(b * ntok * h->d_model);
  w.tmp = take(b * ntok * h->d_model);
  w.ffn = take(b * ntok * std::max(h->cross_hand.ff, h->cross_obj.ff));
  w.mano_a = take(b * h->mano_h1);
  w.mano_b = take(b * h->mano_h2);
  w.mano_o = take(b * 106);
  w.ph_tok = take(b * h->n_force * h->d_model);
  w.ph_mid = take(b * h->n_force * h->phys_hid);
  w.ph_head = take(b * h->n_force * 12);
  if (h->tc_ok) {
    auto take_planes = [&](size_t n_plane) { return reinterpret_cast<unsigned short*>(take(n_plane)); };     // 2 planes x 2 bytes = 4 bytes / element
    w.nF = b * px * (size_t)round_up(h->C + Jmax, 64);
    w.nX = b * px * (size_t)h->enc_hid;
    w.nM = b * px * (size_t)std::max(h->enc_hid / 2, h->heat_hid);
    w.nDEC = b * px * 4 * (size_t)(h->heat_hid / 2);
    w.nP1 = b * (px / 16) * (size_t)h->enc_hid;
    w.nTOK = b * std::max<size_t>({(size_t)ntok * h->d_model, (size_t)h->enc_dim, (size_t)h->mano_h1, (size_t)h->n_force * h->phys_hid});
    w.nFFN = b * std::max<size_t>((size_t)ntok * std::max(h->cross_hand.ff, h->cross_obj.ff), (size_t)h->mano_h2);
    w.pF = take_planes(w.nF); w.pX0 = take_planes(w.nX); w.pX1 = take_planes(w.nX); w.pXA = take_planes(w.nX);
    w.pM1 = take_planes(w.nM); w.pM2 = take_planes(w.nM); w.pDEC = take_planes(w.nDEC);
    w.pP1h = take_planes(w.nP1); w.pP1o = take_planes(w.nP1);
    w.pTOK = take_planes(w.nTOK); w.pATT = take_planes(w.nTOK); w.pFFN = take_planes(w.nFFN);
  }
  w.bytes = off;
  return w;
}

// HeadHeatmap2.forward (head_inplane.py:99-104)
int run_heat(const HeatHead& hh, const float* feat, int bs, int C, int roi, const Ws& w, float* out, cudaStream_t st) {
  const long long px = (long long)roi * roi;
  int rc;
  if ((rc = conv(hh.c0, 3, feat, C * px, bs, C, roi, roi, 1.f, nullptr, nullptr, w.xa, hh.hid * px, st))) return rc;
  if ((rc = conv(hh.c1, 3, w.xa, hh.hid * px, bs, hh.hid, roi, roi, 1.f /* LeakyReLU(True): slope 1 */, nullptr, nullptr, w.mid1,
                 hh.hid * px, st))) return rc;
  const int Co = hh.hid / 2;
  for (int ph = 0; ph < 4; ++ph) {
    const int py = ph >> 1, px_ = ph & 1;
    GemmOp op = base_op(hh.dc[ph], 0.f /* ReLU */);
    op.mode = 1; op.A = w.mid1; op.M = bs * roi * roi; op.Cin = hh.hid; op.H = roi; op.W = roi; op.in_img_stride = hh.hid * px;
    op.ntap = 4;
    const int dys[2] = {py ? 1 : 0, py ? 0 : -1}, dxs[2] = {px_ ? 1 : 0, px_ ? 0 : -1};
    for (int t = 0; t < 4; ++t) { op.dy[t] = (signed char)dys[t >> 1]; op.dx[t] = (signed char)dxs[t & 1]; }
    op.C = w.dec; op.os = 2; op.py = py; op.px = px_; op.out_img_stride = (long long)Co * px * 4;
    if ((rc = run_gemm(op, st))) return rc;
  }
  return conv(hh.fin, 1, w.dec, (long long)Co * px * 4, bs, Co, 2 * roi, 2 * roi, 1.f, nullptr, nullptr, out, (long long)hh.J * px * 4, st);
}

// Encoder.forward (encoding.py:58-73) on the concatenated input in w.cat; pooled map of block 1 -> pool1, flattened output -> enc
int run_encoder(const EncoderW& e, int bs, int roi, const Ws& w, float* pool1, float* enc, cudaStream_t st) {
  int rc, H = roi;
  const int hid = e.hid, mid = e.hid / 2;
  auto other = [&](const float* cur) { return cur == w.xa ? w.xb : w.xa; };
  const float* x = w.xa;
  if ((rc = conv(e.project, 1, w.cat, (long long)e.cin * H * H, bs, e.cin, H, H, 1.f, nullptr, nullptr, w.xa, (long long)hid * H * H, st))) return rc;
  for (int blk = 0; blk < 4; ++blk) {
    const long long s = (long long)H * H;
    for (int j = 0; j < 2; ++j) {
      const ResBlock& r = e.reg[blk * 2 + j];
      float* y = other(x);
      if ((rc = conv(r.c1, 1, x, hid * s, bs, hid, H, H, 0.01f, &r.pre, nullptr, w.mid1, mid * s, st))) return rc;
      if ((rc = conv(r.c2, 3, w.mid1, mid * s, bs, mid, H, H, 0.01f, nullptr, nullptr, w.mid2, mid * s, st))) return rc;
      if ((rc = conv(r.c3, 1, w.mid2, mid * s, bs, mid, H, H, 1.f, nullptr, x, y, hid * s, st))) return rc;
      x = y;
    }
    float* dst = blk == 3 ? enc : (blk == 1 ? pool1 : other(x));      // enc_ls[1] is kept for the cross modules
    const long long total = (long long)bs * hid * (H / 2) * (H / 2);
    VPHO_LAUNCH(k_maxpool2, dim3(grid_for(total)), dim3(256), 0, st, x, (long long)bs * hid, H, H, dst);
    VPHO_CHECK_LAUNCH();
    x = dst;
    H /= 2;
  }
  return VPHO_OK;
}

// CrossModule.forward (cross_module.py:119-137); out_sel 0: y_hand rows, 1: y_obj rows -> result [bs][32][d]
int run_cross(const vpho_heads* h, const CrossW& c, const vpho_heads_args* a, const Ws& w, int sel, float* out, cudaStream_t st) {
  const int bs = a->bs, H = a->roi_size / 4, F = h->n_force, ntok = 2 * F + 1, d = c.d, rows = bs * ntok;
  const long long tok_stride = (long long)ntok * d, s = (long long)H * H;
  int rc;
  if ((rc = conv(c.proj_hand, 3, w.pool1_hand, h->enc_hid * s, bs, h->enc_hid, H, H, 1.f, nullptr, nullptr, w.tok, tok_stride, st))) return rc;
  if ((rc = conv(c.proj_obj, 3, w.pool1_obj, h->enc_hid * s, bs, h->enc_hid, H, H, 1.f, nullptr, nullptr, w.tok + (long long)F * d, tok_stride, st))) return rc;
  VPHO_LAUNCH(k_gravity_pe, dim3(bs, 8), dim3(256), 0, st, a->gravity, a->is_right, c.Wg, c.bg, c.pe, ntok, d, w.tok);
  VPHO_CHECK_LAUNCH();
  if ((rc = linear(c.in_proj, w.tok, rows, d, 1.f, nullptr, w.qkv, 3 * d, st))) return rc;
  const int nhead = 2;
  if ((rc = run_attention(w.qkv, bs, ntok, d, nhead, w.att, st))) return rc;
  if ((rc = linear(c.out_proj, w.att, rows, d, 1.f, w.tok, w.tmp, d, st))) return rc;           // x + sa(x)
  VPHO_LAUNCH(k_layernorm, dim3(rows), dim3(128), 0, st, w.tmp, c.n1w, c.n1b, d, w.tok);
  VPHO_CHECK_LAUNCH();
  if ((rc = linear(c.lin1, w.tok, rows, d, 0.f, nullptr, w.ffn, c.ff, st))) return rc;
  if ((rc = linear(c.lin2, w.ffn, rows, c.ff, 1.f, w.tok, w.tmp, d, st))) return rc;             // x + ff(x)
  VPHO_LAUNCH(k_layernorm, dim3(rows), dim3(128), 0, st, w.tmp, c.n2w, c.n2b, d, w.att);
  VPHO_CHECK_LAUNCH();
  // rows [sel * F, sel * F + F) of every image
  const long long n = (long long)bs * F * d;
  VPHO_LAUNCH(k_copy_cols, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, w.att, bs, ntok * d, sel * F * d, F * d, out);
  VPHO_CHECK_LAUNCH();
  return VPHO_OK;
}


#ifndef VPHO_EMU
// ---------------------------------------------------------------------------------------------------------------------------
// tensor-core path (producers_tc.cu): the same wiring over NHWC FP16 (hi, lo) planes
// ---------------------------------------------------------------------------------------------------------------------------
struct Planes {
  __half* hi;
  __half* lo;
};
Planes planes_of(unsigned short* p, size_t n_plane) { return {reinterpret_cast<__half*>(p), reinterpret_cast<__half*>(p) + n_plane}; }

TcGemm tc_base(const vpho_heads* h, const DevMat& w, float slope) {
  TcGemm p = {};
  p.chunks = w.tc_cp / 64;
  p.ntap = w.tc_ntap;
  p.bias = w.bias;
  p.post_scale = w.post_scale;
  p.post_shift = w.post_shift;
  p.slope = slope;
  p.os = 1;
  p.overflow_flag = h->overflow;
  return p;
}

TcGemm tc_conv_op(const vpho_heads* h, const DevMat& w, int ks, int bs, int H, int W, float slope) {
  TcGemm p = tc_base(h, w, slope);
  p.mode = 1;
  p.n_img = bs; p.H = H; p.W = W;
  for (int t = 0; t < ks * ks; ++t) { p.dy[t] = (signed char)(t / ks - ks / 2); p.dx[t] = (signed char)(t % ks - ks / 2); }
  return p;
}

int run_heat_tc(const vpho_heads* h, const HeatHead& hh, const float* feat, const unsigned char* is_right, int bs, int roi, const Ws& w, float* out,
                cudaStream_t st) {
  int rc;
  const size_t px = (size_t)roi * roi, b = (size_t)bs;
  const Planes F = planes_of(w.pF, b * px * h->C), M1 = planes_of(w.pM1, b * px * hh.hid), M2 = planes_of(w.pM2, b * px * hh.hid);
  const Planes DEC = planes_of(w.pDEC, b * px * 4 * (hh.hid / 2));
  if ((rc = pt_to_planes(feat, nullptr, nullptr, nullptr, is_right, 0, 0, bs, h->C, 0, roi, h->C, F.hi, F.lo, st))) return rc;
  TcGemm p = tc_conv_op(h, hh.c0, 3, bs, roi, roi, 1.f);
  p.out_hi = M1.hi; p.out_lo = M1.lo; p.out_cp = hh.hid;
  if ((rc = pt_gemm(hh.c0.tc, p, F.hi, F.lo, st))) return rc;
  p = tc_conv_op(h, hh.c1, 3, bs, roi, roi, 1.f /* LeakyReLU(True): slope 1 */);
  p.out_hi = M2.hi; p.out_lo = M2.lo; p.out_cp = hh.hid;
  if ((rc = pt_gemm(hh.c1.tc, p, M1.hi, M1.lo, st))) return rc;
  for (int ph = 0; ph < 4; ++ph) {
    const int py = ph >> 1, px_ = ph & 1;
    p = tc_base(h, hh.dc[ph], 0.f /* ReLU */);
    p.mode = 1; p.n_img = bs; p.H = roi; p.W = roi;
    const int dys[2] = {py ? 1 : 0, py ? 0 : -1}, dxs[2] = {px_ ? 1 : 0, px_ ? 0 : -1};
    for (int t = 0; t < 4; ++t) { p.dy[t] = (signed char)dys[t >> 1]; p.dx[t] = (signed char)dxs[t & 1]; }
    p.os = 2; p.py = py; p.px = px_;
    p.out_hi = DEC.hi; p.out_lo = DEC.lo; p.out_cp = hh.hid / 2;
    if ((rc = pt_gemm(hh.dc[ph].tc, p, M2.hi, M2.lo, st))) return rc;
  }
  p = tc_conv_op(h, hh.fin, 1, bs, 2 * roi, 2 * roi, 1.f);
  p.out_f32 = out; p.f32_img_stride = (long long)hh.J * px * 4;
  return pt_gemm(hh.fin.tc, p, DEC.hi, DEC.lo, st);
}

int run_encoder_tc(const vpho_heads* h, const EncoderW& e, const float* feat, const float* hm, const float* bbox, const float* bbox_rect,
                   const unsigned char* is_right, int flip, int J, int bs, int roi, const Ws& w, unsigned short* pool1, float* enc, cudaStream_t st) {
  int rc, H = roi;
  const int hid = e.hid, mid = e.hid / 2, Cp = round_up(h->C + J, 64);
  const size_t b = (size_t)bs;
  const Planes F = planes_of(w.pF, b * roi * roi * Cp);
  if ((rc = pt_to_planes(feat, hm, bbox, bbox_rect, is_right, flip, flip, bs, h->C, J, roi, Cp, F.hi, F.lo, st))) return rc;
  auto other = [&](unsigned short* cur) { return cur == w.pX0 ? w.pX1 : w.pX0; };
  unsigned short* xb = w.pX0;
  {
    const size_t n = b * H * H * hid;
    const Planes X = planes_of(xb, n), XA = planes_of(w.pXA, n);
    TcGemm p = tc_conv_op(h, e.project, 1, bs, H, H, 1.f);
    p.out_hi = X.hi; p.out_lo = X.lo; p.out_cp = hid;
    p.out2_hi = XA.hi; p.out2_lo = XA.lo; p.pre_scale = e.reg[0].pre.scale; p.pre_shift = e.reg[0].pre.shift; p.pre_slope = 0.01f;
    if ((rc = pt_gemm(e.project.tc, p, F.hi, F.lo, st))) return rc;
  }
  for (int blk = 0; blk < 4; ++blk) {
    const size_t n = b * H * H * hid, nm = b * H * H * mid;
    for (int j = 0; j < 2; ++j) {
      const ResBlock& r = e.reg[blk * 2 + j];
      unsigned short* yb = other(xb);
      const Planes X = planes_of(xb, n), Y = planes_of(yb, n), XA = planes_of(w.pXA, n), M1 = planes_of(w.pM1, nm), M2 = planes_of(w.pM2, nm);
      TcGemm p = tc_conv_op(h, r.c1, 1, bs, H, H, 0.01f);
      p.out_hi = M1.hi; p.out_lo = M1.lo; p.out_cp = mid;
      if ((rc = pt_gemm(r.c1.tc, p, XA.hi, XA.lo, st))) return rc;
      p = tc_conv_op(h, r.c2, 3, bs, H, H, 0.01f);
      p.out_hi = M2.hi; p.out_lo = M2.lo; p.out_cp = mid;
      if ((rc = pt_gemm(r.c2.tc, p, M1.hi, M1.lo, st))) return rc;
      p = tc_conv_op(h, r.c3, 1, bs, H, H, 1.f);
      p.res_hi = X.hi; p.res_lo = X.lo;
      p.out_hi = Y.hi; p.out_lo = Y.lo; p.out_cp = hid;
      if (j == 0) {        // the second module of the block starts with its own BatchNorm + LeakyReLU
        const ResBlock& nx = e.reg[blk * 2 + 1];
        p.out2_hi = XA.hi; p.out2_lo = XA.lo; p.pre_scale = nx.pre.scale; p.pre_shift = nx.pre.shift; p.pre_slope = 0.01f;
      }
      if ((rc = pt_gemm(r.c3.tc, p, M2.hi, M2.lo, st))) return rc;
      xb = yb;
    }
    const Planes X = planes_of(xb, n);
    const size_t no = n / 4;
    if (blk == 3) {
      if ((rc = pt_pool(X.hi, X.lo, bs, H, H, hid, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0.f, enc, st))) return rc;
    } else {
      unsigned short* db = blk == 1 ? pool1 : other(xb);      // enc_ls[1] is kept for the cross modules
      const Planes D = planes_of(db, no), XA = planes_of(w.pXA, no);
      const ResBlock& nx = e.reg[(blk + 1) * 2];
      if ((rc = pt_pool(X.hi, X.lo, bs, H, H, hid, D.hi, D.lo, XA.hi, XA.lo, nx.pre.scale, nx.pre.shift, 0.01f, nullptr, st))) return rc;
      xb = db;
    }
    H /= 2;
  }
  return VPHO_OK;
}

int run_cross_tc(const vpho_heads* h, const CrossW& c, const vpho_heads_args* a, const Ws& w, int sel, float* out, cudaStream_t st) {
  const int bs = a->bs, H = a->roi_size / 4, F = h->n_force, ntok = 2 * F + 1, d = c.d, rows = bs * ntok;
  const size_t np1 = (size_t)bs * H * H * h->enc_hid, nt = (size_t)rows * d, nf = (size_t)rows * c.ff;
  const Planes P1h = planes_of(w.pP1h, np1), P1o = planes_of(w.pP1o, np1), TOK = planes_of(w.pTOK, nt), ATT = planes_of(w.pATT, nt),
               FFN = planes_of(w.pFFN, nf);
  int rc;
  TcGemm p = tc_conv_op(h, c.proj_hand, 3, bs, H, H, 1.f);
  p.out_f32 = w.tok; p.f32_img_stride = (long long)ntok * d;
  if ((rc = pt_gemm(c.proj_hand.tc, p, P1h.hi, P1h.lo, st))) return rc;
  p = tc_conv_op(h, c.proj_obj, 3, bs, H, H, 1.f);
  p.out_f32 = w.tok + (long long)F * d; p.f32_img_stride = (long long)ntok * d;
  if ((rc = pt_gemm(c.proj_obj.tc, p, P1o.hi, P1o.lo, st))) return rc;
  VPHO_LAUNCH(k_gravity_pe, dim3(bs, 8), dim3(256), 0, st, a->gravity, a->is_right, c.Wg, c.bg, c.pe, ntok, d, w.tok);
  VPHO_CHECK_LAUNCH();
  auto rows_op = [&](const DevMat& m, float slope) {
    TcGemm q = tc_base(h, m, slope);
    q.mode = 0; q.M = rows;
    return q;
  };
  if ((rc = pt_split_rows(w.tok, (long long)nt, TOK.hi, TOK.lo, st))) return rc;
  p = rows_op(c.in_proj, 1.f);
  p.out_f32 = w.qkv; p.ldc = 3 * d;
  if ((rc = pt_gemm(c.in_proj.tc, p, TOK.hi, TOK.lo, st))) return rc;
  const int nhead = 2;
  if ((rc = run_attention(w.qkv, bs, ntok, d, nhead, w.att, st))) return rc;
  if ((rc = pt_split_rows(w.att, (long long)nt, ATT.hi, ATT.lo, st))) return rc;
  p = rows_op(c.out_proj, 1.f);
  p.res_f32 = w.tok; p.out_f32 = w.tmp; p.ldc = d;                                             // x + sa(x)
  if ((rc = pt_gemm(c.out_proj.tc, p, ATT.hi, ATT.lo, st))) return rc;
  VPHO_LAUNCH(k_layernorm, dim3(rows), dim3(128), 0, st, w.tmp, c.n1w, c.n1b, d, w.tok);
  VPHO_CHECK_LAUNCH();
  if ((rc = pt_split_rows(w.tok, (long long)nt, TOK.hi, TOK.lo, st))) return rc;
  p = rows_op(c.lin1, 0.f);
  p.out_hi = FFN.hi; p.out_lo = FFN.lo; p.out_cp = c.ff;
  if ((rc = pt_gemm(c.lin1.tc, p, TOK.hi, TOK.lo, st))) return rc;
  p = rows_op(c.lin2, 1.f);
  p.res_f32 = w.tok; p.out_f32 = w.tmp; p.ldc = d;                                             // x + ff(x)
  if ((rc = pt_gemm(c.lin2.tc, p, FFN.hi, FFN.lo, st))) return rc;
  VPHO_LAUNCH(k_layernorm, dim3(rows), dim3(128), 0, st, w.tmp, c.n2w, c.n2b, d, w.att);
  VPHO_CHECK_LAUNCH();
  const long long n = (long long)bs * F * d;
  VPHO_LAUNCH(k_copy_cols, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, w.att, bs, ntok * d, sel * F * d, F * d, out);
  VPHO_CHECK_LAUNCH();
  return VPHO_OK;
}
#endif  // !VPHO_EMU

}  // namespace

extern "C" int vpho_heads_create(const vpho_named_tensor* tensors, int n_tensors, vpho_heads_t* out) {
  if (!tensors || n_tensors <= 0 || !out) return VPHO_ERR_INVALID;
  Table T;
  for (int i = 0; i < n_tensors; ++i) {
    if (!tensors[i].name || tensors[i].ndim < 0 || tensors[i].ndim > 4) return VPHO_ERR_INVALID;
    T.m[tensors[i].name] = &tensors[i];
  }
  vpho_heads* h = new vpho_heads();
  int C2 = 0, out_n, in_n;
#ifndef VPHO_EMU
  {   // tensor-core path: every GEMM K (input channels per tap) and every planes output width must be a multiple of 64
    auto dim = [&](const char* k, int i) { auto it = T.m.find(k); return it == T.m.end() || it->second->ndim <= i ? 1 : (int)it->second->shape[i]; };
    const int hh = dim("head_hm_hand.conv_layers.0.weight", 0), Cc = dim("head_hm_hand.conv_layers.0.weight", 1);
    const int eh = dim("encoder_hand.project.weight", 0), dm = dim("cross_hand.gravity_proj.weight", 0);
    const int ff = dim("cross_hand.attn.layers.0.linear1.weight", 0);
    const int m1 = dim("head_mano.base_layer.0.weight", 0), m0 = dim("head_mano.base_layer.0.weight", 1), m2 = dim("head_mano.base_layer.2.weight", 0);
    const int phd = dim("head_physics.fc_scale.0.weight", 0);
    h->tc_ok = pt_available() && Cc % 64 == 0 && hh % 128 == 0 && eh % 128 == 0 && dm % 64 == 0 && ff % 64 == 0 && m0 % 64 == 0 && m1 % 64 == 0 &&
               m2 % 64 == 0 && phd % 64 == 0;
    if (cudaMalloc((void**)&h->overflow, sizeof(int)) != cudaSuccess) { delete h; return VPHO_ERR_ALLOC; }
    h->allocs.push_back(h->overflow);
    cudaMemset(h->overflow, 0, sizeof(int));
  }
#endif
  bool ok = build_heat(h, T, "head_hm_hand", h->hm_hand, h->C) && build_heat(h, T, "head_hm_obj", h->hm_obj, C2) && C2 == h->C &&
            h->hm_hand.hid == h->hm_obj.hid;
  ok = ok && build_encoder(h, T, "encoder_hand", h->enc_hand) && build_encoder(h, T, "encoder_obj", h->enc_obj);
  if (ok) {
    h->Jh = h->hm_hand.J; h->Jo = h->hm_obj.J; h->heat_hid = h->hm_hand.hid; h->enc_hid = h->enc_hand.hid;
    ok = h->enc_hand.cin == h->C + h->Jh && h->enc_obj.cin == h->C + h->Jo && h->enc_obj.hid == h->enc_hid;
  }
  ok = ok && linear_mat(h, T, "head_mano.base_layer.0", h->mano0, h->mano_h1, h->enc_dim, true) &&
       linear_mat(h, T, "head_mano.base_layer.2", h->mano1, h->mano_h2, in_n, true) && in_n == h->mano_h1;
  if (ok) {      // fc_pose (96) and fc_shape (10) share their input: one [h2][106] matrix
    const vpho_named_tensor *wp = T.get("head_mano.fc_pose.weight", 2), *bp = T.get("head_mano.fc_pose.bias", 1),
                            *wsh = T.get("head_mano.fc_shape.weight", 2), *bsh = T.get("head_mano.fc_shape.bias", 1);
    ok = wp && bp && wsh && bsh && wp->shape[0] == 96 && wsh->shape[0] == 10 && wp->shape[1] == h->mano_h2 && wsh->shape[1] == h->mano_h2;
    if (ok) {
      std::vector<float> bias(106);
      for (int i = 0; i < 96; ++i) bias[i] = bp->data[i];
      for (int i = 0; i < 10; ++i) bias[96 + i] = bsh->data[i];
      const float *P = wp->data, *S = wsh->data;
      const int k2 = h->mano_h2;
      ok = make_mat(h, h->mano_out, k2, 106, [=](int k, int n) { return n < 96 ? P[(size_t)n * k2 + k] : S[(size_t)(n - 96) * k2 + k]; }, bias.data(), 1, true);
    }
  }
  ok = ok && build_cross(h, T, "cross_hand", h->cross_hand, h->enc_hid) && build_cross(h, T, "cross_obj", h->cross_obj, h->enc_hid);
  if (ok) {
    h->d_model = h->cross_hand.d;
    ok = h->cross_obj.d == h->d_model && h->d_model % 2 == 0;
  }
  ok = ok && linear_mat(h, T, "head_physics.fc_scale.0", h->phys_scale0, h->phys_hid, in_n, true) && in_n == h->d_model &&
       linear_mat(h, T, "head_physics.fc_weight.0", h->phys_weight0, out_n, in_n, true) && out_n == h->phys_hid && in_n == h->d_model &&
       linear_mat(h, T, "head_physics.fc_CoM.0", h->phys_com0, out_n, in_n, true) && out_n == h->phys_hid && in_n == h->d_model &&
       linear_mat(h, T, "head_physics.fc_scale.2", h->phys_scale2, out_n, in_n, true) && out_n == 1 && in_n == h->phys_hid &&
       linear_mat(h, T, "head_physics.fc_weight.2", h->phys_weight2, out_n, in_n, true) && out_n == 8 && in_n == h->phys_hid &&
       linear_mat(h, T, "head_physics.fc_CoM.2", h->phys_com2, out_n, in_n, true) && out_n == 3 && in_n == h->phys_hid;
  if (ok) {
    const vpho_named_tensor* an = T.get("head_physics.anchor", 2);
    ok = an && an->shape[0] == 8 && an->shape[1] == 3;
    if (ok) h->anchor = upload(h, std::vector<float>(an->data, an->data + 24));
    ok = ok && h->anchor;
  }
  if (!ok || !T.ok || cudaGetLastError() != cudaSuccess) {
    vpho_heads_destroy(h);
    return VPHO_ERR_INVALID;
  }
  *out = h;
  return VPHO_OK;
}

extern "C" int vpho_heads_destroy(vpho_heads_t h) {
  if (!h) return VPHO_OK;
  for (void* p : h->allocs) cudaFree(p);
  delete h;
  return VPHO_OK;
}

extern "C" int vpho_heads_dims(vpho_heads_t h, int32_t* dims) {
  if (!h || !dims) return VPHO_ERR_INVALID;
  const int v[8] = {h->C, h->Jh, h->Jo, h->enc_dim, h->d_model, h->n_force, h->heat_hid, h->enc_hid};
  for (int i = 0; i < 8; ++i) dims[i] = v[i];
  return VPHO_OK;
}

extern "C" size_t vpho_heads_workspace_bytes(vpho_heads_t h, int bs, int roi_size) {
  if (!h || bs <= 0 || roi_size <= 0) return 0;
  return carve(h, bs, roi_size, nullptr).bytes;
}

extern "C" int vpho_heads_forward(vpho_heads_t h, const vpho_heads_args* a, void* workspace, size_t workspace_bytes, void* stream) {
  if (!h || !a || a->bs < 0) return VPHO_ERR_INVALID;
  if (a->bs == 0) return VPHO_OK;
  const int bs = a->bs, roi = a->roi_size;
  // four 2x2 poolings down to (roi/16)^2, the second pooled map (roi/4)^2 must hold n_force tokens of d_model
  if (roi < 16 || roi % 16 || h->enc_hid * (roi / 16) * (roi / 16) != h->enc_dim) return VPHO_ERR_INVALID;
  if (h->cross_hand.proj_dim * (roi / 4) * (roi / 4) != h->n_force * h->d_model || bs > h->cross_hand.pe_rows) return VPHO_ERR_INVALID;
  if (!a->hf_hr || !a->of_or_rect || !a->hf_hr_rect || !a->bbox_hand || !a->bbox_hand_rect || !a->bbox_obj || !a->bbox_obj_rect ||
      !a->is_right || !a->gravity || !a->hand_heatmap || !a->obj_heatmap || !a->encoding_hand || !a->encoding_obj || !a->mano_pose ||
      !a->mano_shape || !a->force_local || !workspace)
    return VPHO_ERR_INVALID;
  const Ws w = carve(h, bs, roi, workspace);
  if (workspace_bytes < w.bytes) return VPHO_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  const bool strict = (a->flags & VPHO_HEADS_STRICT_FP32) != 0;
#ifdef VPHO_EMU
  if (!strict) return VPHO_ERR_INVALID;         // the emulator build has no tensor cores: ask for the FP32 SIMT path explicitly
#else
  if (!strict && !h->tc_ok) return VPHO_ERR_INVALID;   // widths not multiples of 64 (or no TMA descriptor encoder): no silent downgrade
  if (!strict) {
    if (cudaMemsetAsync(h->overflow, 0, sizeof(int), st) != cudaSuccess) return VPHO_ERR_LAUNCH;
    if ((rc = run_heat_tc(h, h->hm_hand, a->hf_hr, a->is_right, bs, roi, w, a->hand_heatmap, st))) return rc;
    if ((rc = run_heat_tc(h, h->hm_obj, a->of_or_rect, a->is_right, bs, roi, w, a->obj_heatmap, st))) return rc;
    if ((rc = run_encoder_tc(h, h->enc_hand, a->hf_hr_rect, a->hand_heatmap, a->bbox_hand, a->bbox_hand_rect, a->is_right, 0, h->Jh, bs, roi, w,
                             w.pP1h, a->encoding_hand, st))) return rc;
    if ((rc = run_encoder_tc(h, h->enc_obj, a->of_or_rect, a->obj_heatmap, a->bbox_obj, a->bbox_obj_rect, a->is_right, 1, h->Jo, bs, roi, w,
                             w.pP1o, a->encoding_obj, st))) return rc;
  } else
#endif
  {
  // heat-maps (VPHO.py:129-130)
  if ((rc = run_heat(h->hm_hand, a->hf_hr, bs, h->C, roi, w, a->hand_heatmap, st))) return rc;
  if ((rc = run_heat(h->hm_obj, a->of_or_rect, bs, h->C, roi, w, a->obj_heatmap, st))) return rc;
  // hand encoder (VPHO.py:132-150): features of the square box + re-aligned, resized heat-maps
  const long long tot_h = (long long)bs * (h->C + h->Jh) * roi * roi, tot_o = (long long)bs * (h->C + h->Jo) * roi * roi;
  VPHO_LAUNCH(k_encoder_input, dim3(grid_for(tot_h)), dim3(256), 0, st, a->hf_hr_rect, a->hand_heatmap, a->bbox_hand, a->bbox_hand_rect,
              a->is_right, 0, 0, bs, h->C, h->Jh, roi, w.cat);
  VPHO_CHECK_LAUNCH();
  if ((rc = run_encoder(h->enc_hand, bs, roi, w, w.pool1_hand, a->encoding_hand, st))) return rc;
  // object encoder: features and heat-maps flipped back for left hands (VPHO.py:138-151)
  VPHO_LAUNCH(k_encoder_input, dim3(grid_for(tot_o)), dim3(256), 0, st, a->of_or_rect, a->obj_heatmap, a->bbox_obj, a->bbox_obj_rect,
              a->is_right, 1, 1, bs, h->C, h->Jo, roi, w.cat);
  VPHO_CHECK_LAUNCH();
  if ((rc = run_encoder(h->enc_obj, bs, roi, w, w.pool1_obj, a->encoding_obj, st))) return rc;
  }
  // regression head (head_mano.py:61-76)
#ifndef VPHO_EMU
  if (!strict) {
    const Planes A0 = planes_of(w.pTOK, (size_t)bs * h->enc_dim), A1 = planes_of(w.pATT, (size_t)bs * h->mano_h1),
                 A2 = planes_of(w.pFFN, (size_t)bs * h->mano_h2);
    if ((rc = pt_split_rows(a->encoding_hand, (long long)bs * h->enc_dim, A0.hi, A0.lo, st))) return rc;
    TcGemm p = tc_base(h, h->mano0, 0.01f);
    p.mode = 0; p.M = bs; p.out_hi = A1.hi; p.out_lo = A1.lo; p.out_cp = h->mano_h1;
    if ((rc = pt_gemm(h->mano0.tc, p, A0.hi, A0.lo, st))) return rc;
    p = tc_base(h, h->mano1, 0.01f);
    p.mode = 0; p.M = bs; p.out_hi = A2.hi; p.out_lo = A2.lo; p.out_cp = h->mano_h2;
    if ((rc = pt_gemm(h->mano1.tc, p, A1.hi, A1.lo, st))) return rc;
    p = tc_base(h, h->mano_out, 1.f);
    p.mode = 0; p.M = bs; p.out_f32 = w.mano_o; p.ldc = 106;
    if ((rc = pt_gemm(h->mano_out.tc, p, A2.hi, A2.lo, st))) return rc;
  } else
#endif
  {
  if ((rc = linear(h->mano0, a->encoding_hand, bs, h->enc_dim, 0.01f, nullptr, w.mano_a, h->mano_h1, st))) return rc;
  if ((rc = linear(h->mano1, w.mano_a, bs, h->mano_h1, 0.01f, nullptr, w.mano_b, h->mano_h2, st))) return rc;
  if ((rc = linear(h->mano_out, w.mano_b, bs, h->mano_h2, 1.f, nullptr, w.mano_o, 106, st))) return rc;
  }
  VPHO_LAUNCH(k_rot6d_rows, dim3((bs * 16 + 255) / 256), dim3(256), 0, st, w.mano_o, bs * 16, 106, a->mano_pose);
  VPHO_CHECK_LAUNCH();
  VPHO_LAUNCH(k_copy_cols, dim3((bs * 10 + 255) / 256), dim3(256), 0, st, w.mano_o, bs, 106, 96, 10, a->mano_shape);
  VPHO_CHECK_LAUNCH();
  // cross modules + physics head (VPHO.py:174-176): hand tokens of cross_hand, object tokens of cross_obj
  const int F = h->n_force, rows = bs * F, d = h->d_model, ph = h->phys_hid;
  auto cross = [&](const CrossW& c, int sel) {
#ifndef VPHO_EMU
    if (!strict) return run_cross_tc(h, c, a, w, sel, w.ph_tok, st);
#endif
    return run_cross(h, c, a, w, sel, w.ph_tok, st);
  };
  if ((rc = cross(h->cross_hand, 0))) return rc;
  // one two-layer MLP of HeadPhysics on the tokens in w.ph_tok -> columns [col, col + N) of the [rows][12] head buffer
  auto mlp = [&](const DevMat& l0, const DevMat& l2, int col) -> int {
    int r;
#ifndef VPHO_EMU
    if (!strict) {
      const Planes A0 = planes_of(w.pTOK, (size_t)rows * d), A1 = planes_of(w.pATT, (size_t)rows * ph);
      if ((r = pt_split_rows(w.ph_tok, (long long)rows * d, A0.hi, A0.lo, st))) return r;
      TcGemm p = tc_base(h, l0, 0.01f);
      p.mode = 0; p.M = rows; p.out_hi = A1.hi; p.out_lo = A1.lo; p.out_cp = ph;
      if ((r = pt_gemm(l0.tc, p, A0.hi, A0.lo, st))) return r;
      p = tc_base(h, l2, 1.f);
      p.mode = 0; p.M = rows; p.out_f32 = w.ph_head + col; p.ldc = 12;
      return pt_gemm(l2.tc, p, A1.hi, A1.lo, st);
    }
#endif
    if ((r = linear(l0, w.ph_tok, rows, d, 0.01f, nullptr, w.ph_mid, ph, st))) return r;
    return linear(l2, w.ph_mid, rows, ph, 1.f, nullptr, w.ph_head + col, 12, st);
  };
  // fc_scale on the hand tokens (physics.py:704)
  if ((rc = mlp(h->phys_scale0, h->phys_scale2, 0))) return rc;
  if (a->enc_phy_hand &&
      cudaMemcpyAsync(a->enc_phy_hand, w.ph_tok, (size_t)rows * d * sizeof(float), cudaMemcpyDeviceToDevice, st) != cudaSuccess) return VPHO_ERR_LAUNCH;
  if ((rc = cross(h->cross_obj, 1))) return rc;
  if (a->enc_phy_obj &&
      cudaMemcpyAsync(a->enc_phy_obj, w.ph_tok, (size_t)rows * d * sizeof(float), cudaMemcpyDeviceToDevice, st) != cudaSuccess) return VPHO_ERR_LAUNCH;
  // fc_weight / fc_CoM on the object tokens (physics.py:706-710)
  if ((rc = mlp(h->phys_weight0, h->phys_weight2, 1))) return rc;
  if ((rc = mlp(h->phys_com0, h->phys_com2, 9))) return rc;
  VPHO_LAUNCH(k_physics_tail, dim3((rows + 127) / 128), dim3(128), 0, st, w.ph_head, h->anchor, rows, a->force_local, a->force_scale,
              a->force_weight, a->CoM);
  VPHO_CHECK_LAUNCH();
  return VPHO_OK;
}

// 1 when an activation of the last tensor-core forward left the FP16 range of the operand planes (|v| >= 60000): its results
// are then invalid and the caller should re-run with VPHO_HEADS_STRICT_FP32.  Synchronises the stream.
extern "C" int vpho_heads_overflow(vpho_heads_t h, int32_t* flag, void* stream) {
  if (!h || !flag) return VPHO_ERR_INVALID;
  *flag = 0;
#ifndef VPHO_EMU
  int v = 0;
  if (cudaMemcpyAsync(&v, h->overflow, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream) != cudaSuccess) return VPHO_ERR_LAUNCH;
  if (cudaStreamSynchronize((cudaStream_t)stream) != cudaSuccess) return VPHO_ERR_LAUNCH;
  *flag = v;
#endif
  return VPHO_OK;
}
