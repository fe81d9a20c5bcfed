// Tensor-core path of the N1 producers (see producers.cu for the module map and reference citations).
//
// Every dense layer of the heat-map heads, the encoders, the cross modules' projections and the transformer layer is one launch
// of k_gemm_tc: D[128 pixels or rows][BN outputs] accumulated in TMEM by tcgen05.mma (cta_group::1, kind::f16) over
// K chunks of 64, with the FP32-parity split of the score network (operands as FP16 hi + lo planes, three UMMAs per K step:
// A_lo B_hi + A_hi B_lo + A_hi B_hi, ~2^-22 relative).  Convolutions are IMPLICIT GEMMs: activations live as channels-last
// (NHWC) __half planes and the A tile of tap (dy, dx) is one 4-D TMA box {64 channels, W, bh rows, bn images} whose start is
// shifted by (dx, dy) -- the zero padding of the convolution is TMA's out-of-bounds fill, nothing is materialised.  The four
// output phases of the stride-2 transposed convolution are the same kernel with 2 x 2 taps and a strided scatter in the
// epilogue.  Epilogue (4 warps, thread = TMEM lane = pixel): un-scale, bias, BatchNorm affine, LeakyReLU / ReLU, residual add,
// then any of: hi/lo planes for the next layer, a second pre-activated copy (the BatchNorm + LeakyReLU that opens the next
// `Residual`), float32 NCHW (heat-maps, tokens) or row-major (transformer).
#include "vpho_common.cuh"
#include "vpho_b200.h"
#include "tc_ptx.cuh"
#include "producers_tc.cuh"

#include <cuda_fp16.h>

namespace vpho {

__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, void* bar, void* dst, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
                   smem_u32(dst)),
               "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}

// v = hi + lo * 2^-11: the residual of the FP16 rounding is stored scaled by 2^11, so that it stays a NORMAL half whenever v
// is one (an unscaled residual of |v| < 0.25 is a subnormal half and loses its low bits).  The cross products A_lo B_hi and
// A_hi B_lo are accumulated in a TMEM accumulator of their own and folded in with the factor 2^-11 in the epilogue.
constexpr float kLoScale = 2048.f, kLoInv = 1.f / 2048.f;
__device__ __forceinline__ void split_half(float v, __half& hi, __half& lo) {
  hi = __float2half_rn(v);
  lo = __float2half_rn((v - __half2float(hi)) * kLoScale);
}
__device__ __forceinline__ float join_half(__half hi, __half lo) { return fmaf(__half2float(lo), kLoInv, __half2float(hi)); }

template <int BN, int NST>
struct PtSmemLayout {
  static constexpr int kA = 128 * 128;                  // one plane of the A tile: 128 rows x 64 halves
  static constexpr int kB = BN * 128;
  static constexpr int kStage = 2 * kA + 2 * kB;
  static constexpr int kStages = NST;
  static constexpr int kParams = 5 * BN * 4;             // bias, post scale / shift, pre scale / shift of the CTA's columns
  static constexpr int kBytes = kStages * kStage + 1024 /* alignment */ + 256 /* barriers */ + kParams;
};

// NST = 3: deep pipeline, one CTA per SM (long K: 3x3 convolutions).  NST = 1: 64 KB of shared memory, two CTAs per SM whose
// load / MMA / epilogue phases interleave (short K: 1x1 convolutions, whose time is the epilogue's).
template <int BN, int NST>
__global__ void __launch_bounds__(256, NST == 1 ? 2 : 1)
k_gemm_tc(const __grid_constant__ CUtensorMap mA_hi, const __grid_constant__ CUtensorMap mA_lo,
          const __grid_constant__ CUtensorMap mB_hi, const __grid_constant__ CUtensorMap mB_lo, const TcGemm p) {
  using L = PtSmemLayout<BN, NST>;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(base + L::kStages * L::kStage);
  unsigned long long* full = bars;
  unsigned long long* empty = bars + L::kStages;
  unsigned long long* acc_full = bars + 2 * L::kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * L::kStages + 1);
  float* par = reinterpret_cast<float*>(base + L::kStages * L::kStage + 256);     // [5][BN]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int mt = blockIdx.x, n0 = blockIdx.y * BN;
  constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  constexpr uint32_t kCols = 2 * BN;                      // accumulator 0: A_hi B_hi; accumulator 1: the two cross products

  // tile origin of a convolution: bn images x bh rows x W columns = 128 pixels
  int img0 = 0, y0 = 0;
  if (p.mode) {
    const int HW = p.H * p.W;
    if (HW >= 128) {
      const int tpi = HW / 128;
      img0 = mt / tpi;
      y0 = (mt - img0 * tpi) * p.bh;
    } else {
      img0 = mt * p.bn;
    }
  }

  if (warp == 1 && lane == 0) {
    for (int i = 0; i < L::kStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int nchunk = p.ntap * p.chunks;

  if (warp == 0) {
    // ===================================================== TMA producer
    if (lane == 0) {
      int st = 0;
      uint32_t phase = 0;
      for (int tap = 0; tap < p.ntap; ++tap)
        for (int c = 0; c < p.chunks; ++c) {
          mbar_wait(&empty[st], phase ^ 1);
          unsigned char* s = base + st * L::kStage;
          mbar_arrive_expect_tx(&full[st], L::kStage);
          if (p.mode) {
            tma_load_4d(&mA_hi, &full[st], s, c * 64, p.dx[tap], y0 + p.dy[tap], img0);
            tma_load_4d(&mA_lo, &full[st], s + L::kA, c * 64, p.dx[tap], y0 + p.dy[tap], img0);
          } else {
            tma_load_2d(&mA_hi, &full[st], s, c * 64, mt * 128);
            tma_load_2d(&mA_lo, &full[st], s + L::kA, c * 64, mt * 128);
          }
          const int kb = (tap * p.chunks + c) * 64;
          tma_load_2d(&mB_hi, &full[st], s + 2 * L::kA, kb, n0);
          tma_load_2d(&mB_lo, &full[st], s + 2 * L::kA + L::kB, kb, n0);
          if (++st == L::kStages) { st = 0; phase ^= 1; }
        }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    if (lane == 0) {
      int st = 0;
      uint32_t phase = 0;
      for (int c = 0; c < nchunk; ++c) {
        mbar_wait(&full[st], phase);
        tc_fence_after();
        unsigned char* s = base + st * L::kStage;
        const uint64_t a_hi = make_kmajor_sw128_desc(s), a_lo = make_kmajor_sw128_desc(s + L::kA);
        const uint64_t b_hi = make_kmajor_sw128_desc(s + 2 * L::kA), b_lo = make_kmajor_sw128_desc(s + 2 * L::kA + L::kB);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t adv = (uint64_t)((k * 32) >> 4);
          umma_f16(tmem_base + BN, a_lo + adv, b_hi + adv, kIdesc, (c | k) != 0 ? 1u : 0u);
          umma_f16(tmem_base + BN, a_hi + adv, b_lo + adv, kIdesc, 1u);
          umma_f16(tmem_base, a_hi + adv, b_hi + adv, kIdesc, (c | k) != 0 ? 1u : 0u);
        }
        umma_commit(&empty[st]);
        if (++st == L::kStages) { st = 0; phase ^= 1; }
      }
      umma_commit(acc_full);
    }
  } else if (warp >= 4) {
    // ===================================================== epilogue: warp q owns TMEM lanes 32 q .. 32 q + 31 (tile rows)
    const int q = warp - 4, row = q * 32 + lane;
    bool valid;
    int img = 0, y = 0, x = 0;
    long long m = (long long)mt * 128 + row;
    if (p.mode) {
      const int per_img = p.bh * p.W;
      const int ni = row / per_img, rem = row - ni * per_img;
      img = img0 + ni;
      y = y0 + rem / p.W;
      x = rem - (rem / p.W) * p.W;
      valid = img < p.n_img;
      m = ((long long)img * p.H + y) * p.W + x;
    } else {
      valid = m < p.M;
    }
    const int OH = p.H * p.os, OW = p.W * p.os, oy = y * p.os + p.py, ox = x * p.os + p.px;
    const long long opix = p.mode ? ((long long)img * OH + oy) * OW + ox : m;        // output pixel / row index
    // per-column epilogue constants -> shared memory while the main loop runs (the epilogue warps are idle until then)
    for (int i = tid - 128; i < BN; i += 128) {
      const int n = n0 + i;
      const bool ok = n < p.N;
      par[i] = (ok && p.bias) ? p.bias[n] : 0.f;
      par[BN + i] = (ok && p.post_scale) ? p.post_scale[n] : 1.f;
      par[2 * BN + i] = (ok && p.post_scale) ? p.post_shift[n] : 0.f;
      par[3 * BN + i] = (ok && p.out2_hi) ? p.pre_scale[n] : 0.f;
      par[4 * BN + i] = (ok && p.out2_hi) ? p.pre_shift[n] : 0.f;
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    mbar_wait(acc_full, 0);
    tc_fence_after();
    bool over = false;
#pragma unroll 1
    for (int g = 0; g < BN / 16; ++g) {
      uint32_t v0[16], v1[16];
      tmem_ld16x2(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * 16), tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(BN + g * 16), v0, v1);
      const int nb = n0 + g * 16;
      if (!valid || nb >= p.N) continue;
      float o[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int n = nb + j;
        float val = 0.f;
        if (n < p.N) {
          val = fmaf(__uint_as_float(v1[j]), kLoInv, __uint_as_float(v0[j])) * p.unscale + par[g * 16 + j];
          val = fmaf(val, par[BN + g * 16 + j], par[2 * BN + g * 16 + j]);
          val = val >= 0.f ? val : val * p.slope;
        }
        o[j] = val;
      }
      if (p.res_hi) {
        const __half* rh = p.res_hi + opix * p.out_cp + nb;
        const __half* rl = p.res_lo + opix * p.out_cp + nb;
        uint4 a[2], b[2];
#pragma unroll
        for (int j8 = 0; j8 < 2; ++j8) { a[j8] = *reinterpret_cast<const uint4*>(rh + j8 * 8); b[j8] = *reinterpret_cast<const uint4*>(rl + j8 * 8); }
        const __half* ah = reinterpret_cast<const __half*>(a);
        const __half* bl = reinterpret_cast<const __half*>(b);
#pragma unroll
        for (int j = 0; j < 16; ++j) o[j] += join_half(ah[j], bl[j]);
      }
      if (p.res_f32) {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (nb + j < p.N) o[j] += p.res_f32[m * p.ldc + nb + j];
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) over |= !(fabsf(o[j]) < 60000.f);
      if (p.out_hi) {
        // 16 channels of this pixel: 32 bytes (one sector) per plane; planes are padded to multiples of 64 channels, zero beyond N
        uint4 ph[2], pl[2];
        __half* hh = reinterpret_cast<__half*>(ph);
        __half* ll = reinterpret_cast<__half*>(pl);
#pragma unroll
        for (int j = 0; j < 16; ++j) split_half(o[j], hh[j], ll[j]);
        uint4* dh = reinterpret_cast<uint4*>(p.out_hi + opix * p.out_cp + nb);
        uint4* dl = reinterpret_cast<uint4*>(p.out_lo + opix * p.out_cp + nb);
        dh[0] = ph[0]; dh[1] = ph[1]; dl[0] = pl[0]; dl[1] = pl[1];
      }
      if (p.out2_hi) {
        uint4 ph[2], pl[2];
        __half* hh = reinterpret_cast<__half*>(ph);
        __half* ll = reinterpret_cast<__half*>(pl);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float t = 0.f;
          if (nb + j < p.N) { t = fmaf(o[j], par[3 * BN + g * 16 + j], par[4 * BN + g * 16 + j]); t = t >= 0.f ? t : t * p.pre_slope; }
          split_half(t, hh[j], ll[j]);
        }
        uint4* dh = reinterpret_cast<uint4*>(p.out2_hi + opix * p.out_cp + nb);
        uint4* dl = reinterpret_cast<uint4*>(p.out2_lo + opix * p.out_cp + nb);
        dh[0] = ph[0]; dh[1] = ph[1]; dl[0] = pl[0]; dl[1] = pl[1];
      }
      if (p.out_f32) {
        if (p.mode) {
          // NCHW: [img][n][oy][ox]; consecutive lanes are consecutive pixels of one channel plane
          float* d = p.out_f32 + (long long)img * p.f32_img_stride + (long long)oy * OW + ox;
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (nb + j < p.N) d[(long long)(nb + j) * OH * OW] = o[j];
        } else {
          float* d = p.out_f32 + m * p.ldc + nb;
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (nb + j < p.N) d[j] = o[j];
        }
      }
    }
    if (over && p.overflow_flag) atomicOr(p.overflow_flag, 1);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kCols));
  }
}

// ---------------------------------------------------------------------------------------------------------------------------
// layout kernels
// ---------------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float pt_aligned_heat(const float* __restrict__ hm, int n, float relw, float relh, int i, int j) {
  const float gx = ((float)i / (float)(n - 1) * 2.f - 1.f) * relw;
  const float gy = ((float)j / (float)(n - 1) * 2.f - 1.f) * relh;
  const float ix = ((gx + 1.f) * (float)n - 1.f) * 0.5f, iy = ((gy + 1.f) * (float)n - 1.f) * 0.5f;
  const float fx = floorf(ix), fy = floorf(iy);
  const int x0 = (int)fx, y0 = (int)fy;
  const float tx = ix - fx, ty = iy - fy;
  auto at = [&](int y, int x) { return (y >= 0 && y < n && x >= 0 && x < n) ? hm[y * n + x] : 0.f; };
  return at(y0, x0) * ((1.f - tx) * (1.f - ty)) + at(y0, x0 + 1) * (tx * (1.f - ty)) + at(y0 + 1, x0) * ((1.f - tx) * ty) +
         at(y0 + 1, x0 + 1) * (tx * ty);
}

// NCHW float32 features (+ optionally the re-aligned, flipped, 2x down-sampled heat-maps as extra channels: the encoder input of
// VPHO.py:136-151, same arithmetic as k_encoder_input) -> NHWC hi/lo planes with Cp channels (zero padded).  One CTA per
// (image, row): the row is transposed through shared memory so that both the reads (along x) and the writes (along c) coalesce.
__global__ void __launch_bounds__(256) k_to_planes(const float* __restrict__ feat, const float* __restrict__ hm, const float* __restrict__ bbox,
                                                   const float* __restrict__ bbox_rect, const unsigned char* __restrict__ is_right,
                                                   int flip_feat, int flip_hm, int C, int J, int roi, int Cp, __half* __restrict__ hi,
                                                   __half* __restrict__ lo) {
  extern __shared__ float tile[];                  // [roi][Cp + 1]
  const int img = blockIdx.y, y = blockIdx.x, ld = Cp + 1;
  const bool flip = !is_right[img];
  for (int i = threadIdx.x; i < Cp * roi; i += blockDim.x) {
    const int c = i / roi, x = i - c * roi;
    float v = 0.f;
    if (c < C) {
      const int xs = (flip && flip_feat) ? roi - 1 - x : x;
      v = feat[(((long long)img * C + c) * roi + y) * roi + xs];
    } else if (c < C + J) {
      const int n = 2 * roi;
      const float* h = hm + ((long long)img * J + (c - C)) * n * n;
      const float bw = bbox[img * 4 + 2] - bbox[img * 4 + 0], bh = bbox[img * 4 + 3] - bbox[img * 4 + 1];
      const float rw = (bbox_rect[img * 4 + 2] - bbox_rect[img * 4 + 0]) / bw, rh = (bbox_rect[img * 4 + 3] - bbox_rect[img * 4 + 1]) / bh;
      float s[2][2];
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          const int ii = 2 * y + a, j0 = 2 * x + b;
          const int jj = (flip && flip_hm) ? n - 1 - j0 : j0;
          s[a][b] = pt_aligned_heat(h, n, rw, rh, ii, jj);
        }
      v = 0.5f * (0.5f * s[0][0] + 0.5f * s[0][1]) + 0.5f * (0.5f * s[1][0] + 0.5f * s[1][1]);
    }
    tile[x * ld + c] = v;
  }
  __syncthreads();
  const long long o0 = ((long long)img * roi + y) * roi * Cp;
  for (int i = threadIdx.x; i < Cp * roi / 8; i += blockDim.x) {       // 8 channels = 16 bytes per plane per thread
    const int x = (i * 8) / Cp, c = i * 8 - x * Cp;
    uint4 ph, pl;
    __half* hh = reinterpret_cast<__half*>(&ph);
    __half* ll = reinterpret_cast<__half*>(&pl);
#pragma unroll
    for (int e = 0; e < 8; ++e) split_half(tile[x * ld + c + e], hh[e], ll[e]);
    *reinterpret_cast<uint4*>(hi + o0 + (long long)i * 8) = ph;
    *reinterpret_cast<uint4*>(lo + o0 + (long long)i * 8) = pl;
  }
}

// 2x2 max-pool over NHWC planes (value = hi + lo) -> any of: pooled raw planes, pre-activated planes (the next Residual's
// BatchNorm + LeakyReLU), float32 NCHW (the flattened encoding).
__global__ void __launch_bounds__(256) k_pool_planes(const __half* __restrict__ ihi, const __half* __restrict__ ilo, int n_img, int H, int W, int C,
                                                     __half* __restrict__ ohi, __half* __restrict__ olo, __half* __restrict__ phi,
                                                     __half* __restrict__ plo, const float* __restrict__ pre_scale,
                                                     const float* __restrict__ pre_shift, float pre_slope, float* __restrict__ f32_nchw) {
  const int OH = H / 2, OW = W / 2;
  const long long total = (long long)n_img * OH * OW * C;
  for (long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x; it < total; it += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(it % C);
    const long long pix = it / C;
    const int x = (int)(pix % OW), y = (int)(pix / OW % OH), img = (int)(pix / ((long long)OW * OH));
    float v = -INFINITY;
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        const long long i = (((long long)img * H + 2 * y + a) * W + 2 * x + b) * C + c;
        v = fmaxf(v, join_half(ihi[i], ilo[i]));
      }
    if (ohi) {
      __half h, l;
      split_half(v, h, l);
      ohi[it] = h;
      olo[it] = l;
    }
    if (phi) {
      float t = fmaf(v, pre_scale[c], pre_shift[c]);
      t = t >= 0.f ? t : t * pre_slope;
      __half h, l;
      split_half(t, h, l);
      phi[it] = h;
      plo[it] = l;
    }
    if (f32_nchw) f32_nchw[(((long long)img * C + c) * OH + y) * OW + x] = v;
  }
}

// row-major float32 [rows][K] -> hi/lo planes with the same shape (K a multiple of 64)
__global__ void __launch_bounds__(256) k_split_rows(const float* __restrict__ x, long long n, __half* __restrict__ hi, __half* __restrict__ lo) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    __half h, l;
    split_half(x[i], h, l);
    hi[i] = h;
    lo[i] = l;
  }
}

// ---------------------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 pt_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* q = nullptr;
    cudaDriverEntryPointQueryResult r;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &r) == cudaSuccess && r == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(q);
  }
  return fn;
}

// NHWC __half plane [n_img][H][W][Cp] -> boxes {64 channels, W, bh, bn}
static bool make_map_nhwc(CUtensorMap* map, const void* base, int n_img, int H, int W, int Cp, int bh, int bn) {
  PFN_cuTensorMapEncodeTiled_v12000 enc = pt_encode();
  if (!enc) return false;
  cuuint64_t gdim[4] = {(cuuint64_t)Cp, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)n_img};
  cuuint64_t gstride[3] = {(cuuint64_t)Cp * 2, (cuuint64_t)W * Cp * 2, (cuuint64_t)H * W * Cp * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)W, (cuuint32_t)bh, (cuuint32_t)bn};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static bool make_map_rows(CUtensorMap* map, const void* base, long long rows, int K, int box_rows) {
  PFN_cuTensorMapEncodeTiled_v12000 enc = pt_encode();
  if (!enc) return false;
  cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool pt_available() { return pt_encode() != nullptr; }

// weights [N][Kp] (Kp = ntap * Cp) as hi/lo __half planes scaled by an exact power of two; rows padded to a multiple of BN
bool pt_make_weights(TcWeights& w, const std::vector<float>& dense /* [N][Kp] */, int N, int Kp) {
  w.N = N;
  w.Kp = Kp;
  w.BN = N > 64 ? 128 : (N > 32 ? 64 : 32);
  w.Npad = (N + w.BN - 1) / w.BN * w.BN;
  float mx = 0.f;
  for (float v : dense) mx = fmaxf(mx, fabsf(v));
  int e = 0;
  if (mx > 0.f) {
    frexpf(mx, &e);                     // mx = f * 2^e, f in [0.5, 1)
    e = 14 - e;                         // scaled maximum in [2^13, 2^14)
  }
  const float scale = ldexpf(1.f, e);
  w.unscale = ldexpf(1.f, -e);
  std::vector<__half> planes((size_t)2 * w.Npad * Kp, __float2half(0.f));
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < Kp; ++k) {
      const float v = dense[(size_t)n * Kp + k] * scale;
      const __half h = __float2half_rn(v);
      planes[(size_t)n * Kp + k] = h;
      planes[(size_t)w.Npad * Kp + (size_t)n * Kp + k] = __float2half_rn((v - __half2float(h)) * 2048.f);
    }
  if (cudaMalloc(&w.planes, planes.size() * sizeof(__half)) != cudaSuccess) return false;
  cudaMemcpy(w.planes, planes.data(), planes.size() * sizeof(__half), cudaMemcpyHostToDevice);
  return make_map_rows(reinterpret_cast<CUtensorMap*>(w.map_hi), w.planes, w.Npad, Kp, w.BN) &&
         make_map_rows(reinterpret_cast<CUtensorMap*>(w.map_lo), static_cast<__half*>(w.planes) + (size_t)w.Npad * Kp, w.Npad, Kp, w.BN);
}

template <int BN, int NST>
static int launch_bn(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const TcWeights& w, const TcGemm& p, int m_tiles, cudaStream_t st) {
  using L = PtSmemLayout<BN, NST>;
  static bool attr_done[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr_done[dev]) {
    if (cudaFuncSetAttribute(k_gemm_tc<BN, NST>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kBytes) != cudaSuccess) return VPHO_ERR_LAUNCH;
    attr_done[dev] = true;
  }
  ++g_launches;
  k_gemm_tc<BN, NST><<<dim3(m_tiles, w.Npad / BN), 256, L::kBytes, st>>>(a_hi, a_lo, *reinterpret_cast<const CUtensorMap*>(w.map_hi),
                                                                         *reinterpret_cast<const CUtensorMap*>(w.map_lo), p);
  return cudaGetLastError() == cudaSuccess ? VPHO_OK : VPHO_ERR_LAUNCH;
}

// A operand: NHWC planes (mode 1) or row-major planes (mode 0).  p must be filled except for the geometry derived here.
int pt_gemm(const TcWeights& w, TcGemm p, const __half* a_hi, const __half* a_lo, cudaStream_t st) {
  alignas(64) CUtensorMap ma_hi, ma_lo;
  int m_tiles;
  p.N = w.N;
  p.unscale = w.unscale;
  if (p.mode) {
    const int HW = p.H * p.W;
    if (p.W > 128 || 128 % p.W) return VPHO_ERR_INVALID;
    p.bh = HW >= 128 ? 128 / p.W : p.H;
    p.bn = HW >= 128 ? 1 : 128 / HW;
    if ((HW >= 128 && HW % 128) || (HW < 128 && 128 % HW)) return VPHO_ERR_INVALID;
    m_tiles = HW >= 128 ? p.n_img * (HW / 128) : (p.n_img + p.bn - 1) / p.bn;
    const int Cp = p.chunks * 64;
    if (w.Kp != p.ntap * Cp) return VPHO_ERR_INVALID;
    if (!make_map_nhwc(&ma_hi, a_hi, p.n_img, p.H, p.W, Cp, p.bh, p.bn) || !make_map_nhwc(&ma_lo, a_lo, p.n_img, p.H, p.W, Cp, p.bh, p.bn))
      return VPHO_ERR_LAUNCH;
  } else {
    m_tiles = (int)((p.M + 127) / 128);
    p.ntap = 1;
    if (w.Kp != p.chunks * 64) return VPHO_ERR_INVALID;
    if (!make_map_rows(&ma_hi, a_hi, p.M, w.Kp, 128) || !make_map_rows(&ma_lo, a_lo, p.M, w.Kp, 128)) return VPHO_ERR_LAUNCH;
  }
  if (m_tiles <= 0) return VPHO_OK;
  // 1x1 convolutions / narrow linears on grids of more than two waves: two CTAs per SM (64 KB each) whose phases interleave;
  // everything else (long K, or grids that leave SMs free anyway): one CTA per SM with the deep pipeline
  const bool short_k = p.ntap * p.chunks <= 4 && (long long)m_tiles * (w.Npad / w.BN) > 2 * 148;
  switch (w.BN) {
    case 128: return short_k ? launch_bn<128, 1>(ma_hi, ma_lo, w, p, m_tiles, st) : launch_bn<128, 3>(ma_hi, ma_lo, w, p, m_tiles, st);
    case 64: return short_k ? launch_bn<64, 1>(ma_hi, ma_lo, w, p, m_tiles, st) : launch_bn<64, 4>(ma_hi, ma_lo, w, p, m_tiles, st);
    default: return short_k ? launch_bn<32, 1>(ma_hi, ma_lo, w, p, m_tiles, st) : launch_bn<32, 4>(ma_hi, ma_lo, w, p, m_tiles, st);
  }
}

int pt_to_planes(const float* feat, const float* hm, const float* bbox, const float* bbox_rect, const unsigned char* is_right, int flip_feat,
                 int flip_hm, int bs, int C, int J, int roi, int Cp, __half* hi, __half* lo, cudaStream_t st) {
  const size_t smem = (size_t)roi * (Cp + 1) * sizeof(float);
  static bool attr_done[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && !attr_done[dev]) {
    if (cudaFuncSetAttribute(k_to_planes, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) return VPHO_ERR_LAUNCH;
    attr_done[dev] = true;
  }
  if (smem > 200 * 1024) return VPHO_ERR_INVALID;
  ++g_launches;
  k_to_planes<<<dim3(roi, bs), 256, smem, st>>>(feat, hm, bbox, bbox_rect, is_right, flip_feat, flip_hm, C, J, roi, Cp, hi, lo);
  return cudaGetLastError() == cudaSuccess ? VPHO_OK : VPHO_ERR_LAUNCH;
}

int pt_pool(const __half* ihi, const __half* ilo, int n_img, int H, int W, int C, __half* ohi, __half* olo, __half* phi, __half* plo,
            const float* pre_scale, const float* pre_shift, float pre_slope, float* f32_nchw, cudaStream_t st) {
  const long long total = (long long)n_img * (H / 2) * (W / 2) * C;
  ++g_launches;
  k_pool_planes<<<(unsigned)std::min<long long>((total + 255) / 256, 148 * 16), 256, 0, st>>>(ihi, ilo, n_img, H, W, C, ohi, olo, phi, plo, pre_scale,
                                                                                            pre_shift, pre_slope, f32_nchw);
  return cudaGetLastError() == cudaSuccess ? VPHO_OK : VPHO_ERR_LAUNCH;
}

int pt_split_rows(const float* x, long long n, __half* hi, __half* lo, cudaStream_t st) {
  ++g_launches;
  k_split_rows<<<(unsigned)std::min<long long>((n + 255) / 256, 148 * 16), 256, 0, st>>>(x, n, hi, lo);
  return cudaGetLastError() == cudaSuccess ? VPHO_OK : VPHO_ERR_LAUNCH;
}

}  // namespace vpho
