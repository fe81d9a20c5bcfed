// Batched MANO layer on the 5th-generation tensor cores.  Replaces `HeadMano.get_hand_verts` -> manopth `ManoLayer.forward`
// (lib/model/head_mano.py:78-87) when vertices are materialised.
//
// The blend  v_posed[c][v][d] = template + sum_k coef[c][k] dirs[k][v][d]  (10 shape + 135 pose-corrective coefficients) is the
// one dense contraction of the layer: 4.3 GFLOP at 6400 candidates, 90 % of the FP32 work of the SIMT kernel (mano.cu).
// Here it runs as a tcgen05 GEMM with FP32-class accuracy by 3xFP16 operand splitting (as the score network):
//   M = 128 vertices of one coordinate plane  (A operand: blend directions [3 d][896 v][192 k], __half (hi, lo) planes scaled
//       by one power of two, streamed by TMA, SWIZZLE_128B K-major, 3-stage mbarrier pipeline of 32 KB chunks)
//   N = 64 candidates of this CTA             (B operand: their coefficients, split per candidate with a power-of-two row
//       scale and written once into swizzled shared memory by the set-up phase)
//   K = 145 -> 10 k-steps of 16;  D[128 x 64] FP32 in TMEM, one accumulator per coordinate, double-buffered per vertex tile.
// Per CTA: ONE pose set-up for its 64 candidates (Rodrigues, joint regression, kinematic chain, skinning transforms -- the
// SIMT kernel repeats it for each of its 4 vertex chunks), then for every vertex tile 30 UMMAs per coordinate while 16
// epilogue warps drain the previous tile: thread = vertex (TMEM lane), 16 candidates per warp-column group; skinning uses the
// vertex's NON-ZERO weights only (ascending joint order, i.e. the dense sum with its exact zeros skipped -- bit-identical to
// it), then wrist-centring; the warp stages its 32 vertices x 3 floats through shared memory and writes them with
// contiguous 8-byte stores (a candidate's 9336-byte vertex block is only 8-byte aligned).
// Roofline: HBM-bound on the 9588 B per candidate it writes.
#include "mano_device.cuh"
#include "tc_ptx.cuh"
#include "vpho_b200.h"

#include <cuda_fp16.h>

#include <cmath>
#include <vector>

namespace vpho {

constexpr int kMtNCMax = 64;               // candidates per CTA = UMMA N: 64 or 48 (template parameter NC, chosen per launch)
constexpr int kMtK = 192;                  // 145 coefficients padded to 3 chunks of 64
constexpr int kMtChunks = 3;
constexpr int kMtVT = 7;                   // vertex tiles of 128 (896 padded slots)
constexpr int kMtVPad = kMtVT * 128;
constexpr int kMtThreads = 640;            // 4 role warps + 16 epilogue warps
constexpr int kMtPlane = 128 * 128;        // one operand plane of a chunk: 128 rows x 64 halves
constexpr int kMtStage = 2 * kMtPlane;     // hi + lo
constexpr int kMtStages = 3;
constexpr uint32_t mt_idesc(int nc) { return (1u << 4) | ((uint32_t)(nc >> 3) << 17) | ((uint32_t)(128 >> 4) << 24); }

struct ManoTcDev {
  const float* tmpl;            // [3][896]
  const float* nz_w;            // [896][4]   non-zero skinning weights, ascending joint order
  const unsigned short* nz_j;   // [896]      their joints, 4 bits each
  const unsigned char* nz_n;    // [896]      how many (<= 4), 255 = more than 4: dense row below
  const float* w_dense;         // [896][16]
  float dirs_inv;               // exact power-of-two un-scaling of the direction planes
};

struct ManoTcHost {
  ManoTcDev dev;
  void* blob = nullptr;
  alignas(64) unsigned char map_hi[128];
  alignas(64) unsigned char map_lo[128];
};

template <int NC>
struct MtSmem {
  unsigned char stage[kMtStages][kMtStage];          // 96 KB; the set-up phase uses it as scratch first
  unsigned char coef[kMtChunks][2][NC * 128];        // NC rows x 64 halves per plane: 48 KB at NC = 64
  float A[NC][16][12];                            // 48 KB skinning transforms
  float center[NC][4];                            // wrist translation, [3] = un-scaling of the candidate's coefficient row
  float outbuf[16][96];                              // per epilogue warp: 32 vertices x 3 floats
  unsigned long long full[kMtStages], empty[kMtStages], tmem_full[2], tmem_empty[2];
  uint32_t tmem_base;
};
// set-up scratch inside the stage region
template <int NC>
struct MtScratch {
  float R[NC][16][9];
  float J[NC][16][3];
  float coef[NC][148];
};
static_assert(sizeof(MtScratch<kMtNCMax>) <= kMtStages * kMtStage, "set-up scratch must fit the pipeline region");

template <int NC>
__global__ void __launch_bounds__(kMtThreads, 1)
k_mano_tc(const __grid_constant__ CUtensorMap tmD_hi, const __grid_constant__ CUtensorMap tmD_lo, ManoModelDev m, ManoTcDev t,
          const float* __restrict__ pose, const float* __restrict__ shape, int pose_stride, int shape_stride, int n,
          float* __restrict__ verts, float* __restrict__ joints, int vt_per_cta, int debug_blend) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  MtSmem<NC>& sm = *reinterpret_cast<MtSmem<NC>*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
  MtScratch<NC>& sc = *reinterpret_cast<MtScratch<NC>*>(sm.stage[0]);
  constexpr int CG = NC / 4;                  // candidates per epilogue warp column group: 16 or 12
  constexpr uint32_t kIdesc = mt_idesc(NC);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int c0 = blockIdx.x * NC;
  const int vt_lo = blockIdx.y * vt_per_cta, vt_hi = min(kMtVT, vt_lo + vt_per_cta);

  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kMtStages; ++i) { mbar_init(&sm.full[i], 1); mbar_init(&sm.empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&sm.tmem_full[i], 1); mbar_init(&sm.tmem_empty[i], 16); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  pdl_wait();          // launched with launch_pdl: the set-up above overlaps the predecessor's tail
  pdl_trigger();

  // ------------------------------------------------------------------ set-up, once per candidate
  // (1) Rodrigues, pose-corrective coefficients, regressed joints: one thread per (candidate, kinematic joint)
  for (int it = tid; it < NC * 16; it += kMtThreads) {
    const int c = it >> 4, j = it & 15;
    const bool valid = c0 + c < n;
    float R[9], beta[10];
    if (valid) {
      VPHO_BOUNDS(c0 + c < n && 3 * j + 2 < 48);
      const float* pp = pose + (size_t)(c0 + c) * pose_stride + 3 * j;
      const float a[3] = {pp[0], pp[1], pp[2]};
      manopth_rodrigues(a, R);
      const float* sp = shape + (size_t)(c0 + c) * shape_stride;
#pragma unroll
      for (int k = 0; k < 10; ++k) beta[k] = sp[k];
    } else {
#pragma unroll
      for (int e = 0; e < 9; ++e) R[e] = (e % 4 == 0) ? 1.f : 0.f;
#pragma unroll
      for (int k = 0; k < 10; ++k) beta[k] = 0.f;
    }
#pragma unroll
    for (int e = 0; e < 9; ++e) sc.R[c][j][e] = R[e];
    if (j >= 1) {
#pragma unroll
      for (int e = 0; e < 9; ++e) sc.coef[c][10 + (j - 1) * 9 + e] = R[e] - ((e % 4 == 0) ? 1.f : 0.f);
    } else {
#pragma unroll
      for (int k = 0; k < 10; ++k) sc.coef[c][k] = beta[k];
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < 10; ++k) acc = fmaf(m.J_shapedirs[(j * 3 + d) * 10 + k], beta[k], acc);
      sc.J[c][j][d] = acc + m.J_template[j * 3 + d];
    }
  }
  __syncthreads();
  // (2) kinematic chain and skinning transforms: one thread per (candidate, finger); the root goes with finger 0
  for (int it = tid; it < NC * 5; it += kMtThreads) {
    const int c = it / 5, f = it % 5;
    const bool store_j = blockIdx.y == 0 && c0 + c < n;
    float G0[12];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
      for (int cc = 0; cc < 3; ++cc) G0[r * 4 + cc] = sc.R[c][0][r * 3 + cc];
      G0[r * 4 + 3] = sc.J[c][0][r];
    }
    auto emit = [&](int j, const float* G) {          // A_j = G_j - [0 | G_j J_j];  kinematic joint -> output slot
      const float* Jj = sc.J[c][j];
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const float gj = (G[r * 4 + 0] * Jj[0] + G[r * 4 + 1] * Jj[1]) + G[r * 4 + 2] * Jj[2];
        sm.A[c][j][r * 4 + 0] = G[r * 4 + 0];
        sm.A[c][j][r * 4 + 1] = G[r * 4 + 1];
        sm.A[c][j][r * 4 + 2] = G[r * 4 + 2];
        sm.A[c][j][r * 4 + 3] = G[r * 4 + 3] - gj;
      }
      if (store_j) {
        VPHO_BOUNDS(c0 + c < n && joint16_to_21(j) < 21);
        float* o = joints + ((size_t)(c0 + c) * 21 + joint16_to_21(j)) * 3;
#pragma unroll
        for (int d = 0; d < 3; ++d) o[d] = mano_center_scale(G[d * 4 + 3], G0[d * 4 + 3]);
      }
    };
    if (f == 0) {
      emit(0, G0);
      sm.center[c][0] = G0[3]; sm.center[c][1] = G0[7]; sm.center[c][2] = G0[11];
    }
    float Gp[12];
#pragma unroll
    for (int e = 0; e < 12; ++e) Gp[e] = G0[e];
    int jp = 0;
#pragma unroll
    for (int lv = 0; lv < 3; ++lv) {
      const int j = 1 + 3 * f + lv;
      const float d[3] = {sc.J[c][j][0] - sc.J[c][jp][0], sc.J[c][j][1] - sc.J[c][jp][1], sc.J[c][j][2] - sc.J[c][jp][2]};
      float Gn[12];
      compose34(Gp, sc.R[c][j], d, Gn);
      emit(j, Gn);
#pragma unroll
      for (int e = 0; e < 12; ++e) Gp[e] = Gn[e];
      jp = j;
    }
  }
  // (3) coefficient rows -> (hi, lo) __half planes scaled per candidate by an exact power of two (peak in [2^13, 2^14)),
  //     written into the swizzled B operand: one warp per candidate at a time
  for (int c = warp; c < NC; c += kMtThreads / 32) {
    float mx = 0.f;
    for (int k = lane; k < kBlendK; k += 32) mx = fmaxf(mx, fabsf(sc.coef[c][k]));
#pragma unroll
    for (int sft = 16; sft > 0; sft >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, sft));
    float scl = 1.f, inv = 1.f;
    if (mx > 0.f && mx < 3.0e38f) {
      int ex = 0;
      frexpf(mx, &ex);
      scl = ldexpf(1.f, 14 - ex);
      inv = ldexpf(1.f, ex - 14);
    }
    if (lane == 0) sm.center[c][3] = inv;
    for (int k = lane; k < kMtK; k += 32) {
      const float x = k < kBlendK ? sc.coef[c][k] * scl : 0.f;
      const __half h = __float2half_rn(x);
      const __half l = __float2half_rn(x - __half2float(h));
      const int ch = k >> 6, kl = k & 63;
      const uint32_t off = (uint32_t)((c >> 3) * 1024 + (c & 7) * 128 + ((((kl >> 3) ^ (c & 7)) & 7) << 4) + (kl & 7) * 2);
      *reinterpret_cast<__half*>(sm.coef[ch][0] + off) = h;
      *reinterpret_cast<__half*>(sm.coef[ch][1] + off) = l;
    }
  }
  fence_proxy_async();          // generic-proxy writes of the B operand -> visible to the tensor-core (async) proxy
  tc_fence_before();
  __syncthreads();              // set-up complete: the scratch (pipeline region) may now be overwritten by TMA
  tc_fence_after();
  const uint32_t tmem_base = sm.tmem_base;

  if (warp == 0) {
    // ===================================================== TMA producer
    if (lane == 0) {
      int st = 0;
      uint32_t phase = 0;
      for (int vt = vt_lo; vt < vt_hi; ++vt)
        for (int dd = 0; dd < 3; ++dd)
          for (int ch = 0; ch < kMtChunks; ++ch) {
            const int d = dd;
            mbar_wait(&sm.empty[st], phase ^ 1);
            mbar_arrive_expect_tx(&sm.full[st], kMtStage);
            tma_load_2d(&tmD_hi, &sm.full[st], sm.stage[st], ch * 64, d * kMtVPad + vt * 128);
            tma_load_2d(&tmD_lo, &sm.full[st], sm.stage[st] + kMtPlane, ch * 64, d * kMtVPad + vt * 128);
            if (++st == kMtStages) { st = 0; phase ^= 1; }
          }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    if (lane == 0) {
      int st = 0;
      uint32_t phase = 0, acc_phase = 0;
      int buf = 0;
      for (int vt = vt_lo; vt < vt_hi; ++vt) {
        mbar_wait(&sm.tmem_empty[buf], acc_phase ^ 1);
        tc_fence_after();
        for (int dd = 0; dd < 3; ++dd) {
          const int d = dd;
          const uint32_t dcol = tmem_base + (uint32_t)(buf * 3 * NC + d * NC);
          for (int ch = 0; ch < kMtChunks; ++ch) {
            mbar_wait(&sm.full[st], phase);
            tc_fence_after();
            const uint64_t a_hi = make_kmajor_sw128_desc(sm.stage[st]), a_lo = make_kmajor_sw128_desc(sm.stage[st] + kMtPlane);
            const uint64_t b_hi = make_kmajor_sw128_desc(sm.coef[ch][0]), b_lo = make_kmajor_sw128_desc(sm.coef[ch][1]);
            const int ksteps = ch < 2 ? 4 : 2;                  // 145 coefficients = 9.06 k-steps of 16
            for (int k = 0; k < ksteps; ++k) {
              const uint64_t adv = (uint64_t)((k * 32) >> 4);
              umma_f16(dcol, a_lo + adv, b_hi + adv, kIdesc, (ch | k) != 0 ? 1u : 0u);
              umma_f16(dcol, a_hi + adv, b_lo + adv, kIdesc, 1u);
              umma_f16(dcol, a_hi + adv, b_hi + adv, kIdesc, 1u);
            }
            umma_commit(&sm.empty[st]);
            if (++st == kMtStages) { st = 0; phase ^= 1; }
          }
        }
        umma_commit(&sm.tmem_full[buf]);
        if (++buf == 2) { buf = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================================================== epilogue: warp e owns TMEM lanes 32 (e & 3) .. (vertices) and the
    // 16 candidates e >> 2 of the tile
    const int e = warp - 4, q = e & 3, cg = e >> 2;
    float* ob = sm.outbuf[e];
    uint32_t acc_phase = 0;
    int buf = 0;
    for (int vt = vt_lo; vt < vt_hi; ++vt) {
      const int v0 = vt * 128 + q * 32, v = v0 + lane;
      const bool vvalid = v < kVerts;
      const int nv = min(32, kVerts - v0);               // vertices of this warp that exist (<= 0: none)
      // per-vertex constants: rest-pose template and the 16 skinning weights.  A vertex with at most 4 non-zero weights
      // (real MANO assets) takes the sparse path when the whole warp can; otherwise the dense sum over the 16 joints
      float tp[3] = {0.f, 0.f, 0.f}, w16[16];
      unsigned jpack = 0;
      int nnz = 0, tip = -1;
#pragma unroll
      for (int j = 0; j < 16; ++j) w16[j] = 0.f;
      if (vvalid) {
        tp[0] = t.tmpl[v]; tp[1] = t.tmpl[kMtVPad + v]; tp[2] = t.tmpl[2 * kMtVPad + v];
        nnz = t.nz_n[v];
        jpack = t.nz_j[v];
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          const float4 ww = *reinterpret_cast<const float4*>(t.w_dense + (size_t)v * 16 + j4 * 4);
          w16[j4 * 4 + 0] = ww.x; w16[j4 * 4 + 1] = ww.y; w16[j4 * 4 + 2] = ww.z; w16[j4 * 4 + 3] = ww.w;
        }
#pragma unroll
        for (int tt = 0; tt < 5; ++tt)
          if (v == tip_vertex(tt)) tip = tt;
      }
      const bool sparse = __all_sync(0xffffffffu, nnz <= 4);
      float ws4[4] = {0.f, 0.f, 0.f, 0.f};
      if (sparse) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int j = (int)((jpack >> (4 * i)) & 15u);
          float wj = 0.f;
#pragma unroll
          for (int jj = 0; jj < 16; ++jj) wj = jj == j ? w16[jj] : wj;
          ws4[i] = i < nnz ? wj : 0.f;
        }
      }
      mbar_wait(&sm.tmem_full[buf], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 3 * NC + cg * CG);
      const uint32_t dstep = (uint32_t)NC;
#pragma unroll 1
      for (int h8 = 0; h8 < 2; ++h8) {
        uint32_t ax[8], ay[8], az[8];
        tmem_ld8x3(taddr + h8 * 8, taddr + dstep + h8 * 8, taddr + 2 * dstep + h8 * 8, ax, ay, az);
        if (h8 == 1) {
          // the accumulators are in registers: the buffer can be refilled while this warp skins
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&sm.tmem_empty[buf]);
        }
        if (nv <= 0) continue;
        // fully unrolled: the accumulator arrays must be indexed statically (they live in registers filled by the asm above)
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) {
          const int c = cg * CG + h8 * 8 + cc;
          if (h8 * 8 + cc >= CG || c0 + c >= n) break;     // uniform across the warp
          const float un = sm.center[c][3] * t.dirs_inv;
          const float vp[3] = {fmaf(__uint_as_float(ax[cc]), un, tp[0]), fmaf(__uint_as_float(ay[cc]), un, tp[1]),
                               fmaf(__uint_as_float(az[cc]), un, tp[2])};
          // T = sum_j w_j A_j as six packed FP32 pairs (FFMA2; each half is the scalar fmaf of mano_skin_point)
          unsigned long long T2[6];
#pragma unroll
          for (int i = 0; i < 6; ++i) T2[i] = 0ull;
          auto add_joint = [&](int j, float wj) {
            const ulonglong2* a2 = reinterpret_cast<const ulonglong2*>(sm.A[c][j]);
            const ulonglong2 a0 = a2[0], a1 = a2[1], a2v = a2[2];
            const unsigned long long w2 = f32x2_pack(wj, wj);
            T2[0] = f32x2_fma(w2, a0.x, T2[0]); T2[1] = f32x2_fma(w2, a0.y, T2[1]);
            T2[2] = f32x2_fma(w2, a1.x, T2[2]); T2[3] = f32x2_fma(w2, a1.y, T2[3]);
            T2[4] = f32x2_fma(w2, a2v.x, T2[4]); T2[5] = f32x2_fma(w2, a2v.y, T2[5]);
          };
          if (sparse) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
              if (i < nnz) add_joint((int)((jpack >> (4 * i)) & 15u), ws4[i]);
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) add_joint(j, w16[j]);
          }
          float T[12];
#pragma unroll
          for (int i = 0; i < 6; ++i) f32x2_unpack(T2[i], T[2 * i], T[2 * i + 1]);
          float o[3];
#pragma unroll
          for (int r = 0; r < 3; ++r) {
            const float raw = ((T[r * 4 + 0] * vp[0] + T[r * 4 + 1] * vp[1]) + T[r * 4 + 2] * vp[2]) + T[r * 4 + 3];
            o[r] = mano_center_scale(raw, sm.center[c][r]);
          }
          if (debug_blend) { o[0] = vp[0]; o[1] = vp[1]; o[2] = vp[2]; }      // VPHO_MANO_DEBUG_BLEND: the blended rest pose
          ob[lane * 3 + 0] = o[0]; ob[lane * 3 + 1] = o[1]; ob[lane * 3 + 2] = o[2];
          if (tip >= 0) {
            VPHO_BOUNDS(c0 + c < n && tip_to_21(tip) < 21);
            float* dst = joints + ((size_t)(c0 + c) * 21 + tip_to_21(tip)) * 3;
            dst[0] = o[0]; dst[1] = o[1]; dst[2] = o[2];
          }
          __syncwarp();
          // 32 vertices x 3 floats of one candidate are contiguous in global memory: 8-byte stores
          float* gdst = verts + ((size_t)(c0 + c) * kVerts + v0) * 3;
          const int nf = nv * 3;
          VPHO_BOUNDS(c0 + c < n && v0 >= 0 && v0 + nv <= kVerts && nf <= 96);
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) {
            const int i2 = lane + 32 * h2;
            if (i2 < 48) {
              if (2 * i2 + 1 < nf) *reinterpret_cast<float2*>(gdst + 2 * i2) = *reinterpret_cast<const float2*>(ob + 2 * i2);
              else if (2 * i2 < nf) gdst[2 * i2] = ob[2 * i2];
            }
          }
          __syncwarp();
        }
      }
      if (++buf == 2) { buf = 0; acc_phase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
  }
}

// ------------------------------------------------------------------------------------------------------ host side
// tables from the reference-layout model tensors (HOST pointers): posedirs / shapedirs -> scaled __half planes, template,
// compact non-zero skinning weights
int mano_tc_create(const float* v_template, const float* shapedirs, const float* posedirs, const float* weights, void** out) {
  *out = nullptr;
  if (!tc_available()) return VPHO_ERR_LAUNCH;
  const size_t n_plane = (size_t)3 * kMtVPad * kMtK;
  float mx = 0.f;
  for (size_t i = 0; i < (size_t)kVerts * 3 * 10; ++i) mx = fmaxf(mx, fabsf(shapedirs[i]));
  for (size_t i = 0; i < (size_t)kVerts * 3 * 135; ++i) mx = fmaxf(mx, fabsf(posedirs[i]));
  float scl = 1.f, inv = 1.f;
  if (mx > 0.f && mx < 3.0e38f) {
    int ex = 0;
    frexpf(mx, &ex);
    scl = ldexpf(1.f, 14 - ex);
    inv = ldexpf(1.f, ex - 14);
  }
  std::vector<unsigned short> planes(2 * n_plane, 0);
  for (int v = 0; v < kVerts; ++v)
    for (int d = 0; d < 3; ++d)
      for (int k = 0; k < kBlendK; ++k) {
        const float x = (k < 10 ? shapedirs[(v * 3 + d) * 10 + k] : posedirs[(v * 3 + d) * 135 + (k - 10)]) * scl;
        const __half h = __float2half_rn(x);
        const size_t idx = ((size_t)d * kMtVPad + v) * kMtK + k;
        planes[idx] = __half_as_ushort(h);
        planes[n_plane + idx] = __half_as_ushort(__float2half_rn(x - __half2float(h)));
      }
  std::vector<float> tmpl(3 * kMtVPad, 0.f), nzw((size_t)kMtVPad * 4, 0.f), wd((size_t)kMtVPad * 16, 0.f);
  std::vector<unsigned short> nzj(kMtVPad, 0);
  std::vector<unsigned char> nzn(kMtVPad, 0);
  for (int v = 0; v < kVerts; ++v) {
    for (int d = 0; d < 3; ++d) tmpl[d * kMtVPad + v] = v_template[v * 3 + d];
    int cnt = 0;
    for (int j = 0; j < 16; ++j) {
      const float w = weights[v * 16 + j];
      wd[(size_t)v * 16 + j] = w;
      if (w != 0.f) {
        if (cnt < 4) { nzw[(size_t)v * 4 + cnt] = w; nzj[v] = (unsigned short)(nzj[v] | (j << (4 * cnt))); }
        ++cnt;
      }
    }
    nzn[v] = cnt <= 4 ? (unsigned char)cnt : (unsigned char)255;
  }
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = (off + bytes + 255) / 256 * 256; return o; };
  const size_t o_pl = take(planes.size() * 2), o_t = take(tmpl.size() * 4), o_w = take(nzw.size() * 4), o_j = take(nzj.size() * 2),
               o_n = take(nzn.size()), o_d = take(wd.size() * 4);
  ManoTcHost* th = new ManoTcHost();
  if (cudaMalloc(&th->blob, off) != cudaSuccess) { delete th; return VPHO_ERR_ALLOC; }
  char* b = static_cast<char*>(th->blob);
  if (cudaMemcpy(b + o_pl, planes.data(), planes.size() * 2, cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(b + o_t, tmpl.data(), tmpl.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(b + o_w, nzw.data(), nzw.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(b + o_j, nzj.data(), nzj.size() * 2, cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(b + o_n, nzn.data(), nzn.size(), cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(b + o_d, wd.data(), wd.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) {
    cudaFree(th->blob); delete th; return VPHO_ERR_ALLOC;
  }
  if (!tc_make_map(th->map_hi, b + o_pl, 3 * kMtVPad, 128, kMtK, true) ||
      !tc_make_map(th->map_lo, b + o_pl + n_plane * 2, 3 * kMtVPad, 128, kMtK, true)) {
    cudaFree(th->blob); delete th; return VPHO_ERR_LAUNCH;
  }
  th->dev.tmpl = (const float*)(b + o_t); th->dev.nz_w = (const float*)(b + o_w); th->dev.nz_j = (const unsigned short*)(b + o_j);
  th->dev.nz_n = (const unsigned char*)(b + o_n); th->dev.w_dense = (const float*)(b + o_d); th->dev.dirs_inv = inv;
  *out = th;
  return VPHO_OK;
}

void mano_tc_destroy(void* h) {
  if (!h) return;
  ManoTcHost* th = static_cast<ManoTcHost*>(h);
  cudaFree(th->blob);
  delete th;
}

int mano_tc_forward(const void* h, const ManoModelDev& m, const float* pose, const float* shape, int pose_stride, int shape_stride,
                    int n, float* verts, float* joints, cudaStream_t stream, int debug_blend) {
  const ManoTcHost* th = static_cast<const ManoTcHost*>(h);
  const int smem64 = (int)sizeof(MtSmem<64>) + 1024, smem48 = (int)sizeof(MtSmem<48>) + 1024;
  int dev = 0, n_sm = 148;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return VPHO_ERR_LAUNCH;
  static bool attr[64] = {};
  static int sms[64] = {};
  if (!attr[dev]) {
    if (cudaFuncSetAttribute(k_mano_tc<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem64) != cudaSuccess ||
        cudaFuncSetAttribute(k_mano_tc<48>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem48) != cudaSuccess)
      return VPHO_ERR_LAUNCH;
    cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev);
    attr[dev] = true;
  }
  if (sms[dev] > 0) n_sm = sms[dev];
  // Grid shape: candidate tiles of 64 or 48 (UMMA N) x groups of vertex tiles.  A CTA costs about (tiles + 1 set-up) x NC; the
  // launch takes ceil(CTAs / SMs) waves of that.  Few candidates: the vertex tiles are split over more CTAs (each repeats the
  // cheap set-up); 6400 candidates: 134 CTAs of 48 instead of 100 of 64 on 148 SMs.
  int best_nc = 64, best_vt = kMtVT;
  long best_cost = -1;
  for (int nc : {64, 48})
    for (int vt = kMtVT; vt >= 1; --vt) {
      const int nb_ = (n + nc - 1) / nc, gy_ = (kMtVT + vt - 1) / vt;
      const long waves = ((long)nb_ * gy_ + n_sm - 1) / n_sm;
      const long cost = waves * nc * (vt + 1);
      if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_nc = nc; best_vt = vt; }
    }
  const int nb = (n + best_nc - 1) / best_nc, vt_per_cta = best_vt, gy = (kMtVT + vt_per_cta - 1) / vt_per_cta;
  profile_begin(VPHO_TAG_MANO_FULL, stream);
  const CUtensorMap& mhi = *reinterpret_cast<const CUtensorMap*>(th->map_hi);
  const CUtensorMap& mlo = *reinterpret_cast<const CUtensorMap*>(th->map_lo);
  const cudaError_t rc =
      best_nc == 64 ? launch_pdl(k_mano_tc<64>, dim3(nb, gy), dim3(kMtThreads), smem64, stream, 1, mhi, mlo, m, th->dev, pose, shape,
                                 pose_stride, shape_stride, n, verts, joints, vt_per_cta, debug_blend)
                    : launch_pdl(k_mano_tc<48>, dim3(nb, gy), dim3(kMtThreads), smem48, stream, 1, mhi, mlo, m, th->dev, pose, shape,
                                 pose_stride, shape_stride, n, verts, joints, vt_per_cta, debug_blend);
  if (rc != cudaSuccess) return VPHO_ERR_LAUNCH;
  profile_end(VPHO_TAG_MANO_FULL, stream);
  return VPHO_OK;
}

}  // namespace vpho
