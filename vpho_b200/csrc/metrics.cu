// Object-pose evaluation metrics on the device: the step right after the hot path in Trainer.evaluate
// (lib/engine/train_diff_hand_obj.py:236-259), `TesterObject` of lib/engine/test.py:196-584, per (image, candidate):
//   MCE / OCE        criterion_MCE_OCE  test.py:354-375   box corners, float64
//   SMCE             criterion_SMCE     test.py:377-399   min over the object's symmetry transforms, float64
//   MCE2             criterion_MCE2     test.py:401-417 -> compute_obj_metrics_dexycb :155-193, axis-aligned boxes, float32
//   ADD / ADD-S / REP criterion_ADD_REP test.py:419-451   float32 distances of the float64-posed clouds, float64 projection
//   F-score x6 / CD  criterion_FSCORE   test.py:453-503
//   ADD01d / ADDS01d / REP5             test.py:505-521
// With these on the device the reference's evaluate loop needs nothing but the per-image metric rows from the GPU (today:
// numpy on the host plus a `.cuda()` round trip per image).  The O(P^2) nearest-point scans are split over 8 CTAs per
// (candidate, image), tiles of the other cloud staged in shared memory; a second kernel adds the partial sums in a fixed
// order and computes the O(P) metrics.
#include "agg_device.cuh"
#include "vpho_b200.h"

#include <cstring>
#include <vector>

namespace vpho {

constexpr int kMetricCols = VPHO_OBJ_METRIC_COLS;
constexpr int kMT = 256;           // threads per CTA
constexpr int kSplit = 8;          // CTAs sharing the nearest-point scans of one (candidate, image): each owns 1/8 of the points
constexpr int kPartCols = 16;      // partial sums per split: ADD-S, pd->gt, gt->pd distance sums, 6 + 6 threshold hit counts

struct ObjMetricDev {
  int n_obj, sym_k, n_fpts;
  const float* bbox3d;     // [n_obj][8][3]
  const float* diameter;   // [n_obj]
  const double* sym_R;     // [n_obj][K][9]
  const double* sym_t;     // [n_obj][K][3] metres
  const int* sym_count;    // [n_obj] transforms that are not identity padding (padding is harmless: it repeats the identity)
  const float* fverts;     // [n_obj][n_fpts][3] cloud of the F-score / Chamfer terms (nullptr: the assets' sampled surface)
};

struct ObjMetricHost {
  ObjMetricDev dev;
  void* blob = nullptr;
  double* part = nullptr;      // [rows][kSplit][kPartCols] partial sums of the scans, grown on demand
  unsigned* colpart = nullptr; // [rows][kSplit][Q] partial column minima of the one-pass scan
  size_t part_rows = 0;
  int sym = -1;                // one-pass scan usable (Q unsigneds of dynamic shared memory fit): decided at first use
};

__device__ __forceinline__ double block_sum_d(double v, double* red) {
  const int tid = threadIdx.x;
  red[tid] = v;
  __syncthreads();
  for (int st = kMT / 2; st > 0; st >>= 1) {
    if (tid < st) red[tid] += red[tid + st];
    __syncthreads();
  }
  const double r = red[0];
  __syncthreads();
  return r;
}
__device__ __forceinline__ float block_min_f(float v, float* red, bool want_max) {
  const int tid = threadIdx.x;
  red[tid] = v;
  __syncthreads();
  for (int st = kMT / 2; st > 0; st >>= 1) {
    if (tid < st) red[tid] = want_max ? fmaxf(red[tid], red[tid + st]) : fminf(red[tid], red[tid + st]);
    __syncthreads();
  }
  const float r = red[0];
  __syncthreads();
  return r;
}

// p R^T + t in float64 (numpy einsum "ni,...ij->...nj" with the transposed rotation, then + t)
__device__ __forceinline__ void pose_point_d(const double* rt, const float* p, double* out) {
#pragma unroll
  for (int j = 0; j < 3; ++j)
    out[j] = (((double)p[0] * rt[j * 4 + 0] + (double)p[1] * rt[j * 4 + 1]) + (double)p[2] * rt[j * 4 + 2]) + rt[j * 4 + 3];
}

// nearest-point scan: every thread owns points of cloud A (posed by rtA) in [own_lo, own_hi) and scans all of cloud B (posed
// by rtB) through shared-memory tiles; `consume(i, d)` receives the distance of point i.
template <typename F>
__device__ __forceinline__ void nearest_scan(const float* base, int n_pts, int own_lo, int own_hi, const double* rtA, const double* rtB,
                                             float4* tile, F consume) {
  const int tid = threadIdx.x;
  for (int p0 = own_lo; p0 < own_hi; p0 += kMT) {
    const int i = p0 + tid;
    float a[3] = {0.f, 0.f, 0.f};
    if (i < own_hi) {
      double ad[3];
      pose_point_d(rtA, base + (size_t)i * 3, ad);
      a[0] = (float)ad[0]; a[1] = (float)ad[1]; a[2] = (float)ad[2];
    }
    float best = INFINITY;
    for (int q0 = 0; q0 < n_pts; q0 += kMT) {
      __syncthreads();
      if (q0 + tid < n_pts) {
        double g[3];
        pose_point_d(rtB, base + (size_t)(q0 + tid) * 3, g);
        tile[tid] = make_float4((float)g[0], (float)g[1], (float)g[2], 0.f);
      }
      __syncthreads();
      const int cnt = min(kMT, n_pts - q0);
#pragma unroll 8
      for (int q = 0; q < cnt; ++q) {
        const float4 t = tile[q];
        const float dx = a[0] - t.x, dy = a[1] - t.y, dz = a[2] - t.z;
        best = fminf(best, (dx * dx + dy * dy) + dz * dz);
      }
    }
    if (i < own_hi) consume(i, sqrtf(best));
  }
  __syncthreads();
}

// Both directions of a nearest-point query in ONE pass over the pair distances: thread = point i of cloud A (rows
// [own_lo, own_hi)), loop over all points j of cloud B.  Row minima stay in a register; the column minimum of j over the 32
// rows of a warp is one `redux.sync.min` on the distance bits (non-negative floats order like unsigned integers) and lands
// in `colmin[j]` (shared, unsigned) with one shared-memory atomic per warp.  Two j per instruction with packed FP32
// (FADD2 / FMUL2 / FFMA2), as the contact scans of aggregate.cu.  colmin must hold +inf bits on entry.
#ifndef VPHO_EMU
__device__ __forceinline__ unsigned long long m_pack(float a, float b) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
template <typename F>
__device__ __forceinline__ void sym_scan(const float* base, int n_pts, int own_lo, int own_hi, const double* rtA, const double* rtB,
                                         float4* tile, unsigned* colmin, F consume_row) {
  const int tid = threadIdx.x, lane = tid & 31;
  float* tx = reinterpret_cast<float*>(tile);
  float* ty = tx + kMT;
  float* tz = ty + kMT;
  for (int p0 = own_lo; p0 < own_hi; p0 += kMT) {
    const int i = p0 + tid;
    float a[3] = {1e18f, 1e18f, 1e18f};          // a row that does not exist: its distances can never be a minimum
    if (i < own_hi) {
      double ad[3];
      pose_point_d(rtA, base + (size_t)i * 3, ad);
      a[0] = (float)ad[0]; a[1] = (float)ad[1]; a[2] = (float)ad[2];
    }
    const unsigned long long A = m_pack(a[0], a[0]), B = m_pack(a[1], a[1]), C = m_pack(a[2], a[2]);
    float best = INFINITY;
    for (int q0 = 0; q0 < n_pts; q0 += kMT) {
      __syncthreads();
      {
        float g[3] = {-1e18f, -1e18f, -1e18f};   // padding column
        if (q0 + tid < n_pts) {
          double gd[3];
          pose_point_d(rtB, base + (size_t)(q0 + tid) * 3, gd);
          g[0] = (float)gd[0]; g[1] = (float)gd[1]; g[2] = (float)gd[2];
        }
        tx[tid] = g[0]; ty[tid] = g[1]; tz[tid] = g[2];
      }
      __syncthreads();
      const int cnt = min(kMT, n_pts - q0);
#pragma unroll 4
      for (int q = 0; q < cnt; q += 2) {
        const unsigned long long X = *reinterpret_cast<const unsigned long long*>(tx + q);
        const unsigned long long Y = *reinterpret_cast<const unsigned long long*>(ty + q);
        const unsigned long long Z = *reinterpret_cast<const unsigned long long*>(tz + q);
        unsigned long long dx, dy, dz, d;
        asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(dx) : "l"(A), "l"(X));
        asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(dy) : "l"(B), "l"(Y));
        asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(dz) : "l"(C), "l"(Z));
        asm("mul.rn.f32x2 %0, %1, %1;" : "=l"(d) : "l"(dy));
        asm("fma.rn.f32x2 %0, %1, %1, %2;" : "=l"(d) : "l"(dx), "l"(d));
        asm("fma.rn.f32x2 %0, %1, %1, %2;" : "=l"(d) : "l"(dz), "l"(d));
        float d0, d1;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(d));
        best = fminf(best, fminf(d0, d1));
        const unsigned m0 = __reduce_min_sync(0xffffffffu, __float_as_uint(d0));
        const unsigned m1 = __reduce_min_sync(0xffffffffu, __float_as_uint(d1));
        if (lane < 2 && q + lane < cnt) atomicMin(&colmin[q0 + q + lane], lane ? m1 : m0);
      }
    }
    if (i < own_hi) consume_row(i, sqrtf(best));
  }
  __syncthreads();
}
#endif

// The O(P^2) part: ADD-S and the two directions of the F-score / Chamfer cloud, each CTA owning 1 / kSplit of the points.
// Partial sums go to part[(b C + c) kSplit + split][kPartCols]; k_object_metrics adds them in split order (deterministic).
// `colpart` [(b C + c) kSplit + split][Q] receives this CTA's partial column minima (squared distances) of the gt -> pd
// direction when the one-pass scan is used (`sym` != 0: Q unsigneds of dynamic shared memory); k_object_metrics takes the
// minimum over the splits.  Otherwise (clouds too large for shared memory, or the CPU emulator build) the gt -> pd
// direction is a second scan and its sums go to part[2], part[9..14] as before.
__global__ void __launch_bounds__(kMT) k_object_metrics_scan(AssetsDev as, ObjMetricDev mt, const double* __restrict__ pd_rt,
                                                             const double* __restrict__ gt_rt, const int* __restrict__ obj_id, int C,
                                                             double* __restrict__ part, unsigned* __restrict__ colpart, int sym) {
#ifndef VPHO_EMU
  extern __shared__ unsigned s_colmin[];
#endif
  __shared__ double red[kMT];
  __shared__ float4 tile[kMT];
  __shared__ double s_pd[12], s_gt[12];
  const int split = blockIdx.x, c = blockIdx.y, b = blockIdx.z, tid = threadIdx.x;
  const int o = obj_index(as, obj_id[b]);
  if (tid < 12) {
    s_pd[tid] = pd_rt[((size_t)b * C + c) * 12 + tid];
    s_gt[tid] = gt_rt[(size_t)b * 12 + tid];
  }
  __syncthreads();
  double* out = part + (((size_t)b * C + c) * kSplit + split) * kPartCols;
  const float* base = as.verts + (size_t)o * as.n_pts * 3;
  const int P = as.n_pts;
  auto range = [&](int n, int& lo, int& hi) {
    const int per = (n + kSplit - 1) / kSplit;
    lo = min(n, split * per);
    hi = min(n, lo + per);
  };
  int lo, hi;
  const float* fbase = mt.fverts ? mt.fverts + (size_t)o * mt.n_fpts * 3 : base;
  const int Q = mt.fverts ? mt.n_fpts : P;
  const bool same_cloud = !mt.fverts;            // ADD-S is then the pd -> gt direction of the F-score scan: scanned once
  double adds = 0.0;
  if (!same_cloud) {
    range(P, lo, hi);
    nearest_scan(base, P, lo, hi, s_pd, s_gt, tile, [&](int, float d) { adds += (double)d; });
  }
  range(Q, lo, hi);
  const float th[6] = {0.002f, 0.005f, 0.010f, 0.020f, 0.050f, 0.100f};
  double s_pg = 0.0, s_gp = 0.0;
  int hit_pg[6] = {0, 0, 0, 0, 0, 0}, hit_gp[6] = {0, 0, 0, 0, 0, 0};
  auto row_pg = [&](int, float d) {
    s_pg += (double)d;
#pragma unroll
    for (int k = 0; k < 6; ++k) hit_pg[k] += d < th[k] ? 1 : 0;
  };
#ifndef VPHO_EMU
  if (sym) {
    for (int j = tid; j < Q; j += kMT) s_colmin[j] = 0x7f800000u;
    __syncthreads();
    sym_scan(fbase, Q, lo, hi, s_pd, s_gt, tile, s_colmin, row_pg);
    unsigned* cp = colpart + (((size_t)b * C + c) * kSplit + split) * (size_t)Q;
    for (int j = tid; j < Q; j += kMT) cp[j] = s_colmin[j];
  } else
#endif
  {
    nearest_scan(fbase, Q, lo, hi, s_pd, s_gt, tile, row_pg);
    nearest_scan(fbase, Q, lo, hi, s_gt, s_pd, tile, [&](int, float d) {
      s_gp += (double)d;
#pragma unroll
      for (int k = 0; k < 6; ++k) hit_gp[k] += d < th[k] ? 1 : 0;
    });
  }
  const double t_pg = block_sum_d(s_pg, red), t_gp = block_sum_d(s_gp, red);
  const double adds_s = same_cloud ? t_pg : block_sum_d(adds, red);
  if (tid == 0) { out[0] = adds_s; out[1] = t_pg; out[2] = t_gp; }
  for (int k = 0; k < 6; ++k) {
    const double np_ = block_sum_d((double)hit_pg[k], red), ng_ = block_sum_d((double)hit_gp[k], red);
    if (tid == 0) { out[3 + k] = np_; out[9 + k] = ng_; }
  }
}

__global__ void __launch_bounds__(kMT) k_object_metrics(AssetsDev as, ObjMetricDev mt, const double* __restrict__ pd_rt,
                                                        const double* __restrict__ gt_rt, const int* __restrict__ obj_id,
                                                        const float* __restrict__ cam_intr, int C, const double* __restrict__ part,
                                                        const unsigned* __restrict__ colpart, int sym, double* __restrict__ out) {
  __shared__ double red[kMT];
  __shared__ float redf[kMT];
  __shared__ double s_pd[12], s_gt[12], s_pdbox[8][3], s_gtbox[8][3];
  const int c = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const int o = obj_index(as, obj_id[b]);
  if (tid < 12) {
    s_pd[tid] = pd_rt[((size_t)b * C + c) * 12 + tid];
    s_gt[tid] = gt_rt[(size_t)b * 12 + tid];
  }
  __syncthreads();
  double* row = out + ((size_t)b * C + c) * kMetricCols;
  // ---- box corners: MCE, OCE
  const float* box = mt.bbox3d + (size_t)o * 24;
  double dcorner = 0.0;
  if (tid < 8) {
    pose_point_d(s_pd, box + tid * 3, s_pdbox[tid]);
    pose_point_d(s_gt, box + tid * 3, s_gtbox[tid]);
    const double dx = s_pdbox[tid][0] - s_gtbox[tid][0], dy = s_pdbox[tid][1] - s_gtbox[tid][1], dz = s_pdbox[tid][2] - s_gtbox[tid][2];
    dcorner = sqrt((dx * dx + dy * dy) + dz * dz);
  }
  const double mce = block_sum_d(dcorner, red) / 8.0;
  if (tid == 0) {
    double pc[3] = {0, 0, 0}, gc[3] = {0, 0, 0};
    for (int k = 0; k < 8; ++k)
      for (int d = 0; d < 3; ++d) { pc[d] += s_pdbox[k][d]; gc[d] += s_gtbox[k][d]; }
    const double dx = pc[0] / 8.0 - gc[0] / 8.0, dy = pc[1] / 8.0 - gc[1] / 8.0, dz = pc[2] / 8.0 - gc[2] / 8.0;
    row[0] = mce;
    row[1] = sqrt((dx * dx + dy * dy) + dz * dz);
  }
  // ---- SMCE: symmetric copies of the box posed by the ground truth; thread k takes transforms k, k+256, ...
  {
    double best = INFINITY;
    for (int k = tid; k < mt.sym_k; k += kMT) {
      const double* sR = mt.sym_R + ((size_t)o * mt.sym_k + k) * 9;
      const double* st = mt.sym_t + ((size_t)o * mt.sym_k + k) * 3;
      double acc = 0.0;
      for (int n = 0; n < 8; ++n) {
        const float* p = box + n * 3;
        double sp[3], g[3];
#pragma unroll
        for (int j = 0; j < 3; ++j)
          sp[j] = (((double)p[0] * sR[j * 3 + 0] + (double)p[1] * sR[j * 3 + 1]) + (double)p[2] * sR[j * 3 + 2]) + st[j];
#pragma unroll
        for (int j = 0; j < 3; ++j) g[j] = ((sp[0] * s_gt[j * 4 + 0] + sp[1] * s_gt[j * 4 + 1]) + sp[2] * s_gt[j * 4 + 2]) + s_gt[j * 4 + 3];
        const double dx = s_pdbox[n][0] - g[0], dy = s_pdbox[n][1] - g[1], dz = s_pdbox[n][2] - g[2];
        acc += sqrt((dx * dx + dy * dy) + dz * dz);
      }
      best = fmin(best, acc / 8.0);
    }
    red[tid] = best;
    __syncthreads();
    for (int st = kMT / 2; st > 0; st >>= 1) {
      if (tid < st) red[tid] = fmin(red[tid], red[tid + st]);
      __syncthreads();
    }
    if (tid == 0) row[3] = red[0];
    __syncthreads();
  }
  // ---- sampled surface: ADD, REP, axis-aligned boxes (MCE2), then ADD-S
  const float* base = as.verts + (size_t)o * as.n_pts * 3;
  const int P = as.n_pts;
  double add = 0.0, rep = 0.0;
  float mn_p[3] = {INFINITY, INFINITY, INFINITY}, mx_p[3] = {-INFINITY, -INFINITY, -INFINITY};
  float mn_g[3] = {INFINITY, INFINITY, INFINITY}, mx_g[3] = {-INFINITY, -INFINITY, -INFINITY};
  double Kd[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) Kd[k] = (double)cam_intr[(size_t)b * 9 + k];
  for (int i = tid; i < P; i += kMT) {
    double pdv[3], gtv[3];
    pose_point_d(s_pd, base + (size_t)i * 3, pdv);
    pose_point_d(s_gt, base + (size_t)i * 3, gtv);
    const float pf[3] = {(float)pdv[0], (float)pdv[1], (float)pdv[2]}, gf[3] = {(float)gtv[0], (float)gtv[1], (float)gtv[2]};
    const float dx = pf[0] - gf[0], dy = pf[1] - gf[1], dz = pf[2] - gf[2];
    add += (double)sqrtf((dx * dx + dy * dy) + dz * dz);
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      mn_p[d] = fminf(mn_p[d], pf[d]); mx_p[d] = fmaxf(mx_p[d], pf[d]);
      mn_g[d] = fminf(mn_g[d], gf[d]); mx_g[d] = fmaxf(mx_g[d], gf[d]);
    }
    // pixel coordinates  K v / (v_z + 1e-7)  in float64
    double pu[2], gu[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      pu[j] = ((pdv[0] * Kd[j * 3 + 0] + pdv[1] * Kd[j * 3 + 1]) + pdv[2] * Kd[j * 3 + 2]) / (pdv[2] + 1e-7);
      gu[j] = ((gtv[0] * Kd[j * 3 + 0] + gtv[1] * Kd[j * 3 + 1]) + gtv[2] * Kd[j * 3 + 2]) / (gtv[2] + 1e-7);
    }
    const double ux = pu[0] - gu[0], uy = pu[1] - gu[1];
    rep += sqrt(ux * ux + uy * uy);
  }
  const double add_m = block_sum_d(add, red) / (double)P;
  const double rep_px = block_sum_d(rep, red) / (double)P;
  float bp[2][3], bg[2][3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    bp[0][d] = block_min_f(mn_p[d], redf, false); bp[1][d] = block_min_f(mx_p[d], redf, true);
    bg[0][d] = block_min_f(mn_g[d], redf, false); bg[1][d] = block_min_f(mx_g[d], redf, true);
  }
  // the scans' partial sums, in split order
  const double* pp = part + ((size_t)b * C + c) * kSplit * kPartCols;
  double tot[kPartCols];
  for (int k = 0; k < kPartCols; ++k) {
    tot[k] = 0.0;
    for (int sp = 0; sp < kSplit; ++sp) tot[k] += pp[sp * kPartCols + k];
  }
  const double adds_m = tot[0] / (double)P;
  if (sym) {
    // gt -> pd direction of the one-pass scan: minimum over the splits' partial column minima, then the same sums
    const int Qn = mt.fverts ? mt.n_fpts : P;
    const unsigned* cp = colpart + ((size_t)b * C + c) * kSplit * (size_t)Qn;
    const float th[6] = {0.002f, 0.005f, 0.010f, 0.020f, 0.050f, 0.100f};
    double s_gp = 0.0;
    int hit[6] = {0, 0, 0, 0, 0, 0};
    for (int j = tid; j < Qn; j += kMT) {
      unsigned m = cp[j];
      for (int sp = 1; sp < kSplit; ++sp) m = min(m, cp[(size_t)sp * Qn + j]);
      const float d = sqrtf(__uint_as_float(m));
      s_gp += (double)d;
#pragma unroll
      for (int k = 0; k < 6; ++k) hit[k] += d < th[k] ? 1 : 0;
    }
    tot[2] = block_sum_d(s_gp, red);
    for (int k = 0; k < 6; ++k) tot[9 + k] = block_sum_d((double)hit[k], red);
  }
  if (tid == 0) {
    // the 8 corners of compute_obj_metrics_dexycb: (x, y, z) picks min (0) or max (1) by these index rows
    const int cx[8] = {0, 1, 0, 0, 1, 0, 1, 1}, cy[8] = {0, 0, 1, 0, 1, 1, 0, 1}, cz[8] = {0, 0, 0, 1, 0, 1, 1, 1};
    float acc = 0.f;
    for (int k = 0; k < 8; ++k) {
      const float dx = bp[cx[k]][0] - bg[cx[k]][0], dy = bp[cy[k]][1] - bg[cy[k]][1], dz = bp[cz[k]][2] - bg[cz[k]][2];
      acc += sqrtf((dx * dx + dy * dy) + dz * dz);
    }
    row[2] = (double)(acc / 8.f);
    row[4] = (double)(float)add_m;
    row[5] = (double)(float)adds_m;
    row[6] = rep_px;
    const float diam = mt.diameter[o];
    row[14] = ((float)add_m <= diam * 0.1f) ? 1.0 : 0.0;
    row[15] = ((float)adds_m <= diam * 0.1f) ? 1.0 : 0.0;
    row[16] = (rep_px < 5.0) ? 1.0 : 0.0;
    // F-score / Chamfer on the F-score cloud, both directions
    const int Q = mt.fverts ? mt.n_fpts : P;
    const double m_pg = tot[1] / (double)Q, m_gp = tot[2] / (double)Q;
    row[7] = (double)(0.5f * ((float)m_pg + (float)m_gp));
    for (int k = 0; k < 6; ++k) {
      const float prec = (float)tot[3 + k] / (float)Q, rec = (float)tot[9 + k] / (float)Q;
      row[8 + k] = (double)((2.f * prec * rec) / (prec + rec + 1e-6f));
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// The whole metric step of Trainer.evaluate for one batch (lib/engine/train_diff_hand_obj.py:224-269) behind one call:
// postprocess (:578-602: un-flip left hands, add the root joint; rot6d + translation -> [R | t + root]), TesterHand on the
// aggregated / first-candidate (/ regression) hands, TesterObject on the aggregated / first-candidate object poses, packed
// into one float64 row per image.
// ------------------------------------------------------------------------------------------------------------
struct EvalDev {
  vpho_eval_record_args a;
  int n_sets;              // hand predictions evaluated per image: 2, or 3 with the regression hand
  float* joint;            // [n_sets][bs][21][3]  post-processed predictions
  float* vert;             // [n_sets][bs][778][3]
  double* obj_rt;          // [bs][2][12]
  float* hand_metrics;     // [n_sets][bs][25]
  double* obj_metrics;     // [bs][2][17]
};

__global__ void __launch_bounds__(256) k_eval_prepare(EvalDev e) {
  const int b = blockIdx.x, tid = threadIdx.x, bs = e.a.bs, S = e.a.S;
  const float sign = e.a.is_right[b] ? 1.f : -1.f;
  const float r[3] = {e.a.root_joint[b * 3 + 0], e.a.root_joint[b * 3 + 1], e.a.root_joint[b * 3 + 2]};
  for (int set = 0; set < e.n_sets; ++set) {
    const float* pj = set == 0 ? e.a.agg_hand_joint + (size_t)b * 63
                               : (set == 1 ? e.a.cand_hand_joint + (size_t)b * S * 63 : e.a.reg_hand_joint + (size_t)b * 63);
    const float* pv = set == 0 ? e.a.agg_hand_vert + (size_t)b * kVerts * 3
                               : (set == 1 ? e.a.cand_hand_vert + (size_t)b * S * kVerts * 3 : e.a.reg_hand_vert + (size_t)b * kVerts * 3);
    float* dj = e.joint + ((size_t)set * bs + b) * 63;
    float* dv = e.vert + ((size_t)set * bs + b) * kVerts * 3;
    for (int i = tid; i < (21 + kVerts) * 3; i += 256) {
      const int d = i % 3;
      const float x = i < 63 ? pj[i] : pv[i - 63];
      const float y = (d == 0 ? sign * x : x) + r[d];       // __postprocess_hand_vert (:598-602)
      if (i < 63) dj[i] = y; else dv[i - 63] = y;
    }
  }
  if (tid < 2) {
    const double* p = tid == 0 ? e.a.agg_obj_6d + (size_t)b * 9 : e.a.cand_obj_6d + (size_t)b * S * 9;
    double R[9];
    rot6d_to_matrix(p, R);                                  // obj_9D_to_mat; translation + root joint (:593-596)
    double* o = e.obj_rt + ((size_t)b * 2 + tid) * 12;
    for (int row = 0; row < 3; ++row) {
      for (int cc = 0; cc < 3; ++cc) o[row * 4 + cc] = R[row * 3 + cc];
      o[row * 4 + 3] = p[6 + row] + (double)r[row];
    }
  }
}

__global__ void k_eval_pack(EvalDev e) {
  const int b = blockIdx.x, bs = e.a.bs, W = e.n_sets * 25 + 2 * kMetricCols;
  double* row = e.a.out + (size_t)b * W;
  for (int i = threadIdx.x; i < W; i += blockDim.x) {
    if (i < e.n_sets * 25) row[i] = (double)e.hand_metrics[((size_t)(i / 25) * bs + b) * 25 + i % 25];
    else row[i] = e.obj_metrics[(size_t)b * 2 * kMetricCols + (i - e.n_sets * 25)];
  }
}

// contact.cu
int launch_hand_metrics_full(const float* pd_joint, const float* gt_joint, const float* pd_vert, const float* gt_vert, int rows, int gt_rows,
                             float* metrics, cudaStream_t st);

}  // namespace vpho

using namespace vpho;

extern "C" int vpho_objmetrics_create(int n_obj, const float* bbox3d, const float* diameter, int sym_k, const double* sym_R,
                                      const double* sym_t, const int32_t* sym_count, int n_fpts, const float* fverts,
                                      vpho_objmetrics_t* out) {
  if (n_obj <= 0 || !bbox3d || !diameter || sym_k <= 0 || !sym_R || !sym_t || !out || (fverts && n_fpts <= 0)) return VPHO_ERR_INVALID;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = (off + bytes + 255) / 256 * 256; return o; };
  const size_t o_box = take((size_t)n_obj * 24 * 4), o_diam = take((size_t)n_obj * 4), o_R = take((size_t)n_obj * sym_k * 9 * 8),
               o_t = take((size_t)n_obj * sym_k * 3 * 8), o_cnt = take((size_t)n_obj * 4),
               o_fv = take(fverts ? (size_t)n_obj * n_fpts * 3 * 4 : 0);
  std::vector<char> h(off, 0);
  memcpy(h.data() + o_box, bbox3d, (size_t)n_obj * 24 * 4);
  memcpy(h.data() + o_diam, diameter, (size_t)n_obj * 4);
  memcpy(h.data() + o_R, sym_R, (size_t)n_obj * sym_k * 9 * 8);
  memcpy(h.data() + o_t, sym_t, (size_t)n_obj * sym_k * 3 * 8);
  if (sym_count) memcpy(h.data() + o_cnt, sym_count, (size_t)n_obj * 4);
  if (fverts) memcpy(h.data() + o_fv, fverts, (size_t)n_obj * n_fpts * 3 * 4);
  ObjMetricHost* mh = new ObjMetricHost();
  if (cudaMalloc(&mh->blob, off) != cudaSuccess) { delete mh; return VPHO_ERR_ALLOC; }
  if (cudaMemcpy(mh->blob, h.data(), off, cudaMemcpyHostToDevice) != cudaSuccess) { cudaFree(mh->blob); delete mh; return VPHO_ERR_ALLOC; }
  char* b = static_cast<char*>(mh->blob);
  mh->dev.n_obj = n_obj; mh->dev.sym_k = sym_k; mh->dev.n_fpts = n_fpts;
  mh->dev.bbox3d = (const float*)(b + o_box); mh->dev.diameter = (const float*)(b + o_diam);
  mh->dev.sym_R = (const double*)(b + o_R); mh->dev.sym_t = (const double*)(b + o_t); mh->dev.sym_count = (const int*)(b + o_cnt);
  mh->dev.fverts = fverts ? (const float*)(b + o_fv) : nullptr;
  *out = mh;
  return VPHO_OK;
}

extern "C" int vpho_objmetrics_destroy(vpho_objmetrics_t h) {
  if (!h) return VPHO_ERR_INVALID;
  ObjMetricHost* mh = static_cast<ObjMetricHost*>(h);
  if (mh->part) cudaFree(mh->part);
  if (mh->colpart) cudaFree(mh->colpart);
  cudaFree(mh->blob);
  delete mh;
  return VPHO_OK;
}

static int launch_object_metrics(const AssetsDev& as, ObjMetricHost* mh, const double* pd_rt, const double* gt_rt, const int32_t* obj_id,
                                 const float* cam_intr, int n, int C, double* out, cudaStream_t st);

extern "C" int vpho_object_metrics(vpho_assets_t assets, vpho_objmetrics_t tables, const double* pd_rt, const double* gt_rt,
                                   const int32_t* obj_id, const float* cam_intr, int n, int C, double* out, void* stream) {
  if (!assets || !tables || n < 0 || C < 0) return VPHO_ERR_INVALID;
  if (n == 0 || C == 0) return VPHO_OK;
  if (!pd_rt || !gt_rt || !obj_id || !cam_intr || !out) return VPHO_ERR_INVALID;
  const AssetsDev& as = static_cast<AssetsHost*>(assets)->dev;
  const ObjMetricDev& mt = static_cast<ObjMetricHost*>(tables)->dev;
  if (mt.n_obj != as.n_obj) return VPHO_ERR_INVALID;
  return launch_object_metrics(as, static_cast<ObjMetricHost*>(tables), pd_rt, gt_rt, obj_id, cam_intr, n, C, out, (cudaStream_t)stream);
}

static int launch_object_metrics(const AssetsDev& as, ObjMetricHost* mh, const double* pd_rt, const double* gt_rt, const int32_t* obj_id,
                                 const float* cam_intr, int n, int C, double* out, cudaStream_t st) {
  const ObjMetricDev& mt = mh->dev;
  const size_t rows = (size_t)n * C;
  const int Q = mt.fverts ? mt.n_fpts : as.n_pts;
#ifdef VPHO_EMU
  mh->sym = 0;
#else
  if (mh->sym < 0) {
    mh->sym = 0;
    const size_t need = (size_t)Q * sizeof(unsigned);
    if (need <= 160 * 1024 &&
        (need <= 32 * 1024 || cudaFuncSetAttribute(k_object_metrics_scan, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need) == cudaSuccess))
      mh->sym = 1;
  }
#endif
  if (rows > mh->part_rows) {
    if (mh->part) { cudaDeviceSynchronize(); cudaFree(mh->part); cudaFree(mh->colpart); mh->part = nullptr; mh->colpart = nullptr; mh->part_rows = 0; }
    if (cudaMalloc((void**)&mh->part, rows * kSplit * kPartCols * sizeof(double)) != cudaSuccess) return VPHO_ERR_ALLOC;
    if (mh->sym && cudaMalloc((void**)&mh->colpart, rows * kSplit * (size_t)Q * sizeof(unsigned)) != cudaSuccess) return VPHO_ERR_ALLOC;
    mh->part_rows = rows;
  }
  const size_t dyn = mh->sym ? (size_t)Q * sizeof(unsigned) : 0;
  VPHO_LAUNCH(k_object_metrics_scan, dim3(kSplit, C, n), dim3(kMT), dyn, st, as, mt, pd_rt, gt_rt, obj_id, C, mh->part, mh->colpart, mh->sym);
  VPHO_LAUNCH(k_object_metrics, dim3(C, n), dim3(kMT), 0, st, as, mt, pd_rt, gt_rt, obj_id, cam_intr, C, mh->part, mh->colpart, mh->sym, out);
  VPHO_CHECK_LAUNCH();
  return VPHO_OK;
}

extern "C" size_t vpho_eval_record_workspace_bytes(int bs) {
  if (bs <= 0) return 0;
  return (size_t)bs * (3 * (21 + kVerts) * 3 * 4 + 2 * 12 * 8 + 3 * 25 * 4 + 2 * kMetricCols * 8) + 5 * 256;
}

extern "C" int vpho_eval_record(vpho_assets_t assets, vpho_objmetrics_t tables, const vpho_eval_record_args* args, void* workspace,
                                size_t workspace_bytes, void* stream) {
  if (!assets || !tables || !args || !workspace) return VPHO_ERR_INVALID;
  const vpho_eval_record_args& a = *args;
  if (a.bs < 0 || a.S <= 0) return VPHO_ERR_INVALID;
  if (a.bs == 0) return VPHO_OK;
  if (!a.agg_hand_joint || !a.agg_hand_vert || !a.cand_hand_joint || !a.cand_hand_vert || !a.agg_obj_6d || !a.cand_obj_6d || !a.root_joint ||
      !a.is_right || !a.gt_joint || !a.gt_vert || !a.gt_obj_rt || !a.cam_intr || !a.obj_id || !a.out || (!a.reg_hand_joint != !a.reg_hand_vert))
    return VPHO_ERR_INVALID;
  if (workspace_bytes < vpho_eval_record_workspace_bytes(a.bs)) return VPHO_ERR_INVALID;
  const AssetsDev& as = static_cast<AssetsHost*>(assets)->dev;
  ObjMetricHost* mh = static_cast<ObjMetricHost*>(tables);
  if (mh->dev.n_obj != as.n_obj) return VPHO_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  EvalDev e;
  e.a = a;
  e.n_sets = a.reg_hand_joint ? 3 : 2;
  char* p = static_cast<char*>(workspace);
  auto take = [&](size_t bytes) { char* q = p; p += (bytes + 255) / 256 * 256; return q; };
  e.obj_rt = reinterpret_cast<double*>(take((size_t)a.bs * 2 * 12 * 8));
  e.obj_metrics = reinterpret_cast<double*>(take((size_t)a.bs * 2 * kMetricCols * 8));
  e.joint = reinterpret_cast<float*>(take((size_t)e.n_sets * a.bs * 63 * 4));
  e.vert = reinterpret_cast<float*>(take((size_t)e.n_sets * a.bs * kVerts * 3 * 4));
  e.hand_metrics = reinterpret_cast<float*>(take((size_t)e.n_sets * a.bs * 25 * 4));
  VPHO_LAUNCH(k_eval_prepare, dim3(a.bs), dim3(256), 0, st, e);
  int rc = launch_hand_metrics_full(e.joint, a.gt_joint, e.vert, a.gt_vert, e.n_sets * a.bs, a.bs, e.hand_metrics, st);
  if (rc) return rc;
  rc = launch_object_metrics(as, mh, e.obj_rt, a.gt_obj_rt, a.obj_id, a.cam_intr, a.bs, 2, e.obj_metrics, st);
  if (rc) return rc;
  VPHO_LAUNCH(k_eval_pack, dim3(a.bs), dim3(128), 0, st, e);
  VPHO_CHECK_LAUNCH();
  return VPHO_OK;
}
