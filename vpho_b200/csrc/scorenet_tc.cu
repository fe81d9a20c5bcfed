// Score network on the 5th-generation tensor cores (tcgen05) with FP32-class accuracy by operand splitting.
//
// Head GEMM (k_head_tc), same contract as k_head_simt (sampler.cu): for every 128-candidate row tile and ParallelLinear head
//     hidden[128][256] = P2[128][256] . Wa_p[head][256][256]      (lib/model/parallel_linear.py:27-35, K = pose features)
//     out[row][0..2]   = relu(hidden + F[img] + Tt) . Wb[head] + bb ;  score = out / (sigma(t) + 1e-7)
// with the contraction run as  A_lo.B_hi + A_hi.B_lo + A_hi.B_hi  on FP16 (hi, lo) planes (kind::f16 UMMA):
//   * operands are scaled by exact powers of two so that every candidate row / head peaks in [2^13, 2^14):
//     hi = half(x s), lo = half(x s - hi) is a 22-bit split; the dropped lo.lo term leaves ~2^-21 per product;
//   * operands staged in shared memory by TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B, K-major 128-byte rows),
//     mbarrier pipeline, one producer lane;
//   * tcgen05.mma issued by one elected lane, FP32 accumulators in TMEM (2 x 256 columns, double-buffered across items);
//   * epilogue warps read the accumulator with tcgen05.ld (one candidate row per thread) and fuse the un-scaling,
//     bias / conditioning / time terms, ReLU, the 256 -> 3 second ParallelLinear and the sigma division.
// Persistent grid: CTA (pair) b processes items b, b + n_units, ... of the (head, row tile) list.
// Pose encoder (k_pose_tc): D -> 256 -> 256 ReLU MLP per 128-row tile, GEMM 1 as 3xTF32, GEMM 2 as 3xFP16.
#include "sampler_device.cuh"
#include "tc_ptx.cuh"
#include "vpho_b200.h"

#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_fp16.h>

namespace vpho {

constexpr int kTcBM = 128, kTcBN = 256, kTcBK = 32, kTcStages = 2, kTcUmmaK = 8;
constexpr int kTcABytes = kTcBM * kTcBK * 4;        // 16 KB per operand plane per stage
constexpr int kTcBBytes = kTcBN * kTcBK * 4;        // 32 KB
constexpr int kTcStageBytes = 2 * kTcABytes + 2 * kTcBBytes;   // 96 KB
constexpr int kTcMaxStages = 3;      // CTA-pair head GEMM: 3 stages of 64 KB in the same 192 KB
constexpr int kHeadThreads = 640;      // head kernel: 4 role warps + 2 x 8 epilogue warps
constexpr int kTcFtImgs = 4;
constexpr int kMaxDevices = 64;       // per-device caches of function attributes
constexpr uint32_t kTcIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kTcBN >> 3) << 17) | ((uint32_t)(kTcBM >> 4) << 24);
// FP16 operands (a_format = b_format = 0), FP32 accumulate: same tile, K = 16 per instruction, twice the TF32 rate
constexpr uint32_t kTcIdescF16 = (1u << 4) | ((uint32_t)(kTcBN >> 3) << 17) | ((uint32_t)(kTcBM >> 4) << 24);

struct TcSmem {
  // operand stages first: 1024-byte aligned swizzle atoms
  unsigned char stage[kTcStages][kTcStageBytes];
  float wb[2][kTcBN][4];
  float tt[2][kTcBN];
  float part[2][kTcBM][4];      // partial 256->3 sums of the upper column half
  float ft[2][kTcFtImgs][kTcBN]; // F[img] + Tt of the images a row tile spans
  unsigned long long full_bar[kTcMaxStages], empty_bar[kTcMaxStages], tmem_full_bar[2], tmem_empty_bar[2];
  uint32_t tmem_base;
};

// Optional timeline instrumentation of CTA 0 (vpho_debug_tc_clocks, tools/diag_tc_timeline.py): %globaltimer stamps of
// the three roles of the head GEMM.  Compiled in only with -DVPHO_TC_TIMELINE; the default build has no stamps.
__device__ unsigned long long g_clk[3 * 2048];
__device__ int g_clk_on = 0;       // 1: every launch stamps; 100 + s: only the pose kernel's calls of RK stage s do
__device__ __forceinline__ void clk_stamp(int role, int idx) {
#ifdef VPHO_TC_TIMELINE
  if (g_clk_on && blockIdx.x == 0 && idx < 2048) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_clk[role * 2048 + idx] = t;
  }
#else
  (void)role; (void)idx;
#endif
}

// Operands are __half planes (64 k per 128-byte row, 4 chunks of K = 256); the epilogue undoes the power-of-two scales.
// kCtas = 2: CTA-pair variant.  The two CTAs of a cluster take the row tiles 2p and 2p + 1 of the same head; each loads its
//                own A tile and HALF of the head's weight tile (128 of the 256 hidden columns), the leader CTA issues one
//                tcgen05.mma.cta_group::2 of M = 256 that reads both halves, so the L2 -> SM operand traffic per CTA drops
//                from 384 KB to 256 KB per work item and the 192 KB of stages hold 3 x 64 KB instead of 2 x 96 KB.
//                TMA completions of both CTAs are counted on the leader's full barrier; tcgen05.commit multicasts the
//                "stage free" and "accumulator ready" arrivals to both CTAs; the peer's epilogue warps arrive remotely
//                on the leader's "accumulator drained" barrier.
// Two samplers in lock-step (n_jobs = 2): the hand and the object integrations of one batch issue the same sequence of
// network calls, so one launch serves both -- job 0's work items followed by job 1's, dealt to the CTAs from opposite ends
// so that the CTAs with one item fewer of job 0 take job 1's first.  Each job has its own operand maps, weights, workspace
// and RK controller; a job whose controller is not active in this call (finished, or a skipped attempt) contributes nothing.
template <int kCtas>
__global__ void __launch_bounds__(kHeadThreads, 1)
k_head_tc(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
          const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
          const __grid_constant__ CUtensorMap tmA_hi1, const __grid_constant__ CUtensorMap tmA_lo1,
          const __grid_constant__ CUtensorMap tmB_hi1, const __grid_constant__ CUtensorMap tmB_lo1, DenoiserDev dn0, SamplerWs ws0,
          DenoiserDev dn1, SamplerWs ws1, int n_jobs, int mode, int s) {
  // A skipped RK attempt (spare launches behind a finished integration) leaves before any set-up.  Reading the status
  // word ahead of pdl_wait is safe as a one-way hint: it is written by k_reduce, which has completed before the kernel
  // that triggered this launch passed its own pdl_wait, and it never returns to "running" within a sample; a stale
  // "running" only sends us down the normal path, which checks again after the wait.  The wait itself is still executed
  // so that the launch chain stays transitively ordered.
  if (!eval_active(*ws0.ctrl, mode) && !(n_jobs > 1 && eval_active(*ws1.ctrl, mode))) {
    pdl_trigger();
    pdl_wait();
    return;
  }
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // align inside the shared window with an offset (an integer round trip would demote every access to a generic load)
  TcSmem& sm = *reinterpret_cast<TcSmem*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr bool kPair = kCtas == 2;
  constexpr int kStages = kPair ? 3 : kTcStages;
  constexpr int kBHalfBytes = kTcBBytes / kCtas;                   // B plane bytes this CTA stages per K chunk
  constexpr int kStageBytes = 2 * kTcABytes + 2 * kBHalfBytes;     // 96 KB, or 64 KB per CTA of a pair
  static_assert(kStages * kStageBytes <= kTcStages * kTcStageBytes, "stages must fit the operand region");
  const uint32_t rank = kPair ? cluster_ctarank() : 0u;
  const int unit = kPair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;           // CTA or CTA pair
  const int n_units = kPair ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  constexpr int kElems = 64;           // __half operand elements per 128-byte row
  constexpr int kChunks = kPDim / kElems;

  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(&sm.full_bar[i], 1); mbar_init(&sm.empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&sm.tmem_full_bar[i], 1); mbar_init(&sm.tmem_empty_bar[i], 8 * kCtas); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    if (kPair) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "r"(512u));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "r"(512u));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
  }
  tc_fence_before();
  if (kPair) cluster_sync_all();                // barrier inits visible to the peer before any remote arrive / complete_tx
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sm.tmem_base;
  // launched with launch_pdl: everything above overlaps the tail of the pose-encoder kernel; nothing it (or any earlier
  // kernel) wrote is read before this point
  pdl_wait();
  pdl_trigger();
  // per job: work items (head, row tile or pair of row tiles) and the first one this CTA / CTA pair takes
  const bool act0 = eval_active(*ws0.ctrl, mode), act1 = n_jobs > 1 && eval_active(*ws1.ctrl, mode);
  const int tiles0 = ws0.Npad / kTcBM, tiles1 = ws1.Npad / kTcBM;
  const int slots0 = kPair ? (tiles0 + 1) / 2 : tiles0, slots1 = kPair ? (tiles1 + 1) / 2 : tiles1;
  const int items0 = act0 ? slots0 * dn0.n_heads : 0, items1 = act1 ? slots1 * dn1.n_heads : 0;
  const int first0 = unit, first1 = n_units - 1 - unit;
  const int cnt0 = items0 > first0 ? (items0 - first0 + n_units - 1) / n_units : 0;     // items of job 0 this CTA takes

  if (warp == 0) {
    // ===================================================== TMA producer (one lane per CTA)
    if (lane == 0) {
      int stage = 0, pc = 0;
      uint32_t phase = 0;
      auto produce = [&](const CUtensorMap* mA_hi, const CUtensorMap* mA_lo, const CUtensorMap* mB_hi, const CUtensorMap* mB_lo,
                         int n_slots, int n_items, int first) {
        for (int it = first; it < n_items; it += n_units) {
          const int slot = it % n_slots, head = it / n_slots;
          const int tile = kPair ? 2 * slot + (int)rank : slot;       // a tile past the end (odd count) loads zeros
          const int brow = head * kTcBN + (int)rank * (kTcBN / kCtas);
          for (int kc = 0; kc < kChunks; ++kc) {
            mbar_wait(&sm.empty_bar[stage], phase ^ 1);
            clk_stamp(0, 2 * pc);
            unsigned char* st = sm.stage[0] + stage * kStageBytes;
            if (kPair) {
              if (rank == 0) mbar_arrive_expect_tx(&sm.full_bar[stage], 2 * kStageBytes);
              const uint32_t lbar = mapa_rank(smem_u32(&sm.full_bar[stage]), 0);
              tma_load_2d_pair(mA_hi, lbar, st, kc * kElems, tile * kTcBM);
              tma_load_2d_pair(mA_lo, lbar, st + kTcABytes, kc * kElems, tile * kTcBM);
              tma_load_2d_pair(mB_hi, lbar, st + 2 * kTcABytes, kc * kElems, brow);
              tma_load_2d_pair(mB_lo, lbar, st + 2 * kTcABytes + kBHalfBytes, kc * kElems, brow);
            } else {
              mbar_arrive_expect_tx(&sm.full_bar[stage], kStageBytes);
              tma_load_2d(mA_hi, &sm.full_bar[stage], st, kc * kElems, tile * kTcBM);
              tma_load_2d(mA_lo, &sm.full_bar[stage], st + kTcABytes, kc * kElems, tile * kTcBM);
              tma_load_2d(mB_hi, &sm.full_bar[stage], st + 2 * kTcABytes, kc * kElems, brow);
              tma_load_2d(mB_lo, &sm.full_bar[stage], st + 2 * kTcABytes + kBHalfBytes, kc * kElems, brow);
            }
            clk_stamp(0, 2 * pc + 1);
            ++pc;
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
      };
      produce(&tmA_hi, &tmA_lo, &tmB_hi, &tmB_lo, slots0, items0, first0);
      produce(&tmA_hi1, &tmA_lo1, &tmB_hi1, &tmB_lo1, slots1, items1, first1);
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer (the leader CTA's lane issues for the pair)
    if (lane == 0 && rank == 0) {
      int stage = 0, acc = 0, mc = 0;
      uint32_t phase = 0, acc_phase = 0;
      constexpr uint32_t kIdesc = kTcIdescF16 + (kPair ? ((uint32_t)(kTcBM >> 4) << 24) : 0u);   // M = 128 * kCtas
      const int cnt1 = items1 > first1 ? (items1 - first1 + n_units - 1) / n_units : 0;
      for (int n_it = 0; n_it < cnt0 + cnt1; ++n_it) {          // every item is the same 128 (x2) x 256 x 256 product
        clk_stamp(1, mc++);
        if (kPair) mbar_wait_cluster(&sm.tmem_empty_bar[acc], acc_phase ^ 1);
        else mbar_wait(&sm.tmem_empty_bar[acc], acc_phase ^ 1);
        clk_stamp(1, mc++);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * kTcBN);
        for (int kc = 0; kc < kChunks; ++kc) {
          mbar_wait(&sm.full_bar[stage], phase);
          clk_stamp(1, mc++);
          tc_fence_after();
          unsigned char* st = sm.stage[0] + stage * kStageBytes;
          const uint64_t a_hi = make_kmajor_sw128_desc(st), a_lo = make_kmajor_sw128_desc(st + kTcABytes);
          const uint64_t b_hi = make_kmajor_sw128_desc(st + 2 * kTcABytes), b_lo = make_kmajor_sw128_desc(st + 2 * kTcABytes + kBHalfBytes);
#pragma unroll
          for (int k = 0; k < kTcBK / kTcUmmaK; ++k) {
            const uint64_t adv = (uint64_t)((k * kTcUmmaK * 4) >> 4);     // 32 bytes per K step inside the 128-byte row
            const uint32_t first = (kc | k) != 0 ? 1u : 0u;
            if (kPair) {
              umma_f16_pair(d_tmem, a_lo + adv, b_hi + adv, kIdesc, first);
              umma_f16_pair(d_tmem, a_hi + adv, b_lo + adv, kIdesc, 1u);
              umma_f16_pair(d_tmem, a_hi + adv, b_hi + adv, kIdesc, 1u);
            } else {
              umma_f16(d_tmem, a_lo + adv, b_hi + adv, kIdesc, first);
              umma_f16(d_tmem, a_hi + adv, b_lo + adv, kIdesc, 1u);
              umma_f16(d_tmem, a_hi + adv, b_hi + adv, kIdesc, 1u);
            }
          }
          // frees the smem slot (of both CTAs) once these MMAs have read it
          if (kPair) umma_commit_pair(&sm.empty_bar[stage]);
          else umma_commit(&sm.empty_bar[stage]);
          clk_stamp(1, mc++);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        // accumulator complete -> epilogue (of both CTAs)
        if (kPair) umma_commit_pair(&sm.tmem_full_bar[acc]);
        else umma_commit(&sm.tmem_full_bar[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================================================== epilogue: two groups of 8 warps.  Group g drains accumulator
    // buffer g, i.e. every other work item, so the two groups' epilogues overlap each other and the MMAs.  Inside a group
    // warp e reads TMEM lanes 32*(e&3).. (its hardware lane quarter, warp_id % 4) and the 128-column half (e>>2)&1; the
    // two partial 256->3 sums of a row are combined through shared memory in a fixed order (low columns first).
    const int e = warp - 4, q = e & 3, half = (e >> 2) & 1, grp = e >> 3;
    const int tg = (threadIdx.x - 128) & 255;            // thread index inside the group
    int ec = 0;
    uint32_t acc_phase = 0;
    const int acc = grp;
    const uint32_t drained_bar = kPair ? mapa_rank(smem_u32(&sm.tmem_empty_bar[acc]), 0) : 0u;
    // Per-item constants of column tg (second-layer weights, time term, conditioning terms of the images the row tile spans)
    // are fetched one item ahead into registers, so their global-memory latency hides behind the previous item's math.
    struct ItemConsts { float4 wb; float tt; float f[kTcFtImgs]; };
    // `seq0`: position of this job's first item in the CTA's item sequence (group g takes the positions of parity g)
    auto drain = [&](const DenoiserDev& dn, const SamplerWs& ws, int n_slots, int n_items, int first, int seq0) {
      RkCtrl& c = *ws.ctrl;
      const EvalTime et = c.et[tt_slot(mode, s)];          // written by the time-term block of this call (or of the attempt's first)
      const int n_rows = (mode == kModeEval) ? ws.eval_rows : c.n_rows;
      const int rpf = (mode == kModeEval) ? ws.eval_rpf : c.rows_per_feat;
      auto item_geometry = [&](int it, int& tile, int& head, int& img0, int& n_img) {
        const int slot = it % n_slots;
        head = it / n_slots;
        tile = kPair ? 2 * slot + (int)rank : slot;
        const int row_lo = tile * kTcBM, row_hi = min(row_lo + kTcBM, n_rows) - 1;
        img0 = row_lo / rpf;
        n_img = row_hi >= row_lo ? row_hi / rpf - img0 + 1 : 0;
      };
      auto fetch_consts = [&](int it, ItemConsts& k) {
        int tile, head, img0, n_img;
        item_geometry(it, tile, head, img0, n_img);
        const int col = head * kHeadHid + tg;
        VPHO_BOUNDS(col < dn.hid && head < dn.n_heads && (n_img > kTcFtImgs || n_img == 0 || img0 + n_img <= ws.R));
        k.tt = ws.Tt[(size_t)tt_slot(mode, s) * dn.hid + col];
        k.wb = __ldg(reinterpret_cast<const float4*>(dn.Wb + (size_t)col * 4));
#pragma unroll
        for (int i = 0; i < kTcFtImgs; ++i) k.f[i] = (n_img <= kTcFtImgs && i < n_img) ? ws.F[(size_t)(img0 + i) * dn.hid + col] : 0.f;
      };
      const int it_first = first + (((seq0 & 1) != grp) ? n_units : 0), it_step = 2 * n_units;
      ItemConsts kc;
      if (it_first < n_items) fetch_consts(it_first, kc);
      for (int it = it_first; it < n_items; it += it_step) {
        int tile, head, img0, n_img;
        item_geometry(it, tile, head, img0, n_img);
        const int hc0 = head * kHeadHid;
        if (tg == 0 && grp == 0) clk_stamp(2, ec++);
        // ft[i][col] = F[img0 + i][col] + Tt[col] for the (at most kTcFtImgs) images this 128-row tile spans; tiles that span
        // more images read F from global memory instead
        const int row_lo = tile * kTcBM;
        const bool ft_smem = n_img <= kTcFtImgs;
        sm.tt[acc][tg] = kc.tt;
        *reinterpret_cast<float4*>(sm.wb[acc][tg]) = kc.wb;
#pragma unroll
        for (int i = 0; i < kTcFtImgs; ++i) sm.ft[acc][i][tg] = kc.f[i] + kc.tt;
        const int row = row_lo + q * 32 + lane;
        const bool valid = row < n_rows;
        const int cbase = half * 128;
        const int img_l = valid ? row / rpf - img0 : 0;
        VPHO_BOUNDS(!valid || row / rpf < ws.R);
        const float* Frow = ws.F + (size_t)(valid ? row / rpf : 0) * dn.hid + hc0 + cbase;
        // exact power-of-two un-scaling of the FP16 operand planes; loaded before the waits
        const float unscale = (row < ws.Npad ? ws.P2scale[row] : 1.f) * dn.Wscale_inv[head];
        if (grp == 0) asm volatile("bar.sync 1, 256;" ::: "memory");
        else asm volatile("bar.sync 3, 256;" ::: "memory");
        if (it + it_step < n_items) fetch_consts(it + it_step, kc);
        if (tg == 0 && grp == 0) clk_stamp(2, ec++);
        mbar_wait(&sm.tmem_full_bar[acc], acc_phase);
        if (tg == 0 && grp == 0) clk_stamp(2, ec++);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * kTcBN + cbase);
        float o0 = 0.f, o1 = 0.f, o2 = 0.f;
        if (tg == 0 && grp == 0) clk_stamp(2, ec++);
        if (ft_smem) {
          // hot path: conditioning + time terms of this row's image and the second-layer weights come from shared memory;
          // the TMEM load of the next 16 columns is in flight while the current 16 are consumed
          const float* ftp = &sm.ft[acc][img_l][cbase];
          const float* wbp = &sm.wb[acc][cbase][0];
          auto consume = [&](const uint32_t (&vv)[16], int cb) {
            float4 fa[4], w[16];
#pragma unroll
            for (int j = 0; j < 4; ++j) fa[j] = *reinterpret_cast<const float4*>(ftp + cb * 16 + j * 4);
#pragma unroll
            for (int j = 0; j < 16; ++j) w[j] = *reinterpret_cast<const float4*>(wbp + (cb * 16 + j) * 4);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float f = (j & 3) == 0 ? fa[j >> 2].x : (j & 3) == 1 ? fa[j >> 2].y : (j & 3) == 2 ? fa[j >> 2].z : fa[j >> 2].w;
              float hval = fmaf(__uint_as_float(vv[j]), unscale, f);
              hval = hval > 0.f ? hval : 0.f;
              o0 = fmaf(hval, w[j].x, o0);
              o1 = fmaf(hval, w[j].y, o1);
              o2 = fmaf(hval, w[j].z, o2);
            }
          };
          uint32_t va[16], vb[16];
          tmem_ld16_issue(taddr, va);
#pragma unroll 1
          for (int cb = 0; cb < 8; cb += 2) {
            tmem_ld16_wait(va);
            tmem_ld16_issue(taddr + (uint32_t)((cb + 1) * 16), vb);
            consume(va, cb);
            if (tg == 0 && grp == 0) clk_stamp(2, ec++);
            tmem_ld16_wait(vb);
            if (cb + 2 < 8) tmem_ld16_issue(taddr + (uint32_t)((cb + 2) * 16), va);
            consume(vb, cb + 1);
          }
        } else {
          // a row tile that spans more than kTcFtImgs images (few candidates per image): F from global memory
#pragma unroll 1
          for (int cb = 0; cb < 16; ++cb) {
            uint32_t vv[8];
            tmem_ld8(taddr + (uint32_t)(cb * 8), vv);
            if (tg == 0 && grp == 0 && (cb & 3) == 3) clk_stamp(2, ec++);
#pragma unroll
            for (int j4 = 0; j4 < 2; ++j4) {
              const float4 f4 = __ldg(reinterpret_cast<const float4*>(Frow + cb * 8 + j4 * 4));
              const float4 t4 = *reinterpret_cast<const float4*>(&sm.tt[acc][cbase + cb * 8 + j4 * 4]);
              const float fa[4] = {f4.x + t4.x, f4.y + t4.y, f4.z + t4.z, f4.w + t4.w};
#pragma unroll
              for (int jj = 0; jj < 4; ++jj) {
                float hval = fmaf(__uint_as_float(vv[j4 * 4 + jj]), unscale, fa[jj]);
                hval = hval > 0.f ? hval : 0.f;
                const float4 w = *reinterpret_cast<const float4*>(sm.wb[acc][cbase + cb * 8 + j4 * 4 + jj]);
                o0 = fmaf(hval, w.x, o0);
                o1 = fmaf(hval, w.y, o1);
                o2 = fmaf(hval, w.z, o2);
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (kPair) mbar_arrive_cluster(drained_bar);        // the leader's MMA lane waits for both CTAs' 8 warps
          else mbar_arrive(&sm.tmem_empty_bar[acc]);
        }
        if (tg == 0 && grp == 0) clk_stamp(2, ec++);
        if (half == 1) {
          float* pp = sm.part[acc][q * 32 + lane];
          pp[0] = o0; pp[1] = o1; pp[2] = o2;
        }
        if (grp == 0) asm volatile("bar.sync 2, 256;" ::: "memory");
        else asm volatile("bar.sync 4, 256;" ::: "memory");
        if (half == 0 && valid) {
          const float* pp = sm.part[acc][q * 32 + lane];
          const float o[3] = {o0 + pp[0], o1 + pp[1], o2 + pp[2]};
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            const float out = o[d] + dn.bb[head * 3 + d];
            VPHO_BOUNDS(head * 3 + d < dn.D && row * dn.D + head * 3 + d < n_rows * dn.D);
            emit_score(ws, c, et, mode, s, row * dn.D + head * 3 + d, __fdiv_rn(out, et.std32));
          }
        }
        if (tg == 0 && grp == 0) clk_stamp(2, ec++);
        acc_phase ^= 1;
      }
    };
    if (items0 > 0) drain(dn0, ws0, slots0, items0, first0, 0);
    if (items1 > 0) drain(dn1, ws1, slots1, items1, first1, cnt0);
  }
  tc_fence_before();
  if (kPair) cluster_sync_all();                // neither CTA leaves (or frees TMEM) while the pair's MMAs / arrivals are in flight
  else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if (kPair) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
  }
}

// =====================================================================================================================
// Pose encoder on tensor cores: P2 = relu(relu(X.W1 + b1).W2 + b2) for one 128-row tile per CTA.
//   input    X = the float64 RK stage combination of the tile's state elements, formed by the compute warps and written as
//            (hi, lo) TF32 planes straight into the swizzled A operand (no stage-input kernel, no global round trip).  The
//            units are dealt chunk by chunk (32 k): every thread issues the loads of BOTH its units of a chunk before any
//            arithmetic, and GEMM 1 starts on chunk kc while the compute warps form chunk kc + 1.  Nothing on this path waits
//            for the controller words: addresses come from the kernel parameters, the scalars (step, weights, activity) are
//            read by an otherwise idle warp and handed over through shared memory while the state loads are in flight.
//   GEMM 1   D1[128x256] (TMEM cols 0..255) = X . W1 as 3xTF32, W1 (hi, lo) chunks of 32 k streamed by TMA;
//   re-stage two passes over D1 (tcgen05.ld is cheap, registers are not: 640 threads leave 96 each): pass A the row maximum
//            of relu(D1 + b1), pass B the scaled (hi, lo) FP16 split written as the next A operand chunk of GEMM 2
//            (double-buffered, reusing the X region) while the MMA lane consumes the previous one;
//   GEMM 2   D2[128x256] (TMEM cols 256..511) = H1 . W2 as 3xFP16;  epilogue: the same two passes -> P2hi / P2lo (row-major).
// =====================================================================================================================
constexpr int kPtMaxK1Chunks = 3;                         // D <= 96
constexpr int kPtRegionA = kPtMaxK1Chunks * 2 * kTcABytes;   // X hi/lo chunks, later 2 x (H1 hi/lo chunk): 96 KB
constexpr int kPtStageB = 2 * kTcBBytes;                  // W (hi, lo) chunk: 64 KB
// One role warpgroup (warp 0: TMEM allocation, controller scalars, TMA producer; warp 1: barrier set-up, MMA issuer; warps
// 2, 3 idle) + 16 compute warps.  Every scheduler holds 5 warps, i.e. 96 registers per thread at launch; the role warpgroup
// hands registers back (setmaxnreg 32) and the compute warpgroups take 112, so that the stage-input phase can keep two
// units' loads (64 registers) in flight without spilling.
constexpr int kPoseThreads = 640;
constexpr int kPoseRoleThreads = 128;
constexpr int kPoseRoleRegs = 48, kPoseComputeRegs = 104;
static_assert(kPoseRoleThreads * kPoseRoleRegs + (kPoseThreads - kPoseRoleThreads) * kPoseComputeRegs <= kPoseThreads * 96,
              "setmaxnreg.inc draws from the registers the CTA was launched with: the budget must close or the kernel deadlocks");

struct PtSmem {
  unsigned char a[kPtRegionA];
  unsigned char b[2][kPtStageB];
  float b1s[kPDim];                // first-layer bias: static weights, staged before the PDL wait
  StageScalars q;                  // controller scalars of this call (warp 3 -> compute warps, q_bar)
  unsigned long long full_bar[2], empty_bar[2], a_full_bar[2], a_empty_bar[2], d1_full_bar, d2_full_bar, x_full_bar[kPtMaxK1Chunks], q_bar;
  uint32_t tmem_base;
  int active;
};

// (hi, lo) __half split of 8 values scaled by the power of two `sc` -> two 16-byte units.  Two values per conversion
// instruction (F2FP.PACK_AB); the products with `sc` are exact, so x - float(hi) cannot be changed by contraction.
__device__ __forceinline__ void split8_f16(const float* x, float sc, uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const float a = x[2 * p] * sc, b = x[2 * p + 1] * sc;
    const __half2 h2 = __floats2half2_rn(a, b);
    const float2 f = __half22float2(h2);
    const __half2 l2 = __floats2half2_rn(a - f.x, b - f.y);
    h[p] = *reinterpret_cast<const uint32_t*>(&h2);
    l[p] = *reinterpret_cast<const uint32_t*>(&l2);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// exact power-of-two scale that puts a row maximum into [2^13, 2^14), and its inverse
__device__ __forceinline__ void row_scale(float rmax, float& sc, float& inv) {
  sc = 1.f; inv = 1.f;
  if (rmax > 0.f && rmax < 3.0e38f) {
    int ex = 0;
    frexpf(rmax, &ex);                           // rmax = m * 2^ex, m in [0.5, 1)
    sc = ldexpf(1.f, 14 - ex);
    inv = ldexpf(1.f, ex - 14);
  }
}

__device__ __forceinline__ void pose_tc_tile(const CUtensorMap* tmW1_hi, const CUtensorMap* tmW1_lo, const CUtensorMap* tmW2_hi,
                                             const CUtensorMap* tmW2_lo, const DenoiserDev& dn, const SamplerWs& ws, int mode, int s,
                                             int tile) {
  const RkCtrl& c = *ws.ctrl;
  if (!eval_active(c, mode)) {           // one-way hint ahead of the wait, see k_head_tc
    pdl_trigger();
    pdl_wait();
    return;
  }
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  PtSmem& sm = *reinterpret_cast<PtSmem*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int D = dn.D, nk1 = (D + kTcBK - 1) / kTcBK;
  const int r0 = tile * kTcBM;
  // GEMM 2 on FP16 planes: 4 chunks of 64 k, kind::f16, H1 re-staged as (hi, lo) halves scaled per row.  A chunk is 128-byte
  // rows: 16 KB per A plane, 32 KB per B plane.
  constexpr int nk2 = 4, kel2 = 64;

  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&sm.full_bar[i], 1); mbar_init(&sm.empty_bar[i], 1);
      mbar_init(&sm.a_full_bar[i], 512); mbar_init(&sm.a_empty_bar[i], 1);
    }
    for (int i = 0; i < kPtMaxK1Chunks; ++i) mbar_init(&sm.x_full_bar[i], 512);
    mbar_init(&sm.d1_full_bar, 1); mbar_init(&sm.d2_full_bar, 1); mbar_init(&sm.q_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    reinterpret_cast<float4*>(sm.b1s)[lane] = __ldg(reinterpret_cast<const float4*>(dn.b1) + lane);
    reinterpret_cast<float4*>(sm.b1s)[32 + lane] = __ldg(reinterpret_cast<const float4*>(dn.b1) + 32 + lane);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sm.tmem_base;
  const uint32_t d1 = tmem_base, d2 = tmem_base + 256;
#ifdef VPHO_TC_TIMELINE
  const int sb_ = (g_clk_on < 100 || (mode == kModeStage && s == g_clk_on - 100)) ? 0 : 4096;   // 4096: stamps dropped
#else
  constexpr int sb_ = 0;
#endif
  int sc_ = 1024 + sb_;                          // timeline stamp slot of this thread (VPHO_TC_TIMELINE builds)
  pdl_wait();                 // launched with launch_pdl: the set-up above overlaps the tail of the preceding kernel
  pdl_trigger();
  if (threadIdx.x == 0) clk_stamp(0, sc_++);
  if (threadIdx.x == kPoseRoleThreads) clk_stamp(2, 1101 + sb_);

  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kPoseRoleRegs));
  if (warp == 0) {
    if (lane == 0) {
      // the controller words of this call: one thread reads them while the compute warps' state loads are in flight
      sm.active = eval_active(c, mode) ? 1 : 0;
      sm.q = stage_scalars(c, mode, s);
      mbar_arrive(&sm.q_bar);
      if (sm.active) {
        int stage = 0;
        uint32_t phase = 0;
        // (the X tile, A operand of GEMM 1, is written into its dedicated region by the compute warps)
        for (int j = 0; j < nk1 + nk2; ++j) {
          mbar_wait(&sm.empty_bar[stage], phase ^ 1);
          clk_stamp(0, sc_++);
          mbar_arrive_expect_tx(&sm.full_bar[stage], kPtStageB);
          if (j < nk1) {
            tma_load_2d(tmW1_hi, &sm.full_bar[stage], sm.b[stage], j * kTcBK, 0);
            tma_load_2d(tmW1_lo, &sm.full_bar[stage], sm.b[stage] + kTcBBytes, j * kTcBK, 0);
          } else {
            tma_load_2d(tmW2_hi, &sm.full_bar[stage], sm.b[stage], (j - nk1) * kel2, 0);
            tma_load_2d(tmW2_lo, &sm.full_bar[stage], sm.b[stage] + kTcBBytes, (j - nk1) * kel2, 0);
          }
          if (++stage == 2) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      mbar_wait(&sm.q_bar, 0);
      if (sm.active) {
        int stage = 0;
        uint32_t phase = 0;
        auto mma_chunk = [&](uint32_t d_tmem, const unsigned char* a_hi_p, const unsigned char* a_lo_p, bool first) {
          const uint64_t a_hi = make_kmajor_sw128_desc(a_hi_p), a_lo = make_kmajor_sw128_desc(a_lo_p);
          const uint64_t b_hi = make_kmajor_sw128_desc(sm.b[stage]), b_lo = make_kmajor_sw128_desc(sm.b[stage] + kTcBBytes);
#pragma unroll
          for (int k = 0; k < kTcBK / kTcUmmaK; ++k) {
            const uint64_t adv = (uint64_t)((k * kTcUmmaK * 4) >> 4);
            umma_tf32(d_tmem, a_lo + adv, b_hi + adv, kTcIdesc, (first && k == 0) ? 0u : 1u);
            umma_tf32(d_tmem, a_hi + adv, b_lo + adv, kTcIdesc, 1u);
            umma_tf32(d_tmem, a_hi + adv, b_hi + adv, kTcIdesc, 1u);
          }
        };
        for (int kc = 0; kc < nk1; ++kc) {
          mbar_wait(&sm.x_full_bar[kc], 0);
          clk_stamp(1, sc_++);
          mbar_wait(&sm.full_bar[stage], phase);
          clk_stamp(1, sc_++);
          tc_fence_after();
          const unsigned char* chunk = sm.a + (size_t)kc * 2 * kTcABytes;
          mma_chunk(d1, chunk, chunk + kTcABytes, kc == 0);
          umma_commit(&sm.empty_bar[stage]);
          if (++stage == 2) { stage = 0; phase ^= 1; }
        }
        umma_commit(&sm.d1_full_bar);
        clk_stamp(1, sc_++);
        for (int kc = 0; kc < nk2; ++kc) {
          const int ab = kc & 1;
          mbar_wait(&sm.full_bar[stage], phase);
          clk_stamp(1, sc_++);
          mbar_wait(&sm.a_full_bar[ab], (uint32_t)((kc >> 1) & 1));
          clk_stamp(1, sc_++);
          tc_fence_after();
          const unsigned char* chunk = sm.a + (size_t)ab * 2 * kTcABytes;
          {
            const uint64_t a_hi = make_kmajor_sw128_desc(chunk), a_lo = make_kmajor_sw128_desc(chunk + kTcABytes);
            const uint64_t b_hi = make_kmajor_sw128_desc(sm.b[stage]), b_lo = make_kmajor_sw128_desc(sm.b[stage] + kTcBBytes);
#pragma unroll
            for (int k = 0; k < 4; ++k) {                    // 4 x K16 inside the 128-byte row
              const uint64_t adv = (uint64_t)((k * 32) >> 4);
              umma_f16(d2, a_lo + adv, b_hi + adv, kTcIdescF16, (kc == 0 && k == 0) ? 0u : 1u);
              umma_f16(d2, a_hi + adv, b_lo + adv, kTcIdescF16, 1u);
              umma_f16(d2, a_hi + adv, b_hi + adv, kTcIdescF16, 1u);
            }
          }
          umma_commit(&sm.empty_bar[stage]);
          umma_commit(&sm.a_empty_bar[ab]);
          if (++stage == 2) { stage = 0; phase ^= 1; }
        }
        umma_commit(&sm.d2_full_bar);
        clk_stamp(1, sc_++);
      }
    }
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kPoseComputeRegs));
    // 16 compute warps: warp w owns TMEM lanes 32*(w&3).. (its hardware lane quarter) and column sub-block cs = (w-4)>>2
    const int q = warp & 3, cs = (warp - 4) >> 2, r = q * 32 + lane;      // r: this thread's row of the tile
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    // ---- stage input: the float64 RK stage combination of every state element of this tile (what scipy hands to `fun`,
    // stage_input), rounded to float32 and split into the (hi, lo) TF32 planes of GEMM 1's A operand, straight into the
    // SWIZZLE_128B chunks (32 k per 128-byte row).  One thread = two 16-byte units (4 consecutive k of a row) per chunk;
    // consecutive threads take consecutive units, so the K-slot reads are coalesced.  Rows past the end and k >= D are zero.
    bool active = true;
    {
      const int tcx = threadIdx.x - kPoseRoleThreads;          // 0..511
      const int n_rows = ws.eval_rows;                         // the sampler's row count (carve): no controller read needed
      const int n_state = n_rows * D;
      const int ns = stage_ns(mode, s);
      const bool vec = (D & 3) == 0;                           // 16-byte aligned rows: vector loads
      for (int kc = 0; kc < nk1; ++kc) {
        StageRaw raw[2];
        bool ok[2];
        int idx[2];
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          const int w = tcx + 512 * b, rr = w >> 3, kq = kc * 8 + (w & 7);
          idx[b] = (r0 + rr) * D + 4 * kq;
          ok[b] = vec && r0 + rr < n_rows && 4 * kq < D;
          if (ok[b]) {
            VPHO_BOUNDS(idx[b] + 3 < n_state);
            stage_load4(ws, mode, ns, idx[b], n_state, raw[b]);
          }
        }
        if (kc == 0) {
          mbar_wait(&sm.q_bar, 0);
          active = sm.active != 0;
          if (threadIdx.x == kPoseRoleThreads) clk_stamp(2, 1102 + sb_);
          if (!active) break;
        }
        unsigned char* chunk = sm.a + (size_t)kc * 2 * kTcABytes;
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          const int w = tcx + 512 * b, rr = w >> 3, ku = w & 7;
          float x4[4] = {0.f, 0.f, 0.f, 0.f}, hi4[4], lo4[4];
          if (ok[b]) {
            stage_math4(ws, sm.q, mode, s, idx[b], raw[b], x4);
          } else if (!vec) {
#pragma unroll
            for (int e4 = 0; e4 < 4; ++e4) {
              const int k = 4 * (kc * 8 + ku) + e4;
              x4[e4] = k < D ? (float)stage_input(ws, c, mode, s, r0 + rr, k, n_rows, D) : 0.f;
            }
          }
#pragma unroll
          for (int e4 = 0; e4 < 4; ++e4) {
            hi4[e4] = tf32_round(x4[e4]);
            lo4[e4] = tf32_round(x4[e4] - hi4[e4]);
          }
          const uint32_t off = (uint32_t)((rr >> 3) * 1024 + (rr & 7) * 128 + (((ku ^ (rr & 7)) & 7) << 4));
          *reinterpret_cast<float4*>(chunk + off) = make_float4(hi4[0], hi4[1], hi4[2], hi4[3]);
          *reinterpret_cast<float4*>(chunk + kTcABytes + off) = make_float4(lo4[0], lo4[1], lo4[2], lo4[3]);
          if (kc == 0 && threadIdx.x == kPoseRoleThreads) clk_stamp(2, 1112 + b + sb_);
        }
        fence_proxy_async();
        mbar_arrive(&sm.x_full_bar[kc]);
        if (threadIdx.x == kPoseRoleThreads) clk_stamp(2, 1103 + (kc == nk1 - 1 ? 1 : 0) + sb_);
      }
    }
    if (active) {
      mbar_wait(&sm.d1_full_bar, 0);
      if (threadIdx.x == kPoseRoleThreads) clk_stamp(2, sc_++);
      tc_fence_after();
      float d2_unscale = 1.f;                        // undoes the operand scaling of GEMM 2
      {
        // ---- re-stage relu(D1 + b1) as FP16 (hi, lo) planes: this thread owns 16 columns of each 64-column chunk.
        // pass A: row maximum (power-of-two row scale, peak in [2^13, 2^14))
        float rmax = 0.f;
#pragma unroll
        for (int kp = 0; kp < 2; ++kp) {
          uint32_t v0[16], v1[16];
          tmem_ld16x2(d1 + lane_addr + (uint32_t)((2 * kp) * 64 + cs * 16), d1 + lane_addr + (uint32_t)((2 * kp + 1) * 64 + cs * 16), v0, v1);
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const float4 ba = *reinterpret_cast<const float4*>(sm.b1s + (2 * kp) * 64 + cs * 16 + j4 * 4);
            const float4 bb4 = *reinterpret_cast<const float4*>(sm.b1s + (2 * kp + 1) * 64 + cs * 16 + j4 * 4);
            rmax = fmaxf(rmax, fmaxf(fmaxf(__uint_as_float(v0[j4 * 4 + 0]) + ba.x, __uint_as_float(v0[j4 * 4 + 1]) + ba.y),
                                     fmaxf(__uint_as_float(v0[j4 * 4 + 2]) + ba.z, __uint_as_float(v0[j4 * 4 + 3]) + ba.w)));
            rmax = fmaxf(rmax, fmaxf(fmaxf(__uint_as_float(v1[j4 * 4 + 0]) + bb4.x, __uint_as_float(v1[j4 * 4 + 1]) + bb4.y),
                                     fmaxf(__uint_as_float(v1[j4 * 4 + 2]) + bb4.z, __uint_as_float(v1[j4 * 4 + 3]) + bb4.w)));
          }
        }
        // second-layer bias into shared memory for the epilogue (the third X chunk is free once D1 is complete; ordered by
        // the row-maximum barrier below)
        float* b2s = reinterpret_cast<float*>(sm.a + 2 * 2 * kTcABytes + 4 * kTcBM * sizeof(float));
        if (threadIdx.x - kPoseRoleThreads < 64)
          *reinterpret_cast<float4*>(b2s + (threadIdx.x - kPoseRoleThreads) * 4) =
              __ldg(reinterpret_cast<const float4*>(dn.b2) + (threadIdx.x - kPoseRoleThreads));
        float* rowmax = reinterpret_cast<float*>(sm.a + 2 * 2 * kTcABytes);      // third X chunk: free once D1 is complete
        rowmax[cs * kTcBM + r] = rmax;
        if (threadIdx.x == kPoseRoleThreads) clk_stamp(2, 1105 + sb_);
        asm volatile("bar.sync 1, 512;" ::: "memory");
        if (threadIdx.x == kPoseRoleThreads) clk_stamp(2, 1106 + sb_);
        rmax = fmaxf(fmaxf(rowmax[r], rowmax[kTcBM + r]), fmaxf(rowmax[2 * kTcBM + r], rowmax[3 * kTcBM + r]));
        float sc, inv;
        row_scale(rmax, sc, inv);
        d2_unscale = inv * dn.W2scale_inv;
        // pass B: the split, chunk by chunk
#pragma unroll
        for (int kc = 0; kc < 4; ++kc) {
          const int ab = kc & 1;
          uint32_t v[16];
          tmem_ld16(d1 + lane_addr + (uint32_t)(kc * 64 + cs * 16), v);
          float hv[16];
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const float4 bv = *reinterpret_cast<const float4*>(sm.b1s + kc * 64 + cs * 16 + j4 * 4);
            hv[j4 * 4 + 0] = fmaxf(__uint_as_float(v[j4 * 4 + 0]) + bv.x, 0.f);
            hv[j4 * 4 + 1] = fmaxf(__uint_as_float(v[j4 * 4 + 1]) + bv.y, 0.f);
            hv[j4 * 4 + 2] = fmaxf(__uint_as_float(v[j4 * 4 + 2]) + bv.z, 0.f);
            hv[j4 * 4 + 3] = fmaxf(__uint_as_float(v[j4 * 4 + 3]) + bv.w, 0.f);
          }
          uint4 hi[2], lo[2];
          split8_f16(hv, sc, hi[0], lo[0]);
          split8_f16(hv + 8, sc, hi[1], lo[1]);
          mbar_wait(&sm.a_empty_bar[ab], (uint32_t)(((kc >> 1) & 1) ^ 1));
          unsigned char* chunk = sm.a + (size_t)ab * 2 * kTcABytes;
#pragma unroll
          for (int u2 = 0; u2 < 2; ++u2) {
            const int u = cs * 2 + u2;                     // 16-byte unit (8 halves) inside the 128-byte row
            const uint32_t off = (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + (((u ^ (r & 7)) & 7) << 4));
            *reinterpret_cast<uint4*>(chunk + off) = hi[u2];
            *reinterpret_cast<uint4*>(chunk + kTcABytes + off) = lo[u2];
          }
          fence_proxy_async();
          mbar_arrive(&sm.a_full_bar[ab]);
          if (threadIdx.x == kPoseRoleThreads) clk_stamp(2, sc_++);
        }
      }
      // ---- epilogue: relu(D2 + b2) -> operand planes of the head GEMM, 64 columns per thread, two passes over TMEM
      mbar_wait(&sm.d2_full_bar, 0);
      if (threadIdx.x == kPoseRoleThreads) clk_stamp(2, sc_++);
      tc_fence_after();
      const int c0 = cs * 64;
      const float* b2s = reinterpret_cast<const float*>(sm.a + 2 * 2 * kTcABytes + 4 * kTcBM * sizeof(float));
      float rmax = 0.f;
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        uint32_t v0[16], v1[16];
        tmem_ld16x2(d2 + lane_addr + (uint32_t)(c0 + 32 * g), d2 + lane_addr + (uint32_t)(c0 + 32 * g + 16), v0, v1);
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          const float4 ba = *reinterpret_cast<const float4*>(b2s + c0 + 32 * g + j4 * 4);
          const float4 bb4 = *reinterpret_cast<const float4*>(b2s + c0 + 32 * g + 16 + j4 * 4);
          rmax = fmaxf(rmax, fmaxf(fmaxf(fmaf(__uint_as_float(v0[j4 * 4 + 0]), d2_unscale, ba.x), fmaf(__uint_as_float(v0[j4 * 4 + 1]), d2_unscale, ba.y)),
                                   fmaxf(fmaf(__uint_as_float(v0[j4 * 4 + 2]), d2_unscale, ba.z), fmaf(__uint_as_float(v0[j4 * 4 + 3]), d2_unscale, ba.w))));
          rmax = fmaxf(rmax, fmaxf(fmaxf(fmaf(__uint_as_float(v1[j4 * 4 + 0]), d2_unscale, bb4.x), fmaf(__uint_as_float(v1[j4 * 4 + 1]), d2_unscale, bb4.y)),
                                   fmaxf(fmaf(__uint_as_float(v1[j4 * 4 + 2]), d2_unscale, bb4.z), fmaf(__uint_as_float(v1[j4 * 4 + 3]), d2_unscale, bb4.w))));
        }
      }
      if (threadIdx.x == kPoseRoleThreads) clk_stamp(2, 1107 + sb_);
      {
        // FP16 planes: the row maximum (over the four column sub-blocks, through shared memory) fixes an exact power-of-two
        // scale so the row peaks in [2^13, 2^14); then (hi, lo) halves
        float* rowmax = reinterpret_cast<float*>(sm.a);      // the A-operand region is free once D2 is complete
        rowmax[cs * kTcBM + r] = rmax;
        if (threadIdx.x == kPoseRoleThreads) clk_stamp(2, 1108 + sb_);
        asm volatile("bar.sync 1, 512;" ::: "memory");
        if (threadIdx.x == kPoseRoleThreads) clk_stamp(2, 1109 + sb_);
        rmax = fmaxf(fmaxf(rowmax[r], rowmax[kTcBM + r]), fmaxf(rowmax[2 * kTcBM + r], rowmax[3 * kTcBM + r]));
        float sc, inv;
        row_scale(rmax, sc, inv);
        VPHO_BOUNDS(r0 + kTcBM <= ws.Npad);
        if (cs == 0) ws.P2scale[r0 + r] = inv;
        // The tile is a contiguous 64 KB block per plane in global memory.  Lanes own rows, so direct stores would touch 32
        // different lines per instruction (partial sectors): stage through shared memory (rows padded to 528 bytes:
        // conflict-free 16-byte stores) and write it out with consecutive threads on consecutive 16-byte units.
        constexpr int kRowPad = 528, kPlane = kTcBM * kRowPad;
        // (operand regions a + b are free now; the staging tile starts past the second-layer bias in the third X chunk,
        // which pass B below still reads)
        unsigned char* ot = sm.a + 72 * 1024;
        static_assert(72 * 1024 >= 2 * 2 * kTcABytes + 4 * kTcBM * 4 + kPDim * 4 && 72 * 1024 + 2 * kTcBM * 528 <= kPtRegionA + 2 * kPtStageB,
                      "staging tile placement");
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint32_t v[16];
          tmem_ld16(d2 + lane_addr + (uint32_t)(c0 + 16 * g), v);
          float pv[16];
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const float4 bv = *reinterpret_cast<const float4*>(b2s + c0 + 16 * g + j4 * 4);
            pv[j4 * 4 + 0] = fmaxf(fmaf(__uint_as_float(v[j4 * 4 + 0]), d2_unscale, bv.x), 0.f);
            pv[j4 * 4 + 1] = fmaxf(fmaf(__uint_as_float(v[j4 * 4 + 1]), d2_unscale, bv.y), 0.f);
            pv[j4 * 4 + 2] = fmaxf(fmaf(__uint_as_float(v[j4 * 4 + 2]), d2_unscale, bv.z), 0.f);
            pv[j4 * 4 + 3] = fmaxf(fmaf(__uint_as_float(v[j4 * 4 + 3]), d2_unscale, bv.w), 0.f);
          }
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            uint4 hi, lo;
            split8_f16(pv + 8 * u, sc, hi, lo);
            *reinterpret_cast<uint4*>(ot + r * kRowPad + (c0 + 16 * g + u * 8) * 2) = hi;
            *reinterpret_cast<uint4*>(ot + kPlane + r * kRowPad + (c0 + 16 * g + u * 8) * 2) = lo;
          }
        }
        if (threadIdx.x == kPoseRoleThreads) clk_stamp(2, 1110 + sb_);
        asm volatile("bar.sync 1, 512;" ::: "memory");
        if (threadIdx.x == kPoseRoleThreads) clk_stamp(2, 1111 + sb_);
        const int tc = threadIdx.x - kPoseRoleThreads;
        unsigned char* ghi = reinterpret_cast<unsigned char*>(ws.P2hi) + (size_t)r0 * kPDim * 2;
        unsigned char* glo = reinterpret_cast<unsigned char*>(ws.P2lo) + (size_t)r0 * kPDim * 2;
#pragma unroll 4
        for (int i = tc; i < 2 * kTcBM * 32; i += 512) {
          const int plane = i >> 12, j = i & 4095, row = j >> 5, unit = j & 31;
          VPHO_BOUNDS(plane < 2 && r0 + row < ws.Npad && unit * 16 + 16 <= kPDim * 2);
          const uint4 val = *reinterpret_cast<const uint4*>(ot + plane * kPlane + row * kRowPad + unit * 16);
          *reinterpret_cast<uint4*>((plane ? glo : ghi) + (size_t)row * 512 + unit * 16) = val;
        }
      }
    }
  }
  if (threadIdx.x == kPoseRoleThreads) clk_stamp(2, sc_++);
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) clk_stamp(0, sc_++);
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
  }
}

// One CTA per 128-row tile; with two samplers in lock-step the grid holds job 0's tiles followed by job 1's.  The jobs'
// parameters are picked by address (grid constants), so the tile body exists once in the binary.
__global__ void __launch_bounds__(kPoseThreads, 1)
k_pose_tc(const __grid_constant__ CUtensorMap tmW1_hi, const __grid_constant__ CUtensorMap tmW1_lo,
          const __grid_constant__ CUtensorMap tmW2_hi, const __grid_constant__ CUtensorMap tmW2_lo,
          const __grid_constant__ CUtensorMap tmW1_hi1, const __grid_constant__ CUtensorMap tmW1_lo1,
          const __grid_constant__ CUtensorMap tmW2_hi1, const __grid_constant__ CUtensorMap tmW2_lo1,
          const __grid_constant__ DenoiserDev dn0, const __grid_constant__ SamplerWs ws0, const __grid_constant__ DenoiserDev dn1,
          const __grid_constant__ SamplerWs ws1, int tiles0, int mode, int s) {
  const bool j1 = (int)blockIdx.x >= tiles0;
  pose_tc_tile(j1 ? &tmW1_hi1 : &tmW1_hi, j1 ? &tmW1_lo1 : &tmW1_lo, j1 ? &tmW2_hi1 : &tmW2_hi, j1 ? &tmW2_lo1 : &tmW2_lo,
               j1 ? dn1 : dn0, j1 ? ws1 : ws0, mode, s, j1 ? (int)blockIdx.x - tiles0 : (int)blockIdx.x);
}

// =====================================================================================================================
// Measured peaks for the rooflines bench.py reports ("TF32/FP32-SIMT peaks to be measured the same way on the
// box"): back-to-back FP32 FMAs on every SM, and back-to-back kind::f16 UMMAs (M128 N256 K16, operands resident in shared
// memory, FP32 accumulate in TMEM) on every SM -- the rates the contact scans and the 3xFP16 score network are bounded by at
// the clock a short kernel actually runs at.
// =====================================================================================================================
__global__ void __launch_bounds__(1024) k_peak_fp32(float* sink, int iters) {
  float a[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = (float)(threadIdx.x + j) * 1e-3f;
  const float b = 1.0000001f, c = 1e-9f;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = fmaf(a[j], b, c);
  }
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += a[j];
  if (s == 123.456f) sink[0] = s;        // never true: keeps the chain alive
}

constexpr int kPeakSmem = 120 * 1024;    // more than half an SM's shared memory: one CTA per SM
__global__ void __launch_bounds__(128, 1) k_peak_umma_f16(int n_groups) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* a = base;                       // [128 rows][64 half]  16 KB, K-major SWIZZLE_128B
  unsigned char* b = base + kTcABytes;           // [256 rows][64 half]  32 KB
  __shared__ unsigned long long bar;
  __shared__ uint32_t tmem_slot;
  for (int i = threadIdx.x; i < (kTcABytes + kTcBBytes) / 16; i += blockDim.x) reinterpret_cast<uint4*>(base)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 1 && lane == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (warp == 1 && lane == 0) {
    const uint64_t ad = make_kmajor_sw128_desc(a), bd = make_kmajor_sw128_desc(b);
    for (int g = 0; g < n_groups; ++g) {
      const uint32_t d = tmem_base + (uint32_t)((g & 1) * kTcBN);
#pragma unroll
      for (int k = 0; k < 4; ++k) {               // 4 x K16 inside the 128-byte rows
        const uint64_t adv = (uint64_t)((k * 32) >> 4);
        umma_f16(d, ad + adv, bd + adv, kTcIdescF16, (g > 1 || k > 0) ? 1u : 0u);
      }
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
  }
}

// -> TFLOP/s of each (best of `reps` timed launches after a warm-up one), measured with CUDA events on `st`
int tc_measure_peaks(float* fp32_tflops, float* f16_tflops, int reps, cudaStream_t st) {
  int dev = 0, n_sm = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return VPHO_ERR_LAUNCH;
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  if (n_sm <= 0) n_sm = 148;
  if (cudaFuncSetAttribute(k_peak_umma_f16, cudaFuncAttributeMaxDynamicSharedMemorySize, kPeakSmem) != cudaSuccess) return VPHO_ERR_LAUNCH;
  float* sink = nullptr;
  if (cudaMalloc((void**)&sink, 256) != cudaSuccess) return VPHO_ERR_ALLOC;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 1 << 15, n_groups = 1 << 13;
  float best32 = 1e30f, best16 = 1e30f;
  for (int r = 0; r <= reps; ++r) {
    float ms = 0.f;
    cudaEventRecord(e0, st);
    k_peak_fp32<<<2 * n_sm, 1024, 0, st>>>(sink, iters);
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    if (r > 0 && ms < best32) best32 = ms;
    cudaEventRecord(e0, st);
    k_peak_umma_f16<<<n_sm, 128, kPeakSmem, st>>>(n_groups);
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    if (r > 0 && ms < best16) best16 = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(sink);
  if (cudaGetLastError() != cudaSuccess) return VPHO_ERR_LAUNCH;
  *fp32_tflops = (float)((double)2 * n_sm * 1024.0 * iters * 8 * 2 / (best32 * 1e-3) / 1e12);
  *f16_tflops = (float)((double)n_sm * n_groups * 4.0 * (2.0 * kTcBM * kTcBN * 16) / (best16 * 1e-3) / 1e12);
  return VPHO_OK;
}

// -------------------------------------------------------------------------------------------------- host side
static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }
  return fn;
}

// 2-D K-major f32 tensor [rows][kdim] -> boxes of {32 k, box_rows}, 128-byte swizzle
bool tc_make_map(void* map_out, const void* base, int rows, int box_rows, int kdim, bool half) {
  CUtensorMap* map = static_cast<CUtensorMap*>(map_out);
  PFN_cuTensorMapEncodeTiled_v12000 enc = get_encode();
  if (!enc) return false;
  const int esz = half ? 2 : 4;
  cuuint64_t gdim[2] = {(cuuint64_t)kdim, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)kdim * esz};
  cuuint32_t box[2] = {(cuuint32_t)(128 / esz), (cuuint32_t)box_rows};      // 128-byte rows
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, half ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

bool tc_available() { return get_encode() != nullptr; }

int tc_debug_clocks(int enable, unsigned long long* out, int n) {
  int on = enable;
  if (cudaMemcpyToSymbol(g_clk_on, &on, sizeof(int)) != cudaSuccess) return VPHO_ERR_LAUNCH;
  if (out && n > 0) {
    if (n > 3 * 2048) n = 3 * 2048;
    if (cudaMemcpyFromSymbol(out, g_clk, (size_t)n * sizeof(unsigned long long)) != cudaSuccess) return VPHO_ERR_LAUNCH;
  }
  return VPHO_OK;
}

// ctas = 2 launches the CTA-pair variant: mapB_* must then have 128-row boxes.  n_jobs = 2 serves two samplers in
// lock-step with one launch (see k_head_tc).  Per-device state (function attributes, SM count) is looked up per device.
int tc_launch_head(const TcHeadJob* jobs, int n_jobs, int mode, int s, int ctas, cudaStream_t st) {
  const int smem = (int)sizeof(TcSmem) + 1024;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return VPHO_ERR_LAUNCH;
  static bool attr[kMaxDevices] = {};
  static int n_sm_of[kMaxDevices] = {};
  if (!attr[dev]) {
    if (cudaFuncSetAttribute(k_head_tc<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess ||
        cudaFuncSetAttribute(k_head_tc<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
      return VPHO_ERR_LAUNCH;
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    n_sm_of[dev] = n > 0 ? n : 148;
    attr[dev] = true;
  }
  const int n_sm = n_sm_of[dev];
  if (n_jobs < 1 || n_jobs > 2 || (ctas != 1 && ctas != 2)) return VPHO_ERR_INVALID;
  const TcHeadJob& j0 = jobs[0];
  const TcHeadJob& j1 = jobs[n_jobs - 1];
  int n_items = 0;
  for (int j = 0; j < n_jobs; ++j) {
    const int n_tiles = jobs[j].ws->Npad / kTcBM;
    n_items += (ctas == 2 ? (n_tiles + 1) / 2 : n_tiles) * jobs[j].dn->n_heads;
  }
  // the fewest CTAs (CTA pairs) that keep the same number of rounds: SMs left over go to other streams
  const int units_max = n_sm / ctas;
  const int rounds = (n_items + units_max - 1) / units_max;
  const int units = (n_items + rounds - 1) / rounds;       // e.g. 150 items: 75 CTAs x 2 rather than 148 CTAs, 2 of them x 2
  auto M = [](const void* p) -> const CUtensorMap& { return *static_cast<const CUtensorMap*>(p); };
  cudaError_t e;
  if (ctas == 2)
    e = launch_pdl(k_head_tc<2>, dim3(2 * units), dim3(kHeadThreads), smem, st, 2, M(j0.mapA_hi), M(j0.mapA_lo), M(j0.mapB_hi),
                   M(j0.mapB_lo), M(j1.mapA_hi), M(j1.mapA_lo), M(j1.mapB_hi), M(j1.mapB_lo), *j0.dn, *j0.ws, *j1.dn, *j1.ws, n_jobs,
                   mode, s);
  else
    e = launch_pdl(k_head_tc<1>, dim3(units), dim3(kHeadThreads), smem, st, 1, M(j0.mapA_hi), M(j0.mapA_lo), M(j0.mapB_hi),
                   M(j0.mapB_lo), M(j1.mapA_hi), M(j1.mapA_lo), M(j1.mapB_hi), M(j1.mapB_lo), *j0.dn, *j0.ws, *j1.dn, *j1.ws, n_jobs,
                   mode, s);
  return e == cudaSuccess ? VPHO_OK : VPHO_ERR_LAUNCH;
}

int tc_launch_pose(const TcPoseJob* jobs, int n_jobs, int mode, int s, cudaStream_t st) {
  const int smem = (int)sizeof(PtSmem) + 1024;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return VPHO_ERR_LAUNCH;
  static bool attr[kMaxDevices] = {};
  if (!attr[dev]) {
    if (cudaFuncSetAttribute(k_pose_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return VPHO_ERR_LAUNCH;
    attr[dev] = true;
  }
  if (n_jobs < 1 || n_jobs > 2) return VPHO_ERR_INVALID;
  const TcPoseJob& j0 = jobs[0];
  const TcPoseJob& j1 = jobs[n_jobs - 1];
  const int tiles0 = j0.ws->Npad / kTcBM, tiles1 = n_jobs > 1 ? j1.ws->Npad / kTcBM : 0;
  auto M = [](const void* p) -> const CUtensorMap& { return *static_cast<const CUtensorMap*>(p); };
  if (launch_pdl(k_pose_tc, dim3(tiles0 + tiles1), dim3(kPoseThreads), smem, st, 1, M(j0.mapW1_hi), M(j0.mapW1_lo), M(j0.mapW2_hi),
                 M(j0.mapW2_lo), M(j1.mapW1_hi), M(j1.mapW1_lo), M(j1.mapW2_hi), M(j1.mapW2_lo), *j0.dn, *j0.ws, *j1.dn, *j1.ws,
                 tiles0, mode, s) != cudaSuccess)
    return VPHO_ERR_LAUNCH;
  return VPHO_OK;
}

}  // namespace vpho
