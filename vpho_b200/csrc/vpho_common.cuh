// Common definitions for the vpho_b200 sm_100a kernels.
// The same sources compile under -DVPHO_EMU against tests/emu/cuda_emu.h (test-only SIMT emulator).
#pragma once
#include <cstdlib>

// -DVPHO_DEBUG_BOUNDS build (libvpho_b200_bounds.so, `python -m vpho_b200.build --bounds`): every global-memory index the
// tensor-core kernels form is asserted against its extent; a violation prints the site and traps (the launch then fails
// with a sticky error, which the C ABI reports).  compute-sanitizer does not run these kernels (tcgen05), hence this build.
#if defined(VPHO_DEBUG_BOUNDS) && !defined(VPHO_EMU)
#include <cstdio>
#define VPHO_BOUNDS(cond)                                                                                                   \
  do {                                                                                                                       \
    if (!(cond)) {                                                                                                           \
      printf("VPHO_BOUNDS violated: %s  (%s:%d, block %d,%d thread %d)\n", #cond, __FILE__, __LINE__, (int)blockIdx.x,    \
             (int)blockIdx.y, (int)threadIdx.x);                                                                             \
      __trap();                                                                                                              \
    }                                                                                                                        \
  } while (0)
#else
#define VPHO_BOUNDS(cond) do { } while (0)
#endif

#ifndef VPHO_EMU
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#define VPHO_DYN_SMEM(type, name)                                 \
  extern __shared__ __align__(16) unsigned char name##_raw_[];    \
  type* name = reinterpret_cast<type*>(name##_raw_)
namespace vpho {
extern unsigned long long g_launches;   // kernels launched by this library since load (vpho_launch_count)
void profile_begin(int tag, cudaStream_t st);   // CUDA-event bracket around one launch when profiling is enabled
void profile_end(int tag, cudaStream_t st);
}  // namespace vpho
#define VPHO_LAUNCH(kern, grid, block, smem, stream, ...) \
  do { ++::vpho::g_launches; kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__); } while (0)
#define VPHO_CONSTANT static __constant__
// launch with programmatic stream serialization (see launch_pdl): the kernel MUST start with pdl_wait()
#define VPHO_LAUNCH_PDL(kern, grid, block, smem, stream, ...) \
  do { if (::vpho::launch_pdl(kern, (grid), (block), (smem), (stream), 1, __VA_ARGS__) != cudaSuccess) return VPHO_ERR_LAUNCH; } while (0)
namespace vpho {
// Programmatic dependent launch: a kernel launched with launch_pdl may be scheduled while its predecessor in the stream is
// still running (after every CTA of the predecessor has executed pdl_trigger or exited).  It must execute pdl_wait() on
// every path before touching anything an earlier kernel wrote; pdl_wait returns once the predecessor grid has completed and
// its writes are visible.  VPHO_NO_PDL=1 falls back to plain stream-ordered launches.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// programmatic dependent launch is on by default; vpho_set_pdl(0) turns it off (bench.py's serialised per-kernel pass)
inline int& pdl_override() {
  static int v = 1;
  return v;
}
inline bool pdl_enabled() { return pdl_override() == 1; }
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster_x, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  int na = 0;
  if (cluster_x > 1) {
    at[na].id = cudaLaunchAttributeClusterDimension;
    at[na].val.clusterDim.x = (unsigned)cluster_x;
    at[na].val.clusterDim.y = 1;
    at[na].val.clusterDim.z = 1;
    ++na;
  }
  if (pdl_enabled()) {
    at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = at;
  cfg.numAttrs = (unsigned)na;
  ++g_launches;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
// 16-byte asynchronous global->shared copy (LDGSTS) and its group fences
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }
}  // namespace vpho
#else
#define VPHO_CONSTANT static
#define VPHO_LAUNCH_PDL VPHO_LAUNCH
namespace vpho {
inline void pdl_wait() {}
inline void pdl_trigger() {}
inline void profile_begin(int, cudaStream_t) {}
inline void profile_end(int, cudaStream_t) {}
inline void cp_async16(void* smem, const void* gmem) { memcpy(smem, gmem, 16); }
inline void cp_async_commit() {}
template <int N>
inline void cp_async_wait() {}
}  // namespace vpho
#endif

// profiling tags (vpho_profile_collect)
#define VPHO_TAG_HEAD_GEMM_HAND 0
#define VPHO_TAG_HEAD_GEMM_OBJ 1
#define VPHO_TAG_POSE_ENCODER 2
#define VPHO_TAG_MANO_FULL 3
#define VPHO_TAG_PHYSICS3 4
#define VPHO_TAG_HAND_SCORE 5
#define VPHO_TAG_STAGE_X 6
#define VPHO_TAG_FEAT_TERM 7
#define VPHO_TAG_RK_CONTROL 8
#define VPHO_TAG_AGGREGATE 9
#define VPHO_TAG_HAND_PHYS 10
#define VPHO_TAG_POSTPROCESS 11
#define VPHO_NUM_TAGS 12

#define VPHO_OK 0
#define VPHO_ERR_INVALID (-1)
#define VPHO_ERR_LAUNCH (-2)
#define VPHO_ERR_ALLOC (-3)

#define VPHO_CHECK_LAUNCH()                                 \
  do {                                                      \
    cudaError_t e__ = cudaGetLastError();                   \
    if (e__ != cudaSuccess) return VPHO_ERR_LAUNCH;         \
  } while (0)

namespace vpho {

// ---- MANO geometry constants (manopth ManoLayer, right hand) ----
constexpr int kVerts = 778;
constexpr int kJoints16 = 16;
constexpr int kJoints21 = 21;
constexpr int kBlendK = 145;       // 10 shape + 135 pose-corrective coefficients
constexpr int kVChunk = 195;       // vertices handled by one CTA of the skinning kernel
constexpr int kVChunkPad = 224;    // packed stride of a chunk (7 warps, 128 B aligned rows)
constexpr int kNumVChunks = 4;
constexpr int kVPad = kVChunkPad * kNumVChunks;  // 896 packed vertex slots

__device__ __forceinline__ int packed_vertex(int v) { return (v / kVChunk) * kVChunkPad + (v % kVChunk); }

}  // namespace vpho
