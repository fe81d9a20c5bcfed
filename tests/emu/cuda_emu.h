// TEST INFRASTRUCTURE -- a tiny single-threaded SIMT emulator so that the SAME kernel sources under
// vpho_b200/csrc/ (everything except the tcgen05/TMA kernels) can be executed on a CPU-only machine at toy
// sizes by tests/ (`-m "not gpu"`).  It is never built into, loaded by, or reachable from the product
// library: vpho_b200/csrc/build.py only ever invokes nvcc for sm_100a, and vpho_b200.capi refuses to load
// anything when no CUDA device is present.  The emulator runs one thread block at a time; every CUDA thread
// is a ucontext fiber, `__syncthreads()` / warp shuffles are cooperative yields.
#pragma once
#ifndef VPHO_EMU
#error "cuda_emu.h is only for -DVPHO_EMU host builds"
#endif

#include <ucontext.h>
#include <sys/mman.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))

struct uint3 { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct float2 { float x, y; };
struct float3 { float x, y, z; };
struct __attribute__((aligned(16))) float4 { float x, y, z, w; };
struct __attribute__((aligned(16))) double2 { double x, y; };
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
struct int2 { int x, y; };
struct __attribute__((aligned(16))) int4 { int x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline float3 make_float3(float x, float y, float z) { return float3{x, y, z}; }
static inline float2 make_float2(float x, float y) { return float2{x, y}; }

typedef int cudaError_t;
typedef void* cudaStream_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorLaunchFailure = 719 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemset(void* p, int v, size_t n) { memset(p, v, n); return cudaSuccess; }
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = malloc(n); return *p ? cudaSuccess : cudaErrorInvalidValue; }
static inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
template <typename T> static inline cudaError_t cudaFuncSetAttribute(T, int, int) { return cudaSuccess; }
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };

namespace emu {

struct Fiber {
  ucontext_t ctx;
  void* stack = nullptr;
  bool done = false;
  uint3 tid{0, 0, 0};
  unsigned linear = 0;
};

struct State {
  std::vector<Fiber> fibers;
  ucontext_t sched;
  Fiber* cur = nullptr;
  dim3 grid, block;
  uint3 bid{0, 0, 0};
  unsigned nthreads = 0;
  // block barrier
  unsigned bar_arrived = 0, bar_gen = 0, exited = 0;
  // per-warp exchange (two generations to separate write/read phases)
  struct Warp { unsigned arrived = 0, gen = 0, size = 0, exited = 0; uint64_t buf[2][32]; };
  std::vector<Warp> warps;
  const std::function<void()>* body = nullptr;
  unsigned char* dyn_smem = nullptr;
  size_t dyn_smem_cap = 0;
};
inline State& S() { static State s; return s; }
static constexpr size_t kStack = 256 * 1024;

inline void yield() { State& s = S(); swapcontext(&s.cur->ctx, &s.sched); }

inline void fiber_entry() {
  State& s = S();
  (*s.body)();
  Fiber* f = s.cur;
  f->done = true;
  s.exited++;
  s.warps[f->linear / 32].exited++;
  swapcontext(&f->ctx, &s.sched);
}

inline void syncthreads() {
  State& s = S();
  unsigned gen = s.bar_gen;
  s.bar_arrived++;
  while (true) {
    if (s.bar_gen != gen) return;
    if (s.bar_arrived + s.exited >= s.nthreads) { s.bar_arrived = 0; s.bar_gen++; return; }
    yield();
  }
}

// all live lanes of the calling warp exchange one 64-bit word; returns pointer to the 32-slot snapshot
inline const uint64_t* warp_exchange(uint64_t v) {
  State& s = S();
  Fiber* f = s.cur;
  State::Warp& w = s.warps[f->linear / 32];
  unsigned gen = w.gen;
  w.buf[gen & 1][f->linear % 32] = v;
  w.arrived++;
  while (true) {
    if (w.gen != gen) break;
    if (w.arrived + w.exited >= w.size) { w.arrived = 0; w.gen++; break; }
    yield();
  }
  return w.buf[gen & 1];
}

inline void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
  State& s = S();
  s.grid = grid; s.block = block; s.body = &body;
  s.nthreads = block.x * block.y * block.z;
  if (s.fibers.size() < s.nthreads) {
    size_t old = s.fibers.size();
    s.fibers.resize(s.nthreads);
    for (size_t i = old; i < s.nthreads; ++i) {
      s.fibers[i].stack = mmap(nullptr, kStack, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
      if (s.fibers[i].stack == MAP_FAILED) { fprintf(stderr, "emu: mmap failed\n"); abort(); }
    }
  }
  if (smem > s.dyn_smem_cap) {
    free(s.dyn_smem);
    s.dyn_smem = (unsigned char*)aligned_alloc(1024, (smem + 1023) / 1024 * 1024);
    s.dyn_smem_cap = smem;
  }
  unsigned nwarps = (s.nthreads + 31) / 32;
  s.warps.assign(nwarps, State::Warp());
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        s.bid = uint3{bx, by, bz};
        s.bar_arrived = 0; s.exited = 0;
        for (unsigned w = 0; w < nwarps; ++w) {
          s.warps[w].arrived = 0; s.warps[w].exited = 0;
          s.warps[w].size = std::min(32u, s.nthreads - w * 32);
        }
        for (unsigned t = 0; t < s.nthreads; ++t) {
          Fiber& f = s.fibers[t];
          f.done = false; f.linear = t;
          f.tid = uint3{t % block.x, (t / block.x) % block.y, t / (block.x * block.y)};
          getcontext(&f.ctx);
          f.ctx.uc_stack.ss_sp = f.stack; f.ctx.uc_stack.ss_size = kStack; f.ctx.uc_link = nullptr;
          makecontext(&f.ctx, (void (*)())fiber_entry, 0);
        }
        unsigned live = s.nthreads;
        while (live) {
          live = 0;
          for (unsigned t = 0; t < s.nthreads; ++t) {
            Fiber& f = s.fibers[t];
            if (f.done) continue;
            s.cur = &f;
            swapcontext(&s.sched, &f.ctx);
            if (!f.done) live++;
          }
        }
      }
  s.cur = nullptr;
}

}  // namespace emu

#define threadIdx (emu::S().cur->tid)
#define blockIdx (emu::S().bid)
#define blockDim (emu::S().block)
#define gridDim (emu::S().grid)
static inline void __syncthreads() { emu::syncthreads(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::warp_exchange(0); }
static inline void __threadfence() {}
static inline void __threadfence_block() {}

template <typename T> static inline uint64_t emu_pack(T v) { uint64_t u = 0; memcpy(&u, &v, sizeof(T)); return u; }
template <typename T> static inline T emu_unpack(uint64_t u) { T v; memcpy(&v, &u, sizeof(T)); return v; }
template <typename T> static inline T __shfl_sync(unsigned, T v, int src, int width = 32) {
  unsigned lane = emu::S().cur->linear % 32;
  const uint64_t* b = emu::warp_exchange(emu_pack(v));
  int base = (lane / width) * width;
  return emu_unpack<T>(b[base + (src % width)]);
}
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int m, int width = 32) {
  unsigned lane = emu::S().cur->linear % 32;
  const uint64_t* b = emu::warp_exchange(emu_pack(v));
  unsigned src = lane ^ (unsigned)m;
  if (src / width != lane / width) src = lane;
  return emu_unpack<T>(b[src]);
}
template <typename T> static inline T __shfl_down_sync(unsigned, T v, unsigned d, int width = 32) {
  unsigned lane = emu::S().cur->linear % 32;
  const uint64_t* b = emu::warp_exchange(emu_pack(v));
  unsigned src = lane + d;
  if (src / width != lane / width) src = lane;
  return emu_unpack<T>(b[src]);
}
template <typename T> static inline T __shfl_up_sync(unsigned, T v, unsigned d, int width = 32) {
  unsigned lane = emu::S().cur->linear % 32;
  const uint64_t* b = emu::warp_exchange(emu_pack(v));
  int src = (int)lane - (int)d;
  if (src < 0 || (unsigned)src / width != lane / width) src = lane;
  return emu_unpack<T>(b[src]);
}
static inline unsigned __ballot_sync(unsigned, int pred) {
  emu::State& s = emu::S();
  unsigned wsz = s.warps[s.cur->linear / 32].size;
  const uint64_t* b = emu::warp_exchange(pred ? 1 : 0);
  unsigned m = 0;
  for (unsigned i = 0; i < wsz; ++i) if (b[i]) m |= 1u << i;
  return m;
}
static inline int __any_sync(unsigned m, int p) { return __ballot_sync(m, p) != 0; }
static inline int __all_sync(unsigned m, int p) {
  emu::State& s = emu::S();
  unsigned wsz = s.warps[s.cur->linear / 32].size;
  unsigned full = wsz == 32 ? 0xffffffffu : ((1u << wsz) - 1);
  return (__ballot_sync(m, p) & full) == full;
}

// atomics: single OS thread => plain read-modify-write
template <typename T> static inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
template <typename T> static inline T atomicMax(T* p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <typename T> static inline T atomicMin(T* p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <typename T> static inline T atomicExch(T* p, T v) { T o = *p; *p = v; return o; }
template <typename T> static inline T atomicOr(T* p, T v) { T o = *p; *p = o | v; return o; }
template <typename T> static inline T atomicCAS(T* p, T c, T v) { T o = *p; if (o == c) *p = v; return o; }
static inline unsigned atomicInc(unsigned* p, unsigned lim) { unsigned o = *p; *p = (o >= lim) ? 0 : o + 1; return o; }

template <typename T> static inline T __ldg(const T* p) { return *p; }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
static inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline float __frcp_rn(float a) { return 1.0f / a; }
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
static inline double __fma_rn(double a, double b, double c) { return fma(a, b, c); }
static inline float rsqrtf(float a) { return 1.0f / sqrtf(a); }
static inline double rsqrt(double a) { return 1.0 / sqrt(a); }
static inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
static inline unsigned __float_as_uint(float f) { unsigned i; memcpy(&i, &f, 4); return i; }
static inline float __uint_as_float(unsigned i) { float f; memcpy(&f, &i, 4); return f; }
static inline long long __double_as_longlong(double d) { long long i; memcpy(&i, &d, 8); return i; }
static inline double __longlong_as_double(long long i) { double d; memcpy(&d, &i, 8); return d; }
static inline int __float2int_rd(float f) { return (int)floorf(f); }
static inline int __float2int_rn(float f) { return (int)nearbyintf(f); }
using std::isnan;
using std::isinf;
using std::isfinite;
using std::max;
using std::min;

#define VPHO_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(emu::S().dyn_smem)
#define VPHO_LAUNCH(kern, grid, block, smem, stream, ...) \
  emu::launch(dim3(grid), dim3(block), (size_t)(smem), [&]() { kern(__VA_ARGS__); })
