// Library-wide bookkeeping: launch counter and optional CUDA-event brackets around tagged kernels (bench.py uses them
// to time the dominant kernel live, on the launching stream, inside its timed region).
#include "vpho_common.cuh"
#include "vpho_b200.h"

#include <utility>
#include <vector>

namespace vpho {

unsigned long long g_launches = 0;
static unsigned g_profile = 0;   // bit t set: bracket kernels tagged t with CUDA events
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_events[VPHO_NUM_TAGS];
static cudaEvent_t g_open[VPHO_NUM_TAGS];
static std::vector<cudaEvent_t> g_pool;     // events are recycled: creating one per launch costs more than the launch

static cudaEvent_t take_event() {
  if (g_pool.empty()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
  }
  cudaEvent_t e = g_pool.back();
  g_pool.pop_back();
  return e;
}

void profile_begin(int tag, cudaStream_t st) {
  if (!(g_profile >> tag & 1u)) return;
  cudaEvent_t e = take_event();
  cudaEventRecord(e, st);
  g_open[tag] = e;
}

void profile_end(int tag, cudaStream_t st) {
  if (!(g_profile >> tag & 1u)) return;
  cudaEvent_t e = take_event();
  cudaEventRecord(e, st);
  g_events[tag].push_back({g_open[tag], e});
}

}  // namespace vpho

using namespace vpho;

namespace vpho { int tc_measure_peaks(float* fp32_tflops, float* f16_tflops, int reps, cudaStream_t st); }

extern "C" int vpho_version(void) { return 200; }

extern "C" int vpho_measure_peaks(float* fp32_fma_tflops, float* f16_umma_tflops, int reps, void* stream) {
  if (!fp32_fma_tflops || !f16_umma_tflops || reps < 1) return VPHO_ERR_INVALID;
  return vpho::tc_measure_peaks(fp32_fma_tflops, f16_umma_tflops, reps, (cudaStream_t)stream);
}

extern "C" int vpho_set_pdl(int enabled) {
  vpho::pdl_override() = enabled ? 1 : 0;
  return VPHO_OK;
}

extern "C" unsigned long long vpho_launch_count(void) { return g_launches; }

extern "C" int vpho_profile_enable(int tag_mask) {
  g_profile = (unsigned)tag_mask;
  return VPHO_OK;
}

extern "C" int vpho_profile_reserve(int n_events) {
  // creating CUDA events lazily inside a timed region occasionally stalls the driver for tens of ms: make them up front
  while ((int)g_pool.size() < n_events) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return VPHO_ERR_ALLOC;
    g_pool.push_back(e);
  }
  return VPHO_OK;
}

extern "C" int vpho_profile_collect_list(int tag, double* total_ms, int* n_launches, float* each_ms, int cap) {
  if (tag < 0 || tag >= VPHO_NUM_TAGS || !total_ms || !n_launches) return VPHO_ERR_INVALID;
  double tot = 0.0;
  int n = 0;
  for (auto& pr : g_events[tag]) {
    if (cudaEventSynchronize(pr.second) != cudaSuccess) return VPHO_ERR_LAUNCH;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, pr.first, pr.second);
    tot += ms;
    if (each_ms && n < cap) each_ms[n] = ms;
    ++n;
    g_pool.push_back(pr.first);
    g_pool.push_back(pr.second);
  }
  g_events[tag].clear();
  *total_ms = tot;
  *n_launches = n;
  return VPHO_OK;
}

extern "C" int vpho_profile_collect(int tag, double* total_ms, int* n_launches) {
  return vpho_profile_collect_list(tag, total_ms, n_launches, nullptr, 0);
}
