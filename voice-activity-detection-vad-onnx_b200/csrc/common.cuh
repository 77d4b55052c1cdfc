// common.cuh -- shared host/device helpers of libvadx (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <atomic>
#include <cstring>

#include "vadx.h"

namespace vadx {

// at or below this many rows the exact-fp32 skinny kernel (gemm_simt.cu) beats a tensor-core launch
constexpr int kSkinnyMaxRows = 16;


// error plumbing (api.cu)
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
extern std::atomic<uint64_t> g_launches;

inline int after_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, what);
  return VADX_OK;
}

// per-stage event timing (api.cu); a no-op unless vadx_profile_enable(1) was called
struct StageTimer {
  int stage;
  cudaStream_t st;
  int slot;
  StageTimer(int stage_, cudaStream_t st_);
  ~StageTimer();
};

#define VADX_REQUIRE(cond, ...)       \
  do {                                \
    if (!(cond)) {                    \
      ::vadx::set_error(__VA_ARGS__); \
      return VADX_EINVAL;             \
    }                                 \
  } while (0)

#define VADX_TRY(expr)            \
  do {                            \
    int _rc = (expr);             \
    if (_rc != VADX_OK) return _rc; \
  } while (0)

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int64_t round_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }
static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
__device__ __forceinline__ bool aligned16_dev(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

__device__ __forceinline__ int64_t ceil_div_dev(int64_t a, int64_t b) { return (a + b - 1) / b; }
__device__ __forceinline__ int64_t min_i64(int64_t a, int64_t b) { return a < b ? a : b; }

__device__ __forceinline__ float apply_act(float v, int act) {
  act &= 15;
  if (act == VADX_ACT_RELU) return fmaxf(v, 0.0f);
  if (act == VADX_ACT_SIGMOID) return 1.0f / (1.0f + expf(-v));
  return v;
}

}  // namespace vadx
