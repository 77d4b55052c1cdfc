// common.cuh -- shared host/device helpers of libvadx (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <mutex>

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

// One-time PER-DEVICE kernel setup: cudaFuncSetAttribute (the > 48 KB dynamic shared-memory opt-in) and the SM count
// belong to a device, not to the process, and ctypes callers release the GIL -- so the "already configured" state is
// keyed by cudaGetDevice() and guarded by a mutex.
//   static PerDevice pd;  int n_sm;  VADX_TRY(pd.ensure(&n_sm, [&] { return cudaFuncSetAttribute(...); }));
struct PerDevice {
  static constexpr int kMaxDevices = 64;
  std::mutex mu;
  bool done[kMaxDevices] = {};
  int sms[kMaxDevices] = {};
  template <class F>
  int ensure(int* n_sm, F&& configure) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
    if (dev < 0 || dev >= kMaxDevices) {
      set_error("device ordinal %d is outside the supported range [0, %d)", dev, kMaxDevices);
      return VADX_EINVAL;
    }
    std::lock_guard<std::mutex> lk(mu);
    if (!done[dev]) {
      int n = 0;
      cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
      e = configure();
      if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute");
      sms[dev] = n > 0 ? n : 148;
      done[dev] = true;
    }
    if (n_sm) *n_sm = sms[dev];
    return VADX_OK;
  }
};

// A/B switches of the measurements quoted in DESIGN.md: compiled in only with -DVADX_AB_SWITCHES (make AB=1); a
// production build has no environment reads on any path.
#ifdef VADX_AB_SWITCHES
inline int ab_env(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}
#else
inline int ab_env(const char*, int dflt) { return dflt; }
#endif

// per-stage / per-kernel event timing (api.cu); a no-op unless vadx_profile_enable(1) was called.  `kernel` names the
// entry point's kernel, `bytes` / `flops` are the ALGORITHMIC work of this call (the tensors it must read and write once,
// 2 x MACs of the contraction it computes) -- what bench.py divides by the measured time for the roofline lines.
struct StageTimer {
  int stage;
  cudaStream_t st;
  int slot;
  StageTimer(int stage_, cudaStream_t st_, const char* kernel = nullptr, double bytes = 0.0, double flops = 0.0);
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
