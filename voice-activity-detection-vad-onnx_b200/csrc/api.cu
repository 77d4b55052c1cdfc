// api.cu -- error plumbing and library-level entry points of libvadx.
#include <cstdarg>
#include <cstdio>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace vadx {

std::atomic<uint64_t> g_launches{0};
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("%s: CUDA error %d (%s)", what, (int)e, cudaGetErrorString(e));
  if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) return VADX_ENODEVICE;
  return VADX_ECUDA;
}

// ---- per-stage profiling ----
namespace {
struct Rec {
  int stage;
  cudaEvent_t a, b;
  const char* kernel;   // string literal of the entry point
  double bytes, flops;
};
struct KernelTotal {
  const char* kernel;
  double ms, bytes, flops;
  uint64_t calls;
};
std::mutex g_prof_mu;
bool g_prof_on = false;
std::vector<Rec> g_recs;
double g_stage_ms[VADX_STAGE_COUNT] = {};
uint64_t g_stage_calls[VADX_STAGE_COUNT] = {};
std::vector<KernelTotal> g_kernels;
const char* const kStageNames[VADX_STAGE_COUNT] = {"prep", "stft", "mel", "linear", "memory", "head", "postproc"};
std::vector<cudaEvent_t> g_pool;
cudaEvent_t take_event() {
  if (!g_pool.empty()) {
    cudaEvent_t e = g_pool.back();
    g_pool.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}
}  // namespace

// synchronise the recorded event pairs and fold them into the per-stage and per-kernel totals (g_prof_mu held)
static int drain_records() {
  for (auto& r : g_recs) {
    cudaError_t e = cudaEventSynchronize(r.b);
    if (e != cudaSuccess) return cuda_fail(e, "vadx_profile_collect");
    float t = 0.f;
    e = cudaEventElapsedTime(&t, r.a, r.b);
    if (e != cudaSuccess) return cuda_fail(e, "vadx_profile_collect");
    g_stage_ms[r.stage] += t;
    g_stage_calls[r.stage] += 1;
    const char* name = r.kernel ? r.kernel : kStageNames[r.stage];
    KernelTotal* k = nullptr;
    for (auto& kt : g_kernels)
      if (kt.kernel == name || !strcmp(kt.kernel, name)) { k = &kt; break; }
    if (!k) {
      g_kernels.push_back(KernelTotal{name, 0.0, 0.0, 0.0, 0});
      k = &g_kernels.back();
    }
    k->ms += t; k->bytes += r.bytes; k->flops += r.flops; k->calls += 1;
    g_pool.push_back(r.a);
    g_pool.push_back(r.b);
  }
  g_recs.clear();
  return VADX_OK;
}

StageTimer::StageTimer(int stage_, cudaStream_t st_, const char* kernel, double bytes, double flops) : stage(stage_), st(st_), slot(-1) {
  if (!g_prof_on) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  Rec r{stage, take_event(), take_event(), kernel, bytes, flops};
  if (!r.a || !r.b) return;
  cudaEventRecord(r.a, st);
  g_recs.push_back(r);
  slot = (int)g_recs.size() - 1;
}
StageTimer::~StageTimer() {
  if (slot < 0) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (slot < (int)g_recs.size()) cudaEventRecord(g_recs[slot].b, st);
}

}  // namespace vadx

extern "C" int vadx_profile_enable(int on) {
  std::lock_guard<std::mutex> lk(vadx::g_prof_mu);
  vadx::g_prof_on = on != 0;
  return VADX_OK;
}

extern "C" int vadx_profile_collect(double* ms, uint64_t* calls, int n_stages) {
  using namespace vadx;
  VADX_REQUIRE(ms && calls && n_stages >= VADX_STAGE_COUNT, "vadx_profile_collect: need %d stage slots",
               VADX_STAGE_COUNT);
  std::lock_guard<std::mutex> lk(g_prof_mu);
  VADX_TRY(drain_records());
  for (int i = 0; i < VADX_STAGE_COUNT; ++i) {
    ms[i] += g_stage_ms[i];
    calls[i] += g_stage_calls[i];
    g_stage_ms[i] = 0.0;
    g_stage_calls[i] = 0;
  }
  return VADX_OK;
}

extern "C" int vadx_profile_collect_kernels(vadx_kernel_stat* out, int capacity, int* n_out) {
  using namespace vadx;
  VADX_REQUIRE(n_out && (out || capacity == 0) && capacity >= 0, "vadx_profile_collect_kernels: bad argument");
  std::lock_guard<std::mutex> lk(g_prof_mu);
  VADX_TRY(drain_records());
  *n_out = (int)g_kernels.size();
  if (capacity < (int)g_kernels.size()) return VADX_OK;      // query: totals are kept until a call with enough room
  for (size_t i = 0; i < g_kernels.size(); ++i) {
    out[i].name = g_kernels[i].kernel;
    out[i].ms = g_kernels[i].ms;
    out[i].calls = g_kernels[i].calls;
    out[i].bytes = g_kernels[i].bytes;
    out[i].flops = g_kernels[i].flops;
  }
  g_kernels.clear();
  return VADX_OK;
}

extern "C" int vadx_abi_version(void) { return VADX_ABI_VERSION; }
extern "C" const char* vadx_last_error(void) { return vadx::g_err; }
extern "C" uint64_t vadx_launch_count(void) { return vadx::g_launches.load(); }
extern "C" int vadx_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}
