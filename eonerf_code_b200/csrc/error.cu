// Error reporting, ABI version and device check of the C ABI (include/eonerf_b200.h).
#include <stdarg.h>

#include "common.cuh"

#include <atomic>
#include <mutex>
#include <vector>

namespace eonerf {
static thread_local char g_err[512] = "";

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

struct ProfRec { cudaEvent_t a, b; int kind; double flops, bytes; };
static bool g_prof_on = false;
static std::vector<ProfRec> g_recs;
static std::vector<cudaEvent_t> g_pool;
static std::mutex g_prof_mu;

static cudaEvent_t take_event() {
  if (!g_pool.empty()) { cudaEvent_t e = g_pool.back(); g_pool.pop_back(); return e; }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}

void profile_begin(int kind, double flops, double bytes, cudaStream_t s) {
  if (!g_prof_on) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  ProfRec r{take_event(), take_event(), kind, flops, bytes};
  cudaEventRecord(r.a, s);
  g_recs.push_back(r);
}
void profile_end(cudaStream_t s) {
  if (!g_prof_on) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (!g_recs.empty()) cudaEventRecord(g_recs.back().b, s);
}

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace eonerf

extern "C" int eonerf_abi_version(void) { return EONERF_ABI_VERSION; }

extern "C" const char* eonerf_last_error(void) { return eonerf::g_err; }

extern "C" int eonerf_check_device(void) {
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    eonerf::set_error("no CUDA device: %s", cudaGetErrorString(e));
    cudaGetLastError();
    return EONERF_EDEVICE;
  }
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (major != 10) {
    eonerf::set_error("device %d has compute capability %d.%d; this library is built for sm_100a only", dev, major, minor);
    return EONERF_EDEVICE;
  }
  return EONERF_OK;
}

extern "C" int64_t eonerf_launch_count(int32_t reset) {
  long long v = eonerf::g_launches.load();
  if (reset) eonerf::g_launches.store(0);
  return v;
}

extern "C" int eonerf_profile_enable(int32_t on) {
  std::lock_guard<std::mutex> lk(eonerf::g_prof_mu);
  eonerf::g_prof_on = on != 0;
  return EONERF_OK;
}

extern "C" int eonerf_profile_read(EonerfProfile* out, int32_t n_kinds) {
  using namespace eonerf;
  EO_REQUIRE(out && n_kinds > 0, "profile_read: bad arguments");
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (int k = 0; k < n_kinds; ++k) out[k] = EonerfProfile{0, 0.0, 0.0, 0.0};
  for (auto& r : g_recs) {
    float ms = 0.f;
    if (cudaEventSynchronize(r.b) == cudaSuccess && cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess && r.kind < n_kinds) {
      out[r.kind].launches += 1;
      out[r.kind].ms += ms;
      out[r.kind].flops += r.flops;
      out[r.kind].bytes += r.bytes;
    }
    g_pool.push_back(r.a);
    g_pool.push_back(r.b);
  }
  g_recs.clear();
  cudaGetLastError();
  return EONERF_OK;
}
