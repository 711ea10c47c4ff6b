// Shared helpers for the eonerf_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/eonerf_b200.h"

namespace eonerf {

void set_error(const char* fmt, ...);

#define EO_REQUIRE(cond, ...)                      \
  do {                                             \
    if (!(cond)) {                                 \
      ::eonerf::set_error(__VA_ARGS__);            \
      return EONERF_EINVAL;                        \
    }                                              \
  } while (0)

#define EO_CUDA(expr)                                                                      \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      ::eonerf::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return EONERF_ECUDA;                                                                 \
    }                                                                                      \
  } while (0)

void count_launch();
#define EO_LAUNCH_CHECK()              \
  do {                                 \
    ::eonerf::count_launch();          \
    EO_CUDA(cudaPeekAtLastError());    \
  } while (0)

// Optional per-launch CUDA-event timing of the GEMM kernels (bench.py's live roofline measurement).
// kind: 0 = gemm_nt_tc, 1 = gemm_tn_tc, 2 = SIMT GEMMs.  No-ops unless eonerf_profile_enable(1) was called.
// cudaFuncSetAttribute applies to the CURRENT device: one-time kernel configuration has to happen once per device, not once per process
struct PerDeviceOnce {
  bool seen[64] = {};
  bool operator()() {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) return true;
    if (seen[d]) return false;
    seen[d] = true;
    return true;
  }
};

void profile_begin(int kind, double flops, double bytes, cudaStream_t s);
void profile_end(cudaStream_t s);

static inline cudaStream_t as_stream(eonerf_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

// inclusive scan across the 32 lanes
__device__ __forceinline__ float warp_inclusive_sum(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float t = __shfl_up_sync(kFull, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

// suffix-inclusive scan (lane i gets sum of lanes i..31)
__device__ __forceinline__ float warp_inclusive_suffix_sum(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float t = __shfl_down_sync(kFull, v, o);
    if (lane + o < 32) v += t;
  }
  return v;
}

static inline int div_up(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

}  // namespace eonerf
