// Device-side helpers shared by the fused forward (field_fused.cu) and backward (field_fused_bwd.cu) kernels:
// the shared-memory map, explicit shared-space accesses, bf16 packing, reduced-range sine / cosine.
#pragma once
#include "field_fused.cuh"
#include "tc_ptx.cuh"

namespace eonerf {

constexpr int kRingStages = 3;
constexpr int kSlotBytes = 5 * kBlkBytes;                  // ACT blocks 0..3 + ENC block 4
constexpr int kOffRing = 0;
constexpr int kOffSlot = kRingStages * kBlkBytes;          // 49152
constexpr int kOffConst = kOffSlot + 2 * kSlotBytes;       // 212992
constexpr int kConstBytes = 15872;
constexpr int kOffPart = kOffConst + kConstBytes;          // [128][2] floats
constexpr int kOffBar = kOffPart + 1024;
constexpr int kSmemFused = kOffBar + 128 + 1024;           // + alignment slack
constexpr int kFusedThreads = 320;
constexpr int kEpiThreads = 256;

static_assert(kCFloats * 4 <= kConstBytes, "constants do not fit");
static_assert(kSmemFused <= 232448, "shared memory budget");

__device__ __forceinline__ float sigmoid_f(float v) { return 1.0f / (1.0f + expf(-v)); }          // nn.Sigmoid
__device__ __forceinline__ float softplus_f(float v) { return v > 20.0f ? v : log1pf(expf(v)); }   // nn.Softplus(beta=1, threshold=20)

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float bf_lo(uint32_t p) { return __uint_as_float(p << 16); }
__device__ __forceinline__ float bf_hi(uint32_t p) { return __uint_as_float(p & 0xFFFF0000u); }

// shared-memory byte offset of (row, 16-byte chunk) inside a [128 x 64] bf16 block
__device__ __forceinline__ uint32_t blk_off(int row, int chunk) { return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4)); }

// explicit shared-space accesses (the 1024-byte alignment arithmetic on the dynamic shared-memory base hides the address
// space from the compiler, which would otherwise emit generic LD/ST)
__device__ __forceinline__ void lds_f4(uint32_t a, float& x, float& y, float& z, float& w) {
  asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x), "=f"(y), "=f"(z), "=f"(w) : "r"(a));
}
__device__ __forceinline__ void lds_f2(uint32_t a, float& x, float& y) {
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(x), "=f"(y) : "r"(a) : "memory");
}
__device__ __forceinline__ void sts_f2(uint32_t a, float x, float y) {
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(a), "f"(x), "f"(y) : "memory");
}
__device__ __forceinline__ void sts_u4(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}

// sin(a) for |a| < ~1e3: two-term Cody-Waite reduction to [-pi, pi] then MUFU.SIN.  Absolute error < 6e-7, far below
// the bf16 resolution the value is rounded to (2^-9 relative), at ~7 instructions instead of ~45 for sinf().
__device__ __forceinline__ float sin_reduced(float a) {
  const float k = rintf(a * 0.15915494309189535f);
  float rr = fmaf(k, -6.2831854820251465f, a);
  rr = fmaf(k, 1.7484555314695172e-7f, rr);
  return __sinf(rr);
}

// positional-encoding column kBase+c (c compile-time after unrolling) of position x
// (mlp.py:199-205: [x, sin(2^k x) k=0..9 (frequency-major), sin(2^k x + pi/2)], column 63 is padding)
template <int kBase>
__device__ __forceinline__ float posenc_col(const float (&x)[3], int c) {
  c += kBase;
  if (c < 3) return x[c];
  if (c >= 63) return 0.f;
  int e = c - 3;
  const int hf = e >= 30;
  e -= hf * 30;
  const float xb = x[e % 3] * (float)(1 << (e / 3));
  return sin_reduced(hf ? __fadd_rn(xb, kHalfPi) : xb);                  // torch adds the scalar in fp32 (mlp.py:203)
}


// cos(a) with the same reduction
__device__ __forceinline__ float cos_reduced(float a) {
  const float k = rintf(a * 0.15915494309189535f);
  float rr = fmaf(k, -6.2831854820251465f, a);
  rr = fmaf(k, 1.7484555314695172e-7f, rr);
  return __cosf(rr);
}

static inline int fused_grid(int64_t n_pairs) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return (int)(n_pairs < sms ? n_pairs : sms);
}

}  // namespace eonerf
