// Micro-benchmark: tcgen05.ld throughput per SM (how fast can the epilogue warps drain a TMEM accumulator?).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tmem_ld tools/microbench/tmem_ld.cu && /tmp/tmem_ld
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int X>
__device__ __forceinline__ void ld(uint32_t taddr, uint32_t& sink) {
  if (X == 32) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) sink ^= r[i];
  } else {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) sink ^= r[i];
  }
}

// two loads in flight per wait
__device__ __forceinline__ void ld2(uint32_t taddr, uint32_t& sink) {
  uint32_t r[32], q[32];
#define LD32(ARR, ADDR)                                                                                                          \
  asm volatile(                                                                                                                  \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                                  \
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                                                  \
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"                                  \
      : "=r"(ARR[0]), "=r"(ARR[1]), "=r"(ARR[2]), "=r"(ARR[3]), "=r"(ARR[4]), "=r"(ARR[5]), "=r"(ARR[6]), "=r"(ARR[7]),         \
        "=r"(ARR[8]), "=r"(ARR[9]), "=r"(ARR[10]), "=r"(ARR[11]), "=r"(ARR[12]), "=r"(ARR[13]), "=r"(ARR[14]), "=r"(ARR[15]),   \
        "=r"(ARR[16]), "=r"(ARR[17]), "=r"(ARR[18]), "=r"(ARR[19]), "=r"(ARR[20]), "=r"(ARR[21]), "=r"(ARR[22]), "=r"(ARR[23]), \
        "=r"(ARR[24]), "=r"(ARR[25]), "=r"(ARR[26]), "=r"(ARR[27]), "=r"(ARR[28]), "=r"(ARR[29]), "=r"(ARR[30]), "=r"(ARR[31])  \
      : "r"(ADDR)                                                                                                                \
      : "memory")
  LD32(r, taddr);
  LD32(q, taddr + 32);
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) sink ^= r[i] ^ q[i];
}

template <int MODE>
__global__ void bench(int iters, long long* cycles, uint32_t* out) {
  __shared__ uint32_t tbase;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tbase)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = tbase + ((uint32_t)((warp & 3) * 32) << 16) + ((warp >> 2) & 1) * 128;
  uint32_t sink = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0) { ld<32>(base, sink); ld<32>(base + 32, sink); ld<32>(base + 64, sink); ld<32>(base + 96, sink); }
    if (MODE == 1) { ld<16>(base, sink); ld<16>(base + 16, sink); ld<16>(base + 32, sink); ld<16>(base + 48, sink);
                     ld<16>(base + 64, sink); ld<16>(base + 80, sink); ld<16>(base + 96, sink); ld<16>(base + 112, sink); }
    if (MODE == 2) { ld2(base, sink); ld2(base + 64, sink); }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = sink;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512) : "memory");
}

int main() {
  long long* cyc; uint32_t* out;
  cudaMalloc(&cyc, 148 * 8); cudaMalloc(&out, 148 * 512 * 4);
  const int iters = 2000;
  for (int mode = 0; mode < 3; ++mode)
    for (int warps : {4, 8, 16}) {
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) bench<0><<<148, warps * 32>>>(iters, cyc, out);
        if (mode == 1) bench<1><<<148, warps * 32>>>(iters, cyc, out);
        if (mode == 2) bench<2><<<148, warps * 32>>>(iters, cyc, out);
        cudaDeviceSynchronize();
      }
      long long h[148];
      cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      // bytes per iteration per CTA: warps x 32 lanes x 128 columns x 4 B
      const double bytes = (double)warps * 32 * 128 * 4;
      printf("mode %d (%s) warps %2d: %8.1f cycles/iter  -> %6.1f B/cycle/SM   (%s)\n", mode,
             mode == 0 ? "x32, wait each" : mode == 1 ? "x16, wait each" : "2 x x32 per wait", warps, (double)h[0] / iters,
             bytes * iters / (double)h[0], cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
