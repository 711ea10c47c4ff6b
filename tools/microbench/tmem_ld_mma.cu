// Micro-benchmark: tcgen05.ld throughput of 8 epilogue warps WHILE the tensor core runs back-to-back MMAs into another
// accumulator of the same SM (the fused-MLP ping-pong).  Also times the MMA stream alone and the loads alone.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I eonerf_code_b200/csrc -I include -o tools/microbench/tmem_ld_mma tools/microbench/tmem_ld_mma.cu
#include <cstdio>
#include "tc_ptx.cuh"

using namespace eonerf;

// mode bit 0: run MMAs, bit 1: run loads (+ optional STS of the packed result, bit 2)
__global__ void __launch_bounds__(320, 1) bench(int mode, int iters, long long* cycles, uint32_t* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint32_t tbase;
  __shared__ uint64_t bar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 1) tmem_alloc(&tbase, 512);
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u;   // bf16 ~0.0078
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tbase;
  long long t0 = 0, t1 = 0;
  uint32_t sink = 0;
  if (warp == 1) {
    if (mode & 1) {
      const bool elected = elect_one_sync();
      const uint32_t idesc = instr_desc(128, 256, 0, 0);
      const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 65536);
      t0 = clock64();
      for (int i = 0; i < iters; ++i) {
        // one "layer": 16 MMAs (K = 256) over 4 A blocks x 2 B halves... B = 256 rows x 64 K = 32 KB per k block
        for (int kb = 0; kb < 4; ++kb) {
          const uint32_t la = desc_lo_k128(a0 + kb * 16384), lb = desc_lo_k128(b0);
          if (elected) {
            umma_k128<1>(tb, la, lb, idesc, 1);
            umma_k128<1>(tb, la + 2, lb + 2, idesc, 1);
            umma_k128<1>(tb, la + 4, lb + 4, idesc, 1);
            umma_k128<1>(tb, la + 6, lb + 6, idesc, 1);
          }
          __syncwarp();
        }
        if (elected) umma_commit(&bar);
        __syncwarp();
        mbar_wait(&bar, i & 1);
      }
      t1 = clock64();
      if (lane == 0) cycles[blockIdx.x * 2] = t1 - t0;
    }
  } else if (warp >= 2) {
    if (mode & 2) {
      const int q = warp & 3, half = (warp - 2) >> 2;
      const uint32_t taddr = tb + 256 + ((uint32_t)(q * 32) << 16) + half * 128;
      const int r = q * 32 + lane;
      t0 = clock64();
      for (int i = 0; i < iters; ++i) {
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t v[32];
          tmem_ld32(taddr + c * 32, v);
          tmem_ld_wait_dep(v);
          if (mode & 4) {
            uint32_t pk[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(pk[j]) : "f"(__uint_as_float(v[2 * j + 1])), "f"(__uint_as_float(v[2 * j])));
            }
            const uint32_t blk = smem_u32(smem + 131072) + (uint32_t)((half * 128 + c * 32) >> 6) * 16384;
            const int ch0 = ((half * 128 + c * 32) & 63) >> 3;
#pragma unroll
            for (int jj = 0; jj < 4; ++jj)
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(blk + (uint32_t)(r * 128 + (((ch0 + jj) ^ (r & 7)) << 4))), "r"(pk[4 * jj]),
                           "r"(pk[4 * jj + 1]), "r"(pk[4 * jj + 2]), "r"(pk[4 * jj + 3])
                           : "memory");
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) sink ^= v[j];
          }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      t1 = clock64();
      if (threadIdx.x == 64) cycles[blockIdx.x * 2 + 1] = t1 - t0;
      out[blockIdx.x * 256 + threadIdx.x - 64] = sink;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tb, 512); }
}

int main() {
  long long* cyc; uint32_t* out;
  cudaMalloc(&cyc, 148 * 16); cudaMalloc(&out, 148 * 256 * 4);
  const int smem_bytes = 200 * 1024;
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  const int iters = 2000;
  const char* names[8] = {"", "MMA only", "loads only", "MMA + loads", "", "", "loads+pack+STS only", "MMA + loads+pack+STS"};
  for (int mode : {1, 2, 3, 6, 7}) {
    cudaMemset(cyc, 0, 148 * 16);
    for (int rep = 0; rep < 2; ++rep) { bench<<<148, 320, smem_bytes>>>(mode, iters, cyc, out); cudaDeviceSynchronize(); }
    long long h[296];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%-22s: MMA layer (16 x 128x256x16) %8.1f cycles | epilogue of one 128x256 accumulator %8.1f cycles   (%s)\n", names[mode],
           (double)h[0] / iters, (double)h[1] / iters, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
