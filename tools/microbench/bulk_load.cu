// Micro-benchmark: how fast can cp.async.bulk (global -> shared) stream the dW GEMM's operands?  Mimics gemm_tn_blocked_kernel's
// producer: persistent CTAs (one per SM), CTA b owns a contiguous run of 64-sample chunks; a chunk = 4 G + 4 X half-blocks of 8 KB
// (the halves of 16 KB tile-blocked blocks), STAGES chunks in flight, the consumer frees a stage as soon as it has landed.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/bulk_load tools/microbench/bulk_load.cu && /tmp/bulk_load
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar, int hint, uint64_t pol) {
  if (hint)
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
  else
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// mode 0: 8 KB half-block boxes, a chunk = one half of 8 blocks; mode 1: 16 KB whole blocks, a "chunk" = 4 whole blocks (same 64 KB per stage)
template <int STAGES>
__global__ void __launch_bounds__(128, 1) k(const uint8_t* G0, const uint8_t* X0, int64_t chunks, int hint, int mode, int n_arr = 1, size_t arr = 0) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[STAGES], empty[STAGES];
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // n_arr > 1 (mode 2): ONE launch over all array pairs, CTA b owns [b, b+1) / gridDim of the concatenated chunk axis (the dW kernel's split)
  const int64_t total = chunks * n_arr;
  const int64_t c0 = total * blockIdx.x / gridDim.x, c1 = total * (blockIdx.x + 1) / gridDim.x;
  if (threadIdx.x == 0) {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    int st = 0; uint32_t ph = 0;
    for (int64_t c = c0; c < c1; ++c) {
      mbar_wait(&empty[st], ph ^ 1);
      uint8_t* s0 = smem + (size_t)st * 65536;
      mbar_expect_tx(&full[st], 65536);
      if (mode == 0 || mode == 2) {
        const int64_t a = c / chunks, cc = c - a * chunks;
        const uint8_t* G = G0 + a * arr; const uint8_t* X = X0 + a * arr;
        const int64_t tile = cc >> 1;
        const size_t hoff = (size_t)(cc & 1) * 8192;
        for (int b = 0; b < 4; ++b) bulk_load(s0 + b * 8192, G + ((size_t)tile * 4 + b) * 16384 + hoff, 8192, &full[st], hint, pol);
        for (int b = 0; b < 4; ++b) bulk_load(s0 + (4 + b) * 8192, X + ((size_t)tile * 4 + b) * 16384 + hoff, 8192, &full[st], hint, pol);
      } else {
        // chunk pairs: even chunk takes the G blocks of the tile, odd chunk the X blocks (whole 16 KB blocks)
        const int64_t tile = c >> 1;
        const uint8_t* src = (c & 1) ? X0 : G0;
        for (int b = 0; b < 4; ++b) bulk_load(s0 + b * 16384, src + ((size_t)tile * 4 + b) * 16384, 16384, &full[st], hint, pol);
      }
      if (++st == STAGES) { st = 0; ph ^= 1; }
    }
  } else if (threadIdx.x == 32) {
    int st = 0; uint32_t ph = 0;
    for (int64_t c = c0; c < c1; ++c) {
      mbar_wait(&full[st], ph);
      mbar_arrive(&empty[st]);
      if (++st == STAGES) { st = 0; ph ^= 1; }
    }
  }
}

__global__ void fill_hash(uint32_t* p, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    uint32_t x = (uint32_t)i * 2654435761u; x ^= x >> 15; x *= 2246822519u; x ^= x >> 13;
    p[i] = x;
  }
}

int main(int argc, char** argv) {
  const int sms = 148;
  const int64_t n_tiles = 11719;                              // 1.5 M samples
  const int64_t chunks = 2 * n_tiles;
  const size_t arr = (size_t)n_tiles * 4 * 16384;             // one [n, 256] bf16 array, tile-blocked
  const int n_arr = 11;                                       // as many G / X pairs as the step has 256-wide dW GEMMs
  uint8_t *G, *X;
  cudaMalloc(&G, arr * n_arr); cudaMalloc(&X, arr * n_arr);
  cudaMemset(G, 1, arr * n_arr); cudaMemset(X, 1, arr * n_arr);
  if (argc > 1) {                                             // any argument: pseudo-random contents instead of a constant byte
    fill_hash<<<1184, 256>>>((uint32_t*)G, arr * n_arr / 4); fill_hash<<<1184, 256>>>((uint32_t*)X, arr * n_arr / 4);
    cudaDeviceSynchronize();
    printf("pseudo-random contents\n");
  }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto run = [&](auto kern, int stages, const char* name, int hint, int mode) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, stages * 65536);
    float best = 1e9;
    for (int it = 0; it < 3; ++it) {
      cudaEventRecord(e0);
      for (int a = 0; a < n_arr; ++a) kern<<<sms, 128, stages * 65536>>>(G + a * arr, X + a * arr, chunks, hint, mode, 1, (size_t)0);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (ms < best) best = ms;
    }
    printf("%-52s hint=%d: %7.3f ms  %7.1f GB/s  (%s)\n", name, hint, best, 2.0 * arr * n_arr / best / 1e6, cudaGetErrorString(cudaGetLastError()));
  };
  printf("%.2f GB per run, %d CTAs\n", 2.0 * arr * n_arr / 1e9, sms);
  for (int hint = 0; hint < 2; ++hint) {
    run(k<1>, 1, "8 KB half-block boxes, 1 x 64 KB in flight", hint, 0);
    run(k<2>, 2, "8 KB half-block boxes, 2 x 64 KB in flight", hint, 0);
    run(k<3>, 3, "8 KB half-block boxes, 3 x 64 KB in flight", hint, 0);
    run(k<3>, 3, "16 KB whole blocks, 3 x 64 KB in flight", hint, 1);
  }
  {
    cudaFuncSetAttribute(k<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * 65536);
    float best = 1e9;
    for (int it = 0; it < 3; ++it) {
      cudaEventRecord(e0);
      k<3><<<sms, 128, 3 * 65536>>>(G, X, chunks, 1, 2, n_arr, arr);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (ms < best) best = ms;
    }
    printf("%-52s hint=1: %7.3f ms  %7.1f GB/s  (%s)\n", "ONE launch, CTA b owns 1/148 of the concatenated axis", best, 2.0 * arr * n_arr / best / 1e6, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
