// Micro-benchmark: how fast can cp.async.bulk (shared -> global) stream the forward kernel's activation stash to HBM?
// Mimics the fused forward: every CTA (one per SM) stores 64 KB per "slot-layer" (4 blocks of 16 KB) into 13 arrays laid out
// like the tile-blocked stash, with at most DEPTH slot-layers in flight (cp.async.bulk.wait_group.read DEPTH-1 before reuse).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/bulk_store tools/microbench/bulk_store.cu && /tmp/bulk_store
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_store(void* dst, const void* src, uint32_t bytes, bool hint, uint64_t pol) {
  if (hint)
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes), "l"(pol) : "memory");
  else
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
template <int N> __device__ __forceinline__ void wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }

// mode 0: bulk stores; mode 1: st.global.v4 by all threads (coalesced 16 KB blocks) for comparison
template <int DEPTH>
__global__ void __launch_bounds__(256, 1) k(uint8_t* out, int64_t tiles_per_cta, int n_layers, int64_t layer_stride, int hint, int blk_bytes, int mode) {
  extern __shared__ __align__(1024) uint8_t smem[];
  for (int i = threadIdx.x; i < 65536 * 2 / 4; i += 256) ((uint32_t*)smem)[i] = i;
  __syncthreads();
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  const int n_blk = 65536 / blk_bytes;
  for (int64_t t = 0; t < tiles_per_cta; ++t) {
    const int64_t tile = (int64_t)blockIdx.x * tiles_per_cta + t;
    for (int l = 0; l < n_layers; ++l) {
      uint8_t* dst = out + l * layer_stride + tile * 65536;
      const uint8_t* src = smem + ((t * n_layers + l) & 1) * 65536;
      if (mode == 0) {
        if (threadIdx.x == 0) {
          wait_read<DEPTH - 1>();
          for (int b = 0; b < n_blk; ++b) bulk_store(dst + (size_t)b * blk_bytes, src + (size_t)b * blk_bytes, blk_bytes, hint, pol);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      } else {
        for (int i = threadIdx.x; i < 65536 / 16; i += 256) ((uint4*)dst)[i] = ((const uint4*)src)[i];
      }
    }
  }
  if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main() {
  const int n_layers = 13, sms = 148;
  const int64_t tiles_per_cta = 52;                       // ~1M samples / 128 / 148
  const int64_t layer_stride = (int64_t)sms * tiles_per_cta * 65536;
  uint8_t* out;
  cudaMalloc(&out, layer_stride * n_layers);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const double bytes = (double)layer_stride * n_layers;
  auto run = [&](auto kern, const char* name, int hint, int blk, int mode) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072);
    float best = 1e9;
    for (int it = 0; it < 4; ++it) {
      cudaEventRecord(e0);
      kern<<<sms, 256, 131072>>>(out, tiles_per_cta, n_layers, layer_stride, hint, blk, mode);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (ms < best) best = ms;
    }
    printf("%-44s hint=%d blk=%5d: %7.3f ms  %7.1f GB/s  (%s)\n", name, hint, blk, best, bytes / best / 1e6, cudaGetErrorString(cudaGetLastError()));
  };
  printf("%.2f GB per run, %d CTAs\n", bytes / 1e9, sms);
  for (int hint = 0; hint < 2; ++hint) {
    run(k<1>, "bulk store, 1 slot-layer in flight", hint, 16384, 0);
    run(k<2>, "bulk store, 2 slot-layers in flight", hint, 16384, 0);
    run(k<2>, "bulk store, 2 in flight, 64 KB copies", hint, 65536, 0);
    run(k<2>, "bulk store, 2 in flight, 4 KB copies", hint, 4096, 0);
  }
  run(k<1>, "st.global.v4, 256 threads", 0, 16384, 1);
  return 0;
}
