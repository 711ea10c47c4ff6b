// bf16 tensor-core GEMMs of the radiance-field MLP for sm_100a: tcgen05.mma with TMEM accumulators, operands
// staged by TMA (cp.async.bulk.tensor, 128-byte swizzle) through an mbarrier ring, warp-specialised
// (1 TMA thread, 1 MMA thread, 4 epilogue warps), persistent over output tiles.
//
//   gemm_nt_tc :  C[M,N] = epi(A[M,K] B[N,K]^T)    forward layers (B = W) and input gradients (B = W^T)
//                 both operands K-major; tile 128 x BLOCK_N(<=256) x 64; two TMEM accumulator stages so the MMA of
//                 tile i+1 overlaps the epilogue of tile i
//   gemm_tn_tc :  D[N,K] += A[M,N]^T X[M,K]        parameter gradients; the contraction runs over the SAMPLE axis,
//                 so both operands are MN-major views of the row-major activations (no transposes are ever
//                 materialised); split over CTAs along the samples, fp32 partial tiles merged with red.global.add
//
// Replaces the cuBLAS SGEMMs behind nn.Linear in /root/reference/radiance_fields/mlp.py:90,99 (forward) and their
// autograd backward.  Shared-memory operand layouts follow the canonical UMMA layouts:
//   K-major  SW128: rows of 64 bf16 (128 B), 8-row groups 1024 B apart (SBO), 16-byte chunks XOR-swizzled by row%8
//   MN-major SW128: the same physical image read the other way round: 64 contiguous MN elements x 8 K-rows per
//                   atom, K groups SBO = 1024 B apart, 64-wide MN blocks LBO = (one TMA box) apart.
#include <stdlib.h>

#include "field_fused.cuh"
#include "tc_ptx.cuh"

namespace eonerf {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;                 // 64 bf16 = one 128-byte swizzle row
constexpr int kBoxBytes = 64 * kBlockK * 2; // a [64 x 64] bf16 box (MN-major operands are loaded box by box)
constexpr int kThreads = 192;               // warp 0: TMA, warp 1: TMEM alloc + MMA, warps 2..5: epilogue

struct NTParams {
  int64_t M; int N, K;
  int block_n, n_tiles, k_blocks;
  int64_t m_tiles;
  __nv_bfloat16* C; int64_t ldc;
  const float* bias;
  const int32_t* row_class; const float* class_bias; int64_t ld_class;
  const __nv_bfloat16* addend; int64_t ld_add;
  const float* rank1_row; int64_t rank1_stride; const float* rank1_col;
  int relu;
  const __nv_bfloat16* mask; int64_t ld_mask; int mask_cols;
};

constexpr int kStoreBox = 32 * 64 * 2;      // one TMA store box: 32 rows x 64 bf16 columns, 128-byte swizzle

template <int kStages>
__global__ void __launch_bounds__(kThreads, 1) gemm_nt_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                  const __grid_constant__ CUtensorMap tmB,
                                                                  const __grid_constant__ CUtensorMap tmC, NTParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int a_bytes = kBlockM * kBlockK * 2;
  const int b_bytes = p.block_n * kBlockK * 2;
  const int stage_bytes = a_bytes + ((b_bytes + 1023) & ~1023);
  __shared__ uint64_t full_bar[kStages], empty_bar[kStages], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 4); }
    fence_barrier_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmC);
  }
  if (warp == 1) tmem_alloc(&tmem_base_s, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const int64_t total_tiles = p.m_tiles * p.n_tiles;

  if (warp == 0 && lane == 0) {
    // ===== TMA producer =====
    int stage = 0; uint32_t phase = 0;
    for (int64_t t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      int64_t mt = t / p.n_tiles; int nt = (int)(t % p.n_tiles);
      for (int kb = 0; kb < p.k_blocks; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + (size_t)stage * stage_bytes;
        mbar_expect_tx(&full_bar[stage], a_bytes + b_bytes);
        tma_load_2d(sa, &tmA, &full_bar[stage], kb * kBlockK, (int)(mt * kBlockM));
        tma_load_2d(sa + a_bytes, &tmB, &full_bar[stage], kb * kBlockK, nt * p.block_n);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ===== MMA issuer =====
    const uint32_t idesc = instr_desc(kBlockM, p.block_n, 0, 0);
    int stage = 0; uint32_t phase = 0;
    int acc = 0; uint32_t acc_phase = 0;
    for (int64_t t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      mbar_wait(&acc_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * 256;
      for (int kb = 0; kb < p.k_blocks; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
        const uint32_t sb = sa + a_bytes;
#pragma unroll
        for (int k = 0; k < kBlockK / 16; ++k) {
          // K-major SW128: 8-row groups 1024 B apart; stepping 16 elements along K moves the start by 32 bytes
          uint64_t da = smem_desc(sa + k * 32, 16, 1024);
          uint64_t db = smem_desc(sb + k * 32, 16, 1024);
          umma_bf16(d_tmem, da, db, idesc, (kb | k) != 0);
        }
        umma_commit(&empty_bar[stage]);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
      umma_commit(&acc_full[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= 2) {
    // ===== epilogue: TMEM -> registers -> bias / addend / rank-1 / ReLU / mask -> bf16 -> swizzled smem -> TMA store =====
    // Each warp owns 32 rows of the tile and streams them out as [32 x 64] boxes through two private 4 KB staging
    // buffers; the TMA store clips rows >= M and columns >= N, so ragged tails need no predication here.
    const int quarter = warp & 3;           // a warp may only touch TMEM lanes [32*(warp%4), +32)
    uint8_t* stg = smem + (size_t)kStages * stage_bytes + (size_t)(warp - 2) * 2 * kStoreBox;
    int buf = 0;
    int acc = 0; uint32_t acc_phase = 0;
    const int groups = p.block_n / 64;
    for (int64_t t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      int64_t mt = t / p.n_tiles; int nt = (int)(t % p.n_tiles);
      const int64_t m = mt * kBlockM + quarter * 32 + lane;
      const bool row_ok = m < p.M;
      const int n0 = nt * p.block_n;
      const float* brow = p.class_bias ? (row_ok ? p.class_bias + (int64_t)__ldg(p.row_class + m) * p.ld_class : p.class_bias) : p.bias;
      const float r1 = (p.rank1_row && row_ok) ? __ldg(p.rank1_row + m * p.rank1_stride) : 0.f;
      mbar_wait(&acc_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * 256 + ((uint32_t)(quarter * 32) << 16);
      for (int g = 0; g < groups; ++g) {
        __syncwarp();
        uint32_t ra[32], rb[32];
        tmem_ld32(taddr + g * 64, ra);
        tmem_ld32(taddr + g * 64 + 32, rb);
        if (lane == 0) tma_store_wait_read<1>();                  // the store that last read stg[buf] has drained it
        tmem_ld_wait();
        if (g == groups - 1) {                                    // accumulator fully read: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[acc]);
        }
        __syncwarp();
        uint8_t* box = stg + buf * kStoreBox;
#pragma unroll
        for (int v = 0; v < 8; ++v) {                             // 8 columns -> one 16-byte chunk
          const int n = n0 + g * 64 + v * 8;
          float f[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(v < 4 ? ra[v * 8 + j] : rb[(v - 4) * 8 + j]);
          if (row_ok && n < p.N) {
            if (brow) {
              const float4 b0 = __ldg((const float4*)(brow + n)), b1 = __ldg((const float4*)(brow + n + 4));
              f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w; f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
            }
            if (p.addend) {
              const uint4 a = __ldg((const uint4*)(p.addend + m * p.ld_add + n));
              const __nv_bfloat162* a2 = (const __nv_bfloat162*)&a;
#pragma unroll
              for (int j = 0; j < 4; ++j) { float2 x = __bfloat1622float2(a2[j]); f[2 * j] += x.x; f[2 * j + 1] += x.y; }
            }
            if (p.rank1_row) {
              const float4 c0 = __ldg((const float4*)(p.rank1_col + n)), c1 = __ldg((const float4*)(p.rank1_col + n + 4));
              f[0] += r1 * c0.x; f[1] += r1 * c0.y; f[2] += r1 * c0.z; f[3] += r1 * c0.w;
              f[4] += r1 * c1.x; f[5] += r1 * c1.y; f[6] += r1 * c1.z; f[7] += r1 * c1.w;
            }
            if (p.relu) {
#pragma unroll
              for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
            }
            if (p.mask && n < p.mask_cols) {
              const uint4 mk = __ldg((const uint4*)(p.mask + m * p.ld_mask + n));
              const __nv_bfloat162* m2 = (const __nv_bfloat162*)&mk;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                float2 x = __bfloat1622float2(m2[j]);
                if (!(x.x > 0.f)) f[2 * j] = 0.f;
                if (!(x.y > 0.f)) f[2 * j + 1] = 0.f;
              }
            }
          }
          uint4 o;
          __nv_bfloat162* o2 = (__nv_bfloat162*)&o;
#pragma unroll
          for (int j = 0; j < 4; ++j) o2[j] = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
          // 128-byte swizzle: chunk v of row `lane` lives at chunk position v ^ (lane % 8)  (conflict-free per quarter warp)
          *(uint4*)(box + lane * 128 + ((v ^ (lane & 7)) << 4)) = o;
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&tmC, box, n0 + g * 64, (int)(mt * kBlockM + quarter * 32));
          tma_store_commit();
        }
        buf ^= 1;
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (lane == 0) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// TN: D[n,k] += sum_m A[m,n] X[m,k]     (MMA "M" = n: output features of dY, MMA "N" = k: input features)
// ------------------------------------------------------------------------------------------------
struct TNParams {
  int64_t M;                 // samples (contraction)
  int N, K;                  // valid output extents
  int mt_count;              // 128-row blocks of N handled by one CTA (1 or 2)
  int block_k;               // width of the K tile (MMA N), multiple of 16, <= 256
  int k_tiles;               // number of K tiles
  int64_t chunks_per_cta;    // 64-sample chunks per CTA
  int64_t chunks;            // total 64-sample chunks
  float* D; int64_t ldd;
};

template <int kStages>
__global__ void __launch_bounds__(kThreads, 1) gemm_tn_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                  const __grid_constant__ CUtensorMap tmX, TNParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int a_boxes = p.mt_count * 2;                   // 64-feature boxes of dY
  const int x_boxes = (p.block_k + 63) / 64;            // 64-feature boxes of X
  const int stage_bytes = (a_boxes + x_boxes) * kBoxBytes;
  __shared__ uint64_t full_bar[kStages], empty_bar[kStages], acc_full;
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&acc_full, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmX);
  }
  if (warp == 1) tmem_alloc(&tmem_base_s, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  const int kt = blockIdx.x % p.k_tiles;                // which K tile
  const int64_t split = blockIdx.x / p.k_tiles;
  const int64_t c_begin = split * p.chunks_per_cta;
  const int64_t c_end = (c_begin + p.chunks_per_cta < p.chunks) ? c_begin + p.chunks_per_cta : p.chunks;
  const int64_t my_chunks = c_end > c_begin ? c_end - c_begin : 0;

  if (warp == 0 && lane == 0) {
    int stage = 0; uint32_t phase = 0;
    for (int64_t c = c_begin; c < c_end; ++c) {
      mbar_wait(&empty_bar[stage], phase ^ 1);
      uint8_t* s0 = smem + (size_t)stage * stage_bytes;
      mbar_expect_tx(&full_bar[stage], stage_bytes);
      for (int b = 0; b < a_boxes; ++b) tma_load_2d(s0 + b * kBoxBytes, &tmA, &full_bar[stage], b * 64, (int)(c * kBlockK));
      for (int b = 0; b < x_boxes; ++b)
        tma_load_2d(s0 + (a_boxes + b) * kBoxBytes, &tmX, &full_bar[stage], kt * p.block_k + b * 64, (int)(c * kBlockK));
      if (++stage == kStages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1 && lane == 0) {
    const uint32_t idesc = instr_desc(kBlockM, p.block_k, 1, 1);
    int stage = 0; uint32_t phase = 0;
    for (int64_t c = 0; c < my_chunks; ++c) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      const uint32_t s0 = smem_u32(smem + (size_t)stage * stage_bytes);
      const uint32_t sx = s0 + a_boxes * kBoxBytes;
#pragma unroll
      for (int k = 0; k < kBlockK / 16; ++k) {
        // MN-major SW128: atom = 64 MN elements x 8 K rows (1024 B); 16 K rows per MMA = 2 atoms = 2048 B per step;
        // 64-wide MN blocks are one box (LBO) apart
        const uint64_t dx = smem_desc(sx + k * 2048, kBoxBytes, 1024);
        for (int mt = 0; mt < p.mt_count; ++mt) {
          const uint64_t da = smem_desc(s0 + mt * 2 * kBoxBytes + k * 2048, kBoxBytes, 1024);
          umma_bf16(tmem_base + mt * 256, da, dx, idesc, (c | k) != 0);
        }
      }
      umma_commit(&empty_bar[stage]);
      if (++stage == kStages) { stage = 0; phase ^= 1; }
    }
    umma_commit(&acc_full);
  } else if (warp >= 2 && my_chunks > 0) {
    const int quarter = warp & 3;
    mbar_wait(&acc_full, 0);
    tc_fence_after();
    for (int mt = 0; mt < p.mt_count; ++mt) {
      const int n = mt * kBlockM + quarter * 32 + lane;          // output row (feature of dY)
      const uint32_t taddr = tmem_base + mt * 256 + ((uint32_t)(quarter * 32) << 16);
      for (int c = 0; c < p.block_k; c += 32) {
        __syncwarp();
        uint32_t r[32];
        const int width = (p.block_k - c) >= 32 ? 32 : 16;
        if (width == 32) tmem_ld32(taddr + c, r); else tmem_ld16(taddr + c, r);
        tmem_ld_wait();
        if (n >= p.N) continue;
        float* drow = p.D + (int64_t)n * p.ldd + kt * p.block_k + c;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < width && kt * p.block_k + c + j < p.K) atomicAdd(drow + j, __uint_as_float(r[j]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// dbias[n] += sum_m A[m,n]:  N/8 lanes x 8 columns own a row (one 16-byte load), 4 rows in flight per thread, smem tree
__global__ void __launch_bounds__(256) colsum_bf16_kernel(const __nv_bfloat16* __restrict__ A, int64_t lda, int64_t M, int N,
                                                          float* __restrict__ out, int64_t rows_per_block) {
  __shared__ float red[256][9];
  const int lpr = N >> 3, rows = 256 / lpr;                      // N in {64, 128, 256}
  const int sub = threadIdx.x % lpr, rsub = threadIdx.x / lpr;
  const int64_t m_begin = (int64_t)blockIdx.x * rows_per_block;
  const int64_t m_end = m_begin + rows_per_block < M ? m_begin + rows_per_block : M;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int64_t m0 = m_begin + rsub; m0 < m_end; m0 += 4 * rows) {
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t m = m0 + (int64_t)u * rows;
      v[u] = m < m_end ? __ldg((const uint4*)(A + m * lda + sub * 8)) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const __nv_bfloat162* h = (const __nv_bfloat162*)&v[u];
#pragma unroll
      for (int j = 0; j < 4; ++j) { float2 x = __bfloat1622float2(h[j]); acc[2 * j] += x.x; acc[2 * j + 1] += x.y; }
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) red[threadIdx.x][e] = acc[e];
  __syncthreads();
  if (threadIdx.x < N) {
    const int n = threadIdx.x, s0 = n >> 3, e = n & 7;
    float v = 0.f;
    for (int r = 0; r < rows; ++r) v += red[r * lpr + s0][e];
    atomicAdd(out + n, v);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// 2-D bf16 row-major [rows, cols] (leading dimension ld elements); box = 64 columns x box_rows rows, 128-byte swizzle;
// out-of-bounds elements read as zero (ragged M / K tails need no special casing in the kernels)
// plain (un-swizzled) map over the pre-tiled weight blob of the fused kernels: rows of 64 bf16 (one 128-byte row of a
// block image), boxes of 64 rows = 8 KB, copied verbatim
int make_blob_map(CUtensorMap* map, const void* base, int64_t n_blocks) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return EONERF_ECUDA; }
  cuuint64_t gdim[2] = {64, (cuuint64_t)n_blocks * 128};
  cuuint64_t gstr[1] = {128};
  cuuint32_t box[2] = {64, 64};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (weight blob) failed (%d)", (int)r); return EONERF_ECUDA; }
  return EONERF_OK;
}

static int make_map(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  if (rows <= 0 || cols <= 0) { set_error("empty tensor map"); return EONERF_EINVAL; }
  EncodeTiledFn fn = encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return EONERF_ECUDA; }
  if (((uintptr_t)base & 15) || (ld * 2) % 16) {
    set_error("tensor-core GEMM operands need 16-byte aligned bases and leading dimensions that are multiples of 8 (ld=%lld)", (long long)ld);
    return EONERF_EINVAL;
  }
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld ld=%lld", (int)r, (long long)rows, (long long)cols, (long long)ld); return EONERF_ECUDA; }
  return EONERF_OK;
}

static int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

constexpr int kNTStages = 4;
constexpr int kTNStages = 3;
constexpr int kSmemNT = kNTStages * (kBlockM * kBlockK * 2 + 256 * kBlockK * 2) + 4 * 2 * kStoreBox + 1024;
constexpr int kSmemTN = kTNStages * 8 * kBoxBytes + 1024;

int gemm_nt_tc(const GemmNT& g, cudaStream_t s) {
  if (g.M <= 0 || g.N <= 0) return EONERF_OK;
  EO_REQUIRE(g.K > 0 && g.K % 8 == 0, "gemm_nt_tc: K must be a positive multiple of 8 (got %d)", g.K);
  EO_REQUIRE(g.N % 8 == 0 && g.ldc % 8 == 0 && ((uintptr_t)g.C & 15) == 0, "gemm_nt_tc: N and ldc must be multiples of 8, C 16-byte aligned");
  EO_REQUIRE(!g.addend || (g.ld_add % 8 == 0 && ((uintptr_t)g.addend & 15) == 0), "gemm_nt_tc: misaligned addend");
  EO_REQUIRE(!g.mask || (g.ld_mask % 8 == 0 && ((uintptr_t)g.mask & 15) == 0 && g.mask_cols % 8 == 0), "gemm_nt_tc: misaligned mask");
  EO_REQUIRE(!g.class_bias || g.ld_class % 4 == 0, "gemm_nt_tc: class_bias rows must be 16-byte aligned");
  static PerDeviceOnce once;
  if (once()) {
    EO_CUDA(cudaFuncSetAttribute(gemm_nt_tc_kernel<kNTStages>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemNT));
  }
  NTParams p{};
  p.M = g.M; p.N = g.N; p.K = g.K;
  // N tiles are multiples of 64 columns (= one TMA store box); columns beyond N are zero weights in, clipped on the way out
  const int n64 = (g.N + 63) / 64;
  p.n_tiles = (n64 + 3) / 4;
  p.block_n = ((n64 + p.n_tiles - 1) / p.n_tiles) * 64;
  p.k_blocks = (g.K + kBlockK - 1) / kBlockK;
  p.m_tiles = (g.M + kBlockM - 1) / kBlockM;
  p.C = (__nv_bfloat16*)g.C; p.ldc = g.ldc; p.bias = g.bias;
  p.row_class = g.row_class; p.class_bias = g.class_bias; p.ld_class = g.ld_class;
  p.addend = (const __nv_bfloat16*)g.addend; p.ld_add = g.ld_add;
  p.rank1_row = g.rank1_row; p.rank1_stride = g.rank1_stride; p.rank1_col = g.rank1_col;
  p.relu = g.relu; p.mask = (const __nv_bfloat16*)g.mask; p.ld_mask = g.ld_mask; p.mask_cols = g.mask_cols;
  CUtensorMap tmA, tmB, tmC;
  int r;
  if ((r = make_map(&tmA, g.A, g.M, g.K, g.lda, kBlockM)) != EONERF_OK) return r;
  if ((r = make_map(&tmB, g.B, g.N, g.K, g.ldb, p.block_n)) != EONERF_OK) return r;
  if ((r = make_map(&tmC, g.C, g.M, g.N, g.ldc, 32)) != EONERF_OK) return r;
  int64_t tiles = p.m_tiles * p.n_tiles;
  int grid = (int)(tiles < sm_count() ? tiles : sm_count());
  const double ak = g.alg_k ? g.alg_k : g.K;
  profile_begin(0, 2.0 * g.M * g.N * ak, 2.0 * (g.M * ak + (double)g.N * ak + (double)g.M * g.N), s);
  gemm_nt_tc_kernel<kNTStages><<<grid, kThreads, kSmemNT, s>>>(tmA, tmB, tmC, p);
  profile_end(s);
  EO_LAUNCH_CHECK();
  return EONERF_OK;
}

int gemm_tn_tc(const GemmTN& g, cudaStream_t s) {
  if (g.M <= 0 || g.N <= 0 || g.K <= 0) return EONERF_OK;
  static PerDeviceOnce once;
  if (once()) {
    EO_CUDA(cudaFuncSetAttribute(gemm_tn_tc_kernel<kTNStages>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTN));
  }
  if (g.dbias) {
    EO_REQUIRE(g.N == 64 || g.N == 128 || g.N == 256, "gemm_tn_tc: dbias supports N in {64,128,256} (got %d)", g.N);
    int64_t rows = 1024;
    while (div_up(g.M, rows) > 4 * sm_count()) rows *= 2;
    colsum_bf16_kernel<<<div_up(g.M, rows), 256, 0, s>>>((const __nv_bfloat16*)g.A, g.lda, g.M, g.N, g.dbias, rows);
    EO_LAUNCH_CHECK();
  }
  if (!g.D) return EONERF_OK;
  EO_REQUIRE(g.N <= 256, "gemm_tn_tc: N (features of dY) must be <= 256 (got %d)", g.N);
  TNParams p{};
  p.M = g.M; p.N = g.N; p.K = g.K;
  p.mt_count = (g.N + kBlockM - 1) / kBlockM;
  // K tiles are whole 64-element swizzle atoms (MN-major operands are read atom by atom); columns beyond K load as zeros
  const int kp = (g.K + 63) / 64 * 64;
  p.k_tiles = (kp + 255) / 256;
  p.block_k = ((kp / 64 + p.k_tiles - 1) / p.k_tiles) * 64;
  p.chunks = (g.M + kBlockK - 1) / kBlockK;
  int64_t splits = sm_count() / p.k_tiles;
  if (splits > p.chunks) splits = p.chunks;
  if (splits < 1) splits = 1;
  p.chunks_per_cta = (p.chunks + splits - 1) / splits;
  splits = (p.chunks + p.chunks_per_cta - 1) / p.chunks_per_cta;
  p.D = g.D; p.ldd = g.ldd;
  CUtensorMap tmA, tmX;
  int r;
  // boxes of [64 samples x 64 features]; the tensor-map column extent is the number of valid features (zeros beyond)
  if ((r = make_map(&tmA, g.A, g.M, g.N, g.lda, kBlockK)) != EONERF_OK) return r;
  if ((r = make_map(&tmX, g.X, g.M, g.K, g.ldx, kBlockK)) != EONERF_OK) return r;
  profile_begin(1, 2.0 * g.M * g.N * g.K, 2.0 * ((double)g.M * g.N * p.k_tiles + (double)g.M * g.K) + 4.0 * g.N * g.K, s);
  gemm_tn_tc_kernel<kTNStages><<<(unsigned)(splits * p.k_tiles), kThreads, kSmemTN, s>>>(tmA, tmX, p);
  profile_end(s);
  EO_LAUNCH_CHECK();
  return EONERF_OK;
}


// ------------------------------------------------------------------------------------------------
// Grouped TN over tile-blocked operands (field_fused.cuh): for every GEMM g of a list
//     D_g[n,k] += sum_m G_g[m,n] X_g[m,k],      db_g[n] += sum_m G_g[m,n]
// in ONE persistent launch.  A 64-sample half of a 16 KB block [128 samples x 64 features] is 8 KB contiguous and already
// the MN-major SW128 image the MMA wants, so operands arrive with plain cp.async.bulk copies (no tensor maps).
//
// Work split: the list is laid out on one axis of "cost" (8 KB boxes to stream); CTA b owns the interval
// [b, b+1) * total / gridDim.x and walks the GEMMs it overlaps.  A GEMM is therefore shared by only ~total/148-sized
// pieces (about 7 CTAs for a 23-GEMM backward pass) instead of by all 148 CTAs: the fp32 partial tiles are merged with
// red.global and the L2 atomic units, not HBM, were the bottleneck when every CTA contributed to every GEMM
// (measured: 4.26 ms -> 2.53 ms for the dW GEMMs of one training step with the merge switched off).
// The four epilogue warps are idle during the main loop: they sum the columns of the G tiles in shared memory (the bias
// gradient) while the tensor core consumes the same tiles.
// ------------------------------------------------------------------------------------------------
#ifndef EONERF_DW_LOAD_HINT
#define EONERF_DW_LOAD_HINT 1
#endif
constexpr int kTNGroupMax = 16;
struct TNBGemm {
  const uint8_t* G; const uint8_t* X;
  int32_t g_nb, g_blk0, mt_count, x_nb, x_blk0, x_cnt;
  int32_t n_valid[2]; int32_t k_valid;
  float* D[2]; int64_t ldd[2]; float* db[2];
  int32_t cost; int32_t pad_;  // cost units of one chunk of this GEMM (tnb_chunk_cost)
  int64_t per0;            // cost units per chunk index summed over the GEMMs before this one (cost0 = chunks * per0)
};
struct TNBParams {
  int32_t n; int32_t dbg_skip_tail;   // timing experiments (EONERF_TN_DBG): bit 0 no red.global tail, bit 2 no MMAs
  int64_t chunks;          // 64-sample chunks of every GEMM of the group (they share the sample axis)
  int64_t per_total;       // cost units per chunk index summed over the group (total cost = chunks * per_total)
  const int64_t* n_pts_dev;   // live sample count on the device: chunks = 2 * ceil(*n_pts_dev / 128)
  TNBGemm g[kTNGroupMax];
};

// The operand ring is 24 boxes (192 KB).  A stage holds one chunk of the CURRENT GEMM, so the stage count follows the chunk size:
// 8 boxes -> 3 stages, 5 -> 4, 4 -> 6, 3 -> 8.  With a fixed three stages the CTAs that own the narrow GEMMs (4 / 5 boxes per chunk)
// had 96-120 KB in flight, ran at 38 GB/s instead of 53 and finished 40 % after everybody else (tools/dw_cta_times.py,
// profiles/r2c_dw_cta_times_*.log).  Every role walks the same (GEMM, chunk) sequence and restarts at stage 0 when it moves to the next
// GEMM; the producer first waits until every stage of the old geometry has been consumed.
constexpr int kTNBRingBoxes = 24;
constexpr int kTNBMaxStages = 8;
constexpr int kSmemTNB = kTNBRingBoxes * kBoxBytes + 1024;
__device__ __forceinline__ int tnb_stages(int boxes) { const int n = kTNBRingBoxes / boxes; return n < kTNBMaxStages ? n : kTNBMaxStages; }

// chunk range [c0, c1) of GEMM gi owned by the CTA whose cost interval is [lo, hi)
__device__ __forceinline__ void tnb_range(const TNBGemm& g, int64_t chunks, int64_t lo, int64_t hi, int64_t& c0, int64_t& c1) {
  const int64_t per = g.cost;                                     // cost units per chunk
  const int64_t cost0 = chunks * g.per0;
  const int64_t end = cost0 + chunks * per;
  const int64_t a = lo > cost0 ? lo : cost0, b = hi < end ? hi : end;
  if (b <= a) { c0 = c1 = 0; return; }
  c0 = (a - cost0 + per - 1) / per;                               // a chunk belongs to the CTA that owns its first cost unit
  c1 = (b - cost0 + per - 1) / per;
  if (c1 > chunks) c1 = chunks;
}

#ifdef EONERF_TIMING
__device__ unsigned long long g_tnb_time[3][256];     // per CTA: globaltimer at start, when the producer has issued its last load, at exit
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#endif
__global__ void __launch_bounds__(kThreads, 1) gemm_tn_blocked_kernel(const __grid_constant__ TNBParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full_bar[kTNBMaxStages], empty_bar[kTNBMaxStages], acc_full, acc_empty;
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#ifdef EONERF_TIMING
  if (threadIdx.x == 0 && blockIdx.x < 256) g_tnb_time[0][blockIdx.x] = gtimer();
#endif
  if (threadIdx.x == 0) {
    for (int s = 0; s < kTNBMaxStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], (p.dbg_skip_tail & 16) ? 1 : 5); }   // MMA commit + 4 epilogue warps
    mbar_init(&acc_full, 1);
    mbar_init(&acc_empty, 4);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_s, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const int64_t chunks = p.n_pts_dev ? 2 * ((__ldg(p.n_pts_dev) + kTileM - 1) / kTileM) : p.chunks;
  const int64_t total_cost = chunks * p.per_total;
  const int64_t lo = total_cost * blockIdx.x / gridDim.x, hi = total_cost * (blockIdx.x + 1) / gridDim.x;

  if (warp == 0) {
    if (lane == 0) {
      // uses: bit s = how often stage s has been filled, mod 2 (fill k + 1 of a stage waits for its k-th release: parity (k - 1) & 1)
      int stage = 0; uint32_t uses = 0;
#if EONERF_DW_LOAD_HINT
      const uint64_t pol = l2_policy_evict_first();            // G and X are streamed once: do not let them displace anything
#define EO_DW_LOAD(dst, src, bytes, bar) bulk_load_hint(dst, src, bytes, bar, pol)
#else
#define EO_DW_LOAD(dst, src, bytes, bar) bulk_load(dst, src, bytes, bar)
#endif
      for (int gi = 0; gi < p.n; ++gi) {
        const TNBGemm& g = p.g[gi];
        int64_t c0, c1;
        tnb_range(g, chunks, lo, hi, c0, c1);
        if (c1 <= c0) continue;
        const int a_boxes = g.mt_count * 2;
        const int boxes = a_boxes + g.x_cnt, n_st = tnb_stages(boxes);
        // new stage geometry: every stage of the previous GEMM must have been consumed before its bytes are overwritten
        for (int st = 0; st < kTNBMaxStages; ++st) mbar_wait(&empty_bar[st], ((uses >> st) & 1u) ^ 1u);
        stage = 0;
        for (int64_t c = c0; c < c1; ++c) {
          mbar_wait(&empty_bar[stage], ((uses >> stage) & 1u) ^ 1u);
          uses ^= 1u << stage;
          uint8_t* s0 = smem + (size_t)stage * boxes * kBoxBytes;
          mbar_expect_tx(&full_bar[stage], (a_boxes + g.x_cnt) * kBoxBytes);
          const int64_t tile = c >> 1;
          const size_t hoff = (size_t)(c & 1) * kBoxBytes;
          for (int b = 0; b < a_boxes; ++b)
            EO_DW_LOAD(s0 + b * kBoxBytes, g.G + ((size_t)tile * g.g_nb + g.g_blk0 + b) * kBlkBytes + hoff, kBoxBytes, &full_bar[stage]);
          for (int b = 0; b < g.x_cnt; ++b)
            EO_DW_LOAD(s0 + (a_boxes + b) * kBoxBytes, g.X + ((size_t)tile * g.x_nb + g.x_blk0 + b) * kBlkBytes + hoff, kBoxBytes, &full_bar[stage]);
          if (++stage == n_st) stage = 0;
        }
      }
#ifdef EONERF_TIMING
      if (blockIdx.x < 256) g_tnb_time[1][blockIdx.x] = gtimer();
#endif
    }
  } else if (warp == 1) {
    // whole warp converged, MMAs predicated on one elected lane: keeps the descriptors in uniform registers (tc_ptx.cuh)
    const bool elected = elect_one_sync();
    int stage = 0; uint32_t uses = 0;                        // bit s = how often stage s has been consumed, mod 2
    uint32_t items = 0;
    for (int gi = 0; gi < p.n; ++gi) {
      const TNBGemm& g = p.g[gi];
      int64_t c0, c1;
      tnb_range(g, chunks, lo, hi, c0, c1);
      if (c1 <= c0) continue;
      const uint32_t idesc = instr_desc(kBlockM, g.x_cnt * 64, 1, 1);
      const int a_boxes = g.mt_count * 2;
      const int boxes = a_boxes + g.x_cnt, n_st = tnb_stages(boxes);
      stage = 0;
      if (items > 0) {                                           // the epilogue has drained the previous item's accumulators
        mbar_wait(&acc_empty, (items - 1) & 1u);
        tc_fence_after();
      }
      for (int64_t c = c0; c < c1; ++c) {
        mbar_wait(&full_bar[stage], (uses >> stage) & 1u);
        uses ^= 1u << stage;
        tc_fence_after();
        const uint32_t s0 = smem_u32(smem + (size_t)stage * boxes * kBoxBytes);
        const uint32_t sx = s0 + a_boxes * kBoxBytes;
        if (elected) {
          if (!(p.dbg_skip_tail & 4))
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            const uint64_t dx = smem_desc(sx + k * 2048, kBoxBytes, 1024);
            umma_bf16(tmem_base, smem_desc(s0 + k * 2048, kBoxBytes, 1024), dx, idesc, (c > c0) || k != 0);
            if (g.mt_count == 2) umma_bf16(tmem_base + 256, smem_desc(s0 + 2 * kBoxBytes + k * 2048, kBoxBytes, 1024), dx, idesc, (c > c0) || k != 0);
          }
          if (p.dbg_skip_tail & 8) mbar_arrive(&empty_bar[stage]); else
          umma_commit(&empty_bar[stage]);
        }
        __syncwarp();
        if (++stage == n_st) stage = 0;
      }
      if (elected) umma_commit(&acc_full);
      __syncwarp();
      ++items;
    }
  } else {
    // ===== epilogue warps: bias-gradient side job during the main loop, then TMEM -> red.global =====
    const int t = threadIdx.x - 64;                       // 0..127: box t/32, 32-bit word `lane` of each 128-byte row
    const int box = t >> 5;
    const int quarter = warp & 3;
    int stage = 0; uint32_t uses = 0;
    uint32_t items = 0;
    for (int gi = 0; gi < p.n; ++gi) {
      const TNBGemm& g = p.g[gi];
      int64_t c0, c1;
      tnb_range(g, chunks, lo, hi, c0, c1);
      if (c1 <= c0) continue;
      const int a_boxes = g.mt_count * 2;
      const int boxes = a_boxes + g.x_cnt, n_st = tnb_stages(boxes);
      stage = 0;
      const int block_k = g.x_cnt * 64;
      const bool do_sum = box < a_boxes && g.db[box >> 1] != nullptr;
      // bias gradient: lane (ro = lane >> 3, lc = lane & 7) sums the eight features of 16-byte chunk lc over rows ro, ro + 4, ...
      // (one LDS.128 per four rows of the box, eight independent accumulators); the four row groups are merged at the end of the item
      float bs[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) bs[e] = 0.f;
      const int lc = lane & 7, ro = lane >> 3;
      for (int64_t c = c0; c < c1 && !(p.dbg_skip_tail & 16); ++c) {
        mbar_wait(&full_bar[stage], (uses >> stage) & 1u);
        uses ^= 1u << stage;
        if (do_sum) {
          const uint32_t base = smem_u32(smem + ((size_t)stage * boxes + box) * kBoxBytes);
#pragma unroll 4
          for (int i = 0; i < 16; ++i) {
            const int rr = 4 * i + ro;
            uint32_t w0, w1, w2, w3;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3) : "r"(base + rr * 128 + ((lc ^ (rr & 7)) << 4)));
            bs[0] += __uint_as_float(w0 << 16); bs[1] += __uint_as_float(w0 & 0xFFFF0000u);
            bs[2] += __uint_as_float(w1 << 16); bs[3] += __uint_as_float(w1 & 0xFFFF0000u);
            bs[4] += __uint_as_float(w2 << 16); bs[5] += __uint_as_float(w2 & 0xFFFF0000u);
            bs[6] += __uint_as_float(w3 << 16); bs[7] += __uint_as_float(w3 & 0xFFFF0000u);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[stage]);
        if (++stage == n_st) stage = 0;
      }
      if (do_sum) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          bs[e] += __shfl_xor_sync(0xffffffffu, bs[e], 8);
          bs[e] += __shfl_xor_sync(0xffffffffu, bs[e], 16);
        }
        if (ro == 0) {
          const int mt = box >> 1;
          const int n0 = (box & 1) * 64 + lc * 8;
#pragma unroll
          for (int e = 0; e < 8; ++e)
            if (n0 + e < g.n_valid[mt]) atomicAdd(g.db[mt] + n0 + e, bs[e]);
        }
      }
      mbar_wait(&acc_full, items & 1u);
      tc_fence_after();
      for (int mt = 0; mt < g.mt_count; ++mt) {
        const int n = quarter * 32 + lane;                 // output row (feature of G) inside this 128-row block
        const uint32_t taddr = tmem_base + mt * 256 + ((uint32_t)(quarter * 32) << 16);
        const bool vec = (g.ldd[mt] & 3) == 0 && ((uintptr_t)g.D[mt] & 15) == 0;
        for (int c = 0; c < block_k; c += 32) {
          __syncwarp();
          uint32_t r[32];
          tmem_ld32(taddr + c, r);
          tmem_ld_wait();
          if (n >= g.n_valid[mt] || g.D[mt] == nullptr || (p.dbg_skip_tail & 1)) continue;
          float* drow = g.D[mt] + (int64_t)n * g.ldd[mt] + c;
          if (vec && c + 32 <= g.k_valid) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(drow + j), "f"(__uint_as_float(r[j])), "f"(__uint_as_float(r[j + 1])),
                           "f"(__uint_as_float(r[j + 2])), "f"(__uint_as_float(r[j + 3]))
                           : "memory");
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (c + j < g.k_valid) atomicAdd(drow + j, __uint_as_float(r[j]));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty);
      ++items;
    }
  }
  tc_fence_before();
  __syncthreads();
#ifdef EONERF_TIMING
  if (threadIdx.x == 0 && blockIdx.x < 256) g_tnb_time[2][blockIdx.x] = gtimer();
#endif
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// Cost of one 64-sample chunk in 1/16-box units.  Bytes (8 KB boxes) are the first-order cost; the measured per-box time of a CTA
// (tools/dw_cta_times.py, 1 M samples, after the variable stage count) deviates from it by GEMM shape: 4-box chunks (six stages, two
// bias boxes) 0.93, 5-box chunks (four stages = 160 KB in flight) 1.02, 5-box chunks that also carry the bias-gradient side job on
// four G boxes 1.14, 8-box chunks 1.00.  Weighting the split by it lets all CTAs finish together.
static int tnb_chunk_cost(const GemmTNBlocked& g) {
  const int boxes = 2 * g.mt_count + g.x_cnt;
  const bool bias = g.db[0] != nullptr || g.db[1] != nullptr;
  double w = 1.0;
  if (boxes <= 4) w = 0.93;
  else if (boxes < 8) w = bias ? 1.14 : 1.02;
  return (int)(16.0 * boxes * w + 0.5);
}

int gemm_tn_blocked_group(const GemmTNBlocked* list, int n, cudaStream_t s) {
  static PerDeviceOnce once;
  if (once()) {
    EO_CUDA(cudaFuncSetAttribute(gemm_tn_blocked_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTNB));
  }
  static int dbg = -1;
  if (dbg < 0) { const char* e = getenv("EONERF_TN_DBG"); dbg = e ? atoi(e) : 0; }
  for (int i0 = 0; i0 < n; i0 += kTNGroupMax) {
    TNBParams p{};
    double flops = 0, bytes = 0;
    int64_t per = 0, chunks_total = 0, n_tiles = -1;
    const int64_t* n_dev = nullptr;
    for (int i = i0; i < n && i < i0 + kTNGroupMax; ++i) {
      const GemmTNBlocked& g = list[i];
      if (g.n_tiles <= 0) continue;
      EO_REQUIRE(g.mt_count >= 1 && g.mt_count <= 2 && g.x_cnt >= 1 && g.x_cnt <= 4, "gemm_tn_blocked: unsupported shape");
      TNBGemm& q = p.g[p.n++];
      q.G = g.G; q.X = g.X; q.g_nb = g.g_nb; q.g_blk0 = g.g_blk0; q.mt_count = g.mt_count;
      q.x_nb = g.x_nb; q.x_blk0 = g.x_blk0; q.x_cnt = g.x_cnt; q.k_valid = g.k_valid;
      for (int j = 0; j < 2; ++j) { q.n_valid[j] = g.n_valid[j]; q.D[j] = g.D[j]; q.ldd[j] = g.ldd[j]; q.db[j] = (dbg & 2) ? nullptr : g.db[j]; }
      EO_REQUIRE(n_tiles < 0 || (n_tiles == g.n_tiles && n_dev == g.n_pts_dev), "gemm_tn_blocked: the GEMMs of a group share the sample axis");
      n_tiles = g.n_tiles; n_dev = g.n_pts_dev;
      q.cost = tnb_chunk_cost(g);
      q.per0 = per;
      per += q.cost;
      chunks_total += g.n_tiles * 2;
      const double M = (double)g.n_tiles * kTileM;
      flops += 2.0 * M * g.mt_count * 128 * g.k_valid;
      bytes += 2.0 * M * (g.mt_count * 128 + g.x_cnt * 64);
    }
    if (p.n == 0) continue;
    p.chunks = n_tiles * 2;
    p.per_total = per;
    p.n_pts_dev = n_dev;
    p.dbg_skip_tail = dbg & 29;
    int64_t grid = sm_count();
    if (grid > chunks_total) grid = chunks_total;
    profile_begin(1, flops, bytes, s);
    gemm_tn_blocked_kernel<<<(unsigned)grid, kThreads, kSmemTNB, s>>>(p);
    profile_end(s);
    EO_LAUNCH_CHECK();
  }
  return EONERF_OK;
}

int gemm_tn_blocked(const GemmTNBlocked& g, cudaStream_t s) { return gemm_tn_blocked_group(&g, 1, s); }

}  // namespace eonerf

#ifdef EONERF_TIMING
extern "C" int eonerf_debug_tnb_time(unsigned long long* out) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, eonerf::g_tnb_time, sizeof(unsigned long long) * 3 * 256);
  return 0;
}
#endif
