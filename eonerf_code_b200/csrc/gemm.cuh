// Internal GEMM interface shared by the SIMT kernels (gemm_simt.cu: fp32 exactness mode, and a bf16-operand
// twin used as the on-device cross-check of the tensor-core path) and the bf16 tcgen05/TMEM/TMA kernels
// (gemm_tc.cu).  Both families implement the two contractions the radiance-field MLP needs
// (/root/reference/radiance_fields/mlp.py:87-101 forward, its autograd backward):
//
//   NT:  C[M,N] = epi( A[M,K] * B[N,K]^T )           forward layer (B = W) and dX (B = W^T)
//          v  = acc + (class_bias ? class_bias[row_class[m], n] : bias ? bias[n] : 0)
//                   + (addend ? addend[m,n] : 0) + (rank1_row ? rank1_row[m*rank1_stride] * rank1_col[n] : 0)
//          v  = relu ? max(v,0) : v
//          v  = (mask && n < mask_cols && !(mask[m,n] > 0)) ? 0 : v        (ReLU backward: threshold at 0)
//   TN:  D[N,K] += A[M,N]^T * X[M,K],  dbias[N] += colsum(A)               parameter gradients (fp32 output)
//
// A, B, C, X, addend, mask share one element type: float or __nv_bfloat16.  Accumulation is fp32.
#pragma once
#include "common.cuh"

namespace eonerf {

struct GemmNT {
  const void* A = nullptr; int64_t lda = 0;
  const void* B = nullptr; int64_t ldb = 0;
  void* C = nullptr; int64_t ldc = 0;
  int64_t M = 0; int N = 0; int K = 0;
  int alg_k = 0;   // un-padded contraction length (0: = K), only used to count algorithmic FLOPs
  const float* bias = nullptr;
  const int32_t* row_class = nullptr; const float* class_bias = nullptr; int64_t ld_class = 0;
  const void* addend = nullptr; int64_t ld_add = 0;
  const float* rank1_row = nullptr; int64_t rank1_stride = 1; const float* rank1_col = nullptr;
  int relu = 0;
  const void* mask = nullptr; int64_t ld_mask = 0; int mask_cols = 0;
};

struct GemmTN {
  const void* A = nullptr; int64_t lda = 0;
  const void* X = nullptr; int64_t ldx = 0;
  int64_t M = 0; int N = 0; int K = 0;
  float* D = nullptr; int64_t ldd = 0;
  float* dbias = nullptr;
};

// element type tags
enum ElemType { kF32 = 0, kBF16 = 1 };

int gemm_nt_simt(ElemType t, const GemmNT& g, cudaStream_t s);
int gemm_tn_simt(ElemType t, const GemmTN& g, cudaStream_t s);
// tensor cores (bf16 only).  Constraints: K % 16 == 0... checked inside, see gemm_tc.cu
int gemm_nt_tc(const GemmNT& g, cudaStream_t s);
int gemm_tn_tc(const GemmTN& g, cudaStream_t s);

// precision code of the C ABI -> which family runs
//   EONERF_PREC_FP32       fp32 storage, SIMT
//   EONERF_PREC_BF16       bf16 storage, tcgen05 tensor cores
//   EONERF_PREC_BF16_SIMT  bf16 storage, SIMT fp32-accumulate (cross-check of the tensor-core kernels)
static inline ElemType elem_of(int precision) { return precision == EONERF_PREC_FP32 ? kF32 : kBF16; }
static inline int elem_size(int precision) { return precision == EONERF_PREC_FP32 ? 4 : 2; }

static inline int gemm_nt(int precision, const GemmNT& g, cudaStream_t s) {
  if (precision == EONERF_PREC_BF16) return gemm_nt_tc(g, s);
  return gemm_nt_simt(elem_of(precision), g, s);
}
static inline int gemm_tn(int precision, const GemmTN& g, cudaStream_t s) {
  if (precision == EONERF_PREC_BF16) return gemm_tn_tc(g, s);
  return gemm_tn_simt(elem_of(precision), g, s);
}

template <class T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <class T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

}  // namespace eonerf
