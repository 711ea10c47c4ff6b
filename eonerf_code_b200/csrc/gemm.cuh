// Internal GEMM interface shared by the fp32 SIMT kernels (gemm_simt.cu, exactness mode) and the
// bf16 tcgen05/TMEM/TMA kernels (gemm_tc.cu).  Both families implement the same two contractions the
// radiance-field MLP needs (radiance_fields/mlp.py:87-101 forward, its autograd backward):
//
//   NT:  C[M,N] = epi( A[M,K] * B[N,K]^T )           forward layer (B = W) and dX (B = W^T)
//          epi(v) = [relu]( v + bias[n] | class_bias[row_class[m], n]  + addend[m,n] ) * [mask[m,n] > 0 for n < mask_cols]
//   TN:  D[N,K] += A[M,N]^T * X[M,K],  dbias[N] += colsum(A)      parameter gradients (fp32 output)
//
// A, B, C, X, addend, mask share one element type: float (SIMT) or __nv_bfloat16 (tensor cores).
#pragma once
#include "common.cuh"

namespace eonerf {

struct GemmNT {
  const void* A = nullptr; int64_t lda = 0;
  const void* B = nullptr; int64_t ldb = 0;
  void* C = nullptr; int64_t ldc = 0;
  int64_t M = 0; int N = 0; int K = 0;
  const float* bias = nullptr;
  const int32_t* row_class = nullptr; const float* class_bias = nullptr;
  const void* addend = nullptr; int64_t ld_add = 0;
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

int gemm_nt_f32(const GemmNT& g, cudaStream_t s);
int gemm_tn_f32(const GemmTN& g, cudaStream_t s);
int gemm_nt_bf16(const GemmNT& g, cudaStream_t s);
int gemm_tn_bf16(const GemmTN& g, cudaStream_t s);

}  // namespace eonerf
