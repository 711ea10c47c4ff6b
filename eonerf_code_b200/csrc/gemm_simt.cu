// SIMT GEMMs of the radiance-field MLP — fp32 storage (EONERF_PREC_FP32: the exactness mode, so that the
// whole rendering path can be compared with the fp32 reference at 1e-5 without bf16 rounding in the way)
// and bf16 storage with fp32 accumulation (EONERF_PREC_BF16_SIMT: the on-device cross-check of the tcgen05
// kernels in gemm_tc.cu, which are the production path).  Plain 64x64 register-tiled kernels.
#include "gemm.cuh"

namespace eonerf {

constexpr int BM = 64, BN = 64, BK = 16;

template <class T>
__global__ void __launch_bounds__(256) gemm_nt_simt_kernel(GemmNT g) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const T* A = (const T*)g.A;
  const T* B = (const T*)g.B;
  int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  int64_t m0 = (int64_t)blockIdx.x * BM;
  int n0 = blockIdx.y * BN;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < g.K; k0 += BK) {
#pragma unroll
    for (int l = 0; l < 4; ++l) {
      int e = tid + l * 256;            // 0..1023
      int row = e >> 4, kk = e & 15;
      int64_t m = m0 + row;
      int n = n0 + row, k = k0 + kk;
      As[kk][row] = (m < g.M && k < g.K) ? to_f32<T>(A[m * g.lda + k]) : 0.f;
      Bs[kk][row] = (n < g.N && k < g.K) ? to_f32<T>(B[(int64_t)n * g.ldb + k]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  T* C = (T*)g.C;
  const T* add = (const T*)g.addend;
  const T* mask = (const T*)g.mask;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int64_t m = m0 + ty * 4 + i;
    if (m >= g.M) continue;
    const float* brow = g.class_bias ? g.class_bias + (int64_t)__ldg(g.row_class + m) * g.ld_class : g.bias;
    float r1 = g.rank1_row ? __ldg(g.rank1_row + m * g.rank1_stride) : 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n >= g.N) continue;
      float v = acc[i][j];
      if (brow) v += __ldg(brow + n);
      if (add) v += to_f32<T>(add[m * g.ld_add + n]);
      if (g.rank1_row) v += r1 * __ldg(g.rank1_col + n);
      if (g.relu) v = fmaxf(v, 0.f);
      if (mask && n < g.mask_cols && !(to_f32<T>(mask[m * g.ld_mask + n]) > 0.f)) v = 0.f;
      C[m * g.ldc + n] = from_f32<T>(v);
    }
  }
}

// D[n,k] += sum_m A[m,n] X[m,k]; grid = (N tiles, K tiles, M splits); split-M partial sums are combined
// with fp32 atomics (order-dependent at the 1e-7 level, like the reference's own index_add_/cuBLAS split-K).
template <class T>
__global__ void __launch_bounds__(256) gemm_tn_simt_kernel(GemmTN g, int64_t rows_per_split) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Xs[BK][BN + 4];
  const T* A = (const T*)g.A;
  const T* X = (const T*)g.X;
  int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  int n0 = blockIdx.x * BM, k0 = blockIdx.y * BN;
  int64_t mb = (int64_t)blockIdx.z * rows_per_split;
  int64_t me = mb + rows_per_split < g.M ? mb + rows_per_split : g.M;
  float acc[4][4] = {};
  float bsum[4] = {};
  bool do_bias = g.dbias && blockIdx.y == 0 && tx == 0;
  for (int64_t m0 = mb; m0 < me; m0 += BK) {
#pragma unroll
    for (int l = 0; l < 4; ++l) {
      int e = tid + l * 256;
      int mm = e >> 6, c = e & 63;
      int64_t m = m0 + mm;
      As[mm][c] = (m < me && n0 + c < g.N) ? to_f32<T>(A[m * g.lda + n0 + c]) : 0.f;
      Xs[mm][c] = (m < me && k0 + c < g.K) ? to_f32<T>(X[m * g.ldx + k0 + c]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int mm = 0; mm < BK; ++mm) {
      float a[4], x[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[mm][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) x[j] = Xs[mm][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (do_bias) bsum[i] += a[i];
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], x[j], acc[i][j]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int n = n0 + ty * 4 + i;
    if (n >= g.N) continue;
    if (do_bias) atomicAdd(g.dbias + n, bsum[i]);
    if (!g.D) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int k = k0 + tx * 4 + j;
      if (k < g.K) atomicAdd(g.D + (int64_t)n * g.ldd + k, acc[i][j]);
    }
  }
}

int gemm_nt_simt(ElemType t, const GemmNT& g, cudaStream_t s) {
  if (g.M <= 0 || g.N <= 0) return EONERF_OK;
  dim3 grid(div_up(g.M, BM), div_up(g.N, BN));
  profile_begin(2, 2.0 * g.M * g.N * (g.alg_k ? g.alg_k : g.K), 0.0, s);
  if (t == kF32) gemm_nt_simt_kernel<float><<<grid, 256, 0, s>>>(g);
  else gemm_nt_simt_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(g);
  profile_end(s);
  EO_LAUNCH_CHECK();
  return EONERF_OK;
}

int gemm_tn_simt(ElemType t, const GemmTN& g, cudaStream_t s) {
  if (g.M <= 0 || g.N <= 0 || g.K <= 0) return EONERF_OK;
  int tn = div_up(g.N, BM), tk = div_up(g.K, BN);
  int64_t want = (2 * 148 + tn * tk - 1) / (tn * tk);
  int64_t max_split = (g.M + 255) / 256;
  int64_t split = want < max_split ? want : max_split;
  if (split < 1) split = 1;
  int64_t rows = (g.M + split - 1) / split;
  rows = (rows + BK - 1) / BK * BK;
  split = (g.M + rows - 1) / rows;
  dim3 grid(tn, tk, (unsigned)split);
  profile_begin(2, 2.0 * g.M * g.N * g.K, 0.0, s);
  if (t == kF32) gemm_tn_simt_kernel<float><<<grid, 256, 0, s>>>(g, rows);
  else gemm_tn_simt_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(g, rows);
  profile_end(s);
  EO_LAUNCH_CHECK();
  return EONERF_OK;
}

}  // namespace eonerf
