// Radiance-field MLP: EONerfMLP.forward / query_density (/root/reference/radiance_fields/eonerf.py:141-170),
// MLP.forward (radiance_fields/mlp.py:87-101), SinusoidalEncoder.forward (mlp.py:190-208) and
// VanillaNeRFRadianceField.forward (mlp.py:245-250) — forward and backward, layer by layer.
//
// Data layout in HBM (row-major, one row per sample, element type T = fp32 or bf16):
//   stash  xf   fp32 [N,3]     sample positions (pos-enc backward needs full precision)
//          cls  i32  [N]       image index per sample (class of the folded transient-embedding bias)
//          H0..H3    T [N,256] post-ReLU trunk activations
//          H4E       T [N,320] = [ h4 (256) | posenc(x) (63) | 0 ]   the skip concat of mlp.py:92-97 is a view:
//                               layer 0 reads columns 256:320, layer 5 reads all 320 (K = 320)
//          H5..H7    T [N,256]
//          BOTT      T [N,256] bottleneck (vanilla: [N,288] = [bottleneck | posenc4(viewdir) (27) | 0])
//          HD0       T [N,256] = [ albedo hidden (128) | transient hidden 0 (128) ]  (vanilla: [N,128])
//          T1..T3    T [N,128] transient hidden 1..3
//   The 4-d transient embedding (eonerf.py:165-167) never becomes a column: W[:,256:260]·emb[img] + b is a
//   per-image bias row of a [n_img,256] table (`class_bias`) selected by `cls` in the GEMM epilogue.
//
// Every dense contraction goes through gemm_nt / gemm_tn (gemm.cuh): tcgen05 tensor cores in bf16 mode.
// The 1- and 3-wide heads (sigma, albedo, transient scalar/beta) are warp-per-row fp32 dot products.
#include <type_traits>

#include "field_layout.cuh"

namespace eonerf {

// ------------------------------------------------------------------------------------------------
// prepare: W fp32 [rows, k] (ld = ldw) -> dst [rows, kp] (zero padded) and dst_t [kp, rows]
// ------------------------------------------------------------------------------------------------
// All weight matrices of a prepare call go through ONE launch: a table of jobs, each thread finds its job from the prefix of
// element counts (14 launches -> 1 per optimiser step).
struct ConvertJob { const float* w; int64_t ldw; int rows, k, kp; void* dst; int dst_row0; void* dst_t; int ld_t; int first; };
constexpr int kMaxConvertJobs = 16;
struct ConvertJobs { ConvertJob j[kMaxConvertJobs]; int n; int total; };

template <class T>
__global__ void __launch_bounds__(256) convert_weights_kernel(const __grid_constant__ ConvertJobs jobs) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= jobs.total) return;
  int ji = 0;
#pragma unroll 1
  while (ji + 1 < jobs.n && idx >= jobs.j[ji + 1].first) ++ji;
  const ConvertJob& q = jobs.j[ji];
  const int e = idx - q.first;
  const int r = e / q.kp, c = e % q.kp;
  const float v = c < q.k ? __ldg(q.w + (int64_t)r * q.ldw + c) : 0.f;
  static_cast<T*>(q.dst)[(int64_t)(q.dst_row0 + r) * q.kp + c] = from_f32<T>(v);
  static_cast<T*>(q.dst_t)[(int64_t)c * q.ld_t + q.dst_row0 + r] = from_f32<T>(v);
}

static thread_local ConvertJobs t_jobs;          // filled by convert_weight(), launched by convert_flush()

static int convert_flush(int precision, cudaStream_t s) {
  if (t_jobs.n == 0) return EONERF_OK;
  const int blocks = div_up(t_jobs.total, 256);
  if (precision == EONERF_PREC_FP32) convert_weights_kernel<float><<<blocks, 256, 0, s>>>(t_jobs);
  else convert_weights_kernel<__nv_bfloat16><<<blocks, 256, 0, s>>>(t_jobs);
  t_jobs.n = 0;
  t_jobs.total = 0;
  EO_LAUNCH_CHECK();
  return EONERF_OK;
}

static int convert_weight(int precision, const float* w, int64_t ldw, int rows, int k, int kp, void* dst, int dst_row0,
                          void* dst_t, int ld_t, cudaStream_t s) {
  if (t_jobs.n == kMaxConvertJobs) {
    int rc = convert_flush(precision, s);
    if (rc != EONERF_OK) return rc;
  }
  t_jobs.j[t_jobs.n] = ConvertJob{w, ldw, rows, k, kp, dst, dst_row0, dst_t, ld_t, t_jobs.total};
  t_jobs.n += 1;
  t_jobs.total += rows * kp;
  return EONERF_OK;
}

// class_bias[img, 0:128] = albedo_mlp.0.bias ; class_bias[img, 128:256] = transient_mlp.0.bias + W[:,256:260] emb[img]
__global__ void class_bias_kernel(const float* __restrict__ ba0, const float* __restrict__ wt0, const float* __restrict__ bt0,
                                  const float* __restrict__ emb, int64_t n_images, float* __restrict__ cb) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_images * 2 * kHid) return;
  int img = idx / (2 * kHid), j = idx % (2 * kHid);
  float v;
  if (j < kHid) {
    v = __ldg(ba0 + j);
  } else {
    int r = j - kHid;
    v = __ldg(bt0 + r);
#pragma unroll
    for (int e = 0; e < 4; ++e) v = fmaf(__ldg(wt0 + r * 260 + 256 + e), __ldg(emb + img * 4 + e), v);
  }
  cb[idx] = v;
}

// ------------------------------------------------------------------------------------------------
// encode: positions (optionally derived from rays), image class, pos-enc columns of H4E
// ------------------------------------------------------------------------------------------------
template <class T>
__global__ void __launch_bounds__(256) encode_kernel(EonerfFieldFwdArgs a, float* __restrict__ xf, int32_t* __restrict__ cls,
                                                     T* __restrict__ h4e) {
  // 64 consecutive threads own one sample: thread c writes pos-enc column c (coalesced 64*sizeof(T) bytes)
  int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t p = gid >> 6;
  int c = (int)(gid & 63);
  if (p >= a.n_pts) return;
  float x[3];
  int64_t ray = a.ray_indices ? __ldg(a.ray_indices + p) : -1;
  if (a.x) {
    x[0] = __ldg(a.x + 3 * p); x[1] = __ldg(a.x + 3 * p + 1); x[2] = __ldg(a.x + 3 * p + 2);
  } else {
    float ts = __ldg(a.t_starts + p), te = __ldg(a.t_ends + p);
    float zm = __fdiv_rn(__fadd_rn(ts, te), 2.0f);                       // eonerf.py:206
    const float* o = a.origins + ray * a.origins_stride;
    const float* d = a.viewdirs + ray * a.viewdirs_stride;
#pragma unroll
    for (int k = 0; k < 3; ++k) x[k] = __fadd_rn(__ldg(o + k), __fmul_rn(__ldg(d + k), zm));   // eonerf.py:207
    if (c == 0 && a.z_mid) a.z_mid[p] = zm;
  }
  if (c < 3) xf[3 * p + c] = x[c];
  if (c == 3 && cls) {
    int64_t img = 0;
    if (a.img_idx) img = a.ray_indices ? __ldg(a.img_idx + ray * a.img_idx_stride) : __ldg(a.img_idx + p * a.img_idx_stride);
    cls[p] = (int32_t)img;
  }
  // mlp.py:199-205: [x, sin(2^k x) (freq-major, xyz-minor) k=0..9, sin(2^k x + pi/2)]
  float v;
  if (c < 3) v = x[c];
  else if (c < 63) {
    int e = c - 3;
    int half = e >= 30;
    e -= half * 30;
    float xb = x[e % 3] * (float)(1 << (e / 3));
    v = sinf(half ? __fadd_rn(xb, kHalfPi) : xb);
  } else v = 0.f;
  h4e[p * kH4E + kW + c] = from_f32<T>(v);
}

// view-direction encoding (L=4) -> BOTT[:, 256:288]  (vanilla field, mlp.py:153-165)
// rays != NULL: dirs holds one row per RAY (looked up through rays[p]); else one row per sample
template <class T>
__global__ void __launch_bounds__(256) encode_dirs_kernel(const float* __restrict__ dirs, int64_t stride, int64_t n,
                                                          T* __restrict__ dst, int64_t ld, int col0, const int64_t* __restrict__ rays) {
  int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t p = gid >> 5;
  int c = (int)(gid & 31);
  if (p >= n) return;
  const int64_t row = rays ? __ldg(rays + p) : p;
  float v = 0.f;
  if (c < 3) v = __ldg(dirs + row * stride + c);
  else if (c < 27) {
    int e = c - 3;
    int half = e >= 12;
    e -= half * 12;
    float xb = __ldg(dirs + row * stride + e % 3) * (float)(1 << (e / 3));
    v = sinf(half ? __fadd_rn(xb, kHalfPi) : xb);
  }
  dst[p * ld + col0 + c] = from_f32<T>(v);
}

// g_x[c] = g_enc[c] + sum_k 2^k ( cos(2^k x_c) g_enc[3+3k+c] + cos(2^k x_c + pi/2) g_enc[33+3k+c] )
// g_enc = ge0 (layer-0 input gradient) + ge5 (columns 256:319 of the layer-5 input gradient)
template <class T>
__global__ void __launch_bounds__(256) posenc_bwd_kernel(const float* __restrict__ xf, const T* __restrict__ ge0, int64_t ld0,
                                                         const T* __restrict__ ge5, int64_t ld5, int64_t n,
                                                         float* __restrict__ g_x) {
  int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t p = gid / 3;
  int c = (int)(gid % 3);
  if (p >= n) return;
  float x = __ldg(xf + 3 * p + c);
  auto ge = [&](int col) { return to_f32<T>(ge0[p * ld0 + col]) + to_f32<T>(ge5[p * ld5 + col]); };
  float g = ge(c);
#pragma unroll
  for (int k = 0; k < 10; ++k) {
    float s = (float)(1 << k), xb = x * s;
    g += s * (cosf(xb) * ge(3 + 3 * k + c) + cosf(__fadd_rn(xb, kHalfPi)) * ge(33 + 3 * k + c));
  }
  g_x[3 * p + c] = g;
}

// ------------------------------------------------------------------------------------------------
// narrow heads: y_j = act_j( x . w_j + b_j ),  j < J <= 3      (warp per row, fp32 math)
// act: 0 identity, 1 relu, 2 sigmoid (nn.Sigmoid), 3 softplus (nn.Softplus beta=1 threshold=20)
// ------------------------------------------------------------------------------------------------
struct Head {
  const float* w; const float* b; int act;
  float* out; int64_t out_stride;          // fwd: out[m*out_stride]
  const float* y; const float* g;          // bwd: forward output and incoming gradient (same stride)
};
struct Heads {
  int J; Head h[3];
};

__device__ __forceinline__ float act_fwd(float v, int act) {
  switch (act) {
    case 1: return fmaxf(v, 0.f);
    case 2: return 1.0f / (1.0f + expf(-v));
    case 3: return v > 20.0f ? v : log1pf(expf(v));
    default: return v;
  }
}
// derivative expressed through the forward output y
__device__ __forceinline__ float act_bwd(float y, int act) {
  switch (act) {
    case 1: return y > 0.f ? 1.f : 0.f;
    case 2: return y * (1.0f - y);
    case 3: return -expm1f(-y);            // sigmoid(pre) = 1 - exp(-softplus(pre)); -> 1 above the threshold
    default: return 1.f;
  }
}

// 8 consecutive elements <-> float[8] (one 16-byte access for bf16, two for fp32)
template <class T> struct Vec8;
template <> struct Vec8<__nv_bfloat16> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&f)[8]) {
    const uint4 v = __ldg((const uint4*)p);
    const __nv_bfloat162* h = (const __nv_bfloat162*)&v;
#pragma unroll
    for (int j = 0; j < 4; ++j) { float2 x = __bfloat1622float2(h[j]); f[2 * j] = x.x; f[2 * j + 1] = x.y; }
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&f)[8]) {
    uint4 v;
    __nv_bfloat162* h = (__nv_bfloat162*)&v;
#pragma unroll
    for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
    *(uint4*)p = v;
  }
};
template <> struct Vec8<float> {
  static __device__ __forceinline__ void load(const float* p, float (&f)[8]) {
    const float4 a = __ldg((const float4*)p), b = __ldg((const float4*)(p + 4));
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&f)[8]) {
    *(float4*)p = make_float4(f[0], f[1], f[2], f[3]);
    *(float4*)(p + 4) = make_float4(f[4], f[5], f[6], f[7]);
  }
};

// The narrow heads read rows of K = 128 or 256 activations.  K/8 lanes own one row (each lane 8 consecutive columns, one
// vector load); the head weights for those 8 columns stay in registers while the warp walks over its rows.
template <class T, int J>
__global__ void __launch_bounds__(256) heads_fwd_kernel(const T* __restrict__ x, int64_t ldx, int K, int64_t M, Heads H) {
  const int lpr = K >> 3;                       // lanes per row: 16 or 32
  const int lane = threadIdx.x & 31;
  const int sub = lane % lpr, rsub = lane / lpr, rpw = 32 / lpr;
  float w[J][8];
#pragma unroll
  for (int j = 0; j < J; ++j)
#pragma unroll
    for (int e = 0; e < 8; ++e) w[j][e] = __ldg(H.h[j].w + sub * 8 + e);
  const int64_t warp_id = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5), n_warps = (int64_t)gridDim.x * 8;
  for (int64_t m0 = warp_id * rpw; m0 < M; m0 += n_warps * rpw) {
    const int64_t m = m0 + rsub;
    float acc[J];
#pragma unroll
    for (int j = 0; j < J; ++j) acc[j] = 0.f;
    if (m < M) {
      float xv[8];
      Vec8<T>::load(x + m * ldx + sub * 8, xv);
#pragma unroll
      for (int j = 0; j < J; ++j)
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[j] = fmaf(xv[e], w[j][e], acc[j]);
    }
#pragma unroll
    for (int j = 0; j < J; ++j) {
      float v = acc[j];
      for (int o = lpr >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
      if (sub == 0 && m < M) H.h[j].out[m * H.h[j].out_stride] = act_fwd(v + __ldg(H.h[j].b), H.h[j].act);
    }
  }
}

// dpre[m, col0+j] = g_j * act'(y_j);  optionally dx[m,k] = (sum_j dpre_j w_j[k]) * (x[m,k] > 0)
template <class T, int J>
__global__ void __launch_bounds__(256) heads_bwd_kernel(const T* __restrict__ x, int64_t ldx, int K, int64_t M, Heads H,
                                                        float* __restrict__ dpre, int col0, T* __restrict__ dx, int64_t lddx) {
  const int lpr = K >> 3;
  const int lane = threadIdx.x & 31;
  const int sub = lane % lpr, rsub = lane / lpr, rpw = 32 / lpr;
  float w[J][8];
#pragma unroll
  for (int j = 0; j < J; ++j)
#pragma unroll
    for (int e = 0; e < 8; ++e) w[j][e] = __ldg(H.h[j].w + sub * 8 + e);
  const int64_t warp_id = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5), n_warps = (int64_t)gridDim.x * 8;
  for (int64_t m0 = warp_id * rpw; m0 < M; m0 += n_warps * rpw) {
    const int64_t m = m0 + rsub;
    if (m >= M) continue;
    float d[J];
#pragma unroll
    for (int j = 0; j < J; ++j) {
      float g = H.h[j].g ? __ldg(H.h[j].g + m * H.h[j].out_stride) : 0.f;
      d[j] = g * act_bwd(__ldg(H.h[j].y + m * H.h[j].out_stride), H.h[j].act);
      if (sub == 0) dpre[m * 8 + col0 + j] = d[j];
    }
    if (!dx) continue;
    float xv[8], o[8];
    Vec8<T>::load(x + m * ldx + sub * 8, xv);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float v = 0.f;
#pragma unroll
      for (int j = 0; j < J; ++j) v = fmaf(d[j], w[j][e], v);
      o[e] = xv[e] > 0.f ? v : 0.f;
    }
    Vec8<T>::store(dx + m * lddx + sub * 8, o);
  }
}

// dw_j[k] += sum_m dpre[m,col0+j] x[m,k];  db_j += sum_m dpre[m,col0+j]
// 256 threads = (256/lpr) rows x lpr column groups per sweep; 4 rows in flight per thread; block-level tree in smem.
struct HeadGrads {
  float* dw[3]; float* db[3];
};
template <class T, int J>
__global__ void __launch_bounds__(256) heads_dw_kernel(const T* __restrict__ x, int64_t ldx, int K, int64_t M,
                                                       const float* __restrict__ dpre, int col0, HeadGrads G, int64_t rows_per_block) {
  __shared__ float red[8][J][264];
  const int lpr = K >> 3, rows = 256 / lpr;
  const int sub = threadIdx.x % lpr, rsub = threadIdx.x / lpr;
  const int64_t m_begin = (int64_t)blockIdx.x * rows_per_block;
  const int64_t m_end = m_begin + rows_per_block < M ? m_begin + rows_per_block : M;
  float acc[J][8], bs[J];
#pragma unroll
  for (int j = 0; j < J; ++j) {
    bs[j] = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[j][e] = 0.f;
  }
  for (int64_t m0 = m_begin + rsub; m0 < m_end; m0 += 4 * rows) {
    float xv[4][8], d[4][J];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t m = m0 + (int64_t)u * rows;
      if (m < m_end) {
        Vec8<T>::load(x + m * ldx + sub * 8, xv[u]);
#pragma unroll
        for (int j = 0; j < J; ++j) d[u][j] = __ldg(dpre + m * 8 + col0 + j);
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) xv[u][e] = 0.f;
#pragma unroll
        for (int j = 0; j < J; ++j) d[u][j] = 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int j = 0; j < J; ++j) {
        bs[j] += d[u][j];
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[j][e] = fmaf(d[u][j], xv[u][e], acc[j][e]);
      }
  }
  // reduce over the row sub-groups: warps hold (32/lpr) row groups each -> shuffle first, then smem across the 8 warps
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lpr == 16) {
#pragma unroll
    for (int j = 0; j < J; ++j) {
      bs[j] += __shfl_xor_sync(kFull, bs[j], 16);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[j][e] += __shfl_xor_sync(kFull, acc[j][e], 16);
    }
  }
  if (lane < lpr) {
#pragma unroll
    for (int j = 0; j < J; ++j) {
#pragma unroll
      for (int e = 0; e < 8; ++e) red[warp][j][lane * 8 + e] = acc[j][e];
      if (lane == 0) red[warp][j][256] = bs[j];
    }
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < J * (K + 1); idx += 256) {
    const int j = idx / (K + 1), k = idx % (K + 1);
    const int col = k < K ? k : 256;
    float v = 0.f;
#pragma unroll
    for (int wv = 0; wv < 8; ++wv) v += red[wv][j][col];
    if (k < K) atomicAdd(G.dw[j] + k, v); else atomicAdd(G.db[j], v);
  }
}

// C[m,n] = (row[m*stride] * col[n]) masked by mask[m,n] > 0    (density-only backward: d h7), 8 columns per thread
template <class T>
__global__ void __launch_bounds__(256) rank1_mask_kernel(const float* __restrict__ row, int64_t stride, const float* __restrict__ col,
                                                         const T* __restrict__ mask, int64_t ld_mask, int64_t M, int N,
                                                         T* __restrict__ C, int64_t ldc) {
  const int per_row = N >> 3;
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t m = gid / per_row;
  const int n = (int)(gid % per_row) * 8;
  if (m >= M) return;
  const float r = __ldg(row + m * stride);
  float mk[8], o[8];
  Vec8<T>::load(mask + m * ld_mask + n, mk);
#pragma unroll
  for (int e = 0; e < 8; ++e) o[e] = mk[e] > 0.f ? r * __ldg(col + n + e) : 0.f;
  Vec8<T>::store(C + m * ldc + n, o);
}

// dcb[img, j] += sum_{rows of class img} g[row, j]  (j < 128).  16 lanes x 8 columns own a row; a thread keeps a running
// sum while consecutive rows stay in the same class (samples are packed ray by ray, a ray has one image index) and
// flushes it into the block-private shared table (or straight to global for very many images) when the class changes.
template <class T>
__global__ void __launch_bounds__(256) class_grad_kernel(const T* __restrict__ g, int64_t ldg, const int32_t* __restrict__ cls,
                                                         int64_t M, int64_t n_images, float* __restrict__ dcb, int64_t rows_per_block,
                                                         int use_smem) {
  extern __shared__ float tab[];
  const int sub = threadIdx.x & 15, rsub = threadIdx.x >> 4;       // 16 rows per sweep
  const int64_t m_begin = (int64_t)blockIdx.x * rows_per_block;
  const int64_t m_end = m_begin + rows_per_block < M ? m_begin + rows_per_block : M;
  if (use_smem) {
    for (int64_t i = threadIdx.x; i < n_images * kHid; i += 256) tab[i] = 0.f;
    __syncthreads();
  }
  float* dst = use_smem ? tab : dcb;
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  int cur = -1;
  for (int64_t m = m_begin + rsub; m < m_end; m += 16) {
    const int c = __ldg(cls + m);
    if (c != cur) {
      if (cur >= 0) {
#pragma unroll
        for (int e = 0; e < 8; ++e) { atomicAdd(dst + (int64_t)cur * kHid + sub * 8 + e, acc[e]); acc[e] = 0.f; }
      }
      cur = c;
    }
    float v[8];
    Vec8<T>::load(g + m * ldg + sub * 8, v);
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] += v[e];
  }
  if (cur >= 0) {
#pragma unroll
    for (int e = 0; e < 8; ++e) atomicAdd(dst + (int64_t)cur * kHid + sub * 8 + e, acc[e]);
  }
  if (use_smem) {
    __syncthreads();
    for (int64_t i = threadIdx.x; i < n_images * kHid; i += 256)
      if (tab[i] != 0.f) atomicAdd(dcb + i, tab[i]);
  }
}

// d emb[img,e] += sum_j dcb[img,j] W[j,256+e];   dW[j,256+e] += sum_img dcb[img,j] emb[img,e]
__global__ void emb_grad_kernel(const float* __restrict__ dcb, const float* __restrict__ wt0, const float* __restrict__ emb,
                                int64_t n_images, float* __restrict__ g_emb, float* __restrict__ g_wt0) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < n_images * 4) {
    int img = idx / 4, e = idx % 4;
    float v = 0.f;
    for (int j = 0; j < kHid; ++j) v = fmaf(__ldg(dcb + img * kHid + j), __ldg(wt0 + j * 260 + 256 + e), v);
    if (g_emb) g_emb[idx] += v;
  }
  if (idx < kHid * 4 && g_wt0) {
    int j = idx / 4, e = idx % 4;
    float v = 0.f;
    for (int64_t img = 0; img < n_images; ++img) v = fmaf(__ldg(dcb + img * kHid + j), __ldg(emb + img * 4 + e), v);
    g_wt0[j * 260 + 256 + e] += v;
  }
}

// ------------------------------------------------------------------------------------------------
// host-side helpers
// ------------------------------------------------------------------------------------------------
#define EO_TRY(expr)            \
  do {                          \
    int _r = (expr);            \
    if (_r != EONERF_OK) return _r; \
  } while (0)

template <class F>
static int by_type(int precision, F&& f) {
  if (precision == EONERF_PREC_FP32) return f((float*)nullptr);
  return f((__nv_bfloat16*)nullptr);
}

template <class F>
static int by_heads(int J, F&& f) {
  switch (J) {
    case 1: return f(std::integral_constant<int, 1>{});
    case 2: return f(std::integral_constant<int, 2>{});
    default: return f(std::integral_constant<int, 3>{});
  }
}

static int heads_grid(int64_t M, int K) {
  const int rpw = 32 / (K >> 3);
  int64_t blocks = (M + 8 * rpw - 1) / (8 * rpw);
  const int64_t cap = 148 * 8;
  return (int)(blocks < cap ? blocks : cap);
}

static int run_heads_fwd(int precision, const void* x, int64_t ldx, int K, int64_t M, const Heads& H, cudaStream_t s) {
  if (M == 0) return EONERF_OK;
  EO_REQUIRE(K == 128 || K == 256, "heads: K must be 128 or 256 (got %d)", K);
  return by_type(precision, [&](auto* tag) {
    using T = std::remove_pointer_t<decltype(tag)>;
    return by_heads(H.J, [&](auto j) {
      heads_fwd_kernel<T, decltype(j)::value><<<heads_grid(M, K), 256, 0, s>>>((const T*)x, ldx, K, M, H);
      EO_LAUNCH_CHECK();
      return EONERF_OK;
    });
  });
}

static int run_heads_bwd(int precision, const void* x, int64_t ldx, int K, int64_t M, const Heads& H, float* dpre, int col0,
                         void* dx, int64_t lddx, cudaStream_t s) {
  if (M == 0) return EONERF_OK;
  EO_REQUIRE(K == 128 || K == 256, "heads: K must be 128 or 256 (got %d)", K);
  return by_type(precision, [&](auto* tag) {
    using T = std::remove_pointer_t<decltype(tag)>;
    return by_heads(H.J, [&](auto j) {
      heads_bwd_kernel<T, decltype(j)::value><<<heads_grid(M, K), 256, 0, s>>>((const T*)x, ldx, K, M, H, dpre, col0, (T*)dx, lddx);
      EO_LAUNCH_CHECK();
      return EONERF_OK;
    });
  });
}

static int run_heads_dw(int precision, const void* x, int64_t ldx, int K, int64_t M, int J, const float* dpre, int col0,
                        const HeadGrads& G, cudaStream_t s) {
  if (M == 0) return EONERF_OK;
  EO_REQUIRE(K == 128 || K == 256, "heads: K must be 128 or 256 (got %d)", K);
  int64_t rows = 1024;
  while (div_up(M, rows) > 4 * 148) rows *= 2;
  return by_type(precision, [&](auto* tag) {
    using T = std::remove_pointer_t<decltype(tag)>;
    return by_heads(J, [&](auto j) {
      heads_dw_kernel<T, decltype(j)::value><<<div_up(M, rows), 256, 0, s>>>((const T*)x, ldx, K, M, dpre, col0, G, rows);
      EO_LAUNCH_CHECK();
      return EONERF_OK;
    });
  });
}

static inline char* at(const void* base, int64_t off) { return (char*)base + off; }

static int field_prepare_impl(int field, int precision, const EonerfFieldParams* p, void* prepared, cudaStream_t s) {
  PrepLayout L = prep_layout(field, precision, p->n_images);
  for (int i = 0; i < 8; ++i)
    EO_TRY(convert_weight(precision, p->trunk_w[i], trunk_k(i), kW, trunk_k(i), trunk_kp(i), at(prepared, L.w[i]), 0,
                          at(prepared, L.wt[i]), kW, s));
  EO_TRY(convert_weight(precision, p->bott_w, kW, kW, kW, kW, at(prepared, L.bott), 0, at(prepared, L.bott_t), kW, s));
  if (field == EONERF_FIELD_EONERF) {
    EO_TRY(convert_weight(precision, p->head0_w, kW, kHid, kW, kW, at(prepared, L.hd0), 0, at(prepared, L.hd0_t), 2 * kHid, s));
    EO_TRY(convert_weight(precision, p->trans_w[0], 260, kHid, kW, kW, at(prepared, L.hd0), kHid, at(prepared, L.hd0_t), 2 * kHid, s));
    for (int i = 0; i < 3; ++i)
      EO_TRY(convert_weight(precision, p->trans_w[i + 1], kHid, kHid, kHid, kHid, at(prepared, L.tr[i]), 0,
                            at(prepared, L.tr_t[i]), kHid, s));
    int n = (int)(p->n_images * 2 * kHid);
    class_bias_kernel<<<div_up(n, 256), 256, 0, s>>>(p->head0_b, p->trans_w[0], p->trans_b[0], p->transient_emb, p->n_images,
                                                      (float*)at(prepared, L.class_bias));
    EO_LAUNCH_CHECK();
  } else {
    EO_TRY(convert_weight(precision, p->head0_w, 283, kHid, 283, kW + kDirEnc, at(prepared, L.hd0), 0, at(prepared, L.hd0_t), kHid, s));
  }
  return convert_flush(precision, s);
}

static int field_fwd_impl(const EonerfFieldFwdArgs* a, cudaStream_t s) {
  const int prec = a->precision, field = a->field;
  const int64_t N = a->n_pts;
  if (N == 0) return EONERF_OK;
  const EonerfFieldParams* p = a->params;
  StashLayout S = stash_layout(field, prec, N, a->density_only);
  PrepLayout W = prep_layout(field, prec, p->n_images);
  const int es = elem_size(prec);
  char* st = (char*)a->stash;
  const char* pr = (const char*)a->prepared;
  bool want_cls = field == EONERF_FIELD_EONERF && !a->density_only;

  EO_TRY(by_type(prec, [&](auto* tag) {
    using T = std::remove_pointer_t<decltype(tag)>;
    encode_kernel<T><<<div_up(N * 64, 256), 256, 0, s>>>(*a, (float*)(st + S.xf), want_cls ? (int32_t*)(st + S.cls) : nullptr,
                                                          (T*)(st + S.h[4]));
    EO_LAUNCH_CHECK();
    return EONERF_OK;
  }));

  // trunk: base_mlp (mlp.py:87-101), skip concat after layer index 4
  for (int i = 0; i < 8; ++i) {
    GemmNT g;
    g.M = N; g.N = kW; g.K = trunk_kp(i);
    if (i == 0) { g.A = st + S.h[4] + (int64_t)kW * es; g.lda = kH4E; }
    else { g.A = st + S.h[i - 1]; g.lda = (i == 5) ? kH4E : kW; }
    g.B = pr + W.w[i]; g.ldb = trunk_kp(i);
    g.C = st + S.h[i]; g.ldc = (i == 4) ? kH4E : kW;
    g.bias = p->trunk_b[i]; g.relu = 1; g.alg_k = trunk_k(i);
    EO_TRY(gemm_nt(prec, g, s));
  }
  {  // sigma head: softplus (eonerf.py:106,145) / relu (vanilla, mlp.py:250)
    Heads H{};
    H.J = 1;
    H.h[0] = Head{p->sigma_w, p->sigma_b, field == EONERF_FIELD_EONERF ? 3 : 1, a->sigma, 1, nullptr, nullptr};
    EO_TRY(run_heads_fwd(prec, st + S.h[7], kW, kW, N, H, s));
  }
  if (a->density_only) return EONERF_OK;

  {  // bottleneck (no activation)
    GemmNT g;
    g.M = N; g.N = kW; g.K = kW;
    g.A = st + S.h[7]; g.lda = kW; g.B = pr + W.bott; g.ldb = kW;
    g.C = st + S.bott; g.ldc = S.ld_bott; g.bias = p->bott_b;
    EO_TRY(gemm_nt(prec, g, s));
  }
  if (field == EONERF_FIELD_EONERF) {
    {  // [albedo_mlp.0 | transient_mlp.0] in one GEMM; the embedding enters through the per-image bias row
      GemmNT g;
      g.M = N; g.N = 2 * kHid; g.K = kW;
      g.A = st + S.bott; g.lda = kW; g.B = pr + W.hd0; g.ldb = kW;
      g.C = st + S.hd0; g.ldc = 2 * kHid;
      g.row_class = (const int32_t*)(st + S.cls); g.class_bias = (const float*)(pr + W.class_bias); g.ld_class = 2 * kHid;
      g.relu = 1;
      EO_TRY(gemm_nt(prec, g, s));
    }
    {
      Heads H{};
      H.J = 3;
      for (int j = 0; j < 3; ++j) H.h[j] = Head{p->head1_w + j * kHid, p->head1_b + j, 2, a->rgb + j, 3, nullptr, nullptr};
      EO_TRY(run_heads_fwd(prec, st + S.hd0, 2 * kHid, kHid, N, H, s));
    }
    for (int i = 0; i < 3; ++i) {
      GemmNT g;
      g.M = N; g.N = kHid; g.K = kHid;
      if (i == 0) { g.A = st + S.hd0 + (int64_t)kHid * es; g.lda = 2 * kHid; }
      else { g.A = st + S.t[i - 1]; g.lda = kHid; }
      g.B = pr + W.tr[i]; g.ldb = kHid;
      g.C = st + S.t[i]; g.ldc = kHid; g.bias = p->trans_b[i + 1]; g.relu = 1;
      EO_TRY(gemm_nt(prec, g, s));
    }
    {
      Heads H{};
      H.J = 2;
      H.h[0] = Head{p->ts_w, p->ts_b, 2, a->transient_s, 1, nullptr, nullptr};
      H.h[1] = Head{p->tb_w, p->tb_b, 3, a->transient_beta, 1, nullptr, nullptr};
      EO_TRY(run_heads_fwd(prec, st + S.t[2], kHid, kHid, N, H, s));
    }
  } else {
    EO_REQUIRE(a->cond_dirs, "field_fwd: the vanilla field needs cond_dirs");
    EO_TRY(by_type(prec, [&](auto* tag) {
      using T = std::remove_pointer_t<decltype(tag)>;
      encode_dirs_kernel<T><<<div_up(N * 32, 256), 256, 0, s>>>(a->cond_dirs, a->cond_dirs_stride, N, (T*)(st + S.bott), S.ld_bott, kW,
                                                                a->cond_dirs_per_ray ? a->ray_indices : nullptr);
      EO_LAUNCH_CHECK();
      return EONERF_OK;
    }));
    GemmNT g;
    g.M = N; g.N = kHid; g.K = kW + kDirEnc;
    g.A = st + S.bott; g.lda = S.ld_bott; g.B = pr + W.hd0; g.ldb = kW + kDirEnc;
    g.C = st + S.hd0; g.ldc = kHid; g.bias = p->head0_b; g.relu = 1;
    EO_TRY(gemm_nt(prec, g, s));
    Heads H{};
    H.J = 3;
    for (int j = 0; j < 3; ++j) H.h[j] = Head{p->head1_w + j * kHid, p->head1_b + j, 2, a->rgb + j, 3, nullptr, nullptr};
    EO_TRY(run_heads_fwd(prec, st + S.hd0, kHid, kHid, N, H, s));
  }
  return EONERF_OK;
}

static int field_bwd_impl(const EonerfFieldBwdArgs* a, cudaStream_t s) {
  const int prec = a->precision, field = a->field;
  const int64_t N = a->n_pts;
  if (N == 0) return EONERF_OK;
  const EonerfFieldParams* p = a->params;
  const EonerfFieldParams* G = a->grads;
  StashLayout S = stash_layout(field, prec, N, a->density_only);
  PrepLayout W = prep_layout(field, prec, p->n_images);
  ScratchLayout C = scratch_layout(field, prec, N, p->n_images);
  const int es = elem_size(prec);
  const char* st = (const char*)a->stash;
  const char* pr = (const char*)a->prepared;
  char* sc = (char*)a->scratch;
  float* dpre = (float*)(sc + C.dpre);
  char *ga = sc + C.ga, *gb = sc + C.gb, *gc = sc + C.gc;

  auto dW = [&](const void* dY, int64_t lddy, int n, const void* X, int64_t ldx, int k, float* dw, int64_t lddw, float* db) {
    if (!G) return (int)EONERF_OK;
    GemmTN t;
    t.A = dY; t.lda = lddy; t.X = X; t.ldx = ldx; t.M = N; t.N = n; t.K = k; t.D = dw; t.ldd = lddw; t.dbias = db;
    return gemm_tn(prec, t, s);
  };

  const int sigma_act = field == EONERF_FIELD_EONERF ? 3 : 1;
  if (!a->density_only) {
    char *gat = sc + C.gat, *g1 = sc + C.g1, *g2 = sc + C.g2;
    if (field == EONERF_FIELD_EONERF) {
      {  // transient_scalar / transient_beta heads -> d t3
        Heads H{};
        H.J = 2;
        H.h[0] = Head{p->ts_w, p->ts_b, 2, nullptr, 1, a->transient_s, a->g_transient_s};
        H.h[1] = Head{p->tb_w, p->tb_b, 3, nullptr, 1, a->transient_beta, a->g_transient_beta};
        EO_TRY(run_heads_bwd(prec, st + S.t[2], kHid, kHid, N, H, dpre, 4, g1, kHid, s));
        if (G) {
          HeadGrads hg{{G->ts_w, G->tb_w, nullptr}, {G->ts_b, G->tb_b, nullptr}};
          EO_TRY(run_heads_dw(prec, st + S.t[2], kHid, kHid, N, 2, dpre, 4, hg, s));
        }
      }
      // transient_mlp.3, .2, .1
      char* cur = g1;
      char* nxt = g2;
      for (int i = 2; i >= 0; --i) {
        const char* X = (i == 0) ? st + S.hd0 + (int64_t)kHid * es : st + S.t[i - 1];
        int64_t ldx = (i == 0) ? 2 * kHid : kHid;
        EO_TRY(dW(cur, kHid, kHid, X, ldx, kHid, G ? G->trans_w[i + 1] : nullptr, kHid, G ? G->trans_b[i + 1] : nullptr));
        GemmNT g;
        g.M = N; g.N = kHid; g.K = kHid; g.A = cur; g.lda = kHid; g.B = pr + W.tr_t[i]; g.ldb = kHid;
        if (i == 0) { g.C = gat + (int64_t)kHid * es; g.ldc = 2 * kHid; }
        else { g.C = nxt; g.ldc = kHid; }
        g.mask = X; g.ld_mask = ldx; g.mask_cols = kHid;
        EO_TRY(gemm_nt(prec, g, s));
        char* t = cur; cur = nxt; nxt = t;
      }
      {  // albedo head -> d a0 (columns 0:128 of GAT)
        Heads H{};
        H.J = 3;
        for (int j = 0; j < 3; ++j) H.h[j] = Head{p->head1_w + j * kHid, p->head1_b + j, 2, nullptr, 3, a->rgb + j, a->g_rgb ? a->g_rgb + j : nullptr};
        EO_TRY(run_heads_bwd(prec, st + S.hd0, 2 * kHid, kHid, N, H, dpre, 1, gat, 2 * kHid, s));
        if (G) {
          HeadGrads hg{{G->head1_w, G->head1_w + kHid, G->head1_w + 2 * kHid}, {G->head1_b, G->head1_b + 1, G->head1_b + 2}};
          EO_TRY(run_heads_dw(prec, st + S.hd0, 2 * kHid, kHid, N, 3, dpre, 1, hg, s));
        }
      }
      if (G) {
        EO_TRY(dW(gat, 2 * kHid, kHid, st + S.bott, kW, kW, G->head0_w, kW, G->head0_b));
        EO_TRY(dW(gat + (int64_t)kHid * es, 2 * kHid, kHid, st + S.bott, kW, kW, G->trans_w[0], 260, G->trans_b[0]));
        // embedding / W[:,256:260] through the per-image bias rows
        float* dcb = (float*)(sc + C.dcb);
        EO_CUDA(cudaMemsetAsync(dcb, 0, p->n_images * kHid * 4, s));
        int64_t tab_bytes = p->n_images * kHid * 4;
        int use_smem = tab_bytes <= 40 * 1024;
        int64_t rows = 2048;
        while (div_up(N, rows) > 4 * 148) rows *= 2;
        EO_TRY(by_type(prec, [&](auto* tag) {
          using T = std::remove_pointer_t<decltype(tag)>;
          class_grad_kernel<T><<<div_up(N, rows), 256, use_smem ? tab_bytes : 0, s>>>(
              (const T*)(gat + (int64_t)kHid * es), 2 * kHid, (const int32_t*)(st + S.cls), N, p->n_images, dcb, rows, use_smem);
          EO_LAUNCH_CHECK();
          return EONERF_OK;
        }));
        int nthreads = (int)(p->n_images * 4 > kHid * 4 ? p->n_images * 4 : kHid * 4);
        emb_grad_kernel<<<div_up(nthreads, 128), 128, 0, s>>>(dcb, p->trans_w[0], p->transient_emb, p->n_images, G->transient_emb, G->trans_w[0]);
        EO_LAUNCH_CHECK();
      }
      {  // d bottleneck = [d a0 | d t0] [W_a0 ; W_t0]
        GemmNT g;
        g.M = N; g.N = kW; g.K = 2 * kHid; g.A = gat; g.lda = 2 * kHid; g.B = pr + W.hd0_t; g.ldb = 2 * kHid;
        g.C = gb; g.ldc = kW;
        EO_TRY(gemm_nt(prec, g, s));
      }
    } else {
      Heads H{};
      H.J = 3;
      for (int j = 0; j < 3; ++j) H.h[j] = Head{p->head1_w + j * kHid, p->head1_b + j, 2, nullptr, 3, a->rgb + j, a->g_rgb ? a->g_rgb + j : nullptr};
      EO_TRY(run_heads_bwd(prec, st + S.hd0, kHid, kHid, N, H, dpre, 1, g1, kHid, s));
      if (G) {
        HeadGrads hg{{G->head1_w, G->head1_w + kHid, G->head1_w + 2 * kHid}, {G->head1_b, G->head1_b + 1, G->head1_b + 2}};
        EO_TRY(run_heads_dw(prec, st + S.hd0, kHid, kHid, N, 3, dpre, 1, hg, s));
        EO_TRY(dW(g1, kHid, kHid, st + S.bott, S.ld_bott, 283, G->head0_w, 283, G->head0_b));
      }
      GemmNT g;   // d [bottleneck | dir enc] — only the first 256 columns are needed
      g.M = N; g.N = kW; g.K = kHid; g.A = g1; g.lda = kHid; g.B = pr + W.hd0_t; g.ldb = kHid;
      g.C = gb; g.ldc = kW;
      EO_TRY(gemm_nt(prec, g, s));
    }
    EO_TRY(dW(gb, kW, kW, st + S.h[7], kW, kW, G ? G->bott_w : nullptr, kW, G ? G->bott_b : nullptr));
  }
  {  // sigma head
    Heads H{};
    H.J = 1;
    H.h[0] = Head{p->sigma_w, p->sigma_b, sigma_act, nullptr, 1, a->sigma, a->g_sigma};
    EO_TRY(run_heads_bwd(prec, st + S.h[7], kW, kW, N, H, dpre, 0, nullptr, 0, s));
    if (G) {
      HeadGrads hg{{G->sigma_w, nullptr, nullptr}, {G->sigma_b, nullptr, nullptr}};
      EO_TRY(run_heads_dw(prec, st + S.h[7], kW, kW, N, 1, dpre, 0, hg, s));
    }
  }
  if (a->density_only) {
    EO_TRY(by_type(prec, [&](auto* tag) {
      using T = std::remove_pointer_t<decltype(tag)>;
      rank1_mask_kernel<T><<<div_up(N * (kW / 8), 256), 256, 0, s>>>(dpre, 8, p->sigma_w, (const T*)(st + S.h[7]), kW, N, kW, (T*)ga, kW);
      EO_LAUNCH_CHECK();
      return EONERF_OK;
    }));
  } else {  // d h7 = (d bott W_b + d sigma_pre (x) w_sigma) masked
    GemmNT g;
    g.M = N; g.N = kW; g.K = kW; g.A = gb; g.lda = kW; g.B = pr + W.bott_t; g.ldb = kW;
    g.C = ga; g.ldc = kW;
    g.rank1_row = dpre; g.rank1_stride = 8; g.rank1_col = p->sigma_w;
    g.mask = st + S.h[7]; g.ld_mask = kW; g.mask_cols = kW;
    EO_TRY(gemm_nt(prec, g, s));
  }

  // trunk backward.  `cur` holds d(pre-activation) of layer i.
  const bool want_x = a->g_x != nullptr;
  char* cur = ga; int64_t ld_cur = kW;
  // layer 7 -> gb (ld 256) ; 6 -> ga ; 5 -> gb (ld 320, incl. d enc) ; 4 -> ga ; 3 -> gc ; 2 -> ga ; 1 -> gc ; 0 -> ga[:, :64]
  char* dst_of[8] = {ga, gc, ga, gc, ga, gb, ga, gb};
  for (int i = 7; i >= 0; --i) {
    const char* X; int64_t ldx;
    if (i == 0) { X = st + S.h[4] + (int64_t)kW * es; ldx = kH4E; }
    else { X = st + S.h[i - 1]; ldx = (i == 5) ? kH4E : kW; }
    EO_TRY(dW(cur, ld_cur, kW, X, ldx, trunk_k(i), G ? G->trunk_w[i] : nullptr, trunk_k(i), G ? G->trunk_b[i] : nullptr));
    if (i == 0 && !want_x) break;
    GemmNT g;
    g.M = N; g.K = kW; g.A = cur; g.lda = ld_cur; g.B = pr + W.wt[i]; g.ldb = kW;
    g.C = dst_of[i];
    if (i == 0) { g.N = kEnc; g.ldc = kW; }
    else if (i == 5) { g.N = want_x ? kH4E : kW; g.ldc = kH4E; g.mask = X; g.ld_mask = kH4E; g.mask_cols = kW; }
    else { g.N = kW; g.ldc = kW; g.mask = X; g.ld_mask = kW; g.mask_cols = kW; }
    EO_TRY(gemm_nt(prec, g, s));
    cur = dst_of[i]; ld_cur = g.ldc;
  }
  if (want_x) {
    EO_TRY(by_type(prec, [&](auto* tag) {
      using T = std::remove_pointer_t<decltype(tag)>;
      posenc_bwd_kernel<T><<<div_up(N * 3, 256), 256, 0, s>>>((const float*)(st + S.xf), (const T*)ga, kW,
                                                               (const T*)(gb + (int64_t)kW * es), kH4E, N, a->g_x);
      EO_LAUNCH_CHECK();
      return EONERF_OK;
    }));
  }
  return EONERF_OK;
}

}  // namespace eonerf

using namespace eonerf;

static bool valid_prec(int p) { return p == EONERF_PREC_FP32 || p == EONERF_PREC_BF16 || p == EONERF_PREC_BF16_SIMT || p == EONERF_PREC_BF16_FUSED; }
static bool fused(int p) { return p == EONERF_PREC_BF16_FUSED; }
static bool valid_field(int f) { return f == EONERF_FIELD_EONERF || f == EONERF_FIELD_VANILLA; }

extern "C" int64_t eonerf_field_prepared_bytes(int32_t field, int32_t precision, int64_t n_images) {
  if (!valid_prec(precision) || !valid_field(field)) return -1;
  if (fused(precision)) return prep_layout(field, EONERF_PREC_BF16, n_images).total + fused_prepared_extra_bytes(n_images);
  return prep_layout(field, precision, n_images).total;
}
extern "C" int64_t eonerf_field_stash_bytes(int32_t field, int32_t precision, int64_t n_pts, int32_t density_only) {
  if (!valid_prec(precision) || !valid_field(field) || n_pts < 0) return -1;
  if (fused(precision)) return fused_stash_bytes(n_pts, density_only);
  return stash_layout(field, precision, n_pts, density_only).total;
}
extern "C" int64_t eonerf_field_scratch_bytes(int32_t field, int32_t precision, int64_t n_pts, int64_t n_images) {
  if (!valid_prec(precision) || !valid_field(field) || n_pts < 0) return -1;
  if (fused(precision)) return fused_scratch_bytes(n_pts, n_images, 0);
  return scratch_layout(field, precision, n_pts, n_images).total;
}

extern "C" int eonerf_field_prepare(int32_t field, int32_t precision, const EonerfFieldParams* params, void* prepared,
                                    eonerf_stream_t stream) {
  EO_REQUIRE(valid_prec(precision) && valid_field(field), "field_prepare: bad field/precision %d/%d", field, precision);
  EO_REQUIRE(params && prepared, "field_prepare: null pointer");
  EO_REQUIRE(field == EONERF_FIELD_VANILLA || (params->n_images > 0 && params->transient_emb), "field_prepare: need n_images > 0");
  if (fused(precision)) {
    EO_TRY(field_prepare_impl(field, EONERF_PREC_BF16, params, prepared, as_stream(stream)));
    return fused_prepare(field, params, prepared, as_stream(stream));
  }
  return field_prepare_impl(field, precision, params, prepared, as_stream(stream));
}

extern "C" int eonerf_field_fwd(const EonerfFieldFwdArgs* a, eonerf_stream_t stream) {
  EO_REQUIRE(a && valid_prec(a->precision) && valid_field(a->field), "field_fwd: bad field/precision");
  EO_REQUIRE(a->n_pts >= 0, "field_fwd: negative n_pts");
  if (a->n_pts == 0) return EONERF_OK;
  EO_REQUIRE(a->params && a->prepared && a->sigma, "field_fwd: null pointer");
  EO_REQUIRE(a->stash || fused(a->precision), "field_fwd: null stash (only the fused mode can run without one: inference)");
  EO_REQUIRE(a->x || (a->origins && a->viewdirs && a->ray_indices && a->t_starts && a->t_ends),
             "field_fwd: give x or (origins, viewdirs, ray_indices, t_starts, t_ends)");
  EO_REQUIRE(a->density_only || a->rgb, "field_fwd: null rgb output");
  EO_REQUIRE(a->density_only || a->field == EONERF_FIELD_VANILLA || (a->transient_s && a->transient_beta && a->img_idx),
             "field_fwd: the eonerf field needs img_idx and the transient outputs");
  if (fused(a->precision)) return fused_field_fwd(a, as_stream(stream));
  return field_fwd_impl(a, as_stream(stream));
}

extern "C" int eonerf_field_bwd(const EonerfFieldBwdArgs* a, eonerf_stream_t stream) {
  EO_REQUIRE(a && valid_prec(a->precision) && valid_field(a->field), "field_bwd: bad field/precision");
  EO_REQUIRE(a->n_pts >= 0, "field_bwd: negative n_pts");
  if (a->n_pts == 0) return EONERF_OK;
  EO_REQUIRE(a->params && a->prepared && a->stash && a->scratch && a->sigma, "field_bwd: null pointer");
  EO_REQUIRE(a->density_only || a->rgb, "field_bwd: null rgb");
  EO_REQUIRE(a->density_only || a->field == EONERF_FIELD_VANILLA || (a->transient_s && a->transient_beta),
             "field_bwd: the eonerf field needs the transient forward outputs");
  if (fused(a->precision)) return fused_field_bwd(a, as_stream(stream));
  return field_bwd_impl(a, as_stream(stream));
}

// ---- per-ray ambient colour (eonerf.py:163-164): enc4(sun) -> 128 relu -> 3 sigmoid, fp32 ----------
// stash floats per ray: [enc 32 | hidden 128]; scratch: [d hidden 128 | dpre 8]
extern "C" int eonerf_ambient_fwd(const EonerfAmbientFwdArgs* a, eonerf_stream_t stream) {
  EO_REQUIRE(a && a->n_rays >= 0, "ambient_fwd: bad arguments");
  if (a->n_rays == 0) return EONERF_OK;
  EO_REQUIRE(a->sundirs && a->w0 && a->b0 && a->w1 && a->b1 && a->stash && a->ambient, "ambient_fwd: null pointer");
  cudaStream_t s = as_stream(stream);
  int64_t B = a->n_rays;
  float* enc = a->stash;
  float* hid = a->stash + B * kDirEnc;
  encode_dirs_kernel<float><<<div_up(B * 32, 256), 256, 0, s>>>(a->sundirs, a->sundirs_stride, B, enc, kDirEnc, 0, nullptr);
  EO_LAUNCH_CHECK();
  GemmNT g;
  g.M = B; g.N = kHid; g.K = 27; g.A = enc; g.lda = kDirEnc; g.B = a->w0; g.ldb = 27; g.C = hid; g.ldc = kHid;
  g.bias = a->b0; g.relu = 1;
  EO_TRY(gemm_nt_simt(kF32, g, s));
  Heads H{};
  H.J = 3;
  for (int j = 0; j < 3; ++j) H.h[j] = Head{a->w1 + j * kHid, a->b1 + j, 2, a->ambient + j, 3, nullptr, nullptr};
  return run_heads_fwd(EONERF_PREC_FP32, hid, kHid, kHid, B, H, s);
}

extern "C" int eonerf_ambient_bwd(const EonerfAmbientBwdArgs* a, eonerf_stream_t stream) {
  EO_REQUIRE(a && a->n_rays >= 0, "ambient_bwd: bad arguments");
  if (a->n_rays == 0) return EONERF_OK;
  EO_REQUIRE(a->w0 && a->w1 && a->stash && a->scratch && a->ambient && a->g_ambient && a->g_w0 && a->g_b0 && a->g_w1 && a->g_b1,
             "ambient_bwd: null pointer");
  cudaStream_t s = as_stream(stream);
  int64_t B = a->n_rays;
  const float* enc = a->stash;
  const float* hid = a->stash + B * kDirEnc;
  float* dh = a->scratch;
  float* dpre = a->scratch + B * kHid;
  Heads H{};
  H.J = 3;
  for (int j = 0; j < 3; ++j) H.h[j] = Head{a->w1 + j * kHid, nullptr, 2, nullptr, 3, a->ambient + j, a->g_ambient + j};
  EO_TRY(run_heads_bwd(EONERF_PREC_FP32, hid, kHid, kHid, B, H, dpre, 0, dh, kHid, s));
  HeadGrads hg{{a->g_w1, a->g_w1 + kHid, a->g_w1 + 2 * kHid}, {a->g_b1, a->g_b1 + 1, a->g_b1 + 2}};
  EO_TRY(run_heads_dw(EONERF_PREC_FP32, hid, kHid, kHid, B, 3, dpre, 0, hg, s));
  GemmTN t;
  t.A = dh; t.lda = kHid; t.X = enc; t.ldx = kDirEnc; t.M = B; t.N = kHid; t.K = 27; t.D = a->g_w0; t.ldd = 27; t.dbias = a->g_b0;
  return gemm_tn_simt(kF32, t, s);
}

// ---- building blocks ---------------------------------------------------------------------------
extern "C" int eonerf_linear_fwd(const EonerfLinearArgs* a, eonerf_stream_t stream) {
  EO_REQUIRE(a && valid_prec(a->precision), "linear_fwd: bad precision");
  EO_REQUIRE(a->m >= 0 && a->n > 0 && a->k > 0, "linear_fwd: bad shape");
  if (a->m == 0) return EONERF_OK;
  EO_REQUIRE(a->x && a->w && a->y, "linear_fwd: null pointer");
  GemmNT g;
  g.A = a->x; g.lda = a->ldx; g.B = a->w; g.ldb = a->ldw; g.C = a->y; g.ldc = a->ldy;
  g.M = a->m; g.N = a->n; g.K = a->k; g.bias = a->bias; g.relu = a->act == 1;
  return gemm_nt(a->precision, g, as_stream(stream));
}

extern "C" int eonerf_linear_dw(const EonerfDwArgs* a, eonerf_stream_t stream) {
  EO_REQUIRE(a && valid_prec(a->precision), "linear_dw: bad precision");
  EO_REQUIRE(a->m >= 0 && a->n > 0 && a->k > 0, "linear_dw: bad shape");
  if (a->m == 0) return EONERF_OK;
  EO_REQUIRE(a->dy && a->x && a->dw, "linear_dw: null pointer");
  GemmTN t;
  t.A = a->dy; t.lda = a->lddy; t.X = a->x; t.ldx = a->ldx; t.M = a->m; t.N = a->n; t.K = a->k;
  t.D = a->dw; t.ldd = a->lddw; t.dbias = a->db;
  return gemm_tn(a->precision, t, as_stream(stream));
}
