// Adam step of the training loop (/root/reference/train_eonerf.py:57,158-160: torch.optim.Adam(lr=5e-4), default betas /
// eps, no weight decay, no amsgrad) over ONE flat fp32 buffer: every parameter, its gradient and both moments are views of
// four flat arrays, so the whole update is a single 128-bit-vectorised pass (680 k floats) instead of torch's per-tensor
// multi-tensor launches.  The step counter lives on the device (a captured CUDA graph replays the same launch); the bias
// corrections are computed from it in double precision, as torch.optim.Adam's default path does from Python floats.
#include "common.cuh"

namespace eonerf {

// one thread: step += 1, then this step's scalars in DOUBLE precision as torch.optim.Adam's default (non-capturable) path
// computes them from Python floats: step[1] = lr / (1 - beta1^t), step[2] = sqrt(1 - beta2^t)
__global__ void adam_tick_kernel(EonerfAdamArgs a) {
  const float t = a.step[0] + 1.0f;
  a.step[0] = t;
  const double lr = a.lr_dev ? *a.lr_dev : a.lr;
  a.step[1] = (float)(lr / (1.0 - pow(a.beta1, (double)t)));
  a.step[2] = (float)sqrt(1.0 - pow(a.beta2, (double)t));
}

__global__ void __launch_bounds__(256) adam_kernel(EonerfAdamArgs a) {
  const float step_size = a.step[1], bc2_sqrt = a.step[2];
  const int64_t n4 = a.n >> 2;
  const float b2 = (float)a.beta2, omb1 = (float)(1.0 - a.beta1), omb2 = (float)(1.0 - a.beta2), gs = a.grad_scale, eps = (float)a.eps;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 p = reinterpret_cast<float4*>(a.param)[i];
    const float4 g = reinterpret_cast<const float4*>(a.grad)[i];
    float4 m = reinterpret_cast<float4*>(a.exp_avg)[i];
    float4 v = reinterpret_cast<float4*>(a.exp_avg_sq)[i];
#define EO_ADAM(c)                                                   \
  {                                                                  \
    const float gc = g.c * gs;                                       \
    m.c = m.c + (gc - m.c) * omb1;                 /* lerp_ */        \
    v.c = v.c * b2 + gc * gc * omb2;               /* mul_ + addcmul_ */ \
    const float denom = sqrtf(v.c) / bc2_sqrt + eps;                 \
    p.c = p.c - step_size * (m.c / denom);         /* addcdiv_ */     \
  }
    EO_ADAM(x) EO_ADAM(y) EO_ADAM(z) EO_ADAM(w)
    reinterpret_cast<float4*>(a.param)[i] = p;
    reinterpret_cast<float4*>(a.exp_avg)[i] = m;
    reinterpret_cast<float4*>(a.exp_avg_sq)[i] = v;
  }
  if (blockIdx.x == 0 && threadIdx.x < (a.n & 3)) {            // tail (n not a multiple of 4)
    const int64_t i = (n4 << 2) + threadIdx.x;
    const float gc = a.grad[i] * gs;
    const float m = a.exp_avg[i] + (gc - a.exp_avg[i]) * omb1;
    const float v = a.exp_avg_sq[i] * b2 + gc * gc * omb2;
    a.exp_avg[i] = m;
    a.exp_avg_sq[i] = v;
    a.param[i] -= step_size * (m / (sqrtf(v) / bc2_sqrt + eps));
  }
#undef EO_ADAM
}

}  // namespace eonerf

using namespace eonerf;

extern "C" int eonerf_adam_step(const EonerfAdamArgs* a, eonerf_stream_t stream) {
  EO_REQUIRE(a && a->n >= 0 && a->param && a->grad && a->exp_avg && a->exp_avg_sq && a->step, "adam_step: bad arguments");
  EO_REQUIRE((((uintptr_t)a->param | (uintptr_t)a->grad | (uintptr_t)a->exp_avg | (uintptr_t)a->exp_avg_sq) & 15) == 0,
             "adam_step: the flat buffers must be 16-byte aligned");
  if (a->n == 0) return EONERF_OK;
  cudaStream_t s = as_stream(stream);
  int64_t blocks = div_up(a->n >> 2, 256);
  if (blocks < 1) blocks = 1;
  if (blocks > 148 * 8) blocks = 148 * 8;
  adam_tick_kernel<<<1, 1, 0, s>>>(*a);
  EO_LAUNCH_CHECK();
  adam_kernel<<<(unsigned)blocks, 256, 0, s>>>(*a);
  EO_LAUNCH_CHECK();
  return EONERF_OK;
}

// ------------------------------------------------------------------------------------------------
// Device-resident ray table: one training batch = rows perm[first .. first + B) of the dataset's all_rays / all_rgbs /
// all_ids_img (/root/reference/datasets/satellite.py:799-807 `__getitem__` + the DataLoader's collate,
// train_eonerf.py:70,99-109), gathered by one kernel instead of B Python `__getitem__` calls, worker IPC and a pageable
// host -> device copy.  One thread per (row, 16-byte-or-less piece): rows are 44 + 12 + 8 bytes.
// ------------------------------------------------------------------------------------------------
namespace eonerf {
__global__ void __launch_bounds__(256) gather_batch_kernel(EonerfGatherBatchArgs a) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t row = t >> 4;
  const int c = (int)(t & 15);
  if (row >= a.batch) return;
  int64_t src = __ldg(a.perm + a.first + row);
  if (src < 0 || src >= a.n_rows) src = 0;                 // a corrupt index must not read out of bounds
  if (c < 11) a.rays_out[row * 11 + c] = __ldg(a.all_rays + src * a.rays_stride + c);
  else if (c < 14) a.rgbs_out[row * 3 + (c - 11)] = __ldg(a.all_rgbs + src * a.rgbs_stride + (c - 11));
  else if (c == 14) a.ts_out[row] = __ldg(a.all_ts + src);
  else if (a.idx_out) a.idx_out[row] = src;
}
}  // namespace eonerf

extern "C" int eonerf_gather_batch(const EonerfGatherBatchArgs* a, eonerf_stream_t stream) {
  EO_REQUIRE(a && a->batch >= 0 && a->n_rows > 0 && a->first >= 0, "gather_batch: bad arguments");
  if (a->batch == 0) return EONERF_OK;
  EO_REQUIRE(a->all_rays && a->all_rgbs && a->all_ts && a->perm && a->rays_out && a->rgbs_out && a->ts_out, "gather_batch: null pointer");
  gather_batch_kernel<<<(unsigned)div_up(a->batch * 16, 256), 256, 0, as_stream(stream)>>>(*a);
  EO_LAUNCH_CHECK();
  return EONERF_OK;
}

// ------------------------------------------------------------------------------------------------
// Training losses on the packed per-ray outputs (out[B,21]: rgb = columns 0:3, beta = column 12), value AND gradient in
// one pass: metrics.mse (train_eonerf.py:139-140, epoch < 2) and metrics.uncertainty_aware_loss (metrics.py:17-22)
//     color = mean_{r,c} (rgb - gt)^2 / (2 beta^2),   logbeta = (3 + mean_r log beta) / 2,   loss = color + logbeta.
// Replaces ~25 element-wise / reduction / slice-backward launches of the autograd graph.  Deterministic: per-block
// partial sums, then one block adds them in a fixed order.
// ------------------------------------------------------------------------------------------------
namespace eonerf {
__global__ void __launch_bounds__(256) loss_rows_kernel(EonerfLossArgs a) {
  __shared__ float red[2][8];
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float color = 0.f, lb = 0.f;
  if (r < a.n_rays) {
    const float* o = a.out + r * EONERF_OUT_COLS;
    float* g = a.g_out + r * EONERF_OUT_COLS;
    const float inv3b = 1.0f / (3.0f * (float)a.n_rays);
    float d[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) d[c] = o[c] - __ldg(a.gt_rgb + r * 3 + c);
    const float sq = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
#pragma unroll
    for (int c = 0; c < EONERF_OUT_COLS; ++c) g[c] = 0.f;
    if (a.mode == 0) {                                        // mse
      color = sq;
#pragma unroll
      for (int c = 0; c < 3; ++c) g[c] = 2.0f * d[c] * inv3b;
    } else {
      const float beta = o[12];
      const float ib2 = 1.0f / (beta * beta);
      color = sq * 0.5f * ib2;
      lb = logf(beta);
#pragma unroll
      for (int c = 0; c < 3; ++c) g[c] = d[c] * ib2 * inv3b;
      g[12] = -sq * ib2 / beta * inv3b + 0.5f / ((float)a.n_rays * beta);
    }
  }
  color = warp_sum(color);
  lb = warp_sum(lb);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { red[0][warp] = color; red[1][warp] = lb; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float c = 0.f, l = 0.f;
    for (int w = 0; w < 8; ++w) { c += red[0][w]; l += red[1][w]; }
    a.partials[2 * blockIdx.x] = c;
    a.partials[2 * blockIdx.x + 1] = l;
  }
}

__global__ void __launch_bounds__(256) loss_finish_kernel(EonerfLossArgs a, int n_blocks) {
  __shared__ float red[2][8];
  float c = 0.f, l = 0.f;
  for (int i = threadIdx.x; i < n_blocks; i += 256) { c += a.partials[2 * i]; l += a.partials[2 * i + 1]; }
  c = warp_sum(c);
  l = warp_sum(l);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { red[0][warp] = c; red[1][warp] = l; }
  __syncthreads();
  if (threadIdx.x == 0) {
    c = 0.f; l = 0.f;
    for (int w = 0; w < 8; ++w) { c += red[0][w]; l += red[1][w]; }
    const float color = c / (3.0f * (float)a.n_rays);
    const float logbeta = a.mode == 0 ? 0.f : (3.0f + l / (float)a.n_rays) * 0.5f;
    a.loss[0] = color + logbeta;
    a.loss[1] = color;
    a.loss[2] = logbeta;
  }
}
}  // namespace eonerf

extern "C" int64_t eonerf_loss_partials(int64_t n_rays) { return 2 * div_up(n_rays > 0 ? n_rays : 1, 256); }

extern "C" int eonerf_loss_fwd_bwd(const EonerfLossArgs* a, eonerf_stream_t stream) {
  EO_REQUIRE(a && a->n_rays > 0 && (a->mode == 0 || a->mode == 1), "loss_fwd_bwd: bad arguments");
  EO_REQUIRE(a->out && a->gt_rgb && a->g_out && a->loss && a->partials, "loss_fwd_bwd: null pointer");
  cudaStream_t s = as_stream(stream);
  const int blocks = (int)div_up(a->n_rays, 256);
  loss_rows_kernel<<<blocks, 256, 0, s>>>(*a);
  EO_LAUNCH_CHECK();
  loss_finish_kernel<<<1, 256, 0, s>>>(*a, blocks);
  EO_LAUNCH_CHECK();
  return EONERF_OK;
}
