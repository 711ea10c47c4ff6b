// Uniform ray marching inside an axis-aligned box (BASELINE configs[1], the vanilla-NeRF benchmark of train_mlp_nerf.py).
//
// /root/reference/train_mlp_nerf.py:155-170 calls render_image_with_occgrid -> estimator.sampling(...) of nerfacc v0.5.2, but the
// helper module it imports (`utils2`, train_mlp_nerf.py:17) is missing from the reference, so that entry point cannot run and
// there is nothing to pin against.  What is implemented is the occupancy-free limit of that sampler (every cell occupied): the
// part of the ray inside the scene box [near_plane, far_plane] is cut into consecutive intervals of render_step_size, shifted by
// one stratified offset per ray in training (nerfacc's `stratified`), packed ray after ray like every other sampler here.
//   t_min / t_max : slab test against the box, clipped to [near, far]
//   t0 = t_min + jitter * step;  n = ceil((t_max - t0) / step);  interval k = [t0 + k step, min(t0 + (k+1) step, t_max))
// fp32 with separately rounded operations (the oracle restates it in torch fp32 and the indices must agree bit for bit).
#include "common.cuh"

namespace eonerf {

__device__ __forceinline__ void march_range(const EonerfMarchArgs& a, int64_t ray, float& t0, int& n) {
  const float* o = a.origins + ray * a.origins_stride;
  const float* d = a.viewdirs + ray * a.viewdirs_stride;
  float tmin = a.near_plane, tmax = a.far_plane;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float oc = __ldg(o + c), dc = __ldg(d + c);
    const float t1 = __fdiv_rn(__fsub_rn(a.aabb[c], oc), dc), t2 = __fdiv_rn(__fsub_rn(a.aabb[3 + c], oc), dc);
    tmin = fmaxf(tmin, fminf(t1, t2));
    tmax = fminf(tmax, fmaxf(t1, t2));
  }
  const float j = a.jitter ? __ldg(a.jitter + ray) : 0.0f;
  t0 = __fadd_rn(tmin, __fmul_rn(j, a.step));
  n = 0;
  if (tmax > t0) n = (int)ceilf(__fdiv_rn(__fsub_rn(tmax, t0), a.step));
  if (n > a.max_per_ray) n = a.max_per_ray;
  a.t_max_out[ray] = tmax;
}

__global__ void __launch_bounds__(256) march_count_kernel(EonerfMarchArgs a) {
  const int64_t ray = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (ray >= a.n_rays) return;
  float t0;
  int n;
  march_range(a, ray, t0, n);
  a.counts[ray] = n;
  a.t0_out[ray] = t0;
}

// one warp per ray, lanes over the intervals: coalesced writes of (ray index, t_start, t_end)
__global__ void __launch_bounds__(256) march_write_kernel(EonerfMarchArgs a) {
  const int lane = threadIdx.x & 31;
  for (int64_t ray = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); ray < a.n_rays; ray += (int64_t)gridDim.x * 8) {
    const int64_t beg = a.ray_offsets[ray], end = a.ray_offsets[ray + 1];
    const float t0 = a.t0_out[ray], tmax = a.t_max_out[ray];
    for (int64_t k = lane; k < end - beg; k += 32) {
      const float ts = __fadd_rn(t0, __fmul_rn((float)k, a.step));
      a.ray_indices[beg + k] = ray;
      a.t_starts[beg + k] = ts;
      a.t_ends[beg + k] = fminf(__fadd_rn(ts, a.step), tmax);
    }
  }
}

}  // namespace eonerf

using namespace eonerf;

extern "C" int eonerf_march_count(const EonerfMarchArgs* a, eonerf_stream_t stream) {
  EO_REQUIRE(a && a->n_rays >= 0 && a->step > 0.f && a->max_per_ray > 0, "march_count: bad arguments");
  if (a->n_rays == 0) return EONERF_OK;
  EO_REQUIRE(a->origins && a->viewdirs && a->counts && a->t0_out && a->t_max_out, "march_count: null pointer");
  march_count_kernel<<<div_up(a->n_rays, 256), 256, 0, as_stream(stream)>>>(*a);
  EO_LAUNCH_CHECK();
  return EONERF_OK;
}

extern "C" int eonerf_march_write(const EonerfMarchArgs* a, eonerf_stream_t stream) {
  EO_REQUIRE(a && a->n_rays >= 0, "march_write: bad arguments");
  if (a->n_rays == 0) return EONERF_OK;
  EO_REQUIRE(a->ray_offsets && a->t0_out && a->t_max_out && a->ray_indices && a->t_starts && a->t_ends, "march_write: null pointer");
  int64_t blocks = div_up(a->n_rays, 8);
  if (blocks > 148 * 16) blocks = 148 * 16;
  march_write_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(*a);
  EO_LAUNCH_CHECK();
  return EONERF_OK;
}
