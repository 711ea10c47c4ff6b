// Evaluation epilogue of the per-ray path, device-resident up to the GeoTIFF write:
//   * utm_points_kernel: rendered depth -> UTM/altitude point cloud, x = (o + d * depth) * scene_scale + scene_offset in
//     fp64 (/root/reference/datasets/satellite.py:502-531, utm_sampling branch) + the fp32 altitude column of the render arm;
//   * dsm_splat_kernel / dsm_finish_kernel: the point cloud rasterised into a DSM grid with the semantics of
//     plyflatten(cloud, xoff, yoff, resolution, xsize, ysize, radius=1, sigma=inf) (satellite.py:548-587).
// HBM-bound element-wise / scatter work: one thread per ray, 128-bit-friendly SoA outputs, fp64 atomics for the cell sums.
#include "common.cuh"

namespace eonerf {

__global__ void __launch_bounds__(256) utm_points_kernel(EonerfUtmPointsArgs a) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n_rays; i += (int64_t)gridDim.x * blockDim.x) {
    const float* r = a.rays + i * a.rays_stride;
    const double t = (double)a.depth[i * a.depth_stride];
    // rays.double(), depth.double(), then o + d * depth, * scale, + offset: separately rounded fp64 operations, as torch
    double p[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const double xn = __dadd_rn((double)r[c], __dmul_rn((double)r[3 + c], t));
      p[c] = __dadd_rn(__dmul_rn(xn, a.scene_scale[c]), a.scene_offset[c]);
    }
    if (a.easts) a.easts[i] = p[0];
    if (a.norths) a.norths[i] = p[1];
    if (a.alts) a.alts[i] = p[2];
    if (a.alt_f32) a.alt_f32[i] = (float)p[2];
  }
}

// plyflatten.c (package `plyflatten`, un-vendored; restated from the published source):
//   i = floor((x - xoff) / res), j = floor((-y + yoff) / res); every cell (i+k1, j+k2) with k1^2 + k2^2 <= radius^2 inside the
//   grid receives the point's height with weight w = 1 (sigma = inf) or exp(-dist^2 / (2 sigma^2)), dist = distance from the
//   point to the cell centre; the cell value is sum(w z) / sum(w), NaN where no point fell.
__global__ void __launch_bounds__(256) dsm_splat_kernel(EonerfDsmArgs a) {
  const bool gauss = isfinite(a.sigma);
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < a.n_points; p += (int64_t)gridDim.x * blockDim.x) {
    if (a.depth && !(a.depth[p * a.depth_stride] >= 0.0f)) continue;      // "negative depths are not allowed" (satellite.py:561)
    const double x = a.easts[p];
    double y = a.norths[p];
    if (y < 0.0) y += a.negative_north_shift;                              // cloud[cloud[:,1] < 0, 1] += 10e6 (satellite.py:559)
    const double z = a.alts[p];
    const int i = (int)floor((x - a.xoff) / a.resolution);
    const int j = (int)floor((-y + a.yoff) / a.resolution);
    for (int k1 = -a.radius; k1 <= a.radius; ++k1)
      for (int k2 = -a.radius; k2 <= a.radius; ++k2) {
        if (k1 * k1 + k2 * k2 > a.radius * a.radius) continue;
        const int ii = i + k1, jj = j + k2;
        if (ii < 0 || jj < 0 || ii >= a.xsize || jj >= a.ysize) continue;
        double w = 1.0;
        if (gauss) {
          const double dx = x - (a.xoff + a.resolution * (0.5 + ii));
          const double dy = y - (a.yoff - a.resolution * (0.5 + jj));
          w = exp(-(dx * dx + dy * dy) / (2.0 * a.sigma * a.sigma));
        }
        const int64_t cell = (int64_t)jj * a.xsize + ii;
        atomicAdd(a.acc + 2 * cell, w * z);
        atomicAdd(a.acc + 2 * cell + 1, w);
      }
  }
}

__global__ void __launch_bounds__(256) dsm_finish_kernel(const double* acc, int64_t n_cells, float* dsm) {
  for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < n_cells; c += (int64_t)gridDim.x * blockDim.x) {
    const double w = acc[2 * c + 1];
    dsm[c] = w > 0.0 ? (float)(acc[2 * c] / w) : __int_as_float(0x7fc00000);
  }
}

static inline unsigned grid_for(int64_t n) {
  int64_t b = div_up(n, 256);
  if (b > 148 * 8) b = 148 * 8;
  return (unsigned)(b < 1 ? 1 : b);
}

}  // namespace eonerf

using namespace eonerf;

extern "C" int eonerf_utm_points(const EonerfUtmPointsArgs* a, eonerf_stream_t stream) {
  EO_REQUIRE(a && a->n_rays >= 0 && a->rays && a->depth && a->rays_stride >= 6 && a->depth_stride >= 1, "utm_points: bad arguments");
  if (a->n_rays == 0) return EONERF_OK;
  utm_points_kernel<<<grid_for(a->n_rays), 256, 0, as_stream(stream)>>>(*a);
  EO_LAUNCH_CHECK();
  return EONERF_OK;
}

extern "C" int eonerf_dsm_rasterize(const EonerfDsmArgs* a, eonerf_stream_t stream) {
  EO_REQUIRE(a && a->n_points >= 0 && a->easts && a->norths && a->alts && a->acc && a->dsm, "dsm_rasterize: bad arguments");
  EO_REQUIRE(a->xsize > 0 && a->ysize > 0 && a->resolution > 0 && a->radius >= 0 && a->radius <= 8, "dsm_rasterize: bad grid");
  cudaStream_t s = as_stream(stream);
  const int64_t cells = (int64_t)a->xsize * a->ysize;
  EO_CUDA(cudaMemsetAsync(a->acc, 0, cells * 2 * sizeof(double), s));
  if (a->n_points > 0) {
    dsm_splat_kernel<<<grid_for(a->n_points), 256, 0, s>>>(*a);
    EO_LAUNCH_CHECK();
  }
  dsm_finish_kernel<<<grid_for(cells), 256, 0, s>>>(a->acc, cells, a->dsm);
  EO_LAUNCH_CHECK();
  return EONERF_OK;
}
