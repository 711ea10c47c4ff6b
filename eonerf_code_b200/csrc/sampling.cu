// Stratified sampling between the ray bounds, cube mask and order-preserving compaction.
//
// Restates satnerf_sampling / perturb_z_vals / filter_pts_outside_cube
// (/root/reference/sat_rendering.py:18-22,46-54,56-84) as three launches:
//   1. count   : one warp per ray evaluates the n-1 intervals and counts the kept ones,
//   2. scan    : one CTA turns the counts into exclusive offsets (packed info) + totals,
//   3. scatter : one warp per ray re-evaluates the intervals and writes the kept ones in order.
// Re-evaluating is cheaper than staging a dense [B, n-1] copy through HBM: the only per-sample
// input is the 4-byte uniform.
//
// Bit-exactness: the reference runs every arithmetic step as its own ATen kernel, so there is never
// an FMA contraction across steps.  All arithmetic below therefore uses the __f*_rn intrinsics
// (never contracted by nvcc) in exactly the reference's order.
#include "common.cuh"

namespace eonerf {

struct RayGeom {
  float ox, oy, oz, dx, dy, dz, near, far;
};

__device__ __forceinline__ float z_linear(const RayGeom& r, const float* __restrict__ zs, int i) {
  // z_vals = near * (1 - z_steps) + far * z_steps          (sat_rendering.py:68)
  float s = __ldg(zs + i);
  return __fadd_rn(__fmul_rn(r.near, __fsub_rn(1.0f, s)), __fmul_rn(r.far, s));
}

__device__ __forceinline__ float z_perturbed(const RayGeom& r, const float* __restrict__ zs,
                                             const float* __restrict__ u_row, int i, int n) {
  // perturb_z_vals (sat_rendering.py:46-54)
  float zi = z_linear(r, zs, i);
  float lower = (i == 0) ? zi : __fmul_rn(0.5f, __fadd_rn(z_linear(r, zs, i - 1), zi));
  float upper = (i == n - 1) ? zi : __fmul_rn(0.5f, __fadd_rn(zi, z_linear(r, zs, i + 1)));
  return __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), __ldg(u_row + i)));
}

// interval i in [0, n-1): t_start, t_end and the keep flag
__device__ __forceinline__ bool interval(const RayGeom& r, const float* __restrict__ zs,
                                         const float* __restrict__ u_row, int i, int n, float& ts, float& te) {
  float z0 = z_perturbed(r, zs, u_row, i, n);
  float z1 = z_perturbed(r, zs, u_row, i + 1, n);
  ts = z0;
  te = __fadd_rn(z0, __fsub_rn(z1, z0));                      // :74  z[:-1] + (z[1:] - z[:-1])
  float zm = __fdiv_rn(__fadd_rn(ts, te), 2.0f);              // :79
  float x = __fadd_rn(r.ox, __fmul_rn(r.dx, zm));             // :80
  float y = __fadd_rn(r.oy, __fmul_rn(r.dy, zm));
  float z = __fadd_rn(r.oz, __fmul_rn(r.dz, zm));
  // :20  keep iff no coordinate has |c| >= 1
  return !((fabsf(x) >= 1.0f) || (fabsf(y) >= 1.0f) || (fabsf(z) >= 1.0f));
}

__device__ __forceinline__ RayGeom load_ray(const EonerfSampleArgs& a, int64_t ray) {
  RayGeom r;
  const float* o = a.origins + ray * a.origins_stride;
  const float* d = a.viewdirs + ray * a.viewdirs_stride;
  r.ox = __ldg(o); r.oy = __ldg(o + 1); r.oz = __ldg(o + 2);
  r.dx = __ldg(d); r.dy = __ldg(d + 1); r.dz = __ldg(d + 2);
  r.near = a.near ? __ldg(a.near + ray * a.near_stride) : 0.0f;
  r.far = __fadd_rn(r.near, 2.0f);                             // :63
  return r;
}

__global__ void __launch_bounds__(256) sample_count_kernel(EonerfSampleArgs a) {
  if (a.run_if && *a.run_if == 0) return;
  int lane = threadIdx.x & 31;
  int64_t ray = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (ray >= a.n_rays) return;
  RayGeom r = load_ray(a, ray);
  const float* u_row = a.u + ray * a.n_samples;
  int S = a.n_samples - 1, cnt = 0;
  for (int base = 0; base < S; base += 32) {
    int i = base + lane;
    float ts, te;
    bool keep = (i < S) && interval(r, a.z_steps, u_row, i, a.n_samples, ts, te);
    cnt += __popc(__ballot_sync(kFull, keep));
  }
  if (lane == 0) a.ray_offsets[ray + 1] = cnt;
}

// single CTA: in-place inclusive scan of ray_offsets[1..B]; ray_offsets[0] = 0; also emits the fp32
// per-ray counts (count_number_of_pts_per_nerfacc_ray, sat_rendering.py:10-16) and the totals.
constexpr int kScanItems = 8;     // rays per thread per sweep: 8192 rays per sweep of the single scan CTA
__global__ void __launch_bounds__(1024) sample_scan_kernel(EonerfSampleArgs a) {
  __shared__ long long warp_tot[32];
  __shared__ long long carry_s;
  __shared__ int empty_s;
  int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const bool redraw = a.run_if != nullptr;                   // conditional second draw: keeps pts_per_ray and stats[1]
  if (redraw && *a.run_if == 0) return;
  if (tid == 0) { carry_s = 0; empty_s = 0; a.ray_offsets[0] = 0; }
  __syncthreads();
  int n_empty = 0;
  for (int64_t base = 0; base < a.n_rays; base += 1024 * kScanItems) {
    const int64_t i0 = base + (int64_t)tid * kScanItems;
    long long c[kScanItems];
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) c[k] = (i0 + k < a.n_rays) ? a.ray_offsets[i0 + k + 1] : 0;
    long long v = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
      if (i0 + k < a.n_rays) {
        if (!redraw) a.pts_per_ray[i0 + k] = (float)c[k];
        n_empty += (c[k] == 0);
      }
      v += c[k];
      c[k] = v;                                               // inclusive prefix inside the thread
    }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      long long t = __shfl_up_sync(kFull, v, o);
      if (lane >= o) v += t;
    }
    if (lane == 31) warp_tot[wid] = v;
    __syncthreads();
    if (wid == 0) {
      long long w = warp_tot[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        long long t = __shfl_up_sync(kFull, w, o);
        if (lane >= o) w += t;
      }
      warp_tot[lane] = w;
    }
    __syncthreads();
    // exclusive prefix of this thread = carry + previous warps + previous lanes
    const long long before = carry_s + (wid ? warp_tot[wid - 1] : 0) + (v - c[kScanItems - 1]);
#pragma unroll
    for (int k = 0; k < kScanItems; ++k)
      if (i0 + k < a.n_rays) a.ray_offsets[i0 + k + 1] = before + c[k];
    __syncthreads();
    if (tid == 1023) carry_s = before + c[kScanItems - 1];
    __syncthreads();
  }
  n_empty = __reduce_add_sync(kFull, n_empty);
  if (lane == 0 && n_empty) atomicAdd(&empty_s, n_empty);
  __syncthreads();
  if (tid == 0) {
    a.stats[0] = carry_s;
    if (!redraw) a.stats[1] = empty_s;
  }
}

__global__ void __launch_bounds__(256) sample_scatter_kernel(EonerfSampleArgs a) {
  if (a.run_if && *a.run_if == 0) return;
  int lane = threadIdx.x & 31;
  int64_t ray = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (ray >= a.n_rays) return;
  RayGeom r = load_ray(a, ray);
  const float* u_row = a.u + ray * a.n_samples;
  int S = a.n_samples - 1;
  int64_t out = a.ray_offsets[ray];
  for (int base = 0; base < S; base += 32) {
    int i = base + lane;
    float ts = 0.f, te = 0.f;
    bool keep = (i < S) && interval(r, a.z_steps, u_row, i, a.n_samples, ts, te);
    unsigned m = __ballot_sync(kFull, keep);
    if (keep) {
      int64_t p = out + __popc(m & ((1u << lane) - 1u));
      a.ray_indices[p] = ray;
      a.t_starts[p] = ts;
      a.t_ends[p] = te;
    }
    out += __popc(m);
  }
}

// ---- pack_info: offsets from sorted ray_indices (lower_bound per ray) -------------------------
__global__ void pack_info_kernel(const int64_t* __restrict__ ri, int64_t n_pts, int64_t n_rays,
                                 int64_t* __restrict__ offs) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r > n_rays) return;
  int64_t lo = 0, hi = n_pts;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (__ldg(ri + mid) < r) lo = mid + 1; else hi = mid;
  }
  offs[r] = lo;
}

__global__ void set_last_kernel(float* __restrict__ t_ends, const int64_t* __restrict__ offs, int64_t n_rays,
                                float value) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rays) return;
  int64_t a = offs[r], b = offs[r + 1];
  if (b > a) t_ends[b - 1] = value;
}

}  // namespace eonerf

using namespace eonerf;

extern "C" int eonerf_sample_compact(const EonerfSampleArgs* a, eonerf_stream_t stream) {
  EO_REQUIRE(a && a->n_samples >= 2 && a->n_rays >= 0, "sample_compact: need n_samples >= 2 and n_rays >= 0");
  EO_REQUIRE(a->ray_offsets && a->stats, "sample_compact: null ray_offsets / stats");
  EO_REQUIRE(a->n_rays == 0 || (a->origins && a->viewdirs && a->u && a->z_steps), "sample_compact: null input");
  EO_REQUIRE(a->n_rays == 0 || (a->ray_indices && a->t_starts && a->t_ends && (a->pts_per_ray || a->run_if)), "sample_compact: null output");
  cudaStream_t s = as_stream(stream);
  if (a->n_rays > 0) {
    int blocks = div_up(a->n_rays, 8);
    sample_count_kernel<<<blocks, 256, 0, s>>>(*a);
    EO_LAUNCH_CHECK();
  }
  sample_scan_kernel<<<1, 1024, 0, s>>>(*a);
  EO_LAUNCH_CHECK();
  if (a->n_rays > 0) {
    int blocks = div_up(a->n_rays, 8);
    sample_scatter_kernel<<<blocks, 256, 0, s>>>(*a);
    EO_LAUNCH_CHECK();
  }
  return EONERF_OK;
}

extern "C" int eonerf_pack_info(const int64_t* ray_indices, int64_t n_pts, int64_t n_rays, int64_t* ray_offsets,
                                eonerf_stream_t stream) {
  EO_REQUIRE(ray_offsets && (ray_indices || n_pts == 0) && n_rays >= 0, "pack_info: bad arguments");
  pack_info_kernel<<<div_up(n_rays + 1, 256), 256, 0, as_stream(stream)>>>(ray_indices, n_pts, n_rays, ray_offsets);
  EO_LAUNCH_CHECK();
  return EONERF_OK;
}

extern "C" int eonerf_set_last_t_end(float* t_ends, const int64_t* ray_offsets, int64_t n_rays, float value,
                                     eonerf_stream_t stream) {
  EO_REQUIRE(ray_offsets && n_rays >= 0, "set_last_t_end: bad arguments");
  if (n_rays == 0) return EONERF_OK;
  EO_REQUIRE(t_ends, "set_last_t_end: null t_ends");
  set_last_kernel<<<div_up(n_rays, 256), 256, 0, as_stream(stream)>>>(t_ends, ray_offsets, n_rays, value);
  EO_LAUNCH_CHECK();
  return EONERF_OK;
}
