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
  float zm = __fmul_rn(__fadd_rn(ts, te), 0.5f);              // :79  "/ 2.0": halving is exact, the product is bit-identical
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

// ------------------------------------------------------------------------------------------------
// One-pass sampler: count, scan and scatter in a single launch.
//   * a CTA owns a tile of kTileRays consecutive rays (tiles are handed out by an atomic ticket, so a tile never waits for a
//     tile that has not started); warp w evaluates rays w, w+8, ... of the tile ONCE: for n_samples <= 129 the (t_start,
//     t_end) of the <= 4 x 32 intervals of a ray stay in registers together with the keep masks;
//   * the tile's per-ray counts are scanned in shared memory; the tile's global offset comes from a decoupled look-back over
//     64-bit tile descriptors (2 status bits + 62 value bits, one word: no fence between flag and value);
//   * the kept intervals are written in ray order (coalesced per ray: a warp writes consecutive positions).
// HBM traffic: the uniforms are read once (4 B per interval), 16 B are written per kept interval.
// ------------------------------------------------------------------------------------------------
// All intervals of one ray, evaluated by one warp with every z value computed ONCE: lane l of round c owns sample index
// j = 32 c + l; neighbours (z_{j-1}, z_{j+1} for the mid-points of perturb_z_vals, the perturbed z_{j+1} for the interval end)
// come from the adjacent lane by shuffle, across rounds from lane 31 / lane 0 of the adjacent round.  Same operations on the
// same operands in the same order as interval() (which re-derives six z_linear values per lane): bit-identical, ~8x fewer
// instructions.  Requires n <= 32 * kRounds.
template <int kRounds>
__device__ __forceinline__ void eval_ray(const RayGeom& r, const float* __restrict__ zs, const float* __restrict__ u_row, int n, int lane,
                                         float (&ts)[kRounds], float (&te)[kRounds], unsigned (&mask)[kRounds]) {
  float zl[kRounds], uu[kRounds], zp[kRounds];
#pragma unroll
  for (int c = 0; c < kRounds; ++c) {
    const int j = c * 32 + lane;
    const bool ok = j < n;
    uu[c] = ok ? __ldg(u_row + j) : 0.f;
    zl[c] = ok ? z_linear(r, zs, j) : 0.f;
  }
#pragma unroll
  for (int c = 0; c < kRounds; ++c) {
    const int j = c * 32 + lane;
    float prev = __shfl_up_sync(kFull, zl[c], 1), next = __shfl_down_sync(kFull, zl[c], 1);
    if (c > 0) { const float t = __shfl_sync(kFull, zl[c - 1], 31); if (lane == 0) prev = t; }
    if (c + 1 < kRounds) { const float t = __shfl_sync(kFull, zl[c + 1], 0); if (lane == 31) next = t; }
    const float lower = (j == 0) ? zl[c] : __fmul_rn(0.5f, __fadd_rn(prev, zl[c]));
    const float upper = (j == n - 1) ? zl[c] : __fmul_rn(0.5f, __fadd_rn(zl[c], next));
    zp[c] = __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), uu[c]));
  }
#pragma unroll
  for (int c = 0; c < kRounds; ++c) {
    const int i = c * 32 + lane;
    float z1 = __shfl_down_sync(kFull, zp[c], 1);
    if (c + 1 < kRounds) { const float t = __shfl_sync(kFull, zp[c + 1], 0); if (lane == 31) z1 = t; }
    const float z0 = zp[c];
    ts[c] = z0;
    te[c] = __fadd_rn(z0, __fsub_rn(z1, z0));
    const float zm = __fmul_rn(__fadd_rn(ts[c], te[c]), 0.5f);
    const float x = __fadd_rn(r.ox, __fmul_rn(r.dx, zm));
    const float y = __fadd_rn(r.oy, __fmul_rn(r.dy, zm));
    const float z = __fadd_rn(r.oz, __fmul_rn(r.dz, zm));
    const bool keep = (i < n - 1) && !((fabsf(x) >= 1.0f) || (fabsf(y) >= 1.0f) || (fabsf(z) >= 1.0f));
    mask[c] = __ballot_sync(kFull, keep);
  }
}

constexpr int kTileWarps = 8;
constexpr int kRegRounds = 4;                 // cooperative path: up to 4 x 32 sample positions per ray (n_samples <= 128)
constexpr int kMaxTileRays = 1024;            // per-ray counts of a tile live in shared memory
constexpr unsigned long long kDescAggregate = 1ull << 62, kDescPrefix = 2ull << 62, kDescMask = (1ull << 62) - 1;
// scratch words: [0] tile ticket, [1] empty-ray count, [2] finished tiles, [3 ...] tile descriptors

// kept intervals of one ray: masks per round of 32 (and, on the cooperative path, the interval bounds)
template <bool kCoop>
__device__ __forceinline__ int ray_masks(const EonerfSampleArgs& a, const RayGeom& r, const float* u_row, int lane, int rounds,
                                         float (&ts)[kRegRounds], float (&te)[kRegRounds], unsigned (&mask)[kRegRounds]) {
  int cnt = 0;
  if (kCoop) {
    eval_ray<kRegRounds>(r, a.z_steps, u_row, a.n_samples, lane, ts, te, mask);
#pragma unroll
    for (int c = 0; c < kRegRounds; ++c) cnt += __popc(mask[c]);
  } else {
    const int S = a.n_samples - 1;
    for (int c = 0; c < rounds; ++c) {
      const int i = c * 32 + lane;
      float t0, t1;
      const bool keep = (i < S) && interval(r, a.z_steps, u_row, i, a.n_samples, t0, t1);
      cnt += __popc(__ballot_sync(kFull, keep));
    }
  }
  return cnt;
}

// One tile = a contiguous range of rays_per_tile rays; the grid is sized to what the GPU holds at once (a few tiles per SM), so
// the look-back chain is a few hundred descriptors long however many rays there are.  Phase A counts, phase B evaluates again
// (the uniforms come from L2 the second time) and writes: nothing but the per-ray counts has to be kept in between.
template <bool kCoop>
__global__ void __launch_bounds__(kTileWarps * 32, 4) sample_onepass_kernel(EonerfSampleArgs a, int rays_per_tile, int64_t n_tiles) {
  if (a.run_if && *a.run_if == 0) return;
  __shared__ long long s_tile, s_base, s_total;
  __shared__ int s_cnt[kMaxTileRays];
  __shared__ long long s_warp[kTileWarps];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const bool redraw = a.run_if != nullptr;
  unsigned long long* scratch = reinterpret_cast<unsigned long long*>(a.scratch);
  volatile unsigned long long* desc = scratch + 3;
  if (tid == 0) s_tile = (long long)atomicAdd(scratch, 1ull);
  __syncthreads();
  const int64_t tile = s_tile;
  if (tile >= n_tiles) return;
  const int64_t ray0 = tile * rays_per_tile;
  const int n_here = (int)((a.n_rays - ray0) < rays_per_tile ? (a.n_rays - ray0) : rays_per_tile);
  const int S = a.n_samples - 1;
  const int rounds = (S + 31) >> 5;
  float ts[kRegRounds], te[kRegRounds];
  unsigned mask[kRegRounds];

  // ---- phase A: per-ray counts ----
  int empties = 0;
  for (int lr = wid; lr < n_here; lr += kTileWarps) {
    const int64_t ray = ray0 + lr;
    const RayGeom r = load_ray(a, ray);
    const int cnt = ray_masks<kCoop>(a, r, a.u + ray * a.n_samples, lane, rounds, ts, te, mask);
    if (lane == 0) s_cnt[lr] = cnt;
    empties += (cnt == 0);
  }
  __syncthreads();
  // block exclusive scan of the counts: thread t owns rays 4t .. 4t+3 (rays_per_tile <= 1024 = 4 x 256)
  long long c4[4], sum = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) { c4[k] = (4 * tid + k < n_here) ? s_cnt[4 * tid + k] : 0; sum += c4[k]; }
  long long inc = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const long long t = __shfl_up_sync(kFull, inc, o); if (lane >= o) inc += t; }
  if (lane == 31) s_warp[wid] = inc;
  __syncthreads();
  long long before = inc - sum;
  for (int w = 0; w < wid; ++w) before += s_warp[w];

  // ---- tile total -> descriptor; decoupled look-back (warp 0) ----
  if (wid == 0) {
    long long total = 0;
    for (int w = 0; w < kTileWarps; ++w) total += s_warp[w];
    if (lane == 0) {
      if (empties && !redraw) atomicAdd(scratch + 1, (unsigned long long)empties);
      desc[tile] = (tile == 0 ? kDescPrefix : kDescAggregate) | (unsigned long long)total;
    }
    long long base = 0;
    if (tile > 0) {
      int64_t end = tile;                                     // look back 32 tiles at a time: lane l inspects tile end-1-l
      while (true) {
        const int64_t t = end - 1 - lane;
        unsigned long long d = kDescPrefix;                   // "tiles" before 0: a zero prefix
        if (t >= 0) { do { d = desc[t]; } while ((d >> 62) == 0); }
        const unsigned has_prefix = __ballot_sync(kFull, (d >> 62) == 2);
        const int first = __ffs(has_prefix) - 1;              // nearest predecessor holding an inclusive prefix (-1: none)
        long long v = (t >= 0 && (first < 0 || lane <= first)) ? (long long)(d & kDescMask) : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
        base += v;
        if (has_prefix) break;
        end -= 32;
      }
      if (lane == 0) desc[tile] = kDescPrefix | (unsigned long long)(base + total);
    }
    if (lane == 0) {
      s_base = base;
      s_total = total;
      if (tile == n_tiles - 1) a.stats[0] = base + total;
      if (tile == 0) a.ray_offsets[0] = 0;
    }
  }
  // other warps: empties of warps 1.. are added by their lane 0
  if (wid != 0 && lane == 0 && empties && !redraw) atomicAdd(scratch + 1, (unsigned long long)empties);
  // publish this thread's four exclusive offsets (relative to the tile) for phase B
  __syncthreads();
  const long long base = s_base;
  {
    long long run = before;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int lr = 4 * tid + k;
      if (lr < n_here) {
        a.ray_offsets[ray0 + lr + 1] = base + run + c4[k];
        if (!redraw) a.pts_per_ray[ray0 + lr] = (float)c4[k];
      }
      run += c4[k];
    }
  }
  // exclusive offsets back into shared memory (s_cnt is reused: offsets relative to the tile fit in 32 bits: <= 1024 x n)
  __syncthreads();
  {
    long long run = before;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int lr = 4 * tid + k;
      if (lr < n_here) s_cnt[lr] = (int)run;
      run += c4[k];
    }
  }
  __syncthreads();

  // ---- phase B: evaluate again and write the kept intervals in order ----
  for (int lr = wid; lr < n_here; lr += kTileWarps) {
    const int64_t ray = ray0 + lr;
    const RayGeom r = load_ray(a, ray);
    const float* u_row = a.u + ray * a.n_samples;
    int64_t out = base + s_cnt[lr];
    if (kCoop) {
      eval_ray<kRegRounds>(r, a.z_steps, u_row, a.n_samples, lane, ts, te, mask);
#pragma unroll
      for (int c = 0; c < kRegRounds; ++c) {
        const unsigned m = mask[c];
        if ((m >> lane) & 1u) {
          const int64_t p = out + __popc(m & ((1u << lane) - 1u));
          a.ray_indices[p] = ray;
          a.t_starts[p] = ts[c];
          a.t_ends[p] = te[c];
        }
        out += __popc(m);
      }
    } else {
      for (int c = 0; c < rounds; ++c) {
        const int i = c * 32 + lane;
        float t0 = 0.f, t1 = 0.f;
        const bool keep = (i < S) && interval(r, a.z_steps, u_row, i, a.n_samples, t0, t1);
        const unsigned m = __ballot_sync(kFull, keep);
        if (keep) {
          const int64_t p = out + __popc(m & ((1u << lane) - 1u));
          a.ray_indices[p] = ray;
          a.t_starts[p] = t0;
          a.t_ends[p] = t1;
        }
        out += __popc(m);
      }
    }
  }

  // ---- the last tile to finish publishes the empty-ray count ----
  if (!redraw) {
    __syncthreads();
    if (tid == 0) {
      __threadfence();
      const unsigned long long done = atomicAdd(scratch + 2, 1ull);
      if (done == (unsigned long long)(n_tiles - 1)) a.stats[1] = (int64_t)atomicAdd(scratch + 1, 0ull);
    }
  }
}

// tiles of the one-pass sampler: as many as the GPU holds at once (4 CTAs per SM), at most kMaxTileRays rays each
static inline void onepass_tiling(int64_t n_rays, int& rays_per_tile, int64_t& n_tiles) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int64_t rpt = (n_rays + 4 * sms - 1) / (4 * sms);
  rpt = (rpt + kTileWarps - 1) / kTileWarps * kTileWarps;
  if (rpt < kTileWarps) rpt = kTileWarps;
  if (rpt > kMaxTileRays) rpt = kMaxTileRays;
  rays_per_tile = (int)rpt;
  n_tiles = (n_rays + rpt - 1) / rpt;
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

extern "C" int64_t eonerf_sample_scratch_bytes(int64_t n_rays) {
  return (3 + (n_rays + kTileWarps - 1) / kTileWarps) * (int64_t)sizeof(unsigned long long);   // upper bound on the tile count
}

extern "C" int eonerf_sample_compact(const EonerfSampleArgs* a, eonerf_stream_t stream) {
  EO_REQUIRE(a && a->n_samples >= 2 && a->n_rays >= 0, "sample_compact: need n_samples >= 2 and n_rays >= 0");
  EO_REQUIRE(a->ray_offsets && a->stats, "sample_compact: null ray_offsets / stats");
  EO_REQUIRE(a->n_rays == 0 || (a->origins && a->viewdirs && a->u && a->z_steps), "sample_compact: null input");
  EO_REQUIRE(a->n_rays == 0 || (a->ray_indices && a->t_starts && a->t_ends && (a->pts_per_ray || a->run_if)), "sample_compact: null output");
  cudaStream_t s = as_stream(stream);
  if (a->scratch && a->n_rays > 0) {
    EO_REQUIRE(((uintptr_t)a->scratch & 7) == 0, "sample_compact: scratch must be 8-byte aligned");
    int rays_per_tile;
    int64_t n_tiles;
    onepass_tiling(a->n_rays, rays_per_tile, n_tiles);
    EO_CUDA(cudaMemsetAsync(a->scratch, 0, (size_t)(3 + n_tiles) * sizeof(unsigned long long), s));
    if (a->n_samples <= 32 * kRegRounds) sample_onepass_kernel<true><<<(unsigned)n_tiles, kTileWarps * 32, 0, s>>>(*a, rays_per_tile, n_tiles);
    else sample_onepass_kernel<false><<<(unsigned)n_tiles, kTileWarps * 32, 0, s>>>(*a, rays_per_tile, n_tiles);
    EO_LAUNCH_CHECK();
    return EONERF_OK;
  }
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
