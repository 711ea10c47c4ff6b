// Per-ray volume rendering kernels: transmittance / weights / accumulation (the nerfacc v0.5.2 operator
// trio the reference calls), the fused EO-NeRF compositing, the sun-ray shadow pass and the irradiance +
// radiometric epilogue — forward and backward.
//
// Reference call sites (relative to /root/reference):
//   radiance_fields/eonerf.py:186-193,229-246   weights + five accumulate_along_rays + beta_min
//   sat_rendering.py:87-118                      compute_geometric_shadows
//   sat_rendering.py:265-312                     irradiance model, radiometric normalisation, 21-column packing
//
// Layout: samples are packed ray after ray; ray r owns [ray_offsets[r], ray_offsets[r+1]).  One warp owns
// one ray: the 32 lanes read 32 consecutive samples (coalesced 128-byte requests), the per-ray exclusive
// prefix sum of sigma*delta is a shuffle scan with a carry between 32-sample chunks, and every per-ray
// reduction is a butterfly — no atomics, deterministic.
//
// The exclusive scan is a *true* exclusive scan (inclusive scan shifted by one lane): the last interval
// of a camera ray is 1e10 long (eonerf.py:220), "inclusive minus self" would cancel catastrophically.
#include <stdlib.h>

#include "common.cuh"

namespace eonerf {

constexpr int kWarpsPerBlock = 8;

__device__ __forceinline__ int64_t warp_ray() {
  return (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
}

// exclusive prefix (within the warp) of v; `total` receives the warp total
__device__ __forceinline__ float warp_exclusive_sum(float v, int lane, float& total) {
  float inc = warp_inclusive_sum(v, lane);
  total = __shfl_sync(kFull, inc, 31);
  float up = __shfl_up_sync(kFull, inc, 1);
  return lane == 0 ? 0.0f : up;
}

struct SampleW {
  float tau, T, alpha, w;
};

// tau = sigma*(te-ts); T = exp(-prefix); alpha = 1-exp(-tau); w = T*alpha      (nerfacc volrend.py)
__device__ __forceinline__ SampleW sample_weight(float ts, float te, float sigma, float prefix) {
  SampleW s;
  s.tau = sigma * (te - ts);
  s.T = expf(-prefix);
  s.alpha = 1.0f - expf(-s.tau);
  s.w = s.T * s.alpha;
  return s;
}

// ------------------------------------------------------------------------------------------------
// render_weight_from_density / render_transmittance_from_density
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) weights_fwd_kernel(EonerfWeightsFwdArgs a) {
  const int lane = threadIdx.x & 31;
  for (int64_t ray = warp_ray(); ray < a.n_rays; ray += (int64_t)gridDim.x * kWarpsPerBlock) {
    const int64_t beg = a.ray_offsets[ray], end = a.ray_offsets[ray + 1];
    float carry = 0.f;
    for (int64_t base = beg; base < end; base += 128) {       // 4 x 32 samples: all 12 loads of a lane issued up front
      float ts[4], te[4], sg[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int64_t i = base + c * 32 + lane;
        const bool ok = i < end;
        ts[c] = ok ? __ldg(a.t_starts + i) : 0.f;
        te[c] = ok ? __ldg(a.t_ends + i) : 0.f;
        sg[c] = ok ? __ldg(a.sigmas + i) : 0.f;
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int64_t i = base + c * 32 + lane;
        const bool ok = i < end;
        float tot;
        const float tau = sg[c] * (te[c] - ts[c]);
        const float pre = carry + warp_exclusive_sum(ok ? tau : 0.f, lane, tot);
        carry += tot;
        if (ok) {
          const SampleW s = sample_weight(ts[c], te[c], sg[c], pre);
          if (a.weights) a.weights[i] = s.w;
          if (a.trans) a.trans[i] = s.T;
          if (a.alphas) a.alphas[i] = s.alpha;
        }
      }
    }
  }
}

// d/d sigma_j = delta_j * ( ga_j * exp(-tau_j) - sum_{i>j} gT_i * T_i )
//   with gT_i = g_trans_i + g_w_i*alpha_i, ga_i = g_alpha_i + g_w_i*T_i.
// Two forward sweeps: the first one accumulates S = sum_i gT_i*T_i, the second one turns the running
// inclusive prefix into the exclusive suffix S - prefix (all terms are O(1): no 1e10 enters this sum).
__global__ void __launch_bounds__(256) weights_bwd_kernel(EonerfWeightsBwdArgs a) {
  int lane = threadIdx.x & 31;
  int64_t ray = warp_ray();
  if (ray >= a.n_rays) return;
  int64_t beg = a.ray_offsets[ray], end = a.ray_offsets[ray + 1];
  float S = 0.f;
  {
    float carry = 0.f;
    for (int64_t base = beg; base < end; base += 32) {
      int64_t i = base + lane;
      bool ok = i < end;
      float ts = ok ? __ldg(a.t_starts + i) : 0.f, te = ok ? __ldg(a.t_ends + i) : 0.f;
      float sg = ok ? __ldg(a.sigmas + i) : 0.f;
      float tau = sg * (te - ts), tot;
      float pre = carry + warp_exclusive_sum(ok ? tau : 0.f, lane, tot);
      carry += tot;
      float v = 0.f;
      if (ok) {
        SampleW s = sample_weight(ts, te, sg, pre);
        float gT = (a.g_trans ? __ldg(a.g_trans + i) : 0.f) + (a.g_weights ? __ldg(a.g_weights + i) * s.alpha : 0.f);
        v = gT * s.T;
      }
      S += warp_sum(v);
    }
  }
  float carry = 0.f, run = 0.f;
  for (int64_t base = beg; base < end; base += 32) {
    int64_t i = base + lane;
    bool ok = i < end;
    float ts = ok ? __ldg(a.t_starts + i) : 0.f, te = ok ? __ldg(a.t_ends + i) : 0.f;
    float sg = ok ? __ldg(a.sigmas + i) : 0.f;
    float tau = sg * (te - ts), tot;
    float pre = carry + warp_exclusive_sum(ok ? tau : 0.f, lane, tot);
    carry += tot;
    SampleW s = sample_weight(ts, te, sg, pre);
    float gw = (ok && a.g_weights) ? __ldg(a.g_weights + i) : 0.f;
    float gT = ((ok && a.g_trans) ? __ldg(a.g_trans + i) : 0.f) + gw * s.alpha;
    float ga = ((ok && a.g_alphas) ? __ldg(a.g_alphas + i) : 0.f) + gw * s.T;
    float v = ok ? gT * s.T : 0.f;
    float inc = warp_inclusive_sum(v, lane);
    // exclusive suffix sum_{k>i} v_k.  The last sample's suffix is 0 by definition: S - prefix would leave a
    // ~1e-8*S rounding residue there, and its interval is 1e10 long (eonerf.py:220)
    float suffix = (i == end - 1) ? 0.f : S - (run + inc);
    run += __shfl_sync(kFull, inc, 31);
    if (ok) a.g_sigmas[i] = (te - ts) * (ga * expf(-s.tau) - suffix);
  }
}

// ------------------------------------------------------------------------------------------------
// accumulate_along_rays   out[r, c] = sum_{i in r} w_i * v_{i,c}
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) accumulate_fwd_kernel(EonerfAccumFwdArgs a) {
  int lane = threadIdx.x & 31;
  int64_t ray = warp_ray();
  if (ray >= a.n_rays) return;
  int64_t beg = a.ray_offsets[ray], end = a.ray_offsets[ray + 1];
  int C = a.n_channels;
  if (C <= 4) {  // lanes over samples, four chunks of loads in flight
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int64_t base = beg + lane; base < end; base += 128) {
      float w[4], v[4][4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int64_t i = base + 32 * u;
        const bool ok = i < end;
        w[u] = ok ? __ldg(a.weights + i) : 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) v[u][c] = (c < C) ? ((ok && a.values) ? __ldg(a.values + i * C + c) : 1.0f) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[c] += w[u] * v[u][c];
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float s = warp_sum(acc[c]);
      if (c < C && lane == 0) a.out[ray * C + c] = s;
    }
  } else {  // lanes over channels
    for (int c0 = 0; c0 < C; c0 += 32) {
      int c = c0 + lane;
      float acc = 0.f;
      if (c < C)
        for (int64_t i = beg; i < end; ++i) acc += __ldg(a.weights + i) * __ldg(a.values + i * C + c);
      if (c < C) a.out[ray * C + c] = acc;
    }
  }
}

__global__ void __launch_bounds__(256) accumulate_bwd_kernel(EonerfAccumBwdArgs a) {
  int lane = threadIdx.x & 31;
  int64_t ray = warp_ray();
  if (ray >= a.n_rays) return;
  int64_t beg = a.ray_offsets[ray], end = a.ray_offsets[ray + 1];
  int C = a.n_channels;
  for (int64_t i = beg + lane; i < end; i += 32) {
    float w = __ldg(a.weights + i), gw = 0.f;
    for (int c = 0; c < C; ++c) {
      float g = __ldg(a.g_out + ray * C + c);
      gw += g * (a.values ? __ldg(a.values + i * C + c) : 1.0f);
      if (a.g_values) a.g_values[i * C + c] = w * g;
    }
    if (a.g_weights) a.g_weights[i] = gw;
  }
}

// ------------------------------------------------------------------------------------------------
// Fused EO-NeRF compositing (eonerf.py:229-246)
// comp row: 0:3 albedo, 3 depth, 4 beta(+beta_min), 5 transient_s, 6:9 ambient, 9 sum(w), 10:12 zero
// ------------------------------------------------------------------------------------------------
// Rays with at most 128 kept samples (every ray when n_samples <= 129, the reference's setting) take the register path:
// a lane issues all of its (up to 4 x 9) loads back to back before the first dependent instruction, so a warp keeps ~4 KB
// in flight instead of ~1 KB, and the backward needs no second sweep over memory.  Longer rays use the chunk loop.
constexpr int kFastChunks = 4;

struct CompIn {           // one lane's samples of a ray: chunk c holds sample beg + 32 c + lane
  float ts[kFastChunks], te[kFastChunks], sg[kFastChunks], z[kFastChunks];
  float a0[kFastChunks], a1[kFastChunks], a2[kFastChunks], tb[kFastChunks], tsc[kFastChunks];
};

template <class Args>
__device__ __forceinline__ void comp_load(const Args& a, int64_t beg, int64_t end, int lane, CompIn& in) {
#pragma unroll
  for (int c = 0; c < kFastChunks; ++c) {
    const int64_t i = beg + c * 32 + lane;
    const bool ok = i < end;
    in.ts[c] = ok ? __ldg(a.t_starts + i) : 0.f;
    in.te[c] = ok ? __ldg(a.t_ends + i) : 0.f;
    in.sg[c] = ok ? __ldg(a.sigma + i) : 0.f;
    in.z[c] = ok ? __ldg(a.z_mid + i) : 0.f;
    in.a0[c] = (ok && a.albedo) ? __ldg(a.albedo + 3 * i) : 0.f;
    in.a1[c] = (ok && a.albedo) ? __ldg(a.albedo + 3 * i + 1) : 0.f;
    in.a2[c] = (ok && a.albedo) ? __ldg(a.albedo + 3 * i + 2) : 0.f;
    in.tb[c] = (ok && a.transient_beta) ? __ldg(a.transient_beta + i) : 0.f;
    in.tsc[c] = (ok && a.transient_s) ? __ldg(a.transient_s + i) : 0.f;
  }
}

__global__ void __launch_bounds__(256) composite_fwd_kernel(EonerfCompositeFwdArgs a) {
  const int lane = threadIdx.x & 31;
  for (int64_t ray = warp_ray(); ray < a.n_rays; ray += (int64_t)gridDim.x * kWarpsPerBlock) {
    const int64_t beg = a.ray_offsets[ray], end = a.ray_offsets[ray + 1];
    float acc[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};  // albedo3, depth, beta, ts, sumw
    float carry = 0.f;
    if (end - beg <= 32 * kFastChunks) {
      CompIn in;
      comp_load(a, beg, end, lane, in);
#pragma unroll
      for (int c = 0; c < kFastChunks; ++c) {
        const bool ok = beg + c * 32 + lane < end;
        float tot;
        const float tau = in.sg[c] * (in.te[c] - in.ts[c]);
        const float pre = carry + warp_exclusive_sum(ok ? tau : 0.f, lane, tot);
        carry += tot;
        if (ok) {
          const SampleW s = sample_weight(in.ts[c], in.te[c], in.sg[c], pre);
          acc[0] += s.w * in.a0[c]; acc[1] += s.w * in.a1[c]; acc[2] += s.w * in.a2[c];
          acc[3] += s.w * in.z[c];
          acc[4] += s.w * in.tb[c];
          acc[5] += s.w * in.tsc[c];
          acc[6] += s.w;
        }
      }
    } else {
      for (int64_t base = beg; base < end; base += 32) {
        int64_t i = base + lane;
        bool ok = i < end;
        float ts = ok ? __ldg(a.t_starts + i) : 0.f, te = ok ? __ldg(a.t_ends + i) : 0.f;
        float sg = ok ? __ldg(a.sigma + i) : 0.f;
        float tau = sg * (te - ts), tot;
        float pre = carry + warp_exclusive_sum(ok ? tau : 0.f, lane, tot);
        carry += tot;
        if (ok) {
          SampleW s = sample_weight(ts, te, sg, pre);
          if (a.albedo) {
            acc[0] += s.w * __ldg(a.albedo + 3 * i);
            acc[1] += s.w * __ldg(a.albedo + 3 * i + 1);
            acc[2] += s.w * __ldg(a.albedo + 3 * i + 2);
          }
          acc[3] += s.w * __ldg(a.z_mid + i);
          if (a.transient_beta) acc[4] += s.w * __ldg(a.transient_beta + i);
          if (a.transient_s) acc[5] += s.w * __ldg(a.transient_s + i);
          acc[6] += s.w;
        }
      }
    }
#pragma unroll
    for (int c = 0; c < 7; ++c) acc[c] = warp_sum(acc[c]);
    if (lane == 0) {
      float* o = a.comp + ray * EONERF_COMP_COLS;
      // ambient is constant along the ray (sun direction is per ray): sum_i w_i*a = a*sum_i w_i
      float am[3];
      for (int c = 0; c < 3; ++c) am[c] = a.ambient_ray ? acc[6] * __ldg(a.ambient_ray + 3 * ray + c) : 0.f;
      *reinterpret_cast<float4*>(o) = make_float4(acc[0], acc[1], acc[2], acc[3]);
      *reinterpret_cast<float4*>(o + 4) = make_float4(acc[4] + a.beta_min, acc[5], am[0], am[1]);
      *reinterpret_cast<float4*>(o + 8) = make_float4(am[2], acc[6], 0.f, 0.f);
    }
  }
}

__global__ void __launch_bounds__(256) composite_bwd_kernel(EonerfCompositeBwdArgs a) {
  const int lane = threadIdx.x & 31;
  for (int64_t ray = warp_ray(); ray < a.n_rays; ray += (int64_t)gridDim.x * kWarpsPerBlock) {
    const int64_t beg = a.ray_offsets[ray], end = a.ray_offsets[ray + 1];
    const float* g = a.g_comp + ray * EONERF_COMP_COLS;
    const float4 g03 = __ldg(reinterpret_cast<const float4*>(g)), g47 = __ldg(reinterpret_cast<const float4*>(g + 4)),
                 g8b = __ldg(reinterpret_cast<const float4*>(g + 8));
    const float ga0 = g03.x, ga1 = g03.y, ga2 = g03.z, gd = g03.w, gb = g47.x, gs = g47.y;
    const float gam[3] = {g47.z, g47.w, g8b.x};
    // per-ray constant part of G_i: ambient and the sum-of-weights column
    float gconst = g8b.y;
    if (a.ambient_ray)
      for (int c = 0; c < 3; ++c) gconst += gam[c] * __ldg(a.ambient_ray + 3 * ray + c);

    if (end - beg <= 32 * kFastChunks) {
      CompIn in;
      comp_load(a, beg, end, lane, in);
      float w[kFastChunks], T1[kFastChunks], v[kFastChunks], G[kFastChunks];
      float S = 0.f, sumw = 0.f, carry = 0.f;
#pragma unroll
      for (int c = 0; c < kFastChunks; ++c) {
        const bool ok = beg + c * 32 + lane < end;
        float tot;
        const float tau = in.sg[c] * (in.te[c] - in.ts[c]);
        const float pre = carry + warp_exclusive_sum(ok ? tau : 0.f, lane, tot);
        carry += tot;
        const SampleW s = sample_weight(in.ts[c], in.te[c], in.sg[c], pre);
        G[c] = gconst + gd * in.z[c] + ga0 * in.a0[c] + ga1 * in.a1[c] + ga2 * in.a2[c] + gb * in.tb[c] + gs * in.tsc[c];
        w[c] = ok ? s.w : 0.f;
        T1[c] = s.T * expf(-s.tau);                       // T_{j+1} = T_j * exp(-tau_j)
        v[c] = ok ? G[c] * s.w : 0.f;
        S += v[c];
        sumw += w[c];
      }
      S = warp_sum(S);
      sumw = warp_sum(sumw);
      if (lane == 0 && a.g_ambient_ray)
        for (int c = 0; c < 3; ++c) a.g_ambient_ray[3 * ray + c] = gam[c] * sumw;
      float run = 0.f;
#pragma unroll
      for (int c = 0; c < kFastChunks; ++c) {
        const int64_t i = beg + c * 32 + lane;
        const float inc = warp_inclusive_sum(v[c], lane);
        const float suffix = (i == end - 1) ? 0.f : S - (run + inc);   // exact 0 for the 1e10-long last interval (see weights_bwd)
        run += __shfl_sync(kFull, inc, 31);
        if (i < end) {
          a.g_sigma[i] = (in.te[c] - in.ts[c]) * (G[c] * T1[c] - suffix);
          if (a.g_albedo) {
            a.g_albedo[3 * i] = w[c] * ga0;
            a.g_albedo[3 * i + 1] = w[c] * ga1;
            a.g_albedo[3 * i + 2] = w[c] * ga2;
          }
          if (a.g_transient_beta) a.g_transient_beta[i] = w[c] * gb;
          if (a.g_transient_s) a.g_transient_s[i] = w[c] * gs;
        }
      }
      continue;
    }

    auto G_of = [&](int64_t i) {
      float G = gconst + gd * __ldg(a.z_mid + i);
      if (a.albedo) G += ga0 * __ldg(a.albedo + 3 * i) + ga1 * __ldg(a.albedo + 3 * i + 1) + ga2 * __ldg(a.albedo + 3 * i + 2);
      if (a.transient_beta) G += gb * __ldg(a.transient_beta + i);
      if (a.transient_s) G += gs * __ldg(a.transient_s + i);
      return G;
    };

    float S = 0.f, sumw = 0.f;
    {
      float carry = 0.f;
      for (int64_t base = beg; base < end; base += 32) {
        int64_t i = base + lane;
        bool ok = i < end;
        float ts = ok ? __ldg(a.t_starts + i) : 0.f, te = ok ? __ldg(a.t_ends + i) : 0.f;
        float sg = ok ? __ldg(a.sigma + i) : 0.f;
        float tau = sg * (te - ts), tot;
        float pre = carry + warp_exclusive_sum(ok ? tau : 0.f, lane, tot);
        carry += tot;
        float v = 0.f, w = 0.f;
        if (ok) {
          SampleW s = sample_weight(ts, te, sg, pre);
          w = s.w;
          v = G_of(i) * s.w;
        }
        S += warp_sum(v);
        sumw += warp_sum(w);
      }
    }
    if (lane == 0 && a.g_ambient_ray)
      for (int c = 0; c < 3; ++c) a.g_ambient_ray[3 * ray + c] = gam[c] * sumw;

    float carry = 0.f, run = 0.f;
    for (int64_t base = beg; base < end; base += 32) {
      int64_t i = base + lane;
      bool ok = i < end;
      float ts = ok ? __ldg(a.t_starts + i) : 0.f, te = ok ? __ldg(a.t_ends + i) : 0.f;
      float sg = ok ? __ldg(a.sigma + i) : 0.f;
      float tau = sg * (te - ts), tot;
      float pre = carry + warp_exclusive_sum(ok ? tau : 0.f, lane, tot);
      carry += tot;
      SampleW s = sample_weight(ts, te, sg, pre);
      float G = ok ? G_of(i) : 0.f;
      float v = ok ? G * s.w : 0.f;
      float inc = warp_inclusive_sum(v, lane);
      float suffix = (i == end - 1) ? 0.f : S - (run + inc);   // exact 0 for the 1e10-long last interval (see weights_bwd)
      run += __shfl_sync(kFull, inc, 31);
      if (ok) {
        // T_{j+1} = T_j * exp(-tau_j)
        a.g_sigma[i] = (te - ts) * (G * s.T * expf(-s.tau) - suffix);
        if (a.g_albedo) {
          a.g_albedo[3 * i] = s.w * ga0;
          a.g_albedo[3 * i + 1] = s.w * ga1;
          a.g_albedo[3 * i + 2] = s.w * ga2;
        }
        if (a.g_transient_beta) a.g_transient_beta[i] = s.w * gb;
        if (a.g_transient_s) a.g_transient_s[i] = s.w * gs;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// 128-bit vectorised forms.  A ray's samples are read through 16-byte-aligned windows of 128 samples: window w starts at
// (beg & ~3) + 128 w, lane l owns the four consecutive samples s = start + 4 l .. s + 3 and fetches them with ONE 128-bit load
// per array (three for the [P,3] albedo: its 12 floats are contiguous and 48-byte aligned).  Samples of the neighbouring rays
// that fall into the first / last group are masked; the per-ray exclusive scan is a 4-step serial prefix inside the lane plus
// a warp shuffle scan of the lane totals.  Used when every array is 16-byte aligned (checked on the host), else the scalar forms.
// ------------------------------------------------------------------------------------------------
struct Quad { float v[4]; };
__device__ __forceinline__ Quad ldq(const float* __restrict__ p, int64_t s, int64_t n_pts) {
  Quad q;
  if (s + 4 <= n_pts) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p + s));
    q.v[0] = t.x; q.v[1] = t.y; q.v[2] = t.z; q.v[3] = t.w;
  } else {
#pragma unroll
    for (int e = 0; e < 4; ++e) q.v[e] = (s + e < n_pts) ? __ldg(p + s + e) : 0.f;
  }
  return q;
}
__device__ __forceinline__ Quad zero_quad() { Quad q; q.v[0] = q.v[1] = q.v[2] = q.v[3] = 0.f; return q; }

// transmittance in front of each of the lane's four samples (exclusive scan over the ray), carry across windows
struct QuadScan { float pre[4]; };
__device__ __forceinline__ QuadScan quad_exclusive(const float (&tau)[4], int lane, float& carry) {
  const float p1 = tau[0], p2 = p1 + tau[1], p3 = p2 + tau[2], tot_lane = p3 + tau[3];
  float tot;
  const float ex = carry + warp_exclusive_sum(tot_lane, lane, tot);
  carry += tot;
  QuadScan r;
  r.pre[0] = ex; r.pre[1] = ex + p1; r.pre[2] = ex + p2; r.pre[3] = ex + p3;
  return r;
}

__global__ void __launch_bounds__(256) weights_fwd_vec_kernel(EonerfWeightsFwdArgs a) {
  const int lane = threadIdx.x & 31;
  for (int64_t ray = warp_ray(); ray < a.n_rays; ray += (int64_t)gridDim.x * kWarpsPerBlock) {
    const int64_t beg = a.ray_offsets[ray], end = a.ray_offsets[ray + 1];
    float carry = 0.f;
    for (int64_t wb = beg & ~(int64_t)3; wb < end; wb += 128) {
      const int64_t s = wb + 4 * lane;
      const bool any = s < end;
      const Quad ts = any ? ldq(a.t_starts, s, a.n_pts) : zero_quad(), te = any ? ldq(a.t_ends, s, a.n_pts) : zero_quad(),
                 sg = any ? ldq(a.sigmas, s, a.n_pts) : zero_quad();
      float tau[4];
      bool ok[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) { ok[e] = s + e >= beg && s + e < end; tau[e] = ok[e] ? sg.v[e] * (te.v[e] - ts.v[e]) : 0.f; }
      const QuadScan sc = quad_exclusive(tau, lane, carry);
      float w[4], T[4], al[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) { const SampleW x = sample_weight(ts.v[e], te.v[e], sg.v[e], sc.pre[e]); w[e] = x.w; T[e] = x.T; al[e] = x.alpha; }
      if (ok[0] && ok[3] && s + 4 <= a.n_pts) {                 // whole group inside the ray: 128-bit stores
        if (a.weights) *reinterpret_cast<float4*>(a.weights + s) = make_float4(w[0], w[1], w[2], w[3]);
        if (a.trans) *reinterpret_cast<float4*>(a.trans + s) = make_float4(T[0], T[1], T[2], T[3]);
        if (a.alphas) *reinterpret_cast<float4*>(a.alphas + s) = make_float4(al[0], al[1], al[2], al[3]);
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (ok[e]) {
            if (a.weights) a.weights[s + e] = w[e];
            if (a.trans) a.trans[s + e] = T[e];
            if (a.alphas) a.alphas[s + e] = al[e];
          }
      }
    }
  }
}

// C == 3 (or values == NULL, C == 1): out[r, c] = sum_i w_i v_{i,c}
__global__ void __launch_bounds__(256) accumulate_fwd_vec_kernel(EonerfAccumFwdArgs a) {
  const int lane = threadIdx.x & 31;
  const int C = a.n_channels;
  for (int64_t ray = warp_ray(); ray < a.n_rays; ray += (int64_t)gridDim.x * kWarpsPerBlock) {
    const int64_t beg = a.ray_offsets[ray], end = a.ray_offsets[ray + 1];
    float acc[3] = {0.f, 0.f, 0.f};
    for (int64_t wb = beg & ~(int64_t)3; wb < end; wb += 128) {
      const int64_t s = wb + 4 * lane;
      if (s >= end) continue;
      const Quad w = ldq(a.weights, s, a.n_pts);
      if (a.values && C == 3) {
        const Quad v0 = ldq(a.values, 3 * s, 3 * a.n_pts), v1 = ldq(a.values, 3 * s + 4, 3 * a.n_pts), v2 = ldq(a.values, 3 * s + 8, 3 * a.n_pts);
        const float vv[12] = {v0.v[0], v0.v[1], v0.v[2], v0.v[3], v1.v[0], v1.v[1], v1.v[2], v1.v[3], v2.v[0], v2.v[1], v2.v[2], v2.v[3]};
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (s + e >= beg && s + e < end) { acc[0] += w.v[e] * vv[3 * e]; acc[1] += w.v[e] * vv[3 * e + 1]; acc[2] += w.v[e] * vv[3 * e + 2]; }
      } else {
        const Quad v = a.values ? ldq(a.values, s, a.n_pts) : zero_quad();
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (s + e >= beg && s + e < end) acc[0] += w.v[e] * (a.values ? v.v[e] : 1.0f);
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float t = warp_sum(acc[c]);
      if (c < C && lane == 0) a.out[ray * C + c] = t;
    }
  }
}

struct CompQuad { Quad ts, te, sg, z, al0, al1, al2, tb, tsc; };
template <class Args>
__device__ __forceinline__ CompQuad comp_ldq(const Args& a, int64_t s) {
  CompQuad q;
  q.ts = ldq(a.t_starts, s, a.n_pts); q.te = ldq(a.t_ends, s, a.n_pts); q.sg = ldq(a.sigma, s, a.n_pts); q.z = ldq(a.z_mid, s, a.n_pts);
  if (a.albedo) { q.al0 = ldq(a.albedo, 3 * s, 3 * a.n_pts); q.al1 = ldq(a.albedo, 3 * s + 4, 3 * a.n_pts); q.al2 = ldq(a.albedo, 3 * s + 8, 3 * a.n_pts); }
  else { q.al0 = zero_quad(); q.al1 = zero_quad(); q.al2 = zero_quad(); }
  q.tb = a.transient_beta ? ldq(a.transient_beta, s, a.n_pts) : zero_quad();
  q.tsc = a.transient_s ? ldq(a.transient_s, s, a.n_pts) : zero_quad();
  return q;
}
// albedo channel c of the lane's sample e: element 3 e + c of the 12 contiguous floats
__device__ __forceinline__ float alb_of(const CompQuad& q, int e, int c) {
  const int k = 3 * e + c;
  return k < 4 ? q.al0.v[k] : (k < 8 ? q.al1.v[k - 4] : q.al2.v[k - 8]);
}

__global__ void __launch_bounds__(256) composite_fwd_vec_kernel(EonerfCompositeFwdArgs a) {
  const int lane = threadIdx.x & 31;
  for (int64_t ray = warp_ray(); ray < a.n_rays; ray += (int64_t)gridDim.x * kWarpsPerBlock) {
    const int64_t beg = a.ray_offsets[ray], end = a.ray_offsets[ray + 1];
    float acc[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};  // albedo3, depth, beta, ts, sumw
    float carry = 0.f;
    for (int64_t wb = beg & ~(int64_t)3; wb < end; wb += 128) {
      const int64_t s = wb + 4 * lane;
      const bool any = s < end;
      CompQuad q;
      if (any) q = comp_ldq(a, s);
      float tau[4];
      bool ok[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) { ok[e] = any && s + e >= beg && s + e < end; tau[e] = ok[e] ? q.sg.v[e] * (q.te.v[e] - q.ts.v[e]) : 0.f; }
      const QuadScan sc = quad_exclusive(tau, lane, carry);
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (ok[e]) {
          const SampleW x = sample_weight(q.ts.v[e], q.te.v[e], q.sg.v[e], sc.pre[e]);
          acc[0] += x.w * alb_of(q, e, 0); acc[1] += x.w * alb_of(q, e, 1); acc[2] += x.w * alb_of(q, e, 2);
          acc[3] += x.w * q.z.v[e];
          acc[4] += x.w * q.tb.v[e];
          acc[5] += x.w * q.tsc.v[e];
          acc[6] += x.w;
        }
    }
#pragma unroll
    for (int c = 0; c < 7; ++c) acc[c] = warp_sum(acc[c]);
    if (lane == 0) {
      float* o = a.comp + ray * EONERF_COMP_COLS;
      float am[3];
      for (int c = 0; c < 3; ++c) am[c] = a.ambient_ray ? acc[6] * __ldg(a.ambient_ray + 3 * ray + c) : 0.f;
      *reinterpret_cast<float4*>(o) = make_float4(acc[0], acc[1], acc[2], acc[3]);
      *reinterpret_cast<float4*>(o + 4) = make_float4(acc[4] + a.beta_min, acc[5], am[0], am[1]);
      *reinterpret_cast<float4*>(o + 8) = make_float4(am[2], acc[6], 0.f, 0.f);
    }
  }
}

// two sweeps over the windows (S = sum_i G_i w_i first, then the exclusive suffixes): the second sweep's loads hit L1 / L2
__global__ void __launch_bounds__(256) composite_bwd_vec_kernel(EonerfCompositeBwdArgs a) {
  const int lane = threadIdx.x & 31;
  for (int64_t ray = warp_ray(); ray < a.n_rays; ray += (int64_t)gridDim.x * kWarpsPerBlock) {
    const int64_t beg = a.ray_offsets[ray], end = a.ray_offsets[ray + 1];
    const float* g = a.g_comp + ray * EONERF_COMP_COLS;
    const float4 g03 = __ldg(reinterpret_cast<const float4*>(g)), g47 = __ldg(reinterpret_cast<const float4*>(g + 4)),
                 g8b = __ldg(reinterpret_cast<const float4*>(g + 8));
    const float ga[3] = {g03.x, g03.y, g03.z}, gd = g03.w, gb = g47.x, gs = g47.y;
    const float gam[3] = {g47.z, g47.w, g8b.x};
    float gconst = g8b.y;
    if (a.ambient_ray)
      for (int c = 0; c < 3; ++c) gconst += gam[c] * __ldg(a.ambient_ray + 3 * ray + c);
    float S = 0.f, sumw = 0.f;
    for (int sweep = 0; sweep < 2; ++sweep) {
      float carry = 0.f, run = 0.f;
      for (int64_t wb = beg & ~(int64_t)3; wb < end; wb += 128) {
        const int64_t s = wb + 4 * lane;
        const bool any = s < end;
        CompQuad q;
        if (any) q = comp_ldq(a, s);
        float tau[4];
        bool ok[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) { ok[e] = any && s + e >= beg && s + e < end; tau[e] = ok[e] ? q.sg.v[e] * (q.te.v[e] - q.ts.v[e]) : 0.f; }
        const QuadScan sc = quad_exclusive(tau, lane, carry);
        float G[4], w[4], T1[4], v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          G[e] = w[e] = T1[e] = v[e] = 0.f;
          if (ok[e]) {
            const SampleW x = sample_weight(q.ts.v[e], q.te.v[e], q.sg.v[e], sc.pre[e]);
            G[e] = gconst + gd * q.z.v[e] + ga[0] * alb_of(q, e, 0) + ga[1] * alb_of(q, e, 1) + ga[2] * alb_of(q, e, 2) + gb * q.tb.v[e] + gs * q.tsc.v[e];
            w[e] = x.w;
            T1[e] = x.T * expf(-x.tau);                       // T_{j+1} = T_j exp(-tau_j)
            v[e] = G[e] * x.w;
          }
        }
        if (sweep == 0) {
          S += (v[0] + v[1]) + (v[2] + v[3]);
          sumw += (w[0] + w[1]) + (w[2] + w[3]);
          continue;
        }
        // inclusive prefix of v inside the ray, then suffix_i = S - prefix_i (exactly 0 for the 1e10-long last interval)
        const float i0 = v[0], i1 = i0 + v[1], i2 = i1 + v[2], i3 = i2 + v[3];
        float tot;
        const float ex = run + warp_exclusive_sum(i3, lane, tot);
        run += tot;
        const float inc[4] = {ex + i0, ex + i1, ex + i2, ex + i3};
        float gsig[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float suffix = (s + e == end - 1) ? 0.f : S - inc[e];
          gsig[e] = (q.te.v[e] - q.ts.v[e]) * (G[e] * T1[e] - suffix);
        }
        if (ok[0] && ok[3] && s + 4 <= a.n_pts) {
          *reinterpret_cast<float4*>(a.g_sigma + s) = make_float4(gsig[0], gsig[1], gsig[2], gsig[3]);
          if (a.g_albedo) {
            float4* ga4 = reinterpret_cast<float4*>(a.g_albedo + 3 * s);
            ga4[0] = make_float4(w[0] * ga[0], w[0] * ga[1], w[0] * ga[2], w[1] * ga[0]);
            ga4[1] = make_float4(w[1] * ga[1], w[1] * ga[2], w[2] * ga[0], w[2] * ga[1]);
            ga4[2] = make_float4(w[2] * ga[2], w[3] * ga[0], w[3] * ga[1], w[3] * ga[2]);
          }
          if (a.g_transient_beta) *reinterpret_cast<float4*>(a.g_transient_beta + s) = make_float4(w[0] * gb, w[1] * gb, w[2] * gb, w[3] * gb);
          if (a.g_transient_s) *reinterpret_cast<float4*>(a.g_transient_s + s) = make_float4(w[0] * gs, w[1] * gs, w[2] * gs, w[3] * gs);
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (ok[e]) {
              a.g_sigma[s + e] = gsig[e];
              if (a.g_albedo) { a.g_albedo[3 * (s + e)] = w[e] * ga[0]; a.g_albedo[3 * (s + e) + 1] = w[e] * ga[1]; a.g_albedo[3 * (s + e) + 2] = w[e] * ga[2]; }
              if (a.g_transient_beta) a.g_transient_beta[s + e] = w[e] * gb;
              if (a.g_transient_s) a.g_transient_s[s + e] = w[e] * gs;
            }
        }
      }
      if (sweep == 0) {
        S = warp_sum(S);
        sumw = warp_sum(sumw);
        if (lane == 0 && a.g_ambient_ray)
          for (int c = 0; c < 3; ++c) a.g_ambient_ray[3 * ray + c] = gam[c] * sumw;
      }
    }
  }
}

static inline bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }
// Measured at 65 536 rays x 128 samples (profiles/r2c_render_vec_ab.log): the 128-bit forms win only where the scalar form was
// instruction-limited (accumulate_fwd: 60 % -> 68 % of the HBM peak); compositing / weights are bound by the per-ray scan + exp
// chain and by register pressure (4 samples per lane: 118 registers), not by the load width, and the scalar forms with four
// 32-sample chunks in flight stay faster (composite_fwd 69 % vs 61 %, composite_bwd 73 % vs 44 %, weights_fwd 71 % vs 66 %).
// EONERF_RENDER_VEC: 1 (default) = accumulate_fwd only, 2 = every vectorised form (A/B timing), 0 = none.
static inline int render_vec_level() {
  static const int lvl = [] { const char* e = getenv("EONERF_RENDER_VEC"); return e ? atoi(e) : 1; }();
  return lvl;
}
static inline bool render_vec_enabled() { return render_vec_level() >= 2; }

// ------------------------------------------------------------------------------------------------
// Sun-ray shadow pass (sat_rendering.py:87-118)
// ------------------------------------------------------------------------------------------------
__global__ void sun_rays_kernel(EonerfSunRaysArgs a) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= a.n_rays) return;
  const float* o = a.origins + r * a.origins_stride;
  const float* d = a.viewdirs + r * a.viewdirs_stride;
  const float* s = a.sundirs + r * a.sundirs_stride;
  float depth = __ldg(a.depth + r * a.depth_stride);
  float* out = a.sun_rays + 6 * r;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    out[c] = __fadd_rn(__ldg(o + c), __fmul_rn(depth, __ldg(d + c)));   // :90 (two separately rounded ops)
    out[3 + c] = __fmul_rn(-1.0f, __ldg(s + c));                        // :91
  }
}

// geo_shadow[r] = transmittance in front of the LAST kept sample of sun ray r (:106-116), 1 if none
__global__ void __launch_bounds__(256) shadow_fwd_kernel(EonerfShadowFwdArgs a) {
  const int lane = threadIdx.x & 31;
  for (int64_t ray = warp_ray(); ray < a.n_rays; ray += (int64_t)gridDim.x * kWarpsPerBlock) {
    const int64_t beg = a.ray_offsets[ray], end = a.ray_offsets[ray + 1];
    float acc = 0.f;
    int64_t i = beg + lane;
    for (; i + 96 < end - 1; i += 128) {                  // 12 loads in flight per lane
      float sg[4], te[4], ts[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) { sg[c] = __ldg(a.sigma + i + 32 * c); te[c] = __ldg(a.t_ends + i + 32 * c); ts[c] = __ldg(a.t_starts + i + 32 * c); }
#pragma unroll
      for (int c = 0; c < 4; ++c) acc += sg[c] * (te[c] - ts[c]);
    }
    {
      float sg[4], te[4], ts[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const bool ok = i + 32 * c < end - 1;
        sg[c] = ok ? __ldg(a.sigma + i + 32 * c) : 0.f; te[c] = ok ? __ldg(a.t_ends + i + 32 * c) : 0.f; ts[c] = ok ? __ldg(a.t_starts + i + 32 * c) : 0.f;
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) acc += sg[c] * (te[c] - ts[c]);
    }
    acc = warp_sum(acc);
    if (lane == 0) a.geo_shadow[ray] = (end > beg) ? expf(-acc) : 1.0f;
  }
}

__global__ void __launch_bounds__(256) shadow_bwd_kernel(EonerfShadowBwdArgs a) {
  const int lane = threadIdx.x & 31;
  for (int64_t ray = warp_ray(); ray < a.n_rays; ray += (int64_t)gridDim.x * kWarpsPerBlock) {
    const int64_t beg = a.ray_offsets[ray], end = a.ray_offsets[ray + 1];
    const float k = -__ldg(a.geo_shadow + ray) * __ldg(a.g_geo_shadow + ray);
    for (int64_t i = beg + lane; i < end; i += 128) {
      float te[4], ts[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const bool ok = i + 32 * c < end;
        te[c] = ok ? __ldg(a.t_ends + i + 32 * c) : 0.f; ts[c] = ok ? __ldg(a.t_starts + i + 32 * c) : 0.f;
      }
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (i + 32 * c < end) a.g_sigma[i + 32 * c] = (i + 32 * c < end - 1) ? k * (te[c] - ts[c]) : 0.f;
    }
  }
}

__global__ void __launch_bounds__(256) sun_origin_bwd_kernel(EonerfSunOriginBwdArgs a) {
  int lane = threadIdx.x & 31;
  int64_t ray = warp_ray();
  if (ray >= a.n_rays) return;
  int64_t beg = a.ray_offsets[ray], end = a.ray_offsets[ray + 1];
  float gx = 0.f, gy = 0.f, gz = 0.f;
  for (int64_t i = beg + lane; i < end; i += 32) {
    gx += __ldg(a.g_x + 3 * i);
    gy += __ldg(a.g_x + 3 * i + 1);
    gz += __ldg(a.g_x + 3 * i + 2);
  }
  gx = warp_sum(gx); gy = warp_sum(gy); gz = warp_sum(gz);
  if (lane == 0) {
    const float* d = a.viewdirs + ray * a.viewdirs_stride;
    a.g_depth[ray * a.g_depth_stride] += gx * __ldg(d) + gy * __ldg(d + 1) + gz * __ldg(d + 2);
  }
}

// ------------------------------------------------------------------------------------------------
// Irradiance + radiometric epilogue (sat_rendering.py:265-312)
// ------------------------------------------------------------------------------------------------
struct Radiometric {
  float A[3], b[3];
  int64_t img;
};

template <class Args>
__device__ __forceinline__ Radiometric load_radiometric(const Args& a, int64_t r) {
  Radiometric m;
  m.img = a.eval_mode ? __ldg(a.img_idx) : __ldg(a.img_idx + r * a.img_idx_stride);   // :288-291
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    m.A[c] = a.radiometric ? __ldg(a.radiometric + m.img * 9 + c) : 1.0f;
    m.b[c] = a.radiometric ? __ldg(a.radiometric + m.img * 9 + 3 + c) : 0.0f;
  }
  return m;
}

__global__ void epilogue_fwd_kernel(EonerfEpilogueFwdArgs a) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= a.n_rays) return;
  const float* c = a.comp + r * EONERF_COMP_COLS;
  float* o = a.out + r * EONERF_OUT_COLS;
  Radiometric m = load_radiometric(a, r);
  float ts = c[5];
  float geo = a.geo_shadow ? __ldg(a.geo_shadow + r) : 1.0f;
  float s = a.geo_shadow ? geo * ts : 1.0f;                       // :269-276
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float alb = c[k], amb = c[6 + k] * 0.2f;                       // :265
    float rgb = alb * s + (1.0f - s) * (amb * alb);                // :294
    rgb = m.A[k] * rgb + m.b[k];                                   // :304
    o[k] = fminf(fmaxf(rgb, 0.0f), 1.0f);                          // :305
    o[4 + k] = alb;
    o[7 + k] = amb;
    o[18 + k] = m.A[k] * alb + m.b[k];                             // :306
  }
  o[3] = c[3];
  o[10] = geo;
  o[11] = ts;
  o[12] = c[4];
  o[13] = 1.0f;                                                    // entropy (eonerf.py:246)
  o[14] = __ldg(a.pts_per_ray + r);
  o[15] = a.sc_pts_per_ray ? __ldg(a.sc_pts_per_ray + r) : 1.0f;   // :272
  o[16] = 1.0f; o[17] = 1.0f;                                      // opacity_after_surface (:283)
}

__global__ void epilogue_bwd_kernel(EonerfEpilogueBwdArgs a) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= a.n_rays) return;
  const float* c = a.comp + r * EONERF_COMP_COLS;
  const float* go = a.g_out + r * EONERF_OUT_COLS;
  float* gc = a.g_comp + r * EONERF_COMP_COLS;
  Radiometric m = load_radiometric(a, r);
  float ts = c[5];
  float geo = a.geo_shadow ? __ldg(a.geo_shadow + r) : 1.0f;
  float s = a.geo_shadow ? geo * ts : 1.0f;
  float g_s = 0.f;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float alb = c[k], amb = c[6 + k] * 0.2f;
    float rgb0 = alb * s + (1.0f - s) * (amb * alb);
    float pre = m.A[k] * rgb0 + m.b[k];
    float g_pre = (pre >= 0.0f && pre <= 1.0f) ? __ldg(go + k) : 0.0f;   // clip backward, bounds inclusive
    float g_sl = __ldg(go + 18 + k);
    float g_rgb0 = g_pre * m.A[k];
    gc[k] = g_rgb0 * (s + (1.0f - s) * amb) + __ldg(go + 4 + k) + g_sl * m.A[k];
    gc[6 + k] = 0.2f * (g_rgb0 * (1.0f - s) * alb + __ldg(go + 7 + k));
    g_s += g_rgb0 * alb * (1.0f - amb);
    if (a.g_radiometric && a.radiometric) {
      atomicAdd(a.g_radiometric + m.img * 9 + k, g_pre * rgb0 + g_sl * alb);
      atomicAdd(a.g_radiometric + m.img * 9 + 3 + k, g_pre + g_sl);
    }
  }
  gc[3] = __ldg(go + 3);
  gc[4] = __ldg(go + 12);
  if (a.geo_shadow) {
    gc[5] = g_s * geo + __ldg(go + 11);
    if (a.g_geo_shadow) a.g_geo_shadow[r] = g_s * ts + __ldg(go + 10);
  } else {
    gc[5] = __ldg(go + 11);
  }
  gc[9] = 0.f; gc[10] = 0.f; gc[11] = 0.f;
}

}  // namespace eonerf

using namespace eonerf;

#define RAY_BLOCKS(n) div_up((n), kWarpsPerBlock)
// grid-stride kernels (compositing, shadows): at most 8 resident blocks per SM, each warp then walks several rays
static inline int ray_blocks_capped(int64_t n_rays) {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  const int64_t want = div_up(n_rays, kWarpsPerBlock), cap = (int64_t)sms * 8;
  return (int)(want < cap ? want : cap);
}

extern "C" int eonerf_weights_fwd(const EonerfWeightsFwdArgs* a, eonerf_stream_t stream) {
  EO_REQUIRE(a && a->ray_offsets && a->n_rays >= 0, "weights_fwd: bad arguments");
  if (a->n_rays == 0) return EONERF_OK;
  EO_REQUIRE(a->n_pts == 0 || (a->t_starts && a->t_ends && a->sigmas), "weights_fwd: null input");
  if (render_vec_enabled() && aligned16(a->t_starts) && aligned16(a->t_ends) && aligned16(a->sigmas) && aligned16(a->weights) && aligned16(a->trans) &&
      aligned16(a->alphas))
    weights_fwd_vec_kernel<<<ray_blocks_capped(a->n_rays), 256, 0, as_stream(stream)>>>(*a);
  else
    weights_fwd_kernel<<<ray_blocks_capped(a->n_rays), 256, 0, as_stream(stream)>>>(*a);
  EO_LAUNCH_CHECK();
  return EONERF_OK;
}

extern "C" int eonerf_weights_bwd(const EonerfWeightsBwdArgs* a, eonerf_stream_t stream) {
  EO_REQUIRE(a && a->ray_offsets && a->n_rays >= 0, "weights_bwd: bad arguments");
  if (a->n_rays == 0) return EONERF_OK;
  EO_REQUIRE(a->n_pts == 0 || (a->t_starts && a->t_ends && a->sigmas && a->g_sigmas), "weights_bwd: null input");
  weights_bwd_kernel<<<RAY_BLOCKS(a->n_rays), 256, 0, as_stream(stream)>>>(*a);
  EO_LAUNCH_CHECK();
  return EONERF_OK;
}

extern "C" int eonerf_accumulate_fwd(const EonerfAccumFwdArgs* a, eonerf_stream_t stream) {
  EO_REQUIRE(a && a->ray_offsets && a->n_rays >= 0 && a->n_channels >= 1, "accumulate_fwd: bad arguments");
  EO_REQUIRE(a->values || a->n_channels == 1, "accumulate_fwd: values==NULL needs n_channels==1");
  if (a->n_rays == 0) return EONERF_OK;
  EO_REQUIRE(a->out && (a->n_pts == 0 || a->weights), "accumulate_fwd: null pointer");
  if (render_vec_level() >= 1 && aligned16(a->weights) && aligned16(a->values) && (a->n_channels == 3 || a->n_channels == 1))
    accumulate_fwd_vec_kernel<<<ray_blocks_capped(a->n_rays), 256, 0, as_stream(stream)>>>(*a);
  else
    accumulate_fwd_kernel<<<RAY_BLOCKS(a->n_rays), 256, 0, as_stream(stream)>>>(*a);
  EO_LAUNCH_CHECK();
  return EONERF_OK;
}

extern "C" int eonerf_accumulate_bwd(const EonerfAccumBwdArgs* a, eonerf_stream_t stream) {
  EO_REQUIRE(a && a->ray_offsets && a->n_rays >= 0 && a->n_channels >= 1, "accumulate_bwd: bad arguments");
  EO_REQUIRE(a->values || a->n_channels == 1, "accumulate_bwd: values==NULL needs n_channels==1");
  if (a->n_rays == 0 || a->n_pts == 0) return EONERF_OK;
  EO_REQUIRE(a->g_out && a->weights, "accumulate_bwd: null pointer");
  accumulate_bwd_kernel<<<RAY_BLOCKS(a->n_rays), 256, 0, as_stream(stream)>>>(*a);
  EO_LAUNCH_CHECK();
  return EONERF_OK;
}

extern "C" int eonerf_composite_fwd(const EonerfCompositeFwdArgs* a, eonerf_stream_t stream) {
  EO_REQUIRE(a && a->ray_offsets && a->comp && a->n_rays >= 0, "composite_fwd: bad arguments");
  if (a->n_rays == 0) return EONERF_OK;
  EO_REQUIRE(a->n_pts == 0 || (a->t_starts && a->t_ends && a->z_mid && a->sigma), "composite_fwd: null input");
  if (render_vec_enabled() && aligned16(a->t_starts) && aligned16(a->t_ends) && aligned16(a->z_mid) && aligned16(a->sigma) && aligned16(a->albedo) &&
      aligned16(a->transient_s) && aligned16(a->transient_beta))
    composite_fwd_vec_kernel<<<ray_blocks_capped(a->n_rays), 256, 0, as_stream(stream)>>>(*a);
  else
    composite_fwd_kernel<<<ray_blocks_capped(a->n_rays), 256, 0, as_stream(stream)>>>(*a);
  EO_LAUNCH_CHECK();
  return EONERF_OK;
}

extern "C" int eonerf_composite_bwd(const EonerfCompositeBwdArgs* a, eonerf_stream_t stream) {
  EO_REQUIRE(a && a->ray_offsets && a->g_comp && a->n_rays >= 0, "composite_bwd: bad arguments");
  if (a->n_rays == 0) return EONERF_OK;
  EO_REQUIRE(a->n_pts == 0 || (a->t_starts && a->t_ends && a->z_mid && a->sigma && a->g_sigma),
             "composite_bwd: null pointer");
  if (render_vec_enabled() && aligned16(a->t_starts) && aligned16(a->t_ends) && aligned16(a->z_mid) && aligned16(a->sigma) && aligned16(a->albedo) &&
      aligned16(a->transient_s) && aligned16(a->transient_beta) && aligned16(a->g_sigma) && aligned16(a->g_albedo) && aligned16(a->g_transient_s) &&
      aligned16(a->g_transient_beta))
    composite_bwd_vec_kernel<<<ray_blocks_capped(a->n_rays), 256, 0, as_stream(stream)>>>(*a);
  else
    composite_bwd_kernel<<<ray_blocks_capped(a->n_rays), 256, 0, as_stream(stream)>>>(*a);
  EO_LAUNCH_CHECK();
  return EONERF_OK;
}

extern "C" int eonerf_sun_rays(const EonerfSunRaysArgs* a, eonerf_stream_t stream) {
  EO_REQUIRE(a && a->n_rays >= 0, "sun_rays: bad arguments");
  if (a->n_rays == 0) return EONERF_OK;
  EO_REQUIRE(a->origins && a->viewdirs && a->sundirs && a->depth && a->sun_rays, "sun_rays: null pointer");
  sun_rays_kernel<<<div_up(a->n_rays, 256), 256, 0, as_stream(stream)>>>(*a);
  EO_LAUNCH_CHECK();
  return EONERF_OK;
}

extern "C" int eonerf_shadow_fwd(const EonerfShadowFwdArgs* a, eonerf_stream_t stream) {
  EO_REQUIRE(a && a->ray_offsets && a->geo_shadow && a->n_rays >= 0, "shadow_fwd: bad arguments");
  if (a->n_rays == 0) return EONERF_OK;
  EO_REQUIRE(a->n_pts == 0 || (a->t_starts && a->t_ends && a->sigma), "shadow_fwd: null input");
  shadow_fwd_kernel<<<ray_blocks_capped(a->n_rays), 256, 0, as_stream(stream)>>>(*a);
  EO_LAUNCH_CHECK();
  return EONERF_OK;
}

extern "C" int eonerf_shadow_bwd(const EonerfShadowBwdArgs* a, eonerf_stream_t stream) {
  EO_REQUIRE(a && a->ray_offsets && a->n_rays >= 0, "shadow_bwd: bad arguments");
  if (a->n_rays == 0 || a->n_pts == 0) return EONERF_OK;
  EO_REQUIRE(a->t_starts && a->t_ends && a->geo_shadow && a->g_geo_shadow && a->g_sigma, "shadow_bwd: null pointer");
  shadow_bwd_kernel<<<ray_blocks_capped(a->n_rays), 256, 0, as_stream(stream)>>>(*a);
  EO_LAUNCH_CHECK();
  return EONERF_OK;
}

extern "C" int eonerf_sun_origin_bwd(const EonerfSunOriginBwdArgs* a, eonerf_stream_t stream) {
  EO_REQUIRE(a && a->ray_offsets && a->n_rays >= 0, "sun_origin_bwd: bad arguments");
  if (a->n_rays == 0) return EONERF_OK;
  EO_REQUIRE(a->viewdirs && a->g_depth && (a->n_pts == 0 || a->g_x), "sun_origin_bwd: null pointer");
  sun_origin_bwd_kernel<<<RAY_BLOCKS(a->n_rays), 256, 0, as_stream(stream)>>>(*a);
  EO_LAUNCH_CHECK();
  return EONERF_OK;
}

extern "C" int eonerf_epilogue_fwd(const EonerfEpilogueFwdArgs* a, eonerf_stream_t stream) {
  EO_REQUIRE(a && a->n_rays >= 0, "epilogue_fwd: bad arguments");
  if (a->n_rays == 0) return EONERF_OK;
  EO_REQUIRE(a->comp && a->pts_per_ray && a->img_idx && a->out, "epilogue_fwd: null pointer");
  epilogue_fwd_kernel<<<div_up(a->n_rays, 256), 256, 0, as_stream(stream)>>>(*a);
  EO_LAUNCH_CHECK();
  return EONERF_OK;
}

extern "C" int eonerf_epilogue_bwd(const EonerfEpilogueBwdArgs* a, eonerf_stream_t stream) {
  EO_REQUIRE(a && a->n_rays >= 0, "epilogue_bwd: bad arguments");
  if (a->n_rays == 0) return EONERF_OK;
  EO_REQUIRE(a->comp && a->img_idx && a->g_out && a->g_comp, "epilogue_bwd: null pointer");
  epilogue_bwd_kernel<<<div_up(a->n_rays, 256), 256, 0, as_stream(stream)>>>(*a);
  EO_LAUNCH_CHECK();
  return EONERF_OK;
}
