// Fused input-gradient chain of the radiance-field MLP for sm_100a (autograd backward of
// /root/reference/radiance_fields/eonerf.py:141-170 / mlp.py:87-101 with respect to the activations and positions).
//
// Same machine as the fused forward (field_fused.cu): one weight-producer warp, one MMA-issuer warp, eight epilogue
// warps, two 128-sample tiles ping-ponging per CTA, accumulators in TMEM, the running gradient G (wrt a layer's
// pre-activation) kept as the bf16 A operand in shared memory.  Per stage   G_prev = (G . W) * relu'(h_prev)   with the
// ReLU derivative taken from the forward's sign bits (32 B per sample per layer instead of re-reading activations).
// Every G is also stored tile-blocked to HBM: the parameter gradients dW = G^T X are separate tcgen05 GEMMs over the
// blocked G and the blocked forward stash (gemm_tn_blocked), because all layers' dW accumulators cannot live in TMEM.
//
// stage   A (G of)      B = W^T of            epilogue                                         output
//  0      T3            transient_mlp.3       mask T2                                          G_T2  -> blocks 2,3
//  1      T2            transient_mlp.2       mask T1                                          G_T1  -> blocks 0,1
//  2      T1            transient_mlp.1       mask HD0[:,128:]; + albedo head part (SIMT)      G_HD0 -> blocks 0..3
//  3      HD0           [albedo_mlp.0;t_mlp.0] none (bottleneck is linear)                     G_BOTT
//  4      BOTT          bottleneck            + dsigma (x) w_sigma, mask H7                    G_H7
//  5,6    H7, H6        base_mlp.7, .6        mask H6, H5                                      G_H6, G_H5
//  7      H5            base_mlp.5[:,256:319] none                                             G_ENC5 -> DENC block (only if g_x wanted)
//  8      H5            base_mlp.5[:,0:256]   mask H4                                          G_H4
//  9..12  H4..H1        base_mlp.4 .. .1      mask H3..H0                                      G_H3..G_H0
//  13     H0            base_mlp.0            + G_ENC5, positional-encoding backward           g_x (only if wanted)
#include <vector>
#include "fused_common.cuh"

// EONERF_STORE_HINT=1: bulk stash stores carry an L2 evict_first policy
#ifndef EONERF_STORE_HINT
#define EONERF_STORE_HINT 1
#endif
#if EONERF_STORE_HINT
#define EO_BULK_STORE(dst, src, bytes) bulk_store_hint(dst, src, bytes, l2_policy_evict_first())
#else
#define EO_BULK_STORE(dst, src, bytes) bulk_store(dst, src, bytes)
#endif

namespace eonerf {

namespace {

struct BStage {
  int8_t halves, nkb, a[4], out_blk, mask, mask_w0, kind, garr, pad;
};
// kind: 0 plain, 1 + albedo head part, 2 + sigma rank-1 term, 3 encoding part of layer 5, 4 final (positions)
__constant__ BStage c_bstage[kBwdStages] = {
    {1, 2, {0, 1, 0, 0}, 2, kMaskT1 + 1, 0, 0, 11, 0},
    {1, 2, {2, 3, 0, 0}, 0, kMaskT1 + 0, 0, 0, 10, 0},
    {1, 2, {0, 1, 0, 0}, 2, kMaskHd0, 4, 1, 9, 0},
    {2, 4, {0, 1, 2, 3}, 0, -1, 0, 0, 8, 0},
    {2, 4, {0, 1, 2, 3}, 0, kMaskH0 + 7, 0, 2, 7, 0},
    {2, 4, {0, 1, 2, 3}, 0, kMaskH0 + 6, 0, 0, 6, 0},
    {2, 4, {0, 1, 2, 3}, 0, kMaskH0 + 5, 0, 0, 5, 0},
    {1, 4, {0, 1, 2, 3}, 4, -1, 0, 3, -1, 0},
    {2, 4, {0, 1, 2, 3}, 0, kMaskH0 + 4, 0, 0, 4, 0},
    {2, 4, {0, 1, 2, 3}, 0, kMaskH0 + 3, 0, 0, 3, 0},
    {2, 4, {0, 1, 2, 3}, 0, kMaskH0 + 2, 0, 0, 2, 0},
    {2, 4, {0, 1, 2, 3}, 0, kMaskH0 + 1, 0, 0, 1, 0},
    {2, 4, {0, 1, 2, 3}, 0, kMaskH0 + 0, 0, 0, 0, 0},
    {1, 4, {0, 1, 2, 3}, 0, -1, 0, 4, -1, 0},
};
static const int kBwdBlkCount[kBwdStages] = {2, 2, 2, 8, 8, 8, 8, 4, 8, 8, 8, 8, 8, 4};

struct FusedBwdParams {
  int64_t M; int64_t n_tiles;
  const int64_t* M_dev;          // live sample count on the device (M, n_tiles are then capacities)
  int n_prog; int8_t prog[kBwdStages];
  int density_only;
  int vanilla;         // vanilla field: sigma = relu(.) (mlp.py:243,250), no transient branch (the chain starts at the HD0 stage)
  const uint8_t* wblob; const float* consts;
  MmaProgram mma;      // MMA side of prog[]
  const uint32_t* mask[kNumMask];
  uint8_t* garr[13];
  const float* xf;
  // forward outputs and incoming gradients (any g_* may be NULL = 0)
  const float* sigma; const float* rgb; const float* ts; const float* tb;
  const float* g_sigma; const float* g_rgb; const float* g_ts; const float* g_tb;
  float* dpre;         // [Mpad, 8]: 0 sigma, 1..3 albedo, 4 transient_s, 5 transient_beta (pre-activation gradients of the heads)
  float* g_x;          // [M, 3] or NULL
};

// shared-memory use of the constants region in this kernel: head weights (kCWSigma.. as in the forward), then per-row
// head gradients [2 slots][128 rows][4] = (dsigma, drgb0, drgb1, drgb2)
constexpr int kOffRows = kOffConst + 4096;
static_assert(4096 + 2 * 128 * 16 <= kConstBytes, "row scalars do not fit");

// keep[j] of a packed bf16 pair: sign-bit layout written by the forward (column 2j -> bit 15-j, column 2j+1 -> bit 31-j)
__device__ __forceinline__ uint32_t apply_mask(uint32_t pk, uint32_t m, int j) {
  // bits 15 and 31 of (m << j) are the keep flags of the pair: PRMT in sign-replicate mode (selector nibble | 8) spreads the
  // top bit of bytes 1 and 3 over the low and the high half -> 0xFFFF per kept half, one instruction
  uint32_t sel;
  asm("prmt.b32 %0, %1, %2, 0xBB99;" : "=r"(sel) : "r"(m << j), "r"(0u));
  return pk & sel;
}

// kRing = 3: slots of 5 blocks (DENC for the position gradients); kRing = 5: no DENC block, a 5-deep weight ring (want_x == false)
template <int kCG, int kMC, int kRing>
__global__ void __launch_bounds__(kFusedThreads, 1) fused_bwd_kernel(const __grid_constant__ FusedBwdParams p, const __grid_constant__ CUtensorMap wmap) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  float* hw = (float*)(smem + kOffConst);                   // head weights: w_sigma 256 | w_alb 384 | w_ts 128 | w_tb 128
  constexpr int kCl = kCG * kMC;                              // CTAs per cluster
  const uint32_t rank = kCl > 1 ? cluster_ctarank() : 0;
  FusedBars B;
  uint32_t* tmem_base_s;
  constexpr int kOffSlotL = off_slot(kRing), kSlotBytesL = slot_bytes(kRing);
  EO_CTA_TIME(0);
  fused_setup<kCG, kMC, kRing>(smem, B, tmem_base_s, rank);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 896; i += kFusedThreads) hw[i] = __ldg(p.consts + kCWSigma + i);
  tc_fence_before();
  if (kCl > 1) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_s;
  const int64_t M = p.M_dev ? __ldg(p.M_dev) : p.M;
  const int64_t n_tiles = p.M_dev ? (M + kTileM - 1) / kTileM : p.n_tiles;
  const int64_t n_items = (n_tiles + 2 * kCl - 1) / (2 * kCl);
  const int64_t it0 = blockIdx.x / kCl, it_stride = gridDim.x / kCl;

  if (warp == 0) {
    if (lane == 0) fused_producer<kCG, kMC, kRing>(p.mma, p.wblob, &wmap, smem, B, it0, n_items, it_stride, rank);
  } else if (warp == 1) {
    if (kCG == 1 || rank == 0) fused_mma_issuer<kCG, kMC, false, kRing>(p.mma, smem, B, tmem_base, it0, n_items, it_stride);       // whole warp, converged
  } else if (kDutyWarp && warp == kDutyWarpId) {
    // ===== duty warp (see field_fused.cu): mirrors the epilogue warps' named barriers, signals the MMA issuer and owns the
    // bulk stores of the G arrays =====
    for (int64_t it = it0; it < n_items; it += it_stride) {
      if (kDutyStores && lane == 0) tma_store_wait_read<0>();
      named_bar_sync(1, kBarThreads);                       // start of the item
      named_bar_sync(1, kBarThreads);                       // head gradients -> first G written
      if (lane == 0) {
        signal_act_ready<kCG>(B, 0, rank);
        signal_act_ready<kCG>(B, 1, rank);
        const int ga = p.density_only ? 7 : 12;
        const int nb = kDutyStores ? (p.density_only ? 4 : 2) : 0;
        for (int slot = 0; slot < 2; ++slot) {
          const int64_t tile = 2 * kCl * it + 2 * rank + slot;
          if (tile < n_tiles)
            for (int bb = 0; bb < nb; ++bb)
              EO_BULK_STORE(p.garr[ga] + ((size_t)tile * nb + bb) * kBlkBytes, smem + kOffSlotL + slot * kSlotBytesL + bb * kBlkBytes, kBlkBytes);
        }
        tma_store_commit();
      }
      __syncwarp();
      for (int i = 0; i < p.n_prog; ++i) {
        const BStage d = c_bstage[p.prog[i]];
        for (int slot = 0; slot < 2; ++slot) {
          if (d.kind == 4) {                                 // final stage: positions only, nothing to hand over or store
            named_bar_sync(1, kBarThreads);
            continue;
          }
          if (kDutyStores && lane == 0) tma_store_wait_read<0>();
          named_bar_sync(1, kBarThreads);
          if (lane == 0) {
            if (i + 1 < p.n_prog) signal_act_ready<kCG>(B, slot, rank);
            const int64_t tile = 2 * kCl * it + 2 * rank + slot;
            if (kDutyStores && d.garr >= 0 && tile < n_tiles) {
              const int nb = d.kind == 1 ? 4 : d.halves * 2;
              const int b0 = d.kind == 1 ? 0 : d.out_blk;
              for (int bb = 0; bb < nb; ++bb)
                EO_BULK_STORE(p.garr[d.garr] + ((size_t)tile * nb + bb) * kBlkBytes, smem + kOffSlotL + slot * kSlotBytesL + (b0 + bb) * kBlkBytes, kBlkBytes);
            }
            tma_store_commit();
          }
          __syncwarp();
        }
      }
    }
    if (kDutyStores && lane == 0) tma_store_wait_all();
  } else {
    // ===== epilogue warps =====
    const int e = threadIdx.x - 64;
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int r = q * 32 + lane;
    const int store_id = kDutyStores ? -1 : store_thread_id(e);
    const uint32_t s_hw = smem_u32(hw);
    uint32_t cph = 0;
    uint64_t* const acc_full = B.acc_full;
    int tr_i = 0; (void)tr_i;
    for (int64_t it = it0; it < n_items; it += it_stride) {
      EO_TRACE(1, tr_i, threadIdx.x == 64);                 // item start
      // ---- head gradients -> first G of the chain ----
      if (store_id >= 0) tma_store_wait_read<0>();
      named_bar_sync(1, kBarThreads);
      // Every global load of the prologue (both slots: forward outputs, incoming gradients, ReLU sign words) is issued before
      // the first dependent instruction: one memory round trip instead of ~8 per slot (the measured per-item prologue was
      // 14 k cycles of a 100 k-cycle item, profiles/r2b_bwd_item_trace_before.log).
      float in_gs[2], in_s[2], in_gr[2][3], in_r[2][3], in_gts[2], in_ts[2], in_gtb[2], in_tb[2];
      uint4 in_mw[2];
#pragma unroll
      for (int slot = 0; slot < 2; ++slot) {
        const int64_t pt = (2 * kCl * it + 2 * rank + slot) * kTileM + r;
        const bool valid = pt < M;
        in_gs[slot] = (valid && p.g_sigma) ? __ldg(p.g_sigma + pt) : 0.f;
        in_s[slot] = valid ? __ldg(p.sigma + pt) : 0.f;
        const bool full = valid && !p.density_only;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          in_gr[slot][c] = (full && p.g_rgb) ? __ldg(p.g_rgb + 3 * pt + c) : 0.f;
          in_r[slot][c] = (full && p.g_rgb) ? __ldg(p.rgb + 3 * pt + c) : 0.f;
        }
        in_gts[slot] = (full && p.g_ts) ? __ldg(p.g_ts + pt) : 0.f;
        in_ts[slot] = (full && p.g_ts) ? __ldg(p.ts + pt) : 0.f;
        in_gtb[slot] = (full && p.g_tb) ? __ldg(p.g_tb + pt) : 0.f;
        in_tb[slot] = (full && p.g_tb) ? __ldg(p.tb + pt) : 0.f;
        in_mw[slot] = make_uint4(0, 0, 0, 0);
        if (valid) {
          if (p.density_only) in_mw[slot] = __ldg((const uint4*)(p.mask[kMaskH0 + 7] + pt * 8 + half * 4));
          else { const uint2 t2 = __ldg((const uint2*)(p.mask[kMaskT1 + 2] + pt * 8 + half * 2)); in_mw[slot].x = t2.x; in_mw[slot].y = t2.y; }
        }
      }
#pragma unroll
      for (int slot = 0; slot < 2; ++slot) {
        const int64_t tile = 2 * kCl * it + 2 * rank + slot;
        const int64_t pt = tile * kTileM + r;
        const bool valid = pt < M;
        const uint32_t act = smem_u32(smem + kOffSlotL + slot * kSlotBytesL);
        const uint32_t s_row = smem_u32(smem + kOffRows) + (uint32_t)(slot * 128 + r) * 16u;
        // derivatives through the forward outputs: softplus' = 1 - exp(-y) (vanilla: relu'), sigmoid' = y (1 - y)
        const float dsig = in_gs[slot] * (p.vanilla ? (in_s[slot] > 0.f ? 1.0f : 0.f) : -expm1f(-in_s[slot]));
        const float d0 = in_gr[slot][0] * in_r[slot][0] * (1.0f - in_r[slot][0]);
        const float d1 = in_gr[slot][1] * in_r[slot][1] * (1.0f - in_r[slot][1]);
        const float d2 = in_gr[slot][2] * in_r[slot][2] * (1.0f - in_r[slot][2]);
        const float dts = in_gts[slot] * in_ts[slot] * (1.0f - in_ts[slot]);
        const float dtb = in_gtb[slot] * (-expm1f(-in_tb[slot]));
        if (half == 0) {
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(s_row), "f"(dsig), "f"(d0), "f"(d1), "f"(d2) : "memory");
          if (tile < n_tiles) {                            // rows of the padded tail are zero: they add nothing to dW
            float* dp = p.dpre + pt * 8;
            *(float4*)dp = make_float4(dsig, d0, d1, d2);
            *(float4*)(dp + 4) = make_float4(dts, dtb, 0.f, 0.f);
          }
        }
        if (p.density_only) {
          // G_H7 = dsigma (x) w_sigma, masked by H7 > 0: thread covers columns half*128 .. +128
          const uint4 mw = in_mw[slot];
#pragma unroll 1
          for (int c = 0; c < 4; ++c) {
            const uint32_t m = c == 0 ? mw.x : (c == 1 ? mw.y : (c == 2 ? mw.z : mw.w));
            const uint32_t sw = s_hw + (uint32_t)(half * 128 + c * 32) * 4u;
            uint32_t pk[16];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float w0, w1, w2, w3;
              lds_f4(sw + j * 16, w0, w1, w2, w3);
              pk[2 * j] = apply_mask(pack_bf16(dsig * w0, dsig * w1), m, 2 * j);
              pk[2 * j + 1] = apply_mask(pack_bf16(dsig * w2, dsig * w3), m, 2 * j + 1);
            }
            const int colg = half * 128 + c * 32;
            const uint32_t blk = act + (uint32_t)(colg >> 6) * kBlkBytes;
            const int ch0 = (colg & 63) >> 3;
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) sts_u4(blk + blk_off(r, ch0 + jj), pk[4 * jj], pk[4 * jj + 1], pk[4 * jj + 2], pk[4 * jj + 3]);
          }
        } else {
          // G_T3 = (dts (x) w_ts + dtb (x) w_tb), masked by T3 > 0: thread covers columns half*64 .. +64 -> blocks 0,1
          const uint2 mw = make_uint2(in_mw[slot].x, in_mw[slot].y);
#pragma unroll 1
          for (int c = 0; c < 2; ++c) {
            const uint32_t m = c == 0 ? mw.x : mw.y;
            const uint32_t sws = s_hw + (uint32_t)(640 + half * 64 + c * 32) * 4u;     // w_ts at float 640, w_tb at 768
            uint32_t pk[16];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float a0, a1, a2, a3, b0, b1, b2, b3;
              lds_f4(sws + j * 16, a0, a1, a2, a3);
              lds_f4(sws + 512 + j * 16, b0, b1, b2, b3);
              pk[2 * j] = apply_mask(pack_bf16(fmaf(dts, a0, dtb * b0), fmaf(dts, a1, dtb * b1)), m, 2 * j);
              pk[2 * j + 1] = apply_mask(pack_bf16(fmaf(dts, a2, dtb * b2), fmaf(dts, a3, dtb * b3)), m, 2 * j + 1);
            }
            const int colg = half * 64 + c * 32;
            const uint32_t blk = act + (uint32_t)(colg >> 6) * kBlkBytes;
            const int ch0 = (colg & 63) >> 3;
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) sts_u4(blk + blk_off(r, ch0 + jj), pk[4 * jj], pk[4 * jj + 1], pk[4 * jj + 2], pk[4 * jj + 3]);
          }
        }
      }
      fence_proxy_async();
      named_bar_sync(1, kBarThreads);
      EO_TRACE(1, tr_i, threadIdx.x == 64);                 // prologue done
      if (!kDutyWarp && e == kSignalThread) {
        signal_act_ready<kCG>(B, 0, rank);
        signal_act_ready<kCG>(B, 1, rank);
      }
      if (store_id >= 0) {                                   // block `store_id` of both slots' first G
        const int ga = p.density_only ? 7 : 12;
        const int nb = p.density_only ? 4 : 2;
        for (int bb = store_id; bb < nb; bb += kStoreThreads)
          for (int slot = 0; slot < 2; ++slot) {
            const int64_t tile = 2 * kCl * it + 2 * rank + slot;
            if (tile < n_tiles)
              EO_BULK_STORE(p.garr[ga] + ((size_t)tile * nb + bb) * kBlkBytes, smem + kOffSlotL + slot * kSlotBytesL + bb * kBlkBytes, kBlkBytes);
          }
        tma_store_commit();
      }

      // ---- the chain ----
      for (int i = 0; i < p.n_prog; ++i) {
        const int s = p.prog[i];
        const BStage d = c_bstage[s];
        const int cpt = d.halves == 2 ? 128 : 64;
        const int col0 = half * cpt;
        for (int slot = 0; slot < 2; ++slot) {
          const int64_t tile = 2 * kCl * it + 2 * rank + slot;
          const int64_t pt = tile * kTileM + r;
          const bool valid = pt < M;
          const uint32_t act = smem_u32(smem + kOffSlotL + slot * kSlotBytesL);
          const uint32_t s_row = smem_u32(smem + kOffRows) + (uint32_t)(slot * 128 + r) * 16u;
          // ReLU sign bits of this thread's columns (and of the albedo half for stage 2), fetched while the MMA runs
          uint4 mw = make_uint4(~0u, ~0u, ~0u, ~0u);
          uint2 mwa = make_uint2(0u, 0u);
          if (d.mask >= 0) {
            mw = make_uint4(0u, 0u, 0u, 0u);
            if (valid) {
              const uint32_t* mrow = p.mask[d.mask] + pt * 8;
              if (cpt == 128) mw = __ldg((const uint4*)(mrow + half * 4));
              else { const uint2 t = __ldg((const uint2*)(mrow + d.mask_w0 + half * 2)); mw.x = t.x; mw.y = t.y; }
              if (d.kind == 1) mwa = __ldg((const uint2*)(mrow + half * 2));
            }
          }
          float dsig = 0.f, d0 = 0.f, d1 = 0.f, d2 = 0.f;
          if (d.kind == 1 || d.kind == 2) asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(dsig), "=f"(d0), "=f"(d1), "=f"(d2) : "r"(s_row) : "memory");
          mbar_wait(&acc_full[slot], (cph >> slot) & 1u);
          cph ^= 1u << slot;
          tc_fence_after();
          const uint32_t taddr = tmem_base + slot * 256 + ((uint32_t)(q * 32) << 16) + col0;
          if (d.kind == 4) {
            // ---- final stage: g_enc = acc[0:64] + G_ENC5, positional-encoding backward (half 0 owns the 64 columns) ----
            if (half == 0) {
              float ge[64];
#pragma unroll
              for (int c = 0; c < 2; ++c) {
                uint32_t v[32];
                tmem_ld32(taddr + c * 32, v);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) ge[c * 32 + j] = __uint_as_float(v[j]);
              }
              const uint32_t denc = act + 4 * kBlkBytes;
#pragma unroll
              for (int ch = 0; ch < 8; ++ch) {
                uint32_t w0, w1, w2, w3;
                asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3) : "r"(denc + blk_off(r, ch)) : "memory");
                ge[ch * 8 + 0] += bf_lo(w0); ge[ch * 8 + 1] += bf_hi(w0); ge[ch * 8 + 2] += bf_lo(w1); ge[ch * 8 + 3] += bf_hi(w1);
                ge[ch * 8 + 4] += bf_lo(w2); ge[ch * 8 + 5] += bf_hi(w2); ge[ch * 8 + 6] += bf_lo(w3); ge[ch * 8 + 7] += bf_hi(w3);
              }
              if (valid) {
                // g_x[c] = g_enc[c] + sum_k 2^k ( cos(2^k x_c) g_enc[3+3k+c] + cos(2^k x_c + pi/2) g_enc[33+3k+c] )
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                  const float x = __ldg(p.xf + 3 * pt + c);
                  float g = ge[c];
#pragma unroll
                  for (int k = 0; k < 10; ++k) {
                    const float sc = (float)(1 << k), xb = x * sc;
                    g += sc * (cos_reduced(xb) * ge[3 + 3 * k + c] + cos_reduced(__fadd_rn(xb, kHalfPi)) * ge[33 + 3 * k + c]);
                  }
                  p.g_x[3 * pt + c] = g;
                }
              }
            }
            tc_fence_before();
            named_bar_sync(1, kBarThreads);
            continue;
          }
          const bool write_out = !(d.kind == 3 && half == 1);   // encoding part: only 64 real columns
          // one 32-column chunk: accumulator (+ sigma rank-1 term) -> bf16 -> ReLU mask -> next A operand in shared memory
          auto chunk = [&](uint32_t (&v)[32], const int c) {
            const uint32_t m = c == 0 ? mw.x : (c == 1 ? mw.y : (c == 2 ? mw.z : mw.w));
            uint32_t pk[16];
            if (d.kind == 2) {
              const uint32_t sw = s_hw + (uint32_t)(col0 + c * 32) * 4u;
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                float w0, w1, w2, w3;
                lds_f4(sw + j * 16, w0, w1, w2, w3);
                pk[2 * j] = apply_mask(pack_bf16(fmaf(dsig, w0, __uint_as_float(v[4 * j])), fmaf(dsig, w1, __uint_as_float(v[4 * j + 1]))), m, 2 * j);
                pk[2 * j + 1] = apply_mask(pack_bf16(fmaf(dsig, w2, __uint_as_float(v[4 * j + 2])), fmaf(dsig, w3, __uint_as_float(v[4 * j + 3]))), m, 2 * j + 1);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) pk[j] = apply_mask(pack_bf16(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])), m, j);
            }
            if (write_out) {
              const int colg = col0 + c * 32;
              const uint32_t blk = act + (uint32_t)(d.out_blk + (colg >> 6)) * kBlkBytes;
              const int ch0 = (colg & 63) >> 3;
#pragma unroll
              for (int jj = 0; jj < 4; ++jj) sts_u4(blk + blk_off(r, ch0 + jj), pk[4 * jj], pk[4 * jj + 1], pk[4 * jj + 2], pk[4 * jj + 3]);
            }
          };
          {   // software pipeline: the tcgen05.ld of chunk c+1 is in flight while chunk c is processed
            uint32_t va[32], vb[32];
            tmem_ld32(taddr, va);
            tmem_ld_wait_dep(va);
            tmem_ld32(taddr + 32, vb);
            chunk(va, 0);
            tmem_ld_wait_dep(vb);
            if (cpt == 128) tmem_ld32(taddr + 64, va);
            chunk(vb, 1);
            if (cpt == 128) {
              tmem_ld_wait_dep(va);
              tmem_ld32(taddr + 96, vb);
              chunk(va, 2);
              tmem_ld_wait_dep(vb);
              chunk(vb, 3);
            }
          }
          if (d.kind == 1) {
            // albedo half of G_HD0: sum_j drgb_j (x) albedo_mlp.output_layer.weight[j], masked by HD0[:, 0:128] > 0 -> blocks 0,1
#pragma unroll 1
            for (int c = 0; c < 2; ++c) {
              const uint32_t m = c == 0 ? mwa.x : mwa.y;
              const uint32_t sw = s_hw + (uint32_t)(256 + half * 64 + c * 32) * 4u;    // w_alb rows at floats 256, 384, 512
              uint32_t pk[16];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                float a0, a1, a2, a3, b0, b1, b2, b3, c0, c1, c2, c3;
                lds_f4(sw + j * 16, a0, a1, a2, a3);
                lds_f4(sw + 512 + j * 16, b0, b1, b2, b3);
                lds_f4(sw + 1024 + j * 16, c0, c1, c2, c3);
                pk[2 * j] = apply_mask(pack_bf16(fmaf(d0, a0, fmaf(d1, b0, d2 * c0)), fmaf(d0, a1, fmaf(d1, b1, d2 * c1))), m, 2 * j);
                pk[2 * j + 1] = apply_mask(pack_bf16(fmaf(d0, a2, fmaf(d1, b2, d2 * c2)), fmaf(d0, a3, fmaf(d1, b3, d2 * c3))), m, 2 * j + 1);
              }
              const int colg = half * 64 + c * 32;
              const uint32_t blk = act + (uint32_t)(colg >> 6) * kBlkBytes;
              const int ch0 = (colg & 63) >> 3;
#pragma unroll
              for (int jj = 0; jj < 4; ++jj) sts_u4(blk + blk_off(r, ch0 + jj), pk[4 * jj], pk[4 * jj + 1], pk[4 * jj + 2], pk[4 * jj + 3]);
            }
          }
          tc_fence_before();
          fence_proxy_async();
          // the OTHER slot's latest G store (issued one epilogue ago) must have finished reading shared memory before the
          // next epilogue overwrites that slot: checked here, so one named barrier per stage is enough
          if (store_id >= 0) tma_store_wait_read<0>();
          named_bar_sync(1, kBarThreads);
          if (!kDutyWarp && e == kSignalThread && i + 1 < p.n_prog) signal_act_ready<kCG>(B, slot, rank);
          if (store_id >= 0 && d.garr >= 0 && tile < n_tiles) {
            const int nb = d.kind == 1 ? 4 : d.halves * 2;
            const int b0 = d.kind == 1 ? 0 : d.out_blk;
            for (int bb = store_id; bb < nb; bb += kStoreThreads)    // 16 KB blocks store_id, store_id + kStoreThreads, ... of this G
              EO_BULK_STORE(p.garr[d.garr] + ((size_t)tile * nb + bb) * kBlkBytes, smem + kOffSlotL + slot * kSlotBytesL + (b0 + bb) * kBlkBytes, kBlkBytes);
            tma_store_commit();
          }
        }
      }
    }
    if (store_id >= 0) tma_store_wait_all();
  }
  fused_teardown<kCG, kMC>(tmem_base);
  EO_CTA_TIME(1);
}

// ---- SIMT side kernels over the blocked layout ------------------------------------------------------------------------
// element (sample m, feature k) of a blocked array with nb blocks per tile
__device__ __forceinline__ const uint4* blk_chunk(const uint8_t* base, int nb, int64_t m, int chunk) {
  const int64_t tile = m >> 7;
  const int rr = (int)(m & 127);
  return (const uint4*)(base + ((size_t)tile * nb + (chunk >> 3)) * kBlkBytes + rr * 128 + (((chunk & 7) ^ (rr & 7)) << 4));
}
__device__ __forceinline__ void unpack8(const uint4 v, float (&f)[8]) {
  f[0] = bf_lo(v.x); f[1] = bf_hi(v.x); f[2] = bf_lo(v.y); f[3] = bf_hi(v.y);
  f[4] = bf_lo(v.z); f[5] = bf_hi(v.z); f[6] = bf_lo(v.w); f[7] = bf_hi(v.w);
}

// dw_j[k] += sum_m dpre[m, col0+j] X[m, k];  db_j += sum_m dpre[m, col0+j]     (narrow heads, K = 128 or 256 features
// starting at 16-byte chunk `chunk0` of the blocked array).  256 threads = (256/lpr) rows x lpr chunks per sweep.
struct HeadGradsB { float* dw[3]; float* db[3]; };
#ifndef EONERF_HEADS_UNROLL
#define EONERF_HEADS_UNROLL 4
#endif
constexpr int kHeadsUnroll = EONERF_HEADS_UNROLL;      // 16-byte loads in flight per thread
template <int J>
__global__ void __launch_bounds__(256) heads_dw_blocked_kernel(const uint8_t* __restrict__ X, int nb, int chunk0, int K, int64_t M,
                                                               const float* __restrict__ dpre, int col0, HeadGradsB G,
                                                               int64_t rows_per_block, const int64_t* __restrict__ M_dev) {
  __shared__ float red[8][J][264];
  if (M_dev) {                     // live count on the device: re-balance the rows over the grid (multiples of 64 rows)
    M = __ldg(M_dev);
    rows_per_block = ((M + gridDim.x - 1) / gridDim.x + 63) & ~(int64_t)63;
  }
  if ((int64_t)blockIdx.x * rows_per_block >= M) return;
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
  for (int64_t m0 = m_begin + rsub; m0 < m_end; m0 += kHeadsUnroll * rows) {
    uint4 xv[kHeadsUnroll];
    float dv[kHeadsUnroll][J];
#pragma unroll
    for (int u = 0; u < kHeadsUnroll; ++u) {
      const int64_t m = m0 + (int64_t)u * rows;
      if (m < m_end) {
        xv[u] = __ldg(blk_chunk(X, nb, m, chunk0 + sub));
#pragma unroll
        for (int j = 0; j < J; ++j) dv[u][j] = __ldg(dpre + m * 8 + col0 + j);
      } else {
        xv[u] = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int j = 0; j < J; ++j) dv[u][j] = 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < kHeadsUnroll; ++u) {
      float x[8];
      unpack8(xv[u], x);
#pragma unroll
      for (int j = 0; j < J; ++j) {
        bs[j] += dv[u][j];
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[j][e] = fmaf(dv[u][j], x[e], acc[j][e]);
      }
    }
  }
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

// dcb[img, j] += sum_{rows of image img} G_HD0[row, 128 + j]   (j < 128): gradient of the per-image bias rows
__global__ void __launch_bounds__(256) class_grad_blocked_kernel(const uint8_t* __restrict__ G, const int32_t* __restrict__ cls, int64_t M,
                                                                 int64_t n_images, float* __restrict__ dcb, int64_t rows_per_block,
                                                                 int use_smem, const int64_t* __restrict__ M_dev) {
  extern __shared__ float tab[];
  if (M_dev) {
    M = __ldg(M_dev);
    rows_per_block = ((M + gridDim.x - 1) / gridDim.x + 63) & ~(int64_t)63;
  }
  if ((int64_t)blockIdx.x * rows_per_block >= M) return;
  const int sub = threadIdx.x & 15, rsub = threadIdx.x >> 4;
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
  for (int64_t m0 = m_begin + rsub; m0 < m_end; m0 += 64) {   // four rows in flight per thread
    int cc[4];
    uint4 gv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t m = m0 + 16 * u;
      const bool ok = m < m_end;
      cc[u] = ok ? __ldg(cls + m) : -1;
      gv[u] = ok ? __ldg(blk_chunk(G, 4, m, 16 + sub)) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (cc[u] < 0) continue;
      if (cc[u] != cur) {
        if (cur >= 0) {
#pragma unroll
          for (int e = 0; e < 8; ++e) { atomicAdd(dst + (int64_t)cur * kHid + sub * 8 + e, acc[e]); acc[e] = 0.f; }
        }
        cur = cc[u];
      }
      float v[8];
      unpack8(gv[u], v);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] += v[e];
    }
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
__global__ void emb_grad_fused_kernel(const float* __restrict__ dcb, const float* __restrict__ wt0, const float* __restrict__ emb,
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

template <int J>
static int run_heads_dw_blocked(const uint8_t* X, int nb, int chunk0, int K, int64_t M, const float* dpre, int col0, const HeadGradsB& G,
                                cudaStream_t s, const int64_t* M_dev) {
  // a whole number of blocks per SM (4 x 148) when there is enough work, rows per block a multiple of 64
  int64_t blocks = 4 * 148;
  if (blocks > div_up(M, 1024)) blocks = div_up(M, 1024);
  const int64_t rows = (div_up(M, blocks) + 63) & ~(int64_t)63;
  heads_dw_blocked_kernel<J><<<(unsigned)div_up(M, rows), 256, 0, s>>>(X, nb, chunk0, K, M, dpre, col0, G, rows, M_dev);
  EO_LAUNCH_CHECK();
  return EONERF_OK;
}

}  // namespace

#define EO_TRY(expr)                \
  do {                              \
    int _r = (expr);                \
    if (_r != EONERF_OK) return _r; \
  } while (0)

// The narrow-head / per-image-bias gradient kernels only depend on the chain kernel and touch other gradient tensors than the
// grouped dW GEMM: they run on a side stream, forked after the chain and joined after the GEMM, so they overlap it (all of them
// are HBM readers that do not saturate the bus on their own).  Fork and join are event edges, which a stream capture turns into
// graph dependencies.  One side stream + two events per device, created on first use; EONERF_SIDE_STREAM=0 keeps one stream.
// Vanilla field: gradient of the view-direction columns of rgb_layer.hidden_layers.0,
//   dW[j, 256 + e] += sum_m G_HD0[m, j] enc4(dir_{cls(m)})[e]      (j < 128, e < 27; G in the blocked layout, first two blocks of four)
// Samples arrive ray by ray, so a 128-sample tile holds a handful of runs of equal class (per-ray directions: one or two; per-sample
// directions: 128).  Thread (j, half) sums G[:, j] over each run of its 64 rows and multiplies the run's sum with the run's encoding
// once: 64 adds + 27 FMAs per run instead of 64 x 27 FMAs.  Persistent CTAs: 27 partial columns per thread in registers, added once.
__global__ void __launch_bounds__(256) vanilla_dir_grad_kernel(const uint8_t* __restrict__ G, const int32_t* __restrict__ cls, int64_t M, int64_t n_tiles,
                                                               const float* __restrict__ dirs, int64_t stride, int64_t n_cond, float* __restrict__ dw,
                                                               const int64_t* __restrict__ M_dev) {
  __shared__ float enc[128][28];                       // rows that start a run only
  __shared__ int s_cls[128];                           // -1: past the end
  __shared__ __align__(16) uint8_t gt[2 * kBlkBytes];
  if (M_dev) { M = __ldg(M_dev); n_tiles = (M + kTileM - 1) / kTileM; }
  const int t = threadIdx.x, j = t & 127, m0 = (t >> 7) * 64;
  float acc[27];
#pragma unroll
  for (int e = 0; e < 27; ++e) acc[e] = 0.f;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    __syncthreads();
    // the tile's G blocks 0, 1 (columns 0..127) as they lie in memory (swizzled images), 32 KB
    const uint4* src = (const uint4*)(G + (size_t)tile * 4 * kBlkBytes);
    for (int i = t; i < 2 * kBlkBytes / 16; i += 256) ((uint4*)gt)[i] = __ldg(src + i);
    if (t < 128) {
      const int64_t pt = tile * kTileM + t;
      int c = -1;
      if (pt < M) {
        c = __ldg(cls + pt);
        if (c < 0 || c >= n_cond) c = 0;
      }
      s_cls[t] = c;
    }
    __syncthreads();
    // enc4 of the direction of every row that starts a run (rows 0 and 64 always do: the two thread halves work independently)
    for (int i = t; i < 128 * 27; i += 256) {
      const int m = i / 27, c = i - m * 27;
      const int k = s_cls[m];
      if (k < 0 || ((m & 63) != 0 && s_cls[m - 1] == k)) continue;
      const float* d = dirs + (int64_t)k * stride;
      float v;
      if (c < 3) v = __ldg(d + c);
      else {
        int e = c - 3;
        const int half = e >= 12;
        e -= half * 12;
        const float xb = __ldg(d + e % 3) * (float)(1 << (e / 3));
        v = sinf(half ? __fadd_rn(xb, kHalfPi) : xb);
      }
      enc[m][c] = v;
    }
    __syncthreads();
    // column j of sample m: block j >> 6, 16-byte chunk (j & 63) >> 3, element j & 7
    const int blk = j >> 6, ch = (j & 63) >> 3, el = j & 7;
    int cur = s_cls[m0], start = m0;
    float run = 0.f;
    for (int m = m0; m < m0 + 64; ++m) {
      const int k = s_cls[m];
      if (k != cur) {
        if (cur >= 0) {
#pragma unroll
          for (int e = 0; e < 27; ++e) acc[e] = fmaf(run, enc[start][e], acc[e]);
        }
        cur = k; start = m; run = 0.f;
      }
      const __nv_bfloat16* row = (const __nv_bfloat16*)(gt + blk * kBlkBytes + m * 128 + ((ch ^ (m & 7)) << 4));
      run += __bfloat162float(row[el]);
    }
    if (cur >= 0) {
#pragma unroll
      for (int e = 0; e < 27; ++e) acc[e] = fmaf(run, enc[start][e], acc[e]);
    }
  }
#pragma unroll
  for (int e = 0; e < 27; ++e) atomicAdd(dw + (int64_t)j * 283 + 256 + e, acc[e]);
}

struct SideStream { int dev = -1; cudaStream_t owner = nullptr; cudaStream_t stream = nullptr; cudaEvent_t fork = nullptr, join = nullptr; };
// One side stream + event pair per (host thread, device, caller stream): two host threads, or one thread driving two streams,
// never share (and re-record) the same events.  thread_local => no locking; entries live as long as the thread.
static SideStream* side_stream(cudaStream_t owner) {
  static const int enabled = [] { const char* e = getenv("EONERF_SIDE_STREAM"); return e ? atoi(e) : 1; }();
  if (!enabled) return nullptr;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  thread_local std::vector<SideStream> table;
  for (SideStream& S : table)
    if (S.dev == dev && S.owner == owner) return &S;
  if (table.size() >= 64) return nullptr;                   // a caller cycling through many streams: stay on one stream
  SideStream S;
  S.dev = dev; S.owner = owner;
  if (cudaStreamCreateWithFlags(&S.stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&S.fork, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&S.join, cudaEventDisableTiming) != cudaSuccess) {
    cudaGetLastError();
    if (S.stream) cudaStreamDestroy(S.stream);
    if (S.fork) cudaEventDestroy(S.fork);
    return nullptr;
  }
  table.reserve(64);                                        // pointers handed out stay valid
  table.push_back(S);
  return &table.back();
}

// Fork / join of the side stream with the join guaranteed on every exit path: an early error return between fork and join
// would otherwise leave the side stream's work unordered against whatever the caller launches next on `s` (and, under stream
// capture, leave the capture with an unjoined branch).
struct SideFork {
  SideStream* side; cudaStream_t s; bool forked = false;
  SideFork(SideStream* side_, cudaStream_t s_) : side(side_), s(s_) {}
  int fork() {
    if (!side) return EONERF_OK;
    EO_CUDA(cudaEventRecord(side->fork, s));
    EO_CUDA(cudaStreamWaitEvent(side->stream, side->fork, 0));
    forked = true;
    return EONERF_OK;
  }
  cudaStream_t stream() const { return forked ? side->stream : s; }
  int join() {
    if (!forked) return EONERF_OK;
    forked = false;
    EO_CUDA(cudaEventRecord(side->join, side->stream));
    EO_CUDA(cudaStreamWaitEvent(s, side->join, 0));
    return EONERF_OK;
  }
  ~SideFork() { if (forked) { cudaEventRecord(side->join, side->stream); cudaStreamWaitEvent(s, side->join, 0); } }
};

int fused_field_bwd(const EonerfFieldBwdArgs* a, cudaStream_t s) {
  const bool vanilla = a->field == EONERF_FIELD_VANILLA;
  const int64_t N = a->n_pts;
  const EonerfFieldParams* prm = a->params;
  const EonerfFieldParams* G = a->grads;
  const PrepLayout W = prep_layout(a->field, EONERF_PREC_BF16, prm->n_images);
  const FusedPrepLayout F = fused_prep_layout(prm->n_images);
  const uint8_t* ext = (const uint8_t*)a->prepared + W.total;
  const FusedStashLayout S = fused_stash_layout(N, a->density_only);
  const FusedScratchLayout C = fused_scratch_layout(N, prm->n_images, 0);
  const uint8_t* st = (const uint8_t*)a->stash;
  uint8_t* sc = (uint8_t*)a->scratch;
  const bool want_x = a->g_x != nullptr;

  FusedBwdParams p{};
  p.M = N; p.n_tiles = S.n_tiles; p.M_dev = a->n_pts_dev; p.density_only = a->density_only; p.vanilla = vanilla;
  p.wblob = ext + F.bblob; p.consts = (const float*)(ext + F.consts);
  {
    static const int8_t halves[kBwdStages] = {1, 1, 1, 2, 2, 2, 2, 1, 2, 2, 2, 2, 2, 1};
    static const int8_t nkb[kBwdStages] = {2, 2, 2, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4};
    static const int8_t ablk[kBwdStages][4] = {{0, 1, 0, 0}, {2, 3, 0, 0}, {0, 1, 0, 0}, {0, 1, 2, 3}, {0, 1, 2, 3}, {0, 1, 2, 3}, {0, 1, 2, 3},
                                               {0, 1, 2, 3}, {0, 1, 2, 3}, {0, 1, 2, 3}, {0, 1, 2, 3}, {0, 1, 2, 3}, {0, 1, 2, 3}, {0, 1, 2, 3}};
    int blk_off[kBwdStages];
    int o = 0;
    for (int i = 0; i < kBwdStages; ++i) { blk_off[i] = o; o += kBwdBlkCount[i]; }
    int n = 0;
    for (int i = a->density_only ? 5 : (vanilla ? 2 : 0); i < kBwdStages; ++i) {   // vanilla: stage 2 = HD0 (its transient half stays zero)
      if ((i == 7 || i == 13) && !want_x) continue;
      p.mma.st[n].halves = halves[i]; p.mma.st[n].nkb = nkb[i]; p.mma.st[n].blk_off = blk_off[i];
      for (int k = 0; k < 4; ++k) p.mma.st[n].a[k] = ablk[i][k];
      p.prog[n++] = (int8_t)i;
    }
    p.n_prog = n;
    p.mma.n = n;
  }
  for (int i = 0; i < kNumMask; ++i) p.mask[i] = S.mask[i] >= 0 ? (const uint32_t*)(st + S.mask[i]) : nullptr;
  for (int i = 0; i < 13; ++i) p.garr[i] = (a->density_only && i >= 8) ? nullptr : sc + C.g[i];
  p.xf = (const float*)(st + S.xf);
  p.sigma = a->sigma; p.rgb = a->rgb; p.ts = a->transient_s; p.tb = a->transient_beta;
  p.g_sigma = a->g_sigma; p.g_rgb = a->g_rgb; p.g_ts = a->g_transient_s; p.g_tb = a->g_transient_beta;
  p.dpre = (float*)(sc + C.dpre);
  p.g_x = a->g_x;
  const int mode = fused_cta_group();
  const int csz = mode == 1 ? 1 : (mode == 14 ? 4 : 2);
  const int n_ctas = fused_ctas(p.n_tiles, csz);
  const double flops = (double)N * (a->density_only ? 982528.0 : (vanilla ? 1186816.0 : 1345280.0));
  int rc = EONERF_OK;
  CUtensorMap wmap;
  if ((rc = make_blob_map(&wmap, p.wblob, kBwdBlocks)) != EONERF_OK) return rc;
#define EO_LAUNCH_BWD(CG, MC)                                                                                                  \
  do {                                                                                                                         \
    static PerDeviceOnce once;                                                                                            \
    if (once()) {                                                                                                         \
      EO_CUDA(cudaFuncSetAttribute(fused_bwd_kernel<CG, MC, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemFused));    \
      EO_CUDA(cudaFuncSetAttribute(fused_bwd_kernel<CG, MC, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemFused));    \
    }                                                                                                                          \
    profile_begin(4, flops, 0.0, s);                                                                                           \
    rc = (want_x || !deep_ring) ? launch_fused(fused_bwd_kernel<CG, MC, 3>, csz, n_ctas, p, wmap, s)                           \
                                : launch_fused(fused_bwd_kernel<CG, MC, 5>, csz, n_ctas, p, wmap, s);                          \
  } while (0)
  static int deep_ring = -1;                                  // EONERF_BWD_RING5=0 switches the 5-deep ring off (A/B)
  if (deep_ring < 0) { const char* e5 = getenv("EONERF_BWD_RING5"); deep_ring = e5 ? atoi(e5) : 1; }
  if (mode == 1) EO_LAUNCH_BWD(1, 1);
  else if (mode == 2) EO_LAUNCH_BWD(2, 1);
  else if (mode == 14) EO_LAUNCH_BWD(1, 4);
  else EO_LAUNCH_BWD(1, 2);
#undef EO_LAUNCH_BWD
  profile_end(s);
  if (rc != EONERF_OK) return rc;
  EO_LAUNCH_CHECK();
  if (!G) return EONERF_OK;

  // ---- parameter gradients: dW = G^T X over the blocked arrays ----
  GemmTNBlocked gemms[16];
  int n_gemms = 0;
  auto garr = [&](int i) { return (const uint8_t*)(sc + C.g[i]); };
  auto sarr = [&](int i) { return st + S.arr[i]; };
  auto dW = [&](const uint8_t* Gp, int g_nb, int g_blk0, int mt, const uint8_t* Xp, int x_nb, int x_blk0, int x_cnt, int k_valid,
                float* d0, int64_t ld0, float* b0, float* d1, int64_t ld1, float* b1) {
    GemmTNBlocked t;
    t.G = Gp; t.g_nb = g_nb; t.g_blk0 = g_blk0; t.mt_count = mt;
    t.X = Xp; t.x_nb = x_nb; t.x_blk0 = x_blk0; t.x_cnt = x_cnt; t.k_valid = k_valid;
    t.n_tiles = S.n_tiles; t.n_pts_dev = a->n_pts_dev;
    t.D[0] = d0; t.ldd[0] = ld0; t.db[0] = b0; t.D[1] = d1; t.ldd[1] = ld1; t.db[1] = b1;
    gemms[n_gemms++] = t;                                     // launched together at the end (one persistent kernel)
    return (int)EONERF_OK;
  };
  const float* dpre = p.dpre;
  SideFork side(side_stream(s), s);
  EO_TRY(side.fork());
  cudaStream_t hs = side.stream();                          // stream of the head / per-image gradient kernels
  if (!a->density_only && vanilla) {
    // rgb_layer.hidden_layers.0[:, :256] : G_HD0[:, :128]^T BOTT;  [:, 256:283] : G_HD0[:, :128]^T enc4(dir) (per-class sums)
    EO_REQUIRE(a->cond_dirs && a->n_cond > 0, "field_bwd: the fused vanilla field needs the forward call's cond_dirs / n_cond");
    EO_TRY(dW(garr(9), 4, 0, 1, sarr(8), 4, 0, 4, kW, G->head0_w, kW + 27, G->head0_b, nullptr, 0, nullptr));
    EO_TRY(dW(garr(8), 4, 0, 2, sarr(7), 4, 0, 4, kW, G->bott_w, kW, G->bott_b, G->bott_w + (int64_t)kHid * kW, kW, G->bott_b + kHid));
    HeadGradsB ha{{G->head1_w, G->head1_w + kHid, G->head1_w + 2 * kHid}, {G->head1_b, G->head1_b + 1, G->head1_b + 2}};
    EO_TRY(run_heads_dw_blocked<3>(sarr(9), 4, 0, kHid, N, dpre, 1, ha, hs, a->n_pts_dev));
    int64_t blocks = 2 * 148;
    if (blocks > S.n_tiles) blocks = S.n_tiles;
    vanilla_dir_grad_kernel<<<(unsigned)blocks, 256, 0, hs>>>(garr(9), (const int32_t*)(st + S.cls), N, S.n_tiles, a->cond_dirs, a->cond_dirs_stride,
                                                              a->n_cond, G->head0_w, a->n_pts_dev);
    EO_LAUNCH_CHECK();
  }
  if (!a->density_only && !vanilla) {
    // transient_mlp.3 / .2 / .1 : G_T3^T T2, G_T2^T T1, G_T1^T HD0[:,128:256]
    EO_TRY(dW(garr(12), 2, 0, 1, sarr(11), 2, 0, 2, kHid, G->trans_w[3], kHid, G->trans_b[3], nullptr, 0, nullptr));
    EO_TRY(dW(garr(11), 2, 0, 1, sarr(10), 2, 0, 2, kHid, G->trans_w[2], kHid, G->trans_b[2], nullptr, 0, nullptr));
    EO_TRY(dW(garr(10), 2, 0, 1, sarr(9), 4, 2, 2, kHid, G->trans_w[1], kHid, G->trans_b[1], nullptr, 0, nullptr));
    // [albedo_mlp.0 ; transient_mlp.0[:, :256]] : G_HD0^T BOTT
    EO_TRY(dW(garr(9), 4, 0, 2, sarr(8), 4, 0, 4, kW, G->head0_w, kW, G->head0_b, G->trans_w[0], 260, G->trans_b[0]));
    // bottleneck : G_BOTT^T H7
    EO_TRY(dW(garr(8), 4, 0, 2, sarr(7), 4, 0, 4, kW, G->bott_w, kW, G->bott_b, G->bott_w + (int64_t)kHid * kW, kW, G->bott_b + kHid));
    {  // narrow heads
      HeadGradsB hg{{G->ts_w, G->tb_w, nullptr}, {G->ts_b, G->tb_b, nullptr}};
      EO_TRY(run_heads_dw_blocked<2>(sarr(12), 2, 0, kHid, N, dpre, 4, hg, hs, a->n_pts_dev));
      HeadGradsB ha{{G->head1_w, G->head1_w + kHid, G->head1_w + 2 * kHid}, {G->head1_b, G->head1_b + 1, G->head1_b + 2}};
      EO_TRY(run_heads_dw_blocked<3>(sarr(9), 4, 0, kHid, N, dpre, 1, ha, hs, a->n_pts_dev));
    }
    {  // transient embedding / W_t0[:,256:260] through the per-image bias rows
      float* dcb = (float*)(sc + C.dcb);
      EO_CUDA(cudaMemsetAsync(dcb, 0, prm->n_images * kHid * 4, hs));
      const int64_t tab_bytes = prm->n_images * kHid * 4;
      const int use_smem = tab_bytes <= 40 * 1024;
      int64_t blocks = 4 * 148;
      if (blocks > div_up(N, 2048)) blocks = div_up(N, 2048);
      const int64_t rows = (div_up(N, blocks) + 63) & ~(int64_t)63;
      class_grad_blocked_kernel<<<(unsigned)div_up(N, rows), 256, use_smem ? tab_bytes : 0, hs>>>(garr(9), (const int32_t*)(st + S.cls), N, prm->n_images,
                                                                                       dcb, rows, use_smem, a->n_pts_dev);
      EO_LAUNCH_CHECK();
      const int nthreads = (int)(prm->n_images * 4 > kHid * 4 ? prm->n_images * 4 : kHid * 4);
      emb_grad_fused_kernel<<<div_up(nthreads, 128), 128, 0, hs>>>(dcb, prm->trans_w[0], prm->transient_emb, prm->n_images, G->transient_emb,
                                                                   G->trans_w[0]);
      EO_LAUNCH_CHECK();
    }
  }
  {  // sigma head
    HeadGradsB hsig{{G->sigma_w, nullptr, nullptr}, {G->sigma_b, nullptr, nullptr}};
    EO_TRY(run_heads_dw_blocked<1>(sarr(7), 4, 0, kW, N, dpre, 0, hsig, hs, a->n_pts_dev));
  }
  // trunk: layer i reads X = H_{i-1} (layer 5: [H4 | enc], layer 0: enc)
  for (int i = 7; i >= 1; --i) {
    float* w = G->trunk_w[i];
    const int64_t ld = trunk_k(i);
    EO_TRY(dW(garr(i), 4, 0, 2, sarr(i - 1), 4, 0, 4, kW, w, ld, G->trunk_b[i], w + (int64_t)kHid * ld, ld, G->trunk_b[i] + kHid));
    if (i == 5)
      EO_TRY(dW(garr(5), 4, 0, 2, sarr(kArrEnc), 1, 0, 1, 63, w + kW, ld, nullptr, w + (int64_t)kHid * ld + kW, ld, nullptr));
  }
  EO_TRY(dW(garr(0), 4, 0, 2, sarr(kArrEnc), 1, 0, 1, 63, G->trunk_w[0], 63, G->trunk_b[0], G->trunk_w[0] + (int64_t)kHid * 63, 63,
            G->trunk_b[0] + kHid));
  rc = gemm_tn_blocked_group(gemms, n_gemms, s);
  EO_TRY(side.join());                                      // whatever follows on `s` also follows the head kernels
  return rc;
}

}  // namespace eonerf

#ifdef EONERF_TIMING
extern "C" int eonerf_debug_trace_bwd(long long* out) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, eonerf::g_fused_trace, sizeof(long long) * 2048);
  return 0;
}
extern "C" int eonerf_debug_timing_bwd(unsigned long long* out, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, eonerf::g_fused_timing, sizeof(unsigned long long) * 16);
  if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(eonerf::g_fused_timing, z, sizeof(z)); }
  return 0;
}
extern "C" int eonerf_debug_cta_time_bwd(unsigned long long* out) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, eonerf::g_fused_cta_time, sizeof(unsigned long long) * 512);
  return 0;
}
#endif
