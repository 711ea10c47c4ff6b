// Fused radiance-field MLP for sm_100a: EONerfMLP.forward / query_density
// (/root/reference/radiance_fields/eonerf.py:141-170, mlp.py:87-111,190-208) as ONE persistent tcgen05 kernel, and its
// input-gradient chain as a second one.  A 128-sample tile enters the kernel once (positions -> positional encoding in
// shared memory) and every layer's activation stays on chip: the accumulator lives in TMEM, the epilogue warps turn
// it into the next layer's bf16 A operand directly in the 128-byte-swizzled shared-memory image the tensor core reads.
//
//   warp 0      weight producer: streams pre-tiled 16 KB weight blocks (cp.async.bulk, 3-deep mbarrier ring)
//   warp 1      MMA issuer (one thread): tcgen05.mma 128x128x16, accumulators in TMEM (2 tiles x 256 columns)
//   warps 2-9   epilogue: TMEM -> registers -> bias / ReLU / heads -> bf16 -> shared memory (next A operand)
//                         -> cp.async.bulk store of the same image to the tile-blocked stash (training only)
//
// Two tiles ("slots") are in flight per CTA and ping-pong: while the tensor core works on one tile's layer, the
// epilogue warps drain the other tile's accumulator, so the MMA pipe only idles when an epilogue is longer than a layer.
//
// Rounding points are those of the layer-by-layer bf16 path (field.cu): bf16 activations and weights, fp32 accumulate,
// fp32 biases and narrow heads.
#include <stdlib.h>

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

// ---- forward program ------------------------------------------------------------------------------------------------
struct FStage {
  int8_t halves, nkb, a[5], out_blk, relu, kind, mask, pad;
  int16_t bias_off;            // float offset into the constants block
};
// kind: 0 plain, 1 + sigma head, 2 + albedo head (class bias), 3 + transient heads
__constant__ FStage c_fstage[kFwdStages] = {
    {2, 1, {4, 0, 0, 0, 0}, 0, 1, 0, 0, 0, kCBiasTrunk + 0 * 256},
    {2, 4, {0, 1, 2, 3, 0}, 0, 1, 0, 1, 0, kCBiasTrunk + 1 * 256},
    {2, 4, {0, 1, 2, 3, 0}, 0, 1, 0, 2, 0, kCBiasTrunk + 2 * 256},
    {2, 4, {0, 1, 2, 3, 0}, 0, 1, 0, 3, 0, kCBiasTrunk + 3 * 256},
    {2, 4, {0, 1, 2, 3, 0}, 0, 1, 0, 4, 0, kCBiasTrunk + 4 * 256},
    {2, 5, {0, 1, 2, 3, 4}, 0, 1, 0, 5, 0, kCBiasTrunk + 5 * 256},
    {2, 4, {0, 1, 2, 3, 0}, 0, 1, 0, 6, 0, kCBiasTrunk + 6 * 256},
    {2, 4, {0, 1, 2, 3, 0}, 0, 1, 1, 7, 0, kCBiasTrunk + 7 * 256},
    {2, 4, {0, 1, 2, 3, 0}, 0, 0, 0, -1, 0, kCBiasBott},
    {2, 4, {0, 1, 2, 3, 0}, 0, 1, 2, kMaskHd0, 0, kCBiasHd0},
    {1, 2, {2, 3, 0, 0, 0}, 0, 1, 0, kMaskT1 + 0, 0, kCBiasTr + 0 * 128},
    {1, 2, {0, 1, 0, 0, 0}, 2, 1, 0, kMaskT1 + 1, 0, kCBiasTr + 1 * 128},
    {1, 2, {2, 3, 0, 0, 0}, 0, 1, 3, kMaskT1 + 2, 0, kCBiasTr + 2 * 128},
};

struct FusedFwdParams {
  int64_t M; int64_t n_tiles; int n_stages; int training;
  // cls_mode: class of a sample = row of `class_delta` added to the HD0 stage's bias.  0: EO-NeRF, image index of the sample's ray
  // (columns 128..255: transient half); 1 / 2: vanilla field, the ray (1) or the sample itself (2): row of the per-direction bias
  // table (columns 0..127: rgb hidden layer).  vanilla also selects sigma = relu(.) instead of softplus(.)  (mlp.py:243,250).
  int cls_mode; int vanilla;
  const int64_t* M_dev;          // live sample count on the device (M, n_tiles are then capacities)
  const float* x;
  const float* origins; int64_t o_stride; const float* viewdirs; int64_t d_stride;
  const int64_t* ray_indices; const float* t_starts; const float* t_ends; float* z_mid;
  const int64_t* img_idx; int64_t img_stride;
  const uint8_t* wblob; const float* consts; const float* class_delta;   // class_delta: [n_img,128] fp32
  MmaProgram prog;
  uint8_t* arr[kNumArr];
  uint32_t* mask[kNumMask];
  float* xf; int32_t* cls;
  float* sigma; float* rgb; float* ts; float* tb;
};

template <bool kTrain, int kCG, int kMC>
__global__ void __launch_bounds__(kFwdThreads, 1) fused_fwd_kernel(const __grid_constant__ FusedFwdParams p, const __grid_constant__ CUtensorMap wmap) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  float* cst = (float*)(smem + kFOffConst);              // shared-memory copy of the constants (kFwdConstG: unused)
  float* part = (float*)(smem + kFOffPart);
  constexpr int kCl = kCG * kMC;                              // CTAs per cluster
  const uint32_t rank = kCl > 1 ? cluster_ctarank() : 0;
  FusedBars B;
  uint32_t* tmem_base_s;
  EO_CTA_TIME(0);
  fused_setup<kCG, kMC, kFwdRing>(smem, B, tmem_base_s, rank, kFOffBar);
  // encoder handshake (see fused_mma_issuer): enc_full[slot] counts the encoder warp of every CTA whose rows the MMA covers
  uint64_t* const enc_full = (uint64_t*)(smem + kFOffBar + 88);
  uint64_t* const enc_free = enc_full + 2;
  static_assert((2 * kFwdRing + 4) * 8 + 4 <= 88 && 88 + 4 * 8 <= 128, "barrier region");
  if (threadIdx.x == 32) {
    for (int s = 0; s < 2; ++s) { mbar_init(&enc_full[s], (kCG == 2 && rank == 0) ? 2 : 1); mbar_init(&enc_free[s], 1); }
    fence_barrier_init();
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (!kFwdConstG)
    for (int i = threadIdx.x; i < kCFloats; i += kFwdThreads) cst[i] = __ldg(p.consts + i);
  tc_fence_before();
  if (kCl > 1) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_s;
  const int64_t M = p.M_dev ? __ldg(p.M_dev) : p.M;
  const int64_t n_tiles = p.M_dev ? (M + kTileM - 1) / kTileM : p.n_tiles;
  // work items: 2 tiles per CTA of the cluster; this CTA owns tiles (2*kCl*it + 2*rank + slot)
  const int64_t n_items = (n_tiles + 2 * kCl - 1) / (2 * kCl);
  const int64_t it0 = blockIdx.x / kCl, it_stride = gridDim.x / kCl;

  if (warp == 0) {
    if (lane == 0) fused_producer<kCG, kMC, kFwdRing>(p.prog, p.wblob, &wmap, smem, B, it0, n_items, it_stride, rank, p.training & 16);
  } else if (warp == 1) {
    if (kCG == 1 || rank == 0)      // whole warp, converged
      fused_mma_issuer<kCG, kMC, true, kFwdRing, 5>(p.prog, smem, B, tmem_base, it0, n_items, it_stride, enc_full, enc_free);
  } else if (warp == kEncWarpId) {
    // ===== encoder warp: sample positions -> positional encoding of both tiles of every item, one item ahead of the layers.
    // Lane l owns rows l, l+32, l+64, l+96 of each tile.  Also writes the per-sample side outputs (z_mid; training: xf, cls)
    // and, in training, stores the ENC block to the stash. =====
    uint32_t fph = 0;                                // bit `slot` = phase of enc_free[slot]
    for (int64_t it = it0; it < n_items; it += it_stride)
      for (int slot = 0; slot < 2; ++slot) {
        const int64_t tile = 2 * kCl * it + 2 * rank + slot;
        if (kTrain && lane == 0) tma_store_wait_read<0>();     // the previous ENC store has read the block
        mbar_wait(&enc_free[slot], ((fph >> slot) & 1u) ^ 1u);  // ... and so have the MMAs of the previous item (passes at once the first time)
        fph ^= 1u << slot;
        __syncwarp();
        const uint32_t enc = smem_u32(smem + kFOffSlot + slot * kFSlotBytes + 4 * kBlkBytes);
#pragma unroll 1
        for (int rr = 0; rr < 4; ++rr) {
          const int r = rr * 32 + lane;
          const int64_t pt = tile * kTileM + r;
          float x[3] = {0.f, 0.f, 0.f};
          if (pt < M) {
            int64_t ray = -1;
            if (p.x) {
              x[0] = __ldg(p.x + 3 * pt); x[1] = __ldg(p.x + 3 * pt + 1); x[2] = __ldg(p.x + 3 * pt + 2);
              if (p.ray_indices) ray = __ldg(p.ray_indices + pt);
            } else {
              ray = __ldg(p.ray_indices + pt);
              const float ts = __ldg(p.t_starts + pt), te = __ldg(p.t_ends + pt);
              const float zm = __fdiv_rn(__fadd_rn(ts, te), 2.0f);                              // eonerf.py:206
              const float* o = p.origins + ray * p.o_stride;
              const float* dd = p.viewdirs + ray * p.d_stride;
#pragma unroll
              for (int k = 0; k < 3; ++k) x[k] = __fadd_rn(__ldg(o + k), __fmul_rn(__ldg(dd + k), zm));   // eonerf.py:207
              if (p.z_mid) p.z_mid[pt] = zm;
            }
            if (kTrain) {
              p.xf[3 * pt] = x[0]; p.xf[3 * pt + 1] = x[1]; p.xf[3 * pt + 2] = x[2];
              if (p.cls && p.cls_mode == 0 && p.img_idx)
                p.cls[pt] = (int32_t)(p.ray_indices ? __ldg(p.img_idx + ray * p.img_stride) : __ldg(p.img_idx + pt * p.img_stride));
              if (p.cls && p.cls_mode != 0) p.cls[pt] = (int32_t)(p.cls_mode == 1 ? ray : pt);
            }
          }
          uint32_t w[16];
#pragma unroll
          for (int u = 0; u < 16; ++u) w[u] = pack_bf16(posenc_col<0>(x, 2 * u), posenc_col<0>(x, 2 * u + 1));
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) sts_u4(enc + blk_off(r, jj), w[4 * jj], w[4 * jj + 1], w[4 * jj + 2], w[4 * jj + 3]);
#pragma unroll
          for (int u = 0; u < 16; ++u) w[u] = pack_bf16(posenc_col<32>(x, 2 * u), posenc_col<32>(x, 2 * u + 1));
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) sts_u4(enc + blk_off(r, 4 + jj), w[4 * jj], w[4 * jj + 1], w[4 * jj + 2], w[4 * jj + 3]);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          if (kCG == 2 && rank != 0) mbar_arrive_remote(&enc_full[slot], 0); else mbar_arrive(&enc_full[slot]);
          if (kTrain) {
            if (tile < n_tiles) EO_BULK_STORE(p.arr[kArrEnc] + (size_t)tile * kBlkBytes, smem + kFOffSlot + slot * kFSlotBytes + 4 * kBlkBytes, kBlkBytes);
            tma_store_commit();
          }
        }
        __syncwarp();
      }
    if (kTrain && lane == 0) tma_store_wait_all();
  } else if (kDutyWarp && warp == kDutyWarpId) {
    // ===== duty warp: mirrors the epilogue warps' named barriers; after each one, lane 0 hands the slot to the MMA issuer =====
    named_bar_sync(1, kBarThreads);                       // layer-0 bias of the first item written to both accumulators
    if (lane == 0 && it0 < n_items) {
      signal_act_ready<kCG>(B, 0, rank);
      signal_act_ready<kCG>(B, 1, rank);
    }
    __syncwarp();
    for (int64_t it = it0; it < n_items; it += it_stride) {
      const bool next_item = it + it_stride < n_items;
      for (int s = 0; s < p.n_stages; ++s)
        for (int slot = 0; slot < 2; ++slot) {
          if (kTrain && kDutyStores && lane == 0) tma_store_wait_read<0>();
          named_bar_sync(1, kBarThreads);
          if (lane == 0) {
            if (s + 1 < p.n_stages || next_item) signal_act_ready<kCG>(B, slot, rank);
            const int64_t tile = 2 * kCl * it + 2 * rank + slot;
            if (kTrain && kDutyStores && tile < n_tiles && !(p.training & 2)) {
              const FStage d = c_fstage[s];
              const int nb = d.halves * 2;
              for (int bb = 0; bb < nb; ++bb)
                EO_BULK_STORE(p.arr[s] + ((size_t)tile * nb + bb) * kBlkBytes, smem + kFOffSlot + slot * kFSlotBytes + (d.out_blk + bb) * kBlkBytes, kBlkBytes);
              tma_store_commit();
            }
          }
          __syncwarp();
        }
    }
    if (kTrain && kDutyStores && lane == 0) tma_store_wait_all();
  } else {
    // ===== epilogue warps =====
    const int e = threadIdx.x - 64;
    const int q = warp & 3;                         // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;
    const int r = q * 32 + lane;                    // row of the tile
    const int store_id = kDutyStores ? -1 : store_thread_id(e);   // >= 0: this epilogue thread owns bulk stash stores
    const uint32_t s_cst = kFwdConstG ? 0u : smem_u32(cst);     // base "address" of the constants for cf4()
    auto cscal = [&](int i) { return kFwdConstG ? __ldg(p.consts + kCScalars + i) : cst[kCScalars + i]; };
    const uint32_t s_part = smem_u32(part);
    uint32_t cph = 0;                               // bit `slot` = phase of acc_full[slot]
    uint64_t* const acc_full = B.acc_full;
    int tr_i = 0; (void)tr_i;
    // The accumulators start from the layer's bias: every epilogue writes the NEXT layer's bias over the columns it has just
    // drained (tcgen05.st), so no bias add sits between the TMEM load and the bf16 pack.  Layer 0's goes in here for the
    // first item (later items: the last stage of the previous item writes it).
    {
      const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16) + half * 128;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t b[32];
        const uint32_t sb = s_cst + (uint32_t)(c_fstage[0].bias_off + half * 128 + c * 32) * 4u;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float b0, b1, b2, b3;
          cf4<kFwdConstG>(p.consts, sb + j * 16, b0, b1, b2, b3);
          b[4 * j] = __float_as_uint(b0); b[4 * j + 1] = __float_as_uint(b1); b[4 * j + 2] = __float_as_uint(b2); b[4 * j + 3] = __float_as_uint(b3);
        }
        tmem_st32(t0 + c * 32, b);
        tmem_st32(t0 + 256 + c * 32, b);
      }
      tmem_st_wait();
    }
    tc_fence_before();
    named_bar_sync(1, kBarThreads);
    if (!kDutyWarp && e == kSignalThread && it0 < n_items) {
      signal_act_ready<kCG>(B, 0, rank);
      signal_act_ready<kCG>(B, 1, rank);
    }
    for (int64_t it = it0; it < n_items; it += it_stride) {
      const bool next_item = it + it_stride < n_items;
      // image index of this row in slot 0 (low 16 bits) / slot 1: selects the per-image bias row of the HD0 stage (fetched now,
      // used eight stages later)
      uint32_t cls0 = 0u, cls1 = 0u;
      if (p.img_idx || p.cls_mode != 0)
        for (int slot = 0; slot < 2; ++slot) {
          const int64_t pt = (2 * kCl * it + 2 * rank + slot) * kTileM + r;
          if (pt < M) {
            int64_t c;
            if (p.cls_mode == 0) c = p.ray_indices ? __ldg(p.img_idx + __ldg(p.ray_indices + pt) * p.img_stride) : __ldg(p.img_idx + pt * p.img_stride);
            else c = p.cls_mode == 1 ? __ldg(p.ray_indices + pt) : pt;
            if (slot == 0) cls0 = (uint32_t)c; else cls1 = (uint32_t)c;
          }
        }

      // ---- the layers ----
      for (int s = 0; s < p.n_stages; ++s) {
        const FStage d = c_fstage[s];
        const int cpt = d.halves == 2 ? 128 : 64;          // columns per thread
        const int col0 = half * cpt;
        const bool want_mask = kTrain && d.mask >= 0 && !(p.training & 4);
        const bool last_stage = s + 1 == p.n_stages;
        const bool has_next = !last_stage || next_item;     // another MMA stage follows on this slot (maybe the next item's stage 0)
        const FStage dn = c_fstage[last_stage ? 0 : s + 1];
        const int next_cols = dn.halves * 128;              // accumulator columns the next layer uses
        for (int slot = 0; slot < 2; ++slot) {
          const int64_t tile = 2 * kCl * it + 2 * rank + slot;
          const int64_t pt = tile * kTileM + r;
          const bool valid = pt < M;
          const uint32_t act = smem_u32(smem + kFOffSlot + slot * kFSlotBytes);
          // per-image part of the HD0 bias (transient half only): W[:,256:260] . emb[img]
          const float* delta_next = (has_next && dn.kind == 2 && p.class_delta) ? p.class_delta + (size_t)(slot == 0 ? cls0 : cls1) * kHid : nullptr;
          const int delta_col0 = p.vanilla ? 0 : kHid;       // first of the 128 accumulator columns the class row is added to
          EO_TN(ta); { EO_T0(); mbar_wait(&acc_full[slot], (cph >> slot) & 1u); if (e == 0) EO_T1(3); }
          cph ^= 1u << slot;
          tc_fence_after();
          EO_TN(tb);
          EO_TRACE(1, tr_i, threadIdx.x == 64);
          const uint32_t taddr = tmem_base + slot * 256 + ((uint32_t)(q * 32) << 16) + col0;
          float h0 = 0.f, h1 = 0.f, h2 = 0.f;               // head partial sums

          // one 32-column chunk: accumulator + bias -> (ReLU) -> bf16 -> next A operand in shared memory; returns the ReLU keep bits
          // (column 2j -> bit 15-j, column 2j+1 -> bit 31-j: the complement of the pre-activation sign bits, funnel-shifted in)
          // next layer's bias for the 32 accumulator columns of chunk c (per-image rows of the HD0 transient half included)
          auto next_bias_cols = [&](const int colg) {
            if (!has_next || colg >= next_cols) return;
            uint32_t b[32];
            const uint32_t sb = s_cst + (uint32_t)(dn.bias_off + colg) * 4u;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float b0, b1, b2, b3;
              cf4<kFwdConstG>(p.consts, sb + j * 16, b0, b1, b2, b3);
              b[4 * j] = __float_as_uint(b0); b[4 * j + 1] = __float_as_uint(b1); b[4 * j + 2] = __float_as_uint(b2); b[4 * j + 3] = __float_as_uint(b3);
            }
            if (delta_next && colg >= delta_col0 && colg < delta_col0 + kHid) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 t = __ldg((const float4*)(delta_next + colg - delta_col0) + j);
                b[4 * j] = __float_as_uint(__uint_as_float(b[4 * j]) + t.x); b[4 * j + 1] = __float_as_uint(__uint_as_float(b[4 * j + 1]) + t.y);
                b[4 * j + 2] = __float_as_uint(__uint_as_float(b[4 * j + 2]) + t.z); b[4 * j + 3] = __float_as_uint(__uint_as_float(b[4 * j + 3]) + t.w);
              }
            }
            tmem_st32(tmem_base + slot * 256 + ((uint32_t)(q * 32) << 16) + colg, b);
          };
          auto next_bias = [&](const int c) { next_bias_cols(col0 + c * 32); };
          auto chunk = [&](uint32_t (&v)[32], const int c) -> uint32_t {
            float x[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(v[j]);
            uint32_t sign = 0u;
            if (want_mask) {                                  // four independent funnel-shift chains of 8
              uint32_t s0 = 0u, s1 = 0u, s2 = 0u, s3 = 0u;
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                s0 = __funnelshift_l(__float_as_uint(x[2 * j + 1]), s0, 1);        // odd columns 1..15   -> bits 31..24
                s1 = __funnelshift_l(__float_as_uint(x[2 * j + 17]), s1, 1);       // odd columns 17..31  -> bits 23..16
                s2 = __funnelshift_l(__float_as_uint(x[2 * j]), s2, 1);            // even columns 0..14  -> bits 15..8
                s3 = __funnelshift_l(__float_as_uint(x[2 * j + 16]), s3, 1);       // even columns 16..30 -> bits 7..0
              }
              sign = __byte_perm(__byte_perm(s3, s2, 0x0040), __byte_perm(s1, s0, 0x0040), 0x5410);
            }
            uint32_t pk[16];
            if (d.relu) {
#pragma unroll
              for (int j = 0; j < 16; ++j) pk[j] = pack_bf16_relu(x[2 * j], x[2 * j + 1]);
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) pk[j] = pack_bf16(x[2 * j], x[2 * j + 1]);
            }
            if (d.kind == 1) {
              const uint32_t sw = s_cst + (uint32_t)(kCWSigma + col0 + c * 32) * 4u;
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                float w0, w1, w2, w3;
                cf4<kFwdConstG>(p.consts, sw + j * 16, w0, w1, w2, w3);
                h0 = fmaf(bf_lo(pk[2 * j]), w0, h0); h0 = fmaf(bf_hi(pk[2 * j]), w1, h0);
                h0 = fmaf(bf_lo(pk[2 * j + 1]), w2, h0); h0 = fmaf(bf_hi(pk[2 * j + 1]), w3, h0);
              }
            } else if (d.kind == 2) {
              if (half == 0) {
                const uint32_t sw = s_cst + (uint32_t)(kCWAlb + c * 32) * 4u;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  float w0, w1, w2, w3;
                  const float a0 = bf_lo(pk[2 * j]), a1 = bf_hi(pk[2 * j]), a2 = bf_lo(pk[2 * j + 1]), a3 = bf_hi(pk[2 * j + 1]);
                  cf4<kFwdConstG>(p.consts, sw + j * 16, w0, w1, w2, w3);
                  h0 = fmaf(a0, w0, h0); h0 = fmaf(a1, w1, h0); h0 = fmaf(a2, w2, h0); h0 = fmaf(a3, w3, h0);
                  cf4<kFwdConstG>(p.consts, sw + 512 + j * 16, w0, w1, w2, w3);
                  h1 = fmaf(a0, w0, h1); h1 = fmaf(a1, w1, h1); h1 = fmaf(a2, w2, h1); h1 = fmaf(a3, w3, h1);
                  cf4<kFwdConstG>(p.consts, sw + 1024 + j * 16, w0, w1, w2, w3);
                  h2 = fmaf(a0, w0, h2); h2 = fmaf(a1, w1, h2); h2 = fmaf(a2, w2, h2); h2 = fmaf(a3, w3, h2);
                }
              }
            } else if (d.kind == 3) {
              const uint32_t sw = s_cst + (uint32_t)(kCWTs + col0 + c * 32) * 4u;
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                float w0, w1, w2, w3;
                const float a0 = bf_lo(pk[2 * j]), a1 = bf_hi(pk[2 * j]), a2 = bf_lo(pk[2 * j + 1]), a3 = bf_hi(pk[2 * j + 1]);
                cf4<kFwdConstG>(p.consts, sw + j * 16, w0, w1, w2, w3);
                h0 = fmaf(a0, w0, h0); h0 = fmaf(a1, w1, h0); h0 = fmaf(a2, w2, h0); h0 = fmaf(a3, w3, h0);
                cf4<kFwdConstG>(p.consts, sw + (kCWTb - kCWTs) * 4 + j * 16, w0, w1, w2, w3);
                h1 = fmaf(a0, w0, h1); h1 = fmaf(a1, w1, h1); h1 = fmaf(a2, w2, h1); h1 = fmaf(a3, w3, h1);
              }
            }
            const int colg = col0 + c * 32;                   // first output column of this chunk
            const uint32_t blk = act + (uint32_t)(d.out_blk + (colg >> 6)) * kBlkBytes;
            const int ch0 = (colg & 63) >> 3;
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) sts_u4(blk + blk_off(r, ch0 + jj), pk[4 * jj], pk[4 * jj + 1], pk[4 * jj + 2], pk[4 * jj + 3]);
            return ~sign;
          };

          // software pipeline over the chunks: the tcgen05.ld of chunk c+1 is in flight while chunk c is processed
          uint32_t va[32], vb[32];
          uint32_t m0 = 0u, m1 = 0u, m2 = 0u, m3 = 0u;
          if (!(p.training & 8)) {                           // ablation knob 8: no accumulator drain at all (timing experiments only)
          tmem_ld32(taddr, va);
          tmem_ld_wait_dep(va);
          tmem_ld32(taddr + 32, vb);
          next_bias(0);                                      // columns of chunk 0 are drained: refill them
          m0 = chunk(va, 0);
          tmem_ld_wait_dep(vb);
          if (cpt == 128) tmem_ld32(taddr + 64, va);
          next_bias(1);
          m1 = chunk(vb, 1);
          if (cpt == 128) {
            tmem_ld_wait_dep(va);
            tmem_ld32(taddr + 96, vb);
            next_bias(2);
            m2 = chunk(va, 2);
            tmem_ld_wait_dep(vb);
            next_bias(3);
            m3 = chunk(vb, 3);
          }
          if (last_stage && cpt == 64) {
            // a 128-wide last stage drains (and refills) accumulator columns 0..127 only; columns 128..255 have been idle since
            // the HD0 stage: give them the next item's layer-0 bias too
            next_bias_cols(128 + col0);
            next_bias_cols(128 + col0 + 32);
          }
          }
          if (has_next) tmem_st_wait();
          EO_TN(tc);
          EO_TRACE(1, tr_i, threadIdx.x == 64);
          if (want_mask && valid) {
            uint32_t* mrow = p.mask[d.mask] + pt * 8 + col0 / 32;
            if (cpt == 128) *(uint4*)mrow = make_uint4(m0, m1, m2, m3);
            else *(uint2*)mrow = make_uint2(m0, m1);
          }
          if (half == 1 && (d.kind == 1 || d.kind == 3)) sts_f2(s_part + slot * 1024 + r * 8, h0, h1);
          tc_fence_before();
          fence_proxy_async();
          // the OTHER slot's latest stash store (issued one epilogue ago) must have finished reading shared memory before the
          // next epilogue overwrites that slot
          EO_TN(td);
          if (kTrain && store_id >= 0) tma_store_wait_read<0>();
          EO_TN(te);
          { EO_T0(); named_bar_sync(1, kBarThreads); if (e == 32) EO_T1(5); }
          EO_TN(tf);
          EO_TRACE(1, tr_i, threadIdx.x == 64);
          if (!kDutyWarp && e == kSignalThread && has_next) signal_act_ready<kCG>(B, slot, rank);
          if (kTrain && store_id >= 0 && tile < n_tiles && !(p.training & 2)) {
            const int nb = d.halves * 2;                     // 16 KB blocks store_id, store_id + kStoreThreads, ... of the layer's output
            for (int bb = store_id; bb < nb; bb += kStoreThreads)
              EO_BULK_STORE(p.arr[s] + ((size_t)tile * nb + bb) * kBlkBytes, smem + kFOffSlot + slot * kFSlotBytes + (d.out_blk + bb) * kBlkBytes, kBlkBytes);
            tma_store_commit();
          }
          if (valid && half == 0 && d.kind != 0) {
            float p0 = 0.f, p1 = 0.f;
            if (d.kind != 2) lds_f2(s_part + slot * 1024 + r * 8, p0, p1);
            if (d.kind == 1) {
              const float pre = h0 + p0 + cscal(0);
              p.sigma[pt] = p.vanilla ? fmaxf(pre, 0.f) : softplus_f(pre);           // eonerf.py:106,145 / mlp.py:243,250
            } else if (d.kind == 2) {
              p.rgb[3 * pt + 0] = sigmoid_f(h0 + cscal(1));
              p.rgb[3 * pt + 1] = sigmoid_f(h1 + cscal(2));
              p.rgb[3 * pt + 2] = sigmoid_f(h2 + cscal(3));
            } else {
              p.ts[pt] = sigmoid_f(h0 + p0 + cscal(4));
              p.tb[pt] = softplus_f(h1 + p1 + cscal(5));
            }
          }
          EO_TN(tg);
          EO_TD(7, tb, tc); EO_TD(8, tc, td); EO_TD(9, td, te); EO_TD(10, te, tf); EO_TD(11, tf, tg); EO_TD(12, ta, tg);
        }
      }
    }
    if (store_id >= 0) tma_store_wait_all();
  }
  fused_teardown<kCG, kMC>(tmem_base);
  EO_CTA_TIME(1);
}

// ---- prepare: weight blocks + constants ----------------------------------------------------------------------------
struct PackJob { const __nv_bfloat16* src; int ld, row0, rows, col0; };
struct PackJobs { PackJob j[kBwdBlocks]; int n; };

// block image [128 x 64]: element (n, k) = src[(row0+n)*ld + col0+k] for n < rows (zero beyond), swizzled
__global__ void __launch_bounds__(256) pack_blocks_kernel(const __grid_constant__ PackJobs jobs, uint8_t* __restrict__ dst) {
  const int b = blockIdx.x >> 2;
  const int idx = (blockIdx.x & 3) * 256 + threadIdx.x;     // 1024 16-byte chunks per block
  const int n = idx >> 3, c = idx & 7;
  const PackJob j = jobs.j[b];
  uint4 v = make_uint4(0, 0, 0, 0);
  if (n < j.rows) v = __ldg((const uint4*)(j.src + (size_t)(j.row0 + n) * j.ld + j.col0 + c * 8));
  *(uint4*)(dst + (size_t)b * kBlkBytes + n * 128 + ((c ^ (n & 7)) << 4)) = v;
}

struct ConstSrc { const float* bt[8]; const float* bb; const float* btr[3]; const float* ws; const float* wa; const float* wts; const float* wtb;
                  const float* bs; const float* ba; const float* bts; const float* btb; const float* bh0; const float* bt0; };
__global__ void pack_consts_kernel(ConstSrc c, float* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kCFloats) return;
  float v = 0.f;
  if (i < kCBiasBott) v = c.bt[i >> 8][i & 255];
  else if (i < kCBiasTr) v = c.bb[i - kCBiasBott];
  else if (i < kCWSigma) v = c.btr[(i - kCBiasTr) >> 7][(i - kCBiasTr) & 127];
  else if (i < kCWAlb) v = c.ws[i - kCWSigma];
  else if (i < kCWTs) v = c.wa[i - kCWAlb];
  else if (i < kCWTb) v = c.wts[i - kCWTs];
  else if (i < kCScalars) v = c.wtb[i - kCWTb];
  else if (i == kCScalars) v = c.bs[0];
  else if (i < kCScalars + 4) v = c.ba[i - kCScalars - 1];
  else if (i == kCScalars + 4) v = c.bts[0];
  else if (i == kCScalars + 5) v = c.btb[0];
  else if (i >= kCBiasHd0 && i < kCBiasHd0 + kHid) v = c.bh0[i - kCBiasHd0];
  else if (i >= kCBiasHd0 + kHid) v = c.bt0[i - kCBiasHd0 - kHid];
  dst[i] = v;
}

// delta[img, j] = sum_e W_t0[j, 256+e] emb[img, e]
__global__ void class_delta_kernel(const float* __restrict__ wt0, const float* __restrict__ emb, int64_t n_images, float* __restrict__ delta) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_images * kHid) return;
  const int img = idx / kHid, j = idx % kHid;
  float v = 0.f;
#pragma unroll
  for (int e = 0; e < 4; ++e) v = fmaf(__ldg(wt0 + j * 260 + 256 + e), __ldg(emb + img * 4 + e), v);
  delta[idx] = v;
}

}  // namespace

int64_t fused_prepared_extra_bytes(int64_t n_images) { return fused_prep_layout(n_images).total; }
int64_t fused_stash_bytes(int64_t n_pts, int density_only) { return fused_stash_layout(n_pts, density_only).total; }
int64_t fused_scratch_bytes(int64_t n_pts, int64_t n_images, int density_only) {
  return fused_scratch_layout(n_pts, n_images, density_only).total;
}

// MMA side of the forward program (host copy of the A-block lists of c_fstage)
static MmaProgram fwd_program(int n_stages) {
  static const int8_t halves[kFwdStages] = {2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 1, 1, 1};
  static const int8_t nkb[kFwdStages] = {1, 4, 4, 4, 4, 5, 4, 4, 4, 4, 2, 2, 2};
  static const int8_t a[kFwdStages][5] = {{4, 0, 0, 0, 0}, {0, 1, 2, 3, 0}, {0, 1, 2, 3, 0}, {0, 1, 2, 3, 0}, {0, 1, 2, 3, 0}, {0, 1, 2, 3, 4},
                                          {0, 1, 2, 3, 0}, {0, 1, 2, 3, 0}, {0, 1, 2, 3, 0}, {0, 1, 2, 3, 0}, {2, 3, 0, 0, 0}, {0, 1, 0, 0, 0},
                                          {2, 3, 0, 0, 0}};
  MmaProgram P{};
  P.n = n_stages;
  int off = 0;
  for (int s = 0; s < kFwdStages; ++s) {
    if (s < n_stages) {
      P.st[s].halves = halves[s]; P.st[s].nkb = nkb[s]; P.st[s].blk_off = off;
      for (int k = 0; k < 5; ++k) P.st[s].a[k] = a[s][k];
    }
    off += halves[s] * nkb[s];
  }
  return P;
}

// How the CTAs of the fused kernels cooperate (EONERF_FUSED_MODE overrides, for A/B timing):
//    1  every CTA on its own
//    2  CTA pairs issuing cta_group::2 MMAs (each CTA streams half of every weight block)            <- default
//   12  clusters of 2 sharing the weight stream by TMA multicast (independent cta_group::1 MMAs)
//   14  clusters of 4 sharing the weight stream
int fused_cta_group() {
  static int mode = 0;
  if (!mode) {
    const char* e = getenv("EONERF_FUSED_MODE");
    mode = e ? atoi(e) : 2;
    if (mode != 1 && mode != 2 && mode != 12 && mode != 14) mode = 2;
  }
  return mode;
}

// dir_bias[r, j] = sum_e W_rgb0[j, 256 + e] enc4(dir_r)[e]   (mlp.py:153-165,199-205: the view-direction columns of
// rgb_layer.hidden_layers.0 folded into a bias row per conditioning row; fp32)
__global__ void __launch_bounds__(128) vanilla_dir_bias_kernel(const float* __restrict__ dirs, int64_t stride, int64_t n, const float* __restrict__ w,
                                                               float* __restrict__ out) {
  __shared__ float enc[32];
  const int64_t r = blockIdx.x;
  if (r >= n) return;
  const int t = threadIdx.x;
  if (t < 32) {
    float v = 0.f;
    if (t < 3) v = __ldg(dirs + r * stride + t);
    else if (t < 27) {
      int e = t - 3;
      const int half = e >= 12;
      e -= half * 12;
      const float xb = __ldg(dirs + r * stride + e % 3) * (float)(1 << (e / 3));
      v = sinf(half ? __fadd_rn(xb, kHalfPi) : xb);
    }
    enc[t] = v;
  }
  __syncthreads();
  float acc = 0.f;
#pragma unroll
  for (int e = 0; e < 27; ++e) acc = fmaf(__ldg(w + t * 283 + 256 + e), enc[e], acc);
  out[r * kHid + t] = acc;
}

int fused_prepare(int field, const EonerfFieldParams* p, void* prepared, cudaStream_t s) {
  const bool vanilla = field == EONERF_FIELD_VANILLA;
  const PrepLayout W = prep_layout(field, EONERF_PREC_BF16, p->n_images);
  const FusedPrepLayout F = fused_prep_layout(p->n_images);
  uint8_t* base = (uint8_t*)prepared;
  uint8_t* ext = base + W.total;
  auto bf = [&](int64_t off) { return (const __nv_bfloat16*)(base + off); };
  // the vanilla field runs the first ten stages of the same program: its rgb hidden layer takes the place of the albedo half of
  // the HD0 stage (the other half, and the three transient stages, are zero blocks)
  {  // forward blocks: stage-major, then output half, then k block;  B = W [out, Kp]
    PackJobs J{};
    int n = 0;
    auto add = [&](const __nv_bfloat16* src, int ld, int out_rows, int kp) {
      for (int h = 0; h < out_rows / 128; ++h)
        for (int kb = 0; kb < kp / 64; ++kb) J.j[n++] = PackJob{src, ld, h * 128, 128, kb * 64};
    };
    auto zeros = [&](int count) { for (int i = 0; i < count; ++i) J.j[n++] = PackJob{bf(W.w[0]), 64, 0, 0, 0}; };
    for (int i = 0; i < 8; ++i) add(bf(W.w[i]), trunk_kp(i), kW, trunk_kp(i));
    add(bf(W.bott), kW, kW, kW);
    if (!vanilla) {
      add(bf(W.hd0), kW, kW, kW);
      for (int i = 0; i < 3; ++i) add(bf(W.tr[i]), kHid, kHid, kHid);
    } else {
      for (int kb = 0; kb < 4; ++kb) J.j[n++] = PackJob{bf(W.hd0), kW + kDirEnc, 0, 128, kb * 64};   // rgb hidden [128, 256 | dir 32]
      zeros(4 + 3 * 2);
    }
    if (n != kFwdBlocks) { set_error("fused_prepare: forward block count %d != %d", n, kFwdBlocks); return EONERF_EINVAL; }
    J.n = n;
    pack_blocks_kernel<<<n * 4, 256, 0, s>>>(J, ext + F.fblob);
    EO_LAUNCH_CHECK();
  }
  {  // backward blocks: B = W^T [in, out]
    PackJobs J{};
    int n = 0;
    auto add = [&](const __nv_bfloat16* src, int ld, int row0, int rows, int out_cols) {
      for (int h = 0; h < (rows + 127) / 128; ++h)
        for (int kb = 0; kb < out_cols / 64; ++kb) {
          const int rr = rows - h * 128 < 128 ? rows - h * 128 : 128;
          J.j[n++] = PackJob{src, ld, row0 + h * 128, rr, kb * 64};
        }
    };
    auto zeros = [&](int count) { for (int i = 0; i < count; ++i) J.j[n++] = PackJob{bf(W.w[0]), 64, 0, 0, 0}; };
    if (!vanilla) {
      for (int i = 2; i >= 0; --i) add(bf(W.tr_t[i]), kHid, 0, kHid, kHid);      // S0..S2
      add(bf(W.hd0_t), 2 * kHid, 0, kW, 2 * kHid);                               // S3
    } else {
      zeros(3 * 2);                                                              // S0..S2: no transient branch
      for (int h = 0; h < 2; ++h)                                                // S3: W_rgb0[:, :256]^T = [256 in, 128 out | 128 zero]
        for (int kb = 0; kb < 4; ++kb) {
          if (kb < 2) J.j[n++] = PackJob{bf(W.hd0_t), kHid, h * 128, 128, kb * 64};
          else zeros(1);
        }
    }
    add(bf(W.bott_t), kW, 0, kW, kW);                                          // S4
    add(bf(W.wt[7]), kW, 0, kW, kW);                                           // S5
    add(bf(W.wt[6]), kW, 0, kW, kW);                                           // S6
    add(bf(W.wt[5]), kW, kW, kEnc, kW);                                        // S7e: encoding rows of W5^T
    add(bf(W.wt[5]), kW, 0, kW, kW);                                           // S7
    for (int i = 4; i >= 1; --i) add(bf(W.wt[i]), kW, 0, kW, kW);              // S8..S11
    add(bf(W.wt[0]), kW, 0, kEnc, kW);                                         // S12
    if (n != kBwdBlocks) { set_error("fused_prepare: backward block count %d != %d", n, kBwdBlocks); return EONERF_EINVAL; }
    J.n = n;
    pack_blocks_kernel<<<n * 4, 256, 0, s>>>(J, ext + F.bblob);
    EO_LAUNCH_CHECK();
  }
  ConstSrc c{};
  for (int i = 0; i < 8; ++i) c.bt[i] = p->trunk_b[i];
  c.bb = p->bott_b;
  c.ws = p->sigma_w; c.wa = p->head1_w; c.bs = p->sigma_b; c.ba = p->head1_b; c.bh0 = p->head0_b;
  if (!vanilla) {
    for (int i = 0; i < 3; ++i) c.btr[i] = p->trans_b[i + 1];
    c.wts = p->ts_w; c.wtb = p->tb_w; c.bts = p->ts_b; c.btb = p->tb_b; c.bt0 = p->trans_b[0];
  } else {
    // no transient branch: its constants read a zeroed 128-float row (the unused per-image table of the extras)
    float* z = (float*)(ext + F.delta);
    EO_CUDA(cudaMemsetAsync(z, 0, kHid * sizeof(float), s));
    for (int i = 0; i < 3; ++i) c.btr[i] = z;
    c.wts = z; c.wtb = z; c.bts = z; c.btb = z; c.bt0 = z;
  }
  pack_consts_kernel<<<div_up(kCFloats, 256), 256, 0, s>>>(c, (float*)(ext + F.consts));
  EO_LAUNCH_CHECK();
  if (!vanilla) {
    class_delta_kernel<<<div_up(p->n_images * kHid, 128), 128, 0, s>>>(p->trans_w[0], p->transient_emb, p->n_images, (float*)(ext + F.delta));
    EO_LAUNCH_CHECK();
  }
  return EONERF_OK;
}

int fused_field_fwd(const EonerfFieldFwdArgs* a, cudaStream_t s) {
  const bool vanilla = a->field == EONERF_FIELD_VANILLA;
  const int64_t N = a->n_pts;
  const EonerfFieldParams* prm = a->params;
  const PrepLayout W = prep_layout(a->field, EONERF_PREC_BF16, prm->n_images);
  const FusedPrepLayout F = fused_prep_layout(prm->n_images);
  const uint8_t* ext = (const uint8_t*)a->prepared + W.total;
  const bool train = a->stash != nullptr;
  if (!train && !vanilla && fused_cta_group() == 2 && fused_ts_enabled()) return fused_field_fwd_ts(a, s);   // inference: activations in TMEM
  FusedFwdParams p{};
  p.M = N;
  p.M_dev = a->n_pts_dev;
  p.n_tiles = (N + kTileM - 1) / kTileM;
  p.n_stages = a->density_only ? 8 : (vanilla ? 10 : kFwdStages);
  p.vanilla = vanilla;
  if (vanilla && !a->density_only) {
    EO_REQUIRE(a->cond_dirs && a->dir_bias && a->n_cond > 0, "field_fwd: the fused vanilla field needs cond_dirs, dir_bias and n_cond");
    EO_REQUIRE(!a->cond_dirs_per_ray || a->ray_indices, "field_fwd: cond_dirs_per_ray needs ray_indices");
    p.cls_mode = a->cond_dirs_per_ray ? 1 : 2;
    vanilla_dir_bias_kernel<<<(unsigned)a->n_cond, 128, 0, s>>>(a->cond_dirs, a->cond_dirs_stride, a->n_cond, prm->head0_w, a->dir_bias);
    EO_LAUNCH_CHECK();
  }
  p.training = train;
  if (const char* dbg = getenv("EONERF_FUSED_DBG")) p.training |= atoi(dbg);   // experiment knobs: 2 = skip stash stores, 4 = skip masks, 8 = skip the accumulator drain, 16 = skip the weight loads
  p.x = a->x;
  p.origins = a->origins; p.o_stride = a->origins_stride; p.viewdirs = a->viewdirs; p.d_stride = a->viewdirs_stride;
  p.ray_indices = a->ray_indices; p.t_starts = a->t_starts; p.t_ends = a->t_ends; p.z_mid = a->z_mid;
  p.img_idx = a->density_only ? nullptr : a->img_idx; p.img_stride = a->img_idx_stride;
  p.wblob = ext + F.fblob; p.consts = (const float*)(ext + F.consts);
  p.class_delta = vanilla ? (a->density_only ? nullptr : a->dir_bias) : (const float*)(ext + F.delta);
  p.prog = fwd_program(p.n_stages);
  if (train) {
    const FusedStashLayout S = fused_stash_layout(N, a->density_only);
    uint8_t* st = (uint8_t*)a->stash;
    for (int i = 0; i < kNumArr; ++i) p.arr[i] = S.arr[i] >= 0 ? st + S.arr[i] : nullptr;
    for (int i = 0; i < kNumMask; ++i) p.mask[i] = S.mask[i] >= 0 ? (uint32_t*)(st + S.mask[i]) : nullptr;
    p.xf = (float*)(st + S.xf);
    p.cls = a->density_only ? nullptr : (int32_t*)(st + S.cls);
  }
  p.sigma = a->sigma; p.rgb = a->rgb; p.ts = a->transient_s; p.tb = a->transient_beta;
  const int mode = fused_cta_group();
  const int csz = mode == 1 ? 1 : (mode == 14 ? 4 : 2);
  const int n_ctas = fused_ctas(p.n_tiles, csz);
  const double flops = (double)N * (a->density_only ? 982528.0 : (vanilla ? 1186816.0 : 1345280.0));
  int rc = EONERF_OK;
  CUtensorMap wmap;
  if ((rc = make_blob_map(&wmap, p.wblob, kFwdBlocks)) != EONERF_OK) return rc;
#define EO_LAUNCH_FWD(CG, MC)                                                                                                   \
  do {                                                                                                                          \
    static PerDeviceOnce once;                                                                                             \
    if (once()) {                                                                                                          \
      EO_CUDA(cudaFuncSetAttribute(fused_fwd_kernel<true, CG, MC>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemFwd));  \
      EO_CUDA(cudaFuncSetAttribute(fused_fwd_kernel<false, CG, MC>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemFwd)); \
    }                                                                                                                           \
    profile_begin(3, flops, 0.0, s);                                                                                            \
    rc = train ? launch_fused(fused_fwd_kernel<true, CG, MC>, csz, n_ctas, p, wmap, s, kSmemFwd, kFwdThreads)                \
               : launch_fused(fused_fwd_kernel<false, CG, MC>, csz, n_ctas, p, wmap, s, kSmemFwd, kFwdThreads);              \
  } while (0)
  if (mode == 1) EO_LAUNCH_FWD(1, 1);
  else if (mode == 2) EO_LAUNCH_FWD(2, 1);
  else if (mode == 14) EO_LAUNCH_FWD(1, 4);
  else EO_LAUNCH_FWD(1, 2);
#undef EO_LAUNCH_FWD
  profile_end(s);
  if (rc != EONERF_OK) return rc;
  EO_LAUNCH_CHECK();
  return EONERF_OK;
}

}  // namespace eonerf

#ifdef EONERF_TIMING
extern "C" int eonerf_debug_trace(long long* out) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, eonerf::g_fused_trace, sizeof(long long) * 2048);
  return 0;
}
extern "C" int eonerf_debug_timing(unsigned long long* out, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, eonerf::g_fused_timing, sizeof(unsigned long long) * 16);
  if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(eonerf::g_fused_timing, z, sizeof(z)); }
  return 0;
}
extern "C" int eonerf_debug_cta_time_fwd(unsigned long long* out) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, eonerf::g_fused_cta_time, sizeof(unsigned long long) * 512);
  return 0;
}
#endif
