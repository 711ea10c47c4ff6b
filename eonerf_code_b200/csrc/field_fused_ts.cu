// Inference form of the fused radiance-field MLP with the ACTIVATIONS IN TENSOR MEMORY ("TS" tcgen05.mma: A from TMEM, B from
// shared memory).  EONerfMLP.forward / query_density without a stash (/root/reference/radiance_fields/eonerf.py:141-170,
// mlp.py:87-111,190-208): what the evaluation render and the occupancy-grid update run.
//
// Why: the shared-memory pipe bounds the SS form (field_fused.cu): per 128-row tile and 256-wide
// layer the tensor core reads A (64 KB) and B (64 KB) from shared memory while the epilogue writes the next A (64 KB) and TMA writes
// the weights (64 KB): 256+ KB against a 128 B/cycle pipe and a 2048-cycle MMA budget.  Here a tile's activations never leave
// tensor memory: the epilogue turns an accumulator into the next layer's bf16 A operand with tcgen05.ld -> cvt -> tcgen05.st, the
// tensor core reads A from TMEM, and shared memory only carries the weight stream (128 KB per layer).
//
// TMEM (512 columns per CTA): accumulator 2 x 128 columns (the two N-halves of a 256-wide layer), A operand 2 x 128 columns
// (bf16 pairs: K = 256; double-buffered between consecutive layers).  One 128-row tile per CTA (256 rows per CTA pair,
// cta_group::2), software-pipelined INSIDE the layer instead of across two tiles:
//   * a layer is issued as two N-halves; with cta_group::2 each CTA supplies 64 weight rows per half, so accumulator half h holds
//     output columns {h*64 .. +64} and {128 + h*64 .. +64}: draining half 0 yields K-blocks 0 and 2 of the next layer's A operand,
//     half 1 yields K-blocks 1 and 3;
//   * while the epilogue warps drain half 0, the tensor core computes half 1; while they drain half 1, it already runs the next
//     layer's half 0 over K-blocks 0 and 2 (K-partial issue), then K-blocks 1 and 3, then half 1.
//   warp 0 weight producer (8 KB sub-blocks, deep ring)   warp 1 MMA issuer   warp 2 encoder (positional encoding of the next
//   tile, one tile ahead)   warp 3 idle   warps 4-11 epilogue (lane quarter = warp & 3, column group = (warp - 4) >> 2)
// Rounding points are those of the other bf16 paths: bf16 activations and weights, fp32 accumulate, fp32 biases and narrow heads.
#include <stdlib.h>

#include "fused_common.cuh"

namespace eonerf {

namespace {

constexpr int kTsThreads = 12 * 32;
constexpr int kTsEpiThreads = 8 * 32;
constexpr int kTsRing = 8;                        // weight ring: 8 slots of two 8 KB sub-blocks
constexpr int kSubBytes = kBlkBytes / 2;          // one sub-block = 64 weight rows x 64 K
constexpr int kTsOffRing = 0;
constexpr int kTsOffEnc = kTsRing * kBlkBytes;    // 131072
constexpr int kTsOffConst = kTsOffEnc + kBlkBytes;
constexpr int kTsOffPart = kTsOffConst + kConstBytes;          // [128 rows][4] floats
constexpr int kTsOffBar = kTsOffPart + 128 * 4 * 4;
constexpr int kSmemTs = kTsOffBar + 512 + 1024;   // + alignment slack
// TMEM columns
constexpr uint32_t kTmAcc = 0, kTmA0 = 256, kTmA1 = 384;

// one MMA group = one ring slot = one or two weight sub-blocks (64 rows per CTA x 64 K), each against one K block of A
struct TsStep {
  int8_t half;        // accumulator half (0 / 1)
  int8_t n;           // sub-blocks in the slot (1 / 2)
  int8_t kb[2];       // weight block index inside the stage (blob order)
  int8_t a_kb[2];     // K block of the A operand in TMEM, or -1: the ENC block in shared memory
  int8_t wait;        // before issuing: 1 wait half_ready[0], 2 wait half_ready[1], 4 wait enc_full
  int8_t commit;      // after issuing: 1 acc_full[0], 2 acc_full[1], 4 enc_free
};
struct TsStage { int8_t n_steps, wide, nkb, pad; int32_t blk_off; TsStep st[6]; };
struct TsProgram { int32_t n; TsStage s[kFwdStages]; };

struct TsStageEpi { int8_t wide, relu, kind, pad; int16_t bias_off; };
__constant__ TsStageEpi c_ts_epi[kFwdStages] = {
    {1, 1, 0, 0, kCBiasTrunk + 0 * 256}, {1, 1, 0, 0, kCBiasTrunk + 1 * 256}, {1, 1, 0, 0, kCBiasTrunk + 2 * 256},
    {1, 1, 0, 0, kCBiasTrunk + 3 * 256}, {1, 1, 0, 0, kCBiasTrunk + 4 * 256}, {1, 1, 0, 0, kCBiasTrunk + 5 * 256},
    {1, 1, 0, 0, kCBiasTrunk + 6 * 256}, {1, 1, 1, 0, kCBiasTrunk + 7 * 256}, {1, 0, 0, 0, kCBiasBott},
    {1, 1, 2, 0, kCBiasHd0},             {0, 1, 0, 0, kCBiasTr + 0 * 128},    {0, 1, 0, 0, kCBiasTr + 1 * 128},
    {0, 1, 3, 0, kCBiasTr + 2 * 128},
};

struct FusedTsParams {
  int64_t M; int64_t n_tiles; int n_stages; int dbg;   // dbg: timing experiments only (1 no drain, 2 no bias prefill)
  const int64_t* M_dev;
  const float* x;
  const float* origins; int64_t o_stride; const float* viewdirs; int64_t d_stride;
  const int64_t* ray_indices; const float* t_starts; const float* t_ends; float* z_mid;
  const int64_t* img_idx; int64_t img_stride;
  const float* consts; const float* class_delta;
  TsProgram prog;
  float* sigma; float* rgb; float* ts; float* tb;
};

struct TsBars { uint64_t* w_full; uint64_t* w_empty; uint64_t* acc_full; uint64_t* half_ready; uint64_t* enc_full; uint64_t* enc_free; };

__device__ __forceinline__ void umma_ts_2cta(uint32_t tmem_d, uint32_t tmem_a, uint32_t lo_b, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
      "setp.ne.b32 p, 1, 0;\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], db, %4, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "r"(lo_b), "r"(kDescHiK128), "r"(idesc)
      : "memory");
}

__global__ void __launch_bounds__(kTsThreads, 1) fused_fwd_ts_kernel(const __grid_constant__ FusedTsParams p, const __grid_constant__ CUtensorMap wmap) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  float* cst = (float*)(smem + kTsOffConst);
  float* part = (float*)(smem + kTsOffPart);
  const uint32_t rank = cluster_ctarank();
  TsBars B;
  B.w_full = (uint64_t*)(smem + kTsOffBar);
  B.w_empty = B.w_full + kTsRing;
  B.acc_full = B.w_empty + kTsRing;
  B.half_ready = B.acc_full + 2;
  B.enc_full = B.half_ready + 2;
  B.enc_free = B.enc_full + 1;
  uint32_t* tmem_base_s = (uint32_t*)(B.enc_free + 1);
  static_assert((2 * kTsRing + 6) * 8 + 4 <= 512, "barrier region");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    const uint32_t n_arr = rank == 0 ? 2 : 1;              // leader: + the peer CTA's remote arrival
    for (int s = 0; s < kTsRing; ++s) { mbar_init(&B.w_full[s], 1); mbar_init(&B.w_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&B.acc_full[s], 1); mbar_init(&B.half_ready[s], n_arr); }
    mbar_init(B.enc_full, n_arr);
    mbar_init(B.enc_free, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2cta(tmem_base_s, 512);
  for (int i = threadIdx.x; i < kCFloats; i += kTsThreads) cst[i] = __ldg(p.consts + i);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_s;
  const int64_t M = p.M_dev ? __ldg(p.M_dev) : p.M;
  const int64_t n_tiles = p.M_dev ? (M + kTileM - 1) / kTileM : p.n_tiles;
  // work items: one 128-row tile per CTA, two per pair; this CTA owns tile 2 * it + rank
  const int64_t n_items = (n_tiles + 1) / 2;
  const int64_t it0 = blockIdx.x / 2, it_stride = gridDim.x / 2;

  if (warp == 0) {
    // ===== weight producer: 8 KB sub-blocks in the order the MMA issuer consumes them =====
    if (lane == 0) {
      int rs = 0; uint32_t rph = 0;
      int filled = 0;
      for (int64_t it = it0; it < n_items; it += it_stride)
        for (int s = 0; s < p.prog.n; ++s) {
          const TsStage& S = p.prog.s[s];
          for (int k = 0; k < S.n_steps; ++k) {
            const TsStep st = S.st[k];
            mbar_wait(&B.w_empty[rs], rph ^ 1);
            if ((p.dbg & 8) && filled >= kTsRing) {
              if (rank == 0) mbar_arrive(&B.w_full[rs]);
              if (++rs == kTsRing) { rs = 0; rph ^= 1; }
              continue;
            }
            ++filled;
            if (rank == 0) mbar_expect_tx(&B.w_full[rs], 2 * st.n * kSubBytes);
            for (int u = 0; u < st.n; ++u) {
              // wide stage: block (rank, kb) of the stage, rows half*64..; 128-wide stage: block kb, rows rank*64..
              const int row0 = S.wide ? (S.blk_off + (int)rank * S.nkb + st.kb[u]) * 128 + st.half * 64 : (S.blk_off + st.kb[u]) * 128 + (int)rank * 64;
              tma_load_2d_2cta(smem + kTsOffRing + rs * kBlkBytes + u * kSubBytes, &wmap, &B.w_full[rs], 0, row0);
            }
            if (++rs == kTsRing) { rs = 0; rph ^= 1; }
          }
        }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (leader CTA of the pair; whole warp converged, one elected lane issues) =====
    if (rank == 0) {
      int rs = 0; uint32_t rph = 0;
      uint32_t hph = 0, eph = 0;
      int ti = 0; (void)ti;
      const uint32_t ring0 = smem_u32(smem + kTsOffRing);
      const uint32_t enc_lo = desc_lo_k128(smem_u32(smem + kTsOffEnc));
      const bool elected = elect_one_sync();
      const uint32_t idesc = instr_desc(256, 128, 0, 0);
      for (int64_t it = it0; it < n_items; it += it_stride)
        for (int s = 0; s < p.prog.n; ++s) {
          const TsStage& S = p.prog.s[s];
          const uint32_t a_in = tmem_base + ((s & 1) ? kTmA0 : kTmA1);       // stage s reads what stage s-1 wrote: A[(s-1) & 1]
          for (int k = 0; k < S.n_steps; ++k) {
            const TsStep st = S.st[k];
            EO_TRACE(0, ti, lane == 0);
            if (p.dbg & 32) {} else
            if (st.wait & 1) { mbar_wait(&B.half_ready[0], hph & 1u); hph ^= 1u; }
            if (p.dbg & 32) {} else
            if (st.wait & 2) { mbar_wait(&B.half_ready[1], (hph >> 1) & 1u); hph ^= 2u; }
            if (p.dbg & 32) {} else
            if (st.wait & 4) { mbar_wait(B.enc_full, eph); eph ^= 1u; }
            mbar_wait(&B.w_full[rs], rph);
            tc_fence_after();
            EO_TRACE(0, ti, lane == 0);
            const uint32_t d_tmem = tmem_base + kTmAcc + st.half * 128;
            const uint32_t d_alt = tmem_base + kTmAcc + (st.half ^ 1) * 128;
            if (elected) {
              for (int u = 0; u < st.n; ++u) {
                const uint32_t lb = desc_lo_k128(ring0 + rs * kBlkBytes + u * kSubBytes);
                if (st.a_kb[u] >= 0 && !(p.dbg & 4)) {
                  const uint32_t a0 = a_in + st.a_kb[u] * 32;
#pragma unroll
                  for (int q = 0; q < 4; ++q) umma_ts_2cta((p.dbg & 16) && (q & 1) ? d_alt : d_tmem, a0 + q * 8, lb + 2 * q, idesc);   // 16 K elements = 8 TMEM columns = 32 B of B
                } else {
#pragma unroll
                  for (int q = 0; q < 4; ++q) umma_k128<2>((p.dbg & 16) && (q & 1) ? d_alt : d_tmem, enc_lo + 2 * q, lb + 2 * q, idesc, 1);
                }
              }
              umma_commit_2cta(&B.w_empty[rs]);
              if (st.commit & 1) umma_commit_2cta(&B.acc_full[0]);
              if (st.commit & 2) umma_commit_2cta(&B.acc_full[1]);
              if (st.commit & 4) umma_commit_2cta(B.enc_free);
            }
            EO_TRACE(0, ti, lane == 0);
            __syncwarp();
            if (++rs == kTsRing) { rs = 0; rph ^= 1; }
          }
        }
    }
  } else if (warp == 2) {
    // ===== encoder warp: positions -> positional encoding of the next tile (shared memory, SS operand of layers 0 and 5) =====
    uint32_t fph = 0;
    for (int64_t it = it0; it < n_items && !(p.dbg & 32); it += it_stride) {
      const int64_t tile = 2 * it + rank;
      mbar_wait(B.enc_free, fph ^ 1u);
      fph ^= 1u;
      __syncwarp();
      const uint32_t enc = smem_u32(smem + kTsOffEnc);
#pragma unroll 1
      for (int rr = 0; rr < 4; ++rr) {
        const int r = rr * 32 + lane;
        const int64_t pt = tile * kTileM + r;
        float x[3] = {0.f, 0.f, 0.f};
        if (pt < M) {
          if (p.x) {
            x[0] = __ldg(p.x + 3 * pt); x[1] = __ldg(p.x + 3 * pt + 1); x[2] = __ldg(p.x + 3 * pt + 2);
          } else {
            const int64_t ray = __ldg(p.ray_indices + pt);
            const float ts = __ldg(p.t_starts + pt), te = __ldg(p.t_ends + pt);
            const float zm = __fdiv_rn(__fadd_rn(ts, te), 2.0f);                              // eonerf.py:206
            const float* o = p.origins + ray * p.o_stride;
            const float* dd = p.viewdirs + ray * p.d_stride;
#pragma unroll
            for (int k = 0; k < 3; ++k) x[k] = __fadd_rn(__ldg(o + k), __fmul_rn(__ldg(dd + k), zm));   // eonerf.py:207
            if (p.z_mid) p.z_mid[pt] = zm;
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
      if (lane == 0) { if (rank != 0) mbar_arrive_remote(B.enc_full, 0); else mbar_arrive(B.enc_full); }
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ===== epilogue warps =====
    const int q = warp & 3;                         // TMEM lane quarter
    const int g = (warp - 4) >> 2;                  // column group: 64 of the 128 accumulator columns of a half
    const int r = q * 32 + lane;
    const int e = threadIdx.x - 128;
    const uint32_t s_cst = smem_u32(cst);
    const uint32_t s_part = smem_u32(part) + (uint32_t)r * 16u;
    const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
    uint32_t cph = 0;
    int te = 0; (void)te;

    // bias of stage `dn` into accumulator half h, this thread's 64 columns (+ the per-image row of the HD0 stage's transient half)
    auto prefill = [&](const TsStageEpi& dn, const int h, const float* delta) {
      const int oc0 = dn.wide ? (g == 0 ? h * 64 : 128 + h * 64) : g * 64;      // output column of this thread's first accumulator column
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        uint32_t b[32];
        const uint32_t sb = s_cst + (uint32_t)(dn.bias_off + oc0 + c * 32) * 4u;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float b0, b1, b2, b3;
          lds_f4(sb + j * 16, b0, b1, b2, b3);
          b[4 * j] = __float_as_uint(b0); b[4 * j + 1] = __float_as_uint(b1); b[4 * j + 2] = __float_as_uint(b2); b[4 * j + 3] = __float_as_uint(b3);
        }
        if (delta && oc0 >= kHid) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 t = __ldg((const float4*)(delta + oc0 - kHid + c * 32) + j);
            b[4 * j] = __float_as_uint(__uint_as_float(b[4 * j]) + t.x); b[4 * j + 1] = __float_as_uint(__uint_as_float(b[4 * j + 1]) + t.y);
            b[4 * j + 2] = __float_as_uint(__uint_as_float(b[4 * j + 2]) + t.z); b[4 * j + 3] = __float_as_uint(__uint_as_float(b[4 * j + 3]) + t.w);
          }
        }
        tmem_st32(tlane + kTmAcc + h * 128 + g * 64 + c * 32, b);
      }
    };

    // first tile: layer 0's bias into both accumulator halves
    prefill(c_ts_epi[0], 0, nullptr);
    prefill(c_ts_epi[0], 1, nullptr);
    tmem_st_wait();
    tc_fence_before();
    named_bar_sync(1, kTsEpiThreads);
    if (e == 0 && it0 < n_items) {
      for (int h = 0; h < 2; ++h) { if (rank != 0) mbar_arrive_remote(&B.half_ready[h], 0); else mbar_arrive(&B.half_ready[h]); }
    }

    for (int64_t it = it0; it < n_items && !(p.dbg & 32); it += it_stride) {
      const bool next_item = it + it_stride < n_items;
      const int64_t pt = (2 * it + rank) * kTileM + r;
      const bool valid = pt < M;
      uint32_t cls = 0;
      if (p.img_idx && valid)
        cls = (uint32_t)(p.ray_indices ? __ldg(p.img_idx + __ldg(p.ray_indices + pt) * p.img_stride) : __ldg(p.img_idx + pt * p.img_stride));
      for (int s = 0; s < p.n_stages; ++s) {
        const TsStageEpi d = c_ts_epi[s];
        const bool last_stage = s + 1 == p.n_stages;
        const bool has_next = !last_stage || next_item;
        const TsStageEpi dn = c_ts_epi[last_stage ? 0 : s + 1];
        const float* delta_next = (has_next && dn.kind == 2) ? p.class_delta + (size_t)cls * kHid : nullptr;
        const uint32_t a_out = tlane + ((s & 1) ? kTmA1 : kTmA0);            // stage s writes A[s & 1]
        const int n_halves = d.wide ? 2 : 1;
        float h0 = 0.f, h1 = 0.f, h2 = 0.f;
        for (int h = 0; h < n_halves; ++h) {
          mbar_wait(&B.acc_full[h], (cph >> h) & 1u);
          cph ^= 1u << h;
          tc_fence_after();
          EO_TRACE(1, te, e == 0);
          const int oc0 = d.wide ? (g == 0 ? h * 64 : 128 + h * 64) : g * 64;   // first output column of this thread's 64
          uint32_t pk[32];
          if (p.dbg & 1) {
#pragma unroll
            for (int j = 0; j < 32; ++j) pk[j] = 0x3c003c00u;
          } else
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t v[32];
            tmem_ld32(tlane + kTmAcc + h * 128 + g * 64 + c * 32, v);
            tmem_ld_wait_dep(v);
            if (d.relu) {
#pragma unroll
              for (int j = 0; j < 16; ++j) pk[c * 16 + j] = pack_bf16_relu(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) pk[c * 16 + j] = pack_bf16(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
            }
          }
          // narrow heads on the rounded activations
          if (d.kind == 1 || d.kind == 3 || (d.kind == 2 && g == 0)) {
            const int wbase = d.kind == 1 ? kCWSigma : (d.kind == 2 ? kCWAlb : kCWTs);
            const uint32_t sw = s_cst + (uint32_t)(wbase + oc0) * 4u;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              float w0, w1, w2, w3;
              const float a0 = bf_lo(pk[2 * j]), a1 = bf_hi(pk[2 * j]), a2 = bf_lo(pk[2 * j + 1]), a3 = bf_hi(pk[2 * j + 1]);
              lds_f4(sw + j * 16, w0, w1, w2, w3);
              h0 = fmaf(a0, w0, h0); h0 = fmaf(a1, w1, h0); h0 = fmaf(a2, w2, h0); h0 = fmaf(a3, w3, h0);
              if (d.kind == 2) {
                lds_f4(sw + 512 + j * 16, w0, w1, w2, w3);
                h1 = fmaf(a0, w0, h1); h1 = fmaf(a1, w1, h1); h1 = fmaf(a2, w2, h1); h1 = fmaf(a3, w3, h1);
                lds_f4(sw + 1024 + j * 16, w0, w1, w2, w3);
                h2 = fmaf(a0, w0, h2); h2 = fmaf(a1, w1, h2); h2 = fmaf(a2, w2, h2); h2 = fmaf(a3, w3, h2);
              } else if (d.kind == 3) {
                lds_f4(sw + (kCWTb - kCWTs) * 4 + j * 16, w0, w1, w2, w3);
                h1 = fmaf(a0, w0, h1); h1 = fmaf(a1, w1, h1); h1 = fmaf(a2, w2, h1); h1 = fmaf(a3, w3, h1);
              }
            }
          }
          // the next layer's A operand: K block oc0 / 64 of A_out, 32 TMEM columns (bf16 pairs)
          if (!last_stage && !(p.dbg & 1)) tmem_st32(a_out + (oc0 >> 1), pk);
          // this half's accumulator columns are drained: refill them with the next stage's bias
          if (has_next && (dn.wide || h == 0) && !(p.dbg & 2)) prefill(dn, h, delta_next);
          if (!d.wide && has_next && dn.wide && !(p.dbg & 2)) prefill(dn, 1, delta_next);        // 128-wide last stage: half 1 has been idle since HD0
          if (h == n_halves - 1) {
            // partial head sums of the two column groups -> shared memory; combined after the barrier below
            if (d.kind == 1) {
              asm volatile("st.shared.f32 [%0], %1;" ::"r"(s_part + g * 4), "f"(h0) : "memory");
            } else if (d.kind == 2 && g == 0) {
              asm volatile("st.shared.f32 [%0], %1;" ::"r"(s_part), "f"(h0) : "memory");
              asm volatile("st.shared.f32 [%0], %1;" ::"r"(s_part + 4), "f"(h1) : "memory");
              asm volatile("st.shared.f32 [%0], %1;" ::"r"(s_part + 8), "f"(h2) : "memory");
            } else if (d.kind == 3) {
              asm volatile("st.shared.f32 [%0], %1;" ::"r"(s_part + g * 8), "f"(h0) : "memory");
              asm volatile("st.shared.f32 [%0], %1;" ::"r"(s_part + g * 8 + 4), "f"(h1) : "memory");
            }
          }
          // hand the half over: its A-operand K blocks are written and its accumulator columns hold the next bias
          EO_TRACE(1, te, e == 0);
          tmem_st_wait();
          tc_fence_before();
          named_bar_sync(1, kTsEpiThreads);
          EO_TRACE(1, te, e == 0);
          if (e == 0 && has_next) {
            if (rank != 0) mbar_arrive_remote(&B.half_ready[h], 0); else mbar_arrive(&B.half_ready[h]);
            if (!d.wide) { if (rank != 0) mbar_arrive_remote(&B.half_ready[1], 0); else mbar_arrive(&B.half_ready[1]); }
          }
        }
        if (valid && g == 0 && d.kind != 0) {
          float q0, q1, q2, q3;
          lds_f4(s_part, q0, q1, q2, q3);
          const float* sc = cst + kCScalars;
          if (d.kind == 1) p.sigma[pt] = softplus_f(q0 + q1 + sc[0]);                // eonerf.py:106,145
          else if (d.kind == 2) {
            p.rgb[3 * pt + 0] = sigmoid_f(q0 + sc[1]);
            p.rgb[3 * pt + 1] = sigmoid_f(q1 + sc[2]);
            p.rgb[3 * pt + 2] = sigmoid_f(q2 + sc[3]);
          } else {
            p.ts[pt] = sigmoid_f(q0 + q2 + sc[4]);
            p.tb[pt] = softplus_f(q1 + q3 + sc[5]);
          }
        }
        // (the partial sums are re-written two or more stages later, behind several of the barriers above)
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, 512);
  }
}

// issue order of one stage (see the header): which accumulator half, which weight block, which A block, what to wait for / commit
static TsProgram ts_program(int n_stages) {
  static const int8_t nkb[kFwdStages] = {1, 4, 4, 4, 4, 5, 4, 4, 4, 4, 2, 2, 2};
  static const int8_t halves[kFwdStages] = {2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 1, 1, 1};
  // A K-blocks read by the 128-wide stages (weight block kb -> A block): T1 reads HD0[:, 128:256], T2 / T3 read the 128 columns before
  static const int8_t narrow_a[3][2] = {{2, 3}, {0, 1}, {0, 1}};
  TsProgram P{};
  P.n = n_stages;
  int off = 0;
  for (int s = 0; s < kFwdStages; ++s) {
    if (s < n_stages) {
      TsStage& S = P.s[s];
      S.wide = halves[s] == 2; S.nkb = nkb[s]; S.blk_off = off;
      int n = 0;
      auto add = [&](int half, int kb0, int a0, int kb1, int a1, int wait, int commit) {      // kb1 < 0: a single sub-block
        TsStep& t = S.st[n++];
        t.half = (int8_t)half; t.n = kb1 < 0 ? 1 : 2;
        t.kb[0] = (int8_t)kb0; t.a_kb[0] = (int8_t)a0; t.kb[1] = (int8_t)kb1; t.a_kb[1] = (int8_t)a1;
        t.wait = (int8_t)wait; t.commit = (int8_t)commit;
      };
      if (s == 0) {                       // K = 64: the encoding only
        add(0, 0, -1, -1, 0, 1 | 4, 1);
        add(1, 0, -1, -1, 0, 2, 2);
      } else if (S.wide) {
        const bool enc = s == 5;          // skip connection: weight block 4 multiplies the encoding
        int w0 = 1;
        if (enc) { add(0, 4, -1, -1, 0, w0, 0); w0 = 0; }
        add(0, 0, 0, 2, 2, w0, 0);
        add(0, 1, 1, 3, 3, 2, 1);
        if (enc) add(1, 4, -1, -1, 0, 0, 0);
        add(1, 0, 0, 1, 1, 0, 0);
        add(1, 2, 2, 3, 3, 0, 2 | (enc ? 4 : 0));
      } else {
        add(0, 0, narrow_a[s - 10][0], 1, narrow_a[s - 10][1], 1 | 2, 1);
      }
      S.n_steps = (int8_t)n;
    }
    off += halves[s] * nkb[s];
  }
  return P;
}

}  // namespace

// EONERF_FUSED_TS=1: inference calls of the EO-NeRF field take this kernel.  Off by default: measured slower than the
// shared-memory-operand kernel (profiles/r2c_ts_experiment.log): one tile per CTA doubles the weight stream per
// FLOP and leaves only a quarter of a layer of independent MMA work to hide the accumulator -> epilogue -> issuer hand-over.
bool fused_ts_enabled() {
  static const int on = [] { const char* e = getenv("EONERF_FUSED_TS"); return e ? atoi(e) : 0; }();
  return on != 0;
}

int fused_field_fwd_ts(const EonerfFieldFwdArgs* a, cudaStream_t s) {
  const int64_t N = a->n_pts;
  const EonerfFieldParams* prm = a->params;
  const PrepLayout W = prep_layout(a->field, EONERF_PREC_BF16, prm->n_images);
  const FusedPrepLayout F = fused_prep_layout(prm->n_images);
  const uint8_t* ext = (const uint8_t*)a->prepared + W.total;
  FusedTsParams p{};
  p.M = N; p.M_dev = a->n_pts_dev; p.n_tiles = (N + kTileM - 1) / kTileM;
  p.n_stages = a->density_only ? 8 : kFwdStages;
  if (const char* dbg = getenv("EONERF_FUSED_DBG")) p.dbg = atoi(dbg);
  p.x = a->x;
  p.origins = a->origins; p.o_stride = a->origins_stride; p.viewdirs = a->viewdirs; p.d_stride = a->viewdirs_stride;
  p.ray_indices = a->ray_indices; p.t_starts = a->t_starts; p.t_ends = a->t_ends; p.z_mid = a->z_mid;
  p.img_idx = a->density_only ? nullptr : a->img_idx; p.img_stride = a->img_idx_stride;
  p.consts = (const float*)(ext + F.consts);
  p.class_delta = (const float*)(ext + F.delta);
  p.prog = ts_program(p.n_stages);
  p.sigma = a->sigma; p.rgb = a->rgb; p.ts = a->transient_s; p.tb = a->transient_beta;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t items = (p.n_tiles + 1) / 2;
  const int n_ctas = (int)((items < sms / 2 ? items : sms / 2) * 2);
  CUtensorMap wmap;
  int rc = make_blob_map(&wmap, ext + F.fblob, kFwdBlocks);
  if (rc != EONERF_OK) return rc;
  static PerDeviceOnce once;
  if (once()) {
    EO_CUDA(cudaFuncSetAttribute(fused_fwd_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTs));
  }
  profile_begin(3, (double)N * (a->density_only ? 982528.0 : 1345280.0), 0.0, s);
  rc = launch_fused(fused_fwd_ts_kernel, 2, n_ctas, p, wmap, s, kSmemTs, kTsThreads);
  profile_end(s);
  if (rc != EONERF_OK) return rc;
  EO_LAUNCH_CHECK();
  return EONERF_OK;
}

}  // namespace eonerf

#ifdef EONERF_TIMING
extern "C" int eonerf_debug_trace_ts(long long* out) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out, eonerf::g_fused_trace, sizeof(long long) * 2048);
  return 0;
}
#endif
