// Device-side helpers shared by the fused forward (field_fused.cu) and backward (field_fused_bwd.cu) kernels:
// the shared-memory map, explicit shared-space accesses, bf16 packing, reduced-range sine / cosine.
#pragma once
#include "field_fused.cuh"
#include "tc_ptx.cuh"

namespace eonerf {

// EONERF_WEIGHT_HINT=1: weight-block loads carry an L2 evict_last policy.  Measured: no effect (A/B on one box), so off; the
// evict_first policy on the stash STORES (EONERF_STORE_HINT in field_fused*.cu) is what helps (+3.7 % on the camera forward).
#ifndef EONERF_WEIGHT_HINT
#define EONERF_WEIGHT_HINT 0
#endif

#ifndef EONERF_RING_STAGES
#define EONERF_RING_STAGES 3
#endif
constexpr int kRingStages = EONERF_RING_STAGES;   // weight ring depth (2 only for the latency experiment)
constexpr int kSlotBytes = 5 * kBlkBytes;                  // ACT blocks 0..3 + ENC block 4
constexpr int kOffRing = 0;
constexpr int kOffSlot = kRingStages * kBlkBytes;          // 49152
constexpr int kOffConst = kOffSlot + 2 * kSlotBytes;       // 212992
// Ring depth R: R weight blocks, then two activation slots of 5 blocks (R = 3: ACT 0..3 + ENC/DENC) or of 4 blocks (R = 5: the
// backward chain without position gradients needs no DENC block, so its weight ring gets the 32 KB).  Same total either way.
__host__ __device__ constexpr int off_slot(int ring) { return ring * kBlkBytes; }
__host__ __device__ constexpr int slot_bytes(int ring) { return (ring > 3 ? 4 : 5) * kBlkBytes; }
static_assert(off_slot(5) + 2 * slot_bytes(5) == kOffConst && off_slot(3) + 2 * slot_bytes(3) == kOffConst, "layouts must end at the constants");
constexpr int kConstBytes = 15872;
constexpr int kOffPart = kOffConst + kConstBytes;          // [2 slots][128][2] floats
constexpr int kOffBar = kOffPart + 2048;
constexpr int kSmemFused = kOffBar + 128 + 1024;           // + alignment slack
// Forward kernel: EONERF_FWD_RING = 4 buys a fourth weight-ring stage with the 15 KB constants block (biases, head weights), which
// is then read through L1 from global memory instead, and with the 1 KB alignment slack (the dynamic shared-memory array is
// declared 1024-byte aligned).  3: the constants are staged in shared memory as in the backward kernel.
#ifndef EONERF_FWD_RING
#define EONERF_FWD_RING 3
#endif
constexpr int kFwdRing = EONERF_FWD_RING;
constexpr bool kFwdConstG = kFwdRing > 3;                  // constants from global memory
constexpr int kFOffSlot = kFwdRing * kBlkBytes;
constexpr int kFSlotBytes = 5 * kBlkBytes;
constexpr int kFOffConst = kFOffSlot + 2 * kFSlotBytes;
constexpr int kFOffPart = kFOffConst + (kFwdConstG ? 0 : kConstBytes);
constexpr int kFOffBar = kFOffPart + 2048;
constexpr int kSmemFwd = kFOffBar + 128 + (kFwdConstG ? 896 : 1024);   // + alignment slack (896: all that is left under 227 KB)
static_assert(kSmemFwd <= 232448, "forward shared memory budget");
// EONERF_DUTY_WARP = 1: an 11th warp joins every end-of-layer named barrier and does the after-barrier duties (signal the MMA
// issuer, issue the bulk stash stores, wait for their shared-memory reads), so no epilogue warp has anything to do between
// the barrier and the next accumulator.  0: two epilogue threads (kSignalThread, the store threads) do them.
#ifndef EONERF_DUTY_WARP
#define EONERF_DUTY_WARP 1
#endif
constexpr bool kDutyWarp = EONERF_DUTY_WARP != 0;
constexpr int kDutyWarpId = 10;
#ifndef EONERF_DUTY_STORES
#define EONERF_DUTY_STORES 0
#endif
constexpr bool kDutyStores = kDutyWarp && EONERF_DUTY_STORES != 0;   // the duty warp also owns the bulk stores (measured slower)
constexpr int kFusedThreads = kDutyWarp ? 352 : 320;
// Forward kernel only: a 12th warp computes the positional encoding of the NEXT work item while the current one is in its
// layers (the ENC block of a slot is free once layer 5's MMAs have read it), so consecutive items form one uninterrupted
// pipeline: no per-item prologue during which the tensor core has nothing to do.  12 warps x 168 registers = 64 512 <= 65 536.
constexpr int kEncWarpId = kDutyWarp ? 11 : 10;
constexpr int kFwdThreads = kFusedThreads + 32;
constexpr int kEpiThreads = 256;
constexpr int kBarThreads = kDutyWarp ? 288 : 256;      // participants of the end-of-layer named barrier
constexpr int kSignalThread = 128;    // epilogue thread (warp 6) that signals act_ready after the end-of-layer barrier
// The bulk stash stores of an epilogue (up to four 16 KB blocks) are issued by kStoreThreads threads of different warps, each
// tracking and waiting for its own bulk groups.  What matters is that the store thread is not the signalling thread and that
// neither sits in a warp with other after-barrier work: 1, 2 and 4 store threads measure the same (A/B on one box).
#ifndef EONERF_STORE_THREADS
#define EONERF_STORE_THREADS 1
#endif
constexpr int kStoreThreads = EONERF_STORE_THREADS;        // 1 (thread 160 stores every block), 2 or 4
__host__ __device__ constexpr int store_thread_id(int e) {
  return kStoreThreads == 4 ? ((e & 63) == 32 ? (e >> 6) : -1)           // e = 32, 96, 160, 224
         : kStoreThreads == 2 ? (e == 160 ? 0 : (e == 224 ? 1 : -1))
                              : (e == 160 ? 0 : -1);
}

static_assert(kCFloats * 4 <= kConstBytes, "constants do not fit");
static_assert(kSmemFused <= 232448, "shared memory budget");

__device__ __forceinline__ float sigmoid_f(float v) { return 1.0f / (1.0f + expf(-v)); }          // nn.Sigmoid
__device__ __forceinline__ float softplus_f(float v) { return v > 20.0f ? v : log1pf(expf(v)); }   // nn.Softplus(beta=1, threshold=20)

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
// {hi, lo} -> packed bf16 pair with ReLU folded into the conversion (one F2FP instead of 2 FMNMX + F2FP)
__device__ __forceinline__ uint32_t pack_bf16_relu(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
__device__ __forceinline__ float bf_lo(uint32_t p) { return __uint_as_float(p << 16); }
__device__ __forceinline__ float bf_hi(uint32_t p) { return __uint_as_float(p & 0xFFFF0000u); }

// shared-memory byte offset of (row, 16-byte chunk) inside a [128 x 64] bf16 block
__device__ __forceinline__ uint32_t blk_off(int row, int chunk) { return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4)); }

// explicit shared-space accesses (the 1024-byte alignment arithmetic on the dynamic shared-memory base hides the address
// space from the compiler, which would otherwise emit generic LD/ST)
__device__ __forceinline__ void lds_f4(uint32_t a, float& x, float& y, float& z, float& w) {
  // volatile: keeps the load where it is written (hoisted above the accumulator wait, 128 bias values would spill)
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x), "=f"(y), "=f"(z), "=f"(w) : "r"(a) : "memory");
}
__device__ __forceinline__ void lds_f2(uint32_t a, float& x, float& y) {
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(x), "=f"(y) : "r"(a) : "memory");
}
__device__ __forceinline__ void sts_f2(uint32_t a, float x, float y) {
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(a), "f"(x), "f"(y) : "memory");
}
__device__ __forceinline__ void sts_u4(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}

// four consecutive constants: from the shared-memory copy (addr = shared address) or, when the forward keeps no copy, through L1
// from the prepared blob (addr = byte offset into it)
template <bool kGlobal>
__device__ __forceinline__ void cf4(const float* gbase, uint32_t addr, float& x, float& y, float& z, float& w) {
  if (kGlobal) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const char*>(gbase) + addr));
    x = t.x; y = t.y; z = t.z; w = t.w;
  } else {
    lds_f4(addr, x, y, z, w);
  }
}

// sin(a) for |a| < ~1e3: two-term Cody-Waite reduction to [-pi, pi] then MUFU.SIN.  Absolute error < 6e-7, far below
// the bf16 resolution the value is rounded to (2^-9 relative), at ~7 instructions instead of ~45 for sinf().
__device__ __forceinline__ float sin_reduced(float a) {
  const float k = rintf(a * 0.15915494309189535f);
  float rr = fmaf(k, -6.2831854820251465f, a);
  rr = fmaf(k, 1.7484555314695172e-7f, rr);
  return __sinf(rr);
}

// positional-encoding column kBase+c (c compile-time after unrolling) of position x
// (mlp.py:199-205: [x, sin(2^k x) k=0..9 (frequency-major), sin(2^k x + pi/2)], column 63 is padding)
template <int kBase>
__device__ __forceinline__ float posenc_col(const float (&x)[3], int c) {
  c += kBase;
  if (c < 3) return x[c];
  if (c >= 63) return 0.f;
  int e = c - 3;
  const int hf = e >= 30;
  e -= hf * 30;
  const float xb = x[e % 3] * (float)(1 << (e / 3));
  return sin_reduced(hf ? __fadd_rn(xb, kHalfPi) : xb);                  // torch adds the scalar in fp32 (mlp.py:203)
}


// cos(a) with the same reduction
__device__ __forceinline__ float cos_reduced(float a) {
  const float k = rintf(a * 0.15915494309189535f);
  float rr = fmaf(k, -6.2831854820251465f, a);
  rr = fmaf(k, 1.7484555314695172e-7f, rr);
  return __cosf(rr);
}

static inline int fused_grid(int64_t n_pairs) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return (int)(n_pairs < sms ? n_pairs : sms);
}

// Optional wait-time instrumentation (build with -DEONERF_TIMING): cycles CTA 0's role threads spend blocked on each
// barrier class.  g_fused_timing: 0 mma<-act_ready, 1 mma<-weights, 2 producer<-ring slot, 3 epilogue<-accumulator,
// 4 epilogue start barrier (incl. stash-store drain), 5 epilogue end barrier, 6 total cycles of the MMA thread
#ifdef EONERF_TIMING
static __device__ unsigned long long g_fused_timing[16];   // one copy per translation unit
// raw clock64 trace of CTA 0: [0][..] MMA thread (per slot-stage: slot handed over, MMAs issued), [1][..] epilogue thread 64 (per
// slot-stage: accumulator seen, chunks done, end barrier passed)
static __device__ long long g_fused_trace[2][1024];
// per-CTA wall time (globaltimer, ns): [0][cta] at entry, [1][cta] at exit
static __device__ unsigned long long g_fused_cta_time[2][256];
__device__ __forceinline__ unsigned long long eo_gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define EO_CTA_TIME(i) do { if (threadIdx.x == 0 && blockIdx.x < 256) g_fused_cta_time[i][blockIdx.x] = eo_gtimer(); } while (0)
#define EO_TRACE(role, idx, cond) do { if (blockIdx.x == 0 && (cond) && (idx) < 1024) g_fused_trace[role][(idx)++] = clock64(); } while (0)
#define EO_T0() const long long _t0 = clock64()
#define EO_TN(name) const long long name = clock64()
#define EO_TD(slot, a, b) do { if (blockIdx.x == 0 && threadIdx.x == 64) g_fused_timing[slot] += (unsigned long long)((b) - (a)); } while (0)
#define EO_T1(slot) do { if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) g_fused_timing[slot] += (unsigned long long)(clock64() - _t0); } while (0)
#else
#define EO_T0() do {} while (0)
#define EO_TN(name) do {} while (0)
#define EO_TD(slot, a, b) do {} while (0)
#define EO_T1(slot) do {} while (0)
#define EO_TRACE(role, idx, cond) do {} while (0)
#define EO_CTA_TIME(i) do {} while (0)
#endif

// ---- the tensor-core side of both fused kernels -------------------------------------------------------------------
// A "program" is a list of GEMM stages; per stage the A operand is a list of 16 KB activation blocks of the slot
// (0..3 = ACT, 4 = ENC) and the B operand a run of weight blocks in the prepared blob ([output half][k block]).
struct StageMma { int8_t halves, nkb, a[5], pad; int32_t blk_off; };
struct MmaProgram { int32_t n; StageMma st[14]; };

struct FusedBars { uint64_t* w_full; uint64_t* w_empty; uint64_t* acc_full; uint64_t* act_ready; };

// kCG = 1: every CTA is on its own.  kCG = 2: CTA pair; work item `it` covers 4 tiles (2 per CTA), the leader (rank 0)
// issues cta_group::2 MMAs over both CTAs' tiles, each CTA streams only its half of every weight block.
// kMC > 1 (with kCG = 1): weight multicast.  The CTAs of a cluster of kMC run independently (own tiles, own MMAs) but share
// the weight stream: each CTA loads 1/kMC of every weight block and multicasts it, so L2 is read once per cluster.  A ring
// slot is refilled only when ALL CTAs of the cluster have consumed it (every CTA's tcgen05.commit arrives on every
// CTA's w_empty), which keeps the rings in lockstep.
template <int kCG, int kMC = 1, int kRing = kRingStages>
__device__ __forceinline__ void fused_setup(uint8_t* smem, FusedBars& B, uint32_t*& tmem_base_s, uint32_t rank, int off_bar = kOffBar) {
  static_assert((2 * kRing + 4) * 8 + 4 <= 128, "barrier region");
  B.w_full = (uint64_t*)(smem + off_bar);
  B.w_empty = B.w_full + kRing;
  B.acc_full = B.w_empty + kRing;
  B.act_ready = B.acc_full + 2;
  tmem_base_s = (uint32_t*)(B.act_ready + 2);
  if (threadIdx.x == 0) {
    const uint32_t n_arr = (kCG == 2 && rank == 0) ? 2 : 1;   // leader of a pair: + the peer epilogue's remote arrival
    for (int s = 0; s < kRing; ++s) { mbar_init(&B.w_full[s], 1); mbar_init(&B.w_empty[s], kMC); }
    for (int s = 0; s < 2; ++s) { mbar_init(&B.acc_full[s], 1); mbar_init(&B.act_ready[s], n_arr); }
    fence_barrier_init();
  }
  if ((threadIdx.x >> 5) == 1) {
    if (kCG == 2) tmem_alloc_2cta(tmem_base_s, 512); else tmem_alloc(tmem_base_s, 512);
  }
}

template <int kCG, int kMC = 1>
__device__ __forceinline__ void fused_teardown(uint32_t tmem_base) {
  tc_fence_before();
  if (kCG == 2 || kMC > 1) cluster_sync_all(); else __syncthreads();
  if ((threadIdx.x >> 5) == 1) {
    __syncwarp();
    tc_fence_after();
    if (kCG == 2) tmem_dealloc_2cta(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
  }
}

// one thread: stream this CTA's weight blocks through the ring
template <int kCG, int kMC = 1, int kRing = kRingStages>
__device__ __forceinline__ void fused_producer(const MmaProgram& prog, const uint8_t* wblob, const CUtensorMap* wmap, uint8_t* smem,
                                               const FusedBars& B, int64_t it0, int64_t n_items, int64_t it_stride, uint32_t rank,
                                               int no_loads = 0) {
  int rs = 0; uint32_t rph = 0;
  int filled = 0;                                    // ablation (no_loads): only the first ring fill is loaded, later blocks are "ready" at once
#if EONERF_WEIGHT_HINT
  const uint64_t pol = l2_policy_evict_last();        // the weight blob is re-read by every CTA for every tile: keep it in L2
#endif
  for (int64_t it = it0; it < n_items; it += it_stride)
    for (int s = 0; s < prog.n; ++s) {
      const StageMma d = prog.st[s];
      const uint8_t* src = wblob + (size_t)d.blk_off * kBlkBytes;
      const int nblk = kCG == 1 ? d.halves * d.nkb : d.nkb;
      const uint32_t bytes = (kCG == 2 && d.halves == 1) ? kBlkBytes / 2 : kBlkBytes;
      const size_t base = kCG == 1 ? 0 : (d.halves == 2 ? (size_t)rank * d.nkb * kBlkBytes : (size_t)rank * (kBlkBytes / 2));
      for (int rep = 0; rep < 2; ++rep)
        for (int b = 0; b < nblk; ++b) {
          { EO_T0(); mbar_wait(&B.w_empty[rs], rph ^ 1); EO_T1(2); }
          if (no_loads && filled >= kRing) {
            if (kCG != 2 || rank == 0) mbar_arrive(&B.w_full[rs]);
            if (++rs == kRing) { rs = 0; rph ^= 1; }
            continue;
          }
          ++filled;
          if (kCG == 2) {
            // pair: both CTAs' copies are counted on the leader's barrier; block (stage, half = rank, kb) or, for the
            // 128-wide stages, rows rank*64.. of block kb
            if (rank == 0) mbar_expect_tx(&B.w_full[rs], 2 * bytes);
            uint8_t* dst = smem + kOffRing + rs * kBlkBytes;
            const int row0 = (d.blk_off + (d.halves == 2 ? (int)rank * d.nkb + b : b)) * 128 + (d.halves == 2 ? 0 : (int)rank * 64);
#if EONERF_WEIGHT_HINT
            tma_load_2d_2cta_hint(dst, wmap, &B.w_full[rs], 0, row0, pol);
            if (d.halves == 2) tma_load_2d_2cta_hint(dst + kBlkBytes / 2, wmap, &B.w_full[rs], 0, row0 + 64, pol);
#else
            tma_load_2d_2cta(dst, wmap, &B.w_full[rs], 0, row0);
            if (d.halves == 2) tma_load_2d_2cta(dst + kBlkBytes / 2, wmap, &B.w_full[rs], 0, row0 + 64);
#endif
            if (++rs == kRing) { rs = 0; rph ^= 1; }
            continue;
          }
          mbar_expect_tx(&B.w_full[rs], bytes);
          if (kMC > 1) {
            constexpr uint32_t piece = kBlkBytes / kMC;          // this CTA's share of the block, broadcast to the whole cluster
            bulk_load_multicast(smem + kOffRing + rs * kBlkBytes + rank * piece, src + (size_t)b * kBlkBytes + rank * piece, piece, &B.w_full[rs],
                                (uint16_t)((1u << kMC) - 1));
          } else
#if EONERF_WEIGHT_HINT
          bulk_load_hint(smem + kOffRing + rs * kBlkBytes, src + base + (size_t)b * kBlkBytes, bytes, &B.w_full[rs], pol);
#else
          bulk_load(smem + kOffRing + rs * kBlkBytes, src + base + (size_t)b * kBlkBytes, bytes, &B.w_full[rs]);
#endif
          if (++rs == kRing) { rs = 0; rph ^= 1; }
        }
    }
}

// the whole MMA warp (kCG = 2: of the leader CTA), converged: issue the MMAs of every stage for both slots
// kAccInit: the epilogue warps have pre-loaded every accumulator (with the layer's bias): the first MMA of a stage accumulates too
// enc_full / enc_free (forward kernel): per-slot handshake with the encoder warp.  Stage 0 of an item waits until the item's
// positional encoding is in the slot's ENC block; the MMAs of stage kEncLastStage (layer 5, the skip connection) are the last
// readers of it, and their completion (tcgen05.commit) hands the block back for the next item's encoding.
constexpr int kEncLastStage = 5;
template <int kCG, int kMC = 1, bool kAccInit = false, int kRing = kRingStages, int kSlotBlk = (kRing > 3 ? 4 : 5)>
__device__ __forceinline__ void fused_mma_issuer(const MmaProgram& prog, uint8_t* smem, const FusedBars& B, uint32_t tmem_base, int64_t it0,
                                                 int64_t n_items, int64_t it_stride, uint64_t* enc_full = nullptr, uint64_t* enc_free = nullptr) {
  int rs = 0; uint32_t rph = 0;
  uint32_t aph = 0;                                  // bit `slot` = phase of act_ready[slot]
  uint32_t eph = 0;                                  // bit `slot` = phase of enc_full[slot]
  int tr_i = 0; (void)tr_i;
  const uint32_t ring0 = smem_u32(smem + kOffRing);
  const bool elected = elect_one_sync();
#ifdef EONERF_TIMING
  const long long t_begin = clock64();
#endif
  for (int64_t it = it0; it < n_items; it += it_stride)
    for (int s = 0; s < prog.n; ++s) {
      const StageMma d = prog.st[s];
      const int n_h = kCG == 1 ? d.halves : 1;                          // kCG = 2: one MMA covers both output halves
      const uint32_t idesc = kCG == 1 ? instr_desc(128, 128, 0, 0) : instr_desc(256, d.halves * 128, 0, 0);
      for (int slot = 0; slot < 2; ++slot) {
        { EO_T0(); mbar_wait(&B.act_ready[slot], (aph >> slot) & 1u); EO_T1(0); }
        aph ^= 1u << slot;
        if (enc_full && s == 0) {
          EO_T0(); mbar_wait(&enc_full[slot], (eph >> slot) & 1u); EO_T1(0);
          eph ^= 1u << slot;
        }
        tc_fence_after();
        EO_TRACE(0, tr_i, elected);
        const uint32_t slot0 = smem_u32(smem + off_slot(kRing) + slot * (kSlotBlk * kBlkBytes));
        for (int h = 0; h < n_h; ++h) {
          const uint32_t d_tmem = tmem_base + slot * 256 + h * 128;
          for (int kb = 0; kb < d.nkb; ++kb) {
            { EO_T0(); mbar_wait(&B.w_full[rs], rph); EO_T1(1); }
            tc_fence_after();
            const uint32_t la = desc_lo_k128(slot0 + d.a[kb] * kBlkBytes);
            const uint32_t lb = desc_lo_k128(ring0 + rs * kBlkBytes);
            if (elected) {
              umma_k128<kCG>(d_tmem, la, lb, idesc, kAccInit || kb != 0);           // 16 K elements = 32 bytes = +2 in the address field
              umma_k128<kCG>(d_tmem, la + 2, lb + 2, idesc, 1);
              umma_k128<kCG>(d_tmem, la + 4, lb + 4, idesc, 1);
              umma_k128<kCG>(d_tmem, la + 6, lb + 6, idesc, 1);
              if (kCG == 2) umma_commit_2cta(&B.w_empty[rs]);
              else if (kMC > 1) umma_commit_multicast(&B.w_empty[rs], (uint16_t)((1u << kMC) - 1));
              else umma_commit(&B.w_empty[rs]);
            }
            __syncwarp();
            if (++rs == kRing) { rs = 0; rph ^= 1; }
          }
        }
        if (elected) {
          if (kCG == 2) umma_commit_2cta(&B.acc_full[slot]); else umma_commit(&B.acc_full[slot]);
          if (enc_free && s == kEncLastStage) {
            if (kCG == 2) umma_commit_2cta(&enc_free[slot]); else umma_commit(&enc_free[slot]);
          }
        }
        EO_TRACE(0, tr_i, elected);
        __syncwarp();
      }
    }
#ifdef EONERF_TIMING
  if (blockIdx.x == 0 && elected) g_fused_timing[6] += (unsigned long long)(clock64() - t_begin);
#endif
}

// epilogue side: "this slot's A operand for the next stage is in shared memory and its accumulator is drained"
template <int kCG>
__device__ __forceinline__ void signal_act_ready(const FusedBars& B, int slot, uint32_t rank) {
  if (kCG == 2 && rank != 0) mbar_arrive_remote(&B.act_ready[slot], 0);
  else mbar_arrive(&B.act_ready[slot]);
}

template <class Kernel, class Params>
static int launch_fused(Kernel kernel, int cg, int n_ctas, const Params& p, const CUtensorMap& wmap, cudaStream_t s, int smem_bytes = kSmemFused,
                        int threads = kFusedThreads) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)n_ctas);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = (size_t)smem_bytes;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cg; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  EO_CUDA(cudaLaunchKernelEx(&cfg, kernel, p, wmap));
  return EONERF_OK;
}

// number of CTAs for n_tiles tiles: `cg` CTAs (cluster size) per work item of 2*cg tiles, at most one CTA per SM
static inline int fused_ctas(int64_t n_tiles, int cg) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t items = (n_tiles + 2 * cg - 1) / (2 * cg);
  const int64_t max_items = sms / cg;
  return (int)((items < max_items ? items : max_items) * cg);
}

}  // namespace eonerf
