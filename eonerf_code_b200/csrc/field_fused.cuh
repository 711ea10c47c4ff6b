// Layouts and launch parameters of the fused radiance-field kernels (field_fused.cu, gemm_tn_blocked in gemm_tc.cu).
//
// Tile-blocked activation layout: an activation / gradient matrix with C columns (a multiple of 64) is stored as
// 16 KB blocks [128 samples x 64 features] of bf16, block (tile t, column block kb) at ((t * C/64) + kb) * 16384.
// Inside a block, sample row r and 16-byte chunk c (8 features) live at r*128 + ((c ^ (r & 7)) << 4): exactly the
// 128-byte-swizzled K-major shared-memory image the tensor core reads, so a block moves between HBM and shared memory
// with one cp.async.bulk and no tensor map.  Read the other way round the same image is the MN-major operand of the
// parameter-gradient GEMM (contraction over the samples).
#pragma once
#include <cuda.h>

#include "field_layout.cuh"

namespace eonerf {

constexpr int kTileM = 128;
constexpr int kBlkBytes = 16384;

// stash arrays written by the fused forward (index = producing stage; 13 = positional encoding)
enum { kArrH0 = 0, kArrBott = 8, kArrHd0 = 9, kArrT1 = 10, kArrEnc = 13, kNumArr = 14 };
// ReLU sign-bit arrays: [Mpad][8] uint32 (256 bits per sample; 128-wide stages use the first 4 words)
enum { kMaskH0 = 0, kMaskHd0 = 8, kMaskT1 = 9, kNumMask = 12 };

static inline int arr_blocks(int arr) { return arr == kArrEnc ? 1 : (arr >= kArrT1 && arr < kArrEnc ? 2 : 4); }

struct FusedStashLayout {
  int64_t n_tiles, mpad;
  int64_t xf, cls, arr[kNumArr], mask[kNumMask], total;
};

static inline FusedStashLayout fused_stash_layout(int64_t n, int density_only) {
  FusedStashLayout L{};
  L.n_tiles = (n + kTileM - 1) / kTileM;
  L.mpad = L.n_tiles * kTileM;
  int64_t off = 0;
  auto take = [&](int64_t bytes) { int64_t o = off; off = align_up(off + bytes, 1024); return o; };
  L.xf = take(L.mpad * 3 * 4);
  L.cls = take(L.mpad * 4);
  for (int a = 0; a < kNumArr; ++a) {
    bool used = a < 8 || a == kArrEnc || !density_only;
    L.arr[a] = used ? take(L.n_tiles * arr_blocks(a) * kBlkBytes) : -1;
  }
  for (int m = 0; m < kNumMask; ++m) {
    bool used = m < 8 || !density_only;
    L.mask[m] = used ? take(L.mpad * 32) : -1;
  }
  L.total = off;
  return L;
}

// backward scratch: gradient arrays wrt the pre-activations, same blocked shapes as the stash arrays (index = the
// forward stage whose pre-activation it is), per-sample head pre-activation gradients, per-image bias gradients
struct FusedScratchLayout {
  int64_t n_tiles, mpad;
  int64_t g[13], dpre, dcb, total;
};

static inline FusedScratchLayout fused_scratch_layout(int64_t n, int64_t n_images, int density_only) {
  FusedScratchLayout L{};
  L.n_tiles = (n + kTileM - 1) / kTileM;
  L.mpad = L.n_tiles * kTileM;
  int64_t off = 0;
  auto take = [&](int64_t bytes) { int64_t o = off; off = align_up(off + bytes, 1024); return o; };
  for (int a = 0; a < 13; ++a) {
    bool used = a < 8 || !density_only;
    L.g[a] = used ? take(L.n_tiles * arr_blocks(a) * kBlkBytes) : -1;
  }
  L.dpre = take(L.mpad * 8 * 4);
  L.dcb = take((n_images > 0 ? n_images : 1) * kHid * 4);
  L.total = off;
  return L;
}

// constants block (fp32) staged in shared memory by the fused kernels
constexpr int kCBiasTrunk = 0;        // 8 x 256
constexpr int kCBiasBott = 2048;      // 256
constexpr int kCBiasTr = 2304;        // 3 x 128  (transient_mlp.{1,2,3}.bias)
constexpr int kCWSigma = 2688;        // 256
constexpr int kCWAlb = 2944;          // 3 x 128
constexpr int kCWTs = 3328;           // 128
constexpr int kCWTb = 3456;           // 128
constexpr int kCScalars = 3584;       // b_sigma, b_alb[3], b_ts, b_tb
constexpr int kCBiasHd0 = 3600;       // 256: [albedo_mlp.0.bias | transient_mlp.0.bias]  (the per-image part is `delta`)
constexpr int kCFloats = 3856;

// weight-block programs: forward 13 stages, backward 15 stages (see field_fused.cu)
constexpr int kFwdStages = 13;
constexpr int kFwdBlocks = 2 + 4 * 8 + 10 + 2 * 8 + 8 + 8 + 3 * 2;   // 82
constexpr int kBwdStages = 14;
constexpr int kBwdBlocks = 3 * 2 + 8 + 8 + 8 + 8 + 4 + 8 + 4 * 8 + 4;  // 86

struct FusedPrepLayout {
  int64_t fblob, bblob, consts, delta, total;   // relative to the start of the fused extras
};
// delta: fp32 [n_img,128] = W_t0[:,256:260] . emb[img]  (eonerf.py:165-167: the transient embedding enters as a bias row)
static inline FusedPrepLayout fused_prep_layout(int64_t n_images) {
  FusedPrepLayout L{};
  int64_t off = 0;
  auto take = [&](int64_t bytes) { int64_t o = off; off = align_up(off + bytes, 1024); return o; };
  L.fblob = take((int64_t)kFwdBlocks * kBlkBytes);
  L.bblob = take((int64_t)kBwdBlocks * kBlkBytes);
  L.consts = take(kCFloats * 4);
  L.delta = take((n_images > 0 ? n_images : 1) * kHid * 4);
  L.total = off;
  return L;
}

// D[n,k] (+)= sum_m G[m,n] X[m,k] over blocked operands (gemm_tc.cu); db[n] += sum_m G[m,n]
struct GemmTNBlocked {
  const uint8_t* G = nullptr; int g_nb = 4; int g_blk0 = 0; int mt_count = 2;   // G features: mt_count*128 from block g_blk0
  const uint8_t* X = nullptr; int x_nb = 4; int x_blk0 = 0; int x_cnt = 4;      // X features: x_cnt*64 from block x_blk0
  int64_t n_tiles = 0;
  const int64_t* n_pts_dev = nullptr;   // live sample count on the device (n_tiles is then the capacity)
  int n_valid[2] = {128, 128};        // valid output rows per 128-row block of G features
  int k_valid = 256;                  // valid X features
  float* D[2] = {nullptr, nullptr}; int64_t ldd[2] = {0, 0};
  float* db[2] = {nullptr, nullptr};
};
int gemm_tn_blocked(const GemmTNBlocked& g, cudaStream_t s);
// all GEMMs of `list` in one persistent launch (at most 16 per launch; longer lists are cut)
int gemm_tn_blocked_group(const GemmTNBlocked* list, int n, cudaStream_t s);
int fused_cta_group();
int make_blob_map(CUtensorMap* map, const void* base, int64_t n_blocks);
bool fused_ts_enabled();                                               // field_fused_ts.cu
int fused_field_fwd_ts(const EonerfFieldFwdArgs* a, cudaStream_t s);

}  // namespace eonerf
