// Memory layouts of the radiance-field MLP shared by the layer-by-layer path (field.cu) and the fused tcgen05 path
// (field_fused.cu): the forward stash, the backward scratch and the prepared-parameter blob.
#pragma once
#include "gemm.cuh"

namespace eonerf {

constexpr int kW = 256;       // trunk width (eonerf.py:73-75; --fc_units is never read)
constexpr int kEnc = 64;      // 63 pos-enc columns + 1 zero pad
constexpr int kH4E = kW + kEnc;
constexpr int kHid = 128;
constexpr int kDirEnc = 32;   // 27 view-enc columns + 5 zero pad
constexpr float kHalfPi = 1.57079637050628662109375f;   // fl32(0.5*pi): torch adds the python scalar in fp32 (mlp.py:203)

static inline int64_t align_up(int64_t v, int64_t a = 256) { return (v + a - 1) / a * a; }

struct StashLayout {
  int64_t xf, cls, h[8], bott, hd0, t[3], total;
  int ld_bott, ld_hd0;
};

static StashLayout stash_layout(int field, int precision, int64_t n, int density_only) {
  StashLayout L{};
  int64_t es = elem_size(precision), off = 0;
  auto take = [&](int64_t bytes) { int64_t o = off; off = align_up(off + bytes); return o; };
  L.xf = take(n * 3 * 4);
  L.cls = take(n * 4);
  for (int i = 0; i < 8; ++i) L.h[i] = take(n * (i == 4 ? kH4E : kW) * es);
  L.ld_bott = field == EONERF_FIELD_VANILLA ? kW + kDirEnc : kW;
  L.ld_hd0 = field == EONERF_FIELD_VANILLA ? kHid : 2 * kHid;
  if (!density_only) {
    L.bott = take(n * L.ld_bott * es);
    L.hd0 = take(n * L.ld_hd0 * es);
    if (field == EONERF_FIELD_EONERF)
      for (int i = 0; i < 3; ++i) L.t[i] = take(n * kHid * es);
  }
  L.total = off;
  return L;
}

struct ScratchLayout {
  int64_t ga, gb, gc, gat, g1, g2, dpre, dcb, total;
};

static ScratchLayout scratch_layout(int field, int precision, int64_t n, int64_t n_images) {
  ScratchLayout L{};
  int64_t es = elem_size(precision), off = 0;
  auto take = [&](int64_t bytes) { int64_t o = off; off = align_up(off + bytes); return o; };
  L.ga = take(n * kW * es);
  L.gb = take(n * kH4E * es);
  L.gc = take(n * (kW + kDirEnc) * es);
  L.gat = take(n * 2 * kHid * es);
  L.g1 = take(n * kHid * es);
  L.g2 = take(n * kHid * es);
  L.dpre = take(n * 8 * 4);
  L.dcb = take((n_images > 0 ? n_images : 1) * kHid * 4);
  L.total = off;
  return L;
}

// prepared blob: for every matrix W [out, Kp] then W^T [Kp, out] in T, then the fp32 class-bias table
struct PrepLayout {
  int64_t w[8], wt[8];       // trunk
  int64_t bott, bott_t;
  int64_t hd0, hd0_t;        // eonerf: [256,256] = [albedo_mlp.0 ; transient_mlp.0[:, :256]]   vanilla: [128,288]
  int64_t tr[3], tr_t[3];    // transient_mlp.{1,2,3}
  int64_t class_bias;        // fp32 [n_img,256] = [ albedo_mlp.0.bias | transient_mlp.0.bias + W[:,256:260] emb[img] ]
  int64_t total;
};

static int trunk_kp(int i) { return i == 0 ? kEnc : (i == 5 ? kH4E : kW); }
static int trunk_k(int i) { return i == 0 ? 63 : (i == 5 ? 319 : kW); }

static PrepLayout prep_layout(int field, int precision, int64_t n_images) {
  PrepLayout L{};
  int64_t es = elem_size(precision), off = 0;
  auto take = [&](int64_t bytes) { int64_t o = off; off = align_up(off + bytes); return o; };
  for (int i = 0; i < 8; ++i) { L.w[i] = take(kW * trunk_kp(i) * es); L.wt[i] = take(kW * trunk_kp(i) * es); }
  L.bott = take(kW * kW * es); L.bott_t = take(kW * kW * es);
  int64_t hd0 = field == EONERF_FIELD_VANILLA ? kHid * (kW + kDirEnc) : 2 * kHid * kW;
  L.hd0 = take(hd0 * es); L.hd0_t = take(hd0 * es);
  for (int i = 0; i < 3; ++i) { L.tr[i] = take(kHid * kHid * es); L.tr_t[i] = take(kHid * kHid * es); }
  L.class_bias = take((n_images > 0 ? n_images : 1) * 2 * kHid * 4);
  L.total = off;
  return L;
}


// ---- fused path (field_fused.cu), precision EONERF_PREC_BF16_FUSED ---------------------------------------------------
int64_t fused_prepared_extra_bytes(int64_t n_images);
int64_t fused_stash_bytes(int64_t n_pts, int density_only);
int64_t fused_scratch_bytes(int64_t n_pts, int64_t n_images, int density_only);
// `prepared` is the layered bf16 blob (PrepLayout) followed by the fused extras at PrepLayout::total
int fused_prepare(int field, const EonerfFieldParams* p, void* prepared, cudaStream_t s);
int fused_field_fwd(const EonerfFieldFwdArgs* a, cudaStream_t s);
int fused_field_bwd(const EonerfFieldBwdArgs* a, cudaStream_t s);

}  // namespace eonerf
