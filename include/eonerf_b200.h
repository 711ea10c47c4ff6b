/*
 * eonerf_b200.h — C ABI of the B200-native EO-NeRF per-ray rendering hot path.
 *
 * Drop-in boundary.  The reference has no FFI of its own for this path: its only
 * native code is the un-vendored nerfacc v0.5.2 extension (/root/reference/setup_env.sh:10) reached
 * through three Python operators, everything else is ATen.  Each entry point below names the
 * reference interface it replaces (file:line relative to /root/reference).
 *
 * Conventions
 *   - plain C, POD argument structs, raw *device* pointers + int64 sizes; no torch types.
 *   - the caller owns ALL memory (inputs, outputs, workspaces); the library never allocates device
 *     memory and never synchronises; work is enqueued on the `stream` argument (a cudaStream_t).
 *   - contiguous row-major fp32 / int64 unless a `*_stride` (in elements) is given.
 *   - return 0 on success, a negative EONERF_E* code on failure; eonerf_last_error() gives the
 *     message of the last failure on the calling thread.  There is no CPU fallback: a missing
 *     device or a device that is not sm_100 is an error.
 */
#ifndef EONERF_B200_H
#define EONERF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EONERF_ABI_VERSION 16

#define EONERF_OK 0
#define EONERF_EINVAL (-1)   /* bad argument / unsupported shape */
#define EONERF_ECUDA (-2)    /* CUDA runtime or launch error     */
#define EONERF_EDEVICE (-3)  /* not an sm_100 device             */

typedef void* eonerf_stream_t; /* cudaStream_t */

int eonerf_abi_version(void);
const char* eonerf_last_error(void);
/* 0 if the current device can run the kernels (compute capability 10.x), else EONERF_EDEVICE. */
int eonerf_check_device(void);

/* Instrumentation (bench.py): number of kernels this library launched since the last reset; per-launch CUDA-event
 * timing of the GEMM kernels.  kinds: 0 = tcgen05 NT GEMM (forward / input gradients), 1 = tcgen05 TN GEMM (parameter
 * gradients), 2 = SIMT GEMMs.  `flops`/`bytes` are the ALGORITHMIC counts of the timed launches (2*M*N*K with the
 * un-padded K; operand + result bytes). */
typedef struct { int64_t launches; double ms; double flops; double bytes; } EonerfProfile;
int64_t eonerf_launch_count(int32_t reset);
int eonerf_profile_enable(int32_t on);
int eonerf_profile_read(EonerfProfile* out, int32_t n_kinds);   /* synchronises the recorded events, then clears them */

/* ------------------------------------------------------------------------------------------------
 * Stratified sampling + cube mask + order-preserving compaction.
 * Replaces satnerf_sampling / perturb_z_vals / filter_pts_outside_cube
 *   (sat_rendering.py:18-22,46-54,56-84) and count_number_of_pts_per_nerfacc_ray (:10-16).
 * Bit-exact with the reference given the same uniforms `u` and the same `z_steps`
 * (= torch.linspace(0,1,n), passed in because its rounding is torch's own).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  const float* origins;   int64_t origins_stride;   /* [B,3] rows, stride in floats (11 for a ray table) */
  const float* viewdirs;  int64_t viewdirs_stride;
  const float* near;      int64_t near_stride;      /* [B] or NULL (=0)  (sat_rendering.py:60-61) */
  const float* u;                                   /* [B,n] uniforms in [0,1) (torch.rand_like, :52) */
  const float* z_steps;                             /* [n] */
  int64_t n_rays;
  int32_t n_samples;                                /* n = int(2/render_step_size) (:64) */
  /* outputs; capacity of the three packed arrays is B*(n-1) */
  int64_t* ray_indices;                             /* [P] */
  float* t_starts;                                  /* [P] */
  float* t_ends;                                    /* [P] */
  float* pts_per_ray;                               /* [B] fp32 counts (the reference returns fp32, :14) */
  int64_t* ray_offsets;                             /* [B+1] exclusive prefix of the counts (packed info) */
  int64_t* stats;                                   /* [2]: P, number of rays with 0 samples */
  /* Device-side condition (NULL: always run).  When set, the whole call is a no-op unless *run_if != 0, and it writes
   * stats[0] only: the sync-free form of the reference's "some ray kept no sample -> draw again" (sat_rendering.py:259-262),
   * called with run_if = &stats[1] of the first draw and the same output buffers. */
  const int64_t* run_if;
  /* Scratch of eonerf_sample_scratch_bytes(n_rays) bytes (8-byte aligned; zeroed by the call).  When given, the sampler runs
   * as ONE pass: each CTA evaluates a tile of rays once (kept intervals stay in registers), a decoupled look-back over
   * per-tile descriptors in the scratch gives the tile its global offset, and the kept intervals are written in order.
   * NULL: the three-launch count / scan / scatter form (evaluates every interval twice). */
  int64_t* scratch;
} EonerfSampleArgs;
int64_t eonerf_sample_scratch_bytes(int64_t n_rays);
int eonerf_sample_compact(const EonerfSampleArgs* a, eonerf_stream_t stream);

/* ray_offsets[B+1] from sorted ray_indices[P] (nerfacc pack_info; used when the operator-level API is
 * called with ray_indices only: radiance_fields/eonerf.py:229-235). */
int eonerf_pack_info(const int64_t* ray_indices, int64_t n_pts, int64_t n_rays, int64_t* ray_offsets,
                     eonerf_stream_t stream);

/* t_ends[last sample of every non-empty ray] = value   (radiance_fields/eonerf.py:218-220: 1e10) */
int eonerf_set_last_t_end(float* t_ends, const int64_t* ray_offsets, int64_t n_rays, float value,
                          eonerf_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * nerfacc v0.5.2 volume rendering operators (flattened samples + packed info).
 * Replace render_transmittance_from_density / render_weight_from_density / accumulate_along_rays
 *   (call sites radiance_fields/eonerf.py:186-193,229-242; sat_rendering.py:106-110).
 * Deterministic: one warp owns one ray, no atomics.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  const float* t_starts; const float* t_ends; const float* sigmas;  /* [P] */
  const int64_t* ray_offsets; int64_t n_rays; int64_t n_pts;
  float* weights; float* trans; float* alphas;                       /* [P] each, any may be NULL */
} EonerfWeightsFwdArgs;
int eonerf_weights_fwd(const EonerfWeightsFwdArgs* a, eonerf_stream_t stream);

typedef struct {
  const float* t_starts; const float* t_ends; const float* sigmas;  /* [P] */
  const int64_t* ray_offsets; int64_t n_rays; int64_t n_pts;
  const float* g_weights; const float* g_trans; const float* g_alphas; /* [P] each, any may be NULL */
  float* g_sigmas;                                                   /* [P] */
} EonerfWeightsBwdArgs;
int eonerf_weights_bwd(const EonerfWeightsBwdArgs* a, eonerf_stream_t stream);

typedef struct {
  const float* weights;       /* [P] */
  const float* values;        /* [P,C] or NULL (C=1, value 1) */
  int32_t n_channels;
  const int64_t* ray_offsets; int64_t n_rays; int64_t n_pts;
  float* out;                 /* [B,C] */
} EonerfAccumFwdArgs;
int eonerf_accumulate_fwd(const EonerfAccumFwdArgs* a, eonerf_stream_t stream);

typedef struct {
  const float* weights; const float* values; int32_t n_channels;
  const int64_t* ray_offsets; int64_t n_rays; int64_t n_pts;
  const float* g_out;         /* [B,C] */
  float* g_weights;           /* [P] or NULL */
  float* g_values;            /* [P,C] or NULL */
} EonerfAccumBwdArgs;
int eonerf_accumulate_bwd(const EonerfAccumBwdArgs* a, eonerf_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Fused EO-NeRF compositing of one chunk of rays.
 * Replaces the weights + five accumulate_along_rays calls + beta_min of EONerfMLP.rendering
 *   (radiance_fields/eonerf.py:229-246).  Per-ray output row `comp[B,12]`:
 *     0:3 albedo, 3 depth, 4 beta (+0.05), 5 transient_s, 6:9 ambient (NOT yet x0.2), 9 sum of weights,
 *     10,11 unused (0).
 * ---------------------------------------------------------------------------------------------- */
#define EONERF_COMP_COLS 12
typedef struct {
  const float* t_starts; const float* t_ends;   /* [P]; t_ends already holds 1e10 at each ray's last sample */
  const float* z_mid;                           /* [P] (t_starts+t_ends)/2 computed BEFORE the 1e10 write (eonerf.py:206) */
  const float* sigma;                           /* [P] */
  const float* albedo;                          /* [P,3] */
  const float* transient_s;                     /* [P] */
  const float* transient_beta;                  /* [P] */
  const float* ambient_ray;                     /* [B,3] per-ray ambient colour (constant along a ray, eonerf.py:163-164,204) */
  const int64_t* ray_offsets; int64_t n_rays; int64_t n_pts;
  float beta_min;                               /* 0.05, eonerf.py:87,243 */
  float* comp;                                  /* [B,12] */
} EonerfCompositeFwdArgs;
int eonerf_composite_fwd(const EonerfCompositeFwdArgs* a, eonerf_stream_t stream);

typedef struct {
  const float* t_starts; const float* t_ends; const float* z_mid; const float* sigma;
  const float* albedo; const float* transient_s; const float* transient_beta; const float* ambient_ray;
  const int64_t* ray_offsets; int64_t n_rays; int64_t n_pts;
  const float* g_comp;                          /* [B,12] */
  float* g_sigma;                               /* [P] */
  float* g_albedo;                              /* [P,3] */
  float* g_transient_s;                         /* [P] */
  float* g_transient_beta;                      /* [P] */
  float* g_ambient_ray;                         /* [B,3] */
} EonerfCompositeBwdArgs;
int eonerf_composite_bwd(const EonerfCompositeBwdArgs* a, eonerf_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Sun-direction shadow pass.  Replaces compute_geometric_shadows (sat_rendering.py:87-118) around the
 * density query: ray set-up (:90-91), transmittance at the last kept sample (:106-116) and their backward.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  const float* origins;  int64_t origins_stride;
  const float* viewdirs; int64_t viewdirs_stride;
  const float* sundirs;  int64_t sundirs_stride;
  const float* depth;    int64_t depth_stride;   /* [B] */
  int64_t n_rays;
  float* sun_rays;                               /* [B,6]: origin = o + depth*d (:90), dir = -1*sun (:91) */
} EonerfSunRaysArgs;
int eonerf_sun_rays(const EonerfSunRaysArgs* a, eonerf_stream_t stream);

typedef struct {
  const float* t_starts; const float* t_ends; const float* sigma;  /* [Q] */
  const int64_t* ray_offsets; int64_t n_rays; int64_t n_pts;
  float* geo_shadow;                              /* [B]: T before the last kept sample; 1 if the ray has none */
} EonerfShadowFwdArgs;
int eonerf_shadow_fwd(const EonerfShadowFwdArgs* a, eonerf_stream_t stream);

typedef struct {
  const float* t_starts; const float* t_ends;     /* [Q] */
  const int64_t* ray_offsets; int64_t n_rays; int64_t n_pts;
  const float* geo_shadow; const float* g_geo_shadow; /* [B] */
  float* g_sigma;                                 /* [Q] */
} EonerfShadowBwdArgs;
int eonerf_shadow_bwd(const EonerfShadowBwdArgs* a, eonerf_stream_t stream);

/* d depth[r] += sum_c (sum_{i in ray r} g_x[i,c]) * viewdir[r,c]   (chain rule through sat_rendering.py:90) */
typedef struct {
  const float* g_x;                               /* [Q,3] gradient wrt the sun-sample positions */
  const int64_t* ray_offsets; int64_t n_rays; int64_t n_pts;
  const float* viewdirs; int64_t viewdirs_stride;
  float* g_depth; int64_t g_depth_stride;         /* [B] (strided), accumulated into */
} EonerfSunOriginBwdArgs;
int eonerf_sun_origin_bwd(const EonerfSunOriginBwdArgs* a, eonerf_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Irradiance model + radiometric normalisation + 21-column packing.
 * Replaces sat_rendering.py:265,269-276,288-312.
 * ---------------------------------------------------------------------------------------------- */
#define EONERF_OUT_COLS 21
typedef struct {
  const float* comp;                /* [B,12] from eonerf_composite_fwd */
  const float* geo_shadow;          /* [B] or NULL (= 1, epoch_idx < 2: s = 1, sat_rendering.py:269-272) */
  const float* pts_per_ray;         /* [B] */
  const float* sc_pts_per_ray;      /* [B] or NULL (= 1) */
  const int64_t* img_idx; int64_t img_idx_stride;  /* [B] */
  int32_t eval_mode;                /* 1: use img_idx[0] for every ray (sat_rendering.py:288-289) */
  const float* radiometric;         /* [n_img,9] or NULL (A=1,b=0) */
  int64_t n_images;
  int64_t n_rays;
  float* out;                       /* [B,21] */
} EonerfEpilogueFwdArgs;
int eonerf_epilogue_fwd(const EonerfEpilogueFwdArgs* a, eonerf_stream_t stream);

typedef struct {
  const float* comp; const float* geo_shadow;
  const int64_t* img_idx; int64_t img_idx_stride; int32_t eval_mode;
  const float* radiometric; int64_t n_images; int64_t n_rays;
  const float* g_out;               /* [B,21] */
  float* g_comp;                    /* [B,12] */
  float* g_geo_shadow;              /* [B] or NULL */
  float* g_radiometric;             /* [n_img,9] accumulated into (fp32 atomics), or NULL */
} EonerfEpilogueBwdArgs;
int eonerf_epilogue_bwd(const EonerfEpilogueBwdArgs* a, eonerf_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Radiance-field MLP.  Replaces EONerfMLP.forward / query_density (radiance_fields/eonerf.py:141-170),
 * MLP.forward (radiance_fields/mlp.py:87-101), SinusoidalEncoder.forward (mlp.py:190-208) and, for
 * BASELINE config 2, VanillaNeRFRadianceField.forward (mlp.py:245-250) — forward and backward.
 *
 * precision: EONERF_PREC_FP32       fp32 storage, SIMT kernels (exactness mode, used for the 1e-5 parity tests)
 *            EONERF_PREC_BF16       bf16 storage, tcgen05/TMEM/TMA tensor-core GEMMs, fp32 accumulation
 *            EONERF_PREC_BF16_SIMT  bf16 storage, SIMT GEMMs (on-device cross-check of the tensor-core kernels)
 *            EONERF_PREC_BF16_FUSED bf16 storage, the whole MLP as one persistent tcgen05 kernel per direction; activations
 *                                   stay in shared memory / TMEM between layers, the stash is tile-blocked (EO-NeRF field
 *                                   only).  eonerf_field_fwd accepts stash == NULL in this mode: inference, nothing kept.
 * ---------------------------------------------------------------------------------------------- */
#define EONERF_PREC_FP32 0
#define EONERF_PREC_BF16 1
#define EONERF_PREC_BF16_SIMT 2
#define EONERF_PREC_BF16_FUSED 3

#define EONERF_FIELD_EONERF 0
#define EONERF_FIELD_VANILLA 1

/* Pointers to the fp32 master parameters (checkpoint layout).  The same struct
 * is used for gradients.  Weight matrices are [out,in] row-major as in nn.Linear. */
typedef struct {
  float* trunk_w[8]; float* trunk_b[8];     /* base_mlp.hidden_layers.{0..7}: (256,63) (256,256)x4 (256,319) (256,256)x2 */
  float* sigma_w; float* sigma_b;           /* (1,256) (1) */
  float* bott_w; float* bott_b;             /* (256,256) (256) */
  float* head0_w; float* head0_b;           /* eonerf: albedo_mlp.hidden_layers.0 (128,256); vanilla: rgb_layer.hidden_layers.0 (128,283) */
  float* head1_w; float* head1_b;           /* (3,128) (3) */
  float* trans_w[4]; float* trans_b[4];     /* transient_mlp.hidden_layers.{0..3}: (128,260) (128,128)x3  (eonerf only) */
  float* ts_w; float* ts_b;                 /* transient_scalar (1,128) (1) */
  float* tb_w; float* tb_b;                 /* transient_beta   (1,128) (1) */
  float* transient_emb;                     /* (n_img,4) */
  int64_t n_images;
} EonerfFieldParams;

/* bytes of the prepared-parameter blob / of the forward stash / of the backward scratch */
int64_t eonerf_field_prepared_bytes(int32_t field, int32_t precision, int64_t n_images);
int64_t eonerf_field_stash_bytes(int32_t field, int32_t precision, int64_t n_pts, int32_t density_only);
int64_t eonerf_field_scratch_bytes(int32_t field, int32_t precision, int64_t n_pts, int64_t n_images);

/* Convert the fp32 master weight matrices into the kernels' operand layouts (K padded, W [out,Kp] and
 * W^T [Kp,out], fp32 or bf16) and build the per-image bias table that folds the 4-d transient embedding
 * into layer transient_mlp.0 (eonerf.py:165-167).  Call after every optimiser step. */
int eonerf_field_prepare(int32_t field, int32_t precision, const EonerfFieldParams* params, void* prepared,
                         eonerf_stream_t stream);

typedef struct {
  int32_t field; int32_t precision;
  const EonerfFieldParams* params;  /* fp32 masters: biases, the 1- and 3-wide heads */
  const void* prepared;
  int64_t n_pts;
  /* sample positions: either explicit x[N,3] ... */
  const float* x;
  /* ... or derived from rays: x = o[ri] + d[ri]*z_mid, z_mid = (t_starts+t_ends)/2 (eonerf.py:202-207) */
  const float* origins;  int64_t origins_stride;
  const float* viewdirs; int64_t viewdirs_stride;
  const int64_t* ray_indices;      /* [N] */
  const float* t_starts; const float* t_ends;   /* [N]; read BEFORE any 1e10 write */
  float* z_mid;                    /* [N] out (may be NULL) */
  /* conditioning: image index per ray [B] (looked up through ray_indices) when ray_indices != NULL,
   * else per sample [N] */
  const int64_t* img_idx; int64_t img_idx_stride;
  const float* cond_dirs; int64_t cond_dirs_stride; /* vanilla field only: per-sample view directions [N,3] */
  int32_t cond_dirs_per_ray;       /* 1: cond_dirs holds one row per RAY, looked up through ray_indices */
  int32_t density_only;            /* 1: query_density (trunk + sigma), eonerf.py:141-145 */
  void* stash;                     /* activations kept for backward */
  /* outputs, fp32 */
  float* sigma;                    /* [N] */
  float* rgb;                      /* [N,3] albedo (eonerf) / rgb (vanilla) */
  float* transient_s;              /* [N] */
  float* transient_beta;           /* [N] */
  /* Sync-free form (EONERF_PREC_BF16_FUSED only): when non-NULL the live sample count is read ON THE DEVICE from
   * *n_pts_dev (<= n_pts); n_pts is then the capacity every buffer (stash included) is sized for, and rows >= *n_pts_dev
   * of the outputs are left untouched.  Lets a whole training step run without a host read of P (CUDA-graph capture). */
  const int64_t* n_pts_dev;
  /* Vanilla field in the fused mode: the view-direction term of rgb_layer.hidden_layers.0 (mlp.py:153-165: columns 256:283 of
   * its weight times enc4(dir)) enters the fused program as a bias row per conditioning row.  dir_bias = caller-provided
   * scratch [n_cond,128] fp32 (written by this call, read again by eonerf_field_bwd's caller only through the stash);
   * n_cond = rows of cond_dirs (rays when cond_dirs_per_ray, else samples). */
  float* dir_bias; int64_t n_cond;
} EonerfFieldFwdArgs;
int eonerf_field_fwd(const EonerfFieldFwdArgs* a, eonerf_stream_t stream);

typedef struct {
  int32_t field; int32_t precision;
  const EonerfFieldParams* params;
  const void* prepared;
  int64_t n_pts;
  int32_t density_only;
  const void* stash;               /* written by the matching eonerf_field_fwd */
  void* scratch;
  const float* sigma; const float* rgb; const float* transient_s; const float* transient_beta; /* forward outputs */
  const float* g_sigma; const float* g_rgb; const float* g_transient_s; const float* g_transient_beta; /* may be NULL (=0) */
  const EonerfFieldParams* grads;  /* fp32 gradients, same shapes as params, ACCUMULATED into; NULL: skip parameter gradients */
  float* g_x;                      /* [N,3] gradient wrt positions, or NULL */
  const int64_t* n_pts_dev;        /* as in EonerfFieldFwdArgs (must match the forward call) */
  const float* cond_dirs; int64_t cond_dirs_stride; int64_t n_cond;   /* vanilla field, fused mode: the forward call's cond_dirs */
} EonerfFieldBwdArgs;
int eonerf_field_bwd(const EonerfFieldBwdArgs* a, eonerf_stream_t stream);

/* Per-ray ambient colour sigmoid(W1 relu(W0 enc4(sun) + b0) + b1) (eonerf.py:163-164), fwd + bwd.
 * Tiny (B rows); always fp32.  stash: B*160 floats, scratch: B*136 floats. */
typedef struct {
  const float* sundirs; int64_t sundirs_stride; int64_t n_rays;
  const float* w0; const float* b0; const float* w1; const float* b1;   /* (128,27) (128) (3,128) (3) */
  float* stash;
  float* ambient;                  /* [B,3] */
} EonerfAmbientFwdArgs;
int eonerf_ambient_fwd(const EonerfAmbientFwdArgs* a, eonerf_stream_t stream);

typedef struct {
  int64_t n_rays;
  const float* w0; const float* w1;
  const float* stash; float* scratch;
  const float* ambient; const float* g_ambient;                         /* [B,3] [B,3] */
  float* g_w0; float* g_b0; float* g_w1; float* g_b1;                    /* accumulated into (atomics) */
} EonerfAmbientBwdArgs;
int eonerf_ambient_bwd(const EonerfAmbientBwdArgs* a, eonerf_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Building blocks exposed for tests and micro-benchmarks:
 *   Y[M,N] = act(X[M,K] W[N,K]^T + b)        and        dW[N,K] += dY[M,N]^T X[M,K], db[N] += colsum(dY)
 * X, W, Y, dY are fp32 (EONERF_PREC_FP32) or bf16 (EONERF_PREC_BF16: tcgen05; EONERF_PREC_BF16_SIMT).
 * Tensor-core constraints: K % 8 == 0, leading dimensions % 8 == 0, 16-byte aligned bases.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t precision;
  const void* x; int64_t ldx; const void* w; int64_t ldw; const float* bias;
  int64_t m; int32_t n; int32_t k;
  int32_t act;                      /* 0 none, 1 relu */
  void* y; int64_t ldy;
} EonerfLinearArgs;
int eonerf_linear_fwd(const EonerfLinearArgs* a, eonerf_stream_t stream);

typedef struct {
  int32_t precision;
  const void* dy; int64_t lddy; const void* x; int64_t ldx;
  int64_t m; int32_t n; int32_t k;
  float* dw; int64_t lddw;          /* accumulated into */
  float* db;                        /* accumulated into, or NULL */
} EonerfDwArgs;
int eonerf_linear_dw(const EonerfDwArgs* a, eonerf_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Optimiser step of the training loop.  Replaces torch.optim.Adam(lr=5e-4).step() (train_eonerf.py:57,158-160; default
 * betas (0.9, 0.999), eps 1e-8, no weight decay / amsgrad) for parameters, gradients and moments that are views of four
 * flat fp32 buffers (16-byte aligned).  `step` points to 4 DEVICE floats: [0] the step counter (torch keeps Adam's step as
 * an fp32 tensor), incremented by one per call; [1], [2] scratch for this step's lr/(1-beta1^t) and sqrt(1-beta2^t), which
 * a one-thread kernel computes in double precision before the update kernel: the same two launches can be replayed from a
 * CUDA graph.  grad_scale multiplies the gradient first (1/world_size after a sum all-reduce).
 * lr_dev (optional): one DEVICE double holding the learning rate; when set it overrides `lr`, so a captured step follows
 * a scheduler (StepLR(gamma=0.9), train_eonerf.py:64,304) through a host->device refresh of that double instead of
 * replaying the value frozen at capture time.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  float* param; const float* grad; float* exp_avg; float* exp_avg_sq;
  int64_t n;
  float* step;
  double lr; double beta1; double beta2; double eps;
  float grad_scale;
  const double* lr_dev;
} EonerfAdamArgs;
int eonerf_adam_step(const EonerfAdamArgs* a, eonerf_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Device-resident ray table.  Replaces the per-ray SatelliteDataset.__getitem__ + DataLoader
 * collate + host->device copies of the training loop (datasets/satellite.py:799-807, train_eonerf.py:70,99-109): a batch
 * is rows perm[first .. first+batch) of all_rays[N,11] / all_rgbs[N,3] / all_ids_img[N] (int64), all resident in HBM.
 * perm is a device-side permutation of 0..N-1 (the DataLoader's shuffle=True epoch order).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  const float* all_rays; int64_t rays_stride;     /* [N,11], row stride in floats */
  const float* all_rgbs; int64_t rgbs_stride;     /* [N,3] */
  const int64_t* all_ts;                          /* [N] image index per ray */
  const int64_t* perm;                            /* [>= first+batch] row indices */
  int64_t n_rows; int64_t first; int64_t batch;
  float* rays_out;                                /* [batch,11] */
  float* rgbs_out;                                /* [batch,3] */
  int64_t* ts_out;                                /* [batch] (viewed as [batch,1]) */
  int64_t* idx_out;                               /* [batch] source rows ("idx" of the reference's sample dict), or NULL */
} EonerfGatherBatchArgs;
int eonerf_gather_batch(const EonerfGatherBatchArgs* a, eonerf_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Training losses on the packed per-ray outputs of eonerf_epilogue_fwd (out[B,21]), value and gradient in one pass.
 * Replaces metrics.uncertainty_aware_loss (metrics.py:17-22; mode 1) / the epoch<2 MSE (train_eonerf.py:139-140; mode 0)
 * and their autograd backward.  loss[0] = total, loss[1] = colour term, loss[2] = log-beta term (the reference's loss_dict).
 * g_out[B,21] receives d loss / d out (zero outside the rgb columns 0:3 and the beta column 12).  Deterministic.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  const float* out;            /* [B,21] */
  const float* gt_rgb;         /* [B,3] */
  int64_t n_rays;
  int32_t mode;                /* 0 mse, 1 uncertainty-aware */
  float* loss;                 /* [3] */
  float* g_out;                /* [B,21] */
  float* partials;             /* scratch, eonerf_loss_partials(B) floats */
} EonerfLossArgs;
int64_t eonerf_loss_partials(int64_t n_rays);
int eonerf_loss_fwd_bwd(const EonerfLossArgs* a, eonerf_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Uniform in-box ray marching: the sampler of BASELINE configs[1] (train_mlp_nerf.py:155-170 -> nerfacc v0.5.2
 * estimator.sampling with every cell occupied; the reference's own helper module `utils2` is missing, train_mlp_nerf.py:17,
 * so this path is a benchmark without a pin).  Two calls around an exclusive scan of `counts` (done by the caller):
 *   eonerf_march_count: counts[r] = number of render_step_size intervals of ray r inside aabb and [near_plane, far_plane]
 *                       (+ t0_out / t_max_out per ray for the second call); jitter[r] in [0,1) shifts the ray's first
 *                       interval (nerfacc's `stratified`), NULL = no shift;
 *   eonerf_march_write: packed (ray_indices, t_starts, t_ends) from ray_offsets[B+1].
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  const float* origins; int64_t origins_stride;   /* [B,3] */
  const float* viewdirs; int64_t viewdirs_stride; /* [B,3] */
  const float* jitter;                            /* [B] or NULL */
  int64_t n_rays;
  float aabb[6];                                  /* minx miny minz maxx maxy maxz */
  float near_plane; float far_plane; float step;
  int32_t max_per_ray;
  int64_t* counts;                                /* [B]   (march_count) */
  float* t0_out; float* t_max_out;                /* [B]   (march_count writes, march_write reads) */
  const int64_t* ray_offsets;                     /* [B+1] (march_write) */
  int64_t* ray_indices; float* t_starts; float* t_ends;   /* [P] (march_write) */
} EonerfMarchArgs;
int eonerf_march_count(const EonerfMarchArgs* a, eonerf_stream_t stream);
int eonerf_march_write(const EonerfMarchArgs* a, eonerf_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Evaluation epilogue.
 * eonerf_utm_points replaces SatelliteDataset.get_utmalt_from_nerf_prediction (datasets/satellite.py:502-531, the
 * utm_sampling branch): x = (o + d * depth) * scene_scale + scene_offset with rays and depth promoted to fp64 first.
 * Outputs are optional planes [N] (fp64 easts / norths / alts, fp32 altitude for the render arm's altitude column).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  const float* rays; int64_t rays_stride;        /* [N,>=6] rows [o(3) d(3) ...], stride in floats */
  const float* depth; int64_t depth_stride;      /* [N] rendered depth */
  int64_t n_rays;
  double scene_scale[3]; double scene_offset[3];
  double* easts; double* norths; double* alts;   /* [N] each, or NULL */
  float* alt_f32;                                /* [N] or NULL */
} EonerfUtmPointsArgs;
int eonerf_utm_points(const EonerfUtmPointsArgs* a, eonerf_stream_t stream);

/* eonerf_dsm_rasterize replaces the plyflatten call of SatelliteDataset.get_dsm_from_nerf_prediction
 * (datasets/satellite.py:548-587): plyflatten(cloud, xoff, yoff, resolution, xsize, ysize, radius, sigma) -> dsm[ysize,xsize]
 * fp32, NaN where no point fell.  Cell of a point: i = floor((x - xoff)/res), j = floor((yoff - y)/res); cells within
 * `radius` (k1^2 + k2^2 <= radius^2) receive its height with weight 1 (sigma = inf) or the Gaussian of the distance to the
 * cell centre.  Points with depth < 0 are dropped (:561) and norths < 0 are shifted by negative_north_shift (:559, 10e6).
 * `acc` = caller-provided scratch of 2 * xsize * ysize doubles (zeroed here).  plyflatten is an un-vendored dependency of
 * the reference: semantics restated from its published source, parity unpinned.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  const double* easts; const double* norths; const double* alts;   /* [N] */
  const float* depth; int64_t depth_stride;                        /* [N] or NULL */
  int64_t n_points;
  double xoff; double yoff; double resolution;
  int32_t xsize; int32_t ysize; int32_t radius;
  double sigma;                                                    /* +inf: unweighted mean */
  double negative_north_shift;
  double* acc;                                                     /* scratch [ysize*xsize*2] */
  float* dsm;                                                      /* [ysize,xsize] */
} EonerfDsmArgs;
int eonerf_dsm_rasterize(const EonerfDsmArgs* a, eonerf_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* EONERF_B200_H */
