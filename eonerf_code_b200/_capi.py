"""ctypes binding of the C-ABI library `libeonerf_b200.so` (include/eonerf_b200.h).

There is no CPU fallback: if the library is missing or the device is not an sm_100 GPU every call raises.
Struct layouts mirror include/eonerf_b200.h field by field (tests/test_capi.py checks the exported symbols).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libeonerf_b200.so")

ABI_VERSION = 16
COMP_COLS = 12
OUT_COLS = 21
PREC_FP32, PREC_BF16, PREC_BF16_SIMT, PREC_BF16_FUSED = 0, 1, 2, 3
FIELD_EONERF, FIELD_VANILLA = 0, 1

P = C.c_void_p
I64 = C.c_int64
I32 = C.c_int32
F32 = C.c_float


class Profile(C.Structure):
    _fields_ = [("launches", I64), ("ms", C.c_double), ("flops", C.c_double), ("bytes", C.c_double)]


class SampleArgs(C.Structure):
    _fields_ = [("origins", P), ("origins_stride", I64), ("viewdirs", P), ("viewdirs_stride", I64),
                ("near", P), ("near_stride", I64), ("u", P), ("z_steps", P), ("n_rays", I64), ("n_samples", I32),
                ("ray_indices", P), ("t_starts", P), ("t_ends", P), ("pts_per_ray", P), ("ray_offsets", P), ("stats", P),
                ("run_if", P), ("scratch", P)]


class WeightsFwdArgs(C.Structure):
    _fields_ = [("t_starts", P), ("t_ends", P), ("sigmas", P), ("ray_offsets", P), ("n_rays", I64), ("n_pts", I64),
                ("weights", P), ("trans", P), ("alphas", P)]


class WeightsBwdArgs(C.Structure):
    _fields_ = [("t_starts", P), ("t_ends", P), ("sigmas", P), ("ray_offsets", P), ("n_rays", I64), ("n_pts", I64),
                ("g_weights", P), ("g_trans", P), ("g_alphas", P), ("g_sigmas", P)]


class AccumFwdArgs(C.Structure):
    _fields_ = [("weights", P), ("values", P), ("n_channels", I32), ("ray_offsets", P), ("n_rays", I64), ("n_pts", I64),
                ("out", P)]


class AccumBwdArgs(C.Structure):
    _fields_ = [("weights", P), ("values", P), ("n_channels", I32), ("ray_offsets", P), ("n_rays", I64), ("n_pts", I64),
                ("g_out", P), ("g_weights", P), ("g_values", P)]


class CompositeFwdArgs(C.Structure):
    _fields_ = [("t_starts", P), ("t_ends", P), ("z_mid", P), ("sigma", P), ("albedo", P), ("transient_s", P),
                ("transient_beta", P), ("ambient_ray", P), ("ray_offsets", P), ("n_rays", I64), ("n_pts", I64),
                ("beta_min", F32), ("comp", P)]


class CompositeBwdArgs(C.Structure):
    _fields_ = [("t_starts", P), ("t_ends", P), ("z_mid", P), ("sigma", P), ("albedo", P), ("transient_s", P),
                ("transient_beta", P), ("ambient_ray", P), ("ray_offsets", P), ("n_rays", I64), ("n_pts", I64),
                ("g_comp", P), ("g_sigma", P), ("g_albedo", P), ("g_transient_s", P), ("g_transient_beta", P),
                ("g_ambient_ray", P)]


class SunRaysArgs(C.Structure):
    _fields_ = [("origins", P), ("origins_stride", I64), ("viewdirs", P), ("viewdirs_stride", I64),
                ("sundirs", P), ("sundirs_stride", I64), ("depth", P), ("depth_stride", I64), ("n_rays", I64),
                ("sun_rays", P)]


class ShadowFwdArgs(C.Structure):
    _fields_ = [("t_starts", P), ("t_ends", P), ("sigma", P), ("ray_offsets", P), ("n_rays", I64), ("n_pts", I64),
                ("geo_shadow", P)]


class ShadowBwdArgs(C.Structure):
    _fields_ = [("t_starts", P), ("t_ends", P), ("ray_offsets", P), ("n_rays", I64), ("n_pts", I64),
                ("geo_shadow", P), ("g_geo_shadow", P), ("g_sigma", P)]


class SunOriginBwdArgs(C.Structure):
    _fields_ = [("g_x", P), ("ray_offsets", P), ("n_rays", I64), ("n_pts", I64), ("viewdirs", P), ("viewdirs_stride", I64),
                ("g_depth", P), ("g_depth_stride", I64)]


class EpilogueFwdArgs(C.Structure):
    _fields_ = [("comp", P), ("geo_shadow", P), ("pts_per_ray", P), ("sc_pts_per_ray", P), ("img_idx", P),
                ("img_idx_stride", I64), ("eval_mode", I32), ("radiometric", P), ("n_images", I64), ("n_rays", I64),
                ("out", P)]


class EpilogueBwdArgs(C.Structure):
    _fields_ = [("comp", P), ("geo_shadow", P), ("img_idx", P), ("img_idx_stride", I64), ("eval_mode", I32),
                ("radiometric", P), ("n_images", I64), ("n_rays", I64), ("g_out", P), ("g_comp", P), ("g_geo_shadow", P),
                ("g_radiometric", P)]


class FieldParams(C.Structure):
    _fields_ = [("trunk_w", P * 8), ("trunk_b", P * 8), ("sigma_w", P), ("sigma_b", P), ("bott_w", P), ("bott_b", P),
                ("head0_w", P), ("head0_b", P), ("head1_w", P), ("head1_b", P), ("trans_w", P * 4), ("trans_b", P * 4),
                ("ts_w", P), ("ts_b", P), ("tb_w", P), ("tb_b", P), ("transient_emb", P), ("n_images", I64)]


class FieldFwdArgs(C.Structure):
    _fields_ = [("field", I32), ("precision", I32), ("params", C.POINTER(FieldParams)), ("prepared", P), ("n_pts", I64),
                ("x", P), ("origins", P), ("origins_stride", I64), ("viewdirs", P), ("viewdirs_stride", I64),
                ("ray_indices", P), ("t_starts", P), ("t_ends", P), ("z_mid", P), ("img_idx", P), ("img_idx_stride", I64),
                ("cond_dirs", P), ("cond_dirs_stride", I64), ("cond_dirs_per_ray", I32), ("density_only", I32), ("stash", P),
                ("sigma", P), ("rgb", P), ("transient_s", P), ("transient_beta", P), ("n_pts_dev", P), ("dir_bias", P), ("n_cond", I64)]


class FieldBwdArgs(C.Structure):
    _fields_ = [("field", I32), ("precision", I32), ("params", C.POINTER(FieldParams)), ("prepared", P), ("n_pts", I64),
                ("density_only", I32), ("stash", P), ("scratch", P),
                ("sigma", P), ("rgb", P), ("transient_s", P), ("transient_beta", P),
                ("g_sigma", P), ("g_rgb", P), ("g_transient_s", P), ("g_transient_beta", P),
                ("grads", C.POINTER(FieldParams)), ("g_x", P), ("n_pts_dev", P), ("cond_dirs", P), ("cond_dirs_stride", I64), ("n_cond", I64)]


class AmbientFwdArgs(C.Structure):
    _fields_ = [("sundirs", P), ("sundirs_stride", I64), ("n_rays", I64), ("w0", P), ("b0", P), ("w1", P), ("b1", P),
                ("stash", P), ("ambient", P)]


class AmbientBwdArgs(C.Structure):
    _fields_ = [("n_rays", I64), ("w0", P), ("w1", P), ("stash", P), ("scratch", P), ("ambient", P), ("g_ambient", P),
                ("g_w0", P), ("g_b0", P), ("g_w1", P), ("g_b1", P)]


class LinearArgs(C.Structure):
    _fields_ = [("precision", I32), ("x", P), ("ldx", I64), ("w", P), ("ldw", I64), ("bias", P), ("m", I64), ("n", I32),
                ("k", I32), ("act", I32), ("y", P), ("ldy", I64)]


class DwArgs(C.Structure):
    _fields_ = [("precision", I32), ("dy", P), ("lddy", I64), ("x", P), ("ldx", I64), ("m", I64), ("n", I32), ("k", I32),
                ("dw", P), ("lddw", I64), ("db", P)]


class AdamArgs(C.Structure):
    _fields_ = [("param", P), ("grad", P), ("exp_avg", P), ("exp_avg_sq", P), ("n", I64), ("step", P),
                ("lr", C.c_double), ("beta1", C.c_double), ("beta2", C.c_double), ("eps", C.c_double), ("grad_scale", F32), ("lr_dev", P)]


class GatherBatchArgs(C.Structure):
    _fields_ = [("all_rays", P), ("rays_stride", I64), ("all_rgbs", P), ("rgbs_stride", I64), ("all_ts", P), ("perm", P),
                ("n_rows", I64), ("first", I64), ("batch", I64), ("rays_out", P), ("rgbs_out", P), ("ts_out", P), ("idx_out", P)]


class LossArgs(C.Structure):
    _fields_ = [("out", P), ("gt_rgb", P), ("n_rays", I64), ("mode", I32), ("loss", P), ("g_out", P), ("partials", P)]


class MarchArgs(C.Structure):
    _fields_ = [("origins", P), ("origins_stride", I64), ("viewdirs", P), ("viewdirs_stride", I64), ("jitter", P), ("n_rays", I64),
                ("aabb", F32 * 6), ("near_plane", F32), ("far_plane", F32), ("step", F32), ("max_per_ray", I32),
                ("counts", P), ("t0_out", P), ("t_max_out", P), ("ray_offsets", P), ("ray_indices", P), ("t_starts", P), ("t_ends", P)]


class UtmPointsArgs(C.Structure):
    _fields_ = [("rays", P), ("rays_stride", I64), ("depth", P), ("depth_stride", I64), ("n_rays", I64),
                ("scene_scale", C.c_double * 3), ("scene_offset", C.c_double * 3), ("easts", P), ("norths", P), ("alts", P),
                ("alt_f32", P)]


class DsmArgs(C.Structure):
    _fields_ = [("easts", P), ("norths", P), ("alts", P), ("depth", P), ("depth_stride", I64), ("n_points", I64),
                ("xoff", C.c_double), ("yoff", C.c_double), ("resolution", C.c_double), ("xsize", I32), ("ysize", I32),
                ("radius", I32), ("sigma", C.c_double), ("negative_north_shift", C.c_double), ("acc", P), ("dsm", P)]


# every symbol include/eonerf_b200.h declares: name -> (restype, argtypes)
_ARGS = lambda T: [C.POINTER(T), P]
SYMBOLS = {
    "eonerf_abi_version": (C.c_int, []),
    "eonerf_last_error": (C.c_char_p, []),
    "eonerf_check_device": (C.c_int, []),
    "eonerf_launch_count": (I64, [I32]),
    "eonerf_profile_enable": (C.c_int, [I32]),
    "eonerf_profile_read": (C.c_int, [C.POINTER(Profile), I32]),
    "eonerf_sample_scratch_bytes": (I64, [I64]),
    "eonerf_sample_compact": (C.c_int, _ARGS(SampleArgs)),
    "eonerf_pack_info": (C.c_int, [P, I64, I64, P, P]),
    "eonerf_set_last_t_end": (C.c_int, [P, P, I64, F32, P]),
    "eonerf_weights_fwd": (C.c_int, _ARGS(WeightsFwdArgs)),
    "eonerf_weights_bwd": (C.c_int, _ARGS(WeightsBwdArgs)),
    "eonerf_accumulate_fwd": (C.c_int, _ARGS(AccumFwdArgs)),
    "eonerf_accumulate_bwd": (C.c_int, _ARGS(AccumBwdArgs)),
    "eonerf_composite_fwd": (C.c_int, _ARGS(CompositeFwdArgs)),
    "eonerf_composite_bwd": (C.c_int, _ARGS(CompositeBwdArgs)),
    "eonerf_sun_rays": (C.c_int, _ARGS(SunRaysArgs)),
    "eonerf_shadow_fwd": (C.c_int, _ARGS(ShadowFwdArgs)),
    "eonerf_shadow_bwd": (C.c_int, _ARGS(ShadowBwdArgs)),
    "eonerf_sun_origin_bwd": (C.c_int, _ARGS(SunOriginBwdArgs)),
    "eonerf_epilogue_fwd": (C.c_int, _ARGS(EpilogueFwdArgs)),
    "eonerf_epilogue_bwd": (C.c_int, _ARGS(EpilogueBwdArgs)),
    "eonerf_field_prepared_bytes": (I64, [I32, I32, I64]),
    "eonerf_field_stash_bytes": (I64, [I32, I32, I64, I32]),
    "eonerf_field_scratch_bytes": (I64, [I32, I32, I64, I64]),
    "eonerf_field_prepare": (C.c_int, [I32, I32, C.POINTER(FieldParams), P, P]),
    "eonerf_field_fwd": (C.c_int, _ARGS(FieldFwdArgs)),
    "eonerf_field_bwd": (C.c_int, _ARGS(FieldBwdArgs)),
    "eonerf_ambient_fwd": (C.c_int, _ARGS(AmbientFwdArgs)),
    "eonerf_ambient_bwd": (C.c_int, _ARGS(AmbientBwdArgs)),
    "eonerf_linear_fwd": (C.c_int, _ARGS(LinearArgs)),
    "eonerf_linear_dw": (C.c_int, _ARGS(DwArgs)),
    "eonerf_adam_step": (C.c_int, _ARGS(AdamArgs)),
    "eonerf_gather_batch": (C.c_int, _ARGS(GatherBatchArgs)),
    "eonerf_loss_partials": (I64, [I64]),
    "eonerf_loss_fwd_bwd": (C.c_int, _ARGS(LossArgs)),
    "eonerf_march_count": (C.c_int, _ARGS(MarchArgs)),
    "eonerf_march_write": (C.c_int, _ARGS(MarchArgs)),
    "eonerf_utm_points": (C.c_int, _ARGS(UtmPointsArgs)),
    "eonerf_dsm_rasterize": (C.c_int, _ARGS(DsmArgs)),
}

_lib = None


def lib():
    """Load the library (once).  Raises if it has not been built: there is no fallback path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -m eonerf_code_b200.build` (no CPU fallback exists)")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        v = l.eonerf_abi_version()
        if v != ABI_VERSION:
            raise RuntimeError(f"libeonerf_b200.so has ABI {v}, the binding expects {ABI_VERSION}: rebuild")
        _lib = l
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().eonerf_last_error().decode()
        raise RuntimeError(f"eonerf_b200 {what} failed ({rc}): {msg}")


_device_ok = set()


def require_device(index=None):
    """Raises unless the CUDA device (default: the current one) is sm_100; checked once per device."""
    import torch
    if index is None:
        index = torch.cuda.current_device()
    if index not in _device_ok:
        with torch.cuda.device(index):
            check(lib().eonerf_check_device(), "check_device")
        _device_ok.add(index)


def call(name, args, stream):
    """Invoke `int eonerf_<name>(const Args*, stream)`."""
    check(getattr(lib(), "eonerf_" + name)(C.byref(args), stream), name)
