"""Tensor-level wrappers over the C ABI (include/eonerf_b200.h) and the autograd functions built on them.

PyTorch is plumbing here: it owns device memory, streams and the autograd graph; every computation on the
hot path is a hand-written sm_100a kernel behind `libeonerf_b200.so`.  Nothing in this module has a CPU
or eager-PyTorch fallback — a CPU tensor or a missing library raises.
"""
import ctypes as C
from collections import OrderedDict

import torch

from . import _capi as K

BETA_MIN = 0.05          # /root/reference/radiance_fields/eonerf.py:87


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    return None if t is None else t.data_ptr()


def _f32(t):
    if t.dtype != torch.float32:
        t = t.float()
    return t


def _rows(t):
    """[B,C] fp32 view with unit inner stride (a column slice of the [B,11] ray table qualifies) -> (tensor, row stride)."""
    t = _f32(t)
    if t.dim() == 1:
        t = t[:, None]
    if t.stride(1) != 1 and t.shape[1] != 1:
        t = t.contiguous()
    return t, t.stride(0)


def _need_cuda(*ts):
    """Every tensor on ONE sm_100 device, and that device current (the kernels launch on its current stream)."""
    dev = None
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("eonerf_code_b200 runs on sm_100 GPUs only: got a CPU tensor (there is no CPU fallback)")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError(f"eonerf_code_b200: tensors on different devices ({dev} and {t.device})")
    if dev is not None and dev.index != torch.cuda.current_device():
        raise RuntimeError(f"eonerf_code_b200: tensors live on {dev} but the current device is cuda:{torch.cuda.current_device()} "
                           "(internal error: the entry points switch to the tensors' device, see on_tensor_device)")
    K.require_device()


def tensor_device(*objs):
    """Device of the first CUDA tensor found in objs (namedtuples / lists / dicts are searched one level deep)."""
    for o in objs:
        if isinstance(o, torch.Tensor):
            if o.is_cuda:
                return o.device
        elif isinstance(o, (tuple, list)):
            for x in o:
                if isinstance(x, torch.Tensor) and x.is_cuda:
                    return x.device
        elif isinstance(o, torch.nn.Module):
            for x in o.parameters():
                if x.is_cuda:
                    return x.device
                break
        elif isinstance(o, FieldEngine):
            for x in o.named.values():
                if x.is_cuda:
                    return x.device
                break
    return None


def on_tensor_device(fn):
    """Run fn with the CUDA device of its first CUDA tensor argument made current.  The reference picks its GPU with
    `device=f"cuda:{args.gpu_id}"` and never calls set_device (train_eonerf.py:39-41): as a drop-in with gpu_id != 0 the
    kernels must be launched on that device's stream, not on device 0's."""
    import functools

    @functools.wraps(fn)
    def wrapped(*a, **k):
        dev = tensor_device(*a, *k.values())
        if dev is None or dev.index == torch.cuda.current_device():
            return fn(*a, **k)
        with torch.cuda.device(dev):
            return fn(*a, **k)
    return wrapped


# one-pass sampler (decoupled look-back; csrc/sampling.cu); EONERF_SAMPLER=3pass selects the count / scan / scatter form (A/B)
import os as _os
ONE_PASS_SAMPLER = _os.environ.get("EONERF_SAMPLER", "1pass") != "3pass"

_Z_STEPS = {}


def z_steps_for(n, device):
    """torch.linspace(0, 1, n) on the device, as /root/reference/sat_rendering.py:67 computes it."""
    key = (n, str(device))
    if key not in _Z_STEPS:
        _Z_STEPS[key] = torch.linspace(0, 1, n, device=device)
    return _Z_STEPS[key]


# ------------------------------------------------------------------------------------------------
# sampling
# ------------------------------------------------------------------------------------------------
@on_tensor_device
def sample_compact(origins, viewdirs, near, u, z_steps=None, redraw_of=None):
    """Stratified samples + cube mask + compaction (sat_rendering.py:56-84, :10-16).
    Returns worst-case-sized buffers (ray_indices, t_starts, t_ends), pts_per_ray[B] fp32, ray_offsets[B+1], stats[2]
    (stats = [P, number of empty rays], still on the device).
    `redraw_of` = the tuple returned by a first call: the sync-free form of the reference's "some ray kept no sample ->
    draw again" (sat_rendering.py:259-262).  The kernels then run only if that first draw left an empty ray (decided on
    the device) and overwrite its packed arrays, offsets and P in place; pts_per_ray and the empty-ray count keep the
    first draw's values, as in the reference."""
    _need_cuda(origins, viewdirs, u)
    B, n = u.shape
    dev = origins.device
    o, os_ = _rows(origins)
    d, ds_ = _rows(viewdirs)
    u = _f32(u).contiguous()
    if z_steps is None:
        z_steps = z_steps_for(n, dev)
    cap = B * (n - 1)
    if redraw_of is not None:
        ri, ts, te, ppr, offs, stats = redraw_of
    else:
        ri = torch.empty(cap, dtype=torch.int64, device=dev)
        ts = torch.empty(cap, dtype=torch.float32, device=dev)
        te = torch.empty(cap, dtype=torch.float32, device=dev)
        ppr = torch.empty(B, dtype=torch.float32, device=dev)
        offs = torch.empty(B + 1, dtype=torch.int64, device=dev)
        stats = torch.empty(2, dtype=torch.int64, device=dev)
    a = K.SampleArgs()
    a.origins, a.origins_stride, a.viewdirs, a.viewdirs_stride = _p(o), os_, _p(d), ds_
    if near is not None:
        nr, ns_ = _rows(near)
        a.near, a.near_stride = _p(nr), ns_
    a.u, a.z_steps, a.n_rays, a.n_samples = _p(u), _p(z_steps), B, n
    a.ray_indices, a.t_starts, a.t_ends = _p(ri), _p(ts), _p(te)
    a.pts_per_ray, a.ray_offsets, a.stats = _p(ppr), _p(offs), _p(stats)
    if redraw_of is not None:
        a.run_if = stats.data_ptr() + 8
    if ONE_PASS_SAMPLER:
        scratch = torch.empty(K.lib().eonerf_sample_scratch_bytes(B) // 8, dtype=torch.int64, device=dev)
        a.scratch = _p(scratch)
    K.call("sample_compact", a, _stream())
    return ri, ts, te, ppr, offs, stats


@on_tensor_device
def pack_info(ray_indices, n_rays):
    """ray_offsets[B+1] from sorted ray_indices (nerfacc pack_info)."""
    _need_cuda(ray_indices)
    ray_indices = ray_indices.contiguous()
    offs = torch.empty(n_rays + 1, dtype=torch.int64, device=ray_indices.device)
    K.check(K.lib().eonerf_pack_info(_p(ray_indices), ray_indices.numel(), n_rays, _p(offs), _stream()), "pack_info")
    return offs


@on_tensor_device
def set_last_t_end(t_ends, ray_offsets, value=1e10):
    """In place: t_ends[last sample of each ray] = 1e10 (eonerf.py:218-220)."""
    n_rays = ray_offsets.numel() - 1
    K.check(K.lib().eonerf_set_last_t_end(_p(t_ends), _p(ray_offsets), n_rays, value, _stream()), "set_last_t_end")


@on_tensor_device
def march_aabb(origins, viewdirs, aabb, near_plane, far_plane, step, jitter=None, max_per_ray=4096, static=False):
    """Uniform marching inside the scene box (csrc/march.cu; BASELINE configs[1]).  -> (ray_indices i64[P], t_starts[P],
    t_ends[P], ray_offsets i64[B+1]).  One host read (the total P), like the reference's samplers.
    static=True: no host read (CUDA-graph capturable): the outputs have the worst-case capacity B * min(max_per_ray, box diagonal / step + 2)
    and a fifth return value holds the live count P as an int64[1] device tensor (a view of ray_offsets[B])."""
    _need_cuda(origins, viewdirs, jitter)
    o, os_ = _rows(origins)
    d, ds_ = _rows(viewdirs)
    B, dev = o.shape[0], o.device
    a = K.MarchArgs()
    a.origins, a.origins_stride, a.viewdirs, a.viewdirs_stride, a.n_rays = _p(o), os_, _p(d), ds_, B
    if jitter is not None:
        jitter = _f32(jitter).contiguous()
        a.jitter = _p(jitter)
    box = [float(x) for x in torch.as_tensor(aabb).flatten().tolist()]
    for i, v in enumerate(box):
        a.aabb[i] = v
    a.near_plane, a.far_plane, a.step, a.max_per_ray = float(near_plane), float(min(far_plane, 3.0e38)), float(step), int(max_per_ray)
    counts = torch.empty(B, dtype=torch.int64, device=dev)
    t0, tmax = torch.empty(B, dtype=torch.float32, device=dev), torch.empty(B, dtype=torch.float32, device=dev)
    a.counts, a.t0_out, a.t_max_out = _p(counts), _p(t0), _p(tmax)
    K.call("march_count", a, _stream())
    offs = torch.zeros(B + 1, dtype=torch.int64, device=dev)
    torch.cumsum(counts, 0, out=offs[1:])
    if static:
        diag = sum((box[3 + k] - box[k]) ** 2 for k in range(3)) ** 0.5
        P = B * min(int(max_per_ray), int(diag / float(step)) + 2)
    else:
        P = int(offs[-1])
    ri = torch.empty(P, dtype=torch.int64, device=dev)
    ts, te = torch.empty(P, dtype=torch.float32, device=dev), torch.empty(P, dtype=torch.float32, device=dev)
    a.ray_offsets, a.ray_indices, a.t_starts, a.t_ends = _p(offs), _p(ri), _p(ts), _p(te)
    K.call("march_write", a, _stream())
    if static:
        return ri, ts, te, offs, offs[B:]
    return ri, ts, te, offs


# ------------------------------------------------------------------------------------------------
# nerfacc v0.5.2 operator trio
# ------------------------------------------------------------------------------------------------
class _WeightsFn(torch.autograd.Function):
    @staticmethod
    @on_tensor_device
    def forward(ctx, t_starts, t_ends, sigmas, ray_offsets):
        _need_cuda(t_starts, t_ends, sigmas)
        ts, te, sg = _f32(t_starts).contiguous(), _f32(t_ends).contiguous(), _f32(sigmas).contiguous()
        n = ts.numel()
        w, T, al = torch.empty_like(ts), torch.empty_like(ts), torch.empty_like(ts)
        a = K.WeightsFwdArgs(_p(ts), _p(te), _p(sg), _p(ray_offsets), ray_offsets.numel() - 1, n, _p(w), _p(T), _p(al))
        K.call("weights_fwd", a, _stream())
        ctx.save_for_backward(ts, te, sg, ray_offsets)
        return w, T, al

    @staticmethod
    @on_tensor_device
    def backward(ctx, gw, gT, ga):
        ts, te, sg, offs = ctx.saved_tensors
        c = lambda g: None if g is None else _f32(g).contiguous()
        gw, gT, ga = c(gw), c(gT), c(ga)
        gs = torch.zeros_like(sg)
        a = K.WeightsBwdArgs(_p(ts), _p(te), _p(sg), _p(offs), offs.numel() - 1, ts.numel(), _p(gw), _p(gT), _p(ga), _p(gs))
        K.call("weights_bwd", a, _stream())
        return None, None, gs, None


class _AccumFn(torch.autograd.Function):
    @staticmethod
    @on_tensor_device
    def forward(ctx, weights, values, ray_offsets):
        _need_cuda(weights)
        w = _f32(weights).contiguous()
        v = None if values is None else _f32(values).contiguous()
        Cn = 1 if v is None else v.shape[-1]
        B = ray_offsets.numel() - 1
        out = torch.empty(B, Cn, dtype=torch.float32, device=w.device)
        a = K.AccumFwdArgs(_p(w), _p(v), Cn, _p(ray_offsets), B, w.numel(), _p(out))
        K.call("accumulate_fwd", a, _stream())
        ctx.save_for_backward(w, v, ray_offsets)
        ctx.has_values = v is not None
        return out

    @staticmethod
    @on_tensor_device
    def backward(ctx, g_out):
        w, v, offs = ctx.saved_tensors
        g_out = _f32(g_out).contiguous()
        gw = torch.zeros_like(w)
        gv = torch.zeros_like(v) if ctx.has_values else None
        Cn = g_out.shape[-1]
        a = K.AccumBwdArgs(_p(w), _p(v), Cn, _p(offs), offs.numel() - 1, w.numel(), _p(g_out), _p(gw), _p(gv))
        K.call("accumulate_bwd", a, _stream())
        return gw, gv, None


# ------------------------------------------------------------------------------------------------
# field engine: parameter marshalling + prepared (operand-layout) weights
# ------------------------------------------------------------------------------------------------
EONERF_PARAM_ORDER = (
    ["transient_encoder.weight", "radiometricT_enc.weight"]
    + [f"base_mlp.hidden_layers.{i}.{k}" for i in range(8) for k in ("weight", "bias")]
    + [f"{m}.{k}" for m in ("sigma_layer.output_layer", "bottleneck_layer.output_layer", "albedo_mlp.hidden_layers.0",
                           "albedo_mlp.output_layer") for k in ("weight", "bias")]
    + [f"transient_mlp.hidden_layers.{i}.{k}" for i in range(4) for k in ("weight", "bias")]
    + [f"{m}.{k}" for m in ("transient_scalar.output_layer", "transient_beta.output_layer", "ambient_mlp.hidden_layers.0",
                           "ambient_mlp.output_layer") for k in ("weight", "bias")])

VANILLA_PARAM_ORDER = (
    [f"mlp.base.hidden_layers.{i}.{k}" for i in range(8) for k in ("weight", "bias")]
    + [f"mlp.{m}.{k}" for m in ("sigma_layer.output_layer", "bottleneck_layer.output_layer", "rgb_layer.hidden_layers.0",
                                "rgb_layer.output_layer") for k in ("weight", "bias")])


def _fill_field_params(st, get, field, n_images):
    """st: K.FieldParams; get(name) -> device pointer (int) of the fp32 tensor registered under the reference's key."""
    if field == K.FIELD_EONERF:
        base, pre = "base_mlp", ""
        st.head0_w, st.head0_b = get("albedo_mlp.hidden_layers.0.weight"), get("albedo_mlp.hidden_layers.0.bias")
        st.head1_w, st.head1_b = get("albedo_mlp.output_layer.weight"), get("albedo_mlp.output_layer.bias")
        for i in range(4):
            st.trans_w[i] = get(f"transient_mlp.hidden_layers.{i}.weight")
            st.trans_b[i] = get(f"transient_mlp.hidden_layers.{i}.bias")
        st.ts_w, st.ts_b = get("transient_scalar.output_layer.weight"), get("transient_scalar.output_layer.bias")
        st.tb_w, st.tb_b = get("transient_beta.output_layer.weight"), get("transient_beta.output_layer.bias")
        st.transient_emb = get("transient_encoder.weight")
    else:
        base, pre = "mlp.base", "mlp."
        st.head0_w, st.head0_b = get("mlp.rgb_layer.hidden_layers.0.weight"), get("mlp.rgb_layer.hidden_layers.0.bias")
        st.head1_w, st.head1_b = get("mlp.rgb_layer.output_layer.weight"), get("mlp.rgb_layer.output_layer.bias")
    for i in range(8):
        st.trunk_w[i] = get(f"{base}.hidden_layers.{i}.weight")
        st.trunk_b[i] = get(f"{base}.hidden_layers.{i}.bias")
    st.sigma_w, st.sigma_b = get(pre + "sigma_layer.output_layer.weight"), get(pre + "sigma_layer.output_layer.bias")
    st.bott_w, st.bott_b = get(pre + "bottleneck_layer.output_layer.weight"), get(pre + "bottleneck_layer.output_layer.bias")
    st.n_images = n_images
    return st


class FieldEngine:
    """Marshals a module's fp32 master parameters (reference state_dict names) into the C ABI, keeps the
    operand-layout copy (`prepared`) fresh, and owns the flat gradient buffer."""

    def __init__(self, named_params, field=K.FIELD_EONERF, precision=K.PREC_BF16, n_images=0):
        self.field, self.precision, self.n_images = field, precision, n_images
        self.order = EONERF_PARAM_ORDER if field == K.FIELD_EONERF else VANILLA_PARAM_ORDER
        self.named = OrderedDict((k, named_params[k]) for k in self.order if k in named_params)
        self.has_radiometric = "radiometricT_enc.weight" in self.named
        self._prepared = None
        self._prepared_key = None
        self.grad_sync = None        # optional callable(flat_grad) run at the end of backward (data parallel all-reduce)
        self.grad_sink = None        # optional {name: fp32 tensor}: backward accumulates straight into these (see use_grad_sink)
        self._sink_struct = None

    # --- parameters -------------------------------------------------------------------------
    def tensors(self):
        return list(self.named.values())

    def _check(self):
        for k, t in self.named.items():
            if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
                raise RuntimeError(f"parameter {k} must be a contiguous fp32 CUDA tensor (got {t.dtype}, {t.device})")

    def params_struct(self):
        self._check()
        return _fill_field_params(K.FieldParams(), lambda n: self.named[n].data_ptr(), self.field, self.n_images)

    def refresh_prepared(self):
        """Rebuild the operand-layout weights unconditionally.  A captured step starts with this: graph replays change the
        parameters without bumping their Python-side versions."""
        self._prepared_key = None
        return self.prepared()

    @on_tensor_device
    def prepared(self):
        """Operand-layout weights, rebuilt when any parameter changed (optimizer steps bump ._version)."""
        key = (self.precision,) + tuple((t.data_ptr(), t._version) for t in self.named.values())
        if key != self._prepared_key:
            _need_cuda(*self.named.values())
            nbytes = K.lib().eonerf_field_prepared_bytes(self.field, self.precision, self.n_images)
            dev = next(iter(self.named.values())).device
            if self._prepared is None or self._prepared.numel() != nbytes or self._prepared.device != dev:
                self._prepared = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            ps = self.params_struct()
            K.check(K.lib().eonerf_field_prepare(self.field, self.precision, C.byref(ps), _p(self._prepared), _stream()),
                    "field_prepare")
            self._prepared_key = key
        return self._prepared

    # --- gradients --------------------------------------------------------------------------
    def use_grad_sink(self, sink):
        """Training fast path: every backward kernel accumulates (atomic adds) directly into the caller's gradient tensors
        `sink[name]` (same shapes as the parameters, typically the parameters' .grad views of one flat buffer) and the
        autograd functions return no parameter gradients, instead of one fresh flat buffer per pass that autograd then
        adds tensor by tensor (~90 tiny kernels per step).  The caller zeroes the sink before the step.  None: off."""
        if sink is not None:
            for k, t in self.named.items():
                g = sink[k]
                if not (g.is_cuda and g.dtype == torch.float32 and g.is_contiguous() and g.shape == t.shape):
                    raise RuntimeError(f"gradient sink for {k} must be a contiguous fp32 CUDA tensor of shape {tuple(t.shape)}")
            sink = OrderedDict((k, sink[k]) for k in self.named)
            self._sink_struct = _fill_field_params(K.FieldParams(), lambda n: sink[n].data_ptr(), self.field, self.n_images)
        self.grad_sink = sink

    def grads_for_backward(self):
        """(flat buffer or None, {name: tensor}, K.FieldParams, direct).  direct=True: the sink is in use."""
        if self.grad_sink is not None:
            return None, self.grad_sink, self._sink_struct, True
        return self.new_grads() + (False,)

    def new_grads(self):
        """(flat fp32 zero buffer, {name: view}, K.FieldParams of the views)."""
        dev = next(iter(self.named.values())).device
        total = sum(t.numel() for t in self.named.values())
        flat = torch.zeros(total, dtype=torch.float32, device=dev)
        views, off = OrderedDict(), 0
        for k, t in self.named.items():
            views[k] = flat[off:off + t.numel()].view_as(t)
            off += t.numel()
        gs = _fill_field_params(K.FieldParams(), lambda n: views[n].data_ptr(), self.field, self.n_images)
        return flat, views, gs

    # --- raw calls --------------------------------------------------------------------------
    def stash_bytes(self, n, density_only):
        return K.lib().eonerf_field_stash_bytes(self.field, self.precision, n, int(density_only))

    def scratch_bytes(self, n):
        return K.lib().eonerf_field_scratch_bytes(self.field, self.precision, n, self.n_images)

    @on_tensor_device
    def fwd(self, n, density_only, x=None, rays=None, img_idx=None, cond_dirs=None, want_z=False, keep=True, n_dev=None,
            cond_dirs_per_ray=False):
        """rays = (origins, viewdirs, ray_indices, t_starts, t_ends).  Returns dict of outputs + stash.
        keep=False (inference): the fused mode keeps no activations at all; the layered modes still need the buffer.
        n_dev: int64[1] device tensor holding the live sample count (n is then the capacity): no host read of P."""
        if n_dev is not None and self.precision != K.PREC_BF16_FUSED:
            raise RuntimeError("device-side sample counts need precision='bf16_fused'")
        dev = next(iter(self.named.values())).device
        f32 = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
        no_stash = not keep and self.precision == K.PREC_BF16_FUSED
        out = dict(sigma=f32(n), stash=None if no_stash else torch.empty(self.stash_bytes(n, density_only), dtype=torch.uint8, device=dev))
        a = K.FieldFwdArgs()
        ps = self.params_struct()
        a.field, a.precision, a.params, a.prepared, a.n_pts = self.field, self.precision, C.pointer(ps), _p(self.prepared()), n
        keep = [ps]
        if x is not None:
            x = _f32(x).contiguous()
            a.x = _p(x)
            keep.append(x)
        if rays is not None:
            o, d, ri, ts, te = rays
            o, os_ = _rows(o)
            d, ds_ = _rows(d)
            a.origins, a.origins_stride, a.viewdirs, a.viewdirs_stride = _p(o), os_, _p(d), ds_
            a.ray_indices, a.t_starts, a.t_ends = _p(ri), _p(ts), _p(te)
            keep += [o, d]
            if want_z:
                out["z_mid"] = f32(n)
                a.z_mid = _p(out["z_mid"])
        if (x is None) == (rays is None):
            raise RuntimeError("field fwd: give exactly one of x / rays")
        if img_idx is not None:
            ii = img_idx if img_idx.dim() == 2 else img_idx[:, None]
            if ii.dtype != torch.int64:
                ii = ii.long()
            a.img_idx, a.img_idx_stride = _p(ii), ii.stride(0)
            keep.append(ii)
        if cond_dirs is not None:
            cd, cs_ = _rows(cond_dirs)
            a.cond_dirs, a.cond_dirs_stride, a.cond_dirs_per_ray = _p(cd), cs_, int(bool(cond_dirs_per_ray))
            keep.append(cd)
            if self.precision == K.PREC_BF16_FUSED and not density_only:
                # the view-direction term of the rgb hidden layer as one bias row per conditioning row (csrc/field_fused.cu)
                out["dir_bias"] = f32(cd.shape[0], 128)
                a.dir_bias, a.n_cond = _p(out["dir_bias"]), cd.shape[0]
                out["cond"] = (cd, cs_)
        a.density_only, a.stash, a.sigma = int(density_only), _p(out["stash"]), _p(out["sigma"])
        a.n_pts_dev = _p(n_dev)
        if not density_only:
            out["rgb"] = f32(n, 3)
            a.rgb = _p(out["rgb"])
            if self.field == K.FIELD_EONERF:
                out["transient_s"], out["transient_beta"] = f32(n), f32(n)
                a.transient_s, a.transient_beta = _p(out["transient_s"]), _p(out["transient_beta"])
        K.call("field_fwd", a, _stream())
        return out

    @on_tensor_device
    def bwd(self, n, density_only, fwd_out, g_sigma=None, g_rgb=None, g_ts=None, g_tb=None, grads_struct=None, want_gx=False,
            n_dev=None):
        dev = fwd_out["sigma"].device
        scratch = torch.empty(self.scratch_bytes(n), dtype=torch.uint8, device=dev)
        gx = torch.empty(n, 3, dtype=torch.float32, device=dev) if want_gx else None
        a = K.FieldBwdArgs()
        ps = self.params_struct()
        a.field, a.precision, a.params, a.prepared, a.n_pts = self.field, self.precision, C.pointer(ps), _p(self.prepared()), n
        a.density_only, a.stash, a.scratch = int(density_only), _p(fwd_out["stash"]), _p(scratch)
        a.sigma, a.rgb = _p(fwd_out["sigma"]), _p(fwd_out.get("rgb"))
        a.transient_s, a.transient_beta = _p(fwd_out.get("transient_s")), _p(fwd_out.get("transient_beta"))
        a.g_sigma, a.g_rgb, a.g_transient_s, a.g_transient_beta = _p(g_sigma), _p(g_rgb), _p(g_ts), _p(g_tb)
        if grads_struct is not None:
            a.grads = C.pointer(grads_struct)
        a.g_x = _p(gx)
        a.n_pts_dev = _p(n_dev)
        if fwd_out.get("cond") is not None:
            cd, cs_ = fwd_out["cond"]
            a.cond_dirs, a.cond_dirs_stride, a.n_cond = _p(cd), cs_, cd.shape[0]
        K.call("field_bwd", a, _stream())
        return gx

    # --- per-ray ambient MLP (eonerf.py:163-164) --------------------------------------------
    @on_tensor_device
    def ambient_fwd(self, sundirs):
        sd, ss_ = _rows(sundirs)
        B = sd.shape[0]
        stash = torch.empty(B * 160, dtype=torch.float32, device=sd.device)
        amb = torch.empty(B, 3, dtype=torch.float32, device=sd.device)
        g = lambda n: self.named[n].data_ptr()
        a = K.AmbientFwdArgs(_p(sd), ss_, B, g("ambient_mlp.hidden_layers.0.weight"), g("ambient_mlp.hidden_layers.0.bias"),
                             g("ambient_mlp.output_layer.weight"), g("ambient_mlp.output_layer.bias"), _p(stash), _p(amb))
        K.call("ambient_fwd", a, _stream())
        return amb, stash

    @on_tensor_device
    def ambient_bwd(self, amb, stash, g_amb, grad_views):
        B = amb.shape[0]
        scratch = torch.empty(B * 136, dtype=torch.float32, device=amb.device)
        g = lambda n: self.named[n].data_ptr()
        v = lambda n: grad_views[n].data_ptr()
        a = K.AmbientBwdArgs(B, g("ambient_mlp.hidden_layers.0.weight"), g("ambient_mlp.output_layer.weight"), _p(stash),
                             _p(scratch), _p(amb), _p(g_amb), v("ambient_mlp.hidden_layers.0.weight"),
                             v("ambient_mlp.hidden_layers.0.bias"), v("ambient_mlp.output_layer.weight"),
                             v("ambient_mlp.output_layer.bias"))
        K.call("ambient_bwd", a, _stream())


def _keep_for_backward(out):
    """The forward dict as the backward needs it, with every tensor replaced by a detached alias (same storage).  The tensors a
    Function RETURNS get `grad_fn = ctx` after forward(): keeping those very objects on ctx would close a reference cycle
    (ctx -> dict -> output -> grad_fn -> ctx) that only Python's cycle collector frees — one whole stash (6 KB per sample) leaked per
    step until the next collection."""
    return {k: (v.detach() if torch.is_tensor(v) else v) for k, v in out.items()}


def _grads_tuple(engine, views, params, direct=False):
    if direct:                      # already accumulated into the sink
        return (None,) * len(params)
    return tuple(views[k] if p.requires_grad else None for (k, _), p in zip(engine.named.items(), params))


class _FieldFn(torch.autograd.Function):
    """EONerfMLP.forward / query_density / VanillaNeRFRadianceField.forward on explicit positions."""

    @staticmethod
    @on_tensor_device
    def forward(ctx, grad_on, engine, density_only, x, img_idx, cond_dirs, *params):
        """grad_on: torch.is_grad_enabled() at the call site (inside forward() autograd always reports it off; under
        torch.no_grad() nothing is stashed)."""
        _need_cuda(x)
        n = x.shape[0]
        out = engine.fwd(n, density_only, x=x, img_idx=img_idx, cond_dirs=cond_dirs, keep=grad_on and any(ctx.needs_input_grad))
        ctx.engine, ctx.density_only, ctx.n, ctx.out = engine, density_only, n, _keep_for_backward(out)
        ctx.x_needs_grad = x.requires_grad
        ctx.params = params
        if density_only:
            return out["sigma"][:, None]
        if engine.field == K.FIELD_EONERF:
            return out["sigma"][:, None], out["rgb"], out["transient_s"][:, None], out["transient_beta"][:, None]
        return out["sigma"][:, None], out["rgb"]

    @staticmethod
    @on_tensor_device
    def backward(ctx, *gs):
        e = ctx.engine
        c = lambda g: None if g is None else _f32(g).contiguous()
        gs = [c(g) for g in gs] + [None] * 4
        flat, views, gstruct, direct = e.grads_for_backward()
        gx = e.bwd(ctx.n, ctx.density_only, ctx.out, g_sigma=gs[0], g_rgb=gs[1], g_ts=gs[2], g_tb=gs[3],
                   grads_struct=gstruct, want_gx=ctx.x_needs_grad)
        ctx.out = None
        if e.grad_sync is not None and not direct:
            e.grad_sync(flat)
        return (None, None, None, gx, None, None) + _grads_tuple(e, views, ctx.params, direct)


class _VanillaRaysFn(torch.autograd.Function):
    """rgb_sigma_fn of nerfacc.rendering for the vanilla field (train_mlp_nerf.py:155-170 via nerfacc examples/utils.py):
    positions x = o[ri] + d[ri] (t_s + t_e) / 2 and the per-sample view directions d[ri] are formed inside the kernels.
    -> (sigma[P,1], rgb[P,3], z_mid[P])."""

    @staticmethod
    @on_tensor_device
    def forward(ctx, grad_on, engine, origins, viewdirs, ri, ts, te, n_dev, *params):
        """n_dev: None, or int64[1] on the device = live sample count (ri / ts / te then have their full capacity)."""
        _need_cuda(origins, viewdirs, ri, ts, te)
        P = ts.numel()
        out = engine.fwd(P, False, rays=(origins, viewdirs, ri.contiguous(), ts, te), cond_dirs=viewdirs, cond_dirs_per_ray=True, want_z=True,
                         keep=grad_on and any(ctx.needs_input_grad), n_dev=n_dev)
        ctx.engine, ctx.n, ctx.out, ctx.params, ctx.n_dev = engine, P, _keep_for_backward(out), params, n_dev
        ctx.mark_non_differentiable(out["z_mid"])
        return out["sigma"][:, None], out["rgb"], out["z_mid"]

    @staticmethod
    @on_tensor_device
    def backward(ctx, g_sigma, g_rgb, _g_z):
        e = ctx.engine
        c = lambda g: None if g is None else _f32(g).contiguous()
        flat, views, gstruct, direct = e.grads_for_backward()
        e.bwd(ctx.n, False, ctx.out, g_sigma=c(g_sigma), g_rgb=c(g_rgb), grads_struct=gstruct, n_dev=ctx.n_dev)
        ctx.out = None
        if e.grad_sync is not None and not direct:
            e.grad_sync(flat)
        return (None,) * 8 + _grads_tuple(e, views, ctx.params, direct)


class _AmbientFn(torch.autograd.Function):
    @staticmethod
    @on_tensor_device
    def forward(ctx, engine, sundirs, *params):
        _need_cuda(sundirs)
        amb, stash = engine.ambient_fwd(sundirs)
        ctx.engine, ctx.amb, ctx.stash, ctx.params = engine, amb.detach(), stash, params     # detached alias: see _keep_for_backward
        return amb

    @staticmethod
    @on_tensor_device
    def backward(ctx, g):
        e = ctx.engine
        flat, views, _, direct = e.grads_for_backward()
        e.ambient_bwd(ctx.amb, ctx.stash, _f32(g).contiguous(), views)
        if e.grad_sync is not None and not direct:
            e.grad_sync(flat)
        return (None, None) + _grads_tuple(e, views, ctx.params, direct)


# ------------------------------------------------------------------------------------------------
# the three stages of one chunk of sat_rendering.py:252-312, one autograd node each
# ------------------------------------------------------------------------------------------------
def _img_idx_2d(img_idx):
    ii = img_idx if img_idx.dim() == 2 else img_idx[:, None]
    return ii if ii.dtype == torch.int64 else ii.long()


class _CameraPassFn(torch.autograd.Function):
    """EONerfMLP.rendering / render_depth (eonerf.py:172-248): gather + positions + MLP + 1e10 last interval +
    weights + the five accumulations, as field_fwd -> ambient_fwd -> composite_fwd.
    Output comp[B,12]: 0:3 albedo, 3 depth, 4 beta(+0.05), 5 transient_s, 6:9 ambient (not yet x0.2), 9 sum(w)."""

    @staticmethod
    @on_tensor_device
    def forward(ctx, grad_on, engine, only_depth, origins, viewdirs, sundirs, img_idx, ri, ts, te, offs, n_dev, *params):
        """n_dev: None, or int64[1] on the device = live sample count P (ri / ts / te then have their full capacity)."""
        _need_cuda(origins, viewdirs, ri, ts, te)
        for t in (ts, te):
            if not (t.dtype == torch.float32 and t.is_contiguous()):
                raise RuntimeError("t_starts / t_ends must be contiguous fp32 (t_ends is updated in place, eonerf.py:220)")
        B, P = origins.shape[0], ts.numel()
        ri = ri.contiguous()
        f = engine.fwd(P, density_only=only_depth, rays=(origins, viewdirs, ri, ts, te),
                       img_idx=None if only_depth else _img_idx_2d(img_idx), want_z=True, keep=grad_on and any(ctx.needs_input_grad),
                       n_dev=n_dev)
        set_last_t_end(te, offs)                                    # after z / positions were taken
        amb = amb_stash = None
        if not only_depth:
            amb, amb_stash = engine.ambient_fwd(sundirs)
        comp = torch.empty(B, K.COMP_COLS, dtype=torch.float32, device=ts.device)
        a = K.CompositeFwdArgs(_p(ts), _p(te), _p(f["z_mid"]), _p(f["sigma"]), _p(f.get("rgb")), _p(f.get("transient_s")),
                               _p(f.get("transient_beta")), _p(amb), _p(offs), B, P, BETA_MIN, _p(comp))
        K.call("composite_fwd", a, _stream())
        ctx.engine, ctx.only_depth, ctx.params = engine, only_depth, params
        ctx.keep = (B, P, ts, te, offs, f, amb, amb_stash, n_dev)
        return comp

    @staticmethod
    @on_tensor_device
    def backward(ctx, g_comp):
        e = ctx.engine
        B, P, ts, te, offs, f, amb, amb_stash, n_dev = ctx.keep
        dev = g_comp.device
        g_comp = _f32(g_comp).contiguous()
        flat, views, gstruct, direct = e.grads_for_backward()
        new = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
        g_sigma = new(P)
        g_alb = g_ts = g_tb = g_amb = None
        if not ctx.only_depth:
            g_alb, g_ts, g_tb, g_amb = new(P, 3), new(P), new(P), new(B, 3)
        a = K.CompositeBwdArgs(_p(ts), _p(te), _p(f["z_mid"]), _p(f["sigma"]), _p(f.get("rgb")), _p(f.get("transient_s")),
                               _p(f.get("transient_beta")), _p(amb), _p(offs), B, P, _p(g_comp), _p(g_sigma), _p(g_alb),
                               _p(g_ts), _p(g_tb), _p(g_amb))
        K.call("composite_bwd", a, _stream())
        if not ctx.only_depth:
            e.ambient_bwd(amb, amb_stash, g_amb, views)
        e.bwd(P, ctx.only_depth, f, g_sigma=g_sigma, g_rgb=g_alb, g_ts=g_ts, g_tb=g_tb, grads_struct=gstruct, n_dev=n_dev)
        ctx.keep = None
        if e.grad_sync is not None and not direct:
            e.grad_sync(flat)
        return (None,) * 12 + _grads_tuple(e, views, ctx.params, direct)


class _SunPassFn(torch.autograd.Function):
    """compute_geometric_shadows (sat_rendering.py:87-118): sun-ray set-up from the rendered depth, sampling, density
    query, transmittance in front of the last kept sample.  Differentiable in `depth` and the parameters."""

    @staticmethod
    @on_tensor_device
    def forward(ctx, grad_on, engine, origins, viewdirs, sundirs, depth, n_samples, u_sun, z_steps, info, static, *params):
        """static=True: no host read of the sun-sample count Q (buffers keep their capacity, kernels read Q on the device)."""
        _need_cuda(origins, viewdirs, sundirs, depth)
        B, dev = origins.shape[0], origins.device
        oo, os_ = _rows(origins)
        dd, ds_ = _rows(viewdirs)
        sd, ss_ = _rows(sundirs)
        dp, dps_ = _rows(depth)
        sun = torch.empty(B, 6, dtype=torch.float32, device=dev)
        a = K.SunRaysArgs(_p(oo), os_, _p(dd), ds_, _p(sd), ss_, _p(dp), dps_, B, _p(sun))
        K.call("sun_rays", a, _stream())
        if u_sun is None:
            u_sun = torch.rand(B, n_samples, dtype=torch.float32, device=dev)      # sat_rendering.py:52 via :93
        ri2, ts2, te2, sc_ppr, offs2, stats2 = sample_compact(sun[:, 0:3], sun[:, 3:6], None, u_sun, z_steps)
        n_dev = stats2[0:1] if static else None
        Q = ts2.numel() if static else int(stats2[0])               # eager: the one host sync of the sun pass
        f2 = engine.fwd(Q, density_only=True, rays=(sun[:, 0:3], sun[:, 3:6], ri2, ts2, te2), keep=grad_on and any(ctx.needs_input_grad),
                        n_dev=n_dev)
        geo = torch.empty(B, 1, dtype=torch.float32, device=dev)
        a = K.ShadowFwdArgs(_p(ts2), _p(te2), _p(f2["sigma"]), _p(offs2), B, Q, _p(geo))
        K.call("shadow_fwd", a, _stream())
        info.update(sc_pts_per_ray=sc_ppr, n_sun_samples=stats2[0] if static else Q, ray_indices=ri2[:Q], t_starts=ts2[:Q],
                    t_ends=te2[:Q], sigma=f2["sigma"], sun_rays=sun)
        ctx.engine, ctx.params = engine, params
        ctx.keep = (B, Q, ts2, te2, offs2, f2, geo.detach(), dd, ds_, n_dev)        # detached alias: see _keep_for_backward
        return geo

    @staticmethod
    @on_tensor_device
    def backward(ctx, g_geo):
        e = ctx.engine
        B, Q, ts2, te2, offs2, f2, geo, dd, ds_, n_dev = ctx.keep
        dev = g_geo.device
        g_geo = _f32(g_geo).contiguous()
        flat, views, gstruct, direct = e.grads_for_backward()
        g_sig2 = torch.empty(Q, dtype=torch.float32, device=dev)
        a = K.ShadowBwdArgs(_p(ts2), _p(te2), _p(offs2), B, Q, _p(geo), _p(g_geo), _p(g_sig2))
        K.call("shadow_bwd", a, _stream())
        gx = e.bwd(Q, True, f2, g_sigma=g_sig2, grads_struct=gstruct, want_gx=True, n_dev=n_dev)
        g_depth = torch.zeros(B, 1, dtype=torch.float32, device=dev)
        a = K.SunOriginBwdArgs(_p(gx), _p(offs2), B, Q, _p(dd), ds_, _p(g_depth), 1)     # chain rule through :90
        K.call("sun_origin_bwd", a, _stream())
        ctx.keep = None
        if e.grad_sync is not None and not direct:
            e.grad_sync(flat)
        return (None, None, None, None, None, g_depth, None, None, None, None, None) + _grads_tuple(e, views, ctx.params, direct)


class _EpilogueFn(torch.autograd.Function):
    """Irradiance model + radiometric normalisation + 21-column packing (sat_rendering.py:265,269-276,288-312)."""

    @staticmethod
    @on_tensor_device
    def forward(ctx, comp, geo, ppr, sc_ppr, img_idx, eval_mode, rad, n_images, rad_sink=None):
        """rad_sink: optional fp32 [n_img,9] tensor the radiometric-embedding gradient is accumulated into directly."""
        _need_cuda(comp)
        B = comp.shape[0]
        comp = _f32(comp).contiguous()
        geo = None if geo is None else _f32(geo).contiguous()
        ii = _img_idx_2d(img_idx)
        out = torch.empty(B, K.OUT_COLS, dtype=torch.float32, device=comp.device)
        a = K.EpilogueFwdArgs(_p(comp), _p(geo), _p(ppr), _p(sc_ppr), _p(ii), ii.stride(0), int(bool(eval_mode)), _p(rad),
                              n_images, B, _p(out))
        K.call("epilogue_fwd", a, _stream())
        ctx.keep = (comp, geo, ii, int(bool(eval_mode)), rad, n_images, rad_sink)
        return out

    @staticmethod
    @on_tensor_device
    def backward(ctx, g_out):
        comp, geo, ii, eval_mode, rad, n_images, rad_sink = ctx.keep
        B, dev = comp.shape[0], comp.device
        g_out = _f32(g_out).contiguous()
        g_comp = torch.empty(B, K.COMP_COLS, dtype=torch.float32, device=dev)
        g_geo = None if geo is None else torch.empty(B, 1, dtype=torch.float32, device=dev)
        g_rad = None if rad is None else (rad_sink if rad_sink is not None else torch.zeros_like(rad))
        a = K.EpilogueBwdArgs(_p(comp), _p(geo), _p(ii), ii.stride(0), eval_mode, _p(rad), n_images, B, _p(g_out),
                              _p(g_comp), _p(g_geo), _p(g_rad))
        K.call("epilogue_bwd", a, _stream())
        return g_comp, g_geo, None, None, None, None, (None if rad_sink is not None else g_rad), None, None
