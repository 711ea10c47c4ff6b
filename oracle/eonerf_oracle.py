"""TEST INFRASTRUCTURE ONLY — CPU (torch fp32) restatement of EO-NeRF's per-ray rendering hot path.

This is the checker the CUDA path is compared against; it is never the thing measured as the product
and nothing under eonerf_code_b200/ imports it.  Only tests/, __graft_entry__.smoke() and bench.py's
`cpu_baseline` / `--impl reference` legs may import it.

Every function cites the reference lines it restates (paths relative to /root/reference).  The
restatement is written in *functional* form (explicit parameter dict, explicit uniforms `u`) so that
the same inputs can be fed to the CUDA kernels; arithmetic that has to match bit for bit (the
sampler) performs the same separately-rounded fp32 torch ops in the same order as the reference.

PARITY STATUS
  * Pinned against the real reference: tests/golden/*.npz were produced in the build container by
    importing the unmodified reference Python from /root/reference (oracle/ref_harness.py,
    oracle/make_golden.py) and tests/test_oracle.py checks this file against them
    (sampling bit-exact; field / compositing / shadows / 12 render outputs / parameter gradients).
  * The reference itself has no tests or golden vectors, and its nerfacc v0.5.2
    dependency is un-vendored: that part is restated in oracle/nerfacc_v052.py ("parity unpinned"
    in isolation, see its header).
"""
import math
from collections import OrderedDict, namedtuple

import torch
import torch.nn.functional as F

from . import nerfacc_v052 as nv

SatRays = namedtuple("SatRays", ("origins", "viewdirs", "sundirs", "img_idx", "t_near", "t_far"))

POS_L = 10   # radiance_fields/eonerf.py:79
VIEW_L = 4   # radiance_fields/eonerf.py:80
BETA_MIN = 0.05  # radiance_fields/eonerf.py:87


def satrays_from_table(rays, ts):
    """datasets/satellite.py:23-26"""
    return SatRays(rays[:, 0:3], rays[:, 3:6], rays[:, 8:11], ts, rays[:, 6:7], rays[:, 7:8])


# --------------------------------------------------------------------------------------------
# parameters
# --------------------------------------------------------------------------------------------
def param_shapes(n_img, radiometric=True):
    """state_dict contract of EONerfMLP (radiance_fields/eonerf.py:69-139),
    in named_parameters() order."""
    s = OrderedDict()
    s["transient_encoder.weight"] = (n_img, 4)
    if radiometric:
        s["radiometricT_enc.weight"] = (n_img, 9)
    ins = [63, 256, 256, 256, 256, 319, 256, 256]
    for i, k in enumerate(ins):
        s[f"base_mlp.hidden_layers.{i}.weight"] = (256, k)
        s[f"base_mlp.hidden_layers.{i}.bias"] = (256,)
    s["sigma_layer.output_layer.weight"] = (1, 256)
    s["sigma_layer.output_layer.bias"] = (1,)
    s["bottleneck_layer.output_layer.weight"] = (256, 256)
    s["bottleneck_layer.output_layer.bias"] = (256,)
    s["albedo_mlp.hidden_layers.0.weight"] = (128, 256)
    s["albedo_mlp.hidden_layers.0.bias"] = (128,)
    s["albedo_mlp.output_layer.weight"] = (3, 128)
    s["albedo_mlp.output_layer.bias"] = (3,)
    for i, k in enumerate([260, 128, 128, 128]):
        s[f"transient_mlp.hidden_layers.{i}.weight"] = (128, k)
        s[f"transient_mlp.hidden_layers.{i}.bias"] = (128,)
    s["transient_scalar.output_layer.weight"] = (1, 128)
    s["transient_scalar.output_layer.bias"] = (1,)
    s["transient_beta.output_layer.weight"] = (1, 128)
    s["transient_beta.output_layer.bias"] = (1,)
    s["ambient_mlp.hidden_layers.0.weight"] = (128, 27)
    s["ambient_mlp.hidden_layers.0.bias"] = (128,)
    s["ambient_mlp.output_layer.weight"] = (3, 128)
    s["ambient_mlp.output_layer.bias"] = (3,)
    return s


def init_params(n_img, seed=0, radiometric=True, bias_scale=0.0, dtype=torch.float32):
    """xavier-uniform weights, zero biases (radiance_fields/mlp.py:22-28,67-85); N(0,1) transient
    embedding; radiometric embedding [1,1,1,0,...] (eonerf.py:91-94).  `bias_scale` > 0 adds small
    random biases so tests exercise the bias paths."""
    g = torch.Generator().manual_seed(seed)
    p = OrderedDict()
    for name, shape in param_shapes(n_img, radiometric).items():
        if name == "transient_encoder.weight":
            t = torch.randn(shape, generator=g)
        elif name == "radiometricT_enc.weight":
            t = torch.cat([torch.ones(n_img, 3), torch.zeros(n_img, 6)], 1)
            if bias_scale > 0:
                t = t + bias_scale * torch.randn(shape, generator=g)
        elif name.endswith(".weight"):
            bound = math.sqrt(6.0 / (shape[0] + shape[1]))
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        else:
            t = bias_scale * torch.randn(shape, generator=g)
        p[name] = t.to(dtype)
    return p


# --------------------------------------------------------------------------------------------
# sampler  (sat_rendering.py:18-22,46-54,56-84) — must be bit-exact
# --------------------------------------------------------------------------------------------
def n_samples_from_step(render_step_size):
    """sat_rendering.py:64"""
    return int(2 / render_step_size)


def stratified_z(near, n_samples, u, z_steps=None):
    """z values per ray before flattening: sat_rendering.py:60-70 + :46-54.  `u` [B,n] replaces
    torch.rand_like.  Returns z [B, n]."""
    if z_steps is None:
        z_steps = torch.linspace(0, 1, n_samples, device=near.device)
    far = near + 2
    z = near * (1 - z_steps) + far * z_steps
    mid = 0.5 * (z[:, :-1] + z[:, 1:])
    upper = torch.cat([mid, z[:, -1:]], -1)
    lower = torch.cat([z[:, :1], mid], -1)
    return lower + (upper - lower) * u


def satnerf_sampling(origins, viewdirs, n_samples, u, near=None, z_steps=None):
    """sat_rendering.py:56-84 with the uniforms passed in.  Returns compacted
    (ray_indices i64[P], t_starts[P], t_ends[P]) plus the dense keep-mask [B, n-1]."""
    # always fp32 (the reference's arithmetic), so that an fp64 run of the rest of the oracle — used by the tests to
    # measure the conditioning of the gradients — sees exactly the same kept samples
    origins, viewdirs, u = origins.float(), viewdirs.float(), u.float()
    near = torch.zeros_like(origins[:, 0:1]) if near is None else near.float()
    z = stratified_z(near, n_samples, u, z_steps)
    n_rays = origins.shape[0]
    t_starts = z[:, :-1].flatten()
    t_ends = (z[:, :-1] + (z[:, 1:] - z[:, :-1])).flatten()
    ray_indices = torch.arange(n_rays, device=origins.device).repeat_interleave(n_samples - 1)
    zm = (t_starts + t_ends)[:, None] / 2.0
    xyz = origins[ray_indices] + viewdirs[ray_indices] * zm
    keep = torch.sum(torch.abs(xyz) >= 1, dim=1) == 0          # :18-22
    return ray_indices[keep], t_starts[keep], t_ends[keep], keep.view(n_rays, n_samples - 1)


def pts_per_ray(ray_indices, n_rays):
    """sat_rendering.py:10-16 (fp32 counts)"""
    return torch.bincount(ray_indices, minlength=n_rays).to(torch.float32)


# --------------------------------------------------------------------------------------------
# field  (radiance_fields/eonerf.py:141-170, radiance_fields/mlp.py:87-101,190-208)
# --------------------------------------------------------------------------------------------
def posenc(x, L):
    """mlp.py:190-208: [x, sin(2^k x) (freq-major, xyz-minor), sin(2^k x + pi/2)]"""
    scales = torch.tensor([2 ** i for i in range(L)], device=x.device)
    xb = (x[..., None, :] * scales[:, None]).reshape(*x.shape[:-1], L * x.shape[-1])
    lat = torch.sin(torch.cat([xb, xb + 0.5 * math.pi], dim=-1))
    return torch.cat([x, lat], dim=-1)


def _lin(p, name, x):
    return F.linear(x, p[name + ".weight"], p[name + ".bias"])


def _r16(t):
    """bf16 round trip (values stay fp32).  Used by the `emulate_bf16` variants below, which restate WHERE the CUDA bf16
    path rounds (operands of every tensor-core GEMM: weights and stored activations) so that the ReLU masks of the two
    computations agree and the backward can be compared tightly."""
    return t.bfloat16().float()


def _lin16(p, name, x, cols=None):
    w = p[name + ".weight"]
    if cols is not None:
        w = w[:, cols[0]:cols[1]]
    return F.linear(x, _r16(w), None)


def trunk(p, x, emulate_bf16=False):
    """posi_encoder + base_mlp (depth 8, width 256, skip concat after layer index 4; mlp.py:87-97)"""
    enc = posenc(x, POS_L)
    if not emulate_bf16:
        h = enc
        for i in range(8):
            h = torch.relu(_lin(p, f"base_mlp.hidden_layers.{i}", h))
            if i == 4:
                h = torch.cat([h, enc], -1)
        return h
    enc = _r16(enc)
    h = enc
    for i in range(8):
        h = _r16(torch.relu(_lin16(p, f"base_mlp.hidden_layers.{i}", h) + p[f"base_mlp.hidden_layers.{i}.bias"]))
        if i == 4:
            h = torch.cat([h, enc], -1)
    return h


def query_density(p, x, emulate_bf16=False):
    """eonerf.py:141-145"""
    return F.softplus(_lin(p, "sigma_layer.output_layer", trunk(p, x, emulate_bf16)))


def field_forward(p, x, sun_dirs, img_indices, emulate_bf16=False):
    """eonerf.py:154-170 → (sigma[N,1], albedo[N,3], ambient[N,3], transient_s[N,1], transient_beta[N,1]).
    emulate_bf16: round where csrc/field.cu rounds in bf16 mode (GEMM operands); the narrow heads, the biases, the
    embedding term and the ambient branch stay fp32 exactly as in the kernels."""
    h = trunk(p, x, emulate_bf16)
    sigma = F.softplus(_lin(p, "sigma_layer.output_layer", h))
    amb = torch.relu(_lin(p, "ambient_mlp.hidden_layers.0", posenc(sun_dirs, VIEW_L)))
    ambient = torch.sigmoid(_lin(p, "ambient_mlp.output_layer", amb))
    emb = p["transient_encoder.weight"][img_indices.reshape(-1)]
    if not emulate_bf16:
        bott = _lin(p, "bottleneck_layer.output_layer", h)
        a = torch.relu(_lin(p, "albedo_mlp.hidden_layers.0", bott))
        t = torch.cat([bott, emb], -1)
        for i in range(4):
            t = torch.relu(_lin(p, f"transient_mlp.hidden_layers.{i}", t))
    else:
        bott = _r16(_lin16(p, "bottleneck_layer.output_layer", h) + p["bottleneck_layer.output_layer.bias"])
        a = _r16(torch.relu(_lin16(p, "albedo_mlp.hidden_layers.0", bott) + p["albedo_mlp.hidden_layers.0.bias"]))
        w0 = p["transient_mlp.hidden_layers.0.weight"]
        t = _r16(torch.relu(_lin16(p, "transient_mlp.hidden_layers.0", bott, (0, 256)) + emb @ w0[:, 256:260].t()
                            + p["transient_mlp.hidden_layers.0.bias"]))
        for i in range(1, 4):
            t = _r16(torch.relu(_lin16(p, f"transient_mlp.hidden_layers.{i}", t) + p[f"transient_mlp.hidden_layers.{i}.bias"]))
    albedo = torch.sigmoid(_lin(p, "albedo_mlp.output_layer", a))
    s = torch.sigmoid(_lin(p, "transient_scalar.output_layer", t))
    beta = F.softplus(_lin(p, "transient_beta.output_layer", t))
    return sigma, albedo, ambient, s, beta


# --------------------------------------------------------------------------------------------
# vanilla NeRF field (mlp.py:114-165, 211-250) — BASELINE config 2
# --------------------------------------------------------------------------------------------
def vanilla_param_shapes():
    s = OrderedDict()
    ins = [63, 256, 256, 256, 256, 319, 256, 256]
    for i, k in enumerate(ins):
        s[f"mlp.base.hidden_layers.{i}.weight"] = (256, k)
        s[f"mlp.base.hidden_layers.{i}.bias"] = (256,)
    s["mlp.sigma_layer.output_layer.weight"] = (1, 256)
    s["mlp.sigma_layer.output_layer.bias"] = (1,)
    s["mlp.bottleneck_layer.output_layer.weight"] = (256, 256)
    s["mlp.bottleneck_layer.output_layer.bias"] = (256,)
    s["mlp.rgb_layer.hidden_layers.0.weight"] = (128, 283)
    s["mlp.rgb_layer.hidden_layers.0.bias"] = (128,)
    s["mlp.rgb_layer.output_layer.weight"] = (3, 128)
    s["mlp.rgb_layer.output_layer.bias"] = (3,)
    return s


def init_vanilla_params(seed=0, bias_scale=0.0):
    g = torch.Generator().manual_seed(seed)
    p = OrderedDict()
    for name, shape in vanilla_param_shapes().items():
        if name.endswith(".weight"):
            bound = math.sqrt(6.0 / (shape[0] + shape[1]))
            p[name] = (torch.rand(shape, generator=g) * 2 - 1) * bound
        else:
            p[name] = bias_scale * torch.randn(shape, generator=g)
    return p


def vanilla_forward(p, x, viewdirs):
    """VanillaNeRFRadianceField.forward (mlp.py:245-250): rgb = sigmoid(rgb_layer(cat[bottleneck, enc4(dir)])),
    sigma = relu(sigma_layer(trunk))."""
    enc = posenc(x, POS_L)
    h = enc
    for i in range(8):
        h = torch.relu(_lin(p, f"mlp.base.hidden_layers.{i}", h))
        if i == 4:
            h = torch.cat([h, enc], -1)
    raw_sigma = _lin(p, "mlp.sigma_layer.output_layer", h)
    bott = _lin(p, "mlp.bottleneck_layer.output_layer", h)
    c = torch.cat([bott, posenc(viewdirs, VIEW_L)], -1)
    c = torch.relu(_lin(p, "mlp.rgb_layer.hidden_layers.0", c))
    rgb = torch.sigmoid(_lin(p, "mlp.rgb_layer.output_layer", c))
    return rgb, torch.relu(raw_sigma)


# --------------------------------------------------------------------------------------------
# per-ray rendering (eonerf.py:196-248) and sun-ray shadows (sat_rendering.py:87-118)
# --------------------------------------------------------------------------------------------
def _last_sample_index(ray_indices):
    """eonerf.py:218-219"""
    _, counts = torch.unique(ray_indices, return_counts=True)
    return torch.cumsum(counts, 0) - 1


def rendering(p, rays, t_starts, t_ends, ray_indices, emulate_bf16=False):
    """eonerf.py:196-248.  NOTE mutates t_ends in place like the reference (:220)."""
    n_rays = rays.origins.shape[0]
    z = ((t_starts + t_ends)[:, None] / 2.0).to(rays.origins.dtype)
    x = rays.origins[ray_indices] + rays.viewdirs[ray_indices] * z
    t_ends[_last_sample_index(ray_indices)] = 1e10
    t_starts, t_ends = t_starts.to(z.dtype), t_ends.to(z.dtype)
    sigma, albedo, ambient, ts, tb = field_forward(p, x, rays.sundirs[ray_indices], rays.img_idx[ray_indices], emulate_bf16)
    w, trans, alphas = nv.render_weight_from_density(t_starts, t_ends, sigma.squeeze(-1),
                                                     ray_indices=ray_indices, n_rays=n_rays)
    acc = lambda v: nv.accumulate_along_rays(w, values=v, ray_indices=ray_indices, n_rays=n_rays)
    depth, albedo_, ambient_, ts_, tb_ = acc(z), acc(albedo), acc(ambient), acc(ts), acc(tb)
    tb_ = tb_ + BETA_MIN
    extras = dict(sigma=sigma.squeeze(-1), weights=w, trans=trans, alphas=alphas, x=x,
                  albedo_pts=albedo, ambient_pts=ambient, ts_pts=ts, tb_pts=tb)
    return albedo_, depth, tb_, ts_, ambient_, torch.ones_like(depth), extras


def render_depth(p, rays, t_starts, t_ends, ray_indices):
    """eonerf.py:172-194"""
    n_rays = rays.origins.shape[0]
    z = (t_starts + t_ends)[:, None] / 2.0
    x = rays.origins[ray_indices] + rays.viewdirs[ray_indices] * z
    t_ends[_last_sample_index(ray_indices)] = 1e10
    sigma = query_density(p, x).squeeze(-1)
    w, _, _ = nv.render_weight_from_density(t_starts, t_ends, sigma, ray_indices=ray_indices, n_rays=n_rays)
    return nv.accumulate_along_rays(w, values=z, ray_indices=ray_indices, n_rays=n_rays)


def geometric_shadows(p, rays, depth, n_samples, u_sun, z_steps=None, emulate_bf16=False):
    """sat_rendering.py:87-118: secondary rays from the rendered surface point towards the sun;
    shadow = transmittance *before* the last kept sample of each sun ray; 1 for rays with no samples."""
    n_rays = rays.origins.shape[0]
    sc_o = rays.origins + torch.hstack([depth, depth, depth]) * rays.viewdirs
    sc_d = -1.0 * rays.sundirs
    ri, ts, te, _ = satnerf_sampling(sc_o, sc_d, n_samples, u_sun, near=None, z_steps=z_steps)
    sc_pts = pts_per_ray(ri, n_rays)
    z = ((ts + te)[:, None] / 2.0).to(sc_o.dtype)
    ts, te = ts.to(sc_o.dtype), te.to(sc_o.dtype)
    x = sc_o[ri] + sc_d[ri] * z
    sigma = query_density(p, x, emulate_bf16).squeeze(-1)
    trans, _ = nv.render_transmittance_from_density(ts, te, sigma, ray_indices=ri, n_rays=n_rays)
    geo = torch.ones((n_rays, 1), dtype=sc_o.dtype)
    if ri.numel():
        uniq, counts = torch.unique(ri, return_counts=True)
        geo = geo.index_put((uniq,), trans.view(-1, 1)[torch.cumsum(counts, 0) - 1])
    return geo, sc_pts, dict(ray_indices=ri, t_starts=ts, t_ends=te, sigma=sigma, trans=trans, x=x, origins=sc_o)


OUT_KEYS = OrderedDict([  # sat_rendering.py:322-334
    ("rgb", (0, 3)), ("depth", (3, 4)), ("albedo_rgb", (4, 7)), ("ambient_rgb", (7, 10)),
    ("geo_shadows", (10, 11)), ("transient_s", (11, 12)), ("beta", (12, 13)), ("entropy", (13, 14)),
    ("pts_per_ray", (14, 15)), ("sc_pts_per_ray", (15, 16)), ("opacity_after_surface", (16, 18)),
    ("shadowless_rgb", (18, 21))])


def render_chunk(p, rays, n_samples, epoch_idx, u_cam, u_sun=None, u_cam2=None, eval=False,
                 radiometric=True, z_steps=None, return_extras=False, emulate_bf16=False, force_redraw=None,
                 eval_img=None):
    """One iteration of the chunk loop of sat_rendering.py:252-313 → out [B, 21], n_samples_rendered.
    emulate_bf16: the MLPs round where the CUDA bf16 path rounds (see field_forward).
    force_redraw / eval_img: for callers that evaluate a chunk of the reference in independent slices of rays
    (render_chunk_sliced): the chunk-wide decisions of :259-262 ("any ray empty -> draw again") and :288-289
    (eval: image index of the chunk's first ray) are then made by the caller over the whole chunk."""
    n_rays = rays.origins.shape[0]
    ri, ts, te, _ = satnerf_sampling(rays.origins, rays.viewdirs, n_samples, u_cam, near=rays.t_near, z_steps=z_steps)
    ppr = pts_per_ray(ri, n_rays)
    if bool(torch.sum(ppr == 0)) if force_redraw is None else force_redraw:      # :260-262
        assert u_cam2 is not None, "a ray kept no samples: the reference re-draws, pass u_cam2"
        ri, ts, te, _ = satnerf_sampling(rays.origins, rays.viewdirs, n_samples, u_cam2, near=None, z_steps=z_steps)
    albedo, depth, beta, tr_s, ambient, entropy, ex = rendering(p, rays, ts, te, ri, emulate_bf16)
    ambient = ambient * 0.2                                                       # :265
    sc_ex = None
    if epoch_idx < 2:                                                             # :269-272
        geo = torch.ones((n_rays, 1), dtype=albedo.dtype)
        s = geo
        sc_ppr = torch.ones_like(ppr)
    else:
        geo, sc_ppr, sc_ex = geometric_shadows(p, rays, depth, n_samples, u_sun, z_steps=z_steps, emulate_bf16=emulate_bf16)
        s = geo * tr_s                                                            # :276
    first_img = rays.img_idx[0] if eval_img is None else eval_img
    img = (torch.ones(n_rays, dtype=torch.long) * first_img) if eval else rays.img_idx.reshape(-1)  # :288-291
    rgb = albedo * s + (1 - s) * (ambient * albedo)                               # :294
    if radiometric:
        emb = p["radiometricT_enc.weight"][img]
        A, b = emb[:, :3], emb[:, 3:6]
    else:
        A, b = torch.ones_like(rgb), torch.zeros_like(rgb)
    rgb = torch.clip(A * rgb + b, 0, 1)                                           # :304-305
    shadowless = A * albedo + b                                                   # :306
    dt = rgb.dtype
    out = torch.cat([rgb, depth, albedo, ambient, geo, tr_s, beta, entropy, ppr[:, None].to(dt), sc_ppr[:, None].to(dt),
                     torch.ones(n_rays, 2, dtype=dt), shadowless], dim=1)         # :311-312
    if return_extras:
        return out, len(ts), dict(cam=dict(ray_indices=ri, t_starts=ts, t_ends=te, **ex), sun=sc_ex)
    return out, len(ts)


def render_image(p, rays, n_samples, epoch_idx, chunk, us, eval=False, radiometric=True):
    """sat_rendering.py:176-335.  `us` = list (one per chunk) of dicts u_cam / u_sun / u_cam2."""
    shape = rays.origins.shape
    if len(shape) == 3:
        n = shape[0] * shape[1]
        rays = SatRays(*[r.reshape([n] + list(r.shape[2:])) for r in rays])
    else:
        n = shape[0]
    outs, total = [], 0
    for ci, i in enumerate(range(0, n, chunk)):
        cr = SatRays(*[r[i:i + chunk] for r in rays])
        o, k = render_chunk(p, cr, n_samples, epoch_idx, eval=eval, radiometric=radiometric, **us[ci])
        outs.append(o)
        total += k
    out = torch.cat(outs, 0)
    return {k: out[:, a:b].view(*shape[:-1], -1) for k, (a, b) in OUT_KEYS.items()}, total


# --------------------------------------------------------------------------------------------
# losses (metrics.py:17-22, 68-69) and one training step (train_eonerf.py:122-161)
# --------------------------------------------------------------------------------------------
def uncertainty_aware_loss(gt, rgb, beta):
    color = ((rgb - gt) ** 2 / (2 * beta ** 2)).mean()
    return color + (3 + torch.log(beta).mean()) / 2


def loss_from_out(out, pixels, epoch_idx):
    if epoch_idx < 2:
        return F.mse_loss(out[:, 0:3], pixels)
    return uncertainty_aware_loss(pixels, out[:, 0:3], out[:, 12:13])


def chunk_needs_redraw(rays, n_samples, u_cam, z_steps=None):
    """sat_rendering.py:258-260: does any ray of the chunk keep no sample under the first draw?"""
    ri, _, _, _ = satnerf_sampling(rays.origins, rays.viewdirs, n_samples, u_cam, near=rays.t_near, z_steps=z_steps)
    return bool(torch.sum(pts_per_ray(ri, rays.origins.shape[0]) == 0))


def _slices(n, k):
    return [slice(i, min(n, i + k)) for i in range(0, n, k)]


def _take(us, sl):
    return None if us is None else us[sl]


def render_chunk_sliced(p, rays, n_samples, epoch_idx, u_cam, u_sun=None, u_cam2=None, eval=False, radiometric=True,
                        emulate_bf16=False, rays_per_slice=1024, z_steps=None):
    """render_chunk on a chunk too large for one CPU pass: rays are independent, so the chunk is evaluated
    in slices; the two chunk-wide decisions of the reference (re-draw, eval image index) are taken over the whole chunk
    first.  Returns (out [B,21], n_rendering_samples) exactly as render_chunk would on the whole chunk."""
    B = rays.origins.shape[0]
    redraw = chunk_needs_redraw(rays, n_samples, u_cam, z_steps)
    outs, total = [], 0
    for sl in _slices(B, rays_per_slice):
        o, k = render_chunk(p, SatRays(*[r[sl] for r in rays]), n_samples, epoch_idx, u_cam[sl], _take(u_sun, sl), _take(u_cam2, sl),
                            eval=eval, radiometric=radiometric, z_steps=z_steps, emulate_bf16=emulate_bf16, force_redraw=redraw,
                            eval_img=rays.img_idx[0])
        outs.append(o)
        total += k
    return torch.cat(outs, 0), total


def train_step_grads_sliced(p, rays, pixels, n_samples, epoch_idx, u_cam, u_sun=None, u_cam2=None, radiometric=True,
                            emulate_bf16=False, rays_per_slice=1024):
    """train_step_grads on a batch too large for one CPU autograd pass: the losses are batch means (metrics.py:17-22), so
    the gradient of the batch is the size-weighted sum of the slices' gradients.  -> (loss, out, grads, n_rendered)."""
    B = rays.origins.shape[0]
    redraw = chunk_needs_redraw(rays, n_samples, u_cam)
    q = OrderedDict((k, v.detach().clone().requires_grad_(True)) for k, v in p.items())
    outs, total, loss_sum = [], 0, 0.0
    for sl in _slices(B, rays_per_slice):
        o, k = render_chunk(q, SatRays(*[r[sl] for r in rays]), n_samples, epoch_idx, u_cam[sl], _take(u_sun, sl), _take(u_cam2, sl),
                            radiometric=radiometric, emulate_bf16=emulate_bf16, force_redraw=redraw)
        loss = loss_from_out(o, pixels[sl], epoch_idx) * ((sl.stop - sl.start) / B)
        loss.backward()
        outs.append(o.detach())
        loss_sum += float(loss.detach())
        total += k
    grads = OrderedDict((k, (v.grad if v.grad is not None else torch.zeros_like(v))) for k, v in q.items())
    return torch.tensor(loss_sum), torch.cat(outs, 0), grads, total


def train_step_grads(p, rays, pixels, n_samples, epoch_idx, u_cam, u_sun=None, u_cam2=None, radiometric=True,
                     dtype=None, emulate_bf16=False):
    """forward + loss + backward; returns (loss, out, {name: grad}, n_rendering_samples).  dtype=torch.float64 runs
    everything but the sampler in double precision (conditioning reference for the gradient tests)."""
    if dtype is not None:
        p = OrderedDict((k, v.to(dtype)) for k, v in p.items())
        rays = SatRays(*[r if r.dtype == torch.int64 else r.to(dtype) for r in rays])
        pixels = pixels.to(dtype)
    q = OrderedDict((k, v.detach().clone().requires_grad_(True)) for k, v in p.items())
    out, n_rendered = render_chunk(q, rays, n_samples, epoch_idx, u_cam, u_sun, u_cam2, radiometric=radiometric,
                                   emulate_bf16=emulate_bf16)
    loss = loss_from_out(out, pixels, epoch_idx)
    loss.backward()
    grads = OrderedDict((k, (v.grad if v.grad is not None else torch.zeros_like(v))) for k, v in q.items())
    return loss.detach(), out.detach(), grads, n_rendered


# --------------------------------------------------------------------------------------------
# BASELINE configs[1]: vanilla NeRF rendering (train_mlp_nerf.py:155-170 -> nerfacc v0.5.2 examples/utils.py
# render_image_with_occgrid + nerfacc.rendering).  The reference's own helper module is missing (train_mlp_nerf.py:17):
# PARITY UNPINNED.  Restated: the occupancy-free limit of the sampler + the published conventions of nerfacc.rendering.
# --------------------------------------------------------------------------------------------
def march_aabb(origins, viewdirs, aabb, near_plane, far_plane, step, jitter=None, max_per_ray=4096):
    """fp32, one torch op per arithmetic step (the CUDA kernel uses the same separately rounded operations)."""
    o, d = origins.float(), viewdirs.float()
    aabb = torch.as_tensor(aabb, dtype=torch.float32)
    t1, t2 = (aabb[:3] - o) / d, (aabb[3:] - o) / d
    tmin = torch.maximum(torch.minimum(t1, t2).max(dim=1).values, torch.tensor(float(near_plane)))
    tmax = torch.minimum(torch.maximum(t1, t2).min(dim=1).values, torch.tensor(min(float(far_plane), 3.0e38)))
    step_t = torch.tensor(step, dtype=torch.float32)
    t0 = tmin + (jitter.float() * step_t if jitter is not None else 0.0)
    n = torch.where(tmax > t0, torch.ceil((tmax - t0) / step_t), torch.zeros_like(t0)).long().clamp(max=max_per_ray)
    ri = torch.repeat_interleave(torch.arange(o.shape[0]), n)
    starts = torch.cumsum(n, 0) - n
    k = (torch.arange(ri.numel()) - starts[ri]).float()
    ts = t0[ri] + k * step_t
    te = torch.minimum(ts + step_t, tmax[ri])
    return ri, ts, te


def vanilla_render(p, origins, viewdirs, aabb, near_plane, far_plane, step, jitter=None, render_bkgd=None):
    """-> (colors[B,3], opacities[B,1], depths[B,1], n_samples) with nerfacc.rendering's conventions."""
    B = origins.shape[0]
    ri, ts, te = march_aabb(origins, viewdirs, aabb, near_plane, far_plane, step, jitter)
    z = (ts + te)[:, None] / 2.0
    x = origins[ri] + viewdirs[ri] * z
    rgb, sigma = vanilla_forward(p, x, viewdirs[ri])
    w, _, _ = nv.render_weight_from_density(ts, te, sigma.squeeze(-1), ray_indices=ri, n_rays=B)
    colors = nv.accumulate_along_rays(w, rgb, ri, B)
    opac = nv.accumulate_along_rays(w, None, ri, B)
    depth = nv.accumulate_along_rays(w, z, ri, B) / opac.clamp_min(torch.finfo(torch.float32).eps)
    if render_bkgd is not None:
        colors = colors + render_bkgd * (1.0 - opac)
    return colors, opac, depth, int(ts.numel())
