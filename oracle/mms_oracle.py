"""TEST INFRASTRUCTURE — CPU restatement (oracle) of the MMS-FW per-ray rendering hot path.

This file is NOT part of the product: only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` may import it, and only as the checker / the CPU arm.  The product
path (`multimodalstudio_b200`) never imports it and has no CPU fallback.

It restates, function by function, the reference's algorithm in plain fp32 PyTorch on the CPU (the
reference itself is PyTorch; gradients come from autograd on this restatement).  Every function cites
the reference lines it follows (paths relative to /root/reference/src).

Pinning: the reference ships no golden vectors or tests.  The oracle is therefore pinned
against OUTPUTS OF THE REFERENCE ITSELF, run in the build container by `oracle/make_golden.py` (which
imports the unmodified reference through `oracle/ref_harness.py`) and committed as `tests/golden/*.npz`;
`tests/test_oracle_golden.py` checks this file against those fixtures everywhere, and
`tests/test_oracle_vs_reference.py` against the live reference where `/root/reference` exists.
Not pinned (tiny-cuda-nn is not in the tree): Smoothstep hash interpolation and tcnn's SH op.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

PRIMES = (1, 2654435761, 805459861)


# ------------------------------------------------------------------------------------------------
# A9/A10 hash grid — field_components/encodings.py:195-197,223-304; feature_structures.py:78-88
# ------------------------------------------------------------------------------------------------
def hash_resolutions(min_res, max_res, num_levels):
    growth = np.exp((np.log(max_res) - np.log(min_res)) / (num_levels - 1))
    return torch.floor(min_res * growth ** torch.arange(num_levels))


def hash_indices(x, res, log2_t):
    """x [...,3] (already rescaled) -> int64 [..., L, 8] (hashed_0..7, encodings.py:274-281), offsets [...,L,3]"""
    t = 2 ** log2_t
    num_levels = res.shape[0]
    scaled = x[..., None, :] * res.view(-1, 1)
    c = torch.ceil(scaled).to(torch.int32)
    f = torch.floor(scaled).to(torch.int32)
    off = scaled - f

    def h(ix, iy, iz):
        v = torch.stack([ix, iy, iz], -1).to(torch.int64) * torch.tensor(PRIMES)
        r = torch.bitwise_xor(torch.bitwise_xor(v[..., 0], v[..., 1]), v[..., 2])
        return r % t + torch.arange(num_levels) * t

    cx, cy, cz = c[..., 0], c[..., 1], c[..., 2]
    fx, fy, fz = f[..., 0], f[..., 1], f[..., 2]
    idx = torch.stack([h(cx, cy, cz), h(cx, fy, cz), h(fx, fy, cz), h(fx, cy, cz),
                       h(cx, cy, fz), h(cx, fy, fz), h(fx, fy, fz), h(fx, cy, fz)], -1)
    return idx, off


def hash_encode(x, table, res, log2_t, radius=None, mask=None, smoothstep=False):
    if radius is not None and radius > 0:
        x = (x + radius) / (2 * radius)
    idx, off = hash_indices(x, res, log2_t)
    f = [table[idx[..., k]] for k in range(8)]
    if smoothstep:
        off = off * off * (3 - 2 * off)
    ox, oy, oz = off[..., 0:1], off[..., 1:2], off[..., 2:3]
    f03 = f[0] * ox + f[3] * (1 - ox)
    f12 = f[1] * ox + f[2] * (1 - ox)
    f56 = f[5] * ox + f[6] * (1 - ox)
    f47 = f[4] * ox + f[7] * (1 - ox)
    f0312 = f03 * oy + f12 * (1 - oy)
    f4756 = f47 * oy + f56 * (1 - oy)
    out = torch.flatten(f0312 * oz + f4756 * (1 - oz), start_dim=-2)
    if mask is not None:
        out = out * mask
    return out


# ------------------------------------------------------------------------------------------------
# A8 / A15 encodings — encodings.py:161-182; utils/math.py:21-82
# ------------------------------------------------------------------------------------------------
def nerf_encode(x, num_frequencies, min_freq, max_freq, include_input=True):
    freqs = 2 ** torch.linspace(min_freq, max_freq, num_frequencies)
    s = (x[..., None] * freqs).reshape(*x.shape[:-1], -1)
    enc = torch.sin(torch.cat([s, s + torch.pi / 2.0], dim=-1))
    return torch.cat([x, enc], dim=-1) if include_input else enc


def sh_encode(levels, d):
    x, y, z = d[..., 0], d[..., 1], d[..., 2]
    xx, yy, zz = x ** 2, y ** 2, z ** 2
    c = [torch.full_like(x, 0.28209479177387814)]
    if levels > 1:
        c += [0.4886025119029199 * y, 0.4886025119029199 * z, 0.4886025119029199 * x]
    if levels > 2:
        c += [1.0925484305920792 * x * y, 1.0925484305920792 * y * z, 0.9461746957575601 * zz - 0.31539156525251999,
              1.0925484305920792 * x * z, 0.5462742152960396 * (xx - yy)]
    if levels > 3:
        c += [0.5900435899266435 * y * (3 * xx - yy), 2.890611442640554 * x * y * z,
              0.4570457994644658 * y * (5 * zz - 1), 0.3731763325901154 * z * (5 * zz - 3),
              0.4570457994644658 * x * (5 * zz - 1), 1.445305721320277 * z * (xx - yy),
              0.5900435899266435 * x * (xx - 3 * yy)]
    if levels > 4:
        c += [2.5033429417967046 * x * y * (xx - yy), 1.7701307697799304 * y * z * (3 * xx - yy),
              0.9461746957575601 * x * y * (7 * zz - 1), 0.6690465435572892 * y * (7 * zz - 3),
              0.10578554691520431 * (35 * zz * zz - 30 * zz + 3), 0.6690465435572892 * x * z * (7 * zz - 3),
              0.47308734787878004 * (xx - yy) * (7 * zz - 1), 1.7701307697799304 * x * z * (xx - 3 * yy),
              0.4425326924449826 * (xx * (xx - 3 * yy) - yy * (3 * xx - yy))]
    return torch.stack(c, dim=-1)


# ------------------------------------------------------------------------------------------------
# A11 MLP — mlp.py:152-171,206-209
# ------------------------------------------------------------------------------------------------
def wn_weight(sd, prefix):
    """weight_norm parametrization (torch.nn.utils.parametrizations.weight_norm, dim=0): W = v * (g / ||v||_row),
    evaluated by the same ATen op the parametrization calls so the weights are bit-identical."""
    g, v = sd[prefix + ".parametrizations.weight.original0"], sd[prefix + ".parametrizations.weight.original1"]
    return torch._weight_norm(v, g, 0)


def _act(name, beta=1.0):
    if name in (None, "None"):
        return lambda t: t
    if name == "ReLU":
        return F.relu
    if name == "Softplus":
        return lambda t: F.softplus(t, beta=beta)
    if name == "Sigmoid":
        return torch.sigmoid
    raise ValueError(name)


def mlp(sd, prefix, x, num_layers, hidden_act, out_act, beta=1.0, skips=(), out_beta=1.0):
    h = x
    for i in range(num_layers):
        if i in skips:
            h = torch.cat([h, x], -1) / np.sqrt(2)
        p = f"{prefix}.layers.{i}"
        h = F.linear(h, wn_weight(sd, p), sd[p + ".bias"])
        if i < num_layers - 1:
            h = _act(hidden_act, beta)(h)
    return _act(out_act, out_beta)(h)


# ------------------------------------------------------------------------------------------------
# A1/A2 ray generation — camera_optimizers.py:86-119; lie_groups.py:28-63; ray_generators.py:54-81;
# cameras.py:534-703; camera_utils.py:279-383; poses.py:53-67
# ------------------------------------------------------------------------------------------------
def exp_map_so3xr3(tv):
    w = tv[:, 3:]
    ang = torch.clamp((w * w).sum(1), 1e-4).sqrt()
    inv = 1.0 / ang
    f1 = inv * ang.sin()
    f2 = inv * inv * (1.0 - ang.cos())
    zero = torch.zeros_like(w[:, 0])
    k = torch.stack([zero, -w[:, 2], w[:, 1], w[:, 2], zero, -w[:, 0], -w[:, 1], w[:, 0], zero], -1).view(-1, 3, 3)
    r = f1[:, None, None] * k + f2[:, None, None] * torch.bmm(k, k) + torch.eye(3)[None]
    return torch.cat([r, tv[:, :3, None]], dim=-1)


def undistort(coords, dp, eps=1e-3, iters=10):
    xd, yd = coords[..., 0], coords[..., 1]
    k1, k2, k3, k4, p1, p2 = [dp[..., i] for i in range(6)]
    x, y = xd, yd
    for _ in range(iters):
        r = x * x + y * y
        d = 1.0 + r * (k1 + r * (k2 + r * (k3 + r * k4)))
        fx = d * x + 2 * p1 * x * y + p2 * (r + 2 * x * x) - xd
        fy = d * y + 2 * p2 * x * y + p1 * (r + 2 * y * y) - yd
        d_r = k1 + r * (2.0 * k2 + r * (3.0 * k3 + r * 4.0 * k4))
        d_x, d_y = 2.0 * x * d_r, 2.0 * y * d_r
        fx_x = d + d_x * x + 2.0 * p1 * y + 6.0 * p2 * x
        fx_y = d_y * x + 2.0 * p1 * x + 2.0 * p2 * y
        fy_x = d_x * y + 2.0 * p2 * y + 2.0 * p1 * x
        fy_y = d + d_y * y + 2.0 * p2 * x + 6.0 * p1 * y
        den = fy_x * fx_y - fx_x * fy_y
        ok = den.abs() > eps
        x = x + torch.where(ok, (fx * fy_y - fy * fx_y) / den, torch.zeros_like(den))
        y = y + torch.where(ok, (fy * fx_x - fx * fy_x) / den, torch.zeros_like(den))
    return torch.stack([x, y], -1)


def raygen(coords, c2w, intr, dist, pose_adjust, pixel_offset=0.0):
    """coords int [R,3] (cam,y,x); c2w [n,3,4]; intr [n,4]=fx,fy,cx,cy; dist [n,6]|None; pose_adjust [1|n,6]|None."""
    cam = coords[:, 0].long()
    y = coords[:, 1].float() + pixel_offset
    x = coords[:, 2].float() + pixel_offset
    fx, fy, cx, cy = [intr[cam, i] for i in range(4)]
    pts = torch.stack([torch.stack([(x - cx) / fx, -(y - cy) / fy], -1),
                       torch.stack([(x - cx + 1) / fx, -(y - cy) / fy], -1),
                       torch.stack([(x - cx) / fx, -(y - cy + 1) / fy], -1)], 0)
    if dist is not None:
        pts = undistort(pts, dist[cam][None])
    dirs = torch.cat([pts, -torch.ones_like(pts[..., :1])], -1)
    pose = c2w[cam]
    if pose_adjust is not None:
        pa = pose_adjust.expand(c2w.shape[0], 6)[cam] if pose_adjust.shape[0] == 1 else pose_adjust[cam]
        delta = exp_map_so3xr3(pa)
        r1, t1, r2, t2 = pose[:, :, :3], pose[:, :, 3:], delta[:, :, :3], delta[:, :, 3:]
        pose = torch.cat([r1 @ r2, t1 + r1 @ t2], -1)
    rot = pose[:, :, :3]
    dw = torch.sum(dirs[..., None, :] * rot, dim=-1)
    dnorm = torch.norm(dw, dim=-1, keepdim=True)[0]
    dw = F.normalize(dw, dim=-1)
    up = rot[:, :, 1]
    dx = torch.sqrt(torch.sum((dw[0] - dw[1]) ** 2, -1))
    dy = torch.sqrt(torch.sum((dw[0] - dw[2]) ** 2, -1))
    return {"origins": pose[:, :, 3], "directions": dw[0], "up_directions": up, "pixel_area": (dx * dy)[:, None],
            "directions_norm": dnorm}


# ------------------------------------------------------------------------------------------------
# A3 collider — scene_colliders.py:60-80,107-113
# ------------------------------------------------------------------------------------------------
def sphere_collide(o, d, radius=1.0):
    b = (d * o).sum(dim=-1, keepdim=True)
    under = b ** 2 - (o.norm(p=2, dim=-1, keepdim=True) ** 2 - radius ** 2)
    mask = (under > 0.01).squeeze(-1)
    sq = torch.sqrt(under.clamp_min(0.01))
    nears = (-sq - b).clamp_min(0.01)
    fars = (sq - b).clamp_min(0.01)
    return nears, fars, mask


def background_near_far(nears, fars, mask):
    return torch.where(mask[:, None], fars, nears), fars + 3.0


# ------------------------------------------------------------------------------------------------
# A4..A7 samplers — ray_samplers.py:38-68,183-233,316-422,448-551; rays.py:201-217
# ------------------------------------------------------------------------------------------------
def spacing_to_euclid(x, nears, fars, disparity=False):
    if disparity:
        return 1 / ((1 / fars) * x + (1 / nears) * (1 - x))
    return fars * x + nears * (1 - x)


def spaced_bins(nears, fars, num_samples, t_rand=None, disparity=False):
    bins = torch.linspace(0.0, 1.0, num_samples + 1)[None]
    if t_rand is not None:
        centers = (bins[..., 1:] + bins[..., :-1]) / 2.0
        upper = torch.cat([centers, bins[..., -1:]], -1)
        lower = torch.cat([bins[..., :1], centers], -1)
        bins = lower + (upper - lower) * t_rand
    bins = bins.expand(nears.shape[0], num_samples + 1)
    return bins, spacing_to_euclid(bins, nears, fars, disparity)


def fixed_inv_s_alphas(sdf, deltas, inv_s):
    """sdf [R,m], deltas [R,m-1]"""
    prev, nxt = sdf[:, :-1], sdf[:, 1:]
    mid = (prev + nxt) * 0.5
    cos = (nxt - prev) / (deltas + 1e-5)
    prev_cos = torch.cat([torch.zeros(sdf.shape[0], 1), cos[:, :-1]], -1)
    cos = torch.minimum(prev_cos, cos).clip(-1e3, 0.0)
    pe = mid - cos * deltas * 0.5
    ne = mid + cos * deltas * 0.5
    pc, nc = torch.sigmoid(pe * inv_s), torch.sigmoid(ne * inv_s)
    return (pc - nc + 1e-5) / (pc + 1e-5)


def weights_from_alphas(alphas):
    t = torch.cumprod(torch.cat([torch.ones(alphas.shape[0], 1), 1.0 - alphas + 1e-7], 1), 1)
    return alphas * t[:, :-1]


def pdf_cdf(weights, hist_pad, eps=1e-5):
    w = weights + hist_pad
    ws = torch.sum(w, dim=-1, keepdim=True)
    pad = torch.relu(eps - ws)
    w = w + pad / w.shape[-1]
    ws = ws + pad
    cdf = torch.min(torch.ones_like(w), torch.cumsum(w / ws, dim=-1))
    return torch.cat([torch.zeros_like(cdf[..., :1]), cdf], dim=-1)


def make_u(num_rays, num_samples, rand=None):
    nb = num_samples + 1
    u = torch.linspace(0.0, 1.0 - (1.0 / nb), steps=nb)
    if rand is not None:
        return (u.expand(num_rays, nb) + rand / nb).contiguous()
    return (u + 1.0 / (2 * nb)).expand(num_rays, nb).contiguous()


def pdf_inverse(cdf, bins, u):
    inds = torch.searchsorted(cdf, u, side="right")
    below = torch.clamp(inds - 1, 0, bins.shape[-1] - 1)
    above = torch.clamp(inds, 0, bins.shape[-1] - 1)
    c0, c1 = torch.gather(cdf, -1, below), torch.gather(cdf, -1, above)
    b0, b1 = torch.gather(bins, -1, below), torch.gather(bins, -1, above)
    t = torch.clip(torch.nan_to_num((u - c0) / (c1 - c0), 0), 0, 1)
    return inds, b0 + t * (b1 - b0)


def merge_bins(bins_old, bins_new):
    """bins_* are edge arrays; merges the starts, keeps max of the last edges (ray_samplers.py:46-53)."""
    starts = torch.cat([bins_old[:, :-1], bins_new[:, :-1]], -1)
    merged, index = torch.sort(starts, dim=-1, stable=True)
    ends = torch.maximum(bins_old[:, -1:], bins_new[:, -1:])
    return torch.cat([merged, ends], -1), index


def upsample_round(bins, sdf, u, nears, fars, inv_s, hist_pad=1e-5):
    e = spacing_to_euclid(bins, nears, fars)
    deltas = (e[:, 1:] - e[:, :-1])[:, :-1]
    alphas = fixed_inv_s_alphas(sdf, deltas, inv_s)
    w = weights_from_alphas(alphas)
    w = torch.cat([w, torch.zeros_like(w[:, :1])], 1)
    cdf = pdf_cdf(w, hist_pad)
    inds, new_bins = pdf_inverse(cdf, bins, u)
    merged, index = merge_bins(bins, new_bins)
    return {"cdf": cdf, "inds": inds, "new_bins": new_bins, "merged_bins": merged, "merged_index": index}


# ------------------------------------------------------------------------------------------------
# A13 / A14 — surface_model.py:137-153,98; volume_rendering.py:177-213; single_variance.py:34-36
# ------------------------------------------------------------------------------------------------
TAPS = torch.tensor([[1., -1, -1], [-1, -1, 1], [-1, 1, -1], [1, 1, 1]])


def taps_gradients(sdf_c, sdf_t, delta, want_hessian=True):
    """sdf_c [n,1], sdf_t [4,n,1]; delta = numerical_gradients_delta / sqrt(3) (python float)."""
    T = TAPS.to(sdf_t.device)
    g = (T[0] * sdf_t[0] + T[1] * sdf_t[1] + T[2] * sdf_t[2] + T[3] * sdf_t[3]) / (4.0 * delta)
    h = None
    if want_hessian:
        hxx = ((sdf_t[0] + sdf_t[1] + sdf_t[2] + sdf_t[3]) / 2.0 - 2 * sdf_c) / delta ** 2
        h = torch.cat([hxx, hxx, hxx], dim=-1) / 3.0
    return g, h, F.normalize(g, p=2, dim=-1)


def neus_alphas(sdf, gradients, dirs, deltas, inv_s, anneal):
    """sdf [R,S,1], gradients [R,S,3], dirs [R,1,3], deltas [R,S,1]"""
    true_cos = (dirs * gradients).sum(-1, keepdim=True)
    iter_cos = -(F.relu(-true_cos * 0.5 + 0.5) * (1.0 - anneal) + F.relu(-true_cos) * anneal)
    nxt = sdf + iter_cos * deltas * 0.5
    prv = sdf - iter_cos * deltas * 0.5
    pc, nc = torch.sigmoid(prv * inv_s), torch.sigmoid(nxt * inv_s)
    return ((pc - nc + 1e-5) / (pc + 1e-5)).clip(0.0, 1.0).squeeze(-1)


def neus_weights(sdf, gradients, dirs, deltas, inv_s, anneal):
    return weights_from_alphas(neus_alphas(sdf, gradients, dirs, deltas, inv_s, anneal)).unsqueeze(-1)


def density_weights(density, deltas):
    """rays.py:138-151 + 201-217; density, deltas [R,S,1]"""
    alphas = 1 - torch.exp(-(deltas * density))
    t = torch.cumprod(torch.cat([torch.ones(alphas.shape[0], 1, 1), 1.0 - alphas + 1e-7], 1), 1)
    return alphas * t[:, :-1, :]


# ------------------------------------------------------------------------------------------------
# A17 polarization — field_heads.py:90-106; polarizer.py:39-101
# ------------------------------------------------------------------------------------------------
def polarization_post(stokes, directions, up):
    s0 = F.leaky_relu(stokes[..., 0:1])
    stokes = torch.cat([s0, stokes[..., 1:]], -1)
    z = torch.tensor([0., 0., 1.])[None].expand(directions.shape)
    nrm = F.normalize(torch.linalg.cross(directions, z), dim=-1)
    cos_t = torch.clamp(torch.sum(nrm * up, dim=-1), min=-1 + 1e-4, max=1 - 1e-4)
    theta = torch.acos(cos_t) - np.pi / 2
    c, s = torch.cos(2 * theta), torch.sin(2 * theta)
    one, zero = torch.ones_like(c), torch.zeros_like(c)
    rot = torch.stack([one, zero, zero, zero, c, s, zero, -s, c], -1).view(-1, 3, 3)
    aligned = (rot @ stokes[..., None]).squeeze(-1)
    m = 0.5 * torch.tensor([[1., 1., 0.], [1., 0., 1.], [1., -1., 0.], [1., 0., -1.]])
    return (m[None] @ aligned[..., None]).squeeze(-1)


# ------------------------------------------------------------------------------------------------
# A19 compositing — renderers.py:149-243
# ------------------------------------------------------------------------------------------------
def composite(weights, values, background):
    comp = torch.sum(weights * values, dim=-2)
    acc = torch.sum(weights, dim=-2)
    return comp + (background * (1.0 - acc) if background is not None else 0.0)


# ------------------------------------------------------------------------------------------------
# A21/A22 — raw_pipeline.py:112-122; datasets.py:229-250; losses.py:97-164,213-265
# ------------------------------------------------------------------------------------------------
def mosaick_band(coords, pattern):
    pattern = torch.as_tensor(pattern)
    ph, pw = pattern.shape
    return pattern[coords[:, 1].long() % ph, coords[:, 2].long() % pw].long()


def mosaick_select(coords, pattern, rendered):
    band = mosaick_band(coords, pattern)
    return band, torch.gather(rendered, 1, band[:, None])


def l1_loss(output, target, sat_threshold=None):
    if sat_threshold is not None:
        m = target > sat_threshold
        if m.any():
            output = output.masked_fill(m, target[m].flatten()[0])
    return F.l1_loss(output, target)


def decimated_l1_loss(output, target, channel_draw, sat_threshold=None):
    """Loss.forward with per_channel_probability (losses.py:87-105) exactly as the reference evaluates it: the index
    shapes [n] and [n, 1] broadcast to an [n, n] gather (element (i, j) = pixel j at the channel drawn for pixel i)."""
    if sat_threshold is not None:
        m = target > sat_threshold
        if m.any():
            output = output.masked_fill(m, target[m].flatten()[0])
    idx = channel_draw.view(-1, 1)
    rows = torch.arange(output.shape[0])
    return F.l1_loss(output[rows, idx.view(-1, 1)], target[rows, idx.view(-1, 1)])


def eikonal_loss(gradients):
    n = torch.norm(gradients, 2, dim=-1)
    return F.mse_loss(n, torch.ones_like(n))


def curvature_loss(hessians):
    lap = hessians.sum(dim=-1)
    return F.l1_loss(lap, torch.zeros_like(lap))


# ------------------------------------------------------------------------------------------------
# the whole model — models/base_model.py:83-161 and the modules it calls (grid presets)
# ------------------------------------------------------------------------------------------------
class GridModelOracle:
    """Functional forward of the `grid` / `grid_raw` presets over a reference-layout state dict.

    cfg keys: modalities {name: channels}, log2_hashmap_size, num_levels, features_per_level, min_res, max_res,
    radius, num_samples, num_samples_importance, num_upsample_steps, bg_samples, base_variance,
    dir_encoding ("nerf"|"sh"), use_n_dot_v, use_reflection_direction, compute_hessian.
    state: level (coarse-to-fine), delta (numerical_gradients_delta), anneal.
    """

    def __init__(self, sd, cfg):
        self.sd, self.cfg = sd, cfg
        self.res = hash_resolutions(cfg["min_res"], cfg["max_res"], cfg["num_levels"])
        self.level, self.delta, self.anneal = cfg["num_levels"], 2.0 / cfg["max_res"], 1.0
        self.training = True

    def set_schedule_state(self, level, delta, anneal):
        self.level, self.delta, self.anneal = level, delta, anneal

    def _mask(self):
        m = torch.ones(self.cfg["num_levels"] * self.cfg["features_per_level"])
        m[self.level * self.cfg["features_per_level"]:] = 0
        return m

    def _grid_mlp(self, prefix, inputs, hidden_act, out_act, beta=1.0, radius=None):
        """FeatureGridAndMLP.forward (feature_structures.py:153-169)"""
        feats = hash_encode(inputs[..., :3], self.sd[prefix + ".feature_grid.encoding.hash_table"], self.res,
                            self.cfg["log2_hashmap_size"], radius=self.cfg["radius"] if radius is None else radius,
                            mask=self._mask())
        return mlp(self.sd, prefix + ".mlp_head", torch.cat([inputs, feats], -1), 3, hidden_act, out_act, beta)

    def sdf_field(self, x):
        """SDFField.forward (surface_field.py:99-116)"""
        if self.cfg.get("field") == "mlp":
            # presets `mlp*` (method_configs.py:304-335): PE -> 8 x 256 MLP, skip at layer 4, Softplus(beta=100)
            out = mlp(self.sd, "surface_model.surface_field.field", nerf_encode(x, 6, 0.0, 5), 8, "Softplus", "None", 100.0,
                      skips=(4,))
            return out[..., :1], out[..., 1:]
        out = self._grid_mlp("surface_model.surface_field.field", nerf_encode(x, 6, 0.0, 5), "Softplus", "None", 100.0)
        return out[..., :1], out[..., 1:]

    def inv_s(self):
        s = self.sd["surface_model.volume_rendering.density_fn.variance_network.s"]
        return torch.exp(s * 10.0).clip(1e-6, 1e6)

    def sample(self, o, d, nears, fars, rand_uniform=None, rand_pdf=None):
        """NeuSSampler.generate_ray_samples (ray_samplers.py:448-514) -> spacing bins [R, S+1]"""
        c = self.cfg
        bins, _ = spaced_bins(nears, fars, c["num_samples"], rand_uniform)
        k = c["num_samples_importance"] // c["num_upsample_steps"]
        sdf, new_bins, index = None, bins, None
        trace = []
        for it in range(c["num_upsample_steps"]):
            with torch.no_grad():
                t = spacing_to_euclid(new_bins[:, :-1], nears, fars)
                new_sdf = self.sdf_field((o[:, None] + d[:, None] * t[..., None]).reshape(-1, 3))[0].reshape(t.shape)
                sdf = new_sdf if sdf is None else torch.gather(torch.cat([sdf, new_sdf], -1), 1, index)
                u = make_u(o.shape[0], k, None if rand_pdf is None else rand_pdf[it])
                r = upsample_round(bins, sdf, u, nears, fars, c["base_variance"] * 2 ** it)
                new_bins, bins, index = r["new_bins"], r["merged_bins"], r["merged_index"]
                trace.append(r)
        return bins.detach(), trace

    def forward_modality(self, mod, o, d, up, rand=None, heads=None):
        c, sd = self.cfg, self.sd
        rand = rand or {}
        nears, fars, mask = sphere_collide(o, d, c["radius_collider"])
        oi, di, upi, ni, fi = o[mask], d[mask], up[mask], nears[mask], fars[mask]
        bins, _ = self.sample(oi, di, ni, fi, rand.get("uniform"), rand.get("pdf"))
        out = {}
        # background (background_model.py:73-111) on all rays
        bn, bf = background_near_far(nears, fars, mask)
        bbins, be = spaced_bins(bn, bf, c["bg_samples"], rand.get("background"), disparity=True)
        bstarts, bends = be[:, :-1, None], be[:, 1:, None]
        bpos = (o[:, None] + d[:, None] * bstarts).reshape(-1, 3)
        mag = torch.linalg.norm(bpos, ord=float("inf"), dim=-1, keepdim=True)
        bpos = torch.where(mag >= 1, (2 - 1 / mag) * (bpos / mag), bpos)
        bdirs = d[:, None].expand(-1, c["bg_samples"], -1).reshape(-1, 3)
        bup = up[:, None].expand(-1, c["bg_samples"], -1).reshape(-1, 3)
        p = "background_model.background_field"
        if c.get("bg_grid"):
            # preset grid_raw_grid_bg_unbalanced (method_configs.py:428-445): hash-grid background field of radius 2,
            # fed cat[x, PE(x)[3:], hash(x)] (nerf_field.py:92-96 -> feature_structures.py:153-166)
            feat = self._grid_mlp(p + ".base_field", nerf_encode(bpos, 6, 0.0, 5), "ReLU", "ReLU", radius=c["bg_radius"])
        else:
            feat = mlp(sd, p + ".base_field", nerf_encode(bpos, 6, 0.0, 5), 4, "ReLU", "ReLU")
        density = mlp(sd, p + ".density_head.field", feat, 1, "ReLU", "Softplus")
        bfeat = mlp(sd, p + ".head_field", torch.cat([feat, nerf_encode(bdirs, 4, 0.0, 3)], -1), 4, "ReLU", "ReLU")
        bw = density_weights(density.view(-1, c["bg_samples"], 1), bends - bstarts)
        head_list = heads if heads is not None else list(c["modalities"])
        bg = {}
        for h in head_list:
            hp = f"background_model.modality_heads.{h}.field"
            nl = 3 if c.get("bg_grid") else 1         # the grid-background preset copies the radiance heads
            if h == "polarization":
                v = polarization_post(mlp(sd, hp, bfeat, nl, "ReLU", "None"), bdirs, bup)
            else:
                v = mlp(sd, hp, bfeat, nl, "ReLU", "Sigmoid")
            bg[h] = torch.sum(bw * v.view(-1, c["bg_samples"], v.shape[-1]), dim=1)
        # surface (surface_model.py:66-127)
        e = spacing_to_euclid(bins, ni, fi)
        starts, ends = e[:, :-1, None], e[:, 1:, None]
        s = starts.shape[1]
        pos = (oi[:, None] + di[:, None] * starts).reshape(-1, 3)
        if c.get("field") == "mlp":
            # use_numerical_gradients=False: autograd gradient with create_graph (surface_model.py:77,193-203); no Hessian
            if not pos.requires_grad:
                pos.requires_grad_(True)
            with torch.enable_grad():
                sdf, geo = self.sdf_field(pos)
                g = torch.autograd.grad(sdf, pos, torch.ones_like(sdf), create_graph=True, retain_graph=True)[0]
            hess, normals = None, F.normalize(g, p=2, dim=-1)
        else:
            sdf, geo = self.sdf_field(pos)
            delta = self.delta / np.sqrt(3)
            sdf_t = torch.stack([self.sdf_field(pos + TAPS[i].to(pos.device) * delta)[0] for i in range(4)], 0)
            want_h = self.training and c["compute_hessian"]
            g, hess, normals = taps_gradients(sdf, sdf_t, delta, want_h)
        g3, n3 = g.view(-1, s, 3), normals.view(-1, s, 3)
        w = neus_weights(sdf.view(-1, s, 1), g3, di[:, None], ends - starts, self.inv_s(), self.anneal)
        # radiance (radiance_model.py:94-151)
        dirs = di[:, None].expand(-1, s, -1).reshape(-1, 3)
        ups = upi[:, None].expand(-1, s, -1).reshape(-1, 3)
        nd = n3.detach().reshape(-1, 3)
        add = [geo]
        ndv = None
        if c["use_n_dot_v"]:
            ndv = torch.sum(nd * -dirs, dim=-1, keepdim=True)
            add.append(ndv)
        dir_in = dirs
        if c["use_reflection_direction"]:
            if ndv is None:
                ndv = torch.sum(nd * -dirs, dim=-1, keepdim=True)
            dir_in = 2 * (ndv * nd) + dir_in
        dir_in = nerf_encode(dir_in, 4, 0.0, 3) if c["dir_encoding"] == "nerf" else sh_encode(5, dir_in)
        rin = torch.cat([pos, dir_in, torch.cat(add, -1)], -1)
        if c.get("field") == "mlp":
            rfeat = mlp(sd, "radiance_model.radiance_field.base_field", rin, 8, "ReLU", "ReLU", skips=(4,))
        else:
            rfeat = self._grid_mlp("radiance_model.radiance_field.base_field", rin, "ReLU", "ReLU")
        for h in head_list:
            hp = f"radiance_model.modality_heads.{h}.field"
            if h == "polarization":
                v = polarization_post(mlp(sd, hp, rfeat, 3, "ReLU", "None"), dirs, ups)
            else:
                v = mlp(sd, hp, rfeat, 3, "ReLU", "Sigmoid")
            col = bg[h].clone()
            col[mask] = composite(w, v.view(-1, s, v.shape[-1]), bg[h][mask])   # renderers.py:102-106
            out[h] = col
        steps = (starts + ends) / 2
        nrm = torch.zeros(o.shape[0], 3); nrm[mask] = torch.sum(w * n3, dim=-2)
        dep = torch.zeros(o.shape[0], 1); dep[mask] = torch.clip(torch.sum(w * steps, dim=-2), steps.min(), steps.max())
        acc = torch.zeros(o.shape[0], 1); acc[mask] = torch.sum(w, dim=-2)
        out.update({"normals": nrm, "depth": dep, "accumulation": acc, "gradients": g3,
                    "hessians": hess.view(-1, s, 3) if hess is not None else None, "weights": w, "bins": bins,
                    "mask": mask, "sdf": sdf.view(-1, s)})
        return out

    def loss(self, outputs, targets, coords, patterns, curvature_weight, sat_threshold=0.998):
        """LossManager.compute_loss (losses.py:213-265) with the raw channel select (raw_pipeline.py:112-122)."""
        total = 0.0
        losses = {}
        for mod in self.cfg["modalities"]:
            rendered = outputs[mod][mod]
            if patterns is not None:
                _, rendered = mosaick_select(coords[mod], patterns[mod], rendered)
            losses[mod] = l1_loss(rendered, targets[mod], sat_threshold if mod == "polarization" else None)
            total = total + losses[mod]
        g = torch.cat([outputs[m]["gradients"] for m in self.cfg["modalities"]], 0)
        losses["eikonal_loss"] = eikonal_loss(g)
        total = total + 0.1 * losses["eikonal_loss"]
        if self.cfg.get("field") != "mlp" and all(outputs[m]["hessians"] is not None for m in self.cfg["modalities"]):
            h = torch.cat([outputs[m]["hessians"] for m in self.cfg["modalities"]], 0)
            losses["curvature_loss"] = curvature_loss(h)
            total = total + curvature_weight * losses["curvature_loss"]
        return losses, total


def default_cfg(modalities=None, log2_hashmap_size=19, num_samples=32, num_samples_importance=32, bg_samples=16,
                dir_encoding="nerf", bg_grid=False, field="grid"):
    """confs/grid_raw.yaml over the `grid_raw` preset with the tcnn-free substitutions of SURVEY §8c."""
    return dict(modalities=modalities or {"rgb": 3, "infrared": 1, "mono": 1, "polarization": 4, "multispectral": 9},
                log2_hashmap_size=log2_hashmap_size, num_levels=16, features_per_level=2, min_res=16, max_res=1024,
                radius=1.0, radius_collider=1.0, num_samples=num_samples, num_samples_importance=num_samples_importance,
                num_upsample_steps=4, bg_samples=bg_samples, base_variance=64, dir_encoding=dir_encoding,
                use_n_dot_v=True, use_reflection_direction=False, compute_hessian=(field != "mlp"), bg_grid=bg_grid, bg_radius=2.0,
                field=field)
