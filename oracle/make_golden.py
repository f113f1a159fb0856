"""TEST INFRASTRUCTURE — generates tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference/src, imported through oracle/ref_harness.py) on seeded synthetic inputs, on the CPU
of the build container.  The fixtures pin the oracle (oracle/mms_oracle.py) and the CUDA path.

    python oracle/make_golden.py            # rewrites every fixture

The reference cannot travel to the GPU box; these committed vectors (a few hundred KB) do.
"""
import math
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from ref_harness import MODALITY_CHANNELS, build_reference_model, import_reference, set_schedule_state  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
PATTERNS = {"rgb": [[1, 2], [0, 1]], "infrared": [[0]], "mono": [[0]], "polarization": [[2, 1], [3, 0]],
            "multispectral": [[4, 5, 6], [2, 1, 0], [3, 8, 7]]}


def save(name, **arrays):
    conv = {}
    for k, v in arrays.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        conv[k] = np.asarray(v)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **conv)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB", {k: v.shape for k, v in conv.items()})


def synth_rays(n, seed, spread=0.45):
    g = torch.Generator().manual_seed(seed)
    o = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1) * 2.5
    d = torch.nn.functional.normalize(-o + spread * torch.randn(n, 3, generator=g), dim=-1)
    up = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1)
    return o, d, up


def golden_hashgrid():
    from field_components.encodings import HashEncodingConfig, HashEncoding
    from field_components.feature_structures import FeatureGridConfig
    cfg = HashEncodingConfig(max_res=1024, log2_hashmap_size=10, interpolation="Linear", implementation="torch")
    torch.manual_seed(11)
    enc = cfg.setup(in_dim=3)
    g = torch.Generator().manual_seed(12)
    x = torch.rand(192, 3, generator=g) * 1.1 - 0.05          # slightly outside [0,1] like the tap positions
    x[:8] = torch.tensor([[0.0, 0.0, 0.0], [1.0, 1.0, 1.0], [0.5, 0.25, 0.75], [0.5, 0.5, 0.5],
                          [0.0625, 0.125, 1.0], [-0.03125, 0.5, 1.03125], [1.0, 0.0, 0.5], [0.25, 0.25, 0.25]])
    x.requires_grad_(True)
    scaled = x[..., None, :] * enc.scalings.view(-1, 1)
    sc, sf = torch.ceil(scaled).type(torch.int32), torch.floor(scaled).type(torch.int32)
    cat = lambda a, b, c: torch.cat([a[..., 0:1], b[..., 1:2], c[..., 2:3]], dim=-1)
    idx = torch.stack([enc.hash_fn(sc), enc.hash_fn(cat(sc, sf, sc)), enc.hash_fn(cat(sf, sf, sc)),
                       enc.hash_fn(cat(sf, sc, sc)), enc.hash_fn(cat(sc, sc, sf)), enc.hash_fn(cat(sc, sf, sf)),
                       enc.hash_fn(sf), enc.hash_fn(cat(sf, sc, sf))], dim=-1)
    feats = enc(x)
    r = torch.randn(feats.shape, generator=g)
    (feats * r).sum().backward()
    # FeatureGrid (rescale + mask at level 5)
    torch.manual_seed(11)
    fg = FeatureGridConfig(encoding=cfg, radius=1.0).setup(input_dim=3)
    fg.update_mask(5)
    xw = (x.detach() * 2 - 1)
    fg_feats = fg(xw)
    save("hashgrid", x=x, indices=idx.to(torch.int32), features=feats, cotangent=r, dtable=enc.hash_table.grad,
         dx=x.grad, table_seed=11, log2_hashmap_size=10, resolutions=enc.scalings, fg_x=xw, fg_features=fg_feats,
         fg_level=5)


def golden_encodings():
    from field_components.encodings import NeRFEncodingConfig
    from utils.math import components_from_spherical_harmonics
    g = torch.Generator().manual_seed(21)
    x = (torch.rand(64, 3, generator=g) * 3 - 1.5).requires_grad_(True)
    pe6 = NeRFEncodingConfig(num_frequencies=6, min_freq_exp=0.0, max_freq_exp=5).setup(in_dim=3)
    pe4 = NeRFEncodingConfig(num_frequencies=4, min_freq_exp=0.0, max_freq_exp=3).setup(in_dim=3)
    y6, y4 = pe6(x), pe4(x)
    r6 = torch.randn(y6.shape, generator=g)
    (y6 * r6).sum().backward()
    d = torch.nn.functional.normalize(torch.randn(64, 3, generator=g), dim=-1)
    sh = components_from_spherical_harmonics(5, d)
    save("encodings", x=x, pe6=y6, pe4=y4, cot6=r6, dx6=x.grad, dirs=d, sh5=sh)


def golden_mlp():
    from field_components.mlp import MLPConfig
    g = torch.Generator().manual_seed(31)
    out = {}
    for name, cfg, din, dout in [
        ("sdf", MLPConfig(num_layers=3, hidden_dim=48, activation="Softplus", activation_params={"beta": 100},
                          out_activation="None", geometric_init=True, geometric_init_bias=0.4, weight_norm=True), 23, 20),
        ("rad", MLPConfig(num_layers=3, hidden_dim=40, out_activation="ReLU", weight_norm=True), 37, 24),
        ("head", MLPConfig(num_layers=3, hidden_dim=16, out_activation="Sigmoid", weight_norm=True), 24, 9),
        ("dens", MLPConfig(num_layers=1, hidden_dim=64, weight_norm=True, out_activation="Softplus"), 24, 1),
    ]:
        torch.manual_seed(32)
        m = cfg.setup(input_dim=din, output_dim=dout)
        x = (torch.randn(50, din, generator=g) * 0.5).requires_grad_(True)
        y = m(x)
        r = torch.randn(y.shape, generator=g)
        (y * r).sum().backward()
        out[name + "_x"], out[name + "_y"], out[name + "_cot"], out[name + "_dx"] = x, y, r, x.grad
        for k, p in m.named_parameters():
            out[f"{name}_grad.{k}"] = p.grad
    save("mlp", seed=32, **out)


def golden_samplers():
    from cameras.rays import RayBundle
    from model_components.ray_samplers import (LinearDisparitySamplerConfig, NeuSSamplerConfig, UniformSamplerConfig)
    n = 40
    o, d, up = synth_rays(n, 41, spread=0.3)
    from model_components.scene_colliders import SphereCollider
    rb = RayBundle(camera_indices=torch.zeros(n, 1, dtype=torch.long), origins=o, directions=d, up_directions=up,
                   pixel_area=torch.ones(n, 1), directions_norm=torch.ones(n, 1))
    rb, mask = SphereCollider(1.0)(rb)
    out = dict(origins=o, directions=d, nears=rb.nears, fars=rb.fars, mask=mask)
    uni = UniformSamplerConfig(num_samples=32).setup(single_jitter=True)
    lin = LinearDisparitySamplerConfig(num_samples=16).setup()
    for tag, smp, ns in (("uni", uni, 32), ("disp", lin, 16)):
        smp.eval()
        s = smp({"m": rb}, num_samples=ns)["m"]
        out[tag + "_eval_sbins"] = torch.cat([s.spacing_starts[..., 0], s.spacing_ends[..., -1:, 0]], -1)
        out[tag + "_eval_ebins"] = torch.cat([s.frustums.starts[..., 0], s.frustums.ends[..., -1:, 0]], -1)
        smp.train()
        torch.manual_seed(42)
        s = smp({"m": rb}, num_samples=ns)["m"]
        torch.manual_seed(42)
        out[tag + "_rand"] = torch.rand((n, 1)) if tag == "uni" else torch.rand((n, ns + 1))
        out[tag + "_train_sbins"] = torch.cat([s.spacing_starts[..., 0], s.spacing_ends[..., -1:, 0]], -1)
        out[tag + "_train_ebins"] = torch.cat([s.frustums.starts[..., 0], s.frustums.ends[..., -1:, 0]], -1)
    # NeuS sampler with an analytic sdf (sphere of radius 0.6 with a ripple): pins A5+A6+A7 end to end
    def sdf_fn(samples):
        p = samples.frustums.get_start_positions()
        return p.norm(dim=-1, keepdim=True) - 0.6 + 0.02 * torch.sin(9.0 * p[..., 0:1])
    neus = NeuSSamplerConfig(num_samples=32, num_samples_importance=32).setup()
    for mode in ("eval", "train"):
        neus.train(mode == "train")
        torch.manual_seed(43)
        s = neus({"m": rb}, sdf_fn=sdf_fn)["ray_samples_per_modality"]["m"]
        out[f"neus_{mode}_sbins"] = torch.cat([s.spacing_starts[..., 0], s.spacing_ends[..., -1:, 0]], -1)
        out[f"neus_{mode}_ebins"] = torch.cat([s.frustums.starts[..., 0], s.frustums.ends[..., -1:, 0]], -1)
    torch.manual_seed(43)
    out["neus_rand_uniform"] = torch.rand((n, 1))
    out["neus_rand_pdf"] = torch.stack([torch.rand((n, 1)) for _ in range(4)], 0)
    # searchsorted / SURVEY appendix example + random rows
    g = torch.Generator().manual_seed(44)
    cdf = torch.cumsum(torch.rand(30, 20, generator=g), -1)
    cdf = torch.cat([torch.zeros(30, 1), cdf / cdf[:, -1:]], -1)
    cdf[0, :5] = torch.tensor([0, .2, .2, .7, 1.0])
    cdf[0, 5:] = 1.0
    u = torch.rand(30, 9, generator=g)
    u[0, :5] = torch.tensor([0, .2, .69999, .7, 1.0])
    u[1, :3] = cdf[1, 3:6]                                    # exact hits
    out["ss_cdf"], out["ss_u"] = cdf, u
    out["ss_inds"] = torch.searchsorted(cdf, u.contiguous(), side="right")
    save("samplers", **out)


def golden_raygen():
    from cameras.cameras import Cameras
    from cameras.camera_optimizers import CameraOptimizerConfig
    from model_components.ray_generators import RayGenerator
    g = torch.Generator().manual_seed(51)
    n_cam, n = 6, 48
    pos = torch.nn.functional.normalize(torch.randn(n_cam, 3, generator=g), dim=-1) * 2.5
    fwd = torch.nn.functional.normalize(-pos + 0.1 * torch.randn(n_cam, 3, generator=g), dim=-1)
    upv = torch.nn.functional.normalize(torch.randn(n_cam, 3, generator=g), dim=-1)
    right = torch.nn.functional.normalize(torch.linalg.cross(fwd, upv), dim=-1)
    upv = torch.linalg.cross(right, fwd)
    c2w = torch.cat([torch.stack([right, upv, -fwd], -1), pos[..., None]], -1)
    W, H = 613, 511
    intr = torch.tensor([600.0, 590.0, 300.5, 250.25])
    dist = torch.tensor([-0.1, 0.01, 0.0, 0.0, 1e-3, -1e-3])
    out = dict(c2w=c2w, intr=intr, dist=dist, width=W, height=H)
    coords = torch.stack([torch.randint(0, n_cam, (n,), generator=g), torch.randint(0, H, (n,), generator=g),
                          torch.randint(0, W, (n,), generator=g)], -1).int()
    out["coords"] = coords
    for tag, shared, use_dist in (("shared", True, True), ("percam", False, True), ("off", None, False)):
        cams = Cameras(camera_to_worlds=c2w, fx=float(intr[0]), fy=float(intr[1]), cx=float(intr[2]), cy=float(intr[3]),
                       width=W, height=H, distortion_params=dist if use_dist else None)
        if shared is None:
            opt = CameraOptimizerConfig(mode="off", modalities_to_optimize={"m": False}).setup(num_cameras=n_cam)
        else:
            opt = CameraOptimizerConfig(mode="SO3xR3", shared_optimization=shared, modalities_to_optimize={"m": True}
                                        ).setup(num_cameras=n_cam)
            with torch.no_grad():
                pa = opt.pose_adjustment["m"]
                pa.copy_(0.05 * torch.randn(pa.shape, generator=g))
                if not shared:
                    pa[0] = 0.0                                # exercises the clamped-angle branch
            out[tag + "_pose"] = opt.pose_adjustment["m"].detach().clone()
        rg = RayGenerator({"m": {"cameras": cams}}, opt, pixel_offset=0.0)
        rb = rg({"m": coords})["m"]
        for k in ("origins", "directions", "up_directions", "pixel_area", "directions_norm"):
            out[f"{tag}_{k}"] = getattr(rb, k)
        if shared is not None:
            ro, rd, ru = (torch.randn(n, 3, generator=g) for _ in range(3))
            ((rb.origins * ro).sum() + (rb.directions * rd).sum() + (rb.up_directions * ru).sum()).backward()
            out[tag + "_cot"] = torch.stack([ro, rd, ru], 0)
            out[tag + "_dpose"] = opt.pose_adjustment["m"].grad
    save("raygen", **out)


def golden_render():
    from cameras.rays import RayBundle
    from field_components.field_heads import PolarizationHeadConfig
    from model_components.renderers import RadianceRenderer, DepthRenderer, NormalsRenderer, AccumulationRenderer
    from model_components.volume_rendering import NeuSVolumeRenderingConfig, NeuSDensityConfig
    g = torch.Generator().manual_seed(61)
    n, s = 24, 64
    o, d, up = synth_rays(n, 62)
    rb = RayBundle(camera_indices=torch.zeros(n, 1, dtype=torch.long), origins=o, directions=d, up_directions=up,
                   pixel_area=torch.ones(n, 1), directions_norm=torch.ones(n, 1))
    edges = torch.sort(torch.rand(n, s + 1, generator=g) * 2 + 1.5, -1)[0]
    samples = rb.get_ray_samples(bin_starts=edges[:, :-1, None], bin_ends=edges[:, 1:, None])
    sdf = (torch.randn(n, s, 1, generator=g) * 0.05 + torch.linspace(0.3, -0.3, s)[None, :, None]).requires_grad_(True)
    grad = torch.nn.functional.normalize(torch.randn(n, s, 3, generator=g), dim=-1) * (1 + 0.1 * torch.randn(n, s, 1, generator=g))
    grad.requires_grad_(True)
    out = dict(dirs=d, up=up, edges=edges, sdf=sdf, grad=grad)
    for tag, anneal in (("a1", 1.0), ("a03", 0.3)):
        vr = NeuSVolumeRenderingConfig(density_fn=NeuSDensityConfig()).setup()
        vr.set_cos_anneal_ratio(anneal)
        w = vr(samples, sdf, gradients=grad)
        cot = torch.randn(w.shape, generator=g)
        gs = torch.autograd.grad((w * cot).sum(), [sdf, grad, vr.density_fn.variance_network.s])
        out.update({f"{tag}_weights": w, f"{tag}_cot": cot, f"{tag}_dsdf": gs[0], f"{tag}_dgrad": gs[1], f"{tag}_ds": gs[2]})
    w = out["a1_weights"].detach().requires_grad_(True)
    vals = torch.rand(n, s, 9, generator=g).requires_grad_(True)
    bg = torch.rand(n, 9, generator=g).requires_grad_(True)
    col = RadianceRenderer.render(vals, w, bg)
    cot = torch.randn(col.shape, generator=g)
    gs = torch.autograd.grad((col * cot).sum(), [w, vals, bg])
    steps = (samples.frustums.starts + samples.frustums.ends) / 2
    out.update(comp_values=vals, comp_bg=bg, comp_color=col, comp_cot=cot, comp_dw=gs[0], comp_dvalues=gs[1], comp_dbg=gs[2],
               comp_depth=DepthRenderer.render(steps, w), comp_normals=NormalsRenderer.render(grad, w),
               comp_acc=AccumulationRenderer.render(w))
    # density weights (background path)
    dens = torch.rand(n, s, 1, generator=g) * 3
    dens.requires_grad_(True)
    bw = samples.get_weights_from_alphas(samples.get_alphas(dens))
    cot = torch.randn(bw.shape, generator=g)
    out.update(bg_density=dens, bg_weights=bw, bg_cot=cot, bg_ddensity=torch.autograd.grad((bw * cot).sum(), dens)[0])
    # polarization head post-processing
    torch.manual_seed(63)
    ph = PolarizationHeadConfig().setup(input_dim=16, output_dim=4)
    feat = torch.randn(n, 16, generator=g)
    stokes = ph.field(feat)
    out.update(pol_stokes=stokes, pol_out=ph(feat, directions=d, up_directions=up))
    save("render", **out)


def golden_losses():
    import_reference()
    from pipelines.raw_pipeline import RawPipeline
    from model_components.losses import LossConfig, SkipSaturationLossConfig, EikonalLossConfig, CurvatureLossConfig
    g = torch.Generator().manual_seed(71)
    n = 64
    H, W = 37, 41
    out = {}
    masks, coords, rendered, targets = {}, {}, {}, {}
    for mod, c in MODALITY_CHANNELS.items():
        pat = torch.tensor(PATTERNS[mod])
        masks[mod] = pat.repeat((math.ceil(H / pat.shape[0]), math.ceil(W / pat.shape[1])))[:H, :W].type(torch.int8)
        coords[mod] = torch.stack([torch.randint(0, 5, (n,), generator=g), torch.randint(0, H, (n,), generator=g),
                                   torch.randint(0, W, (n,), generator=g)], -1).int()
        rendered[mod] = torch.rand(n, c, generator=g).requires_grad_(True)
        targets[mod] = torch.rand(n, 1, generator=g)
    targets["polarization"][::7] = 1.0
    stub = types.SimpleNamespace(datamanager=types.SimpleNamespace(
        modalities=dict(MODALITY_CHANNELS), train_dataset=types.SimpleNamespace(mosaick_mask_per_modality=masks)))
    outputs = {mod: {mod: rendered[mod]} for mod in MODALITY_CHANNELS}
    sel = RawPipeline.select_right_channel_per_pixel(stub, coords, outputs)
    for mod in MODALITY_CHANNELS:
        out[f"{mod}_coords"], out[f"{mod}_rendered"], out[f"{mod}_target"] = coords[mod], rendered[mod], targets[mod]
        out[f"{mod}_selected"] = sel[mod][mod]
        lf = (SkipSaturationLossConfig(saturation_threshold=0.998) if mod == "polarization" else LossConfig()).setup(num_iterations=100)
        loss, weight = lf(sel[mod][mod], targets[mod], 10)
        out[f"{mod}_loss"] = loss
        out[f"{mod}_drendered"] = torch.autograd.grad(loss, rendered[mod])[0]
    grads = (torch.randn(20, 64, 3, generator=g) * 0.3 + torch.tensor([0.0, 0.0, 1.0])).requires_grad_(True)
    hess = torch.randn(20, 64, 3, generator=g).requires_grad_(True)
    eik, _ = EikonalLossConfig().setup(num_iterations=100)(grads, 10)
    lap = hess.sum(dim=-1)
    curv = torch.nn.functional.l1_loss(lap, torch.zeros_like(lap))
    out.update(geo_gradients=grads, geo_hessians=hess, eikonal=eik, curvature=curv,
               d_eikonal=torch.autograd.grad(eik, grads)[0], d_curvature=torch.autograd.grad(curv, hess)[0])
    save("losses", height=H, width=W, **out)


def golden_model(tag, level, delta, anneal, rays=10, seed_in=81, preset="grid_raw", yaml_name="grid_raw.yaml", modalities=None):
    """Whole BaseModel.forward + channel select + LossManager + backward of the real reference (default: grid_raw, 5
    modalities; `gridbg`: preset grid_raw_grid_bg_unbalanced with confs/grid_raw_rgb_all_views_pol_10_views.yaml, RGB +
    polarization)."""
    from cameras.rays import RayBundle
    from pipelines.raw_pipeline import RawPipeline
    MODALITY_CHANNELS = modalities or globals()["MODALITY_CHANNELS"]
    model, tc = build_reference_model(preset=preset, yaml_name=yaml_name, modalities=dict(MODALITY_CHANNELS), log2_hashmap_size=12)
    set_schedule_state(model, level=level, delta=delta, anneal=anneal)
    model.train()
    out = dict(log2_hashmap_size=12, level=level, delta=delta, anneal=anneal, seed=654824, preset=preset,
               modalities=",".join(MODALITY_CHANNELS))
    H, W = 64, 48
    g = torch.Generator().manual_seed(seed_in)
    inputs, coords, targets, masks_m = {}, {}, {}, {}
    for i, (mod, c) in enumerate(MODALITY_CHANNELS.items()):
        inputs[mod] = synth_rays(rays, seed_in + 1 + i)
        coords[mod] = torch.stack([torch.randint(0, 5, (rays,), generator=g), torch.randint(0, H, (rays,), generator=g),
                                   torch.randint(0, W, (rays,), generator=g)], -1).int()
        targets[mod] = torch.rand(rays, 1, generator=g)
        pat = torch.tensor(PATTERNS[mod])
        masks_m[mod] = pat.repeat((math.ceil(H / pat.shape[0]), math.ceil(W / pat.shape[1])))[:H, :W].type(torch.int8)
    if "polarization" in targets:
        targets["polarization"][::4] = 1.0
    bundles = {mod: RayBundle(camera_indices=torch.zeros(rays, 1, dtype=torch.long), origins=o.clone(), directions=d.clone(),
                              up_directions=up.clone(), pixel_area=torch.ones(rays, 1), directions_norm=torch.ones(rays, 1))
               for mod, (o, d, up) in inputs.items()}
    from model_components.scene_colliders import SphereCollider
    hit = {mod: SphereCollider(1.0)(RayBundle(camera_indices=None, origins=o, directions=d))[1] for mod, (o, d, up) in inputs.items()}
    torch.manual_seed(91)
    outputs = model(bundles)
    # the same draws, in the reference's call order
    torch.manual_seed(91)
    ru = {mod: torch.rand((int(hit[mod].sum()), 1)) for mod in MODALITY_CHANNELS}
    rp = {mod: torch.stack([torch.rand((int(hit[mod].sum()), 1)) for _ in range(4)], 0) for mod in MODALITY_CHANNELS}
    rbk = {mod: torch.rand((rays, 17)) for mod in MODALITY_CHANNELS}
    for mod in MODALITY_CHANNELS:
        o, d, up = inputs[mod]
        out.update({f"{mod}_origins": o, f"{mod}_directions": d, f"{mod}_up": up, f"{mod}_coords": coords[mod],
                    f"{mod}_target": targets[mod], f"{mod}_hit": hit[mod], f"{mod}_rand_uniform": ru[mod],
                    f"{mod}_rand_pdf": rp[mod], f"{mod}_rand_bg": rbk[mod]})
        for k, v in outputs[mod].items():
            if isinstance(v, torch.Tensor):
                out[f"{mod}_out_{k}"] = v
    stub = types.SimpleNamespace(datamanager=types.SimpleNamespace(
        modalities=dict(MODALITY_CHANNELS), train_dataset=types.SimpleNamespace(mosaick_mask_per_modality=masks_m)))
    outputs = RawPipeline.select_right_channel_per_pixel(stub, coords, outputs)
    lm = tc.pipeline.loss_manager.setup(modalities=list(MODALITY_CHANNELS), num_iterations=tc.max_num_iterations, model=model)
    step = 60000
    losses, total = lm.compute_loss(outputs, targets, coords, step)
    total.backward()
    for k, v in losses.items():
        out["loss_" + k] = torch.as_tensor(v)
    out["loss_total"] = total
    out["step"] = step
    gp = torch.Generator().manual_seed(99)
    for name, p in model.named_parameters():
        gr = p.grad if p.grad is not None else torch.zeros_like(p)
        if p.numel() <= 4096:
            out["grad." + name] = gr
        else:   # big tensors: norm + a seeded random projection
            r = torch.randn(p.shape, generator=gp)
            out["gradnorm." + name] = gr.norm()
            out["gradproj." + name] = (gr * r).sum()
    save("model_" + tag, height=H, width=W, **out)


def golden_gridbg():
    """BASELINE.json configs[3]: hash-grid background preset, RGB + polarization."""
    golden_model("gridbg", level=16, delta=2.0 / 1024, anneal=1.0, rays=10, seed_in=181, preset="grid_raw_grid_bg_unbalanced",
                 yaml_name="grid_raw_rgb_all_views_pol_10_views.yaml", modalities={"rgb": 3, "polarization": 4})


def golden_decimated():
    """Preset grid_decimated (method_configs.py:410-424): the per_channel_probability losses of the reference on seeded
    [n, C] renderings / targets, with the channel draws it made (torch.multinomial under the same seed)."""
    import_reference()
    from model_components.losses import LossConfig, SkipSaturationLossConfig
    g = torch.Generator().manual_seed(171)
    n = 48
    out = {}
    for mod, probs in (("rgb", [0.25, 0.5, 0.25]), ("multispectral", [0.1111] * 9), ("polarization", [0.25] * 4)):
        c = len(probs)
        rendered = torch.rand(n, c, generator=g).requires_grad_(True)
        target = torch.rand(n, c, generator=g)
        if mod == "polarization":
            target[::5, 1] = 1.0
        cfg = (SkipSaturationLossConfig(saturation_threshold=0.998, per_channel_probability=probs) if mod == "polarization"
               else LossConfig(per_channel_probability=probs))
        lf = cfg.setup(num_iterations=100)
        torch.manual_seed(172)
        loss, weight = lf(rendered, target, 10)
        torch.manual_seed(172)
        draw = torch.multinomial(torch.tensor(probs), n, replacement=True)
        out.update({f"{mod}_rendered": rendered, f"{mod}_target": target, f"{mod}_draw": draw, f"{mod}_loss": loss,
                    f"{mod}_drendered": torch.autograd.grad(loss, rendered)[0]})
    save("losses_decimated", **out)


def golden_mlp_raw():
    """BASELINE.json configs[0]: preset mlp_raw (MLP fields, autograd SDF gradients = double backward), RGB + mono."""
    golden_model("mlp_raw", level=16, delta=2.0 / 1024, anneal=1.0, rays=10, seed_in=281, preset="mlp_raw",
                 yaml_name="mlp_raw.yaml", modalities={"rgb": 3, "mono": 1})


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    import_reference()
    if len(sys.argv) > 1:
        for name in sys.argv[1:]:
            globals()["golden_" + name]()
        sys.exit(0)
    golden_hashgrid()
    golden_encodings()
    golden_mlp()
    golden_samplers()
    golden_raygen()
    golden_render()
    golden_losses()
    golden_model("late", level=16, delta=2.0 / 1024, anneal=1.0)
    golden_model("early", level=1, delta=2.0 / 16, anneal=0.0)
    golden_gridbg()
    golden_mlp_raw()
    golden_decimated()
