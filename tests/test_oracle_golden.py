"""CPU: the oracle (oracle/mms_oracle.py) against the fixtures generated from the UNMODIFIED reference
(oracle/make_golden.py).  Integer outputs bit-exact; floats within 1e-5 relative (they are bit-identical
on the machine that generated them; the band absorbs BLAS / libm differences between hosts)."""
import numpy as np
import pytest
import torch

import mms_oracle as O
from conftest import assert_close, load_golden

MODS = {"rgb": 3, "infrared": 1, "mono": 1, "polarization": 4, "multispectral": 9}


def test_hashgrid_indices_and_features():
    g = load_golden("hashgrid")
    torch.manual_seed(int(g["table_seed"]))
    table = (torch.rand((2 ** 10) * 16, 2) * 2 - 1) * 0.001
    res = O.hash_resolutions(16, 1024, 16)
    assert torch.equal(res, g.t("resolutions"))
    x = g.t("x").requires_grad_(True)
    table.requires_grad_(True)
    idx, _ = O.hash_indices(x, res, 10)
    assert torch.equal(idx, g.t("indices").long())
    feats = O.hash_encode(x, table, res, 10)
    assert_close(feats, g.t("features"), what="features")
    (feats * g.t("cotangent")).sum().backward()
    assert_close(table.grad, g.t("dtable"), what="dtable")
    assert_close(x.grad, g.t("dx"), what="dx")
    mask = torch.ones(32)
    mask[int(g["fg_level"]) * 2:] = 0
    assert_close(O.hash_encode(g.t("fg_x"), table.detach(), res, 10, radius=1.0, mask=mask), g.t("fg_features"), what="fg")


def test_encodings():
    g = load_golden("encodings")
    x = g.t("x").requires_grad_(True)
    y6 = O.nerf_encode(x, 6, 0.0, 5)
    assert_close(y6, g.t("pe6"))
    assert_close(O.nerf_encode(x, 4, 0.0, 3), g.t("pe4"))
    (y6 * g.t("cot6")).sum().backward()
    assert_close(x.grad, g.t("dx6"))
    assert_close(O.sh_encode(5, g.t("dirs")), g.t("sh5"))


def test_samplers():
    g = load_golden("samplers")
    o, d = g.t("origins"), g.t("directions")
    nears, fars, mask = O.sphere_collide(o, d)
    assert torch.equal(mask, g.t("mask"))
    assert torch.equal(nears, g.t("nears")) and torch.equal(fars, g.t("fars"))
    for tag, ns, disp in (("uni", 32, False), ("disp", 16, True)):
        sb, eb = O.spaced_bins(nears, fars, ns, None, disp)
        assert torch.equal(sb, g.t(tag + "_eval_sbins")) and torch.equal(eb, g.t(tag + "_eval_ebins"))
        sb, eb = O.spaced_bins(nears, fars, ns, g.t(tag + "_rand"), disp)
        assert torch.equal(sb, g.t(tag + "_train_sbins")) and torch.equal(eb, g.t(tag + "_train_ebins"))
    inds = torch.searchsorted(g.t("ss_cdf"), g.t("ss_u").contiguous(), side="right")
    assert torch.equal(inds, g.t("ss_inds"))
    assert inds[0, :4].tolist() == [1, 3, 3, 4]         # SURVEY appendix A example (cdf 0,.2,.2,.7,1 ; u 0,.2,.69999,.7)

    def sdf(p):
        return p.norm(dim=-1) - 0.6 + 0.02 * torch.sin(9.0 * p[..., 0])

    for mode in ("eval", "train"):
        ru = g.t("neus_rand_uniform") if mode == "train" else None
        rp = g.t("neus_rand_pdf") if mode == "train" else None
        bins, _ = O.spaced_bins(nears, fars, 32, ru)
        cur, new_bins, index = None, bins, None
        for it in range(4):
            t = O.spacing_to_euclid(new_bins[:, :-1], nears, fars)
            ns_ = sdf(o[:, None] + d[:, None] * t[..., None])
            cur = ns_ if cur is None else torch.gather(torch.cat([cur, ns_], -1), 1, index)
            u = O.make_u(o.shape[0], 8, None if rp is None else rp[it])
            r = O.upsample_round(bins, cur, u, nears, fars, 64 * 2 ** it)
            new_bins, bins, index = r["new_bins"], r["merged_bins"], r["merged_index"]
        assert_close(bins, g.t(f"neus_{mode}_sbins"), rtol=1e-6, what=f"neus {mode} bins")
        assert_close(O.spacing_to_euclid(bins, nears, fars), g.t(f"neus_{mode}_ebins"), rtol=1e-6)


def test_raygen():
    g = load_golden("raygen")
    n_cam = g["c2w"].shape[0]
    intr = g.t("intr")[None].expand(n_cam, 4)
    for tag in ("shared", "percam", "off"):
        dist = None if tag == "off" else g.t("dist")[None].expand(n_cam, 6)
        pa = None if tag == "off" else g.t(tag + "_pose").clone().requires_grad_(True)
        r = O.raygen(g.t("coords"), g.t("c2w"), intr, dist, pa)
        for k in ("origins", "directions", "up_directions", "pixel_area", "directions_norm"):
            assert_close(r[k], g.t(f"{tag}_{k}"), what=f"{tag} {k}")
        if pa is not None:
            cot = g.t(tag + "_cot")
            ((r["origins"] * cot[0]).sum() + (r["directions"] * cot[1]).sum() + (r["up_directions"] * cot[2]).sum()).backward()
            assert_close(pa.grad, g.t(tag + "_dpose"), what=f"{tag} dpose")


def test_render():
    g = load_golden("render")
    edges = g.t("edges")
    starts, ends = edges[:, :-1, None], edges[:, 1:, None]
    dirs = g.t("dirs")[:, None]
    for tag, anneal in (("a1", 1.0), ("a03", 0.3)):
        sdf, grad = g.t("sdf").requires_grad_(True), g.t("grad").requires_grad_(True)
        s = torch.tensor([0.3], requires_grad=True)
        w = O.neus_weights(sdf, grad, dirs, ends - starts, torch.exp(s * 10.0).clip(1e-6, 1e6), anneal)
        assert_close(w, g.t(tag + "_weights"), what="weights")
        gs = torch.autograd.grad((w * g.t(tag + "_cot")).sum(), [sdf, grad, s])
        assert_close(gs[0], g.t(tag + "_dsdf")); assert_close(gs[1], g.t(tag + "_dgrad")); assert_close(gs[2], g.t(tag + "_ds"))
    w = g.t("a1_weights")
    assert_close(O.composite(w, g.t("comp_values"), g.t("comp_bg")), g.t("comp_color"))
    assert_close(O.density_weights(g.t("bg_density"), ends - starts), g.t("bg_weights"))
    assert_close(O.polarization_post(g.t("pol_stokes"), g.t("dirs"), g.t("up")), g.t("pol_out"))


def test_losses():
    g = load_golden("losses")
    pats = __import__("multimodalstudio_b200.models", fromlist=["x"]).MOSAICK_PATTERNS
    for mod in MODS:
        band, sel = O.mosaick_select(g.t(mod + "_coords"), pats[mod], g.t(mod + "_rendered"))
        assert torch.equal(sel, g.t(mod + "_selected"))
        loss = O.l1_loss(sel, g.t(mod + "_target"), 0.998 if mod == "polarization" else None)
        assert_close(loss, g.t(mod + "_loss"))
    assert_close(O.eikonal_loss(g.t("geo_gradients")), g.t("eikonal"))
    assert_close(O.curvature_loss(g.t("geo_hessians")), g.t("curvature"))


@pytest.mark.parametrize("tag", ["late", "early"])
def test_whole_model(tag):
    """BaseModel.forward + raw channel select + LossManager + backward, 5 modalities (grid_raw)."""
    from multimodalstudio_b200.models import MOSAICK_PATTERNS, build_model
    g = load_golden("model_" + tag)
    model = build_model("grid_raw", log2_hashmap_size=int(g["log2_hashmap_size"]), seed=int(g["seed"]))
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in model.state_dict().items()}
    orc = O.GridModelOracle(sd, O.default_cfg(log2_hashmap_size=int(g["log2_hashmap_size"])))
    orc.set_schedule_state(int(g["level"]), float(g["delta"]), float(g["anneal"]))
    outputs, coords, targets = {}, {}, {}
    for mod in MODS:
        rand = {"uniform": g.t(mod + "_rand_uniform"), "pdf": g.t(mod + "_rand_pdf"), "background": g.t(mod + "_rand_bg")}
        outputs[mod] = orc.forward_modality(mod, g.t(mod + "_origins"), g.t(mod + "_directions"), g.t(mod + "_up"), rand)
        coords[mod], targets[mod] = g.t(mod + "_coords"), g.t(mod + "_target")
        for k in list(MODS) + ["normals", "depth", "accumulation", "gradients", "hessians"]:
            assert_close(outputs[mod][k], g.t(f"{mod}_out_{k}"), rtol=2e-5, what=f"{tag} {mod} {k}")
    curv_w = float(g["loss_curvature_loss_weight"])
    losses, total = orc.loss(outputs, targets, coords, MOSAICK_PATTERNS, curv_w)
    assert_close(total, g.t("loss_total"), what="total loss")
    total.backward()
    for k in g:
        if k.startswith("grad."):
            name = k[5:]
            gr = sd[name].grad if sd[name].grad is not None else torch.zeros_like(sd[name])
            assert_close(gr, g.t(k), rtol=5e-5, atol=1e-9, what=k)
        elif k.startswith("gradnorm."):
            name = k[9:]
            assert_close(sd[name].grad.norm(), g.t(k), rtol=5e-5, what=k)


def test_whole_model_grid_background_preset():
    """BASELINE.json configs[3]: preset grid_raw_grid_bg_unbalanced (hash-grid background of radius 2, background heads
    copied from the radiance heads; method_configs.py:428-445) with confs/grid_raw_rgb_all_views_pol_10_views.yaml,
    RGB + polarization: the host mirror's seeded parameters + the oracle reproduce the reference's step."""
    from multimodalstudio_b200.models import MOSAICK_PATTERNS, build_model
    g = load_golden("model_gridbg")
    mods = {"rgb": 3, "polarization": 4}
    model = build_model("grid_raw_grid_bg_unbalanced", modalities=mods, log2_hashmap_size=int(g["log2_hashmap_size"]), seed=int(g["seed"]))
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in model.state_dict().items()}
    assert sd["background_model.background_field.base_field.feature_grid.encoding.hash_table"].shape == (16 * 4096, 2)
    assert sd["background_model.modality_heads.polarization.field.layers.2.bias"].shape == (3,)
    orc = O.GridModelOracle(sd, O.default_cfg(modalities=mods, log2_hashmap_size=int(g["log2_hashmap_size"]), bg_grid=True))
    orc.set_schedule_state(int(g["level"]), float(g["delta"]), float(g["anneal"]))
    outputs, coords, targets = {}, {}, {}
    for mod in mods:
        rand = {"uniform": g.t(mod + "_rand_uniform"), "pdf": g.t(mod + "_rand_pdf"), "background": g.t(mod + "_rand_bg")}
        outputs[mod] = orc.forward_modality(mod, g.t(mod + "_origins"), g.t(mod + "_directions"), g.t(mod + "_up"), rand)
        coords[mod], targets[mod] = g.t(mod + "_coords"), g.t(mod + "_target")
        for k in list(mods) + ["normals", "depth", "accumulation", "gradients", "hessians"]:
            # finite differences with delta' = 2/1024/sqrt(3) amplify BLAS-order noise of the sdf (thread count) by 220
            assert_close(outputs[mod][k], g.t(f"{mod}_out_{k}"), rtol=1e-4 if k in ("gradients", "hessians", "normals") else 2e-5,
                         what=f"gridbg {mod} {k}")
    losses, total = orc.loss(outputs, targets, coords, MOSAICK_PATTERNS, float(g["loss_curvature_loss_weight"]))
    assert_close(total, g.t("loss_total"), what="total loss")
    total.backward()
    n_checked = 0
    for k in g:
        if k.startswith("grad."):
            gr = sd[k[5:]].grad if sd[k[5:]].grad is not None else torch.zeros_like(sd[k[5:]])
            assert_close(gr, g.t(k), rtol=5e-4, atol=1e-9, what=k)      # BLAS thread-count reassociation x the 1/delta amplification
            n_checked += 1
        elif k.startswith("gradnorm."):
            assert_close(sd[k[9:]].grad.norm(), g.t(k), rtol=5e-4, what=k)
            n_checked += 1
    assert n_checked > 60


def test_whole_model_mlp_raw_preset():
    """BASELINE.json configs[0]: preset mlp_raw (PE + 8 x 256 MLP fields with a skip connection, SDF gradients by
    autograd.grad(create_graph=True) -> double backward in the reference, eikonal loss only; method_configs.py:302-400)
    with confs/mlp_raw.yaml, RGB + mono: the host mirror's seeded parameters + the oracle reproduce the reference's step."""
    from multimodalstudio_b200.models import MOSAICK_PATTERNS, build_model
    g = load_golden("model_mlp_raw")
    mods = {"rgb": 3, "mono": 1}
    model = build_model("mlp_raw", modalities=mods, seed=int(g["seed"]))
    assert sum(p.numel() for p in model.parameters()) == 1582101
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in model.state_dict().items()}
    orc = O.GridModelOracle(sd, O.default_cfg(modalities=mods, field="mlp"))
    orc.set_schedule_state(16, float(g["delta"]), float(g["anneal"]))
    outputs, coords, targets = {}, {}, {}
    for mod in mods:
        rand = {"uniform": g.t(mod + "_rand_uniform"), "pdf": g.t(mod + "_rand_pdf"), "background": g.t(mod + "_rand_bg")}
        outputs[mod] = orc.forward_modality(mod, g.t(mod + "_origins"), g.t(mod + "_directions"), g.t(mod + "_up"), rand)
        coords[mod], targets[mod] = g.t(mod + "_coords"), g.t(mod + "_target")
        for k in list(mods) + ["normals", "depth", "accumulation", "gradients"]:
            assert_close(outputs[mod][k], g.t(f"{mod}_out_{k}"), rtol=2e-5, what=f"mlp_raw {mod} {k}")
    losses, total = orc.loss(outputs, targets, coords, MOSAICK_PATTERNS, 0.0)
    assert "curvature_loss" not in losses
    assert_close(total, g.t("loss_total"), what="total loss")
    total.backward()
    n_checked = 0
    for k in g:
        if k.startswith("grad."):
            gr = sd[k[5:]].grad if sd[k[5:]].grad is not None else torch.zeros_like(sd[k[5:]])
            assert_close(gr, g.t(k), rtol=5e-4, atol=1e-9, what=k)      # BLAS thread-count reassociation x the 1/delta amplification
            n_checked += 1
        elif k.startswith("gradnorm."):
            assert_close(sd[k[9:]].grad.norm(), g.t(k), rtol=5e-4, what=k)
            n_checked += 1
    assert n_checked > 60


def test_decimated_losses():
    """Preset grid_decimated: the reference's per_channel_probability loss (an [n, n] broadcast gather, losses.py:87-105),
    its restatement, and the closed form the B200 mirror computes: (1/n) sum_j sum_c f_c |out - tgt|, f_c = draw frequency."""
    g = load_golden("losses_decimated")
    for mod in ("rgb", "multispectral", "polarization"):
        out, tgt, draw = g.t(mod + "_rendered").requires_grad_(True), g.t(mod + "_target"), g.t(mod + "_draw")
        thr = 0.998 if mod == "polarization" else None
        loss = O.decimated_l1_loss(out, tgt, draw, thr)
        assert_close(loss, g.t(mod + "_loss"), rtol=1e-6, what=mod)
        assert_close(torch.autograd.grad(loss, out)[0], g.t(mod + "_drendered"), rtol=1e-6, what=mod + " grad")
        n, c = out.shape
        freq = torch.bincount(draw, minlength=c).float() / n
        o2 = out.detach()
        if thr is not None and (tgt > thr).any():
            o2 = o2.masked_fill(tgt > thr, tgt[tgt > thr].flatten()[0])
        closed = ((o2 - tgt).abs() * freq[None]).sum() / n
        assert_close(closed, g.t(mod + "_loss"), rtol=1e-6, what=mod + " closed form")
