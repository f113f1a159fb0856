import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


class Golden(dict):
    """npz fixture produced by oracle/make_golden.py from the unmodified reference."""

    def t(self, key, device="cpu"):
        v = torch.from_numpy(np.asarray(self[key]))
        return v.to(device)


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return Golden({k: z[k] for k in z.files})


@pytest.fixture(scope="session")
def golden():
    return load_golden


def rel_err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


def assert_close(a, b, rtol=1e-5, atol=None, what=""):
    """max |a-b| <= rtol * max|b| (+ atol): the 1e-5 relative fp32 band of BASELINE.json."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    assert a.shape == b.shape, f"{what}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    tol = rtol * float(b.abs().max()) + (atol or 0.0)
    err = float((a - b).abs().max()) if a.numel() else 0.0
    assert err <= tol, f"{what}: max abs err {err:.3e} > tol {tol:.3e} (rtol {rtol})"


def assert_close_but_kinks(a, b, rtol, max_frac, what=""):
    """Like assert_close, but a fraction max_frac of the elements may miss the band: a pre-activation that lands
    within the rounding error of 0 flips relu' for its unit, which changes that row's gradient by O(1)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    assert a.shape == b.shape, f"{what}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    tol = rtol * float(b.abs().max())
    bad = float(((a - b).abs() > tol).double().mean()) if a.numel() else 0.0
    assert bad <= max_frac, f"{what}: {bad:.2e} of the elements miss tol {tol:.3e} (allowed {max_frac:.1e})"


@pytest.fixture(params=[0, 3], ids=["fp32simt", "3xtf32"])
def mlp_precision(request):
    """Runs a test once on the fp32 SIMT layers (the tight anchor) and once on the tcgen05 3xTF32 layers (default
    product path).  tcgen05 accumulators truncate (round toward zero) on every MMA, so a K=256 product carries a
    systematic ~2e-6 relative error (measured, tests/test_gpu_ops.py::test_tc_layer_products_vs_fp64): bands x3."""
    from multimodalstudio_b200 import ops
    old = ops.MLP_PRECISION
    ops.set_mlp_precision(request.param)
    yield request.param
    ops.set_mlp_precision(old)


def assert_close_elementwise(a, b, rtol, atol, what=""):
    """|a - b| <= rtol * |b| + atol for EVERY element (a relative band with an absolute floor), unlike assert_close
    whose band is scaled by max|b|."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    assert a.shape == b.shape, f"{what}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    excess = (a - b).abs() - (rtol * b.abs() + atol)
    worst = float(excess.max()) if a.numel() else 0.0
    assert worst <= 0, f"{what}: an element misses |a-b| <= {rtol}*|b| + {atol} by {worst:.3e} (max abs err {float((a - b).abs().max()):.3e})"


_MEASURED = os.path.join(ROOT, "gpurun_out", "parity_measured.jsonl")


def record_error(test, quantity, err, band):
    """Logs the MEASURED error of a parity check next to its band: printed (pytest -rP shows it) and appended to
    gpurun_out/parity_measured.jsonl when that directory exists (copied to profiles/ and quoted in README)."""
    import json
    line = json.dumps({"test": test, "quantity": quantity, "measured": float(err), "band": float(band)})
    print("[parity]", line)
    d = os.path.dirname(_MEASURED)
    if os.path.isdir(d):
        with open(_MEASURED, "a") as fh:
            fh.write(line + "\n")
