"""dev: does the radiance MLP's input row adopt the SDF network's geometry-feature buffer in a model forward?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodalstudio_b200 import ops
from multimodalstudio_b200.cameras import RayBundle
from multimodalstudio_b200.models import build_model
dev = "cuda"
mods = {"rgb": 3, "mono": 1}
model = build_model("grid_raw", modalities=mods, log2_hashmap_size=12, seed=3).to(dev)
model.set_schedule_state(16, 2.0 / 1024, 1.0)
model.train()
g = torch.Generator().manual_seed(1)
n = 64
bundles = {}
for m in mods:
    o = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1) * 2.5
    d = torch.nn.functional.normalize(-o + 0.3 * torch.randn(n, 3, generator=g), dim=-1)
    up = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1)
    bundles[m] = RayBundle(None, o.to(dev), d.to(dev), up.to(dev))
for it in range(3):
    ops.clear_pack_cache()
    out = model(bundles)
    loss = sum(v[m].sum() for m, v in out.items() if isinstance(v, dict) and m in v)
    loss.backward()
    print("pass", it, "adoptions so far", ops.ROW_ADOPTIONS, "hints", {k: v for k, v in ops._ROW_HINT.items()})
