"""dev: hash-grid forward / backward at the step's size (2 621 440 look-ups; points along rays, i.e. spatially coherent,
and uniformly random)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodalstudio_b200 import ops
from multimodalstudio_b200.field_components import HashEncodingConfig
torch.manual_seed(0)
dev = "cuda"
enc = HashEncodingConfig(num_levels=16, min_res=16, max_res=1024, log2_hashmap_size=19, features_per_level=2, interpolation="Linear").setup(in_dim=3).to(dev)
desc, tab = enc.desc(1.0), enc.hash_table.detach()
n = 2621440
mask = torch.ones(32, device=dev)
def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
o = torch.nn.functional.normalize(torch.randn(n // 64, 1, 3, device=dev), dim=-1) * 0.9
d = torch.nn.functional.normalize(torch.randn(n // 64, 1, 3, device=dev), dim=-1)
rays = (o + d * torch.linspace(-0.5, 0.5, 64, device=dev)[None, :, None]).reshape(-1, 3).contiguous()
for name, pts in (("random", torch.rand(n, 3, device=dev) * 2 - 1), ("along rays", rays)):
    feat = torch.empty(n, 32, device=dev); dfeat = torch.randn(n, 32, device=dev)
    dtab = torch.zeros_like(tab); dpts = torch.empty(n, 3, device=dev)
    tf = timeit(lambda: ops.hashgrid_fwd_into(desc, pts, tab, mask, feat))
    tb = timeit(lambda: ops.hashgrid_bwd_from(desc, pts, tab, mask, dfeat, 0, dtab, dpts))
    tb2 = timeit(lambda: ops.hashgrid_bwd_from(desc, pts, tab, mask, dfeat, 0, dtab, None))
    print(f"{name:10s}: fwd {tf:.3f} ms ({n * 1024 / tf / 1e6:.0f} GB/s algorithmic) bwd+dx {tb:.3f} ms bwd {tb2:.3f} ms ({n * 2048 / tb2 / 1e6:.0f} GB/s)", flush=True)
