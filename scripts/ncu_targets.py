"""Target of the `ncu --set full` captures (profiles/): one warm-up + one measured launch of each dominant kernel at the
workload's sizes — the three tcgen05 layer products (2 097 152 rows, 256 -> 256, Softplus(100), 3xTF32) and the hash-grid
forward / backward (2 621 440 look-ups, 16 levels x 2^19 x 2 fp32, the centre + tap batch of one step).

    ncu --set full --clock-control none --import-source on -k regex:'tc_|hashgrid' --launch-skip 5 -c 5 -o prof python scripts/ncu_targets.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodalstudio_b200 import ops
from multimodalstudio_b200.field_components import HashEncodingConfig
torch.manual_seed(0)
dev = "cuda"
prec = 3
n, k, o = 2097152, 256, 256
x = torch.randn(n, k, device=dev); w = torch.randn(o, k, device=dev) * 0.1; b = torch.randn(o, device=dev)
y = torch.empty(n, o, device=dev); dz = torch.randn(n, o, device=dev); dx = torch.empty(n, k, device=dev)
dw = torch.zeros(o, k, device=dev); db = torch.zeros(o, device=dev)
pw = ops.pack_weight(w, False, prec); pwt = ops.pack_weight(w, True, prec)
enc = HashEncodingConfig(num_levels=16, min_res=16, max_res=1024, log2_hashmap_size=19, features_per_level=2, interpolation="Linear").setup(in_dim=3).to(dev)
pts = torch.rand(2621440, 3, device=dev)
mask = torch.ones(32, device=dev)
feat = torch.empty(pts.shape[0], 32, device=dev)
dfeat = torch.randn(pts.shape[0], 32, device=dev)
dtab = torch.zeros_like(enc.hash_table)
dpts = torch.empty_like(pts)
desc = enc.desc(1.0)
tab = enc.hash_table.detach()
for _ in range(2):
    ops.linear_fwd_tc(x, pw, b, o, 2, 100.0, prec, out=y)
    ops.linear_bwd_data_tc(dz, pwt, k, x, 2, 100.0, prec, out=dx)
    ops.linear_bwd_weight_tc(dz, x, dw, db, prec)
    ops.hashgrid_fwd_into(desc, pts, tab, mask, feat)
    ops.hashgrid_bwd_from(desc, pts, tab, mask, dfeat, 0, dtab, dpts)
    torch.cuda.synchronize()
print("ok")
