"""dev: ncu target, the narrow first layer (71 -> 256, Softplus(100), 2 621 440 rows, padded rows) forward + its dgrad."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodalstudio_b200 import ops
n, k, o, prec = 2621440, 71, 256, 3
x = torch.randn(n, 72, device="cuda")[:, :k]; w = torch.randn(o, k, device="cuda") * 0.1; b = torch.randn(o, device="cuda")
y = torch.empty(n, o, device="cuda"); dz = torch.randn(n, o, device="cuda"); dx = torch.empty(n, 72, device="cuda")[:, :k]
pw = ops.pack_weight(w, False, prec); pwt = ops.pack_weight(w, True, prec)
for _ in range(2):
    ops.linear_fwd_tc(x, pw, b, o, 2, 100.0, prec, out=y)
    ops.linear_bwd_data_tc(dz, pwt, k, None, 0, 100.0, prec, out=dx)
    torch.cuda.synchronize()
print("ok")
