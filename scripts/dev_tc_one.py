"""dev: a few launches of the tcgen05 layer kernels at one shape (target of an ncu capture)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodalstudio_b200 import ops
prec = int(sys.argv[1]) if len(sys.argv) > 1 else 3
n, k, o = 524288, 256, 256
x = torch.randn(n, k, device="cuda"); w = torch.randn(o, k, device="cuda") * 0.1; b = torch.randn(o, device="cuda")
y = torch.empty(n, o, device="cuda"); dz = torch.randn(n, o, device="cuda"); dx = torch.empty(n, k, device="cuda")
dw = torch.zeros(o, k, device="cuda"); db = torch.zeros(o, device="cuda")
pw = ops.pack_weight(w, False, prec); pwt = ops.pack_weight(w, True, prec)
for _ in range(3):
    ops.linear_fwd_tc(x, pw, b, o, 1, 1.0, prec, out=y)
    ops.linear_bwd_data_tc(dz, pwt, k, x, 1, 1.0, prec, out=dx)
    ops.linear_bwd_weight_tc(dz, x, dw, db, prec)
torch.cuda.synchronize()
print("ok")
