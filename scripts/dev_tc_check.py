"""dev: tcgen05 layer kernels (mmsb_linear_*_tc) vs torch fp64 matmul (run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodalstudio_b200 import ops
torch.manual_seed(0)
dev = "cuda"
def rel(a, b): return float((a.double() - b.double()).abs().max() / b.double().abs().max())
shapes = [(4099, 71, 256), (4099, 256, 256), (300, 256, 64), (4099, 319, 256), (1000, 283, 128), (130000, 256, 256),
          (4099, 39, 256), (77, 256, 48), (50000, 64, 64)]
worst = 0.0
for prec in (3, 1):
    tol = 2e-5 if prec == 3 else 5e-3
    for n, k, o in shapes:
        ldx = (k + 3) // 4 * 4
        xb = torch.randn(n, ldx, device=dev); x = xb[:, :k]
        w = torch.randn(o, k, device=dev) * 0.1; b = torch.randn(o, device=dev)
        pw = ops.pack_weight(w, False, prec); pwt = ops.pack_weight(w, True, prec)
        y = ops.linear_fwd_tc(x, pw, b, o, 1, 1.0, prec)
        ref = torch.relu(x.double() @ w.double().T + b.double())
        e_f = rel(y, ref)
        dz = torch.randn(n, o, device=dev)
        dx = ops.linear_bwd_data_tc(dz, pwt, k, x, 1, 1.0, prec)
        e_d = rel(dx, (dz.double() @ w.double()) * (x.double() > 0))
        dw = torch.zeros(o, k, device=dev); db = torch.zeros(o, device=dev)
        ops.linear_bwd_weight_tc(dz, x, dw, db, prec)
        e_w = rel(dw, dz.double().T @ x.double()); e_b = rel(db, dz.double().sum(0))
        torch.cuda.synchronize()
        ok = max(e_f, e_d, e_w) < tol and e_b < 2e-5
        print(f"prec={prec} n={n} k={k} o={o}: fwd {e_f:.2e} dgrad {e_d:.2e} wgrad {e_w:.2e} bias {e_b:.2e} {'ok' if ok else 'FAIL'}", flush=True)
        if not ok: worst = 1.0
print("ALL OK" if worst == 0 else "SOME FAILED")
sys.exit(0 if worst == 0 else 1)
