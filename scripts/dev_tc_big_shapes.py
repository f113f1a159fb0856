"""dev: the layer products at the north-star micro-batch's big shapes (ms per launch); compare library builds with MMSB_LIB."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodalstudio_b200 import ops
dev = "cuda"
def pad4(k): return (k + 3) // 4 * 4
def timeit(fn, reps=4):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
prec, act = 3, 2
tot = [0.0, 0.0, 0.0]
for n, k, o in [(10489600, 71, 256), (10489600, 256, 256), (2097920, 319, 256), (2097920, 256, 256), (419584, 256, 64)]:
    x = torch.randn(n, pad4(k), device=dev)[:, :k]; w = torch.randn(o, k, device=dev) * 0.1; b = torch.randn(o, device=dev)
    y = torch.empty(n, pad4(o), device=dev)[:, :o]; dz = torch.randn(n, pad4(o), device=dev)[:, :o]
    dx = torch.empty(n, pad4(k), device=dev)[:, :k]
    dw = torch.zeros(o, k, device=dev); db = torch.zeros(o, device=dev)
    pw = ops.pack_weight(w, False, prec); pwt = ops.pack_weight(w, True, prec)
    fl = 2.0 * n * k * o
    t_f = timeit(lambda: ops.linear_fwd_tc(x, pw, b, o, act, 100.0, prec, out=y))
    t_d = timeit(lambda: ops.linear_bwd_data_tc(dz, pwt, k, x, act, 100.0, prec, out=dx))
    t_w = timeit(lambda: ops.linear_bwd_weight_tc(dz, x, dw, db, prec))
    tot[0] += t_f; tot[1] += t_d; tot[2] += t_w
    print(f"n={n} k={k} o={o}: fwd {t_f:.3f} ms ({fl/t_f/1e9:.0f} TF) dgrad {t_d:.3f} ms ({fl/t_d/1e9:.0f} TF) wgrad {t_w:.3f} ms ({fl/t_w/1e9:.0f} TF)", flush=True)
    del x, y, dz, dx
print(f"sum: fwd {tot[0]:.3f} dgrad {tot[1]:.3f} wgrad {tot[2]:.3f} ms  total {sum(tot):.3f}")
