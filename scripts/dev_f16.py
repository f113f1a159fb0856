"""dev: precision 2 (2-term fp16 split) against fp64 and against 3xTF32 — accuracy on rows sampled across the matrix,
time per launch, at the step's big shapes.   python scripts/dev_f16.py [fwd|dgrad|wgrad ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodalstudio_b200 import ops
dev = "cuda"
def pad4(k): return (k + 3) // 4 * 4
def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
def relerr(a, b): return float((a.double() - b).abs().max() / b.abs().max())
what = sys.argv[1:] or ["fwd"]
torch.manual_seed(0)
SHAPES = [(4194304, 256, 256, 1.0), (4194304, 71, 256, 1.0)] if "big" in what else None
for n, k, o, xs in SHAPES or [(20000, 256, 256, 1.0), (20000, 256, 256, 1e-6), (20000, 256, 256, 3e3), (30000, 71, 256, 1.0), (30000, 319, 256, 1.0),
                    (2097920, 256, 256, 1.0), (10489600, 256, 256, 1.0), (10489600, 71, 256, 1.0)]:
    x = torch.randn(n, pad4(k), device=dev).mul_(xs)[:, :k]
    x[:, :8] *= 1e-3                                   # columns of very different magnitude (hash features next to PE)
    w = torch.randn(o, k, device=dev) * 0.1; b = torch.randn(o, device=dev) * xs
    rows = torch.cat([torch.arange(0, 2048), torch.arange(n // 2, n // 2 + 2048), torch.arange(n - 2048, n)]).to(dev)
    if "fwd" in what:
        ref = torch.relu(x[rows].double() @ w.double().T + b.double())
        res = {}
        for prec in (3, 2):
            pw = ops.pack_weight(w, False, prec)
            y = torch.empty(n, o, device=dev)
            am = ops.amax_of(x) if prec == 2 else None
            yam = torch.zeros(1, device=dev)
            f = lambda: ops.linear_fwd_tc(x, pw, b, o, 1, 1.0, prec, out=y, x_amax=am, y_amax=yam)
            t = timeit(f)
            res[prec] = (relerr(y[rows], ref), t)
            if prec == 2:
                assert abs(float(yam) - float(y.abs().max())) <= 1e-6 * float(yam), (float(yam), float(y.abs().max()))
        fl = 2.0 * n * k * o
        print(f"fwd n={n} k={k} o={o} scale={xs:g}: 3xTF32 err {res[3][0]:.2e} {res[3][1]:.3f} ms ({fl / res[3][1] / 1e9:.0f} TF) | "
              f"fp16x2 err {res[2][0]:.2e} {res[2][1]:.3f} ms ({fl / res[2][1] / 1e9:.0f} TF, {4.0 * n * (k + o) / res[2][1] / 1e6:.0f} GB/s) "
              f"speed-up {res[3][1] / res[2][1]:.2f}x", flush=True)
