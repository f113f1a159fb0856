import os, sys
sys.path.insert(0, os.getcwd())
import torch
from multimodalstudio_b200 import ops
dev = "cuda"
def timeit(fn, reps=4):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for n, k, o in [(10489600, 71, 256), (419584, 256, 64)]:
    x = torch.randn(n, (k + 3) // 4 * 4, device=dev)[:, :k]; w = torch.randn(o, k, device=dev) * 0.1
    dz = torch.randn(n, o, device=dev); dx = torch.empty(n, (k + 3) // 4 * 4, device=dev)[:, :k]
    pwt = ops.pack_weight(w, True, 3)
    t_d = timeit(lambda: ops.linear_bwd_data_tc(dz, pwt, k, None, 0, 1.0, 3, out=dx))
    ref = (dz[:4096].double() @ w.double())
    err = float((dx[:4096].double() - ref).abs().max() / ref.abs().max())
    print(f"n={n} k={k} o={o}: dgrad {t_d:.3f} ms  err {err:.2e}")
