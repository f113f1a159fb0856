"""dev: time the three products of a layer at workload sizes (run on the GPU box)."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodalstudio_b200 import ops
dev = "cuda"
def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for prec, act in ((3, 1), (3, 2), (1, 2)):
    print("precision:", prec, "activation:", {1: "ReLU", 2: "Softplus(100)"}[act])
    for n, k, o in [(524288, 71, 256), (524288, 256, 256), (2097152, 256, 256), (524288, 319, 256), (524288, 256, 64),
                    (524288, 256, 1), (524288, 256, 257)]:
        x = torch.randn(n, k, device=dev); w = torch.randn(o, k, device=dev) * 0.1; b = torch.randn(o, device=dev)
        y = torch.empty(n, o, device=dev); dz = torch.randn(n, o, device=dev); dx = torch.empty(n, k, device=dev)
        dw = torch.zeros(o, k, device=dev); db = torch.zeros(o, device=dev)
        pw = ops.pack_weight(w, False, prec); pwt = ops.pack_weight(w, True, prec)
        fl = 2.0 * n * k * o
        t_f = timeit(lambda: ops.linear_fwd_tc(x, pw, b, o, act, 100.0, prec, out=y))
        t_d = timeit(lambda: ops.linear_bwd_data_tc(dz, pwt, k, x, act, 100.0, prec, out=dx))
        t_w = timeit(lambda: ops.linear_bwd_weight_tc(dz, x, dw, db, prec))
        t_p = timeit(lambda: ops.pack_weight(w, False, prec))
        t_t = timeit(lambda: torch.matmul(x, w.T))
        print(f"n={n} k={k} o={o}: fwd {t_f:.3f} ms ({fl/t_f/1e9:.1f} TF) dgrad {t_d:.3f} ms ({fl/t_d/1e9:.1f} TF) wgrad {t_w:.3f} ms ({fl/t_w/1e9:.1f} TF) pack {t_p:.3f} ms | torch fp32 matmul {t_t:.3f} ms ({fl/t_t/1e9:.1f} TF)", flush=True)
