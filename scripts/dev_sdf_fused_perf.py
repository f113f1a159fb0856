"""dev: time of the fused SDF forward per variant (activation, products, stored or not) at one big batch."""
import os, sys
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
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10489600
k = 71
torch.manual_seed(0)
x = torch.randn(n, 72, device=dev).mul_(0.5)[:, :k]
w0 = torch.randn(256, k, device=dev) * 0.1; b0 = torch.randn(256, device=dev) * 0.1
w1 = torch.randn(256, 256, device=dev) * 0.05; b1 = torch.randn(256, device=dev) * 0.1
w2 = torch.randn(257, 256, device=dev) * 0.05; b2 = torch.randn(257, device=dev) * 0.1
h0 = torch.empty(n, 256, device=dev); h1 = torch.empty(n, 256, device=dev)
for act, name in ((2, "softplus"), (1, "relu")):
    for products in (3, 1):
        t_n = timeit(lambda: ops.sdf_net_fwd_fused(x, w0, b0, w1, b1, w2, b2, act, 100.0, products))
        t_s = timeit(lambda: ops.sdf_net_fwd_fused(x, w0, b0, w1, b1, w2, b2, act, 100.0, products, h0=h0, h1=h1))
        t_1 = timeit(lambda: ops.sdf_net_fwd_fused(x, w0, b0, w1, b1, w2, b2, act, 100.0, products, h1=h1))
        print(f"{name} products={products}: not stored {t_n:.3f} ms | h0+h1 stored {t_s:.3f} ms | h1 stored {t_1:.3f} ms", flush=True)
