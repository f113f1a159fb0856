"""dev: NeRF position encoding at the step's size (10 489 600 rows, 6 frequencies, written into columns 32.. of the
72-float assembled row): forward / backward time and bytes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodalstudio_b200 import ops
dev = "cuda"
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10489600
def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
x = torch.rand(n, 3, device=dev) * 2 - 1
freqs = [2.0 ** k for k in range(6)]
row = torch.empty(n, 72, device=dev)
dout = torch.randn(n, 72, device=dev); dx = torch.empty(n, 3, device=dev)
tf = timeit(lambda: ops.nerf_fwd_into(x, freqs, True, row, 32))
tb = timeit(lambda: ops.nerf_bwd_from(x, freqs, True, dout, 32, dx, False))
print(f"n={n}: fwd {tf:.3f} ms ({n * (12 + 156) / tf / 1e6:.0f} GB/s, {n * 36 / tf / 1e6:.1f} G sin/s)  bwd {tb:.3f} ms ({n * (12 + 156 + 12) / tb / 1e6:.0f} GB/s)")
