"""dev: sensitivity of the forward layer kernel to the ring depth (MMSB_TC_MAX_STAGES) at single-pass TF32 (4 stages fit)
and for a narrow 3xTF32 layer (N = 64: 4 stages fit)."""
import os, subprocess, sys
code = r'''
import os, sys, torch
sys.path.insert(0, os.getcwd())
from multimodalstudio_b200 import ops
n = 2097152
for prec, k, o in ((1, 256, 256), (3, 256, 64), (3, 256, 128)):
    x = torch.randn(n, k, device="cuda"); w = torch.randn(o, k, device="cuda") * 0.1; b = torch.randn(o, device="cuda")
    y = torch.empty(n, o, device="cuda"); pw = ops.pack_weight(w, False, prec)
    for _ in range(2): ops.linear_fwd_tc(x, pw, b, o, 1, 1.0, prec, out=y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): ops.linear_fwd_tc(x, pw, b, o, 1, 1.0, prec, out=y)
    e1.record(); torch.cuda.synchronize()
    print(f"stages<={os.environ.get('MMSB_TC_MAX_STAGES','-')} prec {prec} {k}->{o}: {e0.elapsed_time(e1)/5:.3f} ms", flush=True)
'''
for st in ("2", "3", "4"):
    env = dict(os.environ, MMSB_TC_MAX_STAGES=st)
    subprocess.run([sys.executable, "-c", code], env=env, check=False)
