"""dev: forward layer kernel with parts of the pipeline switched off (MMSB_TC_DEBUG bit mask) to find the bound.
    python scripts/dev_tc_dbg.py [n k o act]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodalstudio_b200 import ops
n, k, o, act = (int(a) for a in sys.argv[1:5]) if len(sys.argv) >= 5 else (2097152, 256, 256, 2)
mode = sys.argv[5] if len(sys.argv) > 5 else "fwd"
x = torch.randn(n, k + (-k) % 4, device="cuda")[:, :k]; w = torch.randn(o, k, device="cuda") * 0.1; b = torch.randn(o, device="cuda")
y = torch.empty(n, o, device="cuda")
dz = torch.randn(n, o, device="cuda"); dx = torch.empty(n, k + (-k) % 4, device="cuda")[:, :k]
print(f"{mode} n={n} k={k} o={o} act={act} pair={os.environ.get('MMSB_TC_PAIR', '1')}")
for prec in (3,):
    pw = ops.pack_weight(w, False, prec); pwt = ops.pack_weight(w, True, prec)
    def run():
        if mode == "fwd":
            ops.linear_fwd_tc(x, pw, b, o, act, 100.0, prec, out=y)
        else:
            ops.linear_bwd_data_tc(dz, pwt, k, x if act else None, act, 100.0, prec, out=dx)
    for dbg, what in [(0, "full"), (2, "no epilogue"), (8, "no MMAs"), (10, "no epilogue, no MMAs"), (32, "no conversion math"), (4, "no B copies")]:
        os.environ["MMSB_TC_DEBUG"] = str(dbg)
        for _ in range(2):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            run()
        e1.record(); torch.cuda.synchronize()
        print(f"prec {prec} dbg {dbg:2d} {what:24s}: {e0.elapsed_time(e1) / 5:.3f} ms", flush=True)
