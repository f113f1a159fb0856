"""dev: tcgen05 GEMM products vs torch fp64 matmul (run on the GPU box)."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodalstudio_b200 import ops
from multimodalstudio_b200._lib import call, ptr, stream_ptr
i32, i64, f32 = ctypes.c_int32, ctypes.c_int64, ctypes.c_float
torch.manual_seed(0)
dev = "cuda"
def rel(a, b): return float((a.double() - b.double()).abs().max() / b.double().abs().max())
for n, k, o in [(4099, 71, 256), (4099, 256, 256), (4099, 256, 257), (4099, 319, 256), (1000, 283, 128), (130000, 256, 256), (4099, 39, 256)]:
    x = torch.randn(n, k, device=dev); w = torch.randn(o, k, device=dev) * 0.1; b = torch.randn(o, device=dev)
    y = ops.linear_fwd(x, w, b, 0, 1.0)
    ref = x.double() @ w.double().T + b.double()
    e_f = rel(y, ref)
    dz = torch.randn(n, o, device=dev)
    dx = torch.empty(n, k, device=dev)
    call("mmsb_linear_bwd_data", ptr(dz), i64(o), ptr(w), ptr(dx), i64(k), None, i64(k), i32(0), f32(1.0), i64(n), i32(k), i32(o), stream_ptr())
    e_d = rel(dx, dz.double() @ w.double())
    dw = torch.zeros(o, k, device=dev); db = torch.zeros(o, device=dev)
    call("mmsb_linear_bwd_weight", ptr(dz), i64(o), ptr(x), i64(k), ptr(dw), ptr(db), i64(n), i32(k), i32(o), stream_ptr())
    e_w = rel(dw, dz.double().T @ x.double()); e_b = rel(db, dz.double().sum(0))
    torch.cuda.synchronize()
    # where is the dgrad error?
    err = (dx.double() - dz.double() @ w.double()).abs()
    cols = err.max(0).values
    bad = (cols > 1e-3 * float(cols.max() + 1e-30) + 1e-4).nonzero().flatten().tolist()
    print(f"n={n} k={k} o={o}: fwd {e_f:.2e} dgrad {e_d:.2e} wgrad {e_w:.2e} bias {e_b:.2e}  bad dgrad cols: {bad[:8]}{'...' if len(bad)>8 else ''} ({len(bad)})")
