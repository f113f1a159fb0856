"""dev: time the three products of a layer at workload sizes (run on the GPU box)."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodalstudio_b200 import ops
from multimodalstudio_b200._lib import call, ptr, stream_ptr
i32, i64, f32 = ctypes.c_int32, ctypes.c_int64, ctypes.c_float
dev = "cuda"
def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
print("path:", os.environ.get("MMSB_MLP_PATH", "tf32x3"))
for n, k, o in [(524288, 71, 256), (524288, 256, 256), (524288, 256, 257), (2097152, 256, 256), (524288, 319, 256), (524288, 256, 64)]:
    x = torch.randn(n, k, device=dev); w = torch.randn(o, k, device=dev) * 0.1; b = torch.randn(o, device=dev)
    y = torch.empty(n, o, device=dev); dz = torch.randn(n, o, device=dev); dx = torch.empty(n, k, device=dev)
    dw = torch.zeros(o, k, device=dev)
    fl = 2.0 * n * k * o
    t_f = timeit(lambda: ops.linear_fwd(x, w, b, 1, 1.0, out=y))
    t_d = timeit(lambda: call("mmsb_linear_bwd_data", ptr(dz), i64(o), ptr(w), ptr(dx), i64(k), ptr(x), i64(k), i32(1), f32(1.0), i64(n), i32(k), i32(o), stream_ptr()))
    t_w = timeit(lambda: call("mmsb_linear_bwd_weight", ptr(dz), i64(o), ptr(x), i64(k), ptr(dw), None, i64(n), i32(k), i32(o), stream_ptr()))
    t_b = timeit(lambda: call("mmsb_linear_bwd_weight", ptr(dz), i64(o), ptr(x), i64(k), None, ptr(b), i64(n), i32(k), i32(o), stream_ptr()))
    t_t = timeit(lambda: torch.matmul(x, w.T))
    print(f"n={n} k={k} o={o}: fwd {t_f:.3f} ms ({fl/t_f/1e9:.1f} TF) dgrad {t_d:.3f} ms ({fl/t_d/1e9:.1f} TF) wgrad {t_w:.3f} ms ({fl/t_w/1e9:.1f} TF) bias {t_b:.3f} ms | torch fp32 matmul {t_t:.3f} ms ({fl/t_t/1e9:.1f} TF)")
# hash grid
res = ops.hash_resolutions(16, 1024, 16)
desc = ops.make_hashgrid_desc(16, 2, 19, res, radius=1.0)
table = torch.rand((2**19) * 16, 2, device=dev) * 1e-3
for n in (458752, 2621440):
    x = torch.rand(n, 3, device=dev) * 1.6 - 0.8
    out = torch.empty(n, 32, device=dev); dout = torch.randn(n, 32, device=dev); dtab = torch.zeros_like(table); dxx = torch.empty(n, 3, device=dev)
    t_f = timeit(lambda: ops.hashgrid_fwd_into(desc, x, table, None, out))
    t_b = timeit(lambda: ops.hashgrid_bwd_from(desc, x, table, None, dout, 0, dtab, None))
    t_bx = timeit(lambda: ops.hashgrid_bwd_from(desc, x, table, None, dout, 0, dtab, dxx))
    print(f"hashgrid n={n}: fwd {t_f:.3f} ms ({n*1024/t_f/1e6:.0f} GB/s alg) bwd(table) {t_b:.3f} ms ({n*2048/t_b/1e6:.0f} GB/s alg) bwd(table+dx) {t_bx:.3f} ms")
