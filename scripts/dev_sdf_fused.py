"""dev: the fused SDF-network forward (mmsb_sdf_net_fwd_fused) against fp64 and against the two-kernel path — accuracy on
rows sampled across the batch, stored activations, time per launch.   python scripts/dev_sdf_fused.py [big] [fast]"""
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
def relerr(a, b): return float((a.double() - b).abs().max() / b.abs().max())
what = sys.argv[1:]
torch.manual_seed(0)
beta = 100.0
sizes = [1, 100, 257, 5000, 40000] + ([2097920, 10489600] if "big" in what else [])
for n in sizes:
    k = 71
    x = torch.randn(n, 72, device=dev).mul_(0.5)[:, :k]
    x[:, :32] *= 1e-2                                  # hash features next to positions / PE
    w0 = torch.randn(256, k, device=dev) * 0.1; b0 = torch.randn(256, device=dev) * 0.1
    w1 = torch.randn(256, 256, device=dev) * 0.05; b1 = torch.randn(256, device=dev) * 0.1
    w2 = torch.randn(257, 256, device=dev) * 0.05; b2 = torch.randn(257, device=dev) * 0.1
    idx = torch.arange(n) if n <= 6144 else torch.cat([torch.arange(0, 2048), torch.arange(n // 2, n // 2 + 2048), torch.arange(n - 2048, n)])
    idx = idx.to(dev)
    sp = torch.nn.Softplus(beta=beta)
    h0r = sp(x[idx].double() @ w0.double().T + b0.double())
    h1r = sp(h0r @ w1.double().T + b1.double())
    sdfr = h1r @ w2[0].double() + b2[0].double()
    for products in ([3, 1] if "fast" in what else [3]):
        h0 = torch.full((n, 256), float("nan"), device=dev); h1 = torch.full((n, 256), float("nan"), device=dev)
        sdf = ops.sdf_net_fwd_fused(x, w0, b0, w1, b1, w2, b2, 2, beta, products, h0=h0, h1=h1)
        torch.cuda.synchronize()
        print(f"n={n} products={products}: sdf err {relerr(sdf[idx], sdfr):.2e}  h0 err {relerr(h0[idx], h0r):.2e}  h1 err {relerr(h1[idx], h1r):.2e}"
              f"  nan: {int(torch.isnan(sdf).sum())} {int(torch.isnan(h0).sum())} {int(torch.isnan(h1).sum())}", flush=True)
        sdf2 = ops.sdf_net_fwd_fused(x, w0, b0, w1, b1, w2, b2, 2, beta, products)
        assert torch.equal(sdf, sdf2), "stored / not stored variants differ"
        if n >= 5:
            g = 5
            h1g = torch.full(((n + g - 1) // g, 256), float("nan"), device=dev)
            ops.sdf_net_fwd_fused(x, w0, b0, w1, b1, w2, b2, 2, beta, products, h1=h1g, h1_group=g)
            assert torch.equal(h1g, h1[0::g]), "grouped h1 store differs"
        if n >= 100000:
            t_s = timeit(lambda: ops.sdf_net_fwd_fused(x, w0, b0, w1, b1, w2, b2, 2, beta, products, h0=h0, h1=h1))
            t_n = timeit(lambda: ops.sdf_net_fwd_fused(x, w0, b0, w1, b1, w2, b2, 2, beta, products))
            fl = 2.0 * n * (80 * 256 + 256 * 256)
            print(f"   stored {t_s:.3f} ms ({fl / t_s / 1e9:.0f} TF, {4.0 * n * (72 + 512) / t_s / 1e6:.0f} GB/s) | not stored {t_n:.3f} ms ({fl / t_n / 1e9:.0f} TF)", flush=True)
    if n >= 100000:
        # the two-kernel path (3xTF32)
        pw0, pw1 = ops.pack_weight(w0, False, 3), ops.pack_weight(w1, False, 3)
        h0 = torch.empty(n, 256, device=dev); h1 = torch.empty(n, 256, device=dev); sdf = torch.zeros(n, device=dev)
        def two():
            ops.linear_fwd_tc(x, pw0, b0, 256, 2, beta, 3, out=h0)
            sdf.zero_()
            ops.call("mmsb_linear_fwd_head_tc", ops.ptr(h0), ops._i64(256), ops.ptr(pw1), ops.ptr(b1), ops.ptr(h1), ops._i64(256), ops._i64(n),
                     ops._i32(256), ops._i32(256), ops._i32(2), ops._f32(beta), ops._i32(3), ops.ptr(w2), ops.ptr(b2), ops.ptr(sdf), None, None,
                     ops.stream_ptr())
        t2 = timeit(two)
        print(f"   two kernels (3xTF32) {t2:.3f} ms; sdf err {relerr(sdf[idx], sdfr):.2e}", flush=True)
