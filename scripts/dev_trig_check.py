import os, sys
sys.path.insert(0, os.getcwd())
import torch
from multimodalstudio_b200 import ops
dev = "cuda"
torch.manual_seed(0)
n = 2_000_000
for scale in (1.0, 3.0, 50.0):
    x = (torch.rand(n, 3, device=dev) * 2 - 1) * scale
    freqs = [2.0 ** k for k in range(10)]
    out = torch.empty(n, 63, device=dev)
    ops.nerf_fwd_into(x, freqs, True, out, 0)
    xs = x.double()[:, :, None] * torch.tensor(freqs, device=dev, dtype=torch.float64)     # fl(x f) exact in double? x*2^k exact
    ref = torch.cat([x.double(), torch.sin(xs).reshape(n, -1), torch.sin((xs.float() + 1.5707963267948966).double()).reshape(n, -1)], -1)
    # the kernel computes sin(fl32(fl32(x f) + pi/2)): build the same argument in fp32
    arg2 = (x[:, :, None] * torch.tensor(freqs, device=dev)) + torch.tensor(1.57079632679489661923, device=dev)
    ref2 = torch.sin(arg2.double()).reshape(n, -1)
    err1 = (out[:, 3:33].double() - torch.sin(xs).reshape(n, -1)).abs().max().item()
    err2 = (out[:, 33:].double() - ref2).abs().max().item()
    print(f"scale {scale}: max abs err sin {err1:.3e}  shifted sin {err2:.3e}")
