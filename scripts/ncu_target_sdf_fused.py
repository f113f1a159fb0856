"""ncu target: the fused SDF forward at the north-star micro-batch (10 489 600 rows): launch 0 = not stored, 1 = h0 + h1 stored."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodalstudio_b200 import ops
dev = "cuda"
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10489600
products = int(sys.argv[2]) if len(sys.argv) > 2 else 3
torch.manual_seed(0)
x = torch.randn(n, 72, device=dev).mul_(0.5)[:, :71]
w0 = torch.randn(256, 71, device=dev) * 0.1; b0 = torch.randn(256, device=dev) * 0.1
w1 = torch.randn(256, 256, device=dev) * 0.05; b1 = torch.randn(256, device=dev) * 0.1
w2 = torch.randn(257, 256, device=dev) * 0.05; b2 = torch.randn(257, device=dev) * 0.1
h0 = torch.empty(n, 256, device=dev); h1 = torch.empty(n, 256, device=dev)
ops.sdf_net_fwd_fused(x, w0, b0, w1, b1, w2, b2, 2, 100.0, products)
ops.sdf_net_fwd_fused(x, w0, b0, w1, b1, w2, b2, 2, 100.0, products, h0=h0, h1=h1)
torch.cuda.synchronize()
