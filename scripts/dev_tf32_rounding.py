"""dev: how does tcgen05.mma kind::tf32 read an fp32 operand whose low 13 mantissa bits are not zero?
MMSB_TC_DEBUG=32 makes the single-pass kernel store the raw activations (no conversion); compare with a truncation and
a round-to-nearest model in fp64."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodalstudio_b200 import ops
torch.manual_seed(0)
n, k, o = 4096, 256, 256
x = torch.randn(n, k, device="cuda"); w = torch.randn(o, k, device="cuda") * 0.1
pw = ops.pack_weight(w, False, 1)
def rna(t): return ((t.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)
def trunc(t): return (t.view(torch.int32) & ~0x1FFF).view(torch.float32)
os.environ["MMSB_TC_DEBUG"] = "32"
y = ops.linear_fwd_tc(x, pw, None, o, 0, 1.0, 1)
torch.cuda.synchronize()
wq = rna(w).double()
for name, f in (("truncate", trunc), ("round-nearest-away", rna)):
    ref = f(x).double() @ wq.T
    print(name, "max abs diff", float((y.double() - ref).abs().max()), "of scale", float(ref.abs().max()))
