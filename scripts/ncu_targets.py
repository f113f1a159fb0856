"""Target of the `ncu --set full` captures (profiles/): one warm-up + one measured launch of each dominant kernel at the
sizes of one micro-batch of the default (north-star) workload: 8195 rays x 256 samples x (centre + 4 taps) = 10 489 600
SDF rows.  The three tcgen05 layer products (256 -> 256, Softplus(100), 3xTF32; the forward also as the opt-in fp16
split), the 71 -> 256 first layer, the fused SDF forward (mmsb_sdf_net_fwd_fused, with and without
the activation stores), and the hash-grid forward / backward on ray-coherent points in the grouped layout of
the step (a sample's centre + 4 tap evaluations in adjacent rows, 16 levels x 2^19 x 2 fp32).

    ncu --set full --clock-control none --import-source on -k regex:'tc_|hashgrid|sdf_fused' --launch-skip 9 -c 9 -o prof python scripts/ncu_targets.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodalstudio_b200 import ops
from multimodalstudio_b200.field_components import HashEncodingConfig
torch.manual_seed(0)
dev = "cuda"
n, k, o = 8195 * 256 * 5, 256, 256
x = torch.randn(n, k, device=dev); w = torch.randn(o, k, device=dev) * 0.1; b = torch.randn(o, device=dev)
y = torch.empty(n, o, device=dev); dz = torch.randn(n, o, device=dev); dx = torch.empty(n, k, device=dev)
dw = torch.zeros(o, k, device=dev); db = torch.zeros(o, device=dev)
pw = ops.pack_weight(w, False, 3); pwt = ops.pack_weight(w, True, 3); pw2 = ops.pack_weight(w, False, 2)
x71 = torch.randn(n, 72, device=dev)[:, :71]; w71 = torch.randn(o, 71, device=dev) * 0.1; pw71 = ops.pack_weight(w71, False, 3)
xam = ops.amax_of(x)
w2 = torch.randn(257, 256, device=dev) * 0.05; b2 = torch.randn(257, device=dev) * 0.1; sdf = torch.empty(n, device=dev)
enc = HashEncodingConfig(num_levels=16, min_res=16, max_res=1024, log2_hashmap_size=19, features_per_level=2, interpolation="Linear").setup(in_dim=3).to(dev)
# ray-coherent points: 8195 rays through the unit sphere, 256 samples each, every sample followed by its 4 taps
rays_o = torch.nn.functional.normalize(torch.randn(8195, 3, device=dev), dim=-1) * 2.5
rays_d = torch.nn.functional.normalize(-rays_o + 0.3 * torch.randn(8195, 3, device=dev), dim=-1)
# samples between the ray's entry into and exit from the unit sphere (rays that miss it: a short chord at closest approach)
b_ = (rays_o * rays_d).sum(-1, keepdim=True)
disc = (b_ ** 2 - (rays_o.norm(dim=-1, keepdim=True) ** 2 - 1.0)).clamp_min(0.01).sqrt()
t = (-b_ - disc) + (2 * disc) * torch.linspace(0.0, 1.0, 256, device=dev)[None]
centre = rays_o[:, None] + rays_d[:, None] * t[:, :, None]
offs = torch.tensor([[0, 0, 0], [1, -1, -1], [-1, -1, 1], [-1, 1, -1], [1, 1, 1]], device=dev, dtype=torch.float32) * (2.0 / 1024 / 3 ** 0.5)
pts = (centre[:, :, None, :] + offs).reshape(-1, 3).contiguous()
mask = torch.ones(32, device=dev)
feat = torch.empty(pts.shape[0], 32, device=dev)
dfeat = torch.randn(pts.shape[0], 32, device=dev)
dtab = torch.zeros_like(enc.hash_table)
dpts = torch.empty_like(pts)
desc = enc.desc(1.0)
tab = enc.hash_table.detach()
for _ in range(2):
    ops.linear_fwd_tc(x, pw, b, o, 2, 100.0, 3, out=y)
    ops.linear_bwd_data_tc(dz, pwt, k, x, 2, 100.0, 3, out=dx)
    ops.linear_bwd_weight_tc(dz, x, dw, db, 3)
    ops.linear_fwd_tc(x71, pw71, b, o, 2, 100.0, 3, out=y)
    ops.linear_fwd_tc(x, pw2, b, o, 2, 100.0, 2, out=y, x_amax=xam)
    ops.hashgrid_fwd_into(desc, pts, tab, mask, feat)
    ops.hashgrid_bwd_from(desc, pts, tab, mask, dfeat, 0, dtab, dpts)
    # the fused SDF forward (71 -> 256 -> 256 -> sdf, fp16 split): activations not stored (sampler / inference), then stored
    ops.sdf_net_fwd_fused(x71, w71, b, w, b, w2, b2, 2, 100.0, 3, sdf=sdf)
    ops.sdf_net_fwd_fused(x71, w71, b, w, b, w2, b2, 2, 100.0, 3, h0=y, h1=dx, sdf=sdf)
    torch.cuda.synchronize()
print("ok")
