"""CPU, gloo, world size 2: the multi-GPU scheme of DESIGN.md §6 — rays sharded across ranks, parameters
replicated, ONE flat gradient all-reduce (mean) — reproduces the single-process gradient.  The arithmetic on each
rank is the CPU oracle here (no GPU in this container); the sharding / all-reduce / flat-buffer logic is the
product's (`pipelines.FlatAdamW`-style flat gradient, `all_reduce_gradients` semantics)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _loss_and_flat_grad(rank, world):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import mms_oracle as O
    from multimodalstudio_b200.models import build_model
    mods = {"mono": 1}
    model = build_model("grid_raw", modalities=mods, log2_hashmap_size=8, num_samples=8, num_samples_importance=8, bg_samples=4, seed=3)
    params = [p for p in model.parameters()]
    sd = {k: v for k, v in model.state_dict(keep_vars=True).items()}
    orc = O.GridModelOracle(sd, O.default_cfg(modalities=mods, log2_hashmap_size=8, num_samples=8, num_samples_importance=8, bg_samples=4))
    orc.cfg["num_upsample_steps"] = 4
    orc.training = False
    orc.set_schedule_state(16, 2.0 / 16, 1.0)     # wide taps: keeps the finite-difference amplification of fp32 noise small
    g = torch.Generator().manual_seed(5)
    n = 8
    o = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1) * 2.5
    d = torch.nn.functional.normalize(-o + 0.2 * torch.randn(n, 3, generator=g), dim=-1)
    up = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1)
    tgt = torch.rand(n, 1, generator=g)
    lo, hi = rank * n // world, (rank + 1) * n // world          # contiguous shard of the ray batch
    out = orc.forward_modality("mono", o[lo:hi], d[lo:hi], up[lo:hi], None)
    loss = (out["mono"] - tgt[lo:hi]).abs().mean()               # per-rank mean; ranks are averaged by the all-reduce
    loss.backward()
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in params])
    return loss.detach(), flat


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    loss, flat = _loss_and_flat_grad(rank, world)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)                  # one flat all-reduce, then the mean (DDP semantics)
    flat /= world
    dist.all_reduce(loss, op=dist.ReduceOp.SUM)
    if rank == 0:
        q.put((loss / world, flat))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_rays_plus_allreduce_equal_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    import socket
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    loss2, flat2 = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    torch.set_num_threads(2)
    loss1, flat1 = _loss_and_flat_grad(0, 1)
    assert abs(float(loss1) - float(loss2)) < 1e-6
    err = float((flat1 - flat2).abs().max() / flat1.abs().max())
    assert err < 1e-4, err       # equal up to fp32 reassociation (different batch shapes pick different BLAS kernels)
