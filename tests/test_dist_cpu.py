"""CPU, gloo, world size 2: the multi-GPU scheme of DESIGN.md §6 ("strong" mode) — ONE global ray batch
split contiguously over the ranks by the product's `pipelines.ShardPlan` (ragged modality counts, micro-batches on
each rank), every loss normalised by the GLOBAL count, gradients accumulated over micro-batches in a flat buffer and
combined by the product's `pipelines.all_reduce_optimizers` over `FlatAdamW`'s flat buffer (one summing all-reduce; the
same objects and calls `RawPipeline.all_reduce_gradients` makes on NCCL) — reproduces the gradient of the unsharded batch.  Also the "weak" mode (per-rank mean losses, summing all-reduce, 1 / world size folded into the
optimizer's gradient read = FlatAdamW.prescale).  The arithmetic of a shard is the CPU oracle here (there is no GPU in
this container; the same plan drives the CUDA path in tests/test_gpu_model.py::test_sharded_*)."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COUNTS = {"mono": 7, "rgb": 5}            # ragged on purpose: 7 does not split evenly over 2 ranks x 2 micro-batches


def _setup():
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import mms_oracle as O
    from multimodalstudio_b200.models import build_model
    mods = {"mono": 1, "rgb": 3}
    model = build_model("grid_raw", modalities=mods, log2_hashmap_size=8, num_samples=8, num_samples_importance=8, bg_samples=4, seed=3)
    params = [p for p in model.parameters()]
    sd = {k: v for k, v in model.state_dict(keep_vars=True).items()}
    orc = O.GridModelOracle(sd, O.default_cfg(modalities=mods, log2_hashmap_size=8, num_samples=8, num_samples_importance=8, bg_samples=4))
    orc.training = False
    orc.set_schedule_state(16, 2.0 / 16, 1.0)     # wide taps: keeps the finite-difference amplification of fp32 noise small
    g = torch.Generator().manual_seed(5)
    rays, tgt = {}, {}
    for m, n in COUNTS.items():
        o = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1) * 2.5
        d = torch.nn.functional.normalize(-o + 0.2 * torch.randn(n, 3, generator=g), dim=-1)
        up = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1)
        rays[m] = torch.cat([o, d, up], -1)
        tgt[m] = torch.rand(n, mods[m], generator=g)
    return O, orc, params, rays, tgt


def _flat_grad_of_plan(O, orc, opt, rays, tgt, plan, geometry_count):
    """Runs every micro-batch of `plan` through the oracle with the product's loss normalisation; the gradients are
    accumulated by the PRODUCT's optimizer object (pipelines.FlatAdamW: detach_grads -> backward -> gather_grads into
    its flat buffer, exactly what RawPipeline.forward_backward does around the CUDA path)."""
    total_sum = 0.0
    for j in range(len(plan)):
        part, tpart, scales = plan.slice(j, rays), plan.slice(j, tgt), plan.loss_scales(j)
        total, grads = 0.0, []
        for m in COUNTS:
            if part[m].shape[0] == 0:
                continue
            out = orc.forward_modality(m, part[m][:, 0:3], part[m][:, 3:6], part[m][:, 6:9], None)
            total = total + scales[m] * (out[m] - tpart[m]).abs().mean()            # mean x n_micro / n_global
            grads.append(out["gradients"])
        g = torch.cat(grads, 0)
        eik_sum = ((g.norm(dim=-1) - 1.0) ** 2).sum()
        total = total + 0.1 * eik_sum / geometry_count                              # local sum / GLOBAL count
        opt.detach_grads()
        total.backward()
        opt.gather_grads(accumulate=j > 0)
        total_sum += float(total.detach())
    return total_sum, opt.grad


def _global_count(O, rays):
    return float(sum(int(O.sphere_collide(r[:, 0:3], r[:, 3:6])[2].sum()) for r in rays.values()) * 16)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from multimodalstudio_b200.pipelines import FlatAdamW, ShardPlan, all_reduce_flat, all_reduce_optimizers
    O, orc, params, rays, tgt = _setup()
    opt = FlatAdamW(params, max_norm=None)
    plan = ShardPlan(COUNTS, world, rank, max_rays_per_micro=4)
    assert len(plan) == 2
    # the global in-sphere sample count: local count, summed over the ranks (as RawPipeline.train_step_sharded does)
    local = plan.local_slice(rays)
    cnt = torch.tensor([_global_count(O, local)])
    all_reduce_flat(cnt)
    total, flat = _flat_grad_of_plan(O, orc, opt, rays, tgt, plan, float(cnt))
    all_reduce_optimizers({"fields": opt}, mean=False)         # the ONE data-path collective: a summing all-reduce
    assert opt.prescale == 1.0
    t = torch.tensor([total])
    all_reduce_flat(t)
    flat = flat.clone()
    # weak mode on the same buffers: the all-reduce still SUMS, DDP's mean is the optimizer's prescale
    before = opt.grad.clone()
    all_reduce_optimizers({"fields": opt}, mean=True)
    assert opt.prescale == 1.0 / world and torch.allclose(opt.grad, before * world)
    if rank == 0:
        q.put((float(t), flat, float(cnt)))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_global_batch_plus_allreduce_equals_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    loss2, flat2, cnt2 = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    threads = torch.get_num_threads()
    torch.set_num_threads(2)
    sys.path.insert(0, ROOT)
    from multimodalstudio_b200.pipelines import FlatAdamW, ShardPlan
    O, orc, params, rays, tgt = _setup()
    cnt1 = _global_count(O, rays)
    assert cnt1 == cnt2
    loss1, flat1 = _flat_grad_of_plan(O, orc, FlatAdamW(params, max_norm=None), rays, tgt, ShardPlan(COUNTS), cnt1)      # one rank, one batch
    assert abs(loss1 - loss2) < 1e-6 * max(1.0, abs(loss1)), (loss1, loss2)
    err = float((flat1 - flat2).abs().max() / flat1.abs().max())
    torch.set_num_threads(threads)
    assert err < 1e-4, err       # equal up to fp32 reassociation (different batch shapes pick different BLAS kernels)


def test_shard_plan_covers_every_ray_once():
    sys.path.insert(0, ROOT)
    from multimodalstudio_b200.pipelines import ShardPlan
    counts = {"rgb": 13108, "infrared": 13107, "mono": 13107, "polarization": 13107, "multispectral": 13107}
    for world in (1, 2, 3, 4, 8):
        seen = {m: [] for m in counts}
        scale_sum = {m: 0.0 for m in counts}
        for rank in range(world):
            plan = ShardPlan(counts, world, rank, max_rays_per_micro=8192)
            assert sum(b - a for a, b in plan.local.values()) <= -(-65536 // world) + len(counts)
            for j in range(len(plan)):
                assert sum(b - a for a, b in plan.micro[j].values()) <= 8192
                for m, (a, b) in plan.micro[j].items():
                    seen[m] += list(range(a, b))
                    scale_sum[m] += plan.loss_scales(j)[m]
        for m, n in counts.items():
            assert seen[m] == list(range(n))
            assert abs(scale_sum[m] - 1.0) < 1e-12
    # the bench's even case: every rank and micro-batch has the same shape (one CUDA-graph key)
    even = {m: 13112 for m in counts}
    shapes = {tuple(sorted((m, b - a) for m, (a, b) in mb.items())) for world in (1, 2, 4, 8) for r in range(world)
              for mb in ShardPlan(even, world, r, max_rays_per_micro=8195).micro}
    assert len(shapes) == 1
