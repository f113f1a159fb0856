"""Mirror of the reference's pipeline step for the hot path: ray generation -> BaseModel.forward ->
mosaick channel select + losses -> backward -> global-norm clip -> AdamW.
ref: src/pipelines/raw_pipeline.py:67-82,112-122, src/pipelines/base_pipeline.py:139-153,232-248,
     src/engine/optimizers.py:96-116, src/engine/schedulers.py:249-270, src/engine/trainer.py:86-138

The synthetic scene generator follows SURVEY.md §8(d) (no dataset is available offline).
"""
import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np
import torch

from . import ops
from .cameras import CameraOptimizer, CameraOptimizerConfig, Cameras, RayGenerator
from .configs import TrainingCallbackAttributes, TrainingCallbackLocation
from .models import MODALITY_CHANNELS, MOSAICK_PATTERNS, build_model, grid_loss_config, loss_config_for

# per-modality image geometry: (width, height, focal)
MODALITY_SENSORS = {
    "rgb": (2448, 2048, 2400.0), "infrared": (1224, 1024, 1200.0), "mono": (1224, 1024, 1200.0),
    "polarization": (2448, 2048, 2400.0), "multispectral": (1269, 981, 1250.0),
}


def select_right_channel_per_pixel(pixel_coords_per_modality, outputs, modalities=None, patterns=None):
    """ref: raw_pipeline.py:112-122 — `outputs[mod][mod] [R,C] -> [R,1]` with band = pattern[y % ph, x % pw].
    (The training step does not call this: the select is fused into the loss kernel.)"""
    patterns = patterns or MOSAICK_PATTERNS
    for mod in (modalities or outputs.keys()):
        pat = torch.tensor(patterns[mod], dtype=torch.int32, device=outputs[mod][mod].device)
        _, sel = ops.mosaick_bands(pixel_coords_per_modality[mod], pat.reshape(-1), pat.shape[0], pat.shape[1],
                                   outputs[mod][mod])
        outputs[mod][mod] = sel[:, None]
    return outputs


def look_at_cameras(n_cam: int, radius: float, seed: int, offset=None) -> torch.Tensor:
    """n_cam camera-to-world matrices [n,3,4] on a shell of `radius`, looking at the origin (+ jitter),
    OpenGL convention (camera looks down -z, +y up) like the reference's ray generation."""
    g = torch.Generator().manual_seed(seed)
    pos = torch.nn.functional.normalize(torch.randn(n_cam, 3, generator=g), dim=-1) * radius
    fwd = torch.nn.functional.normalize(-pos + 0.05 * torch.randn(n_cam, 3, generator=g), dim=-1)
    helper = torch.tensor([0.0, 0.0, 1.0]).expand(n_cam, 3)
    right = torch.nn.functional.normalize(torch.linalg.cross(fwd, helper), dim=-1)
    up = torch.linalg.cross(right, fwd)
    c2w = torch.cat([torch.stack([right, up, -fwd], -1), pos[..., None]], -1)
    if offset is not None:   # camera2reference folded into camtoworld offline (preprocessing/utils.py:531-548)
        rot, t = offset
        c2w = torch.cat([c2w[:, :, :3] @ rot, c2w[:, :, 3:] + c2w[:, :, :3] @ t[:, None]], -1)
    return c2w


class SyntheticScene:
    """Cameras, pixel coordinates and targets of the synthetic workload."""

    def __init__(self, modalities: Dict[str, int], rays_per_modality: Dict[str, int], n_cam: int = 50, seed: int = 654824,
                 raw: bool = True):
        self.modalities, self.rays, self.raw = modalities, rays_per_modality, raw
        self.cameras = {}
        for i, mod in enumerate(modalities):
            w, h, f = MODALITY_SENSORS[mod]
            ang = 0.01 * (i + 1)
            rot = torch.tensor([[math.cos(ang), -math.sin(ang), 0.0], [math.sin(ang), math.cos(ang), 0.0], [0.0, 0.0, 1.0]])
            c2w = look_at_cameras(n_cam, 2.5, seed + 17, offset=(rot, torch.tensor([0.02 * i, -0.01 * i, 0.0])))
            dist = torch.tensor([-0.1, 0.01, 0.0, 0.0, 1e-3, -1e-3]).expand(n_cam, 6).contiguous()
            self.cameras[mod] = Cameras(c2w, f, f, w / 2.0, h / 2.0, width=w, height=h, distortion_params=dist)
        self.n_cam = n_cam
        self.gen = torch.Generator().manual_seed(seed)     # trainer.py:64 / pixel_samplers.py:52 (+ rank offset by the caller)

    def sample_batch(self):
        """UniformPixelSampler semantics (pixel_samplers.py:71-89): CPU randint per modality -> coords int32 [R,3]
        (cam, y, x) and synthetic targets."""
        coords, targets = {}, {}
        for mod, c in self.modalities.items():
            w, h, _ = MODALITY_SENSORS[mod]
            r = self.rays[mod]
            coords[mod] = torch.stack([torch.randint(0, self.n_cam, (r,), generator=self.gen),
                                       torch.randint(0, h, (r,), generator=self.gen),
                                       torch.randint(0, w, (r,), generator=self.gen)], -1).int()
            t = torch.rand(r, 1 if self.raw else c, generator=self.gen)
            if mod == "polarization":
                t[torch.rand(r, generator=self.gen) < 0.005] = 1.0       # exercise skip-saturation
            targets[mod] = t
        return coords, targets


class DevicePixelSampler:
    """UniformPixelSampler + the loader's target gather on the device (pixel_samplers.py:71-89,
    dataloaders.py:164-167): `frames` {mod: fp32 [n_cam, H, W, C]} stay resident in HBM (raw mosaicked stacks have
    C = 1), `sample(step)` returns coords {mod: int32 [R,3]} and targets {mod: [R,C]} without touching the host.  Uniform
    like the reference, not the same sequence (counter-based Philox instead of torch's CPU generator); the rank offsets
    the seed like pixel_samplers.py:49-52."""

    def __init__(self, frames: Dict[str, torch.Tensor], rays_per_modality: Dict[str, int], seed: int = 654824, rank: int = 0):
        self.frames, self.rays, self.seed = frames, rays_per_modality, seed + rank

    def sample(self, step: int):
        coords, targets = {}, {}
        for i, (mod, fr) in enumerate(self.frames.items()):
            n_cam, h, w, _ = fr.shape
            coords[mod], targets[mod] = ops.sample_pixels(self.seed, step, i, n_cam, h, w, self.rays[mod], fr.device, fr)
        return coords, targets


class ShardPlan:
    """How the global ray batch of ONE optimizer step is split ("strong" mode): rank r of G takes the
    contiguous slice [r*R_m/G, (r+1)*R_m/G) of every modality m, so the modality mix of every rank is the global one;
    a rank's slice is cut the same way into k micro-batches of at most `max_rays_per_micro` rays (all modalities
    together) that run one after the other and accumulate their gradients.  Every loss is normalised by the GLOBAL count
    (`loss_scales`: a micro-batch's mean times n_micro / n_global), so the sum of all shards' gradients — over
    micro-batches in the flat buffer, over ranks by the summing all-reduce — IS the single-batch gradient.
    Pure Python (no device work): tests/test_dist_cpu.py drives it under gloo."""

    def __init__(self, global_counts: Dict[str, int], world_size: int = 1, rank: int = 0,
                 max_rays_per_micro: Optional[int] = None):
        if not (0 <= rank < world_size):
            raise ValueError(f"rank {rank} outside world size {world_size}")
        self.global_counts, self.world_size, self.rank = dict(global_counts), world_size, rank
        self.local = {m: (rank * n // world_size, (rank + 1) * n // world_size) for m, n in global_counts.items()}
        local_total = sum(b - a for a, b in self.local.values())
        k = 1
        if max_rays_per_micro is not None and local_total > max_rays_per_micro:
            k = -(-local_total // max_rays_per_micro)
            # per-modality rounding: a micro-batch holds at most sum_m ceil(L_m / k) rays
            while sum(-(-(b - a) // k) for a, b in self.local.values()) > max_rays_per_micro:
                k += 1
        self.micro = []
        for j in range(k):
            self.micro.append({m: (a + j * (b - a) // k, a + (j + 1) * (b - a) // k) for m, (a, b) in self.local.items()})

    def __len__(self):
        return len(self.micro)

    def loss_scales(self, j: int) -> Dict[str, float]:
        """n_micro / n_global per modality: what a micro-batch's mean loss is multiplied with."""
        return {m: (b - a) / max(self.global_counts[m], 1) for m, (a, b) in self.micro[j].items()}

    def slice(self, j: int, tensors: Dict[str, torch.Tensor], base: str = "global") -> Dict[str, torch.Tensor]:
        """Rows of micro-batch j out of per-modality tensors indexed globally (`base="global"`) or relative to this
        rank's local slice (`base="local"`)."""
        out = {}
        for m, (a, b) in self.micro[j].items():
            off = self.local[m][0] if base == "local" else 0
            out[m] = tensors[m][a - off:b - off]
        return out

    def local_slice(self, tensors: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        return {m: tensors[m][a:b] for m, (a, b) in self.local.items()}


def all_reduce_flat(flat: torch.Tensor, group=None):
    """The ONE data-path collective of the hot path: summing all-reduce of a flat fp32 gradient buffer
    over NCCL (gloo in the CPU tests).  The division by the world size of DDP's mean is not applied here: the optimizer
    kernel folds it into its gradient read (`FlatAdamW.prescale`)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return flat
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat


def all_reduce_optimizers(optimizers, group=None, mean: bool = True):
    """ONE summing all-reduce per optimizer over its flat gradient buffer (`FlatAdamW.grad`).  `mean=True` (weak scaling,
    every rank a full batch with its own mean losses — the reference's DDP semantics): the 1 / world size of DDP's
    average is folded into the optimizer kernel's gradient read (FlatAdamW.prescale), no extra pass over the 138 MB.
    `mean=False` (strong scaling, losses already normalised by the global counts): the sum IS the single-batch gradient."""
    import torch.distributed as dist
    ws = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    for opt in optimizers.values():
        opt.prescale = 1.0 / ws if (mean and ws > 1) else 1.0
        if ws > 1:
            all_reduce_flat(opt.grad, group)


class FlatAdamW:
    """One AdamW "optimizer" of the reference (engine/optimizers.py, method_configs.py:260-269) over a flat
    fp32 buffer: parameters and gradients are views into two contiguous tensors, so zero-grad is one memset,
    the global-norm clip one reduction, the update one fused kernel and the DDP all-reduce one NCCL call."""

    def __init__(self, params: List[torch.nn.Parameter], lr=1e-3, weight_decay=0.01, eps=1e-15, betas=(0.9, 0.999),
                 max_norm: Optional[float] = 2.0):
        self.params = [p for p in params if p.requires_grad]
        align = 32      # floats: every parameter view starts on a 128-byte boundary (vector loads / atomics)
        n = sum((p.numel() + align - 1) // align * align for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(n, device=dev)
        self.grad = torch.zeros(n, device=dev)
        self.exp_avg = torch.zeros(n, device=dev)
        self.exp_avg_sq = torch.zeros(n, device=dev)
        off = 0
        self.grad_views = []
        for p in self.params:
            k = p.numel()
            self.flat[off:off + k].copy_(p.data.reshape(-1))
            p.data = self.flat[off:off + k].view_as(p)
            self.grad_views.append(self.grad[off:off + k].view_as(p))
            p.grad = self.grad_views[-1]
            off += (k + align - 1) // align * align
        self.lr, self.wd, self.eps, self.betas, self.max_norm = lr, weight_decay, eps, betas, max_norm
        self.step_count = 0
        self.sumsq = torch.zeros(1, device=dev)
        self.scale = torch.ones(1, device=dev)
        # per-step scalars of the graph-replayable update: {lr, 1 - beta1^t, sqrt(1 - beta2^t), gradient prescale}.
        # They are staged through a RING of pinned slots, one per step, each guarded by the event of its last upload:
        # the CPU runs ahead of the GPU, and a single pinned buffer would be overwritten for step t+k before the
        # asynchronous copy of step t has read it.
        self.prescale = 1.0          # 1 / world size when the all-reduce sums per-rank MEAN losses (DDP semantics)
        self.hyper = torch.zeros(4, device=dev)
        self.HYPER_SLOTS = 64
        on_gpu = dev.type == "cuda"
        self.hyper_ring = torch.zeros(self.HYPER_SLOTS, 4).pin_memory() if on_gpu else torch.zeros(self.HYPER_SLOTS, 4)
        self.hyper_events = [None] * self.HYPER_SLOTS

    def zero_grad(self):
        self.grad.zero_()
        for p, g in zip(self.params, self.grad_views):
            p.grad = g

    def detach_grads(self):
        """Before backward: autograd then *assigns* each parameter's gradient instead of launching one in-place add
        per parameter into the flat buffer."""
        for p in self.params:
            p.grad = None

    def gather_grads(self, accumulate: bool = False):
        """After backward: one memset + one multi-tensor copy bring the gradients into the flat buffer (the layout the
        clip, AdamW and the NCCL all-reduce work on).  `accumulate`: add to what the flat buffer holds instead (gradient
        accumulation over the micro-batches of one step; the caller zeroes the buffer before the first one)."""
        if not accumulate:
            self.grad.zero_()
        src, dst = [], []
        for p, g in zip(self.params, self.grad_views):
            if p.grad is not None:
                src.append(p.grad)
                dst.append(g)
            p.grad = g
        if src:
            if accumulate:
                torch._foreach_add_(dst, src)
            else:
                torch._foreach_copy_(dst, src)

    def step(self, lr_factor: float = 1.0):
        self.step_count += 1
        ops.clear_pack_cache()       # the packed layer operands are functions of the parameters this call rewrites
        scale = None
        if self.max_norm is not None:
            # clip_grad_norm_(max_norm, error_if_nonfinite=False): coef = max_norm / (norm + 1e-6), clamped to 1
            self.sumsq.zero_()
            ops.sumsq(self.grad, self.sumsq)
            torch.clamp(self.max_norm / (self.sumsq.sqrt() * self.prescale + 1e-6), max=1.0, out=self.scale)
            if self.prescale != 1.0:
                self.scale.mul_(self.prescale)
            scale = self.scale
        elif self.prescale != 1.0:
            self.scale.fill_(self.prescale)
            scale = self.scale
        ops.adamw_step(self.flat, self.grad, self.exp_avg, self.exp_avg_sq, scale, self.lr * lr_factor, self.betas[0],
                       self.betas[1], self.eps, self.wd, self.step_count)


    def advance(self, lr_factor: float = 1.0):
        """Host side of `step_dev`: bumps the step count and uploads this step's scalars (outside any graph)."""
        self.step_count += 1
        slot = self.step_count % self.HYPER_SLOTS
        ev = self.hyper_events[slot]
        if ev is not None:
            ev.synchronize()         # only ever waits when the CPU is HYPER_SLOTS steps ahead of the GPU
        host = self.hyper_ring[slot]
        host[0] = self.lr * lr_factor
        host[1] = 1.0 - self.betas[0] ** self.step_count
        host[2] = math.sqrt(1.0 - self.betas[1] ** self.step_count)
        host[3] = self.prescale
        self.hyper.copy_(host, non_blocking=True)
        if self.hyper.is_cuda:
            ev = self.hyper_events[slot] or torch.cuda.Event()
            ev.record()
            self.hyper_events[slot] = ev

    def step_dev(self):
        """Clip + AdamW with every step-dependent scalar read from device memory (CUDA-graph replayable)."""
        ops.clear_pack_cache()       # the packed layer operands are functions of the parameters this call rewrites
        if self.max_norm is not None:
            self.sumsq.zero_()
            ops.sumsq(self.grad, self.sumsq)
        ops.adamw_step_dev(self.flat, self.grad, self.exp_avg, self.exp_avg_sq, self.sumsq if self.max_norm is not None else None,
                           self.max_norm, self.hyper, self.betas[0], self.betas[1], self.eps, self.wd)


def multistep_warmup_factor(step, num_iterations, warm_up_ratio=0.1, milestones=(0.5, 0.75, 0.9), gamma=0.4):
    """ref: engine/schedulers.py:249-270"""
    warm_up_end = int(num_iterations * warm_up_ratio)
    if step < warm_up_end:
        return step / warm_up_end
    return gamma ** int(np.searchsorted(milestones, step / num_iterations, side="left"))


class RawPipeline:
    """train_step of the reference's RawPipeline / BasePipeline on the B200 path.
    `raw=True`: mosaicked frames (grid_raw); `raw=False`: demosaicked (grid)."""

    def __init__(self, modalities: Dict[str, int], cameras: Dict[str, Cameras], device="cuda", raw=True,
                 max_num_iterations=100000, pose_mode="SO3xR3", shared_pose=True, render_all_heads=False,
                 process_group=None, preset: Optional[str] = None, **model_kwargs):
        self.device, self.raw, self.modalities = torch.device(device), raw, modalities
        self.max_num_iterations = max_num_iterations
        self.preset = preset or ("grid_raw" if raw else "grid")
        self.model = build_model(self.preset, modalities=modalities, render_all_heads=render_all_heads,
                                 **model_kwargs).to(self.device)
        n_cam = len(next(iter(cameras.values())))
        self.camera_optimizer = CameraOptimizerConfig(mode=pose_mode, shared_optimization=shared_pose,
                                                      modalities_to_optimize={m: True for m in modalities}
                                                      ).setup(num_cameras=n_cam).to(self.device)
        self.ray_generator = RayGenerator({m: {"cameras": c.to(self.device)} for m, c in cameras.items()},
                                          self.camera_optimizer, pixel_offset=0.0)
        self.loss_manager = loss_config_for(self.preset).setup(modalities=list(modalities), num_iterations=max_num_iterations,
                                                               model=self.model)
        self.patterns = {m: torch.tensor(MOSAICK_PATTERNS[m], dtype=torch.int32, device=self.device) for m in modalities} if raw else None
        self.optimizers = {"fields": FlatAdamW(list(self.model.parameters()), lr=1e-3)}
        pose_params = list(self.camera_optimizer.parameters())
        if pose_params:
            self.optimizers["camera_poses"] = FlatAdamW(pose_params, lr=1e-4)

        class _T:
            pass
        t = _T(); t.max_num_iterations = max_num_iterations
        self.callbacks = self.model.get_training_callbacks(TrainingCallbackAttributes(model=self.model, trainer=t))
        self.process_group = process_group
        self.model.train()
        self._graphs = {}           # schedule key -> (static coords, static targets, static count, fwd+bwd graph, losses, total)
        self._g_opt = None          # clip + AdamW graph (no step-dependent launch argument: one capture serves every step)
        self._pool = None           # memory pool shared by the pipeline's graphs
        self._last_key = None
        self.graph_launches = 0     # kernels of libmms_b200.so captured in one (micro-)batch's forward + backward graph
        self.inputs_event = torch.cuda.Event() if self.device.type == "cuda" else None

    def run_callbacks(self, step):
        for cb in self.callbacks:
            cb.run_callback_at_location(step, TrainingCallbackLocation.BEFORE_TRAIN_ITERATION)

    def forward_backward(self, coords, targets, step, loss_scales=None, geometry_count=None, accumulate=False):
        """One forward + loss + backward over a (micro-)batch; gradients land in the optimizers' flat buffers
        (`accumulate`: added to what they hold).  `loss_scales` / `geometry_count`: normalisation by the GLOBAL counts
        when the batch is one shard of a step (see ShardPlan; LossManager.compute_loss)."""
        ops.clear_pack_cache()
        with torch.nn.utils.parametrize.cached():     # one weight-norm evaluation (and one operand pack) per step
            ray_bundles = self.ray_generator(coords)
            outputs = self.model(ray_bundles)
        losses, total = self.loss_manager.compute_loss(outputs, targets, coords, step, mosaick_patterns=self.patterns,
                                                       loss_scales=loss_scales, geometry_count=geometry_count)
        for opt in self.optimizers.values():
            opt.detach_grads()
        total.backward()
        for opt in self.optimizers.values():
            opt.gather_grads(accumulate)
        # detached: a caller holding on to the losses must not keep the autograd graph (and its AccumulateGrad
        # nodes, which remember the stream they were created on) alive into the next step / a graph capture
        return {k: (v.detach() if torch.is_tensor(v) else v) for k, v in losses.items()}, total.detach()

    def _world(self):
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_world_size(self.process_group)
        return 1

    def all_reduce_gradients(self, mean: bool = True):
        """The gradient exchange of the reference's DDP, see `all_reduce_optimizers`."""
        all_reduce_optimizers(self.optimizers, self.process_group, mean)

    def optimizer_step(self, step):
        f = multistep_warmup_factor(step, self.max_num_iterations)
        for opt in self.optimizers.values():
            opt.step(lr_factor=f)

    def train_step(self, step, coords, targets):
        """coords {mod: int32 [R,3]}, targets {mod: [R,1] raw or [R,C]} — already on the device."""
        self.run_callbacks(step)
        losses, total = self.forward_backward(coords, targets, step)
        self.all_reduce_gradients()
        self.optimizer_step(step)
        return losses, total

    @torch.no_grad()
    def count_unmasked_samples(self, coords) -> torch.Tensor:
        """Device fp32 [1]: samples per ray x the number of rays of `coords` (all modalities) that hit the sphere — the
        denominator of the geometry losses (losses.py:113-150 average over the compacted in-sphere samples).  Ray
        generation + collider only: a few small launches."""
        bundles = self.ray_generator(coords)
        total = torch.zeros((), device=self.device)
        for rb in bundles.values():
            if rb is not None:
                total = total + ops.sphere_collide(rb.origins, rb.directions, self.model.collider.collider.radius)[2].sum(dtype=torch.float32)
        s = self.model.ray_sampler.config.num_samples + self.model.ray_sampler.config.num_samples_importance
        return (total * float(s)).reshape(1)

    def train_step_sharded(self, step, coords, targets, plan: ShardPlan, graphed: bool = True):
        """One optimizer step over this rank's slice of a global batch (SURVEY §8(e) "strong" mode): `coords` / `targets`
        hold the rank's LOCAL rows of every modality (device or pinned host memory), `plan` cuts them into micro-batches
        whose gradients accumulate in the flat buffers; every loss is normalised by the global counts (rays per
        modality: static; in-sphere samples: counted on the device and summed over the ranks), so the summing all-reduce
        yields exactly the gradient of the unsharded batch.  Returns (losses of the last micro-batch, this rank's share
        of the total loss)."""
        self.run_callbacks(step)
        dc = {m: c.to(self.device, non_blocking=True) for m, c in coords.items()}
        dt = {m: t.to(self.device, non_blocking=True) for m, t in targets.items()}
        self.inputs_event.record()       # the caller may reuse its pinned buffers once this event has completed
        count = self.count_unmasked_samples(dc)
        all_reduce_flat(count, self.process_group)
        for opt in self.optimizers.values():
            opt.grad.zero_()
        total_sum, losses = torch.zeros((), device=self.device), None
        for j in range(len(plan)):
            cs, ts = plan.slice(j, dc, "local"), plan.slice(j, dt, "local")
            scales = plan.loss_scales(j)
            if graphed:
                losses, total = self._fb_graphed(step, cs, ts, scales, count)
            else:
                losses, total = self.forward_backward(cs, ts, step, scales, count, accumulate=True)
            total_sum = total_sum + total
        self.all_reduce_gradients(mean=False)
        if graphed:
            self._optimizer_graphed(step)
        else:
            self.optimizer_step(step)
        return losses, total_sum

    # ---- full-frame inference (utils/eval_utils.py:31-75, engine/evaluator.py:619-746) ----------
    @torch.no_grad()
    def render(self, coords, chunk_rays: Optional[int] = None):
        """Renders pixel coordinates {mod: int32 [R,3]} in eval mode (deterministic sampling, no Hessian) in chunks of
        `chunk_rays` rays per modality (default: 4 Mi samples per chunk over all modalities), all modalities of a chunk
        as one batch; returns {mod: [R, C]} (the modality's own head; mosaicked pipelines select the pattern's channel
        per pixel like evaluator.py:721-746)."""
        if chunk_rays is None:
            cfg = self.model.ray_sampler.config
            chunk_rays = max(256, (1 << 22) // ((cfg.num_samples + cfg.num_samples_importance) * max(len(coords), 1)))
        was_training = self.model.training
        self.model.eval()
        ops.clear_pack_cache()
        out = {m: [] for m in coords}
        n_max = max(c.shape[0] for c in coords.values())
        try:
            with torch.nn.utils.parametrize.cached():
                for a in range(0, n_max, chunk_rays):
                    part = {m: c[a:a + chunk_rays] for m, c in coords.items() if c.shape[0] > a}
                    bundles = self.ray_generator(part)
                    outputs = self.model(bundles)
                    for m in part:
                        out[m].append(outputs[m][m])
        finally:
            self.model.train(was_training)
        rendered = {m: torch.cat(v, 0) for m, v in out.items() if v}
        if self.raw:
            selected = {}
            for m, r in rendered.items():
                pat = self.patterns[m]
                _, sel = ops.mosaick_bands(coords[m], pat.reshape(-1), pat.shape[0], pat.shape[1], r)
                selected[m] = sel[:, None]
            return selected
        return rendered

    # ---- SDF volume sweep for mesh extraction (utils/marching_cubes.py:34-188) ---------------------
    @torch.no_grad()
    def sdf_volume(self, resolution: int = 256, bounding_box_min=(-1.0, -1.0, -1.0), bounding_box_max=(1.0, 1.0, 1.0),
                   chunk_points: int = 1 << 21):
        """sdf on the regular grid linspace(min, max, resolution)^3 (indexing "ij" like get_surface_sliding) ->
        [resolution]^3 fp32.  A pure consumer of the fused SDF forward (encodings + layer 0 + layer 1 with the sdf head in
        its epilogue; no activation is stored); marching cubes itself stays with the caller."""
        dev = self.device
        axes = [torch.linspace(float(a), float(b), resolution, device=dev) for a, b in zip(bounding_box_min, bounding_box_max)]
        ops.clear_pack_cache()
        out = torch.empty((resolution ** 3,), device=dev, dtype=torch.float32)
        yz = torch.cartesian_prod(axes[1], axes[2])                       # [res^2, 2], y-major
        slabs = max(1, chunk_points // (resolution * resolution))
        field = self.model.surface_model.surface_field
        with torch.nn.utils.parametrize.cached():
            for i0 in range(0, resolution, slabs):
                xs = axes[0][i0:i0 + slabs]
                pts = torch.cat([xs[:, None, None].expand(-1, yz.shape[0], 1), yz[None].expand(xs.shape[0], -1, -1)], -1)
                sdf = field.single_output(pts.reshape(-1, 3))
                out[i0 * resolution * resolution:(i0 + xs.shape[0]) * resolution * resolution] = sdf.reshape(-1)
        return out.reshape(resolution, resolution, resolution)

    # ---- the same step replayed from CUDA graphs ------------------------------------------------------------
    def _schedule_key(self, step, coords, targets, scales=None):
        """Everything a captured step bakes into its launch arguments: the schedule state the callbacks set
        (level mask, delta, anneal), the loss weights of this step, the loss normalisation and the batch shapes.
        lr and the AdamW bias corrections are NOT baked (device scalars, FlatAdamW.advance), nor is the global
        sample count of the geometry losses (a device scalar)."""
        sm = self.model.surface_model
        levels = tuple(int(m.active_level) for m in self.model.modules() if hasattr(m, "active_level"))
        weights = tuple(float(w) for w in self.loss_manager.weights(step))
        shapes = tuple((m, tuple(c.shape), tuple(targets[m].shape)) for m, c in coords.items())
        sc = tuple(sorted(scales.items())) if scales is not None else None
        delta = sm.numerical_gradients_delta              # None for the analytic-gradient presets (`mlp*`)
        return (float(delta) if delta is not None else 0.0, float(sm.volume_rendering._cos_anneal_ratio), levels, weights, shapes,
                sc, ops.MLP_PRECISION)

    def train_step_graphed(self, step, coords, targets):
        """train_step with the ~1000 launches of a step replayed from two CUDA graphs (forward + backward | clip +
        AdamW; the NCCL all-reduce sits between them).  A step whose schedule key differs from the previous step's
        runs eagerly (early training: the anneal ratio moves every step); the second step with the same key captures.
        coords / targets may live in pinned host memory: they are copied into the graphs' static inputs; the caller
        may overwrite its pinned buffers once `self.inputs_event` has completed."""
        self.run_callbacks(step)
        losses, total = self._fb_graphed(step, coords, targets, None, None)
        self.inputs_event.record()
        self.all_reduce_gradients()
        self._optimizer_graphed(step)
        return losses, total

    def _fb_graphed(self, step, coords, targets, scales, count):
        """forward + backward of one (micro-)batch from the graph captured for its key; `scales is None`: a whole
        step's batch (gradients overwrite the flat buffers), otherwise one shard (gradients accumulate)."""
        accumulate = scales is not None
        if not self.loss_manager.graph_capturable:        # host-side random draws every step (preset grid_decimated)
            dc = {m: c.to(self.device, non_blocking=True) for m, c in coords.items()}
            dt = {m: t.to(self.device, non_blocking=True) for m, t in targets.items()}
            return self.forward_backward(dc, dt, step, scales, count, accumulate)
        key = self._schedule_key(step, coords, targets, scales)
        g = self._graphs.get(key)
        if g is None:
            if self._last_key != key:
                self._last_key = key
                dc = {m: c.to(self.device, non_blocking=True) for m, c in coords.items()}
                dt = {m: t.to(self.device, non_blocking=True) for m, t in targets.items()}
                return self.forward_backward(dc, dt, step, scales, count, accumulate)
            g = self._capture(key, step, coords, targets, scales, count)
        sc, st, scount, g_fb, losses, total = g
        for m in sc:
            sc[m].copy_(coords[m], non_blocking=True)
            st[m].copy_(targets[m], non_blocking=True)
        if scount is not None:
            scount.copy_(count)
        g_fb.replay()
        return losses, total

    def _optimizer_graphed(self, step):
        f = multistep_warmup_factor(step, self.max_num_iterations)
        for opt in self.optimizers.values():
            opt.advance(lr_factor=f)
        if self._g_opt is None:
            self._g_opt = torch.cuda.CUDAGraph()
            # warm-up outside the capture (cudaFuncSetAttribute etc.), on values that are restored afterwards
            with torch.cuda.graph(self._g_opt, pool=self._pool):
                for opt in self.optimizers.values():
                    opt.step_dev()
            if self._pool is None:
                self._pool = self._g_opt.pool()
        self._g_opt.replay()

    def _capture(self, key, step, coords, targets, scales, count):
        from . import _lib
        accumulate = scales is not None
        sc = {m: torch.empty(c.shape, dtype=c.dtype, device=self.device) for m, c in coords.items()}
        st = {m: torch.empty(t.shape, dtype=t.dtype, device=self.device) for m, t in targets.items()}
        scount = torch.empty_like(count) if count is not None else None
        for m in sc:
            sc[m].copy_(coords[m])
            st[m].copy_(targets[m])
        if scount is not None:
            scount.copy_(count)
        # warm-up on a side stream (torch's whole-network-capture recipe): first-use initialisation
        # (cudaFuncSetAttribute, constant caches) happens outside the capture.  Its gradients must not count twice:
        # the flat buffers are saved and restored around it when this capture belongs to an accumulating step.
        saved = [opt.grad.clone() for opt in self.optimizers.values()] if accumulate else None
        side = torch.cuda.Stream(self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            self.forward_backward(sc, st, step, scales, scount, accumulate)
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        if saved is not None:
            for opt, g in zip(self.optimizers.values(), saved):
                opt.grad.copy_(g)
            del saved
        if len(self._graphs) >= 4:              # a captured step owns its activations: keep the cache small
            self._graphs.pop(next(iter(self._graphs)))
        l0 = _lib.launch_count()
        g_fb = torch.cuda.CUDAGraph()
        # all graphs of this pipeline share ONE memory pool: they are replayed one after the other on one stream and
        # none reads another's outputs after a later replay (the losses are consumed right behind each replay)
        with torch.cuda.graph(g_fb, pool=self._pool):
            losses, total = self.forward_backward(sc, st, step, scales, scount, accumulate)
        if self._pool is None:
            self._pool = g_fb.pool()
        self.graph_launches = _lib.launch_count() - l0
        torch.cuda.synchronize(self.device)
        if accumulate:
            # the capture itself did not execute: replaying is the caller's job (it returns through the normal path)
            pass
        self._graphs[key] = (sc, st, scount, g_fb, losses, total)
        return self._graphs[key]
