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
from .models import MODALITY_CHANNELS, MOSAICK_PATTERNS, build_model, grid_loss_config

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
        # per-step scalars of the graph-replayable update: {lr, 1 - beta1^t, sqrt(1 - beta2^t)}
        self.hyper = torch.zeros(3, device=dev)
        self.hyper_host = torch.zeros(3).pin_memory() if dev.type == "cuda" else torch.zeros(3)

    def zero_grad(self):
        self.grad.zero_()
        for p, g in zip(self.params, self.grad_views):
            p.grad = g

    def detach_grads(self):
        """Before backward: autograd then *assigns* each parameter's gradient instead of launching one in-place add
        per parameter into the flat buffer."""
        for p in self.params:
            p.grad = None

    def gather_grads(self):
        """After backward: one memset + one multi-tensor copy bring the gradients into the flat buffer (the layout the
        clip, AdamW and the NCCL all-reduce work on)."""
        self.grad.zero_()
        src, dst = [], []
        for p, g in zip(self.params, self.grad_views):
            if p.grad is not None:
                src.append(p.grad)
                dst.append(g)
            p.grad = g
        if src:
            torch._foreach_copy_(dst, src)

    def step(self, lr_factor: float = 1.0):
        self.step_count += 1
        scale = None
        if self.max_norm is not None:
            # clip_grad_norm_(max_norm, error_if_nonfinite=False): coef = max_norm / (norm + 1e-6), clamped to 1
            self.sumsq.zero_()
            ops.sumsq(self.grad, self.sumsq)
            torch.clamp(self.max_norm / (self.sumsq.sqrt() + 1e-6), max=1.0, out=self.scale)
            scale = self.scale
        ops.adamw_step(self.flat, self.grad, self.exp_avg, self.exp_avg_sq, scale, self.lr * lr_factor, self.betas[0],
                       self.betas[1], self.eps, self.wd, self.step_count)


    def advance(self, lr_factor: float = 1.0):
        """Host side of `step_dev`: bumps the step count and uploads this step's scalars (outside any graph)."""
        self.step_count += 1
        self.hyper_host[0] = self.lr * lr_factor
        self.hyper_host[1] = 1.0 - self.betas[0] ** self.step_count
        self.hyper_host[2] = math.sqrt(1.0 - self.betas[1] ** self.step_count)
        self.hyper.copy_(self.hyper_host, non_blocking=True)

    def step_dev(self):
        """Clip + AdamW with every step-dependent scalar read from device memory (CUDA-graph replayable)."""
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
                 process_group=None, **model_kwargs):
        self.device, self.raw, self.modalities = torch.device(device), raw, modalities
        self.max_num_iterations = max_num_iterations
        self.model = build_model("grid_raw" if raw else "grid", modalities=modalities, render_all_heads=render_all_heads,
                                 **model_kwargs).to(self.device)
        n_cam = len(next(iter(cameras.values())))
        self.camera_optimizer = CameraOptimizerConfig(mode=pose_mode, shared_optimization=shared_pose,
                                                      modalities_to_optimize={m: True for m in modalities}
                                                      ).setup(num_cameras=n_cam).to(self.device)
        self.ray_generator = RayGenerator({m: {"cameras": c.to(self.device)} for m, c in cameras.items()},
                                          self.camera_optimizer, pixel_offset=0.0)
        self.loss_manager = grid_loss_config().setup(modalities=list(modalities), num_iterations=max_num_iterations,
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
        self._graph = None          # (key, static coords, static targets, fwd+bwd graph, optimizer graph, losses, total)
        self._last_key = None
        self.graph_launches = 0     # kernels of libmms_b200.so captured in one step's graphs

    def run_callbacks(self, step):
        for cb in self.callbacks:
            cb.run_callback_at_location(step, TrainingCallbackLocation.BEFORE_TRAIN_ITERATION)

    def forward_backward(self, coords, targets, step):
        ops.clear_pack_cache()
        with torch.nn.utils.parametrize.cached():     # one weight-norm evaluation (and one operand pack) per step
            ray_bundles = self.ray_generator(coords)
            outputs = self.model(ray_bundles)
        losses, total = self.loss_manager.compute_loss(outputs, targets, coords, step, mosaick_patterns=self.patterns)
        for opt in self.optimizers.values():
            opt.detach_grads()
        total.backward()
        for opt in self.optimizers.values():
            opt.gather_grads()
        # detached: a caller holding on to the losses must not keep the autograd graph (and its AccumulateGrad
        # nodes, which remember the stream they were created on) alive into the next step / a graph capture
        return {k: (v.detach() if torch.is_tensor(v) else v) for k, v in losses.items()}, total.detach()

    def all_reduce_gradients(self):
        """DDP semantics of the reference (mean over ranks) — one flat NCCL all-reduce per optimizer."""
        import torch.distributed as dist
        if self.process_group is None and not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
            return
        ws = dist.get_world_size(self.process_group)
        for opt in self.optimizers.values():
            dist.all_reduce(opt.grad, op=dist.ReduceOp.SUM, group=self.process_group)
            opt.grad.mul_(1.0 / ws)

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

    # ---- full-frame inference (utils/eval_utils.py:31-75, engine/evaluator.py:619-746) ----------
    @torch.no_grad()
    def render(self, coords, chunk_rays: int = 32768):
        """Renders pixel coordinates {mod: int32 [R,3]} in eval mode (deterministic sampling, no Hessian) in chunks of
        `chunk_rays` rays per modality, all modalities of a chunk as one batch; returns {mod: [R, C]} (the modality's
        own head; mosaicked pipelines select the pattern's channel per pixel like evaluator.py:721-746)."""
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
    def _schedule_key(self, step, coords, targets):
        """Everything a captured step bakes into its launch arguments: the schedule state the callbacks set
        (level mask, delta, anneal), the loss weights of this step and the batch shapes.  lr and the AdamW bias
        corrections are NOT baked (device scalars, FlatAdamW.advance)."""
        sm = self.model.surface_model
        levels = tuple(int(m.active_level) for m in self.model.modules() if hasattr(m, "active_level"))
        weights = tuple(float(w) for w in self.loss_manager.weights(step))
        shapes = tuple((m, tuple(c.shape), tuple(targets[m].shape)) for m, c in coords.items())
        return (float(sm.numerical_gradients_delta), float(sm.volume_rendering._cos_anneal_ratio), levels, weights, shapes,
                ops.MLP_PRECISION)

    def train_step_graphed(self, step, coords, targets):
        """train_step with the ~4000 launches of a step replayed from two CUDA graphs (forward + backward | clip +
        AdamW; the NCCL all-reduce sits between them).  A step whose schedule key differs from the previous step's
        runs eagerly (early training: the anneal ratio moves every step); the second step with the same key captures.
        coords / targets may live in pinned host memory: they are copied into the graphs' static inputs."""
        self.run_callbacks(step)
        key = self._schedule_key(step, coords, targets)
        if self._graph is None or self._graph[0] != key:
            if self._last_key != key:
                self._last_key = key
                self._graph = None
                dc = {m: c.to(self.device, non_blocking=True) for m, c in coords.items()}
                dt = {m: t.to(self.device, non_blocking=True) for m, t in targets.items()}
                losses, total = self.forward_backward(dc, dt, step)
                self.all_reduce_gradients()
                self.optimizer_step(step)
                return losses, total
            self._capture(key, step, coords, targets)
        _, sc, st, g_fb, g_opt, losses, total = self._graph
        for m in sc:
            sc[m].copy_(coords[m], non_blocking=True)
            st[m].copy_(targets[m], non_blocking=True)
        f = multistep_warmup_factor(step, self.max_num_iterations)
        for opt in self.optimizers.values():
            opt.advance(lr_factor=f)
        g_fb.replay()
        self.all_reduce_gradients()
        g_opt.replay()
        return losses, total

    def _capture(self, key, step, coords, targets):
        from . import _lib
        sc = {m: torch.empty(c.shape, dtype=c.dtype, device=self.device) for m, c in coords.items()}
        st = {m: torch.empty(t.shape, dtype=t.dtype, device=self.device) for m, t in targets.items()}
        for m in sc:
            sc[m].copy_(coords[m])
            st[m].copy_(targets[m])
        # warm-up on a side stream (torch's whole-network-capture recipe): first-use initialisation
        # (cudaFuncSetAttribute, constant caches) happens outside the capture
        side = torch.cuda.Stream(self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            self.forward_backward(sc, st, step)
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        l0 = _lib.launch_count()
        g_fb = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g_fb):
            losses, total = self.forward_backward(sc, st, step)
        g_opt = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g_opt, pool=g_fb.pool()):
            for opt in self.optimizers.values():
                opt.step_dev()
        self.graph_launches = _lib.launch_count() - l0
        self._graph = (key, sc, st, g_fb, g_opt, losses, total)
