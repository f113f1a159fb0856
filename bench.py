#!/usr/bin/env python
"""bench.py — train rays/s (forward + backward + clip + AdamW) of the MMS-FW per-ray rendering hot path.

    python bench.py --gpus N --steps K --warmup W [--workload sweep|grid_raw|grid|grid_bg|mlp_raw] [--scaling strong|weak]
                    [--impl reference]

One "step" = one optimizer step of the hot path over one synthetic batch: pixel sampling -> ray generation -> NeuS
sampling -> hash grids + MLPs -> compositing -> (mosaick-aware) losses -> backward -> (gradient all-reduce) ->
global-norm clip -> AdamW: what the reference's `train_step` timer brackets (engine/trainer.py:107-114).

Default workload = the north-star one (BASELINE.json configs[4]): 5 raw modalities, 256 samples per ray, a GLOBAL batch
of 5 x 13112 = 65560 rays per step ("65536" rounded up so that 8 ranks x 5 modalities get equal slices), split
contiguously over the ranks (strong scaling) and, on each rank, into micro-batches of 8195 rays whose
gradients accumulate.  Prints ONE JSON line (see the driver contract).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

FIVE = ["rgb", "infrared", "mono", "polarization", "multispectral"]
WORKLOADS = {
    # BASELINE.json configs[4] (the north-star workload): 5 modalities raw, 65536 rays x 256 samples.
    # rays = GLOBAL batch of one optimizer step (strong) / per-GPU batch (weak: `weak_rays`); micro = rays per micro-batch
    "sweep": dict(modalities=FIVE, rays=5 * 13112, weak_rays=5 * 1639, micro=5 * 1639, n_c=128, n_i=128, bg=16, raw=True),
    # configs[2]: confs/grid_raw.yaml, 5 modalities, mosaick-aware loss, 8192 rays (ragged split), 64 samples
    "grid_raw": dict(modalities=FIVE, rays=8192, weak_rays=8192, micro=8192, n_c=32, n_i=32, bg=16, raw=True),
    # configs[1]: confs/grid.yaml, RGB + 1 extra modality, demosaicked, 4096 rays x 128 samples
    "grid": dict(modalities=["rgb", "infrared"], rays=4096, weak_rays=4096, micro=4096, n_c=64, n_i=64, bg=16, raw=False),
    # configs[3]: confs/grid_raw_rgb_all_views_pol_10_views.yaml (hash-grid background preset), RGB + polarization, 16384 rays
    "grid_bg": dict(modalities=["rgb", "polarization"], rays=16384, weak_rays=2048, micro=16384, n_c=32, n_i=32, bg=16, raw=True,
                    preset="grid_raw_grid_bg_unbalanced", yaml="grid_raw_rgb_all_views_pol_10_views.yaml", oracle=dict(bg_grid=True)),
    # configs[0]: confs/mlp_raw.yaml (MLP fields, analytic SDF gradients), 2 modalities, 1024 rays x 64 samples
    "mlp_raw": dict(modalities=["rgb", "mono"], rays=1024, weak_rays=1024, micro=1024, n_c=32, n_i=32, bg=16, raw=True,
                    preset="mlp_raw", yaml="mlp_raw.yaml", oracle=dict(field="mlp")),
}


def preset_of(wl):
    return wl.get("preset") or ("grid_raw" if wl["raw"] else "grid")
BASE_STEP = 60000       # late in the 100k-iteration schedule: all 16 levels active, delta = 2/1024, anneal = 1


def split_rays(total, mods):
    base, rem = divmod(total, len(mods))
    return {m: base + (1 if i < rem else 0) for i, m in enumerate(mods)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1590.0, 1400.0, "FALLBACK (B200_PROFILING.md figures; MEASURED_PEAKS.json absent)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------------
# CPU arms (test infrastructure used as the reported baseline only: never on the product path)
# ------------------------------------------------------------------------------------------------------------------
def _synthetic_inputs(wl, rays_per_mod, seed=654824):
    from multimodalstudio_b200.models import MODALITY_CHANNELS
    from multimodalstudio_b200.pipelines import SyntheticScene
    mods = {m: MODALITY_CHANNELS[m] for m in wl["modalities"]}
    return mods, SyntheticScene(mods, {m: rays_per_mod for m in mods}, raw=wl["raw"], seed=seed)


def cpu_port_arm(wl, steps, warmup, rays_per_mod, device="cpu"):
    """The reference's algorithm restated in plain PyTorch (oracle/mms_oracle.py, bit-exact against the reference on the
    build container) over a BOUNDED sample of the workload: forward + channel select + losses + backward.
    device="cuda": the same torch program on the GPU (cuBLAS + ATen kernels) = BASELINE.md §3b's "reference torch path on
    the B200" side figure."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import mms_oracle as O
    from multimodalstudio_b200.models import MOSAICK_PATTERNS, build_model
    mods, scene = _synthetic_inputs(wl, rays_per_mod)
    model = build_model(preset_of(wl), modalities=mods, num_samples=wl["n_c"], num_samples_importance=wl["n_i"], bg_samples=wl["bg"])
    dev = torch.device(device)
    sd = {k: v.detach().to(dev).requires_grad_(True) for k, v in model.state_dict().items()}
    del model
    cfg = O.default_cfg(modalities=mods, num_samples=wl["n_c"], num_samples_importance=wl["n_i"], bg_samples=wl["bg"],
                        **wl.get("oracle", {}))
    cams = {m: (c.camera_to_worlds.to(dev), c.intrinsics.to(dev), c.distortion_params.to(dev)) for m, c in scene.cameras.items()}
    times = []
    batches = [scene.sample_batch() for _ in range(warmup + steps)]       # drawn with the CPU generator
    with torch.device(dev):           # the oracle's factory calls (linspace, arange, ...) follow the inputs' device
        orc = O.GridModelOracle(sd, cfg)
        for it in range(warmup + steps):
            coords, targets = batches[it]
            coords = {m: c.to(dev) for m, c in coords.items()}
            targets = {m: t.to(dev) for m, t in targets.items()}
            if dev.type == "cuda":
                torch.cuda.synchronize()
            t0 = time.perf_counter()
            outputs = {}
            for mod in mods:
                c2w, intr, dist = cams[mod]
                r = O.raygen(coords[mod], c2w, intr, dist, None)
                n_hit = int(O.sphere_collide(r["origins"], r["directions"])[2].sum())
                rand = {"uniform": torch.rand(n_hit, 1), "pdf": [torch.rand(n_hit, 1) for _ in range(4)],
                        "background": torch.rand(rays_per_mod, wl["bg"] + 1)}
                outputs[mod] = orc.forward_modality(mod, r["origins"], r["directions"], r["up_directions"], rand)
            _, total = orc.loss(outputs, targets, coords, MOSAICK_PATTERNS if wl["raw"] else None, 5e-4 * 0.0157)
            for v in sd.values():
                v.grad = None
            total.backward()
            if dev.type == "cuda":
                torch.cuda.synchronize()
            if it >= warmup:
                times.append(time.perf_counter() - t0)
    n_rays = rays_per_mod * len(mods)
    ms = 1e3 * sum(times) / len(times)
    return n_rays / (ms / 1e3), ms, n_rays


def cpu_reference_arm(wl, steps, warmup, rays_per_mod):
    """The UNMODIFIED reference (/root/reference imported through oracle/ref_harness.py) on the host CPU: its own
    BaseModel.forward, channel select, LossManager and autograd backward, on the same bounded sample.  Only possible
    where the reference tree exists (the build container); returns None elsewhere."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    try:
        import ref_harness as RH
        if not RH.reference_available():
            return None
        import math
        import types
        RH.import_reference()
        from cameras.rays import RayBundle
        from pipelines.raw_pipeline import RawPipeline as RefRawPipeline
        import mms_oracle as O
        from multimodalstudio_b200.models import MOSAICK_PATTERNS
        mods, scene = _synthetic_inputs(wl, rays_per_mod)
        model, tc = RH.build_reference_model(preset=preset_of(wl), yaml_name=wl.get("yaml") or ("grid_raw.yaml" if wl["raw"] else "grid.yaml"),
                                             modalities=mods, num_samples=wl["n_c"], num_samples_importance=wl["n_i"], bg_samples=wl["bg"])
        RH.set_schedule_state(model, 16, 2.0 / 1024, 1.0)
        model.train()
        lm = tc.pipeline.loss_manager.setup(modalities=list(mods), num_iterations=tc.max_num_iterations, model=model)
        from multimodalstudio_b200.pipelines import MODALITY_SENSORS
        masks = {}
        for m in mods:
            w, h, _ = MODALITY_SENSORS[m]
            pat = torch.tensor(MOSAICK_PATTERNS[m])
            masks[m] = pat.repeat((math.ceil(h / pat.shape[0]), math.ceil(w / pat.shape[1])))[:h, :w].type(torch.int8)
        stub = types.SimpleNamespace(datamanager=types.SimpleNamespace(
            modalities=dict(mods), train_dataset=types.SimpleNamespace(mosaick_mask_per_modality=masks)))
        times = []
        for it in range(warmup + steps):
            coords, targets = scene.sample_batch()
            t0 = time.perf_counter()
            bundles = {}
            for mod in mods:
                cam = scene.cameras[mod]
                r = O.raygen(coords[mod], cam.camera_to_worlds, cam.intrinsics, cam.distortion_params, None)   # ray generation: the port (needs no dataset objects)
                bundles[mod] = RayBundle(camera_indices=coords[mod][:, :1].long(), origins=r["origins"], directions=r["directions"],
                                         up_directions=r["up_directions"], pixel_area=r["pixel_area"], directions_norm=r["directions_norm"])
            outputs = model(bundles)
            if wl["raw"]:
                outputs = RefRawPipeline.select_right_channel_per_pixel(stub, coords, outputs)
            _, total = lm.compute_loss(outputs, targets, coords, BASE_STEP)
            model.zero_grad(set_to_none=True)
            total.backward()
            if it >= warmup:
                times.append(time.perf_counter() - t0)
        n_rays = rays_per_mod * len(mods)
        ms = 1e3 * sum(times) / len(times)
        return n_rays / (ms / 1e3), ms, n_rays
    except Exception as e:      # the reference could not be driven here: fall back to the port and say so
        print(f"[bench] reference arm unavailable ({type(e).__name__}: {e}); using the oracle port", file=sys.stderr)
        return None


def cpu_arm(wl, steps, warmup, rays_per_mod):
    torch.set_num_threads(os.cpu_count() or 1)
    cores = os.cpu_count() or 1
    r = cpu_reference_arm(wl, steps, warmup, rays_per_mod)
    kind, what = "reference", "the unmodified reference (its BaseModel.forward + LossManager + autograd, imported from /root/reference)"
    if r is None:
        r = cpu_port_arm(wl, steps, warmup, rays_per_mod)
        kind, what = "port", "oracle port of the reference (oracle/mms_oracle.py)"
    v, ms, n_rays = r
    return {"value": v, "unit": "rays/s", "cores": cores, "kind": kind, "ms_per_step": ms,
            "sample": f"{n_rays} rays ({rays_per_mod}/modality) x {wl['n_c'] + wl['n_i']} samples of the same workload per step, {what}, "
                      f"torch CPU fp32, {cores} threads, {warmup} warm-up + {steps} timed steps"}


# ------------------------------------------------------------------------------------------------------------------
class HostPixelSampler:
    """The reference's per-step input path (cameras/pixel_samplers.py:71-89, data/dataloaders.py:164-167): CPU
    `torch.randint` draws of (camera, y, x) per modality and an advanced-index gather of the pixel values out of the
    cached frame stack in HOST memory, written to pinned buffers for the host -> device copy.  Frames are synthetic
    U(0,1) stacks of the real sensor geometry ([n_cam, H, W] raw mosaicked, [n_cam, H, W, C] demosaicked).
    `ranges` {mod: (a, b)}: the rows of the GLOBAL draw this rank owns (every rank draws the same global index set from
    the same seed and gathers only its slice)."""

    def __init__(self, modalities, global_counts, ranges, n_cam, raw, seed):
        from multimodalstudio_b200.pipelines import MODALITY_SENSORS
        self.mods, self.global_counts, self.ranges, self.n_cam = modalities, global_counts, ranges, n_cam
        self.gen = torch.Generator().manual_seed(seed)
        g = torch.Generator().manual_seed(seed + 7)
        self.frames, self.dims = {}, {}
        for m, c in modalities.items():
            w, h, _ = MODALITY_SENSORS[m]
            self.frames[m] = torch.empty((n_cam, h, w) if raw else (n_cam, h, w, c)).uniform_(0.0, 1.0, generator=g)
            self.dims[m] = (h, w)
        # two pinned slots: a step's host -> device copy may still be in flight while the next batch is drawn
        self.slots = [({m: torch.empty((b - a, 3), dtype=torch.int32).pin_memory() for m, (a, b) in ranges.items()},
                       {m: torch.empty((b - a, 1 if raw else modalities[m])).pin_memory() for m, (a, b) in ranges.items()})
                      for _ in range(2)]
        self.i = 0

    def sample(self):
        cs, ts = self.slots[self.i % 2]
        self.i += 1
        for m, n in self.global_counts.items():
            h, w = self.dims[m]
            a, b = self.ranges[m]
            cam = torch.randint(0, self.n_cam, (n,), generator=self.gen)[a:b]
            y = torch.randint(0, h, (n,), generator=self.gen)[a:b]
            x = torch.randint(0, w, (n,), generator=self.gen)[a:b]
            cs[m][:, 0], cs[m][:, 1], cs[m][:, 2] = cam, y, x
            ts[m].copy_(self.frames[m][cam, y, x].reshape(ts[m].shape))
        return cs, ts


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="sweep", choices=list(WORKLOADS))
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong: ONE global batch per step split over the ranks (loss = local sum / global count, summing "
                         "all-reduce); weak: every rank its own full batch (the reference's DDP semantics, gradient mean)")
    ap.add_argument("--rays", type=int, default=None, help="override the (global) ray count of a step")
    ap.add_argument("--micro-rays", type=int, default=None, help="override the rays per micro-batch")
    ap.add_argument("--all-heads", action="store_true", help="evaluate all M heads for every modality like the reference")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-side", action="store_true", help="skip the side figures (inference, device sampling, torch-on-GPU baseline)")
    ap.add_argument("--cpu-rays", type=int, default=128, help="rays per modality of the CPU baseline sample")
    ap.add_argument("--mlp-precision", type=int, default=None, choices=[0, 1, 2, 3],
                    help="layer arithmetic: 3 = tcgen05 3xTF32 (fp32-accurate), 2 = tcgen05 2-term fp16 split (fp32-accurate), "
                         "1 = tcgen05 single-pass TF32 (1e-2 band), 0 = fp32 SIMT")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from Python instead of replaying CUDA graphs")
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    if args.rays:
        wl["rays"] = wl["weak_rays"] = args.rays
    if args.micro_rays:
        wl["micro"] = args.micro_rays
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    hbm_peak, tf_burst, tf_sust, peak_src = peaks()
    strong = args.scaling == "strong"
    n_samp = wl["n_c"] + wl["n_i"]
    global_rays = wl["rays"] if strong else wl["weak_rays"] * world
    cfg_out = {"workload": f"{args.workload}: {len(wl['modalities'])} modalities ({'raw mosaicked' if wl['raw'] else 'demosaicked'}), "
                           f"global batch {global_rays} rays x {n_samp} samples (+{wl['bg']} background) per optimizer step, "
                           f"hash grids 16x2^19x2 fp32, pose refinement SO3xR3 shared",
               "global_rays_per_step": global_rays, "samples_per_ray": n_samp,
               "scaling_mode": ("strong: global batch split contiguously over the ranks, loss = local sum / global count, summing all-reduce"
                                if strong else "weak: every rank draws its own batch (reference DDP semantics), gradient mean"),
               "micro_batch_rays": wl["micro"], "parallelism": f"dp{world} (rays sharded, params replicated)",
               "l2": "inputs >> L2 (2x64 MiB tables + >20 GiB of activations per micro-batch)", "peaks": peak_src,
               "launch": "eager" if args.no_graph else "CUDA graphs (forward+backward per micro-batch | clip+AdamW)"}

    # the keys below describe the workload of BOTH arms (the driver compares the two lines' `config`)
    from multimodalstudio_b200 import ops as _ops
    from multimodalstudio_b200.pipelines import ShardPlan
    if args.mlp_precision is not None:
        _ops.set_mlp_precision(args.mlp_precision)
    FUSED = " + the SDF network's forward (layer 0 -> layer 1 -> sdf head) as ONE kernel with h0 kept on chip, " if _ops.SDF_FUSED else ""
    LAYERS = {0: "fp32 SIMT GEMM", 1: "tcgen05 single-pass TF32" + (FUSED + "single fp16 pass" if FUSED else "") + " (1e-2 band; not the reported configuration)",
              2: "tcgen05 2-term fp16 split (hi + lo, per-tensor power-of-two scaling) for the forward products of the CTA-pair shapes, 3xTF32 elsewhere; fp32 in / out",
              3: "tcgen05 3xTF32" + (FUSED + "2-term fp16 split with per-row power-of-two scales (three kind::f16 MMAs per product, fp32-accurate)" if FUSED else "") + ", fp32 in / out (1e-5 band)"}
    cfg_out["layers"] = LAYERS[_ops.MLP_PRECISION]
    cfg_out["micro_batches_per_rank"] = len(ShardPlan(split_rays(wl["rays"] if strong else wl["weak_rays"], wl["modalities"]),
                                                      world if strong else 1, rank if strong else 0, max_rays_per_micro=wl["micro"]))

    if args.impl == "reference":
        if rank != 0:
            return
        steps_run, warm_run = max(1, min(args.steps, 3)), max(0, min(args.warmup, 1))
        cpu = cpu_arm(wl, steps_run, warm_run, 256)
        line = {"metric": "train rays/sec (fwd+bwd)", "value": cpu["value"], "unit": "rays/s", "n_gpus": args.gpus, "steps": steps_run,
                "warmup": warm_run, "steps_requested": args.steps, "warmup_requested": args.warmup,
                "ms_per_step": cpu["ms_per_step"], "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": cfg_out, "impl": "reference", "cpu_baseline": cpu,
                "note": "CPU arm: a bounded sample per step (see cpu_baseline.sample), capped at 3 timed + 1 warm-up steps so that the run ends in minutes",
                "e2e": {"value": cpu["value"], "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch.distributed as dist
    from multimodalstudio_b200 import _lib
    from multimodalstudio_b200.models import MODALITY_CHANNELS
    from multimodalstudio_b200.pipelines import (MODALITY_SENSORS, DevicePixelSampler, RawPipeline, SyntheticScene)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=300))
    mods = {m: MODALITY_CHANNELS[m] for m in wl["modalities"]}
    n_cam = 50
    if strong:
        gcounts = split_rays(wl["rays"], wl["modalities"])
        plan = ShardPlan(gcounts, world, rank, max_rays_per_micro=wl["micro"])
        seed = 654824                                   # one global batch: the same draw on every rank, sliced by the plan
    else:
        gcounts = split_rays(wl["weak_rays"], wl["modalities"])
        plan = ShardPlan(gcounts, 1, 0, max_rays_per_micro=wl["micro"])
        seed = 654824 + rank                            # pixel_samplers.py:49-52: rank-offset seed
    local_counts = {m: b - a for m, (a, b) in plan.local.items()}
    local_rays = sum(local_counts.values())
    scene = SyntheticScene(mods, gcounts, raw=wl["raw"], seed=seed, n_cam=n_cam)
    pipe = RawPipeline(mods, scene.cameras, device=dev, raw=wl["raw"], render_all_heads=args.all_heads, preset=preset_of(wl),
                       num_samples=wl["n_c"], num_samples_importance=wl["n_i"], bg_samples=wl["bg"])
    n_batches = 4
    resident = []
    for _ in range(n_batches):
        cs, ts = scene.sample_batch()
        resident.append(({m: c.to(dev) for m, c in plan.local_slice(cs).items()}, {m: t.to(dev) for m, t in plan.local_slice(ts).items()}))
    graphed = not args.no_graph

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run_step(i, cs, ts):
        if strong or len(plan) > 1:
            return pipe.train_step_sharded(BASE_STEP + i, cs, ts, plan, graphed=graphed)
        if graphed:
            return pipe.train_step_graphed(BASE_STEP + i, cs, ts)
        return pipe.train_step(BASE_STEP + i, {m: c.to(dev, non_blocking=True) for m, c in cs.items()},
                               {m: t.to(dev, non_blocking=True) for m, t in ts.items()})

    def step_resident(i):
        cs, ts = resident[i % n_batches]
        return run_step(i, cs, ts)

    # at least 3 untimed steps: the first (micro-)batch runs eagerly, the second captures the CUDA graphs, then plain replays
    n_warm = max(args.warmup, 3)
    for i in range(n_warm):
        step_resident(i)
    barrier()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    prof_range = bool(os.environ.get("MMSB_PROFILER_RANGE"))      # `ncu --profile-from-start off`: capture the timed steps only
    if prof_range:
        torch.cuda.profiler.start()
    e0.record()
    for i in range(args.steps):
        step_resident(n_warm + i)
    e1.record()
    barrier()
    if prof_range:
        torch.cuda.profiler.stop()
    ms_total = e0.elapsed_time(e1)
    launches = _lib.launch_count() - l0
    if graphed:
        # every replay launches the kernels captured once; + the eager ray-generation / collider pre-pass and optimizer
        launches += pipe.graph_launches * len(plan) * args.steps
    clk = clocks.stop() if rank == 0 else None
    t = torch.tensor([ms_total], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = global_rays / (ms_step / 1e3)

    # ---- end to end: the reference's input path inside the timed region (CPU index draw + gather from the host frame
    # stack -> pinned memory -> device), the step, and the loss read back to the host
    e2e = None
    if not args.no_e2e:
        host = HostPixelSampler(mods, gcounts, plan.local, n_cam, wl["raw"], seed)

        def run_e2e(n_steps):
            """n_steps optimizer steps, every one of them with its own host-side pixel draw + gather, pinned host ->
            device copy and a device -> host read of its loss.  The draw for step i + 1 is made while the GPU works on
            step i (two pinned slots), the way a dataloader worker would; all n_steps draws are inside the timed region."""
            losses = []
            cs, ts = host.sample()
            for i in range(n_steps):
                _, total = run_step(i, cs, ts)              # asynchronous: H2D copies + graph replays are enqueued
                if i + 1 < n_steps:
                    cs, ts = host.sample()                  # next batch on the CPU, overlapped with this step on the GPU
                losses.append(float(total.item()))          # device -> host read of the step's loss
            return losses

        run_e2e(2)
        barrier()
        e0.record()
        run_e2e(args.steps)
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t.item()) / args.steps
        cs0, ts0 = host.slots[0]
        h2d = sum(c.numel() * c.element_size() for c in cs0.values()) + sum(x.numel() * x.element_size() for x in ts0.values())
        e2e = {"value": global_rays / (ms_e2e / 1e3), "unit": "rays/s", "ms_per_step": ms_e2e,
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
               "input_path": "CPU torch.randint pixel draw + gather from the host frame stack (pixel_samplers.py:71-89) -> pinned -> device, per rank and step; the draw of step i+1 overlaps step i on the GPU"}
        del host

    # ---- per-kernel durations for the roofline: one more micro-batch, eager, with CUDA events around every launch of
    # the instrumented entry points (same stream), outside the timed region so it is not perturbed
    roof, roof_hash = None, None
    cs, ts = resident[0]
    cs0, ts0 = plan.slice(0, cs, "local"), plan.slice(0, ts, "local")
    scales0 = plan.loss_scales(0)
    count0 = pipe.count_unmasked_samples(cs)
    for rep in range(2):                                     # the first pass warms the eager allocator
        if rank == 0 and rep == 1:
            _lib.start_kernel_timing(None)
        pipe.forward_backward(cs0, ts0, BASE_STEP, scales0, count0, accumulate=False)
        torch.cuda.synchronize()
    if rank == 0:
        rec = _lib.stop_kernel_timing()
        nmb = len(plan)
        # entry point -> (product class, index of n in the argument list; in_dim and out_dim follow it)
        LAYER = {"mmsb_linear_fwd": ("fwd", 6), "mmsb_linear_fwd_tc": ("fwd", 6), "mmsb_linear_fwd_head_tc": ("fwd", 6),
                 "mmsb_linear_bwd_data": ("dgrad", 9), "mmsb_linear_bwd_data_tc": ("dgrad", 9),
                 "mmsb_linear_bwd_data_rank1_tc": ("dgrad", 9), "mmsb_linear_bwd_data_head_tc": ("dgrad", 13),
                 "mmsb_linear_bwd_weight": ("wgrad", 6), "mmsb_linear_bwd_weight_tc": ("wgrad", 6),
                 "mmsb_linear_bwd_weight_head_tc": ("wgrad", 11)}
        KERNEL = {"fwd": "tc_rows*_kernel<FWD> (mmsb_linear_fwd[_head]_tc)", "dgrad": "tc_rows*_kernel<DGRAD> (mmsb_linear_bwd_data[_head|_rank1]_tc)",
                  "wgrad": "tc_wgrad_kernel (mmsb_linear_bwd_weight[_head]_tc)"}
        tc_ms = 0.0
        if os.environ.get("MMSB_BENCH_TABLE"):
            # per entry point and layer shape: calls, total ms of the instrumented micro-batch (dev aid)
            agg = {}
            for name, ms, a in rec:
                key = name
                if name in LAYER:
                    j = LAYER[name][1]
                    key = f"{name} n={a[j].value} k={a[j + 1].value} o={a[j + 2].value}"
                c = agg.setdefault(key, [0, 0.0])
                c[0] += 1
                c[1] += ms
            with open(os.environ["MMSB_BENCH_TABLE"], "w") as fh:
                tot = sum(v[1] for v in agg.values())
                fh.write(f"instrumented entry points: {tot:.2f} ms per micro-batch ({nmb} micro-batches per step)\n")
                for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                    fh.write(f"{v[1]:9.3f} ms {v[0]:5d} calls  {k}\n")
        flops = {"fwd": 0.0, "dgrad": 0.0, "wgrad": 0.0}
        msk = {k: 0.0 for k in flops}
        cnt = {k: 0 for k in flops}
        hb = {"mmsb_hashgrid_fwd": [0.0, 0.0, 0], "mmsb_hashgrid_bwd": [0.0, 0.0, 0]}
        shapes = {}     # (class, rows, in, out) -> [flops, ms, launches] of the instrumented micro-batch
        fused = [0.0, 0.0, 0, 0.0, 0]       # flops, ms, launches of the fused SDF forward; ms and rows of its largest launch
        for name, ms, a in rec:
            if name == "mmsb_sdf_net_fwd_fused":
                n, k, o = a[2].value, a[3].value, a[4].value
                f_ = 2.0 * n * (k * o + o * o + o)
                flops["fwd"] += f_
                msk["fwd"] += ms
                cnt["fwd"] += 1
                tc_ms += ms
                fused[0] += f_
                fused[1] += ms
                fused[2] += 1
                if n > fused[4]:
                    fused[3], fused[4] = ms, n
            elif name in LAYER:
                cls, j = LAYER[name]
                n, k, o = a[j].value, a[j + 1].value, a[j + 2].value
                flops[cls] += 2.0 * n * k * o
                msk[cls] += ms
                cnt[cls] += 1
                sh = shapes.setdefault((cls, n, k, o), [0.0, 0.0, 0])
                sh[0] += 2.0 * n * k * o
                sh[1] += ms
                sh[2] += 1
                if name.endswith("_tc"):
                    tc_ms += ms
            elif name in hb:
                n = a[-2].value
                hb[name][0] += n * (1024.0 if name.endswith("fwd") else 2048.0)      # L*8*F*4 B per look-up (x2 read-modify-write)
                hb[name][1] += ms
                hb[name][2] += 1
        top = max(msk, key=lambda k: msk[k])
        if msk[top] > 0:
            ach = flops[top] / (msk[top] / 1e3) / 1e12
            path = {0: "fp32 SIMT GEMM", 1: "tcgen05 TF32", 2: "tcgen05 2-term fp16 split (3 kind::f16 MMAs per product, fp32-accurate)",
                    3: "tcgen05 3xTF32 (3 MMAs per product, fp32-accurate)"}[_ops.MLP_PRECISION]
            traffic, traffic_note, traffic_src = None, None, None
            for tname in ("r2_traffic.json", "r1f_traffic.json"):
                tp = os.path.join(ROOT, "profiles", tname)
                if os.path.exists(tp):
                    tj = json.load(open(tp))
                    if top in tj and tj.get("_precision", 3) == _ops.MLP_PRECISION:
                        traffic, traffic_note = tj[top]["dram_bytes_per_launch"], tj[top]["note"]
                        traffic_src = f"profiles/{tname} (STATIC: ncu --set full capture of an earlier run, not measured in this run)"
                        break
            # every fp32-accurate product costs three MMAs; kind::f16 runs at the bf16 rate, kind::tf32 at half of it
            ceil_div_ = {3: 6.0, 2: 3.0, 1: 2.0, 0: None}[_ops.MLP_PRECISION]
            roof = {"kernel": f"{KERNEL[top]} ({path}; {100.0 * tc_ms / max(sum(msk.values()), 1e-9):.0f}% of layer time on tcgen05)",
                    "bound": "tensor", "achieved": ach, "peak": tf_sust, "unit": "TFLOP/s",
                    "frac": ach / tf_sust, "traffic": traffic, "traffic_note": traffic_note, "traffic_source": traffic_src,
                    "launches": cnt[top] * nmb, "ms_per_step": msk[top] * nmb,
                    "peak_source": f"{peak_src}: bf16 sustained (kernel timed inside a long step)",
                    "achieved_is": "2*n*in*out algorithmic FLOP of every launch of this product class in one micro-batch / their summed CUDA-event time (all layer shapes, incl. the narrow HBM-bound ones)",
                    "per_class": {c: {"tflops": flops[c] / max(msk[c], 1e-9) / 1e9, "ms_per_step": msk[c] * nmb, "launches": cnt[c] * nmb} for c in flops}}
            if ceil_div_:
                roof["ceiling_fp32_accurate"] = tf_sust / ceil_div_
                roof["frac_of_fp32_accurate_ceiling"] = ach / (tf_sust / ceil_div_)
            # the same figure for the single most expensive layer shape of every class: what the kernel reaches where it
            # is tensor-bound, measured live like the class averages
            big = {}
            for c in flops:
                cand = [(v[1], key, v) for key, v in shapes.items() if key[0] == c]
                if cand:
                    _, key, v = max(cand)
                    tf = v[0] / max(v[1], 1e-9) / 1e9
                    big[c] = {"rows": key[1], "in": key[2], "out": key[3], "launches_per_micro_batch": v[2], "ms_per_launch": v[1] / v[2],
                              "tflops": tf, "frac_of_bf16_peak": tf / tf_sust}
            roof["largest_shape"] = big
            if fused[2]:
                tf = fused[0] / max(fused[1], 1e-9) / 1e9
                tf_big = 2.0 * fused[4] * (71 * 256 + 256 * 256 + 256) / max(fused[3], 1e-9) / 1e9
                roof["sdf_fused_fwd"] = {"kernel": "sdf_fused_fwd_kernel (mmsb_sdf_net_fwd_fused): x -> h0 -> h1 -> sdf in one launch, h0 on chip",
                                         "tflops": tf, "frac_of_bf16_peak": tf / tf_sust, "launches_per_micro_batch": fused[2],
                                         "ms_per_micro_batch": fused[1], "largest_launch": {"rows": fused[4], "ms": fused[3], "tflops": tf_big,
                                                                                             "frac_of_bf16_peak": tf_big / tf_sust},
                                         "arithmetic": "three kind::f16 MMAs per product" if _ops.MLP_PRECISION != 1 else "one kind::f16 MMA per product",
                                         "tensor_pipe_active_ncu": "47 % (activations not stored) / 27 % (h0 + h1 stored for the backward): profiles/r2b_ncu_full.md, profiles/r2c_ncu_full.md"}
        roof_hash = {}
        for hk, tag in (("mmsb_hashgrid_fwd", "fwd"), ("mmsb_hashgrid_bwd", "bwd")):
            if hb[hk][1] > 0:
                ach = hb[hk][0] / (hb[hk][1] / 1e3) / 1e9
                roof_hash[tag] = {"kernel": hk, "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                                  "traffic": None, "launches": hb[hk][2] * nmb, "ms_per_step": hb[hk][1] * nmb,
                                  "achieved_is": "1024 B (fwd) / 2048 B (bwd, read-modify-write) per look-up x look-ups / summed CUDA-event time"}

    # ---- side figures (not part of the headline) ---------------------------------------------------------------------
    inference, dev_sampling, torch_gpu = None, None, None

    def side_inference():
        # full-frame inference: eval-mode rendering in chunks
        big = {m: resident[0][0][m] for m in mods}
        n_inf = sum(c.shape[0] for c in big.values())
        pipe.render(big)
        torch.cuda.synchronize()
        e0.record()
        pipe.render(big)
        e1.record()
        torch.cuda.synchronize()
        ms_inf = e0.elapsed_time(e1)
        return {"value": n_inf / (ms_inf / 1e3), "unit": "rays/s", "rays": n_inf, "ms": ms_inf,
                "what": "RawPipeline.render: eval mode, no_grad, all modalities per chunk as one batch (eager launches)"}

    def side_device_sampling():
        # pixel draw + target gather on the device from a frame stack resident in HBM (no per-step
        # host -> device copy at all; not an end-to-end number in the contract's sense, hence a side figure)
        frames = {}
        gfr = torch.Generator(device=dev).manual_seed(5)
        for m, c in mods.items():
            w, h, _ = MODALITY_SENSORS[m]
            frames[m] = torch.empty((n_cam, h, w, 1 if wl["raw"] else c), device=dev).uniform_(0.0, 1.0, generator=gfr)
        dps = DevicePixelSampler(frames, local_counts, seed=seed, rank=rank)

        def step_dev(i):
            cs_, ts_ = dps.sample(BASE_STEP + i)
            return run_step(i, cs_, ts_)

        for i in range(2):
            step_dev(i)
        torch.cuda.synchronize()
        k_ = max(2, args.steps // 2)
        e0.record()
        for i in range(k_):
            step_dev(i)
        e1.record()
        torch.cuda.synchronize()
        ms_ds = e0.elapsed_time(e1) / k_
        return {"value": global_rays / (ms_ds / 1e3), "unit": "rays/s", "ms_per_step": ms_ds,
                "frames_resident_bytes": sum(f.numel() * 4 for f in frames.values()),
                "what": "mmsb_sample_pixels (Philox draw + target gather from the HBM-resident frame stack) inside the step"}

    if world == 1 and not args.no_e2e and not args.no_side:
        for name, fn in (("inference", side_inference), ("dev_sampling", side_device_sampling)):
            try:
                r = fn()
            except Exception as e:                 # a side figure must never take the headline line down with it
                r = {"unavailable": f"{type(e).__name__}: {e}"[:300]}
            if name == "inference":
                inference = r
            else:
                dev_sampling = r
            torch.cuda.empty_cache()
    # ---- fast mode, a side line only (never the headline): the same step with single-pass TF32 layer products (the
    # precision class of the reference's own fp16-autocast GPU runs, engine/trainer.py:51-62; 1e-2 band, tests:
    # test_mlp_precision_modes_agree / test_tc_layer_products_vs_fp64[1-...]), still one kernel per layer
    fast_mode = None
    if world == 1 and not args.no_e2e and not args.no_side and _ops.MLP_PRECISION == 3:
        try:
            pipe = None
            torch.cuda.empty_cache()
            _ops.set_mlp_precision(1)
            pipe = RawPipeline(mods, scene.cameras, device=dev, raw=wl["raw"], render_all_heads=args.all_heads, preset=preset_of(wl),
                               num_samples=wl["n_c"], num_samples_importance=wl["n_i"], bg_samples=wl["bg"])
            for i in range(3):
                step_resident(i)
            torch.cuda.synchronize()
            k_ = max(2, args.steps // 2)
            e0.record()
            for i in range(k_):
                step_resident(3 + i)
            e1.record()
            torch.cuda.synchronize()
            ms_f = e0.elapsed_time(e1) / k_
            fast_mode = {"value": global_rays / (ms_f / 1e3), "unit": "rays/s", "ms_per_step": ms_f, "layers": LAYERS[1],
                         "error_band": "1e-2 relative (single-pass TF32 products: 5e-3 measured per layer against fp64; fused single-pass fp16 SDF forward: 5e-4)",
                         "note": "side line, NOT the headline: per-layer kernels become HBM-bound at this arithmetic cost"}
        except Exception as e:
            fast_mode = {"unavailable": f"{type(e).__name__}: {e}"[:300]}
        finally:
            _ops.set_mlp_precision(3)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        del resident
        pipe = None
        torch.cuda.empty_cache()
        if not args.no_side:
            try:
                v, ms, n_rays = cpu_port_arm(wl, 2, 1, 64, device="cuda")
                torch_gpu = {"value": v, "unit": "rays/s", "ms_per_step": ms, "rays": n_rays,
                             "what": "the reference's torch path (oracle/mms_oracle.py: cuBLAS + ATen kernels, fp32, per-modality loop, dense "
                                     "index_put hash-grid backward) on this B200, same workload, bounded sample — BASELINE.md §3b's GPU bar"}
            except Exception as e:
                torch_gpu = {"unavailable": f"{type(e).__name__}: {e}"[:200]}
            torch.cuda.empty_cache()
        cpu = cpu_arm(wl, 2, 1, args.cpu_rays)
    if rank == 0:
        line = {"metric": "train rays/sec (fwd+bwd)", "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps,
                "warmup": n_warm, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong" if strong else "weak",
                "vs_baseline": None, "dtype": {0: "f32", 1: "tf32", 2: "f32", 3: "f32"}[_ops.MLP_PRECISION], "data": "synthetic",
                "config": cfg_out, "clocks": clk, "e2e": e2e, "gpu_launches": launches,
                "roofline": roof, "roofline_hashgrid": roof_hash, "cpu_baseline": cpu, "inference": inference,
                "e2e_device_sampling": dev_sampling, "fast_mode": fast_mode, "torch_gpu_baseline": torch_gpu, "impl": "b200"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
