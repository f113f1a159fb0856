#!/usr/bin/env python
"""bench.py — train rays/s (forward + backward + clip + AdamW) of the MMS-FW per-ray rendering hot path.

    python bench.py --gpus N --steps K --warmup W [--workload grid|grid_raw|sweep] [--impl reference]

One "step" = one pass of the hot path over one synthetic batch: ray generation -> NeuS sampling -> hash grids +
MLPs -> compositing -> (mosaick-aware) losses -> backward -> global-norm clip -> AdamW (what the reference's
`train_step` times, engine/trainer.py:107-114).  Prints ONE JSON line (see the driver contract).
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

WORKLOADS = {
    # BASELINE.json configs[1]: confs/grid.yaml, RGB + 1 extra modality, demosaicked, 4096 rays x 128 samples
    "grid": dict(modalities=["rgb", "infrared"], rays=4096, n_c=64, n_i=64, bg=16, raw=False),
    # configs[2]: confs/grid_raw.yaml, 5 modalities, mosaick-aware loss, 8192 rays (ragged split)
    "grid_raw": dict(modalities=["rgb", "infrared", "mono", "polarization", "multispectral"], rays=8192, n_c=32, n_i=32, bg=16, raw=True),
    # configs[4]: synthetic large-batch sweep, 5 modalities, 65536 rays x 256 samples
    "sweep": dict(modalities=["rgb", "infrared", "mono", "polarization", "multispectral"], rays=65536, n_c=128, n_i=128, bg=16, raw=True),
}
BASE_STEP = 60000       # late in the 100k-iteration schedule: all 16 levels active, delta = 2/1024, anneal = 1


def split_rays(total, mods):
    base, rem = divmod(total, len(mods))
    return {m: base + (1 if i < rem else 0) for i, m in enumerate(mods)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


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


def cpu_reference_arm(wl, steps, warmup, rays_per_mod=24):
    """The reference's algorithm on the host CPU: the oracle port (the reference is Python and cannot travel to the
    GPU box; oracle/mms_oracle.py is bit-exact against it on the build container).  One step =
    forward + channel select + losses + backward over a BOUNDED sample of the workload."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import mms_oracle as O
    from multimodalstudio_b200.models import MODALITY_CHANNELS, MOSAICK_PATTERNS, build_model
    from multimodalstudio_b200.pipelines import SyntheticScene
    torch.set_num_threads(os.cpu_count() or 1)
    mods = {m: MODALITY_CHANNELS[m] for m in wl["modalities"]}
    model = build_model("grid_raw", modalities=mods, num_samples=wl["n_c"], num_samples_importance=wl["n_i"], bg_samples=wl["bg"])
    sd = {k: v.detach().requires_grad_(True) for k, v in model.state_dict().items()}
    cfg = O.default_cfg(modalities=mods, num_samples=wl["n_c"], num_samples_importance=wl["n_i"], bg_samples=wl["bg"])
    orc = O.GridModelOracle(sd, cfg)
    scene = SyntheticScene(mods, {m: rays_per_mod for m in mods}, raw=wl["raw"])
    times = []
    for it in range(warmup + steps):
        coords, targets = scene.sample_batch()
        t0 = time.perf_counter()
        outputs = {}
        for mod in mods:
            cam = scene.cameras[mod]
            r = O.raygen(coords[mod], cam.camera_to_worlds, cam.intrinsics, cam.distortion_params, None)
            n_hit = int(O.sphere_collide(r["origins"], r["directions"])[2].sum())
            rand = {"uniform": torch.rand(n_hit, 1), "pdf": [torch.rand(n_hit, 1) for _ in range(4)],
                    "background": torch.rand(rays_per_mod, wl["bg"] + 1)}
            outputs[mod] = orc.forward_modality(mod, r["origins"], r["directions"], r["up_directions"], rand)
        _, total = orc.loss(outputs, targets, coords, MOSAICK_PATTERNS if wl["raw"] else None, 5e-4 * 0.0157)
        for v in sd.values():
            v.grad = None
        total.backward()
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    n_rays = rays_per_mod * len(mods)
    ms = 1e3 * sum(times) / len(times)
    return n_rays / (ms / 1e3), ms, n_rays


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="grid", choices=list(WORKLOADS))
    ap.add_argument("--rays", type=int, default=None, help="override the total ray count")
    ap.add_argument("--all-heads", action="store_true", help="evaluate all M heads for every modality like the reference")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--mlp-precision", type=int, default=None, choices=[0, 1, 3],
                    help="layer arithmetic: 3 = tcgen05 3xTF32 (fp32-accurate, default, the reported configuration), "
                         "1 = tcgen05 single-pass TF32 (1e-2 band), 0 = fp32 SIMT")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from Python instead of replaying CUDA graphs")
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    if args.rays:
        wl["rays"] = args.rays
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    hbm_peak, tf_burst, tf_sust, peak_src = peaks()
    cfg_out = {"workload": f"{args.workload}: {len(wl['modalities'])} modalities ({'raw mosaicked' if wl['raw'] else 'demosaicked'}), "
                           f"{wl['rays']} rays/GPU x {wl['n_c'] + wl['n_i']} samples (+{wl['bg']} background), hash grids 16x2^19x2 fp32, "
                           f"pose refinement SO3xR3 shared",
               "rays_per_gpu": wl["rays"], "samples_per_ray": wl["n_c"] + wl["n_i"], "parallelism": f"dp{world} (rays sharded, params replicated)",
               "l2": "inputs >> L2 (2x64 MiB tables + >1 GiB of activations per step)", "peaks": peak_src,
               "launch": "eager" if args.no_graph else "CUDA graphs (forward+backward | clip+AdamW)"}

    if args.impl == "reference":
        if rank != 0:
            return
        cores = os.cpu_count() or 1
        v, ms, n_rays = cpu_reference_arm(wl, max(1, min(args.steps, 3)), max(0, min(args.warmup, 1)))
        line = {"metric": "train rays/sec (fwd+bwd)", "value": v, "unit": "rays/s", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": cfg_out, "impl": "reference",
                "cpu_baseline": {"value": v, "unit": "rays/s", "cores": cores, "kind": "port",
                                 "sample": f"{n_rays} rays of the same workload per step ({n_rays // len(wl['modalities'])}/modality), oracle port of the reference on torch CPU, {cores} threads"},
                "e2e": {"value": v, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch.distributed as dist
    from multimodalstudio_b200 import _lib
    from multimodalstudio_b200.models import MODALITY_CHANNELS
    from multimodalstudio_b200.pipelines import RawPipeline, SyntheticScene
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    from multimodalstudio_b200 import ops as _ops0
    if args.mlp_precision is not None:
        _ops0.set_mlp_precision(args.mlp_precision)
    cfg_out["layers"] = {0: "fp32 SIMT GEMM", 1: "tcgen05 single-pass TF32 (1e-2 band; not the reported configuration)",
                         3: "tcgen05 3xTF32, fp32 in / out (1e-5 band)"}[_ops0.MLP_PRECISION]
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
    mods = {m: MODALITY_CHANNELS[m] for m in wl["modalities"]}
    rays = split_rays(wl["rays"], wl["modalities"])
    scene = SyntheticScene(mods, rays, raw=wl["raw"], seed=654824 + rank)      # pixel_samplers.py:49-52 rank-offset seed
    pipe = RawPipeline(mods, scene.cameras, device=dev, raw=wl["raw"], render_all_heads=args.all_heads,
                       num_samples=wl["n_c"], num_samples_importance=wl["n_i"], bg_samples=wl["bg"])
    n_batches = 4
    host = [scene.sample_batch() for _ in range(n_batches)]
    pinned = [({m: c.pin_memory() for m, c in cs.items()}, {m: t.pin_memory() for m, t in ts.items()}) for cs, ts in host]
    resident = [({m: c.to(dev) for m, c in cs.items()}, {m: t.to(dev) for m, t in ts.items()}) for cs, ts in host]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident(i):
        cs, ts = resident[i % n_batches]
        if args.no_graph:
            return pipe.train_step(BASE_STEP + i, cs, ts)
        return pipe.train_step_graphed(BASE_STEP + i, cs, ts)

    def step_e2e(i):
        cs, ts = pinned[i % n_batches]
        if args.no_graph:
            csd = {m: c.to(dev, non_blocking=True) for m, c in cs.items()}
            tsd = {m: t.to(dev, non_blocking=True) for m, t in ts.items()}
            _, total = pipe.train_step(BASE_STEP + i, csd, tsd)
        else:
            _, total = pipe.train_step_graphed(BASE_STEP + i, cs, ts)     # pinned host -> static graph inputs inside
        return float(total.item())                      # device -> host read of the step's loss

    # at least 3 untimed steps: the first runs eagerly, the second captures the CUDA graphs, the third is a plain replay
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
    e0.record()
    for i in range(args.steps):
        step_resident(n_warm + i)
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = _lib.launch_count() - l0
    if not args.no_graph:
        launches = pipe.graph_launches * args.steps      # every replay launches the kernels captured once
    clk = clocks.stop() if rank == 0 else None
    t = torch.tensor([ms_total], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = wl["rays"] * world / (ms_step / 1e3)

    e2e = None
    if not args.no_e2e:
        for i in range(2):
            step_e2e(i)
        barrier()
        e0.record()
        for i in range(args.steps):
            step_e2e(i)
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t.item()) / args.steps
        h2d = sum(c.numel() * c.element_size() for c in pinned[0][0].values()) + sum(x.numel() * x.element_size() for x in pinned[0][1].values())
        e2e = {"value": wl["rays"] * world / (ms_e2e / 1e3), "unit": "rays/s", "ms_per_step": ms_e2e,
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4}

    # per-kernel durations for the roofline: two more steps with CUDA events around every launch of the
    # instrumented entry points (same stream), outside the timed region so it is not perturbed
    roof, roof_hash = None, None
    if rank == 0:
        _lib.start_kernel_timing(None)
    for i in range(2):                                      # every rank steps (the gradient all-reduce is collective)
        cs, ts = resident[i % n_batches]
        pipe.train_step(BASE_STEP + i, cs, ts)              # eager: events cannot be recorded inside a graph replay
    torch.cuda.synchronize()
    if rank == 0:
        rec = _lib.stop_kernel_timing()
        # entry point -> (product class, index of n in the argument list; in_dim and out_dim follow it)
        LAYER = {"mmsb_linear_fwd": ("fwd", 6), "mmsb_linear_fwd_tc": ("fwd", 6), "mmsb_linear_fwd_head_tc": ("fwd", 6),
                 "mmsb_linear_bwd_data": ("dgrad", 9), "mmsb_linear_bwd_data_tc": ("dgrad", 9),
                 "mmsb_linear_bwd_data_rank1_tc": ("dgrad", 9), "mmsb_linear_bwd_data_head_tc": ("dgrad", 13),
                 "mmsb_linear_bwd_weight": ("wgrad", 6), "mmsb_linear_bwd_weight_tc": ("wgrad", 6),
                 "mmsb_linear_bwd_weight_head_tc": ("wgrad", 11)}
        KERNEL = {"fwd": "tc_rows_kernel<.,FWD> (mmsb_linear_fwd[_head]_tc)", "dgrad": "tc_rows_kernel<.,DGRAD> (mmsb_linear_bwd_data[_head|_rank1]_tc)",
                  "wgrad": "tc_wgrad_kernel (mmsb_linear_bwd_weight[_head]_tc)"}
        tc_ms = 0.0
        if os.environ.get("MMSB_BENCH_TABLE"):
            # per entry point and layer shape: calls, total ms over the two instrumented steps (dev aid)
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
                fh.write(f"instrumented entry points: {tot / 2:.2f} ms per step\n")
                for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                    fh.write(f"{v[1] / 2:9.3f} ms/step {v[0] // 2:5d} calls  {k}\n")
        flops = {"fwd": 0.0, "dgrad": 0.0, "wgrad": 0.0}
        msk = {k: 0.0 for k in flops}
        cnt = {k: 0 for k in flops}
        hb = {"mmsb_hashgrid_fwd": [0.0, 0.0, 0], "mmsb_hashgrid_bwd": [0.0, 0.0, 0]}
        shapes = {}     # (class, rows, in, out) -> [flops, ms, launches] over the two instrumented steps
        for name, ms, a in rec:
            if name in LAYER:
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
            from multimodalstudio_b200 import ops as _ops
            path = {0: "fp32 SIMT GEMM", 1: "tcgen05 TF32", 3: "tcgen05 3xTF32 (3 MMAs per product, fp32-accurate)"}[_ops.MLP_PRECISION]
            traffic, traffic_note = None, None
            tp = os.path.join(ROOT, "profiles", "r1f_traffic.json")
            if os.path.exists(tp):
                tj = json.load(open(tp))
                if top in tj:
                    traffic, traffic_note = tj[top]["dram_bytes_per_launch"], tj[top]["note"]
            roof = {"kernel": f"{KERNEL[top]} ({path}; {100.0 * tc_ms / max(sum(msk.values()), 1e-9):.0f}% of layer time on tcgen05)",
                    "bound": "tensor", "achieved": ach, "peak": tf_sust, "unit": "TFLOP/s",
                    "frac": ach / tf_sust, "traffic": traffic, "traffic_note": traffic_note, "launches": cnt[top] // 2,
                    "ms_per_step": msk[top] / 2,
                    "peak_source": f"{peak_src} bf16 sustained (kernel timed inside a long step)",
                    "achieved_is": "2*n*in*out algorithmic FLOP of every launch of this product class in one step / their summed CUDA-event time (all layer shapes, incl. the narrow HBM-bound ones)",
                    # fp32-accurate products cost three TF32 MMAs each and TF32 runs at half the bf16 rate:
                    "ceiling_3xtf32": tf_sust / 6.0, "frac_of_3xtf32_ceiling": ach / (tf_sust / 6.0),
                    "per_class": {c: {"tflops": flops[c] / max(msk[c], 1e-9) / 1e9, "ms_per_step": msk[c] / 2, "launches": cnt[c] // 2} for c in flops}}
            # the same figure for the single most expensive layer shape of every class (a 256 -> 256 layer of the SDF batch):
            # what the kernel reaches where it is tensor-bound, measured live like the class averages
            big = {}
            for c in flops:
                cand = [(v[1], key, v) for key, v in shapes.items() if key[0] == c]
                if cand:
                    _, key, v = max(cand)
                    tf = v[0] / max(v[1], 1e-9) / 1e9
                    big[c] = {"rows": key[1], "in": key[2], "out": key[3], "launches_per_step": v[2] // 2, "ms_per_launch": v[1] / v[2],
                              "tflops": tf, "frac_of_bf16_peak": tf / tf_sust, "frac_of_3xtf32_ceiling": tf / (tf_sust / 6.0)}
            roof["largest_shape"] = big
        hk = max(hb, key=lambda k: hb[k][1])
        if hb[hk][1] > 0:
            ach = hb[hk][0] / (hb[hk][1] / 1e3) / 1e9
            roof_hash = {"kernel": hk, "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                         "traffic": None, "launches": hb[hk][2] // 2, "ms_per_step": hb[hk][1] / 2}

    # full-frame inference: eval-mode rendering of 4 x the training batch in chunks, reported beside
    # the headline metric (not part of it)
    inference = None
    if world == 1 and not args.no_e2e:
        big = {m: torch.cat([resident[i][0][m] for i in range(n_batches)], 0) for m in mods}
        n_inf = sum(c.shape[0] for c in big.values())
        pipe.render(big)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(3):
            pipe.render(big)
        e1.record()
        torch.cuda.synchronize()
        ms_inf = e0.elapsed_time(e1) / 3
        inference = {"value": n_inf / (ms_inf / 1e3), "unit": "rays/s", "rays": n_inf, "ms": ms_inf,
                     "what": "RawPipeline.render: eval mode, no_grad, all modalities per chunk as one batch (eager launches)"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, ms, n_rays = cpu_reference_arm(wl, 2, 1)
        cores = os.cpu_count() or 1
        cpu = {"value": v, "unit": "rays/s", "cores": cores, "kind": "port", "ms_per_step": ms,
               "sample": f"{n_rays} rays of the same workload per step, oracle port of the reference (torch CPU fp32, {cores} threads), 1 warm-up + 2 timed steps"}
    if rank == 0:
        line = {"metric": "train rays/sec (fwd+bwd)", "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps,
                "warmup": n_warm, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": {0: "f32", 1: "tf32", 3: "f32"}[_ops.MLP_PRECISION] if roof else "f32", "data": "synthetic", "config": cfg_out, "clocks": clk, "e2e": e2e, "gpu_launches": launches,
                "roofline": roof, "roofline_hashgrid": roof_hash, "cpu_baseline": cpu, "inference": inference, "impl": "b200"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
