"""dev: kernel-level timeline of the graphed train step with torch.profiler (CUPTI activity records, no replay):
every kernel of the step including torch's glue kernels, summed over `--steps` replays.

    python scripts/profile_step.py --workload sweep --steps 3 --out gpurun_out/kernels.txt

`sweep` profiles ONE micro-batch per step (8195 rays x 256 samples: the per-GPU share of the 65560-ray global batch at 8
GPUs, and what a single GPU repeats 8 times per optimizer step).
"""
import argparse, collections, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from multimodalstudio_b200.models import MODALITY_CHANNELS
from multimodalstudio_b200.pipelines import RawPipeline, ShardPlan, SyntheticScene

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="sweep")
ap.add_argument("--rays", type=int, default=None)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--out", default="gpurun_out/kernels.txt")
ap.add_argument("--eager", action="store_true")
args = ap.parse_args()
wl = dict(bench.WORKLOADS[args.workload])
if args.workload == "sweep":
    wl["rays"] = wl["micro"]
if args.rays:
    wl["rays"] = args.rays
dev = torch.device("cuda", 0)
mods = {m: MODALITY_CHANNELS[m] for m in wl["modalities"]}
scene = SyntheticScene(mods, bench.split_rays(wl["rays"], wl["modalities"]), raw=wl["raw"])
pipe = RawPipeline(mods, scene.cameras, device=dev, raw=wl["raw"], render_all_heads=False, num_samples=wl["n_c"],
                   num_samples_importance=wl["n_i"], bg_samples=wl["bg"])
batches = [tuple({m: t.to(dev) for m, t in d.items()} for d in scene.sample_batch()) for _ in range(2)]
plan = ShardPlan(bench.split_rays(wl["rays"], wl["modalities"]))
step = lambda i, cs, ts: pipe.train_step_sharded(i, cs, ts, plan, graphed=not args.eager)
for i in range(3):
    step(bench.BASE_STEP + i, *batches[i % 2])
torch.cuda.synchronize()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    for i in range(args.steps):
        step(bench.BASE_STEP + 3 + i, *batches[i % 2])
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        a = agg[ev.name]
        a[0] += 1
        a[1] += ev.device_time_total if hasattr(ev, "device_time_total") else ev.cuda_time_total
tot = sum(v[1] for v in agg.values())
with open(args.out, "w") as fh:
    fh.write(f"{args.workload} {wl['rays']} rays, {args.steps} steps, {'eager' if args.eager else 'CUDA graphs'}: "
             f"{tot / args.steps / 1e3:.2f} ms of kernels per step, {sum(v[0] for v in agg.values()) // args.steps} launches per step\n")
    mine = sum(v[1] for k, v in agg.items() if "mmsb" in k)
    fh.write(f"libmms_b200 kernels: {mine / args.steps / 1e3:.2f} ms per step ({100 * mine / tot:.1f}%)\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        fh.write(f"{v[1] / args.steps / 1e3:9.3f} ms/step {v[0] // args.steps:6d} launches {100 * v[1] / tot:5.1f}%  {k[:150]}\n")
print(open(args.out).read()[:3000])
