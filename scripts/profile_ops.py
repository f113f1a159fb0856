"""dev: which torch (aten) ops of the eager train step cost GPU time, with input shapes and the Python frames that
issued them — the map for removing glue kernels.

    python scripts/profile_ops.py --workload grid_raw --out gpurun_out/ops.txt
"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from multimodalstudio_b200.models import MODALITY_CHANNELS
from multimodalstudio_b200.pipelines import RawPipeline, SyntheticScene

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="grid_raw")
ap.add_argument("--out", default="gpurun_out/ops.txt")
ap.add_argument("--top", type=int, default=45)
args = ap.parse_args()
wl = dict(bench.WORKLOADS[args.workload])
dev = torch.device("cuda", 0)
mods = {m: MODALITY_CHANNELS[m] for m in wl["modalities"]}
scene = SyntheticScene(mods, bench.split_rays(wl["rays"], wl["modalities"]), raw=wl["raw"])
pipe = RawPipeline(mods, scene.cameras, device=dev, raw=wl["raw"], render_all_heads=False, num_samples=wl["n_c"],
                   num_samples_importance=wl["n_i"], bg_samples=wl["bg"])
batches = [tuple({m: t.to(dev) for m, t in d.items()} for d in scene.sample_batch()) for _ in range(2)]
for i in range(3):
    pipe.train_step(bench.BASE_STEP + i, *batches[i % 2])
torch.cuda.synchronize()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CPU, torch.profiler.ProfilerActivity.CUDA],
                            record_shapes=True, with_stack=True) as prof:
    pipe.train_step(bench.BASE_STEP + 3, *batches[1])
    torch.cuda.synchronize()
rows = []
for ev in prof.key_averages(group_by_input_shape=True, group_by_stack_n=8):
    t = getattr(ev, "self_device_time_total", None)
    if t is None:
        t = ev.self_cuda_time_total
    if t > 0:
        rows.append((t, ev.count, ev.key, str(ev.input_shapes)[:120], [s for s in ev.stack if "multimodalstudio_b200" in s or "bench" in s][:4]))
rows.sort(key=lambda r: -r[0])
with open(args.out, "w") as fh:
    fh.write(f"total self device time of ops: {sum(r[0] for r in rows) / 1e3:.2f} ms\n")
    for t, c, k, sh, st in rows[: args.top]:
        fh.write(f"{t / 1e3:8.3f} ms {c:4d}x {k}  {sh}\n")
        for s in st:
            fh.write(f"            {s[-110:]}\n")
print(open(args.out).read()[:200])
