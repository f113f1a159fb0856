"""Turns the raw artefacts of scripts/profile_pass.sh (gpurun_out/*_<tag>*) into the summaries under profiles/.

    python scripts/summarize_profiles.py r2
"""
import collections, csv, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
o, p = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")

# ---- plain copies
for src, dst in [(f"kernels_{tag}.txt", f"{tag}_step_timeline_sweep_microbatch.txt"), (f"table_{tag}.txt", f"{tag}_entry_point_table_sweep.txt"),
                 ("parity_measured.jsonl", f"{tag}_parity_measured.jsonl")]:
    if os.path.exists(os.path.join(o, src)):
        shutil.copy(os.path.join(o, src), os.path.join(p, dst))
line = [l for l in open(os.path.join(o, f"bench_{tag}.json")) if l.strip().startswith("{")][-1]
json.dump(json.loads(line), open(os.path.join(p, f"{tag}_bench_sweep.json"), "w"), indent=1)

# ---- ncu launch list of the bench command
rows = list(csv.DictReader(l for l in open(os.path.join(o, f"launches_{tag}.csv")) if l.startswith('"')))
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    if r["Metric Name"] != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    v = v / 1e3 if r["Metric Unit"] in ("ns", "nsecond") else (v * 1e3 if r["Metric Unit"] in ("ms", "msecond") else v)
    a = agg[r["Kernel Name"]]
    a[0] += 1
    a[1] += v
tot = sum(v[1] for v in agg.values())
with open(os.path.join(p, f"{tag}_ncu_launches.csv"), "w") as fh:
    fh.write("kernel,launches,total_us\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        fh.write(f"\"{k}\",{v[0]},{v[1]:.1f}\n")
mine = sum(v[1] for k, v in agg.items() if not k.startswith("void at::") and "memcpy" not in k and "memset" not in k)
with open(os.path.join(p, f"{tag}_ncu_launches.md"), "w") as fh:
    fh.write(f"# Round 2 — ncu launch list of the bench command (per-launch `gpu__time_duration.sum`)\n\n"
             f"Command (`scripts/profile_pass.sh {tag}`): `MMSB_PROFILER_RANGE=1 ncu --metrics gpu__time_duration.sum --clock-control none "
             f"--profile-from-start off --csv --log-file gpurun_out/launches_{tag}.csv python bench.py --steps 1 --warmup 3 --no-e2e "
             f"--no-cpu-baseline --no-side`\n"
             f"(default workload: 5 raw modalities, 65 560-ray global batch x 256 samples in 8 micro-batches of 8195 rays; the same command "
             f"exited 0 without ncu first; aggregated list: `profiles/{tag}_ncu_launches.csv`.  bench.py brackets its timed region with "
             f"cudaProfilerStart/Stop when MMSB_PROFILER_RANGE is set, so the capture is EXACTLY the one timed optimizer step: the "
             f"ray-generation / collider pre-pass, 8 graph-replayed micro-batches (kernel nodes profiled one by one), the optimizer "
             f"graph.  Times are cold-cache and serialised under ncu — compare SHARES with `profiles/{tag}_step_timeline_sweep_microbatch.txt` "
             f"(torch.profiler / CUPTI, no replay) and with bench.py's `roofline.per_class`.)\n\n"
             f"{sum(v[0] for v in agg.values())} launches, {tot / 1e3:.2f} ms of kernel time, {100 * mine / tot:.1f} % in libmms_b200 kernels\n\n"
             f"| time (us) | launches | share | kernel |\n|---:|---:|---:|---|\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]:
        fh.write(f"| {v[1]:.1f} | {v[0]} | {100 * v[1] / tot:.1f}% | `{k[:110]}` |\n")

# ---- ncu --set full captures -> one row per kernel
def raw_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(out.splitlines()))
    hdr = rr[0]
    return [dict(zip(hdr, r)) for r in rr[2:]]
M = {"time_ms": "gpu__time_duration.sum", "dram_rd": "dram__bytes_read.sum", "dram_wr": "dram__bytes_write.sum",
     "dram_pct": "dram__throughput.avg.pct_of_peak_sustained_elapsed", "tensor_pct": "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
     "sm_pct": "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex_pct": "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
     "lts_pct": "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_hit": "lts__t_sector_hit_rate.pct",
     "issue_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active", "regs": "launch__registers_per_thread",
     "smem_conflicts": "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_wavefronts": "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"}
def fl(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return float("nan")
res = []
for rep in (f"prof_{tag}.ncu-rep",):
    path = os.path.join(o, rep)
    if os.path.exists(path):
        for r in raw_rows(path):
            res.append({"kernel": r["Kernel Name"], "grid": r.get("Grid Size", ""), **{k: fl(r.get(m, "nan")) for k, m in M.items()}})
json.dump(res, open(os.path.join(p, f"{tag}_ncu_full_raw.json"), "w"), indent=1)
for r in res:
    print(f'{r["kernel"][:60]:60s} {r["time_ms"]:.3f} ms  dram {r["dram_rd"]:.3f}+{r["dram_wr"]:.3f} GB ({r["dram_pct"]:.0f}%)  tensor {r["tensor_pct"]:.0f}%  sm {r["sm_pct"]:.0f}%  '
          f'l1tex {r["l1tex_pct"]:.0f}%  lts {r["lts_pct"]:.0f}%  L2hit {r["l2_hit"]:.0f}%  issue {r["issue_pct"]:.0f}%  regs {r["regs"]:.0f}  confl {r["smem_conflicts"]/1e6:.0f}M/{r["smem_wavefronts"]/1e6:.0f}M')
