#!/bin/bash
# The profile pass behind profiles/<tag>_*: GPU tests, `ncu --set full` of the dominant kernels, the ncu launch list of
# the bench command (after the same command exited 0 without ncu), the CUPTI step timeline, the entry-point table and the
# bench line.  Run on the GPU box from the repo root:  bash scripts/profile_pass.sh r2
tag=${1:-rX}
o=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -2 > $o/pytest_$tag.log
python scripts/ncu_targets.py > $o/ncu_targets_$tag.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:"tc_rows|tc_wgrad|hashgrid|sdf_fused" --launch-skip 9 -c 9 \
      -o $o/prof_$tag python scripts/ncu_targets.py > $o/ncu_$tag.log 2>&1
python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-side > $o/bench_for_ncu_$tag.log 2>&1 && \
  MMSB_PROFILER_RANGE=1 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $o/launches_$tag.csv \
      python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-side > $o/ncu_launches_$tag.log 2>&1
python scripts/profile_step.py --steps 3 --out $o/kernels_$tag.txt > /dev/null 2>&1
MMSB_BENCH_TABLE=$o/table_$tag.txt python bench.py > $o/bench_$tag.json 2> $o/bench_$tag.err
cat $o/pytest_$tag.log; tail -2 $o/ncu_$tag.log; tail -c 300 $o/bench_$tag.json
