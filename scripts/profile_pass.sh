#!/bin/bash
# The profile pass behind profiles/<tag>_*: GPU tests, layer-product tables, `ncu --set full` of the dominant kernels,
# the ncu launch list of the bench command (after the same command exited 0 without ncu), CUPTI step timelines and the
# two bench lines.  Run on the GPU box from the repo root:  bash scripts/profile_pass.sh r1f
tag=${1:-rX}
o=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -2 > $o/pytest_$tag.log
python scripts/dev_tc_step_shapes.py > $o/tc_shapes_$tag.log 2>&1
python scripts/ncu_targets.py > $o/ncu_targets_$tag.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:"tc_rows|tc_wgrad|hashgrid" --launch-skip 5 -c 5 \
      -o $o/prof_$tag python scripts/ncu_targets.py > $o/ncu_$tag.log 2>&1
python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > $o/bench_for_ncu_$tag.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 3000 -c 900 --csv --log-file $o/launches_$tag.csv \
      python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > $o/ncu_launches_$tag.log 2>&1
python scripts/profile_step.py --out $o/kernels_$tag.txt > /dev/null 2>&1
python scripts/profile_step.py --workload grid --out $o/kernels_grid_$tag.txt > /dev/null 2>&1
MMSB_BENCH_TABLE=$o/table_grid_$tag.txt python bench.py > $o/bench_${tag}_grid.log 2>&1
MMSB_BENCH_TABLE=$o/table_graw_$tag.txt python bench.py --workload grid_raw --no-cpu-baseline > $o/bench_${tag}_graw.log 2>&1
cat $o/pytest_$tag.log; tail -2 $o/ncu_$tag.log
