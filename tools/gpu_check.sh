#!/bin/bash
# usage: tools/gpu_check.sh <tag>   - parity tests, launch-rate table, ncu instruction counts of the step kernel
tag=$1
python __graft_entry__.py smoke 2>&1 | tail -2
python -m pytest tests -x -q -m gpu 2>&1 | tail -8
python tools/launch_rate.py 2>&1 | tail -12
BENCH="python bench.py --steps 3 --warmup 3 --chunk 20 --no-cpu-baseline --secondary 1 --e2e-steps 1"
$BENCH > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -s 70 -c 20 --csv --log-file gpurun_out/launches_$tag.csv $BENCH > gpurun_out/ncu1.log 2>&1
python tools/ncu_summary.py gpurun_out/launches_$tag.csv
