#!/bin/bash
# usage: tools/ncu_full.sh <tag> [workload]  - one `ncu --set full` capture (with source) of the step kernel
tag=$1; wl=${2:-c3}
BENCH="python bench.py --workload $wl --steps 3 --warmup 3 --chunk 20 --no-cpu-baseline --secondary 0 --e2e-steps 1"
$BENCH > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:ballenv_kernel -s ${3:-3} -c 1 -o gpurun_out/prof_$tag $BENCH > gpurun_out/ncu2.log 2>&1
tail -2 gpurun_out/ncu2.log
