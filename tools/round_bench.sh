#!/bin/bash
# usage: tools/round_bench.sh <tag>  - default bench line, launch list and one full ncu capture of the rollout kernel
tag=$1
python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; tail -c 4000 gpurun_out/bench_$tag.json; tail -3 gpurun_out/bench_$tag.err
SHORT="python bench.py --steps 3 --warmup 3 --chunk 20 --no-cpu-baseline --secondary 0 --e2e-steps 1 --closed-loop-steps 1"
$SHORT > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv $SHORT > gpurun_out/ncu1.log 2>&1
python tools/ncu_summary.py gpurun_out/launches_$tag.csv
$SHORT > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:ballenv_kernel -s 3 -c 2 -o gpurun_out/prof_$tag $SHORT > gpurun_out/ncu2.log 2>&1
tail -2 gpurun_out/ncu2.log
