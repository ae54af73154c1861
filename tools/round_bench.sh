#!/bin/bash
# usage: tools/round_bench.sh <tag>  - default bench line, then the ncu launch list of a short form of the same command
# (every launch with its device time: compare SHARES, the times are cold-cache and serialised)
tag=$1
python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; tail -c 1500 gpurun_out/bench_$tag.json; tail -3 gpurun_out/bench_$tag.err
SHORT="python bench.py --steps 3 --warmup 3 --chunk 20 --no-cpu-baseline --secondary 0 --e2e-steps 1 --closed-loop-steps 1 --c5 0 --e2e-u8 0"
$SHORT > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$tag.csv $SHORT > gpurun_out/ncu1.log 2>&1
python tools/ncu_summary.py gpurun_out/launches_$tag.csv
