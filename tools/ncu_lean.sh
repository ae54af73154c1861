#!/bin/bash
# usage: tools/ncu_lean.sh <tag>  - `ncu --set full` captures of the thread-per-environment kernels: C3 rollout (20 steps
# per launch), C3 single step, W=5 rollout.  tools/rollout_rate.py with T=20 launches, per workload, 6 rollout kernels and
# then 120 single-step kernels.
tag=$1
export T=20 NS=65536
CMD="python tools/rollout_rate.py"
$CMD > gpurun_out/plain.log 2>&1 || { tail -5 gpurun_out/plain.log; exit 1; }
for cap in "c3roll 3" "c3step 60" "w5roll 128" "w5step 190"; do
  set -- $cap
  ncu --set full --clock-control none --import-source on -k regex:ballenv_lean_kernel -s $2 -c 1 -f -o gpurun_out/prof_${tag}_$1 $CMD > gpurun_out/ncu_$1.log 2>&1
  tail -1 gpurun_out/ncu_$1.log
done
