#!/bin/bash
# usage: tools/ncu_patch.sh <tag>  - `ncu --set full` capture of the rgb patch kernel (16 384 environments) with source lines
tag=$1
python tools/patch_rate.py > gpurun_out/plain_patch.log 2>&1 || { tail -3 gpurun_out/plain_patch.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:ballenv_patch_kernel -s 3 -c 1 -f -o gpurun_out/prof_${tag}_patch python tools/patch_rate.py > gpurun_out/ncu_patch.log 2>&1; tail -1 gpurun_out/ncu_patch.log
