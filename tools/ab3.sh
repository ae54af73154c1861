#!/bin/bash
# usage: tools/ab3.sh <n variants>  - baseline library and the prebuilt variant libraries libballenv_b200_V<i>.so
# (built here with BALLENV_EXTRA_DEFS / BALLENV_MINBLOCKS, they travel with the snapshot), two alternating passes
n=${1:-1}
for pass in 1 2; do
  echo "--- A"; python tools/rollout_rate.py | grep "n=  65536"
  for i in $(seq 1 $n); do
    echo "--- V$i"; BALLENV_LIB_PATH=$PWD/gym_ballenv_b200/libballenv_b200_V$i.so python tools/rollout_rate.py | grep "n=  65536"
  done
done
