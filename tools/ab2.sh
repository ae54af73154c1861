#!/bin/bash
# usage: tools/ab2.sh "<defs B>" "<defs C>" ...  - baseline A and variants on the same box, two alternating passes
python gym_ballenv_b200/build.py > /dev/null 2>&1
i=0
for d in "$@"; do
  i=$((i+1))
  BALLENV_EXTRA_DEFS="$d" BALLENV_LIB_NAME=libballenv_b200_V$i.so BALLENV_OBJ_SUFFIX=_V$i python gym_ballenv_b200/build.py > /dev/null 2>&1
done
for pass in 1 2; do
  echo "--- A"; python tools/rollout_rate.py | grep "65536 rollout"
  i=0
  for d in "$@"; do
    i=$((i+1))
    echo "--- V$i ($d)"; BALLENV_LIB_PATH=$PWD/gym_ballenv_b200/libballenv_b200_V$i.so python tools/rollout_rate.py | grep "65536 rollout"
  done
done
