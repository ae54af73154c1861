#!/bin/bash
# usage: tools/ab.sh "<extra defs for variant B>"   - A/B of two kernel builds on the same box, alternating runs
python gym_ballenv_b200/build.py > /dev/null 2>&1
BALLENV_EXTRA_DEFS="$1" BALLENV_LIB_NAME=libballenv_b200_B.so BALLENV_OBJ_SUFFIX=_B python gym_ballenv_b200/build.py > /dev/null 2>&1
for i in 1 2 3; do
  echo "--- A"; python tools/rollout_rate.py | grep 65536
  echo "--- B ($1)"; BALLENV_LIB_PATH=$PWD/gym_ballenv_b200/libballenv_b200_B.so python tools/rollout_rate.py | grep 65536
done
