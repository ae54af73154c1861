#!/bin/bash
# usage (here): tools/lean_variants.sh build "<defs of V1>" "<defs of V2>" ...   - builds libballenv_b200_V<i>.so
#       (GPU):  tools/lean_variants.sh run <n>                                    - rollout / per-step rates of A and V1..Vn
if [ "$1" = build ]; then
  shift; i=1
  for d in "$@"; do
    BALLENV_EXTRA_DEFS="$d" BALLENV_OBJ_SUFFIX=_V$i BALLENV_LIB_NAME=libballenv_b200_V$i.so python gym_ballenv_b200/build.py > /tmp/build_V$i.log 2>&1 || tail -5 /tmp/build_V$i.log
    grep -E "lean_w10_s8_d24_g2.*(Used|spill)" /tmp/build_V$i.log | grep -v " 0 bytes spill" | tail -2
    i=$((i+1))
  done
else
  n=${2:-1}
  for pass in 1 2; do
    echo "--- A"; NS=65536 python tools/rollout_rate.py | grep -v "^$"
    for i in $(seq 1 $n); do
      echo "--- V$i"; BALLENV_LIB_PATH=$PWD/gym_ballenv_b200/libballenv_b200_V$i.so NS=65536 python tools/rollout_rate.py
    done
  done
fi
