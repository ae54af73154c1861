#!/bin/bash
# lanes per environment of the lean kernels, A/B: tools/rollout_rate.py for G = 1, 2 and the block-of-roles kernel
for g in 1 2; do echo "--- lean G=$g"; BALLENV_LEAN_G=$g NS=${NS:-65536} python tools/rollout_rate.py 2>&1 | tail -4; done
echo "--- roles"; BALLENV_NO_LEAN=1 NS=${NS:-65536} python tools/rollout_rate.py 2>&1 | tail -4
