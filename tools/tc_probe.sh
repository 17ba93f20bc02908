#!/bin/bash
# each case in its own process with a timeout so a hung kernel cannot eat the GPU budget
for c in "center_identity 8 1" "center_random 8 1" "tap_z 8 1" "tap_y 8 1" "tap_x 8 1" "random 8 2" "random 24 1" "random 48 1" "random 20 1" "random 30 1"; do
  timeout 90 python tools/tc_probe.py $c 2>&1 | tail -8 || echo "CASE $c: timeout/fail rc=$?"
done
