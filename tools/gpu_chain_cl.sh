#!/bin/bash
# chained forward launches as CTA pairs (multicast weight taps): bit-identity tests, batch-1 / batch-8 forward A/B
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q --timeout 120 -k "chain" 2>&1 | tail -4
for v in 0 1; do
  echo "== SR4D_CHAIN_CLUSTER=$v"
  SR4D_CHAIN_CLUSTER=$v timeout 200 python tools/time_fwd.py 1 2>&1 | grep -v simt | tail -2
  SR4D_CHAIN_CLUSTER=$v timeout 200 python tools/time_fwd.py 8 2>&1 | grep -v simt | tail -2
done
SR4D_NO_CHAIN=1 timeout 200 python tools/time_fwd.py 1 2>&1 | grep -v simt | tail -2
