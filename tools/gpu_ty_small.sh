#!/bin/bash
# batch-1 LR grid: tile height of the chained forward (two rounds of half-height tiles overlap a tile's epilogue with the next tile's MMAs)
mkdir -p gpurun_out
for ty in 0 6 8 12; do
  echo "== SR4D_FWD_TY_SMALL=$ty"
  SR4D_FWD_TY_SMALL=$ty timeout 200 python tools/time_fwd.py 1 2>&1 | grep -v simt | tail -2
done
SR4D_FWD_TY_SMALL=6 timeout 300 python -m pytest tests -m gpu -x -q --timeout 120 -k "chain or forward" 2>&1 | tail -4
