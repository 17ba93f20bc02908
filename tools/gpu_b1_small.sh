#!/bin/bash
# batch-1 forward: single-chunk epilogue of the TY = 12 tiles, x-segmented head sum, merged stems / 8-channel small-grid kernels
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q --timeout 150 -k "forward or chain or predictor or edge or golden or c_client" 2>&1 | tail -4
timeout 120 python tools/time_fwd.py 1 2>&1 | tail -3
timeout 120 python tools/time_fwd.py 8 2>&1 | tail -3
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_fwd_b1.csv python tools/fwd_once.py 1 3 > gpurun_out/ncu_b1.log 2>&1; tail -1 gpurun_out/ncu_b1.log
python tools/ncu_summary.py launches gpurun_out/launches_fwd_b1.csv | head -24
