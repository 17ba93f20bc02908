#!/bin/bash
mkdir -p gpurun_out
SR4D_TC_DEBUG=1 timeout -s KILL 300 python tools/train_once.py 8 2 2>&1 | grep "head_bwd dbg" | tail -1 | tee gpurun_out/head_bwd_dbg.txt
SR4D_RECORD_BARS=gpurun_out/test_bars_head.jsonl timeout -s KILL 300 python -m pytest tests/test_gpu_backward.py -k "head_layer_bwd" -m gpu -x -q --timeout 120 2>&1 | tail -5
cat gpurun_out/test_bars_head.jsonl | tail -30
