#!/bin/bash
mkdir -p gpurun_out
for c in 1 3 5; do timeout 280 python tools/bench_configs.py --config $c 2>&1 | tail -2; done | tee gpurun_out/configs_n1.log
