#!/bin/bash
# parity tests + a short bench (both in one GPU call)
mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -x -q --timeout 120 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
