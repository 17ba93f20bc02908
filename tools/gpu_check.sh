#!/bin/bash
# session check: new reference-graph golden tests first, then the whole gpu suite, then the default bench
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_graph_golden.py -m gpu -q --timeout 120 2>&1 | tail -15 > gpurun_out/pytest_golden.log; cat gpurun_out/pytest_golden.log
timeout 600 python -m pytest tests -m gpu -x -q --timeout 200 --durations=8 2>&1 | tail -25 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; cut -c1-600 gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
