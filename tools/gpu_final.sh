#!/bin/bash
# final check of a session: smoke, the whole gpu suite, both bench arms (no profiler)
mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python -m pytest tests -m gpu -x -q --timeout 200 2>&1 | tail -4 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -2 gpurun_out/bench_n1.err; cut -c1-200 gpurun_out/bench_n1.json
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -1 gpurun_out/bench_ref.err; cut -c1-300 gpurun_out/bench_ref.json
