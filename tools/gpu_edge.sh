#!/bin/bash
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_edge_cases.py -x -q --timeout 200 2>&1 | tail -30 > gpurun_out/pytest_edge.log; cat gpurun_out/pytest_edge.log
