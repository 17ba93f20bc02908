#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_integration.py -x -q --timeout 150 2>&1 | tail -30 > gpurun_out/pytest_integ.log; cat gpurun_out/pytest_integ.log
