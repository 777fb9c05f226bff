#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --timeout 300 2>&1 | tail -4
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 2>&1 | tail -40 > gpurun_out/parity.log
grep -E "passed|failed|Error|assert" gpurun_out/parity.log | tail -12
