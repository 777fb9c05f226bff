#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 2>&1 | tail -40 > gpurun_out/parity.log
grep -E "passed|failed|Error|assert" gpurun_out/parity.log | tail -8
timeout 600 python tests/bench_configs.py 2>&1 | tail -12
