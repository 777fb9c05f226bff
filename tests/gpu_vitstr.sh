#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_vitstr.py -m gpu -q --timeout 600 2>&1 | tail -40 > gpurun_out/vitstr.log
tail -30 gpurun_out/vitstr.log
timeout 300 python tests/bench_vitstr.py 2>&1 | tail -4
