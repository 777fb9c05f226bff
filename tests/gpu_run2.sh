#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tests/diag_numerics.py > gpurun_out/diag.log 2>&1; tail -30 gpurun_out/diag.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 2>&1 | tail -60 > gpurun_out/parity.log
tail -40 gpurun_out/parity.log
