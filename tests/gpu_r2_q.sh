#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu --timeout 300 -k "enc_attention" 2>&1 | tail -8
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_vitstr.py -q -m gpu --timeout 300 -k "p16 or b48 or vitstr" 2>&1 | tail -8
timeout 300 python tests/bench_vitstr.py 2>&1 | tail -3
timeout 600 python tests/bench_configs.py 2>&1 | grep -i "C5\|patch16" | tail -5
