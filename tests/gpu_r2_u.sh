#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_kernels.py -q -m gpu --timeout 600 -k "cta_pair or b48 or gemm or full_size or p16" 2>&1 | tail -6
timeout 600 python tests/bench_configs.py 2>&1 | grep -i "C5\|C2\|patch16\|C4" | tail -6
