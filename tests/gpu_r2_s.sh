#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu --timeout 300 -k "gemm" 2>&1 | tail -5
timeout 300 python tests/bench_gemm.py 2>&1 | head -22 | tee gpurun_out/r2s_gemm_microbench.txt
