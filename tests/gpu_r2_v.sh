#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py -q -m gpu --timeout 300 -k "gemm or cta_pair" 2>&1 | tail -4
timeout 600 python tests/bench_gemm_cg.py 2>&1 | tail -12 | tee gpurun_out/r2v_gemm_cg.txt
