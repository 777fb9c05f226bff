#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tests/bench_tma_stream.py 2>&1 | tail -60 | tee gpurun_out/r2h_tma_stream.txt
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "batch_invariant" 2>&1 | tail -4
