#!/bin/bash
# First-contact script for the GPU box: building blocks first, then the model, each under a timeout.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --timeout 300 2>&1 | tail -40 > gpurun_out/kernels.log
tail -15 gpurun_out/kernels.log
timeout 900 python __graft_entry__.py smoke 2>&1 | tail -30 > gpurun_out/smoke.log
tail -20 gpurun_out/smoke.log
