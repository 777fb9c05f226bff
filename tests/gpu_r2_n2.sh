#!/bin/bash
# 2 GPUs: the NCCL id all-gather test and the N=2 bench line (torchrun, one rank per GPU)
mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu --timeout 500 2>&1 | tail -4 | tee gpurun_out/r2_n2_test.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-configs 2>gpurun_out/bench_n2.err | tee gpurun_out/r2_bench_n2.json | cut -c1-300
tail -3 gpurun_out/bench_n2.err
