#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tests/bench_configs.py 2>&1 | tail -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 2> gpurun_out/bench2.err > gpurun_out/bench_n2.json
cut -c1-400 gpurun_out/bench_n2.json; tail -2 gpurun_out/bench2.err
