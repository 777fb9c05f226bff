#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/r2_compute_sanitizer.txt
bash tests/gpu_sanitize.sh
timeout 900 python bench.py 2>gpurun_out/bench.err | tee gpurun_out/r2n_bench.json | cut -c1-300
tail -3 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>gpurun_out/bench_ref.err | tee gpurun_out/r2n_bench_ref.json | cut -c1-600
