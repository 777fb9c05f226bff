#!/bin/bash
# Round-end style validation: full GPU test suite, smoke, default bench.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu --timeout 900 2>&1 | tail -5 | tee gpurun_out/r2_final_tests.txt
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/r2_final_smoke.txt
timeout 900 python bench.py 2>gpurun_out/bench.err | tee gpurun_out/r2_bench_final_n1.json | cut -c1-300
tail -2 gpurun_out/bench.err
