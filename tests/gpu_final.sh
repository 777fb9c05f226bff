#!/bin/bash
# Round-end style validation: full GPU test suite, smoke, default bench, the side benches that feed DESIGN.md section 6.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/ -x -q -m gpu --timeout 600 2>&1 | tail -5
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python bench.py 2>gpurun_out/bench.err | tee gpurun_out/bench_final.json | cut -c1-400
tail -2 gpurun_out/bench.err
timeout 300 python tests/bench_vitstr.py 2>&1 | tail -3
timeout 600 python tests/bench_configs.py 2>&1 | tail -12
