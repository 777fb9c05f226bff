#!/bin/bash
# round 2, first GPU pass: full GPU suite on the refactored build, AR phase stamps, default bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -q -m gpu --timeout 900 2>&1 | tail -25 | tee gpurun_out/r2a_tests.txt
timeout 300 python tests/prof_ar.py 2>&1 | tail -6 | tee gpurun_out/r2a_prof_ar.txt
timeout 600 python bench.py 2>gpurun_out/bench.err | tee gpurun_out/r2a_bench.json | cut -c1-600
tail -2 gpurun_out/bench.err
