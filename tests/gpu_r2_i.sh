#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tests/prof_ar.py 512 2 2>&1 | tail -7 | tee gpurun_out/r2i_prof_ar2.txt
timeout 200 python tests/prof_ar.py 1 2 2>&1 | tail -6 | tee gpurun_out/r2i_prof_ar2_bs1.txt
timeout 1500 python -m pytest tests/ -q -m gpu --timeout 900 2>&1 | tail -12 | tee gpurun_out/r2i_tests.txt
timeout 900 python bench.py --no-configs 2>gpurun_out/bench.err | tee gpurun_out/r2i_bench.json | cut -c1-400
tail -3 gpurun_out/bench.err
