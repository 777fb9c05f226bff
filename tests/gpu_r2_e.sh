#!/bin/bash
# round 2, run E: 16 co-resident clusters, per-warp slot release in the cluster AR kernel
mkdir -p gpurun_out
timeout 200 python tests/prof_ar.py 512 2 2>&1 | tail -5 | tee gpurun_out/r2e_prof_ar2.txt
timeout 200 python tests/prof_ar.py 1 2 2>&1 | tail -4 | tee gpurun_out/r2e_prof_ar2_bs1.txt
timeout 1500 python -m pytest tests/ -q -m gpu --timeout 900 2>&1 | tail -30 | tee gpurun_out/r2e_tests.txt
timeout 900 python bench.py 2>gpurun_out/bench.err | tee gpurun_out/r2e_bench.json | cut -c1-900
tail -3 gpurun_out/bench.err
