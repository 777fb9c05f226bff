#!/bin/bash
# round 2: first contact of the cluster AR kernel (dec_ar2): a few goldens, phase stamps, then the whole suite + bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 300 -x -k "teacher_forced and (s_ar1_b2 or ti_ar1_b3 or b48_ar1 or sharp)" 2>&1 | tail -15 | tee gpurun_out/r2b_first.txt
timeout 200 python tests/prof_ar.py 512 2 2>&1 | tail -8 | tee gpurun_out/r2b_prof_ar2.txt
timeout 200 python tests/prof_ar.py 512 1 2>&1 | tail -8 | tee gpurun_out/r2b_prof_ar1.txt
timeout 200 python tests/prof_ar.py 1 2 2>&1 | tail -8 | tee gpurun_out/r2b_prof_ar2_bs1.txt
timeout 1500 python -m pytest tests/ -q -m gpu --timeout 900 2>&1 | tail -25 | tee gpurun_out/r2b_tests.txt
timeout 600 python bench.py 2>gpurun_out/bench.err | tee gpurun_out/r2b_bench.json | cut -c1-700
tail -2 gpurun_out/bench.err
