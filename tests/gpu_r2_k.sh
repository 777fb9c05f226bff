#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 200 -x -k "teacher_forced and (s_ar1_b2 or ti_ar1_b3 or b48_ar1)" 2>&1 | tail -4
timeout 200 python tests/prof_ar.py 512 2 2>&1 | tail -7 | tee gpurun_out/r2k_prof_ar2.txt
timeout 200 python tests/prof_ar.py 1 2 2>&1 | tail -6 | tee gpurun_out/r2k_prof_ar2_bs1.txt
timeout 1500 python -m pytest tests/ -q -m gpu --timeout 900 2>&1 | tail -12 | tee gpurun_out/r2k_tests.txt
timeout 900 python bench.py --no-configs 2>gpurun_out/bench.err | tee gpurun_out/r2k_bench.json | cut -c1-400
tail -3 gpurun_out/bench.err
