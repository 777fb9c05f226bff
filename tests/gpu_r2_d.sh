#!/bin/bash
# round 2, run D: ILP in the cluster AR kernel's mma chains; decode-mask fix; bench
mkdir -p gpurun_out
timeout 300 python tests/diag_golden.py s_sharp_ar1_b2 2>&1 | tail -12 | tee gpurun_out/r2d_diag.txt
timeout 300 python tests/diag_golden.py s_ar1_b2 2>&1 | tail -12 | tee -a gpurun_out/r2d_diag.txt
timeout 200 python tests/prof_ar.py 512 2 2>&1 | tail -8 | tee gpurun_out/r2d_prof_ar2.txt
timeout 200 python tests/prof_ar.py 1 2 2>&1 | tail -8 | tee gpurun_out/r2d_prof_ar2_bs1.txt
timeout 1500 python -m pytest tests/ -q -m gpu --timeout 900 2>&1 | tail -40 | tee gpurun_out/r2d_tests.txt
timeout 900 python bench.py 2>gpurun_out/bench.err | tee gpurun_out/r2d_bench.json | cut -c1-1500
tail -3 gpurun_out/bench.err
