#!/bin/bash
# round 2, run C: blocked K/V cache + rotating TMA issuer in the cluster AR kernel, packed-FFMA2 GELU epilogue
mkdir -p gpurun_out
timeout 300 python tests/diag_golden.py b48_sharp_ar1_b2 2>&1 | tail -8 | tee gpurun_out/r2c_diag.txt
timeout 300 python tests/diag_golden.py s_sharp_ar1_b2 2>&1 | tail -8 | tee -a gpurun_out/r2c_diag.txt
timeout 200 python tests/prof_ar.py 512 2 2>&1 | tail -8 | tee gpurun_out/r2c_prof_ar2.txt
timeout 200 python tests/prof_ar.py 1 2 2>&1 | tail -8 | tee gpurun_out/r2c_prof_ar2_bs1.txt
timeout 1500 python -m pytest tests/ -q -m gpu --timeout 900 2>&1 | tail -25 | tee gpurun_out/r2c_tests.txt
timeout 900 python bench.py 2>gpurun_out/bench.err | tee gpurun_out/r2c_bench.json | cut -c1-1500
tail -3 gpurun_out/bench.err
timeout 300 python tests/bench_gemm.py 2>&1 | tail -30 | tee gpurun_out/r2c_gemm_microbench.txt
