#!/bin/bash
mkdir -p gpurun_out /tmp/rep
timeout 600 ncu --profile-from-start off --set full --clock-control none -k regex:gemm_bf16_tcgen05 -o /tmp/rep/cg python tests/prof_gemm_cg.py > gpurun_out/r2_ncu_cg.log 2>&1
tail -2 gpurun_out/r2_ncu_cg.log
ncu -i /tmp/rep/cg.ncu-rep --page raw --csv > gpurun_out/r2_raw_cg.csv 2>/dev/null
ncu -i /tmp/rep/cg.ncu-rep --page details > gpurun_out/r2_details_cg.txt 2>/dev/null
ls -la gpurun_out/r2_raw_cg.csv gpurun_out/r2_details_cg.txt
