#!/bin/bash
# ncu evidence after the GEMM+LN fusion: launch list of one bs=512 forward + full captures of the four encoder GEMM kinds of block 0.
mkdir -p gpurun_out
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bs512_v5.csv python tests/profile_step.py 512 > gpurun_out/ncu_list.log 2>&1
tail -1 gpurun_out/ncu_list.log
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm -s 1 -c 4 -o gpurun_out/prof_gemm_v5 -f python tests/profile_step.py 512 > gpurun_out/ncu_gemm.log 2>&1
tail -1 gpurun_out/ncu_gemm.log
ls -la gpurun_out/*.ncu-rep
