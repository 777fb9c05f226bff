#!/bin/bash
# ncu evidence for profiles/: launch list of one bs=512 forward + captures of the top kernels.
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bs512.csv python tests/profile_step.py 512 > gpurun_out/ncu_list.log 2>&1
tail -1 gpurun_out/ncu_list.log
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_bf16_tcgen05 -s 1 -c 4 -o gpurun_out/prof_gemm python tests/profile_step.py 512 > gpurun_out/ncu_gemm.log 2>&1
tail -1 gpurun_out/ncu_gemm.log
timeout 900 ncu --profile-from-start off --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --section WarpStateStats --clock-control none -k regex:dec_ar_kernel -c 1 -o gpurun_out/prof_dec_ar python tests/profile_step.py 512 > gpurun_out/ncu_decar.log 2>&1
tail -1 gpurun_out/ncu_decar.log
timeout 600 ncu --profile-from-start off --set full --clock-control none -k regex:enc_attention_tc -c 1 -o gpurun_out/prof_attn python tests/profile_step.py 512 > gpurun_out/ncu_attn.log 2>&1
tail -1 gpurun_out/ncu_attn.log
