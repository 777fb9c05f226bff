#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1_v4_bs512.csv python tests/profile_step.py 512 > gpurun_out/ncu_list.log 2>&1
tail -2 gpurun_out/ncu_list.log
