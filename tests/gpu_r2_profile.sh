#!/bin/bash
# round 2 ncu evidence: launch list of one bs=512 forward + --set full captures of the kernels >= 10 % of the step.
# The .ncu-rep files exceed what gpurun copies back: export the raw / details pages to CSV / text on the box, drop the reports.
mkdir -p gpurun_out /tmp/rep
P="python tests/profile_step.py 512"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_bs512.csv $P > gpurun_out/r2_ncu_list.log 2>&1
tail -1 gpurun_out/r2_ncu_list.log
cap() {   # name, kernel regex, extra args
  timeout 900 ncu --profile-from-start off --set full --clock-control none -k regex:$2 $3 -o /tmp/rep/$1 $P > gpurun_out/r2_ncu_$1.log 2>&1
  tail -1 gpurun_out/r2_ncu_$1.log
  ncu -i /tmp/rep/$1.ncu-rep --page raw --csv > gpurun_out/r2_raw_$1.csv 2>/dev/null
  ncu -i /tmp/rep/$1.ncu-rep --page details > gpurun_out/r2_details_$1.txt 2>/dev/null
}
cap ar2 dec_ar2 "-c 1"
cap gemm_ln gemm_ln_fused "-c 2"
cap gemm gemm_bf16_tcgen05 "-s 1 -c 2"
cap attn enc_attention_tc "-c 1"
# the opt-in one-kernel MLP + LayerNorm inside the same step (single-CTA variant): what it does to the DRAM bytes of a block
export PQ_FUSE_MLP=1 PQ_MLP_CTA_GROUP=1
cap mlp_ln mlp_ln_fused "-c 1"
unset PQ_FUSE_MLP PQ_MLP_CTA_GROUP
ls -la gpurun_out/ | tail -20
du -sh gpurun_out
