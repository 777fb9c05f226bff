#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 2>&1 | tail -150 > gpurun_out/parity.log
grep -E "passed|failed" gpurun_out/parity.log | tail -3
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 300 python bench.py --chunk 64 --no-cpu-baseline --no-latency > gpurun_out/bench_c64.json 2>&1; cat gpurun_out/bench_c64.json
timeout 300 python bench.py --chunk 256 --no-cpu-baseline --no-latency > gpurun_out/bench_c256.json 2>&1; cat gpurun_out/bench_c256.json
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1.csv python tests/profile_step.py 128 > gpurun_out/ncu_list.log 2>&1
tail -2 gpurun_out/ncu_list.log
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_bf16_tcgen05 -s 8 -c 4 -o gpurun_out/prof_gemm_r1 python tests/profile_step.py 128 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
