#!/bin/bash
# fused GEMM+LN on CTA pairs: bit-identity test, microbenchmark, whole-step bench with the three settings
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -m gpu -k "gemm_ln" > gpurun_out/r2z_tests.txt 2>&1; echo "tests rc=$?" >> gpurun_out/r2z_tests.txt
tail -5 gpurun_out/r2z_tests.txt
timeout 200 python tests/bench_gemm_ln.py > gpurun_out/r2z_gemm_ln.txt 2>&1; cat gpurun_out/r2z_gemm_ln.txt
for v in "1 0" "2 0" "2 1"; do set -- $v
  timeout 300 python bench.py --steps 20 --warmup 5 --no-configs --no-parity --no-cpu-baseline --no-latency --ln-cta-group $1 --pair-pdl $2 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ln_cta_group=$1 pair_pdl=$2', d['value'], d['ms_per_step'], {k: round(v, 3) for k, v in d['roofline']['by_category_ms'].items()})"
done
