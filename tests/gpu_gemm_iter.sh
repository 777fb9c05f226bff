#!/bin/bash
# GEMM iteration: kernel unit tests, microbenchmark (incl. ring-depth sweep), short bench.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --timeout 300 -k "gemm" 2>&1 | tail -3
timeout 600 python tests/bench_gemm.py 65536 > gpurun_out/gemm_microbench.txt 2>&1
cat gpurun_out/gemm_microbench.txt
timeout 300 python bench.py --no-cpu-baseline 2>gpurun_out/bench.err | tee gpurun_out/bench_last.json | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(d['value'], 'img/s', d['ms_per_step'], 'ms/step | e2e', d['e2e']['value'], '| e2e_u8', d.get('e2e_u8', {}).get('value'))"
tail -2 gpurun_out/bench.err
