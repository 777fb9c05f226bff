#!/bin/bash
# gemm_ln2 (column-split CTA pair) with the x prefetch: unit tests, sanitizer, microbenchmark, whole step (auto rule)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -m gpu -k "gemm_ln" > gpurun_out/r2n_tests.txt 2>&1; echo "tests rc=$?" >> gpurun_out/r2n_tests.txt
tail -6 gpurun_out/r2n_tests.txt
if grep -q "tests rc=0" gpurun_out/r2n_tests.txt; then
  timeout 300 compute-sanitizer --tool memcheck python tests/sanitize_new_kernels.py 2>&1 | grep -E "ERROR SUMMARY|Error|error|ok:|done|Traceback" | head -12 | tee gpurun_out/r2n_sanitizer.txt
  timeout 200 python tests/bench_gemm_ln.py > gpurun_out/r2n_gemm_ln.txt 2>&1; grep -A3 "ln_split=2\|ln_split=1 (" gpurun_out/r2n_gemm_ln.txt | grep "65536\|---"
  for v in 1 0; do
    timeout 300 python bench.py --steps 20 --warmup 5 --no-configs --no-cpu-baseline --no-latency --no-two-in-flight --ln-split $v 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ln_split=$v', round(d['value']), round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value']), {k: round(v, 3) for k, v in d['roofline']['by_category_ms'].items()}, (d.get('parity') or {}).get('ok'))" | tee -a gpurun_out/r2n_step.txt
  done
fi
