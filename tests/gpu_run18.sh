#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --timeout 300 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 2>&1 | tail -30 > gpurun_out/parity.log
grep -E "passed|failed|Error" gpurun_out/parity.log | tail -5
timeout 300 python tests/bench_modes.py 2>&1 | tail -2
timeout 300 python tests/prof_ar.py 2>&1 | tail -2
for cfg in ""; do
  echo "== bench $cfg"
  timeout 300 python bench.py --no-cpu-baseline $cfg 2>gpurun_out/bench.err | tee gpurun_out/bench_last.json | python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()); continue
    print(d['value'], 'img/s', d['ms_per_step'], 'ms/step e2e', d['e2e']['value'], 'launches', d['gpu_launches'], d['roofline']['by_category_ms'], 'gemm TF', round(d['roofline']['achieved'],1), d['latency_bs1'])
"
  tail -3 gpurun_out/bench.err
done
