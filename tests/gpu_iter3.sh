#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 2>&1 | tail -3
timeout 300 python tests/prof_ar.py 2>&1 | tail -5
for i in 1 2; do
timeout 300 python bench.py --no-cpu-baseline 2>gpurun_out/bench.err | tee gpurun_out/bench_last.json | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(d['value'], 'img/s', d['ms_per_step'], 'ms/step | e2e', d['e2e']['value'], '| e2e_u8', d.get('e2e_u8', {}).get('value'), d['roofline']['by_category_ms'])"
done
