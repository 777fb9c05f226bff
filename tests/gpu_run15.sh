#!/bin/bash
mkdir -p gpurun_out
for cfg in "--dec-chunk 512" "--dec-chunk 256" "--chunk 256" ; do
  echo "== bench $cfg"
  timeout 300 python bench.py --no-cpu-baseline --no-latency $cfg 2>gpurun_out/bench.err | python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()); continue
    print(d['value'], 'img/s', d['ms_per_step'], 'ms/step e2e', d['e2e']['value'], 'launches', d['gpu_launches'], d['roofline']['by_category_ms'], 'gemm TF', round(d['roofline']['achieved'],1))
"
  tail -3 gpurun_out/bench.err
done
timeout 300 python tests/bench_modes.py 2>&1 | tail -2
