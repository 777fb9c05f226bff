#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -k "uint8 or postprocess or module_api or full_size" 2>&1 | tail -15
timeout 600 python bench.py --no-cpu-baseline --no-latency 2>gpurun_out/bench.err | python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()); continue
    print(d['value'], 'img/s', d['ms_per_step'], 'ms/step | e2e', d['e2e']['value'], '| e2e_u8', d['e2e_u8']['value'])
"
tail -3 gpurun_out/bench.err
