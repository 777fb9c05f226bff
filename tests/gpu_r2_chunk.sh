#!/bin/bash
# encoder chunk size sweep: smaller chunks keep the per-layer intermediates (qkv, hidden) inside the 126 MB L2
mkdir -p gpurun_out
for c in 512 256 192 148 128 96 64; do
  timeout 300 python bench.py --steps 20 --warmup 5 --no-configs --no-parity --no-cpu-baseline --no-latency --chunk $c 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('chunk=$c', round(d['value']), round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value']), {k: round(v, 3) for k, v in d['roofline']['by_category_ms'].items()})" | tee -a gpurun_out/r2_chunk_sweep.txt
done
