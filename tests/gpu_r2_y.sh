#!/bin/bash
mkdir -p gpurun_out
for f in 3 1 2; do
  timeout 600 python bench.py --fuse-ln $f --no-configs --no-cpu-baseline --no-latency --no-parity 2>gpurun_out/bench.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('fuse_ln', $f, round(d['value']), round(d['ms_per_step'],3), d['roofline']['by_category_ms'])"
done
