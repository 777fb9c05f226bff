#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 300 -k "host_entry or graph_replay or full_size" 2>&1 | tail -6
timeout 900 python bench.py --no-configs --no-cpu-baseline 2>gpurun_out/bench.err | tee gpurun_out/r2o_bench.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], 'e2e', d['e2e'], 'u8', d['e2e_u8']['value'])"
tail -3 gpurun_out/bench.err
