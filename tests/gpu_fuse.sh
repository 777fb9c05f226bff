#!/bin/bash
# Fused residual-GEMM + LayerNorm: unit tests first (bounded waits trap instead of hanging), then the model.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x --timeout 120 -k "gemm_ln" 2>&1 | tail -25 > gpurun_out/fuse_unit.log
tail -25 gpurun_out/fuse_unit.log
if grep -q "failed\|error" gpurun_out/fuse_unit.log; then echo "UNIT FAILED - stopping"; exit 0; fi
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
for f in "--fuse-ln 3" "--fuse-ln 0"; do
echo "== bench $f"
timeout 300 python bench.py --no-cpu-baseline $f 2>gpurun_out/bench.err | tee gpurun_out/bench_last.json | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(d['value'], 'img/s', d['ms_per_step'], 'ms/step | e2e', d['e2e']['value'], '| e2e_u8', d.get('e2e_u8', {}).get('value'), d['roofline']['by_category_ms'])"
tail -1 gpurun_out/bench.err
done
