#!/bin/bash
# one-kernel MLP + LayerNorm: bit-identity test (timeouts: a pipeline deadlock must not hang the box), microbenchmark, whole step
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_kernels.py -q -x -m gpu -k "mlp_ln" > gpurun_out/r2m_tests.txt 2>&1; echo "tests rc=$?" >> gpurun_out/r2m_tests.txt
tail -25 gpurun_out/r2m_tests.txt
if grep -q "tests rc=0" gpurun_out/r2m_tests.txt; then
  timeout 200 python tests/bench_mlp_ln.py > gpurun_out/r2m_mlp_ln.txt 2>&1; cat gpurun_out/r2m_mlp_ln.txt
  for v in "0 0" "1 0" "1 1"; do set -- $v
    timeout 300 python bench.py --steps 20 --warmup 5 --no-configs --no-cpu-baseline --no-latency --fuse-mlp $1 --pair-pdl $2 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('fuse_mlp=$1 pair_pdl=$2', round(d['value']), round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value']), {k: round(v, 3) for k, v in d['roofline']['by_category_ms'].items()}, (d.get('parity') or {}).get('ok'))" | tee -a gpurun_out/r2m_step.txt
  done
fi
