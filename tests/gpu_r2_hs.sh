#!/bin/bash
# head-split cross-attention of the AR kernel for tiny batches (template flag): phase stamps at bs=1 / 512, AR tests, latency
mkdir -p gpurun_out
for b in 1 512; do timeout 200 python tests/prof_ar.py $b 2 > gpurun_out/r2h_prof_ar_bs$b.txt 2>&1; tail -5 gpurun_out/r2h_prof_ar_bs$b.txt; done
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "ar_ or invariant or golden" > gpurun_out/r2h_tests.txt 2>&1; echo "tests rc=$?" >> gpurun_out/r2h_tests.txt
tail -4 gpurun_out/r2h_tests.txt
timeout 300 python bench.py --steps 20 --warmup 5 --no-configs --no-cpu-baseline --no-parity --no-two-in-flight 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench', round(d['value']), round(d['ms_per_step'], 3), {k: round(v, 3) for k, v in d['roofline']['by_category_ms'].items()}, d.get('latency_bs1'), d.get('latency_bs1_module'))" | tee gpurun_out/r2h_step.txt
