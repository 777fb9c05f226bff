#!/bin/bash
# late round 2: sanitizer over the new kernels; bench with the two-in-flight extra
mkdir -p gpurun_out
for tool in memcheck synccheck; do
  echo "== compute-sanitizer --tool $tool python tests/sanitize_new_kernels.py" | tee -a gpurun_out/r2_compute_sanitizer_new_kernels.txt
  timeout 600 compute-sanitizer --tool $tool python tests/sanitize_new_kernels.py 2>&1 | grep -E "ERROR SUMMARY|Error|error|hazard|ok:|done|Traceback" | head -20 | tee -a gpurun_out/r2_compute_sanitizer_new_kernels.txt
done
timeout 400 python bench.py --steps 20 --warmup 5 --no-configs --no-cpu-baseline --no-latency > gpurun_out/r2_bench_two_in_flight.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err
python -c "
import json
d = json.loads(open('gpurun_out/r2_bench_two_in_flight.json').read().strip().splitlines()[-1])
print('value', round(d['value']), 'e2e', round(d['e2e']['value']), 'u8', round(d['e2e_u8']['value']), 'two in flight', d['e2e_two_in_flight'])"
