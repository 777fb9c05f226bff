#!/bin/bash
# compute-sanitizer over the final build: all decode modes at B=3 (fused residual-GEMM + LayerNorm forced, tcgen05 attention,
# cluster AR kernel incl. the producer warp / DSMEM exchanges), the grid-barrier AR kernel, the decode API.
mkdir -p gpurun_out
for tool in memcheck synccheck racecheck; do
  echo "== compute-sanitizer --tool $tool" | tee -a gpurun_out/r2_compute_sanitizer.txt
  timeout 900 compute-sanitizer --tool $tool python tests/sanitize_small.py 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Error|error|hazard|ok:|Traceback" | head -20 | tee -a gpurun_out/r2_compute_sanitizer.txt
done
