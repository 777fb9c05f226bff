#!/bin/bash
# Round-end style validation: full GPU test suite, smoke, default bench (tests/gpu_final.sh runs the same + the full bench)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/ -q -m gpu --timeout 900 2>&1 | tail -6 | tee gpurun_out/r2_final_tests.txt
