#!/bin/bash
mkdir -p gpurun_out
timeout 400 python tests/diag_ar_clusters.py 2>&1 | tail -20 | tee gpurun_out/r2f_clusters.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 600 -k "graph_replay or full_size or super_chunks or ar_loop or batch_invariant" 2>&1 | tail -8 | tee gpurun_out/r2f_tests.txt
