#!/bin/bash
# sanitizers over the kernels added since batch 6: planes backward (shared-memory atomics, barriers), fused kernels (redux.sync), forward general path
mkdir -p gpurun_out
K="planes or variant or aggregating or fused or golden or finite"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "$K" > gpurun_out/r2_sanitizer_memcheck_b.txt 2>&1
tail -4 gpurun_out/r2_sanitizer_memcheck_b.txt
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "planes or variant or aggregating" > gpurun_out/r2_sanitizer_racecheck_b.txt 2>&1
tail -4 gpurun_out/r2_sanitizer_racecheck_b.txt
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 1 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "planes or fused" > gpurun_out/r2_sanitizer_synccheck_b.txt 2>&1
tail -4 gpurun_out/r2_sanitizer_synccheck_b.txt
