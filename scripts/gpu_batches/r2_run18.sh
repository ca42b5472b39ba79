#!/bin/bash
# fused module kernels without IEEE divisions, softmax max by redux.sync: parity + A/B against the plain kernels
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "fused or module or decoder or layer" 2>&1 | tail -3
timeout 300 python scripts/fused_ab.py --out gpurun_out/r2_fused_ab2.json
timeout 300 python scripts/module_bench.py 2>&1 | tail -5 | cut -c1-400
