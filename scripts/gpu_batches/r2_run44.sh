#!/bin/bash
# planes backward with the fused point source: parity, module A/B, whole suite
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "fused or module or non_finite" 2>&1 | tail -4
timeout 300 python scripts/module_bench.py 2>&1 | tail -12 | cut -c1-300
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
