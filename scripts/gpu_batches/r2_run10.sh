#!/bin/bash
# round-2 GPU batch 10: pair-packed bf16 forward -- parity, timing
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "pair_packed or variant or random_problems or golden" 2>&1 | tail -6
timeout 300 python scripts/bwd_modes.py --workloads detr_encoder_800x1333_bf16,grit_encoder_384x640,detr_encoder_800x1333 --modes 1 --out gpurun_out/r2_pairs.json 2>&1 | tail -3 | cut -c1-900
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
