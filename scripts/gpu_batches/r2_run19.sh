#!/bin/bash
# forward row kernel: general path with predicated FMAs (no zero-fill), 40-register cap
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or random_problems or variant or finite or extreme" 2>&1 | tail -2
timeout 400 python scripts/bwd_modes.py --workloads detr_encoder_800x1333,grit_encoder_384x640,detr_encoder_800x1333_bf16,grit_decoder_800x1333_f32 --modes 1 2>&1 | grep -o '"\(fwd_variant[05]\)": {[^}]*}' | cut -c1-200
echo "== detector"
timeout 400 python scripts/bwd_modes.py --workloads detr_encoder_800x1333 --modes 1 --loc-dist detector 2>&1 | grep -o '"\(fwd_variant[05]\)": {[^}]*}' | cut -c1-200
timeout 300 python scripts/fused_ab.py | cut -c1-700
