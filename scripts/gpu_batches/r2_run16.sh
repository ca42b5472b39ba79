#!/bin/bash
# (a) forward row kernel with the per-iteration fast-path vote in its general path; (b) planes backward on the other shapes
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or random_problems or variant" 2>&1 | tail -2
timeout 400 python scripts/bwd_modes.py --workloads detr_encoder_800x1333,grit_encoder_384x640,detr_encoder_800x1333_bf16 --modes 1,4 --out gpurun_out/r2_run16_modes.json 2>&1 | grep -o '"\(fwd_variant[05]\|bwd_mode[14]\)": {[^}]*}' | cut -c1-200
echo "== t768"
timeout 400 python scripts/bwd_modes.py --workloads grit_encoder_384x640,detr_encoder_800x1333_bf16 --modes 4 --skip-fwd --tuning planes_threads=768 2>&1 | grep -o '"bwd_mode4": {[^}]*}' | cut -c1-200
echo "== no value planes"
timeout 400 python scripts/bwd_modes.py --workloads detr_encoder_800x1333,grit_encoder_384x640 --modes 4 --skip-fwd --tuning planes_value=-1 2>&1 | grep -o '"bwd_mode4": {[^}]*}' | cut -c1-200
echo "== detector"
timeout 400 python scripts/bwd_modes.py --workloads detr_encoder_800x1333 --modes 1,4 --loc-dist detector 2>&1 | grep -o '"\(fwd_variant[05]\|bwd_mode[14]\)": {[^}]*}' | cut -c1-200
