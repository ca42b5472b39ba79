#!/bin/bash
# planes backward, small CTAs by default: all dense workloads, both distributions, vs the row kernel and the big-CTA form; parity
mkdir -p gpurun_out
timeout 600 python scripts/bwd_modes.py --workloads detr_encoder_800x1333,grit_encoder_384x640,detr_encoder_800x1333_bf16 --modes 1,4 --skip-fwd --out gpurun_out/r2_planes_small.json 2>&1 | grep -o '^[a-z_0-9]* \|"bwd_\(mode[14]\|auto\)": {[^}]*}' | cut -c1-150
echo "== t768"
timeout 400 python scripts/bwd_modes.py --workloads detr_encoder_800x1333_bf16 --modes 4 --skip-fwd --tuning planes_threads=768 2>&1 | grep -o '"bwd_mode4": {[^}]*}' | cut -c1-110
echo "== detector"
timeout 400 python scripts/bwd_modes.py --workloads detr_encoder_800x1333,grit_encoder_384x640 --modes 1,4 --skip-fwd --loc-dist detector 2>&1 | grep -o '"bwd_mode[14]": {[^}]*}' | cut -c1-150
timeout 900 python -m pytest tests -m gpu -x -q -k "planes or strateg or variant or fullsize or full_size" 2>&1 | tail -3
