#!/bin/bash
# planes backward, small CTAs with right-sized shared memory: dense workloads, both distributions
mkdir -p gpurun_out
timeout 600 python scripts/bwd_modes.py --workloads detr_encoder_800x1333,grit_encoder_384x640,detr_encoder_800x1333_bf16 --modes 1,4 --skip-fwd --out gpurun_out/r2_planes_small.json 2>&1 | grep -o '^[a-z_0-9]* \|"bwd_\(mode[14]\)": {[^}]*}' | cut -c1-150
echo "== bf16 t768 / t512"
timeout 400 python scripts/bwd_modes.py --workloads detr_encoder_800x1333_bf16 --modes 4 --skip-fwd --tuning planes_threads=768 2>&1 | grep -o '"bwd_mode4": {[^}]*}' | cut -c1-110
echo "== detector"
timeout 400 python scripts/bwd_modes.py --workloads detr_encoder_800x1333,grit_encoder_384x640,detr_encoder_800x1333_bf16 --modes 1,4 --skip-fwd --loc-dist detector 2>&1 | grep -o '"bwd_mode[14]": {[^}]*}' | cut -c1-150
