#!/bin/bash
# planes backward (coarse levels in shared-memory int32 fixed point): parity, then timing against the row kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "planes" 2>&1 | tail -15
timeout 400 python scripts/bwd_modes.py --workloads detr_encoder_800x1333,grit_encoder_384x640,detr_encoder_800x1333_bf16 --modes 1,4 --out gpurun_out/r2_planes_modes.json 2>&1 | tail -4 | cut -c1-1500
