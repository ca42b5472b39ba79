#!/bin/bash
# forward row kernel after the general-path / register-cap changes: every bench workload, both distributions; fused A/B
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or random_problems or variant or finite or extreme or fused" 2>&1 | tail -2
timeout 400 python scripts/bwd_modes.py --modes 1 2>&1 | grep -o '^[a-z_0-9]* \|"\(fwd_variant[05]\)": {[^}]*}' | cut -c1-120
echo "== detector"
timeout 400 python scripts/bwd_modes.py --workloads detr_encoder_800x1333,detr_encoder_800x1333_bf16 --modes 1 --loc-dist detector 2>&1 | grep -o '"\(fwd_variant[05]\)": {[^}]*}' | cut -c1-120
timeout 300 python scripts/fused_ab.py --out gpurun_out/r2_fused_ab3.json | cut -c1-700
