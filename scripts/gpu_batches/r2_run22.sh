#!/bin/bash
# same-box A/B: bf16 general path zero-fill (build/libmsda_tapif.so) vs predicated tap blocks (in-tree); then fused A/B, parity
mkdir -p gpurun_out
for rep in 1 2; do
for lib in build/libmsda_tapif.so grit_b200/libmsda_b200.so; do
  echo "== $lib"
  GRIT_B200_LIB=$PWD/$lib timeout 300 python scripts/bwd_modes.py --workloads detr_encoder_800x1333,detr_encoder_800x1333_bf16,grit_decoder_384x640_bf16,grit_decoder_800x1333_bf16 --modes 1 --iters 20 2>&1 | grep -o '"\(fwd_variant5\)": {[^}]*}' | cut -c1-90
  GRIT_B200_LIB=$PWD/$lib timeout 300 python scripts/bwd_modes.py --workloads detr_encoder_800x1333_bf16 --modes 1 --iters 20 --loc-dist detector 2>&1 | grep -o '"\(fwd_variant5\)": {[^}]*}' | cut -c1-90
done
done
timeout 300 python scripts/fused_ab.py --out gpurun_out/r2_fused_ab3.json | cut -c1-700
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
