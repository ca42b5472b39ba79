#!/bin/bash
# same-box A/B, bf16 forward general path: zero-fill (build/libmsda_prev.so) vs predicated raw loads + predicated unpack/FMA (in-tree)
mkdir -p gpurun_out
for rep in 1 2; do
for lib in build/libmsda_prev.so grit_b200/libmsda_b200.so; do
  echo "== $lib"
  GRIT_B200_LIB=$PWD/$lib timeout 300 python scripts/bwd_modes.py --workloads detr_encoder_800x1333_bf16,grit_decoder_384x640_bf16,grit_decoder_800x1333_bf16 --modes 1 --iters 20 2>&1 | grep -o '"\(fwd_variant5\)": {[^}]*}' | cut -c1-100
  GRIT_B200_LIB=$PWD/$lib timeout 300 python scripts/bwd_modes.py --workloads detr_encoder_800x1333_bf16 --modes 1 --iters 20 --loc-dist detector 2>&1 | grep -o '"\(fwd_variant5\)": {[^}]*}' | cut -c1-100
done
done
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or random_problems or variant or finite or extreme" 2>&1 | tail -2
