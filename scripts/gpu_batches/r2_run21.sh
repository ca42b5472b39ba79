#!/bin/bash
# same-box A/B of three builds of the forward row kernel's general path: HEAD (zero-fill), predicated tap blocks, predicated loads then FMAs
mkdir -p gpurun_out
for rep in 1 2; do
for lib in build/libmsda_head.so build/libmsda_tapif.so grit_b200/libmsda_b200.so; do
  echo "== $lib"
  GRIT_B200_LIB=$PWD/$lib timeout 300 python scripts/bwd_modes.py --workloads detr_encoder_800x1333,detr_encoder_800x1333_bf16,grit_decoder_384x640_bf16,grit_decoder_800x1333_f32 --modes 1 --iters 20 2>&1 | grep -o '"\(fwd_variant5\)": {[^}]*}' | cut -c1-90
  GRIT_B200_LIB=$PWD/$lib timeout 300 python scripts/bwd_modes.py --workloads detr_encoder_800x1333 --modes 1 --iters 20 --loc-dist detector 2>&1 | grep -o '"\(fwd_variant5\)": {[^}]*}' | cut -c1-90
done
done
