#!/bin/bash
# upper bound of removing the shape-staging prologue (smem + barrier) from the forward row kernel: hard-coded pyramid
mkdir -p gpurun_out
for rep in 1 2; do
for lib in grit_b200/libmsda_b200.so build/libmsda_hack.so; do
  echo "== $lib"
  GRIT_B200_LIB=$PWD/$lib timeout 300 python scripts/bwd_modes.py --workloads detr_encoder_800x1333 --modes 1 --iters 20 2>&1 | grep -o '"\(fwd_variant5\)": {[^}]*}' | cut -c1-100
done
done
