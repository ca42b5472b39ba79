#!/bin/bash
# same-box A/B of the staged forward's general path (HEAD zero-fill vs predicated tap blocks), then the whole GPU suite
mkdir -p gpurun_out
for rep in 1 2; do
for lib in build/libmsda_head.so grit_b200/libmsda_b200.so; do
  echo "== $lib"
  GRIT_B200_LIB=$PWD/$lib timeout 300 python scripts/bwd_modes.py --workloads grit_encoder_384x640,detr_encoder_800x1333 --modes 1 --iters 20 2>&1 | grep -o '"\(fwd_variant[35]\)": {[^}]*}' | cut -c1-100
  GRIT_B200_LIB=$PWD/$lib timeout 300 python scripts/bwd_modes.py --workloads grit_encoder_384x640 --modes 1 --iters 20 --loc-dist detector 2>&1 | grep -o '"\(fwd_variant[35]\)": {[^}]*}' | cut -c1-100
done
done
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
