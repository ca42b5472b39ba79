#!/bin/bash
# fused forward with / without the min-blocks register cap (same-box A/B of two library builds)
mkdir -p gpurun_out
for rep in 1 2; do
for lib in grit_b200/libmsda_b200.so build/libmsda_alt.so; do
  echo "== $lib"
  for a in "--shape 800x1333" "--shape 800x1333 --dtype bf16 --n 32" "--shape 384x640"; do
  GRIT_B200_LIB=$PWD/$lib timeout 300 python scripts/fused_ab.py $a --iters 20 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print({k: round(v,4) for k,v in d.items() if k.startswith('fwd') and k.endswith('_ms')})"
  done
done
done
