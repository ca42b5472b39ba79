#!/bin/bash
# planes backward: item length vs per-item bubbles under wave alignment (chunks = 18 / 37 / 74 per image)
mkdir -p gpurun_out
W=detr_encoder_800x1333
for t in "planes_rows=1235" "planes_rows=601" "planes_rows=301"; do
  echo "== $t"
  timeout 200 python scripts/bwd_modes.py --workloads $W --modes 4 --skip-fwd --iters 20 --tuning $t 2>&1 | grep -o '"bwd_mode4": {[^}]*}' | cut -c1-110
done
