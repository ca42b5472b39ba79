#!/bin/bash
# planes backward with SMALL CTAs: only the smallest level(s) on chip, 256-thread CTAs, 4 per SM (occupancy of the row kernel,
# fewer reds) -- sweep of plane budget and rows per item
mkdir -p gpurun_out
W=detr_encoder_800x1333
echo "== row / planes t768 reference"
timeout 200 python scripts/bwd_modes.py --workloads $W,grit_encoder_384x640 --modes 1,4 --skip-fwd 2>&1 | grep -o '"bwd_mode[14]": {[^}]*}' | cut -c1-110
for t in "planes_threads=256,planes_budget=36000,planes_rows=256" "planes_threads=256,planes_budget=36000,planes_rows=512" "planes_threads=256,planes_budget=36000,planes_rows=1024" "planes_threads=256,planes_budget=36000,planes_rows=128"; do
  echo "== $t"
  timeout 200 python scripts/bwd_modes.py --workloads $W --modes 4 --skip-fwd --tuning $t 2>&1 | grep -o '"bwd_mode4": {[^}]*}' | cut -c1-110
done
for t in "planes_threads=256,planes_budget=40000,planes_rows=256" "planes_threads=256,planes_budget=40000,planes_rows=512" "planes_threads=256,planes_budget=8000,planes_rows=256"; do
  echo "== 384x640 $t"
  timeout 200 python scripts/bwd_modes.py --workloads grit_encoder_384x640 --modes 4 --skip-fwd --tuning $t 2>&1 | grep -o '"bwd_mode4": {[^}]*}' | cut -c1-110
done
