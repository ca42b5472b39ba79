#!/bin/bash
# what bounds msda_bwd_planes<t768>: timing with the shared-memory adds removed (1), the fine-level reds removed (2), both (3)
mkdir -p gpurun_out
W=detr_encoder_800x1333
for t in "planes_dbg_skip=0" "planes_dbg_skip=1" "planes_dbg_skip=2" "planes_dbg_skip=3"; do
  echo "== $t"
  timeout 200 python scripts/bwd_modes.py --workloads $W --modes 4 --skip-fwd --iters 20 --tuning $t 2>&1 | grep -o '"bwd_mode4": {[^}]*}' | cut -c1-80
done
