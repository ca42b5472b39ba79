#!/bin/bash
# planes backward v3: item-size / CTA-size sweep, full ncu capture
mkdir -p gpurun_out
W=detr_encoder_800x1333
for t in "planes_rows=256" "planes_rows=512" "planes_rows=2048" "planes_rows=3072" "planes_threads=768" "planes_threads=768,planes_rows=2048"; do
  echo "== $t"
  timeout 200 python scripts/bwd_modes.py --workloads $W --modes 4 --skip-fwd --tuning $t 2>&1 | grep -o '"bwd_mode4": {[^}]*}'
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:msda_bwd_planes -c 1 -f -o gpurun_out/r2_planes_full python scripts/bwd_modes.py --workloads $W --iters 1 --modes 4 --skip-fwd > gpurun_out/r2_ncu_planes.log 2>&1
tail -2 gpurun_out/r2_ncu_planes.log | cut -c1-200
