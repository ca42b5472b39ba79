#!/bin/bash
# planes backward CTA size under wave-aligned items: 768 / 896 / 1024 threads, alone and in bench context
mkdir -p gpurun_out
W=detr_encoder_800x1333
for t in "planes_threads=768" "planes_threads=896" "planes_threads=1024"; do
  echo "== $t"
  timeout 200 python scripts/bwd_modes.py --workloads $W,grit_encoder_384x640 --modes 4 --skip-fwd --tuning $t 2>&1 | grep -o '"bwd_mode4": {[^}]*}' | cut -c1-110
  python bench.py --steps 20 --warmup 3 --no-extras --no-e2e --no-cpu-baseline --tuning $t 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(round(d['ms_per_step'],3), 'fwd', round(d['roofline_fwd']['avg_launch_ms'],4), d['roofline']['kernel'], round(d['roofline']['avg_launch_ms'],4), 'min', round(d['roofline']['min_launch_ms'],4), d['clocks']['sm_mhz'], d['clocks']['power_w_max'])"
done
