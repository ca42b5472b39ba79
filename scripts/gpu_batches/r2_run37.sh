#!/bin/bash
# planes backward: wave-aligned item counts (items per image a multiple of the CTA slots) -> one image in flight at a time?
mkdir -p gpurun_out
W=detr_encoder_800x1333
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct
for t in "planes_rows=601" "planes_rows=1024" "planes_threads=256,planes_rows=301" "planes_threads=256,planes_rows=256"; do
  echo "== $t"
  timeout 200 python scripts/bwd_modes.py --workloads $W --modes 4 --skip-fwd --tuning $t 2>&1 | grep -o '"bwd_mode4": {[^}]*}' | cut -c1-110
  timeout 300 ncu --metrics $M --clock-control none -k regex:msda_bwd_planes -c 1 python scripts/bwd_modes.py --workloads $W --iters 1 --modes 4 --skip-fwd --tuning $t 2>&1 | grep -E "^\s+(gpu__|dram__|lts__)"
done
for t in "planes_auto=0" "planes_auto=1,planes_rows=601" "planes_auto=1,planes_threads=256,planes_rows=301"; do
  echo "== bench $t"
  python bench.py --steps 20 --warmup 3 --no-extras --no-e2e --no-cpu-baseline --tuning $t 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(round(d['ms_per_step'],3), 'fwd', round(d['roofline_fwd']['avg_launch_ms'],4), d['roofline']['kernel'], round(d['roofline']['avg_launch_ms'],4), 'min', round(d['roofline']['min_launch_ms'],4), d['clocks']['sm_mhz'], d['clocks']['power_w_max'])"
done
