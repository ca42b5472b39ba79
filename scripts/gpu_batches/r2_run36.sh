#!/bin/bash
# planes backward with streaming (evict-first) row inputs/outputs: DRAM traffic, alone timing, bench-context A/B, rows per item
mkdir -p gpurun_out
W=detr_encoder_800x1333
timeout 300 python -m pytest tests -m gpu -x -q -k "planes" 2>&1 | tail -2
timeout 200 python scripts/bwd_modes.py --workloads $W,grit_encoder_384x640,detr_encoder_800x1333_bf16 --modes 1,4 --skip-fwd 2>&1 | grep -o '"bwd_mode[14]": {[^}]*}' | cut -c1-110
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:msda_bwd_planes -c 1 python scripts/bwd_modes.py --workloads $W --iters 1 --modes 4 --skip-fwd 2>&1 | grep -E "^\s+(gpu__|dram__|lts__)"
for t in "planes_auto=0" "planes_auto=1" "planes_auto=1,planes_rows=512" "planes_auto=1,planes_threads=256"; do
  echo "== $t"
  python bench.py --steps 20 --warmup 3 --no-extras --no-e2e --no-cpu-baseline --tuning $t 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(round(d['ms_per_step'],3), 'fwd', round(d['roofline_fwd']['avg_launch_ms'],4), d['roofline']['kernel'], round(d['roofline']['avg_launch_ms'],4), 'min', round(d['roofline']['min_launch_ms'],4), d['clocks']['sm_mhz'], d['clocks']['power_w_max'])"
done
