#!/bin/bash
# bench-context A/B (six layers back to back on distinct input sets, sustained clocks): row backward vs planes backward
mkdir -p gpurun_out
for rep in 1 2; do
for t in "planes_auto=0" "planes_auto=1" "planes_auto=1,planes_threads=768"; do
  echo "== $t"
  python bench.py --steps 20 --warmup 3 --no-extras --no-e2e --no-cpu-baseline --tuning $t 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(round(d['ms_per_step'],3), 'fwd', round(d['roofline_fwd']['avg_launch_ms'],4), d['roofline']['kernel'], round(d['roofline']['avg_launch_ms'],4), 'min', round(d['roofline']['min_launch_ms'],4), d['clocks'])"
done
done
