#!/bin/bash
# bench-context A/B on the other dense workloads: row vs planes (256-thread CTAs x4 per SM) vs planes (one 768-thread CTA per SM)
mkdir -p gpurun_out
for w in grit_encoder_384x640 detr_encoder_800x1333_bf16; do
for t in "planes_auto=0" "planes_auto=1" "planes_auto=1,planes_threads=768"; do
  echo "== $w $t"
  python bench.py --workload $w --steps 20 --warmup 3 --no-extras --no-e2e --no-cpu-baseline --tuning $t 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(round(d['ms_per_step'],3), 'fwd', round(d['roofline_fwd']['avg_launch_ms'],4), d['roofline']['kernel'], round(d['roofline']['avg_launch_ms'],4), 'min', round(d['roofline']['min_launch_ms'],4), d['clocks']['sm_mhz'], d['clocks']['power_w_max'])"
done
done
echo "== detector 800x1333"
for t in "planes_auto=0" "planes_auto=1" "planes_auto=1,planes_threads=768"; do
  echo "== $t"
  python bench.py --loc-dist detector --steps 20 --warmup 3 --no-extras --no-e2e --no-cpu-baseline --tuning $t 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(round(d['ms_per_step'],3), 'fwd', round(d['roofline_fwd']['avg_launch_ms'],4), d['roofline']['kernel'], round(d['roofline']['avg_launch_ms'],4), 'min', round(d['roofline']['min_launch_ms'],4), d['clocks']['sm_mhz'], d['clocks']['power_w_max'])"
done
