#!/bin/bash
# planes backward with wave-aligned item counts + streaming hints: A/B in bench context on all dense workloads, then suite + bench of record + ncu
mkdir -p gpurun_out
for w in detr_encoder_800x1333 grit_encoder_384x640 detr_encoder_800x1333_bf16; do
for t in "planes_auto=0" "planes_auto=1"; do
  echo "== $w $t"
  python bench.py --workload $w --steps 20 --warmup 3 --no-extras --no-e2e --no-cpu-baseline --tuning $t 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(round(d['ms_per_step'],3), 'fwd', round(d['roofline_fwd']['avg_launch_ms'],4), d['roofline']['kernel'], round(d['roofline']['avg_launch_ms'],4), 'min', round(d['roofline']['min_launch_ms'],4), d['clocks']['sm_mhz'], d['clocks']['power_w_max'])"
done
done
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_n1_e.json 2> gpurun_out/r2_bench_n1_e.err; tail -c 200 gpurun_out/r2_bench_n1_e.err; cut -c1-300 gpurun_out/r2_bench_n1_e.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench_e.csv python bench.py --steps 2 --warmup 1 --no-extras --no-e2e --no-cpu-baseline > gpurun_out/r2_e_ncu.log 2>&1 || tail -3 gpurun_out/r2_e_ncu.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"msda_(fwd_v5|bwd_planes)" -c 2 -f -o gpurun_out/r2_prof_bench_e python bench.py --steps 1 --warmup 0 --no-extras --no-e2e --no-cpu-baseline > gpurun_out/r2_e_ncu2.log 2>&1 || tail -3 gpurun_out/r2_e_ncu2.log
