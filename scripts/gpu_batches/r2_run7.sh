#!/bin/bash
# round-2 GPU batch 7: opaque-base addressing (one IMAD.WIDE per tap) -- timing of every hot kernel, parity, bench
set -x
mkdir -p gpurun_out
timeout 300 python scripts/staged_ab.py --reps 2 2>&1 | tail -13 | head -5 | cut -c1-200
timeout 400 python scripts/bwd_modes.py --modes 1 --out gpurun_out/r2_modes_after_addr.json 2>&1 | tail -8 | cut -c1-520
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 900 python bench.py > gpurun_out/r2_bench_n1_d.json 2> gpurun_out/r2_bench_n1_d.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2_bench_n1_d.json'))
print(d['value'], d['ms_per_step'], d['hbm_frac_step'], d['kernels'])
print('fwd', d['roofline_fwd']['avg_launch_ms'], d['roofline_fwd']['min_launch_ms'], 'bwd', d['roofline']['avg_launch_ms'], d['roofline']['frac'])
print(d['extras']['module_fwd_bwd_ms'])
PY
