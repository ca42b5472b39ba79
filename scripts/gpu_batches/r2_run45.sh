#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/fused_ab.py --out gpurun_out/r2_fused_ab4.json | cut -c1-1200
python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_n1_f.json 2> gpurun_out/r2_bench_n1_f.err; tail -c 200 gpurun_out/r2_bench_n1_f.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_n1_f.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['kernel'], d['roofline']['avg_launch_ms'], d['roofline_fwd']['avg_launch_ms'], d['clocks'])
print(json.dumps(d['extras']['module_fwd_bwd_ms']))
print(d['e2e']['value'], d['e2e']['frac_of_copy_ceiling'], d['cpu_baseline']['value'])
PY
