#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2_bench_extras.json 2>gpurun_out/r2_bench_extras.err; tail -c 300 gpurun_out/r2_bench_extras.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_extras.json').read().strip().splitlines()[-1])
for k,v in d['extras'].items():
    if '_N' in k: print(k, json.dumps(v)[:700])
PY
