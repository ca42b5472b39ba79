#!/bin/bash
# round-2 GPU batch 5: lean microbenchmarks, L2 window experiment, GroupNorm-pack test, bench line with the lean probes
set -x
mkdir -p gpurun_out
timeout 120 build/sm100_gather_paths > gpurun_out/r2_micro_gather_paths.txt 2>&1
cat gpurun_out/r2_micro_gather_paths.txt
timeout 600 python -m pytest tests/test_gpu_decoder.py -x -q -m gpu 2>&1 | tail -8
timeout 300 python scripts/l2_window_experiment.py --out gpurun_out/r2_l2_window.json 2>&1 | tail -6 | cut -c1-700
timeout 900 python bench.py > gpurun_out/r2_bench_n1_b.json 2> gpurun_out/r2_bench_n1_b.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2_bench_n1_b.json'))
print(d['value'], d['ms_per_step'], d['hbm_frac_step'])
print(d['roofline'].get('on_chip')); print(d['roofline_fwd'].get('on_chip'))
print(d['e2e']['value'], d['e2e']['frac_of_copy_ceiling'])
PY
