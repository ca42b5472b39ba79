#!/bin/bash
# end-of-round validation: whole GPU suite, smoke, bench line of record, reference arm
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_n1_g.json 2> gpurun_out/r2_bench_n1_g.err; tail -c 200 gpurun_out/r2_bench_n1_g.err
python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_n1_g.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['kernel'], d['roofline']['avg_launch_ms'], d['roofline']['frac'], d['roofline_fwd']['avg_launch_ms'], d['roofline_fwd']['frac'], d['clocks'])
print(d['e2e']['value'], d['e2e']['frac_of_copy_ceiling'], d['cpu_baseline']['value'], d['gpu_launches'])
for k in ('grit_encoder_384x640','detr_encoder_800x1333_bf16','grit_decoder_384x640_bf16','module_fwd_bwd_ms','reference_cuda_kernels_on_this_gpu'):
    if k in d['extras']: print(k, json.dumps(d['extras'][k])[:420])
PY
