#!/bin/bash
# round-2 GPU batch 6: dynamic staged forward (timing + parity), sanitizers over the round-2 kernels
set -x
mkdir -p gpurun_out
timeout 300 python scripts/bwd_modes.py --workloads detr_encoder_800x1333,grit_encoder_384x640,detr_encoder_800x1333_bf16 --modes 1 --out gpurun_out/r2_fwd_variants.json 2>&1 | tail -3 | cut -c1-700
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r2_bench_n1_c.json 2> gpurun_out/r2_bench_n1_c.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2_bench_n1_c.json'))
print(d['value'], d['ms_per_step'], d['hbm_frac_step'], d['kernels'])
print('fwd', d['roofline_fwd']['avg_launch_ms'], d['roofline_fwd']['min_launch_ms'], 'bwd', d['roofline']['avg_launch_ms'])
PY
K="aggregating or variant or add_dropout or groupnorm or decoder_layer or valid_ratio or nonfinite"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_parity.py tests/test_gpu_decoder.py -x -q -m gpu -k "$K" > gpurun_out/r2_sanitizer_memcheck.txt 2>&1
tail -5 gpurun_out/r2_sanitizer_memcheck.txt
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_gpu_parity.py tests/test_gpu_decoder.py -x -q -m gpu -k "aggregating or variant or add_dropout or groupnorm" > gpurun_out/r2_sanitizer_racecheck.txt 2>&1
tail -5 gpurun_out/r2_sanitizer_racecheck.txt
