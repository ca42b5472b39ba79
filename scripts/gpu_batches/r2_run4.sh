#!/bin/bash
# round-2 GPU batch 4: full GPU suite, sm_100 gather-path microbenchmark, L2 window experiment, decoder bench after the baddbmm hoist
set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -12 > gpurun_out/r2_pytest_all_b.log
cat gpurun_out/r2_pytest_all_b.log
timeout 120 build/sm100_gather_paths > gpurun_out/r2_micro_gather_paths.txt 2>&1
cat gpurun_out/r2_micro_gather_paths.txt
timeout 300 python scripts/l2_window_experiment.py --out gpurun_out/r2_l2_window.json 2>&1 | tail -8
for py in 384x640 800x1333; do
  timeout 600 python scripts/decoder_bench.py --pyramid $py --out gpurun_out/r2_decoder_bench_$py.json 2>&1 | tail -3 | cut -c1-400
done
timeout 300 python scripts/bwd_modes.py --workloads grit_decoder_384x640_bf16,grit_decoder_800x1333_bf16,grit_decoder_800x1333_f32 --out gpurun_out/r2_bwd_modes_decoder.json 2>&1 | tail -3 | cut -c1-900
