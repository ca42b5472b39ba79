#!/bin/bash
# round-2 GPU batch 1: new backward strategies -- parity first, then timings, then a launch list
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "aggregating or variant or nonfinite or fused_deterministic or caller_supplied or validates or reference_test_py" 2>&1 | tail -25 > gpurun_out/r2_pytest_new.log
cat gpurun_out/r2_pytest_new.log
timeout 600 python scripts/bwd_modes.py --out gpurun_out/r2_bwd_modes.json 2>&1 | tail -20 > gpurun_out/r2_bwd_modes.log
cat gpurun_out/r2_bwd_modes.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches_bwd_modes.csv python scripts/bwd_modes.py --workloads detr_encoder_800x1333 --iters 1 > gpurun_out/r2_ncu_modes.log 2>&1
tail -3 gpurun_out/r2_ncu_modes.log
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 > gpurun_out/r2_pytest_all.log
cat gpurun_out/r2_pytest_all.log
