#!/bin/bash
# round-2 GPU batch 2: profile the binned backward (why is it slow?), re-time after the uniform-branch skip, full GPU suite
set -x
mkdir -p gpurun_out
timeout 600 python scripts/bwd_modes.py --workloads detr_encoder_800x1333,grit_encoder_384x640 --out gpurun_out/r2_bwd_modes_b.json 2>&1 | tail -5 > gpurun_out/r2_bwd_modes_b.log
cat gpurun_out/r2_bwd_modes_b.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:msda_bwd -c 2 -f -o gpurun_out/r2_binned_full python scripts/bwd_modes.py --workloads detr_encoder_800x1333 --iters 1 --modes 2 > gpurun_out/r2_ncu_binned.log 2>&1
tail -3 gpurun_out/r2_ncu_binned.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r2_launches_bwd_modes_b.csv python scripts/bwd_modes.py --workloads detr_encoder_800x1333 --iters 1 > gpurun_out/r2_ncu_modes_b.log 2>&1
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 > gpurun_out/r2_pytest_all.log
cat gpurun_out/r2_pytest_all.log
