#!/bin/bash
# full ncu capture of the forward row kernel after the general-path change (source-level instruction mix)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:msda_fwd_v5 -c 1 -f -o gpurun_out/r2_fwd_full python scripts/bwd_modes.py --workloads detr_encoder_800x1333 --iters 1 --modes 1 > gpurun_out/r2_ncu_fwd.log 2>&1
tail -2 gpurun_out/r2_ncu_fwd.log | cut -c1-200
