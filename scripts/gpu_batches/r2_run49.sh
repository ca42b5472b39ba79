#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:msda_fwd_fused -c 1 -f -o gpurun_out/r2_fwd_fused_full python scripts/fused_ab.py --iters 1 > gpurun_out/r2_ncu_fused.log 2>&1
tail -2 gpurun_out/r2_ncu_fused.log | cut -c1-200
