#!/bin/bash
# planes backward v2 (per-iteration fast-path vote, branch-free general path): parity, timing, instruction count
mkdir -p gpurun_out
W=detr_encoder_800x1333
timeout 600 python -m pytest tests -m gpu -x -q -k "planes" 2>&1 | tail -3
timeout 200 python scripts/bwd_modes.py --workloads $W,grit_encoder_384x640 --modes 1,4 --skip-fwd 2>&1 | grep -o '"bwd_mode[14]": {[^}]*}'
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum --clock-control none -k regex:msda_bwd_planes -c 1 python scripts/bwd_modes.py --workloads $W --iters 1 --modes 4 --skip-fwd 2>&1 | grep -E "^\s+(gpu__|smsp__|l1tex__|dram__)"
