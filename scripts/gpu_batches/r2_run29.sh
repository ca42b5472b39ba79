#!/bin/bash
# small-CTA planes backward: structure-only control (budget 0), row counts, ncu metrics
mkdir -p gpurun_out
W=detr_encoder_800x1333
for t in "planes_threads=256,planes_budget=0,planes_rows=256" "planes_threads=256,planes_budget=36000,planes_rows=192" "planes_threads=256,planes_budget=36000,planes_rows=320"; do
  echo "== $t"
  timeout 200 python scripts/bwd_modes.py --workloads $W --modes 4 --skip-fwd --tuning $t 2>&1 | grep -o '"bwd_mode4": {[^}]*}' | cut -c1-110
done
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_sectors_srcunit_tex_op_red.sum,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed
timeout 300 ncu --metrics $M --clock-control none -k regex:msda_bwd_planes -c 1 python scripts/bwd_modes.py --workloads $W --iters 1 --modes 4 --skip-fwd --tuning planes_threads=256,planes_budget=36000,planes_rows=256 2>&1 | grep -E "^\s+(gpu__|smsp__|l1tex__|dram__|sm__|lts__)"
