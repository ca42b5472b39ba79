"""Summarise an .ncu-rep (from `ncu --set full`) into the handful of numbers the design discussion uses.

    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep [> profiles/rNN_name.txt]
"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__cycles_elapsed.avg", "SM cycles"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("lts__t_sectors_srcunit_tex_op_read.sum", "L2 sectors read by SMs"),
    ("lts__t_sectors_srcunit_tex_op_red.sum", "L2 sectors reduced (red) by SMs"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX throughput %"),
    ("l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed", "L1->XBAR request path busy %"),
    ("l1tex__m_xbar2l1tex_read_sectors.sum", "XBAR->L1 sectors"),
    ("l1tex__data_pipe_lsu_wavefronts.sum", "LSU data wavefronts"),
    ("l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed", "LSU writeback busy %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__inst_executed_pipe_lsu.sum", "LSU instructions"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_throttle / issue"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_throttle / issue"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait / issue"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected / issue"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier / issue"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio_throttle / issue"),
]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    print(f"# {path}: {len(rows) - 2} kernel launch(es); ncu --set full --clock-control none")
    for r in rows[2:]:
        print()
        print("kernel:", r[hdr.index("Kernel Name")])
        for key, label in KEYS:
            if key in hdr:
                i = hdr.index(key)
                print(f"  {label:38s} {r[i]:>18s} {units[i]:12s} [{key}]")


if __name__ == "__main__":
    main(sys.argv[1])
