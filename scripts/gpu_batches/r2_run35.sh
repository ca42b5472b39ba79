#!/bin/bash
# final default (planes backward, 768-thread CTAs): whole GPU suite, bench line of record, ncu launch list + full capture
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_n1_d.json 2> gpurun_out/r2_bench_n1_d.err; tail -c 200 gpurun_out/r2_bench_n1_d.err; cut -c1-300 gpurun_out/r2_bench_n1_d.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref_d.json 2>/dev/null; cut -c1-600 gpurun_out/r2_bench_ref_d.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench_d.csv python bench.py --steps 2 --warmup 1 --no-extras --no-e2e --no-cpu-baseline > gpurun_out/r2_d_ncu.log 2>&1 || tail -3 gpurun_out/r2_d_ncu.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"msda_(fwd_v5|bwd_planes)" -c 2 -f -o gpurun_out/r2_prof_bench_d python bench.py --steps 1 --warmup 0 --no-extras --no-e2e --no-cpu-baseline > gpurun_out/r2_d_ncu2.log 2>&1 || tail -3 gpurun_out/r2_d_ncu2.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
