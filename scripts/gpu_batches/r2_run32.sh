#!/bin/bash
# planes backward as the default for dense D=32: whole GPU suite, bench line, ncu launch list + full capture of the bench kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_n1_c.json 2> gpurun_out/r2_bench_n1_c.err; tail -c 300 gpurun_out/r2_bench_n1_c.err; cut -c1-400 gpurun_out/r2_bench_n1_c.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench_c.csv python bench.py --steps 2 --warmup 1 --no-extras --no-e2e --no-cpu-baseline > gpurun_out/r2_c_ncu.log 2>&1 || tail -3 gpurun_out/r2_c_ncu.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"msda_(fwd_v5|bwd_planes)" -c 2 -f -o gpurun_out/r2_prof_bench_c python bench.py --steps 1 --warmup 0 --no-extras --no-e2e --no-cpu-baseline > gpurun_out/r2_c_ncu2.log 2>&1 || tail -3 gpurun_out/r2_c_ncu2.log
