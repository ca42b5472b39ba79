#!/bin/bash
# bench line of record candidate + ncu launch list of the same command + full capture of the two bench kernels
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_n1_b.json 2> gpurun_out/r2_bench_n1_b.err; tail -c 600 gpurun_out/r2_bench_n1_b.err; cut -c1-1500 gpurun_out/r2_bench_n1_b.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench_b.csv python bench.py --steps 2 --warmup 1 --no-extras --no-e2e --no-cpu-baseline > gpurun_out/r2_b_ncu.log 2>&1 || tail -3 gpurun_out/r2_b_ncu.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"msda_(fwd|bwd)_v5" -c 2 -f -o gpurun_out/r2_prof_bench_b python bench.py --steps 1 --warmup 0 --no-extras --no-e2e --no-cpu-baseline > gpurun_out/r2_b_ncu2.log 2>&1 || tail -3 gpurun_out/r2_b_ncu2.log
ls -la gpurun_out | tail -5
