#!/bin/bash
# round-2 GPU batch 8: numbers of record -- bench line, ncu launch list of the same command, full capture of the
# forward and backward of the bench workload, reference CUDA kernels beside ours on every shape
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 900 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
tail -c 400 gpurun_out/r2_bench_n1.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > gpurun_out/r2_bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:msda_ -c 2 -f -o gpurun_out/r2_prof_bench python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > gpurun_out/r2_ncu_full.log 2>&1
tail -2 gpurun_out/r2_ncu_full.log
for wl in grit_decoder_384x640_f32 grit_decoder_800x1333_f32; do
  for n in 4 16 64; do
    timeout 300 python scripts/ref_cuda_bench.py --workload $wl --batch $n 2>&1 | tail -1 | cut -c1-400
  done
done
for wl in detr_encoder_800x1333 grit_encoder_384x640; do
  timeout 300 python scripts/ref_cuda_bench.py --workload $wl 2>&1 | tail -1 | cut -c1-400
done
