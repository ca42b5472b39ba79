#!/bin/bash
# round-2 GPU batch 3: decoder-layer tests, GRIT operating point (decoder bench + reference CUDA kernels beside), bench line
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_decoder.py -x -q -m gpu 2>&1 | tail -30 > gpurun_out/r2_pytest_decoder.log
cat gpurun_out/r2_pytest_decoder.log
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -8 > gpurun_out/r2_pytest_parity.log
cat gpurun_out/r2_pytest_parity.log
for py in 384x640 800x1333; do
  timeout 600 python scripts/decoder_bench.py --pyramid $py --out gpurun_out/r2_decoder_bench_$py.json 2>&1 | tail -4
done
for wl in grit_decoder_384x640_f32 grit_decoder_800x1333_f32; do
  for n in 4 16 64; do
    timeout 300 python scripts/ref_cuda_bench.py --workload $wl --batch $n 2>&1 | tail -1 | cut -c1-600
  done
done
for wl in detr_encoder_800x1333 grit_encoder_384x640; do
  timeout 300 python scripts/ref_cuda_bench.py --workload $wl 2>&1 | tail -1 | cut -c1-600
done
timeout 900 python bench.py > gpurun_out/r2_bench_n1_a.json 2> gpurun_out/r2_bench_n1_a.err
tail -c 3000 gpurun_out/r2_bench_n1_a.json; tail -5 gpurun_out/r2_bench_n1_a.err
