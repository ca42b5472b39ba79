#!/bin/bash
set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
echo "rc=$?"
wc -c gpurun_out/r2_bench_n2.json gpurun_out/r2_bench_n2.err
tail -c 2500 gpurun_out/r2_bench_n2.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench_n2.json"))
print(d["value"], d["ms_per_step"], d["n_gpus"], d["multi_gpu_check"])
print(d["e2e"])
PY
