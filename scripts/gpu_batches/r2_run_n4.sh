#!/bin/bash
set -x
nvidia-smi topo -m | head -12
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r2_bench_n4.json 2> gpurun_out/r2_bench_n4.err
echo "rc=$?"
tail -c 1500 gpurun_out/r2_bench_n4.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_n4.json").read().strip().split("\n")[-1])
print(d["value"], d["ms_per_step"], d["n_gpus"], d["multi_gpu_check"])
print(d["e2e"])
PY
