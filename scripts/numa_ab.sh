#!/bin/bash
# GPU (2 ranks): A/B the NUMA binding of bench.py's e2e leg on the same box.
for bind in 0 1 0 1; do
  MSDA_BENCH_NUMA_BIND=$bind timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2952$bind bench.py --gpus 2 --steps 3 --warmup 3 --no-extras 2>/dev/null | tail -1 > /tmp/line.json
  python - "$bind" <<'PY'
import json, sys
d = json.load(open("/tmp/line.json"))
print("numa_bind", sys.argv[1], "e2e Mq/s", round(d["e2e"]["value"] / 1e6, 2), "ms/step", round(d["e2e"]["ms_per_step"], 1), d["e2e"].get("numa"))
PY
done
nvidia-smi topo -m | head -8
