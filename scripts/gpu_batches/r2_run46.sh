#!/bin/bash
mkdir -p gpurun_out
for a in "--shape 384x640" "--shape 800x1333 --dtype bf16 --n 32" "--shape 384x640 --dtype bf16"; do
  echo "== $a"
  timeout 300 python scripts/fused_ab.py $a | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print({k: round(v,4) for k,v in d.items() if k.endswith('_ms')})"
done
