#!/bin/bash
# planes backward, final form: CTA-size sweep on the three dense shapes, then the whole GPU suite
mkdir -p gpurun_out
for th in 512 768 1024; do
  echo "== planes_threads=$th"
  timeout 400 python scripts/bwd_modes.py --workloads detr_encoder_800x1333,grit_encoder_384x640,detr_encoder_800x1333_bf16 --modes 4 --skip-fwd --tuning planes_threads=$th 2>&1 | grep -o '"bwd_mode4": {[^}]*}' | cut -c1-160
done
timeout 400 python scripts/bwd_modes.py --modes 1,4 --out gpurun_out/r2_planes_modes.json 2>&1 | cut -c1-60
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
