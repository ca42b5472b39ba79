#!/bin/bash
# GPU: sweep the host-buffer pipeline depth (MSDA_HOST_SLOTS) and chunk size of bench.py's e2e leg.
for slots in 3 4 6; do
  for chunk in 1 2 4; do
    MSDA_HOST_SLOTS=$slots timeout 200 python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline --e2e-chunk $chunk 2>/dev/null | tail -1 > /tmp/e2e_line.json
    python - "$slots" "$chunk" <<'PY'
import json, sys
d = json.load(open("/tmp/e2e_line.json"))
print("slots", sys.argv[1], "chunk", sys.argv[2], round(d["e2e"]["value"] / 1e6, 2), "Mq/s", round(d["e2e"]["ms_per_step"], 1), "ms/step")
PY
  done
done
