#!/bin/bash
# round-2 GPU batch 9: smoke(), new tests, register-cap / CTA-size experiments on the row forward
set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "submit_wait or strategy_selection or non_finite or host_session" 2>&1 | tail -5
timeout 300 python scripts/staged_ab.py --reps 2 --configs row 2>&1 | tail -8 | cut -c1-200
