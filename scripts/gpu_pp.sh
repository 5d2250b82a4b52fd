#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
P='import sys,json; d=json.loads([l for l in sys.stdin if l.startswith("{")][-1]); print(round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"],4), round(d["roofline"]["frac"],4), round(d["roofline"]["step_aggregate"]["frac"],4), d["batch_latency"]["ms_median"])'
timeout 300 python -m pytest tests/test_gpu_engine.py -m gpu -x -q -k "bf16_out" 2>&1 | tail -8
echo "=== B64 s1 pair-persistent"; timeout 400 python bench.py --steps 100 --streams 1 --no-cpu-baseline 2>&1 | python -c "$P"
echo "=== B64 s1 pairs"; TMAE_NO_PAIR_PERSISTENT=1 timeout 400 python bench.py --steps 100 --streams 1 --no-cpu-baseline 2>&1 | python -c "$P"
timeout 300 python scripts/profile_steps.py B64 2>&1 | grep -E "blk3\.|sum of launches"
