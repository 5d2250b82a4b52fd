#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
P='import sys,json; d=json.loads([l for l in sys.stdin if l.startswith("{")][-1]); print(round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"],4), d["gpu_launches"]//d["steps"], round(d["roofline"]["frac"],4), round(d["roofline"]["step_aggregate"]["frac"],4), d["batch_latency"]["ms_median"])'
timeout 900 python -m pytest tests/test_gpu_forward.py tests/test_gpu_entropy.py -m gpu -x -q -k "fused or switch or entropy or compress or sweep" 2>&1 | tail -4
timeout 300 python scripts/profile_steps.py B64 2>&1 | grep -E "cc\.0\.|cc\.6\.8|sum of launches"
for r in 1 2; do
echo "=== B64 fused"; timeout 400 python bench.py --steps 200 --no-cpu-baseline 2>&1 | python -c "$P"
echo "=== B64 unfused"; TMAE_NO_GC_FUSE=1 timeout 400 python bench.py --steps 200 --no-cpu-baseline 2>&1 | python -c "$P"
done
