#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
P='import sys,json; d=json.loads([l for l in sys.stdin if l.startswith("{")][-1]); print(round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"],4), d["roofline"]["families"].get("attention"), d["batch_latency"]["ms_median"])'
timeout 300 python -m pytest tests/test_gpu_engine.py -m gpu -x -q -k attention 2>&1 | tail -3
TMAE_ATTN_TIMING=1 python scripts/attn_timing.py 64 65 12 2>&1 | tail -3 | cut -c1-330
for r in 1 2; do
echo "=== B64 s4 duo (default)"; timeout 400 python bench.py --steps 200 --no-cpu-baseline 2>&1 | python -c "$P"
echo "=== B64 s4 mma.sync"; TMAE_NO_TC_ATTN=1 timeout 400 python bench.py --steps 200 --no-cpu-baseline 2>&1 | python -c "$P"
done
echo "=== B64 s1 duo"; timeout 400 python bench.py --steps 200 --streams 1 --no-cpu-baseline 2>&1 | python -c "$P"
echo "=== KODAK24_K64 duo"; timeout 400 python bench.py --workload KODAK24_K64 --steps 200 --no-cpu-baseline 2>&1 | python -c "$P"
