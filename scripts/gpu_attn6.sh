#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
P='import sys,json; d=json.loads([l for l in sys.stdin if l.startswith("{")][-1]); print(round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"],4), d["roofline"]["families"].get("attention"))'
timeout 300 python -m pytest tests/test_gpu_engine.py -m gpu -x -q -k attention 2>&1 | tail -5
TMAE_ATTN_TIMING=1 python scripts/attn_timing.py 64 65 12 2>&1 | tail -6
TMAE_ATTN_TIMING=1 python scripts/attn_timing.py 64 145 12 2>&1 | tail -4
TMAE_ATTN_TIMING=1 python scripts/attn_timing.py 32 257 16 2>&1 | tail -4
for WL in B64 B144 L256; do
echo "=== $WL tc s4"; timeout 400 python bench.py --workload $WL --steps 150 --no-cpu-baseline 2>&1 | python -c "$P"
echo "=== $WL mma s4"; TMAE_NO_TC_ATTN=1 timeout 400 python bench.py --workload $WL --steps 150 --no-cpu-baseline 2>&1 | python -c "$P"
done
