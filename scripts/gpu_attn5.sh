#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
P='import sys,json; d=json.loads([l for l in sys.stdin if l.startswith("{")][-1]); print(round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"],4), d["roofline"]["families"].get("attention"))'
for r in 1 2; do
echo "=== tc s1 steps 300"; timeout 400 python bench.py --steps 300 --streams 1 --no-cpu-baseline 2>&1 | python -c "$P"
echo "=== mma s1 steps 300"; TMAE_NO_TC_ATTN=1 timeout 400 python bench.py --steps 300 --streams 1 --no-cpu-baseline 2>&1 | python -c "$P"
done
echo "=== tc s4 steps 300"; timeout 400 python bench.py --steps 300 --no-cpu-baseline 2>&1 | python -c "$P"
echo "=== mma s4 steps 300"; TMAE_NO_TC_ATTN=1 timeout 400 python bench.py --steps 300 --no-cpu-baseline 2>&1 | python -c "$P"
echo "=== tc NO_GRAPH steps 100"; TMAE_NO_GRAPH=1 timeout 400 python bench.py --steps 100 --streams 1 --no-cpu-baseline 2>&1 | python -c "$P"
echo "=== mma NO_GRAPH steps 100"; TMAE_NO_TC_ATTN=1 TMAE_NO_GRAPH=1 timeout 400 python bench.py --steps 100 --streams 1 --no-cpu-baseline 2>&1 | python -c "$P"
