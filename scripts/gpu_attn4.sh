#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
P='import sys,json; d=json.loads([l for l in sys.stdin if l.startswith("{")][-1]); print(round(d["value"]), round(d["e2e"]["value"]), d["ms_per_step"], d["roofline"]["families"].get("attention"))'
echo "=== tc default"; timeout 400 python bench.py --steps 20 --streams 1 --no-cpu-baseline 2>&1 | python -c "$P"
echo "=== tc NO_PDL"; TMAE_NO_PDL=1 timeout 400 python bench.py --steps 20 --streams 1 --no-cpu-baseline 2>&1 | python -c "$P"
echo "=== tc NO_GRAPH"; TMAE_NO_GRAPH=1 timeout 400 python bench.py --steps 20 --streams 1 --no-cpu-baseline 2>&1 | python -c "$P"
echo "=== tc NO_GRAPH NO_PDL"; TMAE_NO_GRAPH=1 TMAE_NO_PDL=1 timeout 400 python bench.py --steps 20 --streams 1 --no-cpu-baseline 2>&1 | python -c "$P"
echo "=== mma NO_PDL"; TMAE_NO_TC_ATTN=1 TMAE_NO_PDL=1 timeout 400 python bench.py --steps 20 --streams 1 --no-cpu-baseline 2>&1 | python -c "$P"
echo "=== tc default steps 100"; timeout 400 python bench.py --steps 100 --streams 1 --no-cpu-baseline 2>&1 | python -c "$P"
