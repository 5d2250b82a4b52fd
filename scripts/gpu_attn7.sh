#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
P='import sys,json; d=json.loads([l for l in sys.stdin if l.startswith("{")][-1]); print(round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"],4), d["roofline"]["families"].get("attention"), d.get("batch_latency"))'
timeout 300 python -m pytest tests/test_gpu_engine.py -m gpu -x -q -k attention 2>&1 | tail -3
for r in 1 2; do
echo "=== B64 s4 lite (default)"; timeout 400 python bench.py --steps 200 --no-cpu-baseline 2>&1 | python -c "$P"
echo "=== B64 s4 full-SM tc"; TMAE_ATTN_LITE=0 timeout 400 python bench.py --steps 200 --no-cpu-baseline 2>&1 | python -c "$P"
echo "=== B64 s4 mma.sync"; TMAE_NO_TC_ATTN=1 timeout 400 python bench.py --steps 200 --no-cpu-baseline 2>&1 | python -c "$P"
done
echo "=== B64 s1 lite forced"; TMAE_ATTN_LITE=1 timeout 400 python bench.py --steps 200 --streams 1 --no-cpu-baseline 2>&1 | python -c "$P"
echo "=== B64 s1 default (full)"; timeout 400 python bench.py --steps 200 --streams 1 --no-cpu-baseline 2>&1 | python -c "$P"
