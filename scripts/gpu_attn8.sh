#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
P='import sys,json; d=json.loads([l for l in sys.stdin if l.startswith("{")][-1]); print(round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"],4), d["roofline"]["families"].get("attention"), d["batch_latency"]["ms_median"])'
timeout 300 python -m pytest tests/test_gpu_engine.py -m gpu -x -q -k attention 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_forward.py -m gpu -x -q -k "vitL or large or L256 or sweep" 2>&1 | tail -3
for r in 1 2; do
echo "=== L256 tail"; timeout 400 python bench.py --workload L256 --steps 60 --no-cpu-baseline 2>&1 | python -c "$P"
echo "=== L256 no tail"; TMAE_NO_ATTN_TAIL=1 timeout 400 python bench.py --workload L256 --steps 60 --no-cpu-baseline 2>&1 | python -c "$P"
done
echo "=== DIV2K_L256 tail"; timeout 400 python bench.py --workload DIV2K_L256 --steps 60 --no-cpu-baseline 2>&1 | python -c "$P"
echo "=== DIV2K_L256 no tail"; TMAE_NO_ATTN_TAIL=1 timeout 400 python bench.py --workload DIV2K_L256 --steps 60 --no-cpu-baseline 2>&1 | python -c "$P"
