#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
P='import sys,json; d=json.loads([l for l in sys.stdin if l.startswith("{")][-1]); print(round(d["value"]), round(d["e2e"]["value"]), d["ms_per_step"], round(d["roofline"]["frac"],4), round(d["roofline"]["step_aggregate"]["frac"],4), d["gpu_launches"]//d["steps"], d["roofline"]["families"].get("attention"))'
echo "=== forward tests"; timeout 900 python -m pytest -q -m gpu tests/test_gpu_forward.py -x -k "not precise" 2>&1 | tail -4
for S in 4 1; do
  echo "=== B64 streams $S tc attention"; timeout 400 python bench.py --steps 40 --streams $S --no-cpu-baseline 2>&1 | python -c "$P"
  echo "=== B64 streams $S mma.sync attention"; TMAE_NO_TC_ATTN=1 timeout 400 python bench.py --steps 40 --streams $S --no-cpu-baseline 2>&1 | python -c "$P"
done
echo "=== B144 tc"; timeout 400 python bench.py --workload B144 --steps 30 --no-cpu-baseline 2>&1 | python -c "$P"
echo "=== B144 mma.sync"; TMAE_NO_TC_ATTN=1 timeout 400 python bench.py --workload B144 --steps 30 --no-cpu-baseline 2>&1 | python -c "$P"
echo "=== L256 tc"; timeout 400 python bench.py --workload L256 --steps 20 --no-cpu-baseline 2>&1 | python -c "$P"
echo "=== L256 mma.sync"; TMAE_NO_TC_ATTN=1 timeout 400 python bench.py --workload L256 --steps 20 --no-cpu-baseline 2>&1 | python -c "$P"
