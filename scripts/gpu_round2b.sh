#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
P='import sys,json; d=json.loads([l for l in sys.stdin if l.startswith("{")][-1]); print(round(d["value"]), round(d["e2e"]["value"]), d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["step_aggregate"]["frac"], d["gpu_launches"]//d["steps"])'
echo "=== tests (engine + forward quick)"; timeout 900 python -m pytest -q -m gpu tests/test_gpu_engine.py tests/test_gpu_forward.py -x 2>&1 | tail -8
echo "=== host overhead S=1"; timeout 300 python scripts/host_overhead.py 1 2>&1 | head -30
for S in 1 4; do
  echo "=== now streams $S"; timeout 400 python bench.py --steps 60 --streams $S --no-cpu-baseline 2>&1 | python -c "$P"
  echo "=== now streams $S no LN fold"; TMAE_NO_LN_FOLD=1 timeout 400 python bench.py --steps 60 --streams $S --no-cpu-baseline 2>&1 | python -c "$P"
done
echo "=== r1 streams 4"; (cd _r1 && timeout 400 python bench.py --steps 60 --streams 4 --no-cpu-baseline 2>&1 | python -c 'import sys,json; d=json.loads([l for l in sys.stdin if l.startswith("{")][-1]); print(round(d["value"]), round(d["e2e"]["value"]), d["ms_per_step"])')
