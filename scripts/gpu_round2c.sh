#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
P='import sys,json; d=json.loads([l for l in sys.stdin if l.startswith("{")][-1]); print(round(d["value"]), round(d["e2e"]["value"]), d["ms_per_step"], round(d["roofline"]["frac"],4), round(d["roofline"]["step_aggregate"]["frac"],4), d["gpu_launches"]//d["steps"])'
echo "=== engine tests"; timeout 600 python -m pytest -q -m gpu tests/test_gpu_engine.py -x 2>&1 | tail -6
echo "=== forward tests"; timeout 900 python -m pytest -q -m gpu tests/test_gpu_forward.py tests/test_gpu_recon.py tests/test_gpu_entropy.py -x 2>&1 | tail -8
for S in 4 1; do
  echo "=== streams $S (pair + pair conv + LN fold)"; timeout 400 python bench.py --steps 60 --streams $S --no-cpu-baseline 2>&1 | python -c "$P"
  echo "=== streams $S no pair conv"; TMAE_NO_PAIR_CONV=1 timeout 400 python bench.py --steps 60 --streams $S --no-cpu-baseline 2>&1 | python -c "$P"
  echo "=== streams $S no LN fold"; TMAE_NO_LN_FOLD=1 timeout 400 python bench.py --steps 60 --streams $S --no-cpu-baseline 2>&1 | python -c "$P"
done
echo "=== steps share"; timeout 300 python scripts/profile_steps.py B64 64 share 2>&1 | grep -E "sum of launches|^  " | head -14
