#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
P='import sys,json; d=json.loads([l for l in sys.stdin if l.startswith("{")][-1]); print(round(d["value"]), round(d["e2e"]["value"]), d["ms_per_step"])'
for S in 1 4; do
  echo "=== r1 streams $S"; (cd _r1 && timeout 400 python bench.py --steps 60 --streams $S --no-cpu-baseline 2>&1 | python -c "$P")
  echo "=== now nopair streams $S"; TMAE_NO_PAIR=1 timeout 400 python bench.py --steps 60 --streams $S --no-cpu-baseline 2>&1 | python -c "$P"
done
echo "=== r1 steps"; (cd _r1 && timeout 300 python scripts/profile_steps.py B64 64 2>&1 | grep -E "sum of launches|^  " | head -14)
echo "=== now steps (nopair)"; TMAE_NO_PAIR=1 timeout 300 python scripts/profile_steps.py B64 64 2>&1 | grep -E "sum of launches|^  " | head -14
