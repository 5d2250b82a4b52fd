#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
for mode in pair nopair; do
  for share in "" share; do
    if [ $mode == nopair ]; then export TMAE_NO_PAIR=1; else unset TMAE_NO_PAIR; fi
    echo "=== $mode $share"
    timeout 300 python scripts/profile_steps.py B64 64 $share 2>&1 | grep -E "blk\.0|blk0|sum of launches|blk\.5 " | head -8
  done
done
unset TMAE_NO_PAIR
echo "=== bench pair (streams 4)"; timeout 400 python bench.py --steps 60 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['step_aggregate']['frac'])"
echo "=== bench nopair (streams 4)"; TMAE_NO_PAIR=1 timeout 400 python bench.py --steps 60 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['step_aggregate']['frac'])"
echo "=== bench pair (streams 1)"; timeout 400 python bench.py --steps 60 --streams 1 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['step_aggregate']['frac'])"
echo "=== bench nopair (streams 1)"; TMAE_NO_PAIR=1 timeout 400 python bench.py --steps 60 --streams 1 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['step_aggregate']['frac'])"
