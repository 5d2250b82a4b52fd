#!/bin/bash
# Full GPU check: parity tests, smoke, per-launch profile, bench, ncu launch list. Each stage bounded.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > $O/$name.log 2>&1; echo "rc=$? $name"; tail -n "${TAILN:-6}" $O/$name.log; }
TMO=900 TAILN=12 run pytest_gpu python -m pytest -p no:cacheprovider -q -m gpu tests
TMO=300 TAILN=2 run smoke python __graft_entry__.py smoke
TMO=300 TAILN=80 run steps_B64 python scripts/profile_steps.py B64
TMO=600 TAILN=3 run bench python bench.py --steps 20 --warmup 3
if [ "$1" == "ncu" ]; then
  echo "=== ncu launch list"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/launches_B64.csv \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_launches.log 2>&1
  echo "rc=$? ncu"; tail -3 $O/ncu_launches.log
fi
