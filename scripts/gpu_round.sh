#!/bin/bash
# Full GPU check: parity tests, smoke, per-launch profile, bench, optional ncu captures. Each stage bounded.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > $O/$name.log 2>&1; echo "rc=$? $name"; tail -n "${TAILN:-6}" $O/$name.log; }
if [ "$1" != "quick" ]; then
TMO=900 TAILN=12 run pytest_gpu python -m pytest -p no:cacheprovider -q -m gpu tests
TMO=300 TAILN=2 run smoke python __graft_entry__.py smoke
fi
TMO=300 TAILN=40 run steps_B64 python scripts/profile_steps.py B64
TMO=600 TAILN=3 run bench python bench.py
if [ "$1" == "ncu" ] || [ "$2" == "ncu" ]; then
  bash scripts/gpu_profile.sh B64          # launch list + --set full captures (launch counts are read from the run itself)
fi
