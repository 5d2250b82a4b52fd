#!/bin/bash
# Round-2 final GPU check.  usage: gpu_round3.sh tests,smoke,bench | workloads,ncu
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > $O/$name.log 2>&1; echo "rc=$? $name"; tail -n "${TAILN:-6}" $O/$name.log | cut -c1-600; }
STAGES=${1:-tests,smoke,bench}
if [[ $STAGES == *tests* ]]; then
  TMO=2000 TAILN=30 run pytest_gpu python -m pytest -p no:cacheprovider -q -m gpu tests
  cp $O/parity_report.json $O/r02_parity_report.json 2>/dev/null
fi
if [[ $STAGES == *smoke* ]]; then TMO=300 TAILN=2 run smoke python __graft_entry__.py smoke; fi
if [[ $STAGES == *bench* ]]; then
  TMO=600 TAILN=1 run bench_B64 python bench.py
  TMO=600 TAILN=1 run bench_B64_precise python bench.py --precise all --steps 30 --no-cpu-baseline
  TMO=600 TAILN=1 run bench_ref python bench.py --impl reference --steps 5 --warmup 1
fi
if [[ $STAGES == *workloads* ]]; then
  for W in KODAK24 KODAK24_K64 TILES288 B144 L256 DIV2K_L400 DIV2K_L256 DIV2K_L144; do
    TMO=600 TAILN=1 run bench_$W python bench.py --workload $W --steps 40 --no-cpu-baseline
  done
  TMO=600 TAILN=1 run bench_KODAK24_precise python bench.py --workload KODAK24 --precise all --steps 20 --no-cpu-baseline
  TMO=600 TAILN=1 run bench_B64_batch512 python bench.py --batch 512 --streams 1 --steps 10 --no-cpu-baseline
  TMO=300 TAILN=60 run steps_B64 python scripts/profile_steps.py B64
fi
if [[ $STAGES == *ncu* ]]; then bash scripts/gpu_profile.sh B64; fi
if [[ $STAGES == *scores* ]]; then
  TMO=600 TAILN=4 run scores_test python -m pytest -p no:cacheprovider -q -m gpu tests/test_gpu_scores.py
  TMO=300 TAILN=2 run scores_bench python scripts/bench_scores.py
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none -k regex:score_ -c 40 --csv --log-file $O/scores_launches.csv python scripts/bench_scores.py > $O/scores_ncu.log 2>&1
  echo "rc=$? scores ncu"
fi
