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
TMO=600 TAILN=3 run bench python bench.py --steps 20 --warmup 3
if [ "$1" == "ncu" ] || [ "$2" == "ncu" ]; then
  echo "=== ncu"
  TMO=300 run ncu_plain python scripts/ncu_target.py B64 2
  # forward kernels only (skip weight prepack): per-launch durations of the 2nd forward
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'gemm_tc|attention|layernorm|mask_select|gather_patches|bottleneck|gaussian_slice|rate_finalize' \
      --launch-skip 236 --launch-count 236 --csv --log-file $O/launches_B64.csv python scripts/ncu_target.py B64 2 > $O/ncu_launches.log 2>&1
  echo "rc=$? ncu launches"
  # full sections for representative GEMM launches of the 2nd forward: patch-embed, blk0 qkv/proj/fc1/fc2
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc --launch-skip 183 --launch-count 5 \
      -o $O/prof_encoder -f python scripts/ncu_target.py B64 2 > $O/ncu_full1.log 2>&1
  echo "rc=$? ncu full encoder"
  # h_s.8, cc.0.0 .. cc.0.8, lrp.0.0 .. lrp.0.8
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc --launch-skip 245 --launch-count 11 \
      -o $O/prof_slices -f python scripts/ncu_target.py B64 2 > $O/ncu_full2.log 2>&1
  echo "rc=$? ncu full slices"
fi
