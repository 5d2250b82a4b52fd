#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
P='import sys,json; d=json.loads([l for l in sys.stdin if l.startswith("{")][-1]); print(round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"],4), d["roofline"]["families"].get("attention"), d["batch_latency"]["ms_median"])'
timeout 300 python -m pytest tests/test_gpu_engine.py -m gpu -x -q -k attention 2>&1 | tail -3
for WL in B64 B144 L256 KODAK24; do
echo "=== $WL"; timeout 400 python bench.py --workload $WL --steps 100 --no-cpu-baseline 2>&1 | tee gpurun_out/bench_$WL.log | python -c "$P"
done
