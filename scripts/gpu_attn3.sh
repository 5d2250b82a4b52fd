#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
P='import sys,json; d=json.loads([l for l in sys.stdin if l.startswith("{")][-1]); print(round(d["value"]), round(d["e2e"]["value"]), d["ms_per_step"], round(d["roofline"]["frac"],4), round(d["roofline"]["step_aggregate"]["frac"],4), d["gpu_launches"]//d["steps"], d["roofline"]["families"].get("attention"))'
timeout 300 python -m pytest tests/test_gpu_engine.py -m gpu -x -q -k attention 2>&1 | tail -5
for WL in B64 B144 L256; do
  echo "=== $WL streams 1 tc"; timeout 400 python bench.py --workload $WL --steps 20 --streams 1 --no-cpu-baseline 2>&1 | python -c "$P"
  echo "=== $WL streams 1 mma.sync"; TMAE_NO_TC_ATTN=1 timeout 400 python bench.py --workload $WL --steps 20 --streams 1 --no-cpu-baseline 2>&1 | python -c "$P"
done
export TMAE_NO_GRAPH=1
for WL in B64 L256; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_tc --launch-skip 13 --launch-count 1 \
    -o gpurun_out/prof_attn_tc_$WL -f python scripts/ncu_target.py $WL 2 > gpurun_out/ncu_attn_$WL.log 2>&1
echo "rc=$? ncu $WL"
done
