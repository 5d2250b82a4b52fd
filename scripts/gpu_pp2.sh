#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
P='import sys,json; d=json.loads([l for l in sys.stdin if l.startswith("{")][-1]); print(d["config"]["streams_per_gpu"], round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"],4), round(d["roofline"]["step_aggregate"]["frac"],4), d["batch_latency"]["ms_median"])'
for S in 4 2; do
echo "=== s$S default"; timeout 400 python bench.py --steps 200 --streams $S --no-cpu-baseline 2>&1 | python -c "$P"
echo "=== s$S persistent pairs in share mode"; TMAE_PP_SHARE=1 timeout 400 python bench.py --steps 200 --streams $S --no-cpu-baseline 2>&1 | python -c "$P"
done
echo "=== s1 persistent pairs"; timeout 400 python bench.py --steps 200 --streams 1 --no-cpu-baseline 2>&1 | python -c "$P"
echo "=== s1 NO_PAIR (one-CTA persistent)"; TMAE_NO_PAIR=1 timeout 400 python bench.py --steps 200 --streams 1 --no-cpu-baseline 2>&1 | python -c "$P"
