#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
P='import sys,json; d=json.loads([l for l in sys.stdin if l.startswith("{")][-1]); print(d["config"]["streams_per_gpu"], round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"],4), round(d["roofline"]["step_aggregate"]["frac"],4), d["clocks"]["sm_mhz"], d["clocks"]["power_w_max"])'
for S in 2 3 4 5 6 8; do
timeout 400 python bench.py --steps 240 --streams $S --no-cpu-baseline 2>&1 | python -c "$P"
done
