#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
P='import sys,json; d=json.loads([l for l in sys.stdin if l.startswith("{")][-1]); print(round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"],4), d["gpu_launches"]//d["steps"], d["roofline"]["families"].get("entropy_elementwise"), round(d["roofline"]["frac"],4), round(d["roofline"]["step_aggregate"]["frac"],4))'
python __graft_entry__.py smoke 2>&1 | tail -3
timeout 1200 python -m pytest tests/test_gpu_forward.py tests/test_gpu_entropy.py tests/test_gpu_recon.py -m gpu -x -q 2>&1 | tail -8
echo "=== B64 fused"; timeout 400 python bench.py --steps 150 --no-cpu-baseline 2>&1 | python -c "$P"
echo "=== B64 unfused"; TMAE_NO_GC_FUSE=1 timeout 400 python bench.py --steps 150 --no-cpu-baseline 2>&1 | python -c "$P"
