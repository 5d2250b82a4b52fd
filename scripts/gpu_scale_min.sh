#!/bin/bash
# usage: scripts/gpu_scale_min.sh N   (inside a gpurun --gpus N call): the headline workload only, weak scaling
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
N=$1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --warmup 12 --steps 60 --no-cpu-baseline > gpurun_out/scale_${N}gpu_B64_weak.log 2>&1
echo "rc=$?"; grep '^{' gpurun_out/scale_${N}gpu_B64_weak.log | tail -1 | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["n_gpus"], d["scaling"], round(d["value"]), round(d["e2e"]["value"]), d["ms_per_step"])' || tail -5 gpurun_out/scale_${N}gpu_B64_weak.log
