#!/bin/bash
# usage: scripts/gpu_scale.sh N   (inside a gpurun --gpus N call)
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --warmup 12 --no-cpu-baseline"
run() { name=$1; shift; timeout 600 $TR "$@" > gpurun_out/$name.log 2>&1; echo "rc=$? $name"; grep '^{' gpurun_out/$name.log | tail -1 | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["n_gpus"], d["scaling"], d["config"]["name"], d["config"]["per_gpu_batch"], d["config"]["global_batch"], round(d["value"]), round(d["e2e"]["value"]), d["ms_per_step"])' 2>/dev/null || tail -5 gpurun_out/$name.log; }
run scale_${N}gpu_B64_weak --steps 60
if [ "$N" == "8" ]; then
  run scale_${N}gpu_L256_strong --workload L256 --batch 256 --scaling strong --steps 30
  run scale_${N}gpu_TILES288_strong --workload TILES288 --scaling strong --steps 40
  run scale_${N}gpu_KODAK24_strong --workload KODAK24 --scaling strong --steps 60
  run scale_${N}gpu_DIV2K_L256_strong --workload DIV2K_L256 --scaling strong --steps 60
fi
