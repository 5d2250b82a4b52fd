#!/bin/bash
# Staged GPU bring-up: each stage in its own process (a trapped kernel must not poison later stages), each bounded.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,memory.total --format=csv > $O/gpu.txt 2>&1
python -c "import os,torch;print('cpus',os.cpu_count(),'torch',torch.__version__,torch.backends.cpu.get_cpu_capability())" >> $O/gpu.txt 2>&1
grep -m1 "model name" /proc/cpuinfo >> $O/gpu.txt
PT="python -m pytest -p no:cacheprovider -q -m gpu"
run() { name=$1; shift; echo "=== $name"; timeout "$TMO" "$@" > $O/$name.log 2>&1; echo "rc=$? $name"; tail -n "${TAILN:-6}" $O/$name.log; }
TMO=600 run engine_checker $PT tests/test_gpu_engine.py -k checker
TMO=300 run engine_tcgen05 $PT tests/test_gpu_engine.py -k tcgen05
TMO=300 run mask $PT tests/test_gpu_mask.py
TMO=300 run entropy $PT tests/test_gpu_entropy.py
TMO=900 TAILN=25 run forward $PT tests/test_gpu_forward.py -s
TMO=300 run smoke python __graft_entry__.py smoke
TMO=600 TAILN=3 run bench python bench.py --steps 10 --warmup 3
