#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_engine.py -m gpu -x -q -k attention 2>&1 | tail -15 | tee gpurun_out/attn_test.log
timeout 120 python scripts/bench_attention.py 2>&1 | tail -8 | tee gpurun_out/attn_bench.log
