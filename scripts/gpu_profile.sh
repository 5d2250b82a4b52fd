#!/bin/bash
# ncu evidence for profiles/: (1) per-launch list of one forward (durations, DRAM bytes, tensor-pipe activity),
# (2) --set full capture of representative GEMM-engine launches.  Plain launches (no CUDA graph) so every kernel is a
# separate ncu launch.  One GPU, single process.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
O=gpurun_out
export TMAE_NO_GRAPH=1
WL=${1:-B64}
python scripts/ncu_target.py $WL 2 > $O/ncu_plain.log 2>&1 || { echo "plain run failed"; tail -5 $O/ncu_plain.log; exit 1; }
NL=$(grep -o "launches_per_forward [0-9]*" $O/ncu_plain.log | awk '{print $2}')
echo "kernel launches per forward: $NL"
KRE='regex:gemm_tc|attention|layernorm_kernel|mask_select|gather_patches|bottleneck_kernel|gaussian_slice|rate_finalize|copy_outputs'
# --cache-control none: the caches are NOT flushed between kernels, so L2-resident activations and the just-prefetched
# next-layer weights count as hits like in a real forward; few metrics -> one or two replay passes
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed \
    --clock-control none --cache-control none -k "$KRE" --launch-skip $NL --launch-count $NL --csv --log-file $O/launches_$WL.csv \
    python scripts/ncu_target.py $WL 2 > $O/ncu_launches.log 2>&1
echo "rc=$? ncu launch list"
# representative launches of the 2nd forward (gemm-only indices): 1=blk0.qkv 2=proj 3=fc1 4=fc2
NG=$(grep -o "gemm_per_forward [0-9]*" $O/ncu_plain.log | awk '{print $2}')
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc --launch-skip $((NG + 1)) --launch-count 4 \
    -o $O/prof_encoder_$WL -f python scripts/ncu_target.py $WL 2 > $O/ncu_full1.log 2>&1
echo "rc=$? ncu full encoder"
# cc.0.0 .. cc.0.8 (first slice): gemm index 49 + 4 g_a + 5 h_a + 5 h_s = 63
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc --launch-skip $((NG + 63)) --launch-count 5 \
    -o $O/prof_slice0_$WL -f python scripts/ncu_target.py $WL 2 > $O/ncu_full2.log 2>&1
echo "rc=$? ncu full slice0"
# the tcgen05 attention kernel of block 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_tc --launch-skip 13 --launch-count 1 \
    -o $O/prof_attn_tc_$WL -f python scripts/ncu_target.py $WL 2 > $O/ncu_attn.log 2>&1
echo "rc=$? ncu full attention"
ls -la $O/*.ncu-rep $O/launches_$WL.csv
