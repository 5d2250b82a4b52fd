#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
python -m pytest tests/test_gpu_scores.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/scores_test.log
python scripts/bench_scores.py 2>&1 | tail -3 | tee gpurun_out/scores_bench.log
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:score_ -c 40 --csv --log-file gpurun_out/scores_launches.csv python scripts/bench_scores.py > gpurun_out/scores_ncu.log 2>&1
python - <<'P'
import csv
rows=[r for r in csv.reader(open('gpurun_out/scores_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value'); ii=hdr.index('ID')
seen={}
for r in rows[1:]:
    if r[mi]=='gpu__time_duration.sum': seen.setdefault(r[ki].split('(')[0],[]).append(float(r[vi].replace(',','')))
for k,v in seen.items(): print(k, len(v), sum(v)/len(v))
P
