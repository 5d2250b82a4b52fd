"""Aggregate an ncu per-launch CSV (gpu__time_duration, dram bytes, tensor-pipe activity) into per-kernel shares.
usage: python scripts/summarize_launches.py gpurun_out/launches_B64.csv profiles/r01_launches_B64"""
import csv, json, sys
from collections import defaultdict
src, dst = sys.argv[1], sys.argv[2]
rows = list(csv.reader(l for l in open(src) if l.startswith('"')))
hdr, rows = rows[0], rows[1:]
iK, iM, iV, iID, iG = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID"), hdr.index("Grid Size")
launch = defaultdict(dict)
for r in rows:
    d = launch[int(r[iID])]
    d["kernel"] = r[iK].split("(")[0].replace("tmae::", "").replace("<unnamed>::", "").replace("unnamed>::", "").replace("void ", "")
    d["grid"] = r[iG]
    d[r[iM]] = float(r[iV].replace(",", ""))
fam = defaultdict(lambda: {"launches": 0, "us": 0.0, "dram_read_MB": 0.0, "dram_write_MB": 0.0, "tensor_pct_x_us": 0.0})
for i in sorted(launch):
    d = launch[i]
    name = d["kernel"].split("<")[0] if "gemm_tc" not in d["kernel"] else "gemm_tc_kernel"
    f = fam[name]
    us = d.get("gpu__time_duration.sum", 0.0) / 1e3
    f["launches"] += 1; f["us"] += us
    f["dram_read_MB"] += d.get("dram__bytes_read.sum", 0.0) / 1e6
    f["dram_write_MB"] += d.get("dram__bytes_write.sum", 0.0) / 1e6
    f["tensor_pct_x_us"] += d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 0.0) * us
tot = sum(f["us"] for f in fam.values())
out = {"source": src, "total_us_sum_of_kernels": tot, "families": {}}
for k, f in sorted(fam.items(), key=lambda kv: -kv[1]["us"]):
    out["families"][k] = {"launches": f["launches"], "us": round(f["us"], 1), "share": round(f["us"] / tot, 4),
                          "dram_read_MB": round(f["dram_read_MB"], 2), "dram_write_MB": round(f["dram_write_MB"], 2),
                          "dram_MB_per_launch": round((f["dram_read_MB"] + f["dram_write_MB"]) / f["launches"], 3),
                          "tensor_pipe_active_pct_time_weighted": round(f["tensor_pct_x_us"] / max(f["us"], 1e-9), 2)}
json.dump(out, open(dst + ".json", "w"), indent=1)
with open(dst + ".csv", "w") as fo:
    fo.write("launch,kernel,grid,duration_us,dram_read_bytes,dram_write_bytes,tensor_pipe_active_pct\n")
    for i in sorted(launch):
        d = launch[i]
        fo.write(f'{i},"{d["kernel"]}","{d["grid"]}",{d.get("gpu__time_duration.sum", 0) / 1e3:.2f},{d.get("dram__bytes_read.sum", 0):.0f},'
                 f'{d.get("dram__bytes_write.sum", 0):.0f},{d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 0):.2f}\n')
print(json.dumps(out, indent=1))
