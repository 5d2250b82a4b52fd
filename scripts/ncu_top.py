"""Summarise an .ncu-rep: key raw metrics + the most-sampled SASS instructions with stall reasons."""
import csv, subprocess, sys, io
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(raw)))
hdr, units, rows = r[0], r[1], r[2:]
want = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_bytes.sum", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "launch__grid_size", "launch__block_size", "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct"]
for i, h in enumerate(hdr):
    if h in want: print(f"{h:70s} {units[i]:14s} {[row[i] for row in rows]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = None; body = []
for x in rows:
    if x and x[0] == "Address": h = x; continue
    if h is not None and x and x[0].startswith("0x"): body.append(x)
iS = h.index("# Samples"); iSrc = h.index("Source"); iEx = h.index("Instructions Executed")
stall = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
tot = sum(int(x[iS]) for x in body); print("total samples", tot, "instructions", len(body))
agg = {}
for x in body:
    for i in stall:
        if x[i] not in ("", "0"): agg[h[i]] = agg.get(h[i], 0) + int(x[i])
print(sorted(agg.items(), key=lambda kv: -kv[1]))
order = sorted(range(len(body)), key=lambda j: -int(body[j][iS]))[:topn]
for j in sorted(order):
    x = body[j]
    print(f"{j:5d} smp {x[iS]:>5s} exe {x[iEx]:>7s}  {x[iSrc][:70]:70s} " + " ".join(f"{h[i][6:]}={x[i]}" for i in stall if x[i] not in ("", "0")))
