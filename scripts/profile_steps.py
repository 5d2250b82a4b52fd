"""Per-launch device times of one forward (CUDA events inside the library) -> gpurun_out/steps_<workload>.json"""
import json, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from bench import WORKLOADS, make_inputs
from textmae_image_compression_b200 import MCM, PathConfig, make_state_dict

wl = sys.argv[1] if len(sys.argv) > 1 else "B64"
batch_override = int(sys.argv[2]) if len(sys.argv) > 2 else None
kwargs, batch, desc, kind = WORKLOADS[wl]
batch = batch_override or batch
cfg = PathConfig(**kwargs)
share = len(sys.argv) > 3 and sys.argv[3] == "share"      # third argument 'share': the half-smem configs of multi-stream mode
m = MCM(**kwargs, share_sm=share); m.load_state_dict(make_state_dict(cfg, 0)); m.cuda().eval()
imgs, scores = make_inputs(kwargs, batch, 0, 2, kind)
imgs = [t.cuda() for t in imgs]; scores = [t.cuda() for t in scores]
for i in range(3): m(imgs[i % 2], scores[i % 2])
torch.cuda.synchronize()
m.profile(True)
acc = None
R = 5
for r in range(R):
    m(imgs[r % 2], scores[r % 2]); torch.cuda.synchronize()
    st = m.profile_read_steps()
    if acc is None: acc = st
    else:
        for a, b in zip(acc, st): a["ms"] += b["ms"]
for a in acc: a["ms"] /= R
m.profile(False)
tot = sum(a["ms"] for a in acc)
out = {"workload": desc, "batch": batch, "total_ms_sum_of_launches": tot, "steps": acc}
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / f"steps_{wl}_{batch}.json").write_text(json.dumps(out, indent=0))
print(f"{wl} batch {batch}: sum of launches {tot:.3f} ms over {len(acc)} launches")
groups = {}
for a in acc:
    key = a["name"].split(".")[0] if not a["name"].startswith("blk") else "blk." + a["name"].split(".")[1]
    g = groups.setdefault(key, [0.0, 0.0, 0]); g[0] += a["ms"]; g[1] += a["flops"]; g[2] += 1
for k, (ms, fl, n) in sorted(groups.items(), key=lambda kv: -kv[1][0]):
    print(f"  {k:24s} {ms:8.3f} ms  {n:4d} launches  {fl / ms / 1e9 if ms > 0 else 0:8.1f} TFLOP/s")
for a in acc[:60]:
    print(f"    {a['name']:28s} {a['ms']*1e3:8.1f} us  ctas {a['ctas']:4d} bn {a['block_n']:3d}  {a['flops']/max(a['ms'],1e-9)/1e9:8.1f} TF/s")
