"""Target for ncu: builds the B64 model, runs warm-up forwards, then one profiled forward (batch 64)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from bench import WORKLOADS, make_inputs
from textmae_image_compression_b200 import MCM, PathConfig, make_state_dict
wl = sys.argv[1] if len(sys.argv) > 1 else "B64"
nfwd = int(sys.argv[2]) if len(sys.argv) > 2 else 2
kwargs, batch, desc, kind = WORKLOADS[wl]
import os
m = MCM(**kwargs, precise=os.environ.get("TMAE_BENCH_PRECISE") or None); m.load_state_dict(make_state_dict(PathConfig(**kwargs), 0)); m.cuda().eval()
imgs, scores = make_inputs(kwargs, batch, 0, 1, kind)
imgs = imgs[0].cuda(); scores = scores[0].cuda()
for _ in range(nfwd):
    out = m(imgs, scores)
torch.cuda.synchronize()
print("bpp", out["bpp"][:4].tolist())
n = m.launch_count(batch)
print("launches_per_forward", n - 1)          # every plan step is one kernel except the rate-accumulator memset
m.profile(True); m(imgs, scores); torch.cuda.synchronize()
print("gemm_per_forward", sum(1 for s_ in m.profile_read_steps() if s_["ctas"] > 0))
