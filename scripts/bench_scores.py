"""Throughput of the GPU score generator (tmae_generate_scores) on Kodak-sized images; prints one JSON line.
The CPU figure beside it is the oracle (vectorised numpy restatement; the reference's own Python loops take 0.3-1.0 s/image)."""
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import ref_scores  # noqa: E402  (cpu baseline leg only)
from textmae_image_compression_b200.scores import generate_scores  # noqa: E402

n, h, w = 24, 512, 768
imgs = np.stack([ref_scores.synthetic_gray(k % 5, h, w, k) for k in range(n)])
g = torch.from_numpy(imgs).cuda()
for _ in range(5):
    generate_scores(g)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 50
e0.record()
for _ in range(reps):
    generate_scores(g)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
t0 = time.perf_counter()
for k in range(4):
    ref_scores.generate_scores(imgs[k])
cpu_s = (time.perf_counter() - t0) / 4
peaks = json.load(open(ROOT / "MEASURED_PEAKS.json")) if (ROOT / "MEASURED_PEAKS.json").exists() else {"hbm_gbs": 6446.9}
alg_bytes = n * (3 * h * w + 224 * 224 * 0 + 196 * 4)      # grey read by the judge and the segment pass, segmented image written
print(json.dumps({"metric": "images/s patch-score generation", "value": n / (ms * 1e-3), "ms_per_batch": ms, "batch": n,
                  "image": [h, w], "launches_per_batch": 5, "algorithmic_MB_per_batch": alg_bytes / 1e6,
                  "GB/s": alg_bytes / (ms * 1e-3) / 1e9, "frac_of_hbm_peak": alg_bytes / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                  "cpu_oracle_images_per_s": 1.0 / cpu_s}))
