"""testing.py's `--entropy-estimation` flow (testing.py:99-121, 214-233) on the committed Kodak fixtures, every piece from this
package: grey image -> GPU patch scores -> MCM.forward (rate half in libtmae_b200, reconstruction half in stock PyTorch) ->
bpp / PSNR, plus the ids_restore side information the reference's default path adds (testing.py:73-76, 89).

    python examples/kodak_entropy_estimation.py [--keep 144] [--precise all]

Weights are the seeded synthetic checkpoint (no trained checkpoint ships with the reference), so the PSNR is meaningless; the
point is that the call sequence of the reference's CLI runs unchanged."""
import argparse
import math
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from textmae_image_compression_b200 import MCM, make_state_dict, vit_base  # noqa: E402
from textmae_image_compression_b200.huffman import HuffmanCoding  # noqa: E402
from textmae_image_compression_b200.scores import generate_scores  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--keep", type=int, default=144)
    ap.add_argument("--precise", default=None)
    args = ap.parse_args()
    golden = ROOT / "tests" / "golden"
    gray = np.load(golden / "kodak_gray6.npz")
    names = sorted(gray.files)
    rgb = torch.from_numpy(np.load(golden / "kodak_224.npz")["imgs"][:len(names)]).permute(0, 3, 1, 2).float() / 255.0

    cfg = vit_base(args.keep)
    model = MCM(num_keep_patches=args.keep, precise=args.precise)          # testing.py:125 MCM().from_state_dict(...)
    model.load_state_dict(make_state_dict(cfg, seed=0, include_decoder=True))
    model = model.cuda().eval()
    model.update(force=True)                                               # testing.py:223

    for i, name in enumerate(names):
        x = rgb[i:i + 1].cuda()
        total_score = generate_scores(torch.from_numpy(gray[name]).cuda())[None]   # generate_scores_file.py:19-31
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = model(x, total_score)                                        # testing.py:101
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        num_pixels = x.size(0) * x.size(2) * x.size(3)
        bpp = sum(torch.log(lk).sum() / (-math.log(2) * num_pixels) for lk in out["likelihoods"].values()).item()   # testing.py:110-112
        side, _, _ = HuffmanCoding().compress(out["ids_restore"])          # testing.py:73-76
        line = f"{name}: bpp {bpp:.4f} (+ ids_restore {len(side) / num_pixels:.4f}), forward {dt * 1e3:.2f} ms"
        if "x_hat" in out:
            mse = torch.mean((x - out["x_hat"].clamp(0, 1)) ** 2).item()
            line += f", psnr {10 * math.log10(1.0 / max(mse, 1e-12)):.2f} dB (synthetic weights)"
        print(line)


if __name__ == "__main__":
    main()
