"""Freezes outputs of the REFERENCE's own `MCM.forward` (models/Compression/MCM.py executed verbatim from /root/reference,
third-party leaf classes from oracle/ref_stubs.py - see oracle/ref_exec.py) into tests/golden/refexec_small.pt, so that the
GPU box (where /root/reference does not exist) can check the CUDA path and the reconstruction half against reference
outputs.  Run in the build container:

    python tests/golden/make_refexec_golden.py
"""
from __future__ import annotations

import sys
from pathlib import Path

import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))

from oracle import ref_exec, ref_pins  # noqa: E402
from textmae_image_compression_b200 import PathConfig, make_state_dict  # noqa: E402

SMALL = dict(img_size=64, encoder_embed_dim=128, encoder_depth=2, encoder_num_heads=2, num_keep_patches=16,
             decoder_embed_dim=64, decoder_depth=2, decoder_num_heads=2)
MID = dict(img_size=128, encoder_embed_dim=128, encoder_depth=2, encoder_num_heads=2, num_keep_patches=64,
           decoder_embed_dim=128, decoder_depth=2, decoder_num_heads=2)


def one(kw, seed, n):
    cfg = PathConfig(**kw)
    sd = make_state_dict(cfg, seed=seed, include_decoder=True)
    model, missing, unexpected = ref_exec.build_reference_model(cfg, sd)
    assert not missing and not unexpected, (missing, unexpected)
    g = torch.Generator().manual_seed(100 + seed)
    imgs = torch.rand(n, 3, cfg.img_size, cfg.img_size, generator=g)
    scores = torch.rand(n, cfg.num_patches, generator=g)
    out = ref_exec.reference_forward(model, imgs, scores)
    x_remain, ids_restore = model.forward_encoder(imgs, scores)
    crit = ref_pins.load_rate_distortion_loss()(lmbda=1e-2)
    rd = crit(out, imgs)
    return {"kwargs": kw, "seed": seed, "imgs": imgs, "scores": scores,
            "y_lik": out["likelihoods"]["y"], "z_lik": out["likelihoods"]["z"], "x_hat": out["x_hat"],
            "loss": torch.stack([l.float() for l in out["loss"]]), "x_remain": x_remain, "ids_restore": ids_restore,
            "rd_loss": {k: float(v) for k, v in rd.items()}, "aux_loss": float(model.aux_loss())}


def main():
    assert ref_exec.reference_available(), "needs /root/reference"
    blob = {"small": one(SMALL, 3, 3), "mid": one(MID, 5, 2)}
    torch.save(blob, HERE / "refexec_small.pt")
    print({k: {kk: (tuple(vv.shape) if torch.is_tensor(vv) else vv) for kk, vv in v.items() if kk not in ("kwargs",)} for k, v in blob.items()})


if __name__ == "__main__":
    main()
