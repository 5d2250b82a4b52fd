"""Patch-score generation on the GPU: the host-side mirror of the reference's `generate_scores_file.py`.

`preprocess_image_scores(dataset_path, output_file)` keeps the reference's signature and output file
(generate_scores_file.py:13-36: sorted rglob, one fp32 [L] score vector per image, torch.save of the stack); the pixel work -
`utils/map.py` Division_Merge_Segmented / laplacian, `utils/distribution.py` cal_patch_score, the normalisation - runs in
libtmae_b200.so (`tmae_generate_scores`, csrc/scores.cu) on all images of one size at once.  Decoding the files to
grayscale (`cv2.imread(..., IMREAD_GRAYSCALE)`, :19) stays on the host.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import torch

from . import _native


def generate_scores(gray: torch.Tensor, out_side: int = 224, return_maps: bool = False, stream=None):
    """gray: uint8 CUDA tensor [N, H, W] (or [H, W]) -> fp32 [N, (out_side/16)^2] (`total_score` of generate_scores_file.py:24-31).

    return_maps=True also returns (s_map, t_map, segmented): Division_Merge_Segmented(img, (S, S)), laplacian(img, (S, S)) as the
    reference computes them (the Laplacian of the already segmented image) and the segmented image itself.
    """
    if gray.dtype != torch.uint8:
        raise TypeError(f"grayscale images must be uint8 (cv2.IMREAD_GRAYSCALE), got {gray.dtype}")
    if not gray.is_cuda:
        raise RuntimeError("generate_scores runs on the GPU: pass a CUDA tensor (there is no CPU fallback)")
    squeeze = gray.dim() == 2
    if squeeze:
        gray = gray[None]
    if gray.dim() != 3:
        raise ValueError("expected [N, H, W] or [H, W]")
    gray = gray.contiguous()
    n, h, w = gray.shape
    lib = _native.load()
    need = lib.tmae_scores_workspace_bytes(n, h, w, out_side)
    if need == 0:
        raise ValueError(f"unsupported geometry: image {h} x {w} (need >= 8 x 8), out_side {out_side} (a positive multiple of 16)")
    dev = gray.device
    with torch.cuda.device(dev):
        ws = torch.empty(need, dtype=torch.uint8, device=dev)
        L = (out_side // 16) ** 2
        scores = torch.empty(n, L, dtype=torch.float32, device=dev)
        out = _native.TmaeScoreOutputs()
        out.scores = scores.data_ptr()
        maps = None
        if return_maps:
            maps = (torch.empty(n, out_side, out_side, dtype=torch.uint8, device=dev),
                    torch.empty(n, out_side, out_side, dtype=torch.uint8, device=dev),
                    torch.empty(n, h, w, dtype=torch.uint8, device=dev))
            out.s_map, out.t_map, out.segmented = (m.data_ptr() for m in maps)
        st = stream if stream is not None else torch.cuda.current_stream(dev)
        rc = lib.tmae_generate_scores(C.c_void_p(gray.data_ptr()), n, h, w, out_side, C.byref(out), C.c_void_p(ws.data_ptr()),
                                      need, C.c_void_p(st.cuda_stream))
        _native.check(rc, None, ValueError)
        ws.record_stream(st)
    if squeeze:
        scores = scores[0]
        if maps is not None:
            maps = tuple(m[0] for m in maps)
    return (scores, *maps) if return_maps else scores


def preprocess_image_scores(dataset_path, output_file, device="cuda", out_side: int = 224, batch: int = 64):
    """generate_scores_file.py:13-36 with the per-image pixel loops on the GPU.  Returns the saved [n_images, L] tensor."""
    import cv2  # host-side decode only
    import numpy as np

    img_paths = sorted(Path(dataset_path).rglob("*.*"))
    grays = [cv2.imread(str(p), cv2.IMREAD_GRAYSCALE) for p in img_paths]
    L = (out_side // 16) ** 2
    scores = torch.empty(len(grays), L, dtype=torch.float32)
    by_size: dict = {}
    for i, g in enumerate(grays):
        if g is None:
            raise ValueError(f"cannot read {img_paths[i]} as an image")
        by_size.setdefault(g.shape, []).append(i)
    for ids in by_size.values():
        for b0 in range(0, len(ids), batch):
            sel = ids[b0:b0 + batch]
            stack = torch.from_numpy(np.stack([grays[i] for i in sel])).to(device)
            scores[sel] = generate_scores(stack, out_side).cpu()
    print("Shape of list_total_score: ", scores.shape)
    torch.save(scores, output_file)
    return scores
