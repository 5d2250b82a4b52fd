"""TEST INFRASTRUCTURE - CPU restatement of the reference's patch-score generator.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; the product path
(textmae_image_compression_b200/scores.py -> libtmae_b200.so) never does.

What it restates, in numpy, integer / float64 arithmetic exactly as the reference's numpy + OpenCV calls perform it:

  generate_scores_file.py:19-31   per image: s_map, t_map, patch scores, product, min-max normalisation, fp32 cast.
                                  NB the reference passes the SAME array to both map builders and the first one works in
                                  place, so the Laplacian (`:22`) is taken of the already segmented image - restated as such.
  utils/map.py:6-23   Division_Judge  mean, std(ddof=1) (numpy float64), count of (v - mean) < 2 std, ratio >= 0.95
  utils/map.py:27-31  Merge           60 < v < 150 -> 0, else 255, in place
  utils/map.py:35-42  Recursion       split while not judged uniform and min(h, w) > 5; children int(h/2) x int(w/2) at
                                      offsets {0, int(h/2)} x {0, int(w/2)} (an odd last row / column is never visited)
  utils/map.py:46-53  Division_Merge_Segmented   recursion, crop [1:-1, 1:-1], cv2.resize (bilinear)
  utils/map.py:56-60  laplacian       cv2.Laplacian(CV_16S, ksize=3) -> convertScaleAbs -> cv2.resize
  utils/distribution.py:5-16  cal_patch_score   int(mean) of every 16 x 16 block

Third-party arithmetic restated from OpenCV 4.x (opencv-python 4.13 in this image; the reference pins no version):
  * cv2.Laplacian ksize=3: the 3x3 kernel [[2,0,2],[0,-8,0],[2,0,2]], BORDER_REFLECT_101; convertScaleAbs = min(|v|, 255).
  * cv2.resize INTER_LINEAR on uint8: 11-bit fixed-point coefficients (cvRound((1-f)*2048), cvRound(f*2048), f a float),
    horizontal pass in int32 with the edge taps' fraction zeroed, vertical pass
    ((b0*(r0>>4))>>16) + ((b1*(r1>>4))>>16) + 2 >> 2 with clamped rows.
Pinned: `tests/test_scores_oracle.py` checks both restatements against cv2 itself (bit-exact over 30+ geometries incl.
up-scaling) and the whole generator against the reference's own functions executed from /root/reference on the Kodak
images and synthetic images, and against the committed reference-generated `tests/golden/kodak_scores.pt`.
"""
from __future__ import annotations

import numpy as np


def laplacian_abs_u8(img: np.ndarray) -> np.ndarray:
    """cv2.convertScaleAbs(cv2.Laplacian(img, cv2.CV_16S, ksize=3)) (utils/map.py:58-59)."""
    h, w = img.shape
    p = np.pad(img.astype(np.int64), 1, mode="reflect")          # BORDER_REFLECT_101
    lap = 2 * (p[0:h, 0:w] + p[0:h, 2:w + 2] + p[2:h + 2, 0:w] + p[2:h + 2, 2:w + 2]) - 8 * p[1:h + 1, 1:w + 1]
    return np.minimum(np.abs(lap), 255).astype(np.uint8)


def _resize_coeffs(ssize: int, dsize: int, zero_edges: bool):
    scale = 1.0 / (float(dsize) / float(ssize))                 # hal::resize: scale = 1. / inv_scale (double)
    idx = np.arange(dsize, dtype=np.float64)
    f = ((idx + 0.5) * scale - 0.5).astype(np.float32)          # float fx = (float)((dx + 0.5) * scale_x - 0.5)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(np.float32)).astype(np.float32)
    if zero_edges:                                              # the horizontal pass only
        lo = s < 0
        f[lo] = 0
        s[lo] = 0
        hi = s >= ssize - 1
        f[hi] = 0
        s[hi] = ssize - 1
    a0 = np.rint((np.float32(1.0) - f).astype(np.float32) * np.float32(2048)).astype(np.int64)
    a1 = np.rint(f * np.float32(2048)).astype(np.int64)
    return np.clip(s, 0, ssize - 1), np.clip(s + 1, 0, ssize - 1), a0, a1


def resize_linear_u8(src: np.ndarray, dw: int, dh: int) -> np.ndarray:
    """cv2.resize(src, (dw, dh)) for a uint8 single-channel image (default INTER_LINEAR)."""
    sh, sw = src.shape
    sx0, sx1, ax0, ax1 = _resize_coeffs(sw, dw, True)
    sy0, sy1, ay0, ay1 = _resize_coeffs(sh, dh, False)
    s = src.astype(np.int64)
    rows = s[:, sx0] * ax0[None, :] + s[:, sx1] * ax1[None, :]
    r0, r1 = rows[sy0], rows[sy1]
    out = (((ay0[:, None] * (r0 >> 4)) >> 16) + ((ay1[:, None] * (r1 >> 4)) >> 16) + 2) >> 2
    return out.astype(np.uint8)


def division_judge(area: np.ndarray) -> bool:
    """utils/map.py:6-23 with the pixel loop vectorised (same float64 values, same comparison)."""
    mean = np.mean(area)
    with np.errstate(invalid="ignore", divide="ignore"):
        std = np.std(area, ddof=1)
    operated = int(np.count_nonzero((area - mean) < 2 * std))
    return operated / area.size >= 0.95


def segment_inplace(img: np.ndarray) -> list:
    """Recursion + Merge of utils/map.py:27-42 on `img` in place.  Returns the visited leaves (h0, w0, h, w)."""
    leaves = []
    stack = [(0, 0, img.shape[0], img.shape[1])]
    while stack:
        h0, w0, h, w = stack.pop()
        area = img[h0:h0 + h, w0:w0 + w]
        if not division_judge(area) and min(h, w) > 5:
            hh, ww = int(h / 2), int(w / 2)
            stack += [(h0 + hh, w0 + ww, hh, ww), (h0 + hh, w0, hh, ww), (h0, w0 + ww, hh, ww), (h0, w0, hh, ww)]
        else:
            mask = (60 < area) & (area < 150)
            area[mask] = 0
            area[~mask] = 255
            leaves.append((h0, w0, h, w))
    return leaves


def cal_patch_score(img: np.ndarray, crop_sz: int = 16, step: int = 16) -> np.ndarray:
    """utils/distribution.py:5-16: int(mean) of every block (row-major over blocks)."""
    h, w = img.shape
    nh, nw = (h - crop_sz) // step + 1, (w - crop_sz) // step + 1
    assert crop_sz == step and nh * step <= h and nw * step <= w
    blocks = img[:nh * step, :nw * step].astype(np.int64).reshape(nh, step, nw, step).sum(axis=(1, 3))
    return (blocks // (crop_sz * crop_sz)).reshape(-1)


def generate_scores(gray: np.ndarray, out_side: int = 224, return_maps: bool = False):
    """generate_scores_file.py:19-31 for one grayscale uint8 image [H, W] -> float32 [ (out_side/16)^2 ]."""
    img = np.array(gray, dtype=np.uint8, copy=True)
    segment_inplace(img)                                              # Division_Merge_Segmented works in place (:21)
    s_map = resize_linear_u8(img[1:-1, 1:-1], out_side, out_side)
    t_map = resize_linear_u8(laplacian_abs_u8(img), out_side, out_side)   # :22 sees the segmented image
    total = cal_patch_score(t_map) * cal_patch_score(s_map)
    with np.errstate(invalid="ignore", divide="ignore"):
        norm = (total - total.min()) / (total.max() - total.min())      # float64; 0/0 -> NaN like the reference
    scores = norm.astype(np.float32)
    if return_maps:
        return scores, s_map, t_map, img
    return scores


def synthetic_gray(kind: int, h: int, w: int, seed: int) -> np.ndarray:
    """Seeded grayscale test images with natural-image-like structure (flat areas, edges, texture)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    if kind == 0:        # smooth gradients + a few rectangles + mild noise
        img = 110 + 60 * np.sin(xx / (17 + seed % 5)) * np.cos(yy / (23 + seed % 7))
        for _ in range(12 if min(h, w) >= 32 else 0):
            y0, x0 = rng.integers(0, h - 8), rng.integers(0, w - 8)
            hh, ww = rng.integers(8, h // 3), rng.integers(8, w // 3)
            img[y0:y0 + hh, x0:x0 + ww] = rng.integers(0, 256)
        img += rng.normal(0, 3, (h, w))
    elif kind == 1:      # pure noise
        img = rng.integers(0, 256, (h, w)).astype(np.float64)
    elif kind == 2:      # piecewise constant, 4 levels: two-valued regions, zero-variance nodes
        img = (rng.integers(0, 4, (h // 16 + 1, w // 16 + 1)) * 70).repeat(16, 0).repeat(16, 1)[:h, :w].astype(np.float64)
    elif kind == 3:      # constant
        img = np.full((h, w), float(seed % 256))
    else:                # text-like strokes on a bright background
        img = np.full((h, w), 235.0)
        for _ in range(200 if min(h, w) >= 32 else 0):
            y0, x0 = rng.integers(0, h - 4), rng.integers(0, w - 20)
            img[y0:y0 + rng.integers(1, 4), x0:x0 + rng.integers(4, 20)] = rng.integers(0, 90)
        img += rng.normal(0, 2, (h, w))
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)
