"""Tiling policy for inputs that are not `img_size` x `img_size` (an EXTENSION: the reference resizes every image to 224 x 224,
utils/dataloader.py:69-73, and `MCM` asserts the size).  An image is zero-padded on the bottom / right to the
next multiple of the tile side and cut into row-major tiles; every tile is an independent unit of the path (own scores, own
mask, own rate), so tiles shard across GPUs like images do (`distributed.shard_round_robin`).  The rate of the image is the
sum of its tiles' bits over the ORIGINAL pixel count.

BASELINE configs: Kodak 768x512 -> 896x672 -> 4 x 3 = 12 tiles of 224 (config 3); 2048x1080 -> 2048x1536 -> 4 x 3 = 12 tiles
of 512 (config 5)."""
from __future__ import annotations

from typing import Tuple

import torch


def tile_grid(height: int, width: int, tile: int) -> Tuple[int, int]:
    """(rows, cols) of tiles covering a height x width image."""
    return (height + tile - 1) // tile, (width + tile - 1) // tile


def tile_image(img: torch.Tensor, tile: int) -> torch.Tensor:
    """img [C, H, W] (or [N, C, H, W]) -> tiles [rows*cols, C, tile, tile] (or [N*rows*cols, ...]), row-major, zero padded."""
    squeeze = img.dim() == 3
    if squeeze:
        img = img.unsqueeze(0)
    n, c, h, w = img.shape
    rows, cols = tile_grid(h, w, tile)
    padded = img.new_zeros((n, c, rows * tile, cols * tile))
    padded[:, :, :h, :w] = img
    t = padded.reshape(n, c, rows, tile, cols, tile).permute(0, 2, 4, 1, 3, 5).reshape(n * rows * cols, c, tile, tile)
    return t.contiguous()


def untile_image(tiles: torch.Tensor, height: int, width: int) -> torch.Tensor:
    """Inverse of `tile_image` for ONE image: tiles [rows*cols, C, tile, tile] -> [C, height, width] (padding dropped)."""
    tile = tiles.shape[-1]
    rows, cols = tile_grid(height, width, tile)
    c = tiles.shape[1]
    full = tiles.reshape(rows, cols, c, tile, tile).permute(2, 0, 3, 1, 4).reshape(c, rows * tile, cols * tile)
    return full[:, :height, :width].contiguous()


def image_bpp(tile_bpp: torch.Tensor, height: int, width: int, tile: int) -> torch.Tensor:
    """Rate of the whole image from its tiles' per-tile bpp (each over tile*tile pixels): total bits / original pixels."""
    return tile_bpp.double().sum() * (tile * tile) / float(height * width)
