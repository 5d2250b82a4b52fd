"""Geometry of the compression forward path, derived from the reference constructor arguments.

Mirrors the channel arithmetic of `MCM.__init__`
(/root/reference/models/Compression/MCM.py:34-361) so that every layer shape the reference builds
is available without constructing torch modules.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Tuple


@dataclass(frozen=True)
class PathConfig:
    # same names / defaults as MCM.__init__ (MCM.py:34-52)
    img_size: int = 224
    patch_size: int = 16
    in_chans: int = 3
    encoder_embed_dim: int = 768
    encoder_depth: int = 12
    encoder_num_heads: int = 12
    decoder_embed_dim: int = 512
    decoder_depth: int = 8
    decoder_num_heads: int = 16
    mlp_ratio: float = 4.0
    latent_depth: int = 384
    hyperprior_depth: int = 192
    num_slices: int = 12
    num_keep_patches: int = 144
    ln_eps: float = 1e-6                      # norm_layer=partial(nn.LayerNorm, eps=1e-6), MCM.py:46

    # ---- derived -------------------------------------------------------------------------------
    @property
    def grid(self) -> int:
        return self.img_size // self.patch_size

    @property
    def num_patches(self) -> int:             # L
        return self.grid * self.grid

    @property
    def tokens(self) -> int:                  # T = K + 1 (cls)
        return self.num_keep_patches + 1

    @property
    def side(self) -> int:                    # s = sqrt(K)
        return int(round(math.sqrt(self.num_keep_patches)))

    @property
    def head_dim(self) -> int:
        return self.encoder_embed_dim // self.encoder_num_heads

    @property
    def mlp_hidden(self) -> int:
        return int(self.encoder_embed_dim * self.mlp_ratio)

    @property
    def patch_dim(self) -> int:               # in_chans * p * p, im2col order (c, p, q)
        return self.in_chans * self.patch_size * self.patch_size

    @property
    def slice_ch(self) -> int:
        return self.latent_depth // self.num_slices

    @property
    def max_support_slices(self) -> int:      # MCM.py:73
        return self.num_slices // 2

    def g_a_channels(self) -> List[int]:      # MCM.py:77-93
        e, d = self.encoder_embed_dim, self.decoder_embed_dim
        return [e, int(d + (e - d) * 3 / 4), int(d + (e - d) * 2 / 4), d, self.latent_depth]

    def h_a_layers(self) -> List[Tuple[int, int, int]]:   # (cin, cout, stride)  MCM.py:115-129
        m, h = self.latent_depth, self.hyperprior_depth
        c1 = int(h + (m - h) * 3 / 4)
        c2 = int(h + (m - h) * 2 / 4)
        c3 = int(h + (m - h) / 4)
        return [(m, m, 1), (m, c1, 1), (c1, c2, 2), (c2, c3, 1), (c3, h, 2)]

    def h_s_layers(self) -> List[Tuple[int, int, int]]:   # (cin, cout, upscale r)  MCM.py:132-162
        m, h = self.latent_depth, self.hyperprior_depth
        c1 = int(h + (m - h) / 4)
        c2 = int(h + (m - h) * 2 / 4)
        c3 = int(h + (m - h) * 3 / 4)
        return [(h, c1, 1), (c1, c2, 2), (c2, c3, 1), (c3, m, 2), (m, m, 1)]

    def cc_channels(self, i: int) -> List[int]:           # MCM.py:165-249
        sc, ns = self.slice_ch, self.num_slices
        cin = int(self.latent_depth + sc * min(i, ns // 2))
        return [cin, int(sc * (ns // 2 + 1)), int(sc * (ns // 2 * 3 / 4 + 1)),
                int(sc * (ns // 2 * 2 / 4 + 1)), int(sc * (ns // 2 * 1 / 4 + 1)), int(sc)]

    def lrp_channels(self, i: int) -> List[int]:          # MCM.py:252-293
        sc, ns = self.slice_ch, self.num_slices
        cin = int(self.latent_depth + sc * min(i + 1, ns // 2 + 1))
        return [cin] + self.cc_channels(i)[1:]

    def validate(self) -> None:
        """Same failure classes as the reference (SURVEY 8b 'Errors')."""
        if self.num_keep_patches > self.num_patches:
            # MCM.py:374-376
            raise ValueError("Number of patches should not be greater than the length of scores")
        s = self.side
        if s * s != self.num_keep_patches:
            # MCM.py:729-732 view(-1, sqrt(K), sqrt(K), C) fails
            raise RuntimeError(f"num_keep_patches={self.num_keep_patches} is not a perfect square "
                               f"(reference: view() shape mismatch at MCM.py:729)")
        if s % 4 != 0:
            # h_a halves twice, h_s doubles twice; cat at MCM.py:761/780 needs equal sizes
            raise RuntimeError(f"sqrt(num_keep_patches)={s} must be a multiple of 4 "
                               f"(reference: torch.cat size mismatch at MCM.py:761)")
        if self.img_size % self.patch_size != 0:
            raise AssertionError("image size must be divisible by patch size")
        if self.encoder_embed_dim % self.encoder_num_heads != 0:
            raise AssertionError("embed dim must be divisible by heads")


def vit_base(num_keep_patches: int = 144, img_size: int = 224) -> PathConfig:
    return PathConfig(img_size=img_size, num_keep_patches=num_keep_patches)


def vit_large(num_keep_patches: int = 256, img_size: int = 512) -> PathConfig:
    return PathConfig(img_size=img_size, encoder_embed_dim=1024, encoder_depth=24, encoder_num_heads=16,
                      num_keep_patches=num_keep_patches)
