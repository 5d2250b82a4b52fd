"""Seeded synthetic checkpoint with the reference's state-dict names and shapes.

No trained checkpoint ships with the reference (`test.sh:3` expects a download), and the default
`MCM()` initialisation is degenerate for rate work (y-std 0.03 -> every symbol 0).  This
module produces a *non-degenerate* random state dict: same tensor names/shapes as
`MCM.state_dict()` for every tensor on the compression forward path
(/root/reference/models/Compression/MCM.py:71-93,115-293,300-323), values drawn from a seeded
torch CPU generator, then rescaled so that latents span several quantisation bins, scales spread
above the 0.11 floor and the factorized-prior medians are non-zero.

It is used by bench.py (random-init weights of the named architecture) and by the parity tests
(both the CUDA path and the oracle load the same dict).
"""
from __future__ import annotations

import math
from typing import Dict

import numpy as np
import torch

from .config import PathConfig


def sincos_pos_embed(embed_dim: int, grid_size: int) -> torch.Tensor:
    """Fixed 2-D sin-cos table with a zero cls row: fp32 [1, 1 + grid^2, embed_dim].

    Same construction as the reference (`common/pos_embed.py:23-94`, copied into the
    `encoder_pos_embed` parameter at MCM.py:457-464): float64 numpy, w-coordinate first,
    [sin | cos] halves, then cast to fp32."""
    assert embed_dim % 4 == 0
    gh = np.arange(grid_size, dtype=np.float32)
    gw = np.arange(grid_size, dtype=np.float32)
    grid = np.stack(np.meshgrid(gw, gh), axis=0).reshape(2, 1, grid_size, grid_size)

    def one_d(dim, pos):
        omega = np.arange(dim // 2, dtype=np.float64)
        omega /= dim / 2.0
        omega = 1.0 / 10000 ** omega
        out = np.einsum("m,d->md", pos.reshape(-1), omega)
        return np.concatenate([np.sin(out), np.cos(out)], axis=1)

    emb = np.concatenate([one_d(embed_dim // 2, grid[0]), one_d(embed_dim // 2, grid[1])], axis=1)
    emb = np.concatenate([np.zeros([1, embed_dim]), emb], axis=0)
    return torch.from_numpy(emb).float().unsqueeze(0)


# Gains chosen (once, by running the fp32 oracle on U[0,1) images) so that y-std ~ 3 bins, z-std ~ 3,
# mu-std ~ 1, sigma ~ U(0.3,2.3) +- 1: symbols in about -12..10, ~10 % of likelihoods at the 1e-9 floor,
# ~10 % of sigmas below the 0.11 bound (both clamps exercised).
GAIN = 1.5
LAST = {"g_a": 1.5, "h_a": 1.0, "h_s": 0.5, "cc_mean": 0.4, "cc_scale": 0.3, "lrp": 0.7}


def _uniform(gen, shape, bound):
    return (torch.rand(shape, generator=gen, dtype=torch.float32) * 2.0 - 1.0) * bound


def _linear(sd, name, cout, cin, gen, bias_scale=0.02):
    bound = math.sqrt(6.0 / (cin + cout))                     # xavier-uniform, as MCM._init_weights
    sd[name + ".weight"] = _uniform(gen, (cout, cin), bound)
    sd[name + ".bias"] = _uniform(gen, (cout,), bias_scale)


def _conv(sd, name, cout, cin, k, gen, gain=1.0):
    fan_in = cin * k * k
    bound = gain * math.sqrt(3.0 / fan_in)                    # unit-variance-preserving uniform
    sd[name + ".weight"] = _uniform(gen, (cout, cin, k, k), bound)
    sd[name + ".bias"] = _uniform(gen, (cout,), 0.05)


def make_decoder_state_dict(cfg: PathConfig, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Reconstruction-side tensors of the reference state dict: g_s (MCM.py:96-112, ConvTranspose2d
    1x1: weight [cin, cout, 1, 1]), decoder_embed / mask_token / decoder_pos_embed / decoder_blocks / decoder_norm /
    decoder_pred (MCM.py:325-354).  Drawn from its own generator so the hot-path tensors of `make_state_dict` (and the
    goldens frozen from them) do not depend on whether the decoder is requested."""
    g = torch.Generator().manual_seed(seed + 10007)
    sd: Dict[str, torch.Tensor] = {}
    E, Dd = cfg.encoder_embed_dim, cfg.decoder_embed_dim
    ch = cfg.g_a_channels()                                   # [E, c1, c2, Dd, Cy]; g_s walks it backwards
    gs = [ch[4], ch[3], ch[2], ch[1], ch[0]]
    for li, idx in enumerate((0, 2, 4, 6)):
        cin, cout = gs[li], gs[li + 1]
        bound = GAIN * math.sqrt(3.0 / cin)
        sd[f"g_s.{idx}.weight"] = _uniform(g, (cin, cout, 1, 1), bound)
        sd[f"g_s.{idx}.bias"] = _uniform(g, (cout,), 0.05)
    sd["g_s.0.weight"] *= 0.3                                  # y_hat has a std of a few bins
    _linear(sd, "decoder_embed", Dd, E, g)
    sd["mask_token"] = torch.randn(1, 1, Dd, generator=g) * 0.02
    sd["decoder_pos_embed"] = sincos_pos_embed(Dd, cfg.grid)
    hidden = int(Dd * cfg.mlp_ratio)
    for i in range(cfg.decoder_depth):
        pre = f"decoder_blocks.{i}"
        sd[pre + ".norm1.weight"] = 1.0 + _uniform(g, (Dd,), 0.1)
        sd[pre + ".norm1.bias"] = _uniform(g, (Dd,), 0.05)
        _linear(sd, pre + ".attn.qkv", 3 * Dd, Dd, g)
        sd[pre + ".attn.qkv.weight"] *= 2.0
        _linear(sd, pre + ".attn.proj", Dd, Dd, g)
        sd[pre + ".norm2.weight"] = 1.0 + _uniform(g, (Dd,), 0.1)
        sd[pre + ".norm2.bias"] = _uniform(g, (Dd,), 0.05)
        _linear(sd, pre + ".mlp.fc1", hidden, Dd, g)
        _linear(sd, pre + ".mlp.fc2", Dd, hidden, g)
    sd["decoder_norm.weight"] = 1.0 + _uniform(g, (Dd,), 0.1)
    sd["decoder_norm.bias"] = _uniform(g, (Dd,), 0.05)
    _linear(sd, "decoder_pred", cfg.patch_size ** 2 * cfg.in_chans, Dd, g)
    sd["decoder_pred.bias"] = sd["decoder_pred.bias"] + 0.5    # predictions centred in the [0, 1] pixel range
    return sd


def make_state_dict(cfg: PathConfig, seed: int = 0, include_decoder: bool = False) -> Dict[str, torch.Tensor]:
    """fp32 CPU tensors keyed by the reference's parameter names (hot-path subset; + the reconstruction side when
    `include_decoder`)."""
    cfg.validate()
    if include_decoder:
        sd = make_state_dict(cfg, seed)
        sd.update(make_decoder_state_dict(cfg, seed))
        return sd
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    C, D = cfg.encoder_embed_dim, cfg.encoder_depth
    p = cfg.patch_size

    # --- MAE encoder (MCM.py:300-323) ---
    sd["cls_token"] = torch.randn(1, 1, C, generator=g) * 0.02
    sd["encoder_pos_embed"] = sincos_pos_embed(C, cfg.grid)
    bound = math.sqrt(6.0 / (cfg.patch_dim + C))
    sd["encoder_embed.proj.weight"] = _uniform(g, (C, cfg.in_chans, p, p), bound * 4.0)
    sd["encoder_embed.proj.bias"] = _uniform(g, (C,), 0.02)
    for i in range(D):
        pre = f"encoder_blocks.{i}"
        sd[pre + ".norm1.weight"] = 1.0 + _uniform(g, (C,), 0.1)
        sd[pre + ".norm1.bias"] = _uniform(g, (C,), 0.05)
        _linear(sd, pre + ".attn.qkv", 3 * C, C, g)
        sd[pre + ".attn.qkv.weight"] *= 2.0                  # sharper (non-uniform) attention maps
        _linear(sd, pre + ".attn.proj", C, C, g)
        sd[pre + ".norm2.weight"] = 1.0 + _uniform(g, (C,), 0.1)
        sd[pre + ".norm2.bias"] = _uniform(g, (C,), 0.05)
        _linear(sd, pre + ".mlp.fc1", cfg.mlp_hidden, C, g)
        _linear(sd, pre + ".mlp.fc2", C, cfg.mlp_hidden, g)
    sd["encoder_norm.weight"] = 1.0 + _uniform(g, (C,), 0.1)
    sd["encoder_norm.bias"] = _uniform(g, (C,), 0.05)

    # --- g_a: 1x1 convs (MCM.py:77-93) ---
    ch = cfg.g_a_channels()
    for li, idx in enumerate((0, 2, 4, 6)):
        _conv(sd, f"g_a.{idx}", ch[li + 1], ch[li], 1, g, gain=GAIN)
    sd["g_a.6.weight"] *= LAST["g_a"]                         # y-std of a few bins (H3)

    # --- h_a (MCM.py:115-129) ---
    for (cin, cout, _s), idx in zip(cfg.h_a_layers(), (0, 2, 4, 6, 8)):
        _conv(sd, f"h_a.{idx}", cout, cin, 3, g, gain=GAIN)
    sd["h_a.8.weight"] *= LAST["h_a"]

    # --- h_s_mean / h_s_scale (MCM.py:132-162); subpel = Sequential(conv, PixelShuffle) -> ".N.0." ---
    for net in ("h_s_mean", "h_s_scale"):
        for (cin, cout, r), idx in zip(cfg.h_s_layers(), (0, 2, 4, 6, 8)):
            name = f"{net}.{idx}.0" if r > 1 else f"{net}.{idx}"
            _conv(sd, name, cout * r * r, cin, 3, g, gain=GAIN)
        sd[f"{net}.8.weight"] *= LAST["h_s"]

    # --- slice networks (MCM.py:165-293) ---
    for i in range(cfg.num_slices):
        for net, chans in (("cc_transform_mean", cfg.cc_channels(i)),
                           ("cc_transform_scale", cfg.cc_channels(i)),
                           ("lrp_transform", cfg.lrp_channels(i))):
            for li, idx in enumerate((0, 2, 4, 6, 8)):
                _conv(sd, f"{net}.{i}.{idx}", chans[li + 1], chans[li], 3, g, gain=GAIN)
        sd[f"cc_transform_mean.{i}.8.weight"] *= LAST["cc_mean"]
        sd[f"lrp_transform.{i}.8.weight"] *= LAST["lrp"]
        # sigma spread above the 0.11 floor, some below it
        sd[f"cc_transform_scale.{i}.8.bias"] = torch.rand(cfg.slice_ch, generator=g) * 2.0 + 0.3
        sd[f"cc_transform_scale.{i}.8.weight"] *= LAST["cc_scale"]

    # --- entropy bottleneck, compressai 1.2.4 naming (filters (3,3,3,3), init_scale 10) ---
    Cz = cfg.hyperprior_depth
    filters = (1, 3, 3, 3, 3, 1)
    scale = 10.0 ** (1.0 / 5.0)
    for i in range(5):
        init = float(np.log(np.expm1(1.0 / scale / filters[i + 1])))
        sd[f"entropy_bottleneck._matrix{i}"] = torch.full((Cz, filters[i + 1], filters[i]), init) \
            + torch.randn(Cz, filters[i + 1], filters[i], generator=g) * 0.3
        sd[f"entropy_bottleneck._bias{i}"] = _uniform(g, (Cz, filters[i + 1], 1), 0.5)
        if i < 4:
            sd[f"entropy_bottleneck._factor{i}"] = torch.randn(Cz, filters[i + 1], 1, generator=g) * 0.5
    q = torch.tensor([-10.0, 0.0, 10.0]).repeat(Cz, 1, 1)
    q[:, :, 1] = torch.randn(Cz, 1, generator=g)              # non-zero medians
    sd["entropy_bottleneck.quantiles"] = q
    return sd
