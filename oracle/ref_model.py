"""ORACLE (test infrastructure only - never imported by the product path).

CPU restatement, in plain torch fp32 (or fp64), of the reference's compression forward path:
`MCM.forward_encoder` + the rate half of `MCM.forward`
(/root/reference/models/Compression/MCM.py:590-634, 714-787) and the caller-side rate
(/root/reference/models/Compression/loss/rd_loss.py:15-20).

The reference module cannot be imported here (timm 0.4.5 / compressai 1.2.4 / pytorch_msssim are
absent and there is no network), so the arithmetic that lives in those packages is
restated from their published definitions, anchored on the reference's call sites:

  timm 0.4.5  PatchEmbed  (MCM.py:300-302, 615): Conv2d(k=p, s=p) -> flatten(2).transpose(1, 2)
              Block       (MCM.py:313-322, 629-630): x += proj(softmax(q k^T * hd^-0.5) v);
                          x += fc2(gelu_erf(fc1(LN(x)))); qkv reshape (B, T, 3, H, hd)
  compressai 1.2.4  EntropyBottleneck.forward (eval) (MCM.py:741): v = round(z - med) + med,
                          lik = max(|sigmoid(s*U) - sigmoid(s*Lo)|, 1e-9), U/Lo = logits(v +- 0.5),
                          logits = 5 x (softplus(M) @ x + b [+ tanh(f) * tanh(x)]), s = -sign(Lo + U)
              GaussianConditional.forward (eval) (MCM.py:771-772): v = round(y - mu) + mu,
                          lik = max(Phi((.5 - |v - mu|)/s) - Phi((-.5 - |v - mu|)/s), 1e-9),
                          s = max(sigma, 0.11), Phi(x) = 0.5 erfc(-x / sqrt 2)
              quantize_ste(x) = round(x) in the forward pass (MCM.py:744, 776)
              conv3x3 = Conv2d(k3, pad 1, stride), subpel_conv3x3 = conv(cin, cout*r^2) + PixelShuffle(r)

Parity status: the reference has no tests, golden vectors or checkpoints, so this
restatement is **pinned only where reference code can execute here**: the patch ordering is checked
against the verbatim `MCM.get_ids_shuffle` (oracle/ref_mask.py, tests/test_mask_oracle.py).  For
the transformer / entropy-model arithmetic: **parity unpinned** against reference outputs; it is
pinned against independent torch.nn modules built from the same published definitions
(tests/test_oracle_model.py) and frozen in tests/golden/*.pt.

Every function takes a state dict with the reference's parameter names.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

from . import ref_mask

LIKELIHOOD_BOUND = 1e-9          # compressai EntropyModel(likelihood_bound=1e-9)
SCALE_BOUND = 0.11               # compressai GaussianConditional(scale_bound=0.11)


def _w(sd, name, dtype):
    return sd[name].to(dtype)


# ------------------------------------------------------------------------------------------------
# encoder
# ------------------------------------------------------------------------------------------------
def patch_embed(sd, cfg, imgs, dtype):
    """timm PatchEmbed.forward (MCM.py:615): [N,3,S,S] -> [N,L,C]."""
    x = F.conv2d(imgs.to(dtype), _w(sd, "encoder_embed.proj.weight", dtype),
                 _w(sd, "encoder_embed.proj.bias", dtype), stride=cfg.patch_size)
    return x.flatten(2).transpose(1, 2)


def vit_block(sd, cfg, i, x, dtype):
    """timm 0.4.5 Block.forward, eval mode (no dropout / drop-path)."""
    pre = f"encoder_blocks.{i}"
    C, H = cfg.encoder_embed_dim, cfg.encoder_num_heads
    hd = C // H
    B, T, _ = x.shape
    h = F.layer_norm(x, (C,), _w(sd, pre + ".norm1.weight", dtype), _w(sd, pre + ".norm1.bias", dtype), cfg.ln_eps)
    qkv = F.linear(h, _w(sd, pre + ".attn.qkv.weight", dtype), _w(sd, pre + ".attn.qkv.bias", dtype))
    qkv = qkv.reshape(B, T, 3, H, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    attn = (q @ k.transpose(-2, -1)) * (hd ** -0.5)
    attn = attn.softmax(dim=-1)
    a = (attn @ v).transpose(1, 2).reshape(B, T, C)
    x = x + F.linear(a, _w(sd, pre + ".attn.proj.weight", dtype), _w(sd, pre + ".attn.proj.bias", dtype))
    h = F.layer_norm(x, (C,), _w(sd, pre + ".norm2.weight", dtype), _w(sd, pre + ".norm2.bias", dtype), cfg.ln_eps)
    h = F.gelu(F.linear(h, _w(sd, pre + ".mlp.fc1.weight", dtype), _w(sd, pre + ".mlp.fc1.bias", dtype)))
    x = x + F.linear(h, _w(sd, pre + ".mlp.fc2.weight", dtype), _w(sd, pre + ".mlp.fc2.bias", dtype))
    return x


def forward_encoder(sd, cfg, imgs, total_scores, dtype=torch.float32, ids_shuffle=None, taps=None):
    """MCM.forward_encoder (MCM.py:590-634).  Returns (x_remain [N,K,C], ids_restore, ids_keep)."""
    x = patch_embed(sd, cfg, imgs, dtype)
    pos = _w(sd, "encoder_pos_embed", dtype)
    x = x + pos[:, 1:, :]                                              # :618
    if ids_shuffle is None:
        ids_shuffle = ref_mask.ids_shuffle_spec_batch(total_scores, cfg.num_keep_patches)   # :573
    ids_keep, ids_restore = ref_mask.masking_from_shuffle(ids_shuffle, cfg.num_keep_patches)
    D = x.shape[-1]
    x = torch.gather(x, 1, ids_keep.unsqueeze(-1).repeat(1, 1, D))    # :585-586
    cls = _w(sd, "cls_token", dtype) + pos[:, :1, :]                   # :624
    x = torch.cat((cls.expand(x.shape[0], -1, -1), x), dim=1)         # :626
    if taps is not None:
        taps["tokens_in"] = x.clone()
    for i in range(cfg.encoder_depth):                                 # :629-630
        x = vit_block(sd, cfg, i, x, dtype)
        if taps is not None and i == 0:
            taps["block0_out"] = x.clone()
    C = cfg.encoder_embed_dim
    x = F.layer_norm(x, (C,), _w(sd, "encoder_norm.weight", dtype), _w(sd, "encoder_norm.bias", dtype), cfg.ln_eps)
    return x[:, 1:, :], ids_restore, ids_keep                           # :631-634


# ------------------------------------------------------------------------------------------------
# entropy model pieces
# ------------------------------------------------------------------------------------------------
def _seq_convs(sd, prefix, x, dtype, strides=(1, 1, 1, 1, 1), idxs=(0, 2, 4, 6, 8), k3=True):
    """nn.Sequential(conv, GELU, conv, GELU, ..., conv): GELU between, none after the last."""
    n = len(idxs)
    for li, idx in enumerate(idxs):
        x = F.conv2d(x, _w(sd, f"{prefix}.{idx}.weight", dtype), _w(sd, f"{prefix}.{idx}.bias", dtype),
                     stride=strides[li], padding=1 if k3 else 0)
        if li < n - 1:
            x = F.gelu(x)
    return x


def g_a(sd, cfg, y, dtype):                                            # MCM.py:77-93, 735
    return _seq_convs(sd, "g_a", y, dtype, strides=(1, 1, 1, 1), idxs=(0, 2, 4, 6), k3=False)


def h_a(sd, cfg, y, dtype):                                            # MCM.py:115-129, 739
    return _seq_convs(sd, "h_a", y, dtype, strides=tuple(s for _, _, s in cfg.h_a_layers()))


def h_s(sd, cfg, net, z_hat, dtype):                                   # MCM.py:132-162, 747-748
    x = z_hat
    layers = cfg.h_s_layers()
    for li, ((_cin, _cout, r), idx) in enumerate(zip(layers, (0, 2, 4, 6, 8))):
        name = f"{net}.{idx}.0" if r > 1 else f"{net}.{idx}"
        x = F.conv2d(x, _w(sd, name + ".weight", dtype), _w(sd, name + ".bias", dtype), padding=1)
        if r > 1:
            x = F.pixel_shuffle(x, r)
        if li < len(layers) - 1:
            x = F.gelu(x)
    return x


def eb_logits_cumulative(sd, x, dtype):
    """compressai EntropyBottleneck._logits_cumulative; x: [C, 1, M]."""
    logits = x
    for i in range(5):
        m = F.softplus(_w(sd, f"entropy_bottleneck._matrix{i}", dtype))
        logits = torch.matmul(m, logits)
        logits = logits + _w(sd, f"entropy_bottleneck._bias{i}", dtype)
        if i < 4:
            f = _w(sd, f"entropy_bottleneck._factor{i}", dtype)
            logits = logits + torch.tanh(f) * torch.tanh(logits)
    return logits


def entropy_bottleneck_eval(sd, z, dtype):
    """compressai EntropyBottleneck.forward(training=False) -> (z_hat, likelihood), both [N,C,h,w]."""
    N, C = z.shape[:2]
    med = _w(sd, "entropy_bottleneck.quantiles", dtype)[:, :, 1:2]      # _get_medians(): [C,1,1]
    v = z.permute(1, 0, 2, 3).contiguous()
    shape = v.shape
    v = v.reshape(C, 1, -1)
    out = torch.round(v - med) + med                                   # quantize("dequantize", medians)
    lower = eb_logits_cumulative(sd, out - 0.5, dtype)
    upper = eb_logits_cumulative(sd, out + 0.5, dtype)
    sign = -torch.sign(lower + upper)
    lik = torch.abs(torch.sigmoid(sign * upper) - torch.sigmoid(sign * lower))
    lik = torch.clamp_min(lik, LIKELIHOOD_BOUND)
    out = out.reshape(shape).permute(1, 0, 2, 3).contiguous()
    lik = lik.reshape(shape).permute(1, 0, 2, 3).contiguous()
    return out, lik


def gaussian_conditional_eval(y, sigma, mu):
    """compressai GaussianConditional.forward(training=False) -> (y_dequant, likelihood)."""
    out = torch.round(y - mu) + mu
    values = torch.abs(out - mu)
    s = torch.clamp_min(sigma, SCALE_BOUND)
    c = -(2 ** -0.5)
    upper = 0.5 * torch.erfc(c * ((0.5 - values) / s))
    lower = 0.5 * torch.erfc(c * ((-0.5 - values) / s))
    lik = torch.clamp_min(upper - lower, LIKELIHOOD_BOUND)
    return out, lik


# ------------------------------------------------------------------------------------------------
# full path
# ------------------------------------------------------------------------------------------------
def rate_from_latent(sd, cfg, y, dtype=torch.float32, force: Optional[Dict[str, torch.Tensor]] = None):
    """MCM.forward lines 739-787 given y = g_a(...) [N,Cy,s,s].  Returns a dict of every intermediate."""
    out: Dict[str, torch.Tensor] = {"y": y}
    z = h_a(sd, cfg, y, dtype)                                         # :739
    out["z"] = z
    _, z_lik = entropy_bottleneck_eval(sd, z, dtype)                   # :741
    med = _w(sd, "entropy_bottleneck.quantiles", dtype)[:, :, 1:2]      # [C,1,1] broadcasts per channel
    z_sym = torch.round(z - med)
    z_hat = z_sym + med                                                # :742-744
    out.update(z_lik=z_lik, z_sym=z_sym.to(torch.int32), z_hat=z_hat)
    latent_scales = h_s(sd, cfg, "h_s_scale", z_hat, dtype)            # :747
    latent_means = h_s(sd, cfg, "h_s_mean", z_hat, dtype)              # :748
    out.update(latent_scales=latent_scales, latent_means=latent_means)
    s0, s1 = y.shape[2:]
    y_slices = y.chunk(cfg.num_slices, 1)
    y_hat_slices, liks, mus, sigmas, syms, y_hat_pre = [], [], [], [], [], []
    for i, y_slice in enumerate(y_slices):                             # :755-784
        support = y_hat_slices[: cfg.max_support_slices]
        mean_support = torch.cat([latent_means] + support, dim=1)
        mu = _seq_convs(sd, f"cc_transform_mean.{i}", mean_support, dtype)[:, :, :s0, :s1]
        scale_support = torch.cat([latent_scales] + support, dim=1)
        sigma = _seq_convs(sd, f"cc_transform_scale.{i}", scale_support, dtype)[:, :, :s0, :s1]
        _, lik = gaussian_conditional_eval(y_slice, sigma, mu)          # :771-772
        sym = torch.round(y_slice - mu)
        y_hat_slice = sym + mu                                          # :776
        y_hat_pre.append(y_hat_slice.clone())
        lrp_support = torch.cat([mean_support, y_hat_slice], dim=1)     # :780
        lrp = _seq_convs(sd, f"lrp_transform.{i}", lrp_support, dtype)
        y_hat_slice = y_hat_slice + 0.5 * torch.tanh(lrp)               # :782-783
        y_hat_slices.append(y_hat_slice)
        liks.append(lik); mus.append(mu); sigmas.append(sigma); syms.append(sym.to(torch.int32))
    out.update(y_hat=torch.cat(y_hat_slices, 1), y_lik=torch.cat(liks, 1), mu=torch.cat(mus, 1),
               sigma=torch.cat(sigmas, 1), y_sym=torch.cat(syms, 1), y_hat_pre=torch.cat(y_hat_pre, 1))
    return out


def bpp_per_image(y_lik, z_lik, img_size):
    """rd_loss.py:15-20 with N = 1 per image: -(sum ln lik_y + sum ln lik_z) / (ln 2 * H * W)."""
    n = y_lik.shape[0]
    num_pixels = img_size * img_size
    tot = torch.log(y_lik.double()).reshape(n, -1).sum(1) + torch.log(z_lik.double()).reshape(n, -1).sum(1)
    return (tot / (-math.log(2) * num_pixels)).float()


def bpp_batch(y_lik, z_lik, img_size):
    """rd_loss.py:15-20 verbatim (batch aggregate, fp32 like the reference)."""
    num_pixels = y_lik.shape[0] * img_size * img_size
    return sum(torch.log(l).sum() / (-math.log(2) * num_pixels) for l in (y_lik, z_lik))


def forward_rate(sd, cfg, imgs, total_scores, dtype=torch.float32, ids_shuffle=None, taps=None):
    """imgs [N,3,S,S], total_scores [N,L] -> dict with the reference's `likelihoods` plus every
    intermediate on the path (MCM.py:714-787)."""
    cfg.validate()
    if imgs.shape[-1] != cfg.img_size or imgs.shape[-2] != cfg.img_size:
        raise AssertionError(f"Input image size ({imgs.shape[-2]}*{imgs.shape[-1]}) doesn't match model "
                             f"({cfg.img_size}*{cfg.img_size}).")           # timm PatchEmbed assert
    with torch.no_grad():
        x_remain, ids_restore, ids_keep = forward_encoder(sd, cfg, imgs, total_scores, dtype, ids_shuffle, taps)
        s = cfg.side
        y_in = x_remain.reshape(-1, s, s, cfg.encoder_embed_dim).permute(0, 3, 1, 2).contiguous()   # :729-732
        y = g_a(sd, cfg, y_in, dtype)                                   # :735  (.float() is a no-op in fp32)
        out = rate_from_latent(sd, cfg, y, dtype)
        out.update(x_remain=x_remain, ids_restore=ids_restore, ids_keep=ids_keep)
        out["likelihoods"] = {"y": out["y_lik"], "z": out["z_lik"]}      # :801
        out["bpp"] = bpp_per_image(out["y_lik"], out["z_lik"], cfg.img_size)
        out["bpp_batch"] = bpp_batch(out["y_lik"], out["z_lik"], cfg.img_size)
    return out
