"""ORACLE (test infrastructure only - never imported by the product path).

Stand-ins for the third-party LEAF classes the reference's `MCM.py` imports (MCM.py:8-14) and that are
absent from this image with no network: timm 0.4.5 `PatchEmbed` / `Block`, compressai 1.2.4
`EntropyBottleneck` / `GaussianConditional` / `conv3x3` / `subpel_conv3x3` / `quantize_ste` /
`CompressionModel`, pytorch_msssim `SSIM`.  Each is restated from the package's published definition
(same constructor signature, parameter names and eval-mode arithmetic) so that the reference file itself -
`MCM.__init__` topology and channel arithmetic, `forward_encoder`, `random_masking`, `forward`,
`forward_decoder`, `unpatchify`, `forward_loss` - can be EXECUTED VERBATIM from /root/reference
(oracle/ref_exec.py).  Nothing here restates a line of MCM.py.

Not covered (cannot be, no network): the pretrained VGG16 feature loss (loss/vgg.py:99 downloads weights
and calls .cuda()); `cal_features_loss` is replaced by a zero, and the VGG term is reported as 0.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F


# ---- timm 0.4.5 models/vision_transformer.py ---------------------------------------------------------
class Mlp(nn.Module):
    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.0):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop = nn.Dropout(drop)

    def forward(self, x):
        return self.drop(self.fc2(self.drop(self.act(self.fc1(x)))))


class Attention(nn.Module):
    def __init__(self, dim, num_heads=8, qkv_bias=False, qk_scale=None, attn_drop=0.0, proj_drop=0.0):
        super().__init__()
        self.num_heads = num_heads
        head_dim = dim // num_heads
        self.scale = qk_scale or head_dim ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)

    def forward(self, x):
        B, N, C = x.shape
        qkv = self.qkv(x).reshape(B, N, 3, self.num_heads, C // self.num_heads).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0], qkv[1], qkv[2]
        attn = (q @ k.transpose(-2, -1)) * self.scale
        attn = self.attn_drop(attn.softmax(dim=-1))
        x = (attn @ v).transpose(1, 2).reshape(B, N, C)
        return self.proj_drop(self.proj(x))


class Block(nn.Module):
    def __init__(self, dim, num_heads, mlp_ratio=4.0, qkv_bias=False, qk_scale=None, drop=0.0, attn_drop=0.0,
                 drop_path=0.0, act_layer=nn.GELU, norm_layer=nn.LayerNorm):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = Attention(dim, num_heads=num_heads, qkv_bias=qkv_bias, qk_scale=qk_scale, attn_drop=attn_drop,
                              proj_drop=drop)
        self.drop_path = nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)

    def forward(self, x):
        x = x + self.drop_path(self.attn(self.norm1(x)))
        x = x + self.drop_path(self.mlp(self.norm2(x)))
        return x


class PatchEmbed(nn.Module):
    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768, norm_layer=None):
        super().__init__()
        img_size = (img_size, img_size)
        patch_size = (patch_size, patch_size)
        self.img_size = img_size
        self.patch_size = patch_size
        self.grid_size = (img_size[0] // patch_size[0], img_size[1] // patch_size[1])
        self.num_patches = self.grid_size[0] * self.grid_size[1]
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)
        self.norm = norm_layer(embed_dim) if norm_layer else nn.Identity()

    def forward(self, x):
        B, C, H, W = x.shape
        assert H == self.img_size[0] and W == self.img_size[1], \
            f"Input image size ({H}*{W}) doesn't match model ({self.img_size[0]}*{self.img_size[1]})."
        x = self.proj(x).flatten(2).transpose(1, 2)
        return self.norm(x)


# ---- compressai 1.2.4 ------------------------------------------------------------------------------
def quantize_ste(x):
    """compressai.ops.quantize_ste"""
    return (torch.round(x) - x).detach() + x


def conv3x3(in_ch, out_ch, stride=1):
    """compressai.layers.conv3x3"""
    return nn.Conv2d(in_ch, out_ch, kernel_size=3, stride=stride, padding=1)


def subpel_conv3x3(in_ch, out_ch, r=1):
    """compressai.layers.subpel_conv3x3"""
    return nn.Sequential(nn.Conv2d(in_ch, out_ch * r ** 2, kernel_size=3, padding=1), nn.PixelShuffle(r))


class _EntropyModel(nn.Module):
    def __init__(self, likelihood_bound=1e-9):
        super().__init__()
        self.likelihood_bound = float(likelihood_bound)

    def quantize(self, inputs, mode, means=None):
        if mode == "noise":
            half = 0.5
            return inputs + torch.empty_like(inputs).uniform_(-half, half)
        outputs = inputs.clone()
        if means is not None:
            outputs = outputs - means
        outputs = torch.round(outputs)
        if mode == "dequantize":
            if means is not None:
                outputs = outputs + means
            return outputs
        assert mode == "symbols", mode
        return outputs.int()


class EntropyBottleneck(_EntropyModel):
    """compressai.entropy_models.EntropyBottleneck (factorized prior), training / eval forward."""

    def __init__(self, channels, tail_mass=1e-9, init_scale=10, filters=(3, 3, 3, 3)):
        super().__init__()
        self.channels = int(channels)
        self.filters = tuple(int(f) for f in filters)
        self.init_scale = float(init_scale)
        self.tail_mass = float(tail_mass)
        filters = (1,) + self.filters + (1,)
        scale = self.init_scale ** (1 / (len(self.filters) + 1))
        channels = self.channels
        for i in range(len(self.filters) + 1):
            init = math.log(math.expm1(1 / scale / filters[i + 1]))
            matrix = torch.Tensor(channels, filters[i + 1], filters[i])
            matrix.data.fill_(init)
            self.register_parameter(f"_matrix{i:d}", nn.Parameter(matrix))
            bias = torch.Tensor(channels, filters[i + 1], 1)
            nn.init.uniform_(bias, -0.5, 0.5)
            self.register_parameter(f"_bias{i:d}", nn.Parameter(bias))
            if i < len(self.filters):
                factor = torch.Tensor(channels, filters[i + 1], 1)
                nn.init.zeros_(factor)
                self.register_parameter(f"_factor{i:d}", nn.Parameter(factor))
        self.quantiles = nn.Parameter(torch.Tensor(channels, 1, 3))
        init = torch.Tensor([-self.init_scale, 0, self.init_scale])
        self.quantiles.data = init.repeat(self.quantiles.size(0), 1, 1)
        target = math.log(2 / self.tail_mass - 1)
        self.register_buffer("target", torch.Tensor([-target, 0, target]), persistent=False)

    def _get_medians(self):
        return self.quantiles[:, :, 1:2]

    def _logits_cumulative(self, inputs, stop_gradient=False):
        logits = inputs
        for i in range(len(self.filters) + 1):
            matrix = getattr(self, f"_matrix{i:d}")
            logits = torch.matmul(F.softplus(matrix), logits)
            bias = getattr(self, f"_bias{i:d}")
            logits = logits + bias
            if i < len(self.filters):
                factor = getattr(self, f"_factor{i:d}")
                logits = logits + torch.tanh(factor) * torch.tanh(logits)
        return logits

    def _likelihood(self, inputs):
        half = 0.5
        lower = self._logits_cumulative(inputs - half)
        upper = self._logits_cumulative(inputs + half)
        sign = -torch.sign(lower + upper).detach()
        return torch.abs(torch.sigmoid(sign * upper) - torch.sigmoid(sign * lower))

    def loss(self):
        logits = self._logits_cumulative(self.quantiles, stop_gradient=True)
        return torch.abs(logits - self.target).sum()

    def forward(self, x, training=None):
        if training is None:
            training = self.training
        perm = [1, 0] + list(range(2, x.dim()))
        inv_perm = perm                                   # (1, 0, 2, 3) is its own inverse
        x = x.permute(*perm).contiguous()
        shape = x.size()
        values = x.reshape(x.size(0), 1, -1)
        outputs = self.quantize(values, "noise" if training else "dequantize", self._get_medians())
        likelihood = self._likelihood(outputs)
        likelihood = torch.clamp_min(likelihood, self.likelihood_bound)     # LowerBound in eval == clamp
        outputs = outputs.reshape(shape).permute(*inv_perm).contiguous()
        likelihood = likelihood.reshape(shape).permute(*inv_perm).contiguous()
        return outputs, likelihood


class GaussianConditional(_EntropyModel):
    """compressai.entropy_models.GaussianConditional(scale_table=None, scale_bound=0.11, tail_mass=1e-9)."""

    def __init__(self, scale_table, scale_bound=0.11, tail_mass=1e-9):
        super().__init__()
        self.tail_mass = float(tail_mass)
        self.scale_bound = 0.11 if scale_bound is None else float(scale_bound)

    @staticmethod
    def _standardized_cumulative(inputs):
        half = float(0.5)
        const = float(-(2 ** -0.5))
        return half * torch.erfc(const * inputs)

    def _likelihood(self, inputs, scales, means=None):
        half = float(0.5)
        values = inputs - means if means is not None else inputs
        scales = torch.clamp_min(scales, self.scale_bound)
        values = torch.abs(values)
        upper = self._standardized_cumulative((half - values) / scales)
        lower = self._standardized_cumulative((-half - values) / scales)
        return upper - lower

    def forward(self, inputs, scales, means=None, training=None):
        if training is None:
            training = self.training
        outputs = self.quantize(inputs, "noise" if training else "dequantize", means)
        likelihood = self._likelihood(outputs, scales, means)
        likelihood = torch.clamp_min(likelihood, self.likelihood_bound)
        return outputs, likelihood


class CompressionModel(nn.Module):
    """compressai.models.CompressionModel, the part MCM relies on."""

    def aux_loss(self):
        return sum(m.loss() for m in self.modules() if isinstance(m, EntropyBottleneck))

    def update(self, force=False):
        raise NotImplementedError("CDF table construction needs compressai's C++ extension")


class BufferedRansEncoder:          # compressai.ans: C++ extension, never reached on the forward path
    def __init__(self, *a, **k):
        raise NotImplementedError


class RansDecoder(BufferedRansEncoder):
    pass


# ---- pytorch_msssim ---------------------------------------------------------------------------------
def _fspecial_gauss_1d(size, sigma):
    coords = torch.arange(size, dtype=torch.float)
    coords -= size // 2
    g = torch.exp(-(coords ** 2) / (2 * sigma ** 2))
    g /= g.sum()
    return g.unsqueeze(0).unsqueeze(0)


def _gaussian_filter(x, win):
    C = x.shape[1]
    out = x
    for i, s in enumerate(x.shape[2:]):
        if s >= win.shape[-1]:
            out = F.conv2d(out, weight=win.transpose(2 + i, -1), stride=1, padding=0, groups=C)
    return out


def ssim_value(X, Y, data_range=1.0, win_size=11, win_sigma=1.5, size_average=True, K=(0.01, 0.03)):
    """pytorch_msssim.ssim for 4-D inputs (nonnegative_ssim=False)."""
    win = _fspecial_gauss_1d(win_size, win_sigma).repeat([X.shape[1]] + [1] * (len(X.shape) - 1)).to(X.device, X.dtype)
    K1, K2 = K
    C1, C2 = (K1 * data_range) ** 2, (K2 * data_range) ** 2
    mu1, mu2 = _gaussian_filter(X, win), _gaussian_filter(Y, win)
    mu1_sq, mu2_sq, mu1_mu2 = mu1.pow(2), mu2.pow(2), mu1 * mu2
    sigma1_sq = _gaussian_filter(X * X, win) - mu1_sq
    sigma2_sq = _gaussian_filter(Y * Y, win) - mu2_sq
    sigma12 = _gaussian_filter(X * Y, win) - mu1_mu2
    cs_map = (2 * sigma12 + C2) / (sigma1_sq + sigma2_sq + C2)
    ssim_map = ((2 * mu1_mu2 + C1) / (mu1_sq + mu2_sq + C1)) * cs_map
    ssim_per_channel = torch.flatten(ssim_map, 2).mean(-1)
    return ssim_per_channel.mean() if size_average else ssim_per_channel.mean(1)


class SSIM(nn.Module):
    def __init__(self, data_range=255, size_average=True, win_size=11, win_sigma=1.5, channel=3, spatial_dims=2,
                 K=(0.01, 0.03), nonnegative_ssim=False):
        super().__init__()
        self.win_size, self.win_sigma = win_size, win_sigma
        self.size_average, self.data_range, self.K = size_average, data_range, K

    def forward(self, X, Y):
        return ssim_value(X, Y, data_range=self.data_range, win_size=self.win_size, win_sigma=self.win_sigma,
                          size_average=self.size_average, K=self.K)
