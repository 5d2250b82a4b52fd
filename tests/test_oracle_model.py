"""CPU: pins the functional oracle (oracle/ref_model.py) two ways:
 (a) against an independent nn.Module build of the same published definitions (timm 0.4.5 Block / PatchEmbed,
     compressai 1.2.4 EntropyBottleneck / GaussianConditional, MCM.__init__ topology), loaded through
     load_state_dict with the reference's parameter names and shapes;
 (b) against the committed golden intermediates (tests/golden/model_B64.pt, model_B144.pt)."""
import math

import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from oracle import ref_model
from textmae_image_compression_b200.config import PathConfig, vit_base
from textmae_image_compression_b200.synthetic import make_state_dict


# ---- independent module-style build -----------------------------------------------------------------
class Attention(nn.Module):
    def __init__(self, dim, heads):
        super().__init__()
        self.num_heads = heads
        self.scale = (dim // heads) ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.proj = nn.Linear(dim, dim)

    def forward(self, x):
        B, N, C = x.shape
        qkv = self.qkv(x).reshape(B, N, 3, self.num_heads, C // self.num_heads).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0], qkv[1], qkv[2]
        attn = ((q @ k.transpose(-2, -1)) * self.scale).softmax(dim=-1)
        return self.proj((attn @ v).transpose(1, 2).reshape(B, N, C))


class Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1, self.act, self.fc2 = nn.Linear(dim, hidden), nn.GELU(), nn.Linear(hidden, dim)

    def forward(self, x):
        return self.fc2(self.act(self.fc1(x)))


class Block(nn.Module):
    def __init__(self, dim, heads, ratio):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = Attention(dim, heads)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = Mlp(dim, int(dim * ratio))

    def forward(self, x):
        x = x + self.attn(self.norm1(x))
        return x + self.mlp(self.norm2(x))


class PatchEmbed(nn.Module):
    def __init__(self, p, cin, dim):
        super().__init__()
        self.proj = nn.Conv2d(cin, dim, kernel_size=p, stride=p)

    def forward(self, x):
        return self.proj(x).flatten(2).transpose(1, 2)


def conv3(cin, cout, stride=1):
    return nn.Conv2d(cin, cout, 3, stride, 1)


def subpel(cin, cout, r=2):
    return nn.Sequential(nn.Conv2d(cin, cout * r * r, 3, padding=1), nn.PixelShuffle(r))


def stack(chs, mk):
    layers = []
    for i in range(len(chs) - 1):
        layers.append(mk(i, chs[i], chs[i + 1]))
        if i < len(chs) - 2:
            layers.append(nn.GELU())
    return nn.Sequential(*layers)


class EB(nn.Module):
    def __init__(self, ch):
        super().__init__()
        f = (1, 3, 3, 3, 3, 1)
        for i in range(5):
            setattr(self, f"_matrix{i}", nn.Parameter(torch.zeros(ch, f[i + 1], f[i])))
            setattr(self, f"_bias{i}", nn.Parameter(torch.zeros(ch, f[i + 1], 1)))
            if i < 4:
                setattr(self, f"_factor{i}", nn.Parameter(torch.zeros(ch, f[i + 1], 1)))
        self.quantiles = nn.Parameter(torch.zeros(ch, 1, 3))

    def logits(self, x):
        for i in range(5):
            x = torch.matmul(F.softplus(getattr(self, f"_matrix{i}")), x) + getattr(self, f"_bias{i}")
            if i < 4:
                x = x + torch.tanh(getattr(self, f"_factor{i}")) * torch.tanh(x)
        return x

    def forward(self, z):
        med = self.quantiles[:, :, 1:2]
        C = z.shape[1]
        v = z.transpose(0, 1).reshape(C, 1, -1)
        out = torch.round(v - med) + med
        lo, up = self.logits(out - 0.5), self.logits(out + 0.5)
        sign = -torch.sign(lo + up)
        lik = torch.abs(torch.sigmoid(sign * up) - torch.sigmoid(sign * lo)).clamp_min(1e-9)
        shp = z.transpose(0, 1).shape
        return out.reshape(shp).transpose(0, 1), lik.reshape(shp).transpose(0, 1)


class ModuleMCM(nn.Module):
    def __init__(self, cfg: PathConfig):
        super().__init__()
        self.cfg = cfg
        C = cfg.encoder_embed_dim
        self.encoder_embed = PatchEmbed(cfg.patch_size, cfg.in_chans, C)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, C))
        self.encoder_pos_embed = nn.Parameter(torch.zeros(1, cfg.num_patches + 1, C), requires_grad=False)
        self.encoder_blocks = nn.ModuleList([Block(C, cfg.encoder_num_heads, cfg.mlp_ratio) for _ in range(cfg.encoder_depth)])
        self.encoder_norm = nn.LayerNorm(C, eps=1e-6)
        self.g_a = stack(cfg.g_a_channels(), lambda i, a, b: nn.Conv2d(a, b, 1))
        ha = cfg.h_a_layers()
        self.h_a = stack([ha[0][0]] + [l[1] for l in ha], lambda i, a, b: conv3(a, b, ha[i][2]))
        hs = cfg.h_s_layers()
        mk_hs = lambda i, a, b: subpel(a, b) if hs[i][2] == 2 else conv3(a, b)
        self.h_s_mean = stack([hs[0][0]] + [l[1] for l in hs], mk_hs)
        self.h_s_scale = stack([hs[0][0]] + [l[1] for l in hs], mk_hs)
        mk = lambda i, a, b: conv3(a, b)
        self.cc_transform_mean = nn.ModuleList([stack(cfg.cc_channels(i), mk) for i in range(cfg.num_slices)])
        self.cc_transform_scale = nn.ModuleList([stack(cfg.cc_channels(i), mk) for i in range(cfg.num_slices)])
        self.lrp_transform = nn.ModuleList([stack(cfg.lrp_channels(i), mk) for i in range(cfg.num_slices)])
        self.entropy_bottleneck = EB(cfg.hyperprior_depth)

    @torch.no_grad()
    def forward(self, imgs, ids_keep):
        cfg = self.cfg
        x = self.encoder_embed(imgs) + self.encoder_pos_embed[:, 1:, :]
        x = torch.gather(x, 1, ids_keep.unsqueeze(-1).repeat(1, 1, x.shape[-1]))
        cls = (self.cls_token + self.encoder_pos_embed[:, :1, :]).expand(x.shape[0], -1, -1)
        x = torch.cat((cls, x), 1)
        for b in self.encoder_blocks:
            x = b(x)
        x = self.encoder_norm(x)[:, 1:, :]
        s = cfg.side
        y = self.g_a(x.view(-1, s, s, cfg.encoder_embed_dim).permute(0, 3, 1, 2).contiguous())
        z = self.h_a(y)
        _, z_lik = self.entropy_bottleneck(z)
        med = self.entropy_bottleneck.quantiles[:, :, 1:2]
        z_hat = torch.round(z - med) + med
        ls, lm = self.h_s_scale(z_hat), self.h_s_mean(z_hat)
        hats, liks = [], []
        for i, ys in enumerate(y.chunk(cfg.num_slices, 1)):
            sup = hats[: cfg.max_support_slices]
            ms = torch.cat([lm] + sup, 1)
            mu = self.cc_transform_mean[i](ms)
            sg = self.cc_transform_scale[i](torch.cat([ls] + sup, 1))
            v = torch.round(ys - mu) + mu
            sc = sg.clamp_min(0.11)
            d = torch.abs(v - mu)
            lik = (0.5 * torch.erfc(-(2 ** -0.5) * ((0.5 - d) / sc)) - 0.5 * torch.erfc(-(2 ** -0.5) * ((-0.5 - d) / sc))).clamp_min(1e-9)
            yh = torch.round(ys - mu) + mu
            yh = yh + 0.5 * torch.tanh(self.lrp_transform[i](torch.cat([ms, yh], 1)))
            hats.append(yh); liks.append(lik)
        return x, y, z, torch.cat(liks, 1), z_lik, torch.cat(hats, 1)


SMALL = PathConfig(img_size=64, encoder_embed_dim=128, encoder_depth=2, encoder_num_heads=2, num_keep_patches=16)


def test_functional_oracle_equals_module_build():
    cfg = SMALL
    sd = make_state_dict(cfg, seed=3)
    m = ModuleMCM(cfg).eval()
    missing, unexpected = m.load_state_dict(sd, strict=True), None     # names AND shapes must match exactly
    g = torch.Generator().manual_seed(0)
    imgs = torch.rand(3, 3, 64, 64, generator=g)
    scores = torch.rand(3, cfg.num_patches, generator=g)
    out = ref_model.forward_rate(sd, cfg, imgs, scores)
    x, y, z, y_lik, z_lik, y_hat = m(imgs, out["ids_keep"])
    for name, a, b in (("x_remain", out["x_remain"], x), ("y", out["y"], y), ("z", out["z"], z),
                       ("y_lik", out["y_lik"], y_lik), ("z_lik", out["z_lik"], z_lik), ("y_hat", out["y_hat"], y_hat)):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-6), name


def test_state_dict_names_follow_reference_layout():
    sd = make_state_dict(vit_base(144), seed=0)
    # spot-check the names the reference's state_dict carries for the path
    for k, shape in (("cls_token", (1, 1, 768)), ("encoder_pos_embed", (1, 197, 768)),
                     ("encoder_embed.proj.weight", (768, 3, 16, 16)), ("encoder_blocks.11.attn.qkv.weight", (2304, 768)),
                     ("encoder_blocks.0.mlp.fc2.weight", (768, 3072)), ("g_a.0.weight", (704, 768, 1, 1)),
                     ("g_a.6.weight", (384, 512, 1, 1)), ("h_a.4.weight", (288, 336, 3, 3)), ("h_a.8.weight", (192, 240, 3, 3)),
                     ("h_s_mean.2.0.weight", (1152, 240, 3, 3)), ("h_s_scale.6.0.weight", (1536, 336, 3, 3)),
                     ("cc_transform_mean.0.0.weight", (224, 384, 3, 3)), ("cc_transform_mean.7.0.weight", (224, 576, 3, 3)),
                     ("cc_transform_scale.3.8.weight", (32, 80, 3, 3)), ("lrp_transform.0.0.weight", (224, 416, 3, 3)),
                     ("lrp_transform.11.0.weight", (224, 608, 3, 3)), ("entropy_bottleneck._matrix1", (192, 3, 3)),
                     ("entropy_bottleneck.quantiles", (192, 1, 3))):
        assert tuple(sd[k].shape) == shape, k


def test_pos_embed_first_row_zero_and_range():
    pe = make_state_dict(SMALL, 0)["encoder_pos_embed"]
    assert torch.all(pe[0, 0] == 0) and pe.abs().max() <= 1.0


@pytest.mark.parametrize("K", [64, 144])
def test_oracle_reproduces_committed_goldens(K, golden_dir, kodak):
    blob = torch.load(golden_dir / f"model_B{K}.pt")
    n = blob["n_img"]
    imgs, scores = kodak
    cfg = vit_base(K)
    sd = make_state_dict(cfg, seed=0)
    out = ref_model.forward_rate(sd, cfg, imgs[:n], scores[:n])
    assert torch.equal(out["ids_keep"], blob["ids_keep"])
    assert torch.equal(out["ids_restore"], blob["ids_restore"])
    # fp32 CPU GEMM summation order depends on the host's thread count / ISA: tolerances, not bit equality
    assert torch.allclose(out["y"], blob["y"], rtol=1e-3, atol=2e-3)
    assert torch.allclose(out["z"], blob["z"], rtol=1e-3, atol=2e-3)
    flips = (out["y_sym"] != blob["y_sym"]).float().mean().item()
    assert flips < 2e-3, flips
    assert torch.allclose(out["bpp"], blob["bpp"], rtol=2e-3)


def test_bpp_matches_rd_loss_formula():
    y_lik = torch.rand(2, 8, 4, 4) * 0.9 + 0.05
    z_lik = torch.rand(2, 4, 1, 1) * 0.9 + 0.05
    num_pixels = 2 * 64 * 64
    expect = sum(torch.log(l).sum() / (-math.log(2) * num_pixels) for l in (y_lik, z_lik))    # rd_loss.py:19-20
    assert torch.allclose(ref_model.bpp_batch(y_lik, z_lik, 64), expect)
    assert torch.allclose(ref_model.bpp_per_image(y_lik, z_lik, 64).mean(), expect, rtol=1e-5)


def test_invalid_geometry_raises_like_reference():
    with pytest.raises(ValueError):
        PathConfig(num_keep_patches=400).validate()          # K > L (MCM.py:374-376)
    with pytest.raises(RuntimeError):
        PathConfig(num_keep_patches=50).validate()           # not a square (view fails, MCM.py:729)
    with pytest.raises(RuntimeError):
        PathConfig(num_keep_patches=49).validate()           # sqrt not multiple of 4 (cat fails, MCM.py:761)
