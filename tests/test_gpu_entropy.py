"""GPU, teacher-forced: the quantiser / likelihood kernels get the ORACLE's inputs for their stage, so symbols must be
bit-exact and likelihoods equal to fp32 round-off."""
import ctypes as C

import pytest
import torch

from oracle import ref_model
from tests import gpu_util as G
from textmae_image_compression_b200 import MCM, PathConfig, _native, make_state_dict

pytestmark = pytest.mark.gpu

SMALL = dict(img_size=64, encoder_embed_dim=128, encoder_depth=2, encoder_num_heads=2, num_keep_patches=16)


def test_gaussian_rate_operator(cuda_dev):
    g = torch.Generator().manual_seed(0)
    n = 200_000
    y = torch.randn(n, generator=g) * 3
    mu = torch.randn(n, generator=g)
    sigma = torch.rand(n, generator=g) * 2.5 - 0.3           # some below the 0.11 bound, some negative
    sigma[:1000] = torch.rand(1000, generator=g) * 30
    y[-500:] = mu[-500:] + torch.randint(-3, 4, (500,), generator=g).float() + 0.5   # exact rounding ties
    out_ref, lik_ref = ref_model.gaussian_conditional_eval(y, sigma, mu)
    sym_ref = torch.round(y - mu).to(torch.int32)
    lik = torch.empty(n, device=cuda_dev); sym = torch.empty(n, dtype=torch.int32, device=cuda_dev); yh = torch.empty(n, device=cuda_dev)
    yd, md, sd_ = y.to(cuda_dev), mu.to(cuda_dev), sigma.to(cuda_dev)        # keep the device copies alive
    rc = _native.load().tmae_gaussian_rate(G.ptr(yd), G.ptr(md), G.ptr(sd_), n,
                                           G.ptr(lik), G.ptr(sym), G.ptr(yh), G.stream())
    _native.check(rc)
    torch.cuda.synchronize()
    assert torch.equal(sym.cpu(), sym_ref)                                    # bit-exact symbols (half-to-even ties too)
    assert torch.equal(yh.cpu(), out_ref)
    rel = ((lik.cpu() - lik_ref).abs() / lik_ref).max().item()
    assert rel < 2e-4, rel
    assert (lik_ref <= 1e-9).any() and (lik.cpu()[lik_ref <= 1e-9] == lik_ref[lik_ref <= 1e-9]).all()   # floor clamp


def test_bottleneck_rate_operator(cuda_dev):
    cfg = PathConfig(**SMALL)
    sd = make_state_dict(cfg, seed=1)
    m = MCM(**SMALL)
    m.load_state_dict(sd)
    m.cuda()
    m._ensure_handle()
    g = torch.Generator().manual_seed(3)
    z = torch.randn(6, cfg.hyperprior_depth, 5, 7, generator=g) * 4
    zh_ref, lik_ref = ref_model.entropy_bottleneck_eval(sd, z, torch.float32)
    med = sd["entropy_bottleneck.quantiles"][:, :, 1:2]
    sym_ref = torch.round(z - med).to(torch.int32)
    zc = z.permute(0, 2, 3, 1).contiguous().to(cuda_dev)
    rows = zc.numel() // cfg.hyperprior_depth
    lik = torch.empty_like(zc); zh = torch.empty_like(zc); sym = torch.empty(zc.shape, dtype=torch.int32, device=cuda_dev)
    rc = _native.load().tmae_bottleneck_rate(m._handle, G.ptr(zc), rows, G.ptr(lik), G.ptr(sym), G.ptr(zh), G.stream())
    _native.check(rc, m._handle, RuntimeError)
    torch.cuda.synchronize()
    assert torch.equal(sym.cpu().permute(0, 3, 1, 2), sym_ref)
    assert torch.equal(zh.cpu().permute(0, 3, 1, 2), zh_ref)
    rel = ((lik.cpu().permute(0, 3, 1, 2) - lik_ref).abs() / lik_ref).max().item()
    assert rel < 5e-4, rel
