"""GPU: (1) the CUDA path and the reconstruction half against outputs of the REFERENCE's own MCM.forward, frozen in
tests/golden/refexec_small.pt by tests/golden/make_refexec_golden.py (the reference file executed verbatim in the build
container; /root/reference does not exist on this box); (2) the call pattern of utils/engine.py:189-199 val_one_epoch and
loss/rd_loss.py:14-28 RateDistortionLoss restated against those goldens; (3) symbol / index emission for the range coder
(MCM.compress front half, MCM.py:805-873) against the oracle + a restatement of compressai's build_indexes."""
import math

import pytest
import torch

from oracle import ref_model
from textmae_image_compression_b200 import MCM, PathConfig, make_state_dict

pytestmark = pytest.mark.gpu


def _model(kw, seed, **flags):
    cfg = PathConfig(**kw)
    sd = make_state_dict(cfg, seed=seed, include_decoder=True)
    m = MCM(**kw, softmax_isa=16, **flags)
    m.load_state_dict(sd)
    m.cuda().eval()
    return cfg, sd, m


def rd_loss_restated(out, target, lmbda=1e-2):
    """loss/rd_loss.py:14-28 (the class itself is executed against this module in tests/test_reference_exec.py)."""
    N, _, H, W = target.size()
    num_pixels = N * H * W
    r = {"bpp_loss": sum(torch.log(l).sum() / (-math.log(2) * num_pixels) for l in out["likelihoods"].values()),
         "ssim_loss": out["loss"][0], "L1_loss": out["loss"][1], "vgg_loss": out["loss"][2]}
    r["loss"] = lmbda * (0.25 * r["ssim_loss"] + 10 * r["L1_loss"] + 0.1 * r["vgg_loss"]) + r["bpp_loss"]
    return r


@pytest.mark.parametrize("key", ["small", "mid"])
@pytest.mark.parametrize("precise", ["all-x6", None])
def test_forward_against_reference_mcm_goldens(cuda_dev, golden_dir, key, precise):
    blob = torch.load(golden_dir / "refexec_small.pt")[key]
    cfg, sd, m = _model(blob["kwargs"], blob["seed"], precise=precise)
    imgs, scores = blob["imgs"].cuda(), blob["scores"].cuda()
    out = m(imgs, scores)                                    # decoder weights are loaded -> reference-shaped dict
    torch.cuda.synchronize()
    assert set(("loss", "likelihoods", "x_hat")) <= set(out)                       # MCM.py:799-803
    assert torch.equal(out["ids_restore"].cpu(), blob["ids_restore"])
    assert out["likelihoods"]["y"].shape == blob["y_lik"].shape and out["x_hat"].shape == blob["x_hat"].shape
    rd = rd_loss_restated(out, imgs)
    want = blob["rd_loss"]
    if precise is not None:
        ylik = out["likelihoods"]["y"].cpu()
        rel = ((ylik - blob["y_lik"]).abs() / blob["y_lik"])
        assert rel.median().item() < 5e-3
        assert torch.allclose(out["likelihoods"]["z"].cpu(), blob["z_lik"], rtol=5e-3)
        assert abs(rd["bpp_loss"].item() - want["bpp_loss"]) < 5e-3 * want["bpp_loss"]
        assert abs(rd["loss"].item() - want["loss"]) < 5e-3 * want["loss"]
        # images whose symbols all agree with the reference's (no rounding-boundary flip -> no cascade, see
        # tests/test_gpu_forward.py::_compare_precise): likelihoods and the reconstruction match element-wise
        clean = [n for n in range(rel.shape[0]) if (rel[n] <= 5e-3).all()]
        print(f"{key}: {len(clean)} of {rel.shape[0]} images without any symbol flip")
        for n in clean:
            assert (out["x_hat"][n].cpu() - blob["x_hat"][n]).abs().max().item() < 1e-3
        if len(clean) == rel.shape[0]:
            assert abs(rd["L1_loss"].item() - want["L1_loss"]) < 1e-4 and abs(rd["ssim_loss"].item() - want["ssim_loss"]) < 1e-4
        else:
            assert abs(rd["L1_loss"].item() - want["L1_loss"]) < 2e-2 and abs(rd["ssim_loss"].item() - want["ssim_loss"]) < 2e-2
    else:
        assert abs(rd["bpp_loss"].item() - want["bpp_loss"]) < 2e-2 * want["bpp_loss"]
        assert abs(rd["L1_loss"].item() - want["L1_loss"]) < 2e-2 and abs(rd["ssim_loss"].item() - want["ssim_loss"]) < 2e-2
    assert abs(m.aux_loss().item() - blob["aux_loss"]) < 1e-3 * blob["aux_loss"]   # engine.py:194
    x_remain, ids_restore = m.forward_encoder(imgs, scores)                          # MCM.forward_encoder contract
    tol = 1e-4 if precise is not None else 2e-2
    assert ((x_remain.cpu() - blob["x_remain"]).norm() / blob["x_remain"].norm()).item() < tol


def test_need_recon_switch(cuda_dev, golden_dir):
    blob = torch.load(golden_dir / "refexec_small.pt")["small"]
    cfg, sd, m = _model(blob["kwargs"], blob["seed"])
    imgs, scores = blob["imgs"].cuda(), blob["scores"].cuda()
    assert "loss" not in m(imgs, scores, need_recon=False)
    assert "x_hat" in m(imgs, scores, need_recon=True)
    m2 = MCM(**blob["kwargs"], softmax_isa=16)
    m2.load_state_dict(make_state_dict(cfg, seed=blob["seed"]))                       # hot-path tensors only
    m2.cuda().eval()
    assert "loss" not in m2(imgs, scores)
    with pytest.raises(RuntimeError):
        m2(imgs, scores, need_recon=True)


def build_indexes_restated(scales, table):
    """compressai 1.2.4 GaussianConditional.build_indexes: scales lower-bounded (0.11), then
    indexes = len(table) - 1 - sum(scales <= s for s in table[:-1])."""
    scales = torch.clamp_min(scales, 0.11)
    idx = scales.new_full(scales.size(), len(table) - 1).int()
    for s in table[:-1]:
        idx -= (scales <= s).int()
    return idx


@pytest.mark.parametrize("kw,n", [(dict(img_size=64, encoder_embed_dim=128, encoder_depth=2, encoder_num_heads=2, num_keep_patches=16), 3),
                                  (dict(img_size=128, encoder_embed_dim=128, encoder_depth=2, encoder_num_heads=2, num_keep_patches=64), 2)])
def test_compress_symbols_match_oracle(cuda_dev, kw, n):
    cfg = PathConfig(**kw)
    sd = make_state_dict(cfg, seed=7)
    m = MCM(**kw, softmax_isa=16, precise="all-x6", extra_outputs=True)
    m.load_state_dict(sd)
    m.cuda().eval()
    g = torch.Generator().manual_seed(13)
    imgs, scores = torch.rand(n, 3, cfg.img_size, cfg.img_size, generator=g), torch.rand(n, cfg.num_patches, generator=g)
    m.update(force=True)                                                             # testing.py:223
    res = m.compress_symbols(imgs.cuda(), scores.cuda())
    fwd = m(imgs.cuda(), scores.cuda())
    torch.cuda.synchronize()
    table = MCM.get_scale_table()
    assert table.numel() == 64 and abs(table[0].item() - 0.11) < 1e-6 and abs(table[-1].item() - 256.0) < 1e-3
    # packing order: per image, slice by slice, each flattened (c, y, x)  ==  NCHW flatten (MCM.py:872-873)
    assert torch.equal(res["y_symbols"].cpu(), fwd["latents"]["y_sym"].contiguous().reshape(n, -1).cpu())
    # indexes of OUR sigma through the restated build_indexes: bit-exact
    want_idx = build_indexes_restated(fwd["sigma"].cpu().contiguous(), table)
    assert torch.equal(res["y_indexes"].cpu(), want_idx.reshape(n, -1))
    # against the fp32 oracle: slice 0 has no upstream symbols (no cascade), so apart from rounding-boundary / bucket-edge
    # cases its symbols and indexes are the oracle's
    ref = ref_model.forward_rate(sd, cfg, imgs, scores)
    assert torch.equal(res["z_symbols"].cpu(), fwd["latents"]["z_sym"].contiguous().cpu())
    n0 = cfg.slice_ch * cfg.side * cfg.side
    clean = (res["z_symbols"].cpu() == ref["z_sym"]).flatten(1).all(1)
    oracle_idx = build_indexes_restated(ref["sigma"], table).reshape(n, -1)
    assert (res["y_indexes"].cpu()[clean, :n0] != oracle_idx[clean, :n0]).float().mean().item() < 2e-3
    assert (res["y_symbols"].cpu()[clean, :n0] != ref["y_sym"].reshape(n, -1)[clean, :n0]).float().mean().item() < 2e-3
    assert torch.equal(res["z_indexes"][0, :, 0, 0].cpu(), torch.arange(cfg.hyperprior_depth, dtype=torch.int32))
    assert res["shape"] == (cfg.side // 4, cfg.side // 4) and torch.equal(res["ids_restore"].cpu(), ref["ids_restore"])
