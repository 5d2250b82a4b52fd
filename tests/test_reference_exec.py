"""CPU: pins against REFERENCE CODE EXECUTED WHERE IT LIES (/root/reference; skipped on boxes without it):

 * the in-repo oracle (oracle/ref_model.py) against the reference's own `MCM` class - `__init__` topology, `forward_encoder`,
   `random_masking`, `get_ids_shuffle`, the whole `forward` incl. the slice loop - run verbatim from
   models/Compression/MCM.py with only the absent third-party leaf classes (timm / compressai / pytorch_msssim)
   supplied by oracle/ref_stubs.py (oracle/ref_exec.py);
 * `synthetic.sincos_pos_embed` against `common/pos_embed.py:get_2d_sincos_pos_embed`, bit for bit;
 * the per-image / batch rate (`bpp`, `rate_sums`) against `loss/rd_loss.py:RateDistortionLoss`;
 * the reconstruction half of this package (recon.py, stock PyTorch) against the reference forward's `x_hat` / `loss`;
 * the reference's `utils/engine.py:val_one_epoch` and `RateDistortionLoss` driven, unmodified, against this package's
   `MCM` module (host contract: dict keys, `aux_loss()`, `parameters()`, `eval()`); the native forward is replaced by
   a test double built from the oracle because this container has no GPU - the GPU twin of this test is
   tests/test_gpu_recon.py against the goldens frozen here.
"""
import importlib.util
import math
import sys
import types

import pytest
import torch

from oracle import ref_exec, ref_model, ref_pins
from textmae_image_compression_b200 import MCM, PathConfig, make_state_dict, recon
from textmae_image_compression_b200.config import vit_base
from textmae_image_compression_b200.synthetic import sincos_pos_embed

pytestmark = pytest.mark.skipif(not ref_exec.reference_available(), reason="/root/reference is not mounted")

SMALL = dict(img_size=64, encoder_embed_dim=128, encoder_depth=2, encoder_num_heads=2, num_keep_patches=16,
             decoder_embed_dim=64, decoder_depth=2, decoder_num_heads=2)


def _inputs(cfg, n, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(n, 3, cfg.img_size, cfg.img_size, generator=g), torch.rand(n, cfg.num_patches, generator=g)


# ---------------------------------------------------------------------------------------------- oracle vs reference MCM
@pytest.mark.parametrize("kw,n", [(SMALL, 3), (dict(img_size=128, encoder_embed_dim=128, encoder_depth=2, encoder_num_heads=2,
                                                    num_keep_patches=64), 2)])
def test_oracle_equals_reference_mcm_forward(kw, n):
    cfg = PathConfig(**kw)
    sd = make_state_dict(cfg, seed=3, include_decoder=True)
    model, missing, unexpected = ref_exec.build_reference_model(cfg, sd)
    assert unexpected == [], unexpected            # every synthetic tensor name exists in the reference module ...
    assert missing == [], missing                  # ... and every reference parameter is provided (names AND shapes)
    imgs, scores = _inputs(cfg, n, 0)
    out = ref_exec.reference_forward(model, imgs, scores)
    ref = ref_model.forward_rate(sd, cfg, imgs, scores)
    assert torch.equal(out["likelihoods"]["y"], ref["y_lik"])      # same torch ops in the same order: bit-equal
    assert torch.equal(out["likelihoods"]["z"], ref["z_lik"])
    x_remain, ids_restore = model.forward_encoder(imgs, scores)
    assert torch.equal(x_remain, ref["x_remain"]) and torch.equal(ids_restore, ref["ids_restore"])


def test_oracle_equals_reference_mcm_forward_vit_base():
    cfg = vit_base(64)
    sd = make_state_dict(cfg, seed=0)
    model, missing, unexpected = ref_exec.build_reference_model(cfg, sd)
    assert unexpected == [] and all(k.startswith(recon.DECODER_PREFIXES) for k in missing), (unexpected, missing[:5])
    imgs, scores = _inputs(cfg, 1, 5)
    out = ref_exec.reference_forward(model, imgs, scores)
    ref = ref_model.forward_rate(sd, cfg, imgs, scores)
    assert torch.equal(out["likelihoods"]["y"], ref["y_lik"]) and torch.equal(out["likelihoods"]["z"], ref["z_lik"])


def test_reference_rejects_what_the_config_rejects():
    """K with sqrt(K) % 4 != 0 fails inside the reference forward (torch.cat size mismatch, MCM.py:761)."""
    cfg = PathConfig(img_size=128, encoder_embed_dim=128, encoder_depth=1, encoder_num_heads=2, num_keep_patches=64)
    bad = PathConfig(img_size=128, encoder_embed_dim=128, encoder_depth=1, encoder_num_heads=2, num_keep_patches=36)
    with pytest.raises(RuntimeError):
        bad.validate()
    MCMref = ref_exec.load_reference_mcm_class()
    with ref_exec._np_float_shim():
        m = MCMref(img_size=128, encoder_embed_dim=128, encoder_depth=1, encoder_num_heads=2, num_keep_patches=36,
                   decoder_embed_dim=64, decoder_depth=1, decoder_num_heads=2).eval()
    imgs, scores = _inputs(cfg, 1, 2)
    with pytest.raises(RuntimeError):
        with torch.no_grad():
            m(imgs, scores)


# ---------------------------------------------------------------------------------------------- pos-embed / rd_loss pins
@pytest.mark.parametrize("dim,grid", [(768, 14), (1024, 32), (128, 4), (512, 14)])
def test_sincos_pos_embed_is_bit_equal_to_reference(dim, grid):
    assert torch.equal(sincos_pos_embed(dim, grid), ref_pins.reference_pos_embed_parameter(dim, grid))


def test_bpp_equals_reference_rate_distortion_loss():
    cfg = PathConfig(**SMALL)
    sd = make_state_dict(cfg, seed=3)
    imgs, scores = _inputs(cfg, 4, 1)
    ref = ref_model.forward_rate(sd, cfg, imgs, scores)
    bpp_loss = ref_pins.reference_bpp_loss(ref["y_lik"], ref["z_lik"], cfg.img_size)
    assert torch.allclose(ref["bpp_batch"], bpp_loss, rtol=0, atol=0)          # oracle restatement == reference class
    assert torch.allclose(ref["bpp"].mean(), bpp_loss, rtol=2e-6)               # per-image form averages to it
    # the library's rate_sums pair {sum log2 lik, pixels} is the same quantity: bpp_loss == -sums[0] / sums[1]
    sum_log2 = (torch.log2(ref["y_lik"].double()).sum() + torch.log2(ref["z_lik"].double()).sum()).item()
    assert abs(-sum_log2 / (4 * cfg.img_size ** 2) - bpp_loss.item()) < 2e-6 * bpp_loss.item()


# ---------------------------------------------------------------------------------------------- reconstruction half
def test_recon_matches_reference_forward():
    cfg = PathConfig(**SMALL)
    sd = make_state_dict(cfg, seed=3, include_decoder=True)
    model, _, _ = ref_exec.build_reference_model(cfg, sd)
    imgs, scores = _inputs(cfg, 3, 0)
    out = ref_exec.reference_forward(model, imgs, scores)
    ref = ref_model.forward_rate(sd, cfg, imgs, scores)
    tokens = ref["y_hat"].permute(0, 2, 3, 1).reshape(3, cfg.num_keep_patches, cfg.latent_depth)
    loss, x_hat = recon.reconstruct(sd, cfg, tokens, ref["ids_restore"], imgs)
    assert torch.allclose(x_hat, out["x_hat"], rtol=1e-5, atol=1e-6)
    assert torch.allclose(loss[0], out["loss"][0], rtol=1e-5, atol=1e-6)        # 1 - SSIM
    assert torch.allclose(loss[1], out["loss"][1], rtol=1e-6)                   # L1
    assert float(loss[2]) == 0.0


# ---------------------------------------------------------------------------------------------- the reference's callers
class _OracleBackedMCM(MCM):
    """Test double for a GPU-less container: the C-ABI call is replaced by the fp32 oracle; everything around it -
    argument checks, result dict, reconstruction half, aux_loss, parameters - is the product code."""

    def _ensure_handle(self):
        self._handle_device = torch.device("cpu")

    def _native_forward(self, imgs, total_scores, indexes=False):
        ref = ref_model.forward_rate(self._weights, self.cfg, imgs, total_scores)
        nhwc = lambda t: t.permute(0, 2, 3, 1).contiguous()
        return {"y_likelihoods": nhwc(ref["y_lik"]), "z_likelihoods": nhwc(ref["z_lik"]), "y_symbols": nhwc(ref["y_sym"]),
                "z_symbols": nhwc(ref["z_sym"]), "y_hat": nhwc(ref["y_hat"]), "z_hat": nhwc(ref["z_hat"]), "bpp": ref["bpp"],
                "rate_sums": torch.zeros(2, dtype=torch.float64), "ids_restore": ref["ids_restore"], "ids_keep": ref["ids_keep"],
                "ids_shuffle": ref["ids_restore"]}


def _load_reference_engine():
    """utils/engine.py imports `models.Compression.common.{distributed, logger}` (plain torch): import as written."""
    sys.path.insert(0, str(ref_exec.REFERENCE_ROOT))
    try:
        for k in [k for k in sys.modules if k == "models" or k.startswith("models.") or k == "utils" or k.startswith("utils.")]:
            del sys.modules[k]
        spec = importlib.util.spec_from_file_location("_ref_engine", str(ref_pins.ENGINE))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod
    finally:
        sys.path.remove(str(ref_exec.REFERENCE_ROOT))


def test_reference_val_one_epoch_and_rd_loss_run_against_this_module(capsys):
    cfg = PathConfig(**SMALL)
    sd = make_state_dict(cfg, seed=3, include_decoder=True)
    ours = _OracleBackedMCM(**SMALL)
    ours.load_state_dict(sd)
    ours.eval()
    ref_model_, _, _ = ref_exec.build_reference_model(cfg, sd)
    criterion = ref_pins.load_rate_distortion_loss()(lmbda=1e-2)
    engine = _load_reference_engine()
    batches = [(*_inputs(cfg, 2, s)[:1], torch.tensor([[64, 64]] * 2), _inputs(cfg, 2, s)[1]) for s in (1, 2)]   # (img, size, scores)
    stats_ours = engine.val_one_epoch(0, batches, ours, criterion)
    stats_ref = engine.val_one_epoch(0, batches, ref_model_, criterion)
    capsys.readouterr()
    assert stats_ours.keys() == stats_ref.keys() and "bpp_loss" in stats_ours
    for k in stats_ref:
        assert abs(stats_ours[k] - stats_ref[k]) <= 0.011 * max(1.0, abs(stats_ref[k])), (k, stats_ours[k], stats_ref[k])
    # and the criterion directly on one forward (rd_loss.py:14-28)
    imgs, scores = _inputs(cfg, 2, 7)
    a = criterion(ours(imgs, scores), imgs)
    b = criterion(ref_model_(imgs, scores), imgs)
    for k in ("bpp_loss", "ssim_loss", "L1_loss", "vgg_loss", "loss"):
        assert torch.allclose(a[k].float(), b[k].float(), rtol=1e-4, atol=1e-6), k
