"""CPU: host-side planning logic behind the C ABI - no device needed (the functions only compute geometry)."""
import ctypes as C

import pytest

from textmae_image_compression_b200 import _native


def _plan(T, H=12, N=4, mode=1):
    out = (C.c_int * 12)()
    assert _native.load().tmae_attention_plan(T, H, N, mode, out) == 0
    keys = ("ok", "Tp", "q_tiles", "tail_rows", "items", "nbuf", "nst", "tmem_cols", "smem", "q_rows", "kv_rows", "form")
    return dict(zip(keys, list(out)))


@pytest.mark.parametrize("mode", [1, 2, 3])
def test_attention_plan_invariants(mode):
    """Every supported T: the tiles fit shared memory and tensor memory, every query row is covered exactly once (tiles of 128
    rows + tail rows, or two 64-row halves in duo form), TMA boxes stay within the 256-row limit."""
    for T in list(range(1, 420)):
        for H in (1, 2, 12, 16):
            p = _plan(T, H, 3, mode)
            if not p["ok"]:
                assert p["Tp"] > 384 or p["smem"] == 0            # only long sequences are refused (mma.sync kernel serves them)
                continue
            assert p["Tp"] % 16 == 0 and p["Tp"] >= T
            assert p["smem"] <= 227 * 1024 and p["tmem_cols"] in (128, 256, 512)
            assert 1 <= p["q_rows"] <= 128 and 8 <= p["kv_rows"] <= 256
            assert p["nst"] >= 1 and p["nbuf"] in (1, 2)
            if p["nbuf"] == 2:
                assert p["nst"] >= 2                               # the look-ahead S MMA needs a second stage
            if p["form"] == 2:                                     # duo: T - 1 = 64 patch queries per head + the cls row on the tail warp
                assert T == 65 and H % 2 == 0 and p["items"] == 3 * H // 2 and p["tail_rows"] == 2
            else:
                assert p["q_tiles"] * 128 + p["tail_rows"] >= T and p["tail_rows"] <= 8
                assert (p["q_tiles"] - 1) * 128 < T
                assert p["items"] == 3 * H * p["q_tiles"]
                if p["tail_rows"]:
                    assert p["q_tiles"] * 128 + p["tail_rows"] == T
            if p["form"] == 1:
                assert p["Tp"] <= 96 and p["smem"] <= 113 * 1024 and p["tmem_cols"] <= 256      # two CTAs per SM


def test_attention_plan_named_geometries():
    assert _plan(65)["form"] == 0 and _plan(65, mode=2)["form"] == 1 and _plan(65, mode=3)["form"] == 2
    vl = _plan(257, H=16)                                          # ViT-L, K = 256: two items per head + one tail row
    assert (vl["q_tiles"], vl["tail_rows"], vl["nbuf"]) == (2, 1, 1)
    b144 = _plan(145)
    assert (b144["q_tiles"], b144["tail_rows"], b144["nbuf"]) == (2, 0, 2)
    assert not _plan(401, H=16)["ok"]                              # K = 400: served by the mma.sync kernel


def test_score_workspace_geometry():
    lib = _native.load()
    assert lib.tmae_scores_workspace_bytes(1, 4, 300, 224) == 0          # H < 8
    assert lib.tmae_scores_workspace_bytes(1, 64, 64, 100) == 0          # out_side not a multiple of 16
    one = lib.tmae_scores_workspace_bytes(1, 512, 768, 224)
    many = lib.tmae_scores_workspace_bytes(24, 512, 768, 224)
    assert 512 * 768 <= one < 4 * 512 * 768 and 20 * one < many <= 24 * one + 4096
