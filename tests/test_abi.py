"""CPU: the C-ABI library builds, loads, and exports every symbol include/tmae.h declares; argument validation
and the host module's error classes match the reference's.  No compute calls (no GPU here)."""
import ctypes as C
import re
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib():
    from textmae_image_compression_b200 import build, _native
    build.build()
    return _native.load()


def test_header_symbols_are_exported(lib):
    from textmae_image_compression_b200 import _native
    header = (ROOT / "include" / "tmae.h").read_text()
    declared = set(re.findall(r"TMAE_API\s+[\w\s\*]+?\b(tmae_\w+)\s*\(", header))
    assert len(declared) >= 18
    assert declared == set(_native.SIGNATURES), declared ^ set(_native.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.tmae_abi_version() == 3


def test_struct_layouts_match_header():
    from textmae_image_compression_b200 import _native
    assert C.sizeof(_native.TmaeConfig) == 15 * 4
    assert C.sizeof(_native.TmaeOutputs) == 19 * 8
    assert C.sizeof(_native.TmaeHostOutputs) == 7 * 8
    assert C.sizeof(_native.TmaeProfileEntry) == 32 + 4 + 4 + 8 + 8 + 8
    header = (ROOT / "include" / "tmae.h").read_text()
    body = header[header.index("typedef struct {\n    float*   y_likelihoods"):header.index("} tmae_outputs;")]
    fields = re.findall(r"\*\s+(\w+);", body)
    assert tuple(fields) == _native.OUTPUT_FIELDS


@pytest.mark.parametrize("K,msg", [(400, "greater than the length"), (50, "perfect square"), (49, "multiple of 4")])
def test_create_rejects_reference_invalid_geometry(lib, K, msg):
    from textmae_image_compression_b200 import _native
    cfg = _native.TmaeConfig(224, 16, 3, 768, 12, 12, 512, 4.0, 384, 192, 12, K, 1e-6, 16, 0)
    hp = C.c_void_p()
    rc = lib.tmae_create(C.byref(cfg), C.byref(hp))
    assert rc == _native.TMAE_EINVAL
    assert msg in _native.last_error()


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only check")
def test_no_cpu_fallback(lib):
    from textmae_image_compression_b200 import _native, MCM, make_state_dict, PathConfig
    cfg = _native.TmaeConfig(224, 16, 3, 768, 12, 12, 512, 4.0, 384, 192, 12, 64, 1e-6, 16, 0)
    hp = C.c_void_p()
    assert lib.tmae_create(C.byref(cfg), C.byref(hp)) == _native.TMAE_ECUDA
    assert "no CPU fallback" in _native.last_error()
    small = dict(img_size=64, encoder_embed_dim=128, encoder_depth=1, encoder_num_heads=2, num_keep_patches=16)
    m = MCM(**small)
    m.load_state_dict(make_state_dict(PathConfig(**small), 0))
    with pytest.raises(RuntimeError, match="CUDA only"):
        m(torch.rand(1, 3, 64, 64), torch.rand(1, 16))


def test_module_mirrors_reference_error_classes():
    from textmae_image_compression_b200 import MCM
    with pytest.raises(ValueError):
        MCM(num_keep_patches=400)
    with pytest.raises(RuntimeError):
        MCM(num_keep_patches=49)
    m = MCM(num_keep_patches=64)
    with pytest.raises(NotImplementedError):
        m.train()
    assert m.eval() is m
    with pytest.raises(NotImplementedError):
        m.compress(None, None)


def test_product_path_never_imports_oracle():
    """The shipped package must not reach into oracle/ (oracle is test infrastructure)."""
    pkg = ROOT / "textmae_image_compression_b200"
    for f in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")) + list(pkg.rglob("*.h")):
        txt = f.read_text()
        assert "import oracle" not in txt and "from oracle" not in txt and "oracle/" not in txt.replace("oracle/mask_oracle.c for the pinned", ""), f


def test_conv_tiling_geometry_invariants():
    """Host logic of the conv engine (no GPU): for every grid side and batch size the tiles cover every pixel exactly once,
    fit 128 accumulator rows, and the haloed-box reuse is only claimed when a dy shift is a whole number of swizzle atoms."""
    import ctypes as C
    from textmae_image_compression_b200 import _native
    lib = _native.load()
    out = (C.c_int * 6)()
    assert lib.tmae_conv_geometry(0, 4, out) != 0 and lib.tmae_conv_geometry(129, 1, out) != 0
    for s in list(range(1, 41)) + [48, 64, 96, 128]:
        for n in (1, 2, 3, 5, 7, 16, 64, 257):
            assert lib.tmae_conv_geometry(s, n, out) == 0, (s, n)
            box_y, box_n, y_tiles, m_tiles, rows_used, reuse = list(out)
            assert 1 <= box_y <= s and box_n >= 1 and rows_used == s * box_y * box_n <= 128, (s, n, list(out))
            assert y_tiles == -(-s // box_y) and m_tiles == y_tiles * -(-n // box_n), (s, n, list(out))
            covered = set()
            for t in range(m_tiles):
                n0, y0 = (t // y_tiles) * box_n, (t % y_tiles) * box_y
                for nl in range(box_n):
                    for yl in range(box_y):
                        if n0 + nl < n and y0 + yl < s:
                            assert (n0 + nl, y0 + yl) not in covered
                            covered.add((n0 + nl, y0 + yl))
            assert len(covered) == n * s, (s, n, list(out))
            if reuse:
                assert (box_n * s) % 8 == 0 and box_y >= 2, (s, n, list(out))
    # the shapes of the shipped configurations keep all 128 accumulator rows busy
    for s in (2, 4, 8, 16):
        assert lib.tmae_conv_geometry(s, 64, out) == 0 and out[4] == 128 and out[5] == 1, (s, list(out))
