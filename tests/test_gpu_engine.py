"""GPU: the tcgen05 GEMM / conv engine against torch fp32 on the same bf16-rounded operands, and against the
CUDA-core checker kernel that shares its parameter block and epilogue."""
import pytest
import torch
import torch.nn.functional as F

from tests import gpu_util as G

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M,N,K,bn", [(128, 64, 64, 64), (128, 128, 256, 128), (300, 224, 384, 0), (1000, 2304, 768, 256),
                                       (4160, 768, 3072, 0), (257, 80, 176, 0), (130, 32, 80, 32), (512, 704, 768, 192)])
@pytest.mark.parametrize("impl", [1, 0], ids=["checker", "tcgen05"])
def test_gemm_matches_torch(cuda_dev, M, N, K, bn, impl):
    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    A = (torch.randn(M, K, generator=g) * 0.5).to(cuda_dev).bfloat16()
    B = (torch.randn(N, K, generator=g) * 0.1).to(cuda_dev).bfloat16()
    bias = torch.randn(N, generator=g).to(cuda_dev)
    ref = A.float() @ B.float().t() + bias
    out = G.gemm(A, B, bias, block_n=bn, impl=impl)
    err = G.rel_err(out, ref)
    assert err < 2e-5, f"impl={impl} rel err {err}"          # fp32 accumulate of identical bf16 operands


@pytest.mark.parametrize("N,s,Cin,Cout,gelu", [(2, 8, 64, 32, False), (3, 12, 384, 224, True), (2, 4, 192, 240, True),
                                                  (5, 8, 80, 32, False), (1, 16, 576, 224, True),
                                                  # tile geometries of the compact 4-D TMA conv: ragged image groups (box_n
                                                  # tails), row tiles of one image, tiny and odd grids, the B64 slice shape
                                                  (7, 12, 64, 48, False), (9, 6, 128, 64, True), (20, 3, 64, 32, False),
                                                  (40, 2, 192, 96, False), (1, 8, 64, 32, False), (2, 32, 64, 32, False),
                                                  (64, 8, 352, 224, True),
                                                  # haloed-box A reuse with partial tiles: 96 of 128 accumulator rows (s = 6, 2 rows
                                                  # x 8 images), ragged image groups at s = 4 / 2, one image of 16 x 16
                                                  (40, 6, 128, 64, True), (21, 4, 64, 48, False), (70, 2, 64, 32, False), (3, 16, 64, 96, True)])
@pytest.mark.parametrize("impl", [1, 0], ids=["checker", "tcgen05"])
def test_conv3x3_matches_torch(cuda_dev, N, s, Cin, Cout, gelu, impl):
    g = torch.Generator(device="cpu").manual_seed(N * 1000 + s * 100 + Cin)
    x = torch.randn(N, s, s, Cin, generator=g).to(cuda_dev).bfloat16()
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) * (1.0 / (3 * Cin ** 0.5))).to(cuda_dev)
    b = torch.randn(Cout, generator=g).to(cuda_dev)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.bfloat16().float(), b, padding=1)
    if gelu:
        ref = F.gelu(ref)
    ref = ref.permute(0, 2, 3, 1)
    out = G.conv3x3(x, w, b, gelu=gelu, impl=impl)
    err = G.rel_err(out, ref)
    assert err < 3e-5, f"impl={impl} rel err {err}"


# ---- precise configuration (split-bf16 planes, three tensor-core terms per product): fp32 operands, compared with an
# fp64 product of the SAME fp32 operands.  Expected error ~2^-17 per product (the dropped lo*lo term and the rounding of
# the lo planes), i.e. 100x below the 2^-9 of plain bf16 operands.
@pytest.mark.parametrize("M,N,K,bn", [(128, 64, 64, 64), (300, 224, 384, 0), (1000, 2304, 768, 256), (4160, 768, 3072, 0),
                                       (257, 80, 176, 0), (130, 32, 80, 32)])
@pytest.mark.parametrize("impl", [1, 0], ids=["checker", "tcgen05"])
@pytest.mark.parametrize("planes", [2, 3])
def test_gemm_split_matches_fp64(cuda_dev, M, N, K, bn, impl, planes):
    g = torch.Generator(device="cpu").manual_seed(M + N + K + 7)
    A = (torch.randn(M, K, generator=g) * 0.5).to(cuda_dev)
    B = (torch.randn(N, K, generator=g) * 0.1).to(cuda_dev)
    bias = torch.randn(N, generator=g).to(cuda_dev)
    ref = (A.double() @ B.double().t() + bias.double())
    out = G.gemm_split(A, B, bias, block_n=bn, impl=impl, planes=planes)
    err = G.rel_err(out, ref)
    bf16_err = G.rel_err(A.bfloat16().double() @ B.bfloat16().double().t() + bias.double(), ref)
    fp32_err = G.rel_err(A @ B.t() + bias, ref)                    # what plain fp32 arithmetic gives on the same operands
    if planes == 2:
        assert err < 2e-5 and err < bf16_err / 50, f"impl={impl} rel err {err} (plain bf16 operands: {bf16_err})"
    else:
        # three planes carry the fp32 operands exactly and every product is exact; what remains is the fp32 accumulation -
        # sequential FMA in the checker (~1e-6), truncating adds in the tensor core (measured 2-3e-6 at K = 3072, growing
        # linearly with K) - a few times torch's blocked fp32 GEMM, 5-10x below the two-plane form
        assert err < 6e-6 and err < 8 * fp32_err + 1e-6, f"impl={impl} rel err {err} (torch fp32: {fp32_err})"


@pytest.mark.parametrize("N,s,Cin,Cout,gelu", [(2, 8, 64, 32, False), (3, 12, 384, 224, True), (5, 8, 80, 32, False),
                                                  (1, 16, 576, 224, True), (7, 12, 64, 48, False), (40, 2, 192, 96, False),
                                                  (64, 8, 352, 224, True), (40, 6, 128, 64, True), (21, 4, 64, 48, False)])
@pytest.mark.parametrize("impl", [1, 0], ids=["checker", "tcgen05"])
@pytest.mark.parametrize("planes", [2, 3])
def test_conv3x3_split_matches_fp64(cuda_dev, N, s, Cin, Cout, gelu, impl, planes):
    g = torch.Generator(device="cpu").manual_seed(N * 1000 + s * 100 + Cin + 3)
    x = torch.randn(N, s, s, Cin, generator=g).to(cuda_dev)
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) * (1.0 / (3 * Cin ** 0.5))).to(cuda_dev)
    b = torch.randn(Cout, generator=g).to(cuda_dev)
    ref = F.conv2d(x.double().permute(0, 3, 1, 2), w.double(), b.double(), padding=1)
    if gelu:
        ref = F.gelu(ref)
    ref = ref.permute(0, 2, 3, 1)
    out = G.conv3x3_split(x, w, b, gelu=gelu, impl=impl, planes=planes)
    err = G.rel_err(out, ref)
    tol = 2e-5 if planes == 2 else 2e-6 + 5e-9 * 9 * Cin          # accumulation error of the fp32 tensor-core adder grows with K
    assert err < tol, f"impl={impl} planes={planes} rel err {err} (tol {tol})"


# ---- CTA-pair kernel (tcgen05.mma.cta_group::2): 256-row tiles over two SMs, each CTA stages half of the weight tile.
# Same k order as the one-CTA kernel -> bit-identical results; ragged M (last pair half empty / partly filled), several N tiles.
@pytest.mark.parametrize("M,N,K,bn", [(256, 64, 64, 64), (4160, 768, 768, 192), (4160, 768, 3072, 256), (1000, 256, 384, 128),
                                       (130, 96, 128, 96), (385, 512, 192, 256), (2049, 2304, 768, 192)])
def test_gemm_pair_matches_single_cta_and_torch(cuda_dev, M, N, K, bn):
    g = torch.Generator(device="cpu").manual_seed(M + N + K + 1)
    A = (torch.randn(M, K, generator=g) * 0.5).to(cuda_dev).bfloat16()
    B = (torch.randn(N, K, generator=g) * 0.1).to(cuda_dev).bfloat16()
    bias = torch.randn(N, generator=g).to(cuda_dev)
    resid = torch.randn(M, N, generator=g).to(cuda_dev)
    ref = resid + A.float() @ B.float().t() + bias
    one = G.gemm_resid(A, B, bias, resid, block_n=bn, pair=0)
    two = G.gemm_resid(A, B, bias, resid, block_n=bn, pair=1)
    assert G.rel_err(one, ref) < 2e-5
    assert G.rel_err(two, ref) < 2e-5, f"pair rel err {G.rel_err(two, ref)}"
    assert torch.equal(one, two)


@pytest.mark.parametrize("N,s,Cin,Cout,gelu", [(64, 8, 352, 224, True), (64, 8, 384, 224, False), (3, 12, 384, 224, True), (33, 8, 128, 80, True),
                                                  (9, 16, 64, 96, True), (40, 6, 128, 64, True), (21, 4, 64, 48, False), (1, 16, 576, 224, True)])
def test_conv3x3_pair_matches_single_cta(cuda_dev, N, s, Cin, Cout, gelu):
    """CTA-pair launch of the 3x3 conv (two 128-row tiles per MMA, each CTA stages half of the weight tile), with and
    without the haloed-box A reuse, odd tile counts (an empty second CTA) and ragged image groups: bit-identical to the
    one-CTA launch."""
    g = torch.Generator(device="cpu").manual_seed(N * 1000 + s * 100 + Cin + 9)
    x = torch.randn(N, s, s, Cin, generator=g).to(cuda_dev).bfloat16()
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) * (1.0 / (3 * Cin ** 0.5))).to(cuda_dev)
    b = torch.randn(Cout, generator=g).to(cuda_dev)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.bfloat16().float(), b, padding=1)
    if gelu:
        ref = F.gelu(ref)
    ref = ref.permute(0, 2, 3, 1)
    one = G.conv3x3(x, w, b, gelu=gelu, impl=0)
    two = G.conv3x3(x, w, b, gelu=gelu, impl=2)
    assert G.rel_err(two, ref) < 3e-5, G.rel_err(two, ref)
    assert torch.equal(one, two)


@pytest.mark.parametrize("N,T,H", [(1, 17, 2), (3, 65, 2), (64, 65, 12), (2, 145, 12), (5, 129, 3), (1, 192, 1), (4, 128, 4),
                                   (64, 145, 12), (3, 257, 16), (2, 193, 2), (2, 325, 3), (1, 384, 2), (7, 1, 1),
                                   (2, 130, 2), (2, 264, 3), (1, 136, 1), (32, 257, 16), (2, 137, 2), (5, 65, 3), (9, 65, 4)])
@pytest.mark.parametrize("impl", [0, 1, 2, 3], ids=["mma_sync", "tcgen05", "tcgen05_lite", "tcgen05_duo"])
def test_attention_matches_torch(cuda_dev, N, T, H, impl):
    """Both attention kernels against fp32 softmax(q k^T / sqrt(64)) v of the same bf16 q, k, v (P is rounded to bf16 before
    the second product in both kernels)."""
    g = torch.Generator(device="cpu").manual_seed(N * 100 + T + H)
    qkv = (torch.randn(N * T, 3 * H * 64, generator=g) * 1.5).to(cuda_dev).bfloat16()
    q, k, v = qkv.float().reshape(N, T, 3, H, 64).permute(2, 0, 3, 1, 4)
    ref = (((q @ k.transpose(-2, -1)) * 0.125).softmax(-1) @ v).transpose(1, 2).reshape(N * T, H * 64)
    out = G.attention(qkv, N, T, H, impl=impl).float()
    err = G.rel_err(out, ref)
    assert err < 6e-3, f"impl={impl} rel err {err}"
    assert torch.isfinite(out).all()


@pytest.mark.parametrize("M,N,K,bn,gelu", [(4160, 2304, 768, 192, False), (4160, 3072, 768, 256, True), (2100, 1536, 512, 128, True),
                                           (9252, 1024, 1024, 256, False), (4160, 768, 256, 64, False), (300, 512, 128, 128, True)])
def test_gemm_bf16_out_variants_agree(cuda_dev, M, N, K, bn, gelu):
    """The bf16 store phases of the engine (QKV / fc1 style): one tile per CTA or one-CTA persistent (0), CTA pairs (1) and the
    persistent CTA-pair kernel (2: double-buffered tensor memory, tiles round-robin over 74 clusters) run the same k order, so
    they must agree bit for bit; all against the fp32 product of the same bf16 operands."""
    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g).to(cuda_dev).bfloat16()
    B = (torch.randn(N, K, generator=g) * 0.05).to(cuda_dev).bfloat16()
    bias = torch.randn(N, generator=g).to(cuda_dev)
    ref = A.float() @ B.float().t() + bias
    if gelu:
        ref = F.gelu(ref)
    outs = [G.gemm_bf16_out(A, B, bias, block_n=bn, gelu=gelu, variant=v) for v in (0, 1, 2)]
    for o in outs:
        assert G.rel_err(o.float(), ref) < 4e-3, G.rel_err(o.float(), ref)
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
