"""Bring-up micro-benchmark of the tcgen05 engine: per-phase timing of representative layer shapes."""
import os, sys
os.environ["TMAE_GEMM_TIMING"] = "1"
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from tests import gpu_util as G
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
shapes = [(4160, 2304, 768, 256), (4160, 2304, 768, 128), (4160, 768, 768, 192), (4160, 768, 768, 128), (4160, 768, 768, 64),
          (4160, 3072, 768, 256), (4160, 768, 3072, 192), (4096, 32, 128, 32), (4096, 224, 576, 112), (4096, 224, 576, 224)]
for (M, N, K, bn) in shapes:
    A = torch.randn(M, K, generator=g).to(dev).bfloat16(); B = (torch.randn(N, K, generator=g) * 0.05).to(dev).bfloat16()
    bias = torch.randn(N, generator=g).to(dev)
    out = G.gemm(A, B, bias, block_n=bn, impl=0)
    ref = A.float() @ B.float().t() + bias
    print(f"M={M} N={N} K={K} bn={bn} rel_err={G.rel_err(out, ref):.2e}", flush=True)
for (N, s, Cin, Cout) in [(64, 8, 576, 224), (64, 8, 80, 32), (64, 8, 384, 384)]:
    x = torch.randn(N, s, s, Cin, generator=g).to(dev).bfloat16()
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) * 0.02).to(dev); b = torch.randn(Cout, generator=g).to(dev)
    out = G.conv3x3(x, w, b, gelu=True, impl=0)
    print(f"conv N={N} s={s} Cin={Cin} Cout={Cout} ok", flush=True)
