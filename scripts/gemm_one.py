"""One engine launch for ncu: python scripts/gemm_one.py gemm M N K bn | conv N s Cin Cout"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from tests import gpu_util as G
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
if sys.argv[1] == "gemm":
    M, N, K, bn = map(int, sys.argv[2:6])
    A = torch.randn(M, K, generator=g).to(dev).bfloat16(); B = (torch.randn(N, K, generator=g) * 0.05).to(dev).bfloat16()
    bias = torch.randn(N, generator=g).to(dev)
    for _ in range(2):
        out = G.gemm(A, B, bias, block_n=bn, impl=0)
else:
    N, s, Cin, Cout = map(int, sys.argv[2:6])
    x = torch.randn(N, s, s, Cin, generator=g).to(dev).bfloat16()
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) * 0.02).to(dev); b = torch.randn(Cout, generator=g).to(dev)
    for _ in range(2):
        out = G.conv3x3(x, w, b, gelu=True, impl=0)
print("ok", float(out.float().abs().mean()))
