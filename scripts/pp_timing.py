"""TMAE_GEMM_TIMING=1 python scripts/pp_timing.py : phase stamps of the persistent CTA-pair kernel on the encoder shapes."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from tests import gpu_util as G
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
for (M, N, K, bn, gelu) in [(4160, 2304, 768, 192, False), (4160, 2304, 768, 256, False), (4160, 3072, 768, 256, True), (4160, 3072, 768, 256, False)]:
    A = torch.randn(M, K, generator=g).to(dev).bfloat16(); B = (torch.randn(N, K, generator=g) * 0.05).to(dev).bfloat16()
    bias = torch.randn(N, generator=g).to(dev)
    G.gemm_bf16_out(A, B, bias, block_n=bn, gelu=gelu, variant=2)
