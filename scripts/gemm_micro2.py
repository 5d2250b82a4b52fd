import os, sys
os.environ["TMAE_GEMM_TIMING"] = "1"; os.environ["TMAE_TIMING_BF16"] = "1"
if len(sys.argv) > 1 and sys.argv[1] == "gelu": os.environ["TMAE_TIMING_GELU"] = "1"
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from tests import gpu_util as G
dev = torch.device("cuda:0"); g = torch.Generator().manual_seed(0)
for (M, N, K, bn) in [(4160, 2304, 768, 192), (4160, 2304, 768, 256), (4160, 3072, 768, 240), (4160, 3072, 768, 256), (4160, 3072, 768, 192)]:
    A = torch.randn(M, K, generator=g).to(dev).bfloat16(); B = (torch.randn(N, K, generator=g) * 0.05).to(dev).bfloat16()
    bias = torch.randn(N, generator=g).to(dev)
    G.gemm(A, B, bias, block_n=bn, impl=0)
