"""TMAE_ATTN_TIMING=1 python scripts/attn_timing.py N T H : per-phase clock64 stamps of CTA 0 of the tcgen05 attention kernel."""
import ctypes as C, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from textmae_image_compression_b200 import _native
lib = _native.load()
N, T, H = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (64, 65, 12)
qkv = (torch.randn(N * T, 3 * H * 64, device="cuda") * 1.5).bfloat16()
out = torch.zeros(N * T, H * 64, dtype=torch.bfloat16, device="cuda")
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for r in range(2):
    print(f"--- call {r}", file=sys.stderr, flush=True)
    rc = lib.tmae_attention_bf16(C.c_void_p(qkv.data_ptr()), C.c_void_p(out.data_ptr()), N, T, H, 1, st)
    assert rc == 0
