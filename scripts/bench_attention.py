"""Device time of the two attention kernels at the bench geometries (CUDA events, one launch per call incl. its sync)."""
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import ctypes as C  # noqa: E402

from textmae_image_compression_b200 import _native  # noqa: E402

lib = _native.load()
for (N, T, H) in [(64, 65, 12), (64, 145, 12), (32, 257, 16), (24, 145, 12), (12, 257, 16)]:
    qkv = (torch.randn(N * T, 3 * H * 64, device="cuda") * 1.5).bfloat16()
    out = torch.zeros(N * T, H * 64, dtype=torch.bfloat16, device="cuda")
    res = {}
    for impl in (0, 1):
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        for _ in range(3):
            rc = lib.tmae_attention_bf16(C.c_void_p(qkv.data_ptr()), C.c_void_p(out.data_ptr()), N, T, H, impl, st)
            assert rc == 0, _native.last_error()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e9
        for _ in range(20):
            e0.record()
            lib.tmae_attention_bf16(C.c_void_p(qkv.data_ptr()), C.c_void_p(out.data_ptr()), N, T, H, impl, st)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) * 1e3)
        res[impl] = best
    print(f"N={N} T={T} H={H}: mma.sync {res[0]:.1f} us, tcgen05 {res[1]:.1f} us")
