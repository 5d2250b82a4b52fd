"""How fast does cuBLAS (torch.matmul, bf16) run the encoder GEMM shapes of batch 64 / 256?  Context for DESIGN.md only -
the product path never calls it."""
import torch
torch.backends.cuda.matmul.allow_bf16_reduced_precision_reduction = True
dev = torch.device("cuda:0")
def run(M, N, K, reps=50):
    a = torch.randn(M, K, device=dev, dtype=torch.bfloat16); b = torch.randn(N, K, device=dev, dtype=torch.bfloat16)
    for _ in range(5): c = a @ b.t()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps): c = a @ b.t()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    print(f"M={M:6d} N={N:5d} K={K:5d}: {us:7.2f} us  {2.0*M*N*K/us/1e6:7.1f} TFLOP/s")
for M in (4160, 16640):
    for (N, K) in ((2304, 768), (768, 768), (3072, 768), (768, 3072)):
        run(M, N, K)
