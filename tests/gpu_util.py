"""Helpers shared by the -m gpu tests: raw ctypes calls into libtmae_b200.so on torch CUDA tensors."""
import ctypes as C

import torch

from textmae_image_compression_b200 import _native


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def gemm(A_bf16, B_bf16, bias, block_n=0, impl=0):
    """C = A @ B^T + bias on the engine; A [M,K], B [N,K] bf16 -> f32 [M,N]."""
    M, K = A_bf16.shape
    N = B_bf16.shape[0]
    out = torch.zeros(M, N, dtype=torch.float32, device=A_bf16.device)
    rc = _native.load().tmae_gemm_bf16(ptr(A_bf16), ptr(B_bf16), ptr(bias), ptr(out), M, N, K, block_n, impl, stream())
    _native.check(rc, None, RuntimeError)
    torch.cuda.synchronize()
    return out


def conv3x3(x_nhwc_bf16, w, bias, gelu=False, impl=0):
    N, s, _, Cin = x_nhwc_bf16.shape
    Cout = w.shape[0]
    out = torch.zeros(N, s, s, Cout, dtype=torch.float32, device=w.device)
    rc = _native.load().tmae_conv3x3_bf16(ptr(x_nhwc_bf16.contiguous()), ptr(w.contiguous()), ptr(bias), ptr(out), N, s,
                                          Cin, Cout, 1 if gelu else 0, impl, stream())
    _native.check(rc, None, RuntimeError)
    torch.cuda.synchronize()
    return out


def attention(qkv_bf16, N, T, H, impl=0):
    """softmax(q k^T / 8) v over qkv bf16 [N*T, 3*H*64]; impl 0 = mma.sync kernel, 1 = tcgen05 kernel."""
    out = torch.zeros(N * T, H * 64, dtype=torch.bfloat16, device=qkv_bf16.device)
    rc = _native.load().tmae_attention_bf16(ptr(qkv_bf16.contiguous()), ptr(out), N, T, H, impl, stream())
    _native.check(rc, None, RuntimeError)
    torch.cuda.synchronize()
    return out


def gemm_bf16_out(A_bf16, B_bf16, bias, block_n=0, gelu=False, variant=0):
    """bf16 C = act(A @ B^T + bias): variant 0 one tile per CTA / one-CTA persistent, 1 CTA pairs, 2 persistent CTA pairs, 3 checker."""
    M, K = A_bf16.shape
    N = B_bf16.shape[0]
    out = torch.zeros(M, N, dtype=torch.bfloat16, device=A_bf16.device)
    rc = _native.load().tmae_gemm_bf16_out(ptr(A_bf16), ptr(B_bf16), ptr(bias), ptr(out), M, N, K, block_n, 1 if gelu else 0, variant, stream())
    _native.check(rc, None, RuntimeError)
    torch.cuda.synchronize()
    return out


def gemm_resid(A_bf16, B_bf16, bias, resid, block_n=0, pair=0, impl=0):
    """C = resid + A @ B^T + bias (proj / fc2 store phase); pair=1 -> CTA-pair (cta_group::2) kernel."""
    M, K = A_bf16.shape
    N = B_bf16.shape[0]
    out = torch.zeros(M, N, dtype=torch.float32, device=A_bf16.device)
    rc = _native.load().tmae_gemm_bf16_resid(ptr(A_bf16), ptr(B_bf16), ptr(bias), ptr(resid.contiguous()), ptr(out), M, N, K,
                                             block_n, pair, impl, stream())
    _native.check(rc, None, RuntimeError)
    torch.cuda.synchronize()
    return out


def gemm_split(A_f32, B_f32, bias, block_n=0, impl=0, planes=2):
    """Precise engine configuration: fp32 operands, split-bf16 planes, three tensor-core terms per product."""
    M, K = A_f32.shape
    N = B_f32.shape[0]
    out = torch.zeros(M, N, dtype=torch.float32, device=A_f32.device)
    rc = _native.load().tmae_gemm_split(ptr(A_f32.contiguous()), ptr(B_f32.contiguous()), ptr(bias), ptr(out), M, N, K, block_n,
                                        planes, impl, stream())
    _native.check(rc, None, RuntimeError)
    torch.cuda.synchronize()
    return out


def conv3x3_split(x_nhwc_f32, w, bias, gelu=False, impl=0, planes=2):
    N, s, _, Cin = x_nhwc_f32.shape
    Cout = w.shape[0]
    out = torch.zeros(N, s, s, Cout, dtype=torch.float32, device=w.device)
    rc = _native.load().tmae_conv3x3_split(ptr(x_nhwc_f32.contiguous()), ptr(w.contiguous()), ptr(bias), ptr(out), N, s,
                                           Cin, Cout, 1 if gelu else 0, planes, impl, stream())
    _native.check(rc, None, RuntimeError)
    torch.cuda.synchronize()
    return out


def mask_select(scores, K, isa=16):
    N, L = scores.shape
    dev = scores.device
    sh = torch.empty(N, L, dtype=torch.int64, device=dev)
    rs = torch.empty(N, L, dtype=torch.int64, device=dev)
    kp = torch.empty(N, K, dtype=torch.int64, device=dev)
    rc = _native.load().tmae_mask_select(ptr(scores.contiguous()), N, L, K, isa, ptr(sh), ptr(rs), ptr(kp), stream())
    _native.check(rc)
    torch.cuda.synchronize()
    return sh, rs, kp


def rel_err(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()
