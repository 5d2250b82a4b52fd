"""`MCM` - host-side mirror of the reference module for the compression forward path.

Same constructor keywords, state-dict names and `forward(imgs, total_scores)` contract as
`models.Compression.MCM.MCM` (/root/reference/models/Compression/MCM.py:25-52, 445-452, 590-634, 714-803),
but everything up to the rate runs in libtmae_b200.so (hand-written sm_100a kernels) through the C ABI of
include/tmae.h.  PyTorch is used for device memory, streams and the module plumbing only.

`forward` returns a superset of the reference dict:
    "likelihoods": {"y": f32 [N,Cy,s,s], "z": f32 [N,Cz,s/4,s/4]}   (same keys / shapes as MCM.py:801,
                    channels-last strides) - `RateDistortionLoss` (loss/rd_loss.py:15-20) consumes it unchanged
    "latents":     {"y_sym": i32, "z_sym": i32, "y_hat": f32, "z_hat": f32}
    "bpp":         f32 [N]   per-image rate (rd_loss.py formula with N = 1)
    "rate_sums":   f64 [2]   {sum log2 likelihood, pixels} for the data-parallel aggregate
    "ids_restore": i64 [N,L], "ids_keep": i64 [N,K], "ids_shuffle": i64 [N,L]
The reconstruction half ("loss", "x_hat": g_s, MAE decoder, SSIM/L1/VGG) is outside this path.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Iterator, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _native
from .config import PathConfig


class MCM(nn.Module):
    def __init__(self, img_size=224, patch_size=16, in_chans=3, encoder_embed_dim=768, encoder_depth=12,
                 encoder_num_heads=12, decoder_embed_dim=512, decoder_depth=8, decoder_num_heads=16, mlp_ratio=4.0,
                 norm_layer=None, norm_pix_loss=False, latent_depth=384, hyperprior_depth=192, num_slices=12,
                 num_keep_patches=144, *, skip_dead_lrp: bool = False, debug_simt: bool = False,
                 softmax_isa: Optional[int] = None, extra_outputs: bool = False, share_sm: bool = False,
                 precise: Optional[str] = None):
        super().__init__()
        self.cfg = PathConfig(img_size=img_size, patch_size=patch_size, in_chans=in_chans,
                              encoder_embed_dim=encoder_embed_dim, encoder_depth=encoder_depth,
                              encoder_num_heads=encoder_num_heads, decoder_embed_dim=decoder_embed_dim,
                              decoder_depth=decoder_depth, decoder_num_heads=decoder_num_heads, mlp_ratio=mlp_ratio,
                              latent_depth=latent_depth, hyperprior_depth=hyperprior_depth, num_slices=num_slices,
                              num_keep_patches=num_keep_patches)
        self.cfg.validate()                      # same failure classes as the reference (ValueError / RuntimeError)
        self.num_keep_patches = num_keep_patches
        self.norm_pix_loss = norm_pix_loss
        self.skip_dead_lrp = skip_dead_lrp
        self.debug_simt = debug_simt
        self.share_sm = share_sm                 # several handles/streams in flight on this GPU (see TMAE_FLAG_SHARE_SM)
        self.extra_outputs = extra_outputs       # also return y, z, mu, sigma, x_remain (parity tests)
        # accuracy mode: None = bf16 operands everywhere (throughput); "rate" = split-bf16 (fp32-equivalent products) for
        # g_a and every entropy-model conv; "all" = the encoder too -> symbols match the fp32 reference up to ties
        if precise not in (None, "rate", "all"):
            raise ValueError("precise must be None, 'rate' or 'all'")
        self.precise = precise
        if softmax_isa is None:
            # lane order of the ATen CPU softmax the reference's host routine would have used on this machine
            softmax_isa = 16 if "AVX512" in torch.backends.cpu.get_cpu_capability().upper() else 8
        self.softmax_isa = softmax_isa
        self._weights: Dict[str, torch.Tensor] = {}
        self._handle = None
        self._handle_device = None
        self._dirty = True

    # ------------------------------------------------------------------ module plumbing
    def parameters(self, recurse: bool = True) -> Iterator[torch.Tensor]:
        return iter(self._weights.values())

    def named_parameters(self, prefix: str = "", recurse: bool = True, remove_duplicate: bool = True):
        return iter(self._weights.items())

    def state_dict(self, *args, **kwargs):
        return dict(self._weights)

    def load_state_dict(self, state_dict, strict: bool = True):
        self._weights = {k: v.detach().clone().float() if v.is_floating_point() else v.detach().clone()
                         for k, v in state_dict.items()}
        self._dirty = True
        return self

    @classmethod
    def from_state_dict(cls, num_keep_patches, state_dict, **kwargs):     # MCM.py:448-452
        net = cls(num_keep_patches=num_keep_patches, **kwargs)
        net.load_state_dict(state_dict)
        return net

    def _apply(self, fn, recurse=True):
        self._weights = {k: fn(v) for k, v in self._weights.items()}
        self._dirty = True
        return self

    def train(self, mode: bool = True):
        if mode:
            raise NotImplementedError("MCM (B200 path) is inference-only: additive-noise quantisation and backward "
                                      "are outside the compression forward path (SURVEY H7)")
        return super().train(False)

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def _release(self):
        if self._handle is not None:
            _native.load().tmae_destroy(self._handle)
            self._handle = None

    # ------------------------------------------------------------------ native handle
    def _device(self) -> torch.device:
        for v in self._weights.values():
            return v.device
        raise RuntimeError("no weights loaded (call load_state_dict first)")

    def _ensure_handle(self):
        dev = self._device()
        if dev.type != "cuda":
            raise RuntimeError("MCM (B200 path) runs on CUDA only - move the module with .cuda(); there is no CPU path")
        if self._handle is not None and not self._dirty and self._handle_device == dev:
            return
        lib = _native.load()
        self._release()
        c = self.cfg
        flags = ((_native.FLAG_SKIP_DEAD_LRP if self.skip_dead_lrp else 0) | (_native.FLAG_DEBUG_SIMT if self.debug_simt else 0)
                 | (_native.FLAG_SHARE_SM if self.share_sm else 0)
                 | {None: 0, "rate": _native.FLAG_PRECISE_RATE, "all": _native.FLAG_PRECISE_ALL}[self.precise])
        cfg = _native.TmaeConfig(c.img_size, c.patch_size, c.in_chans, c.encoder_embed_dim, c.encoder_depth,
                                 c.encoder_num_heads, c.decoder_embed_dim, c.mlp_ratio, c.latent_depth,
                                 c.hyperprior_depth, c.num_slices, c.num_keep_patches, c.ln_eps, self.softmax_isa, flags)
        with torch.cuda.device(dev):
            hp = C.c_void_p()
            _native.check(lib.tmae_create(C.byref(cfg), C.byref(hp)), None, RuntimeError)
            self._handle = hp
            self._handle_device = dev
            ign = C.c_int(0)
            for name, t in self._weights.items():
                if not t.is_floating_point():
                    continue
                t = t.contiguous()
                shape = (C.c_int64 * max(t.dim(), 1))(*t.shape)
                _native.check(lib.tmae_set_weight(hp, name.encode(), C.c_void_p(t.data_ptr()), 0, t.dim(), shape,
                                                  C.byref(ign)), hp)
            _native.check(lib.tmae_finalize_weights(hp), hp, RuntimeError)
        self._dirty = False

    # ------------------------------------------------------------------ helpers
    def _check_inputs(self, imgs, total_scores):
        c = self.cfg
        if imgs.dim() != 4 or imgs.shape[1] != c.in_chans:
            raise AssertionError(f"expected imgs [N,{c.in_chans},S,S], got {tuple(imgs.shape)}")
        if imgs.shape[2] != c.img_size or imgs.shape[3] != c.img_size:
            raise AssertionError(f"Input image size ({imgs.shape[2]}*{imgs.shape[3]}) doesn't match model "
                                 f"({c.img_size}*{c.img_size}).")                      # timm PatchEmbed assert
        if c.num_keep_patches > total_scores.shape[1]:
            raise ValueError("Number of patches should not be greater than the length of scores")   # MCM.py:374-376
        if total_scores.shape[1] != c.num_patches or total_scores.shape[0] != imgs.shape[0]:
            raise ValueError(f"total_scores must be [N,{c.num_patches}], got {tuple(total_scores.shape)}")

    def _alloc_outputs(self, N, dev, encoder=True, rate=True, extra=False):
        c = self.cfg
        s, s4 = c.side, c.side // 4
        t: Dict[str, torch.Tensor] = {}
        if rate:
            t["y_likelihoods"] = torch.empty((N, s, s, c.latent_depth), dtype=torch.float32, device=dev)
            t["z_likelihoods"] = torch.empty((N, s4, s4, c.hyperprior_depth), dtype=torch.float32, device=dev)
            t["y_symbols"] = torch.empty((N, s, s, c.latent_depth), dtype=torch.int32, device=dev)
            t["z_symbols"] = torch.empty((N, s4, s4, c.hyperprior_depth), dtype=torch.int32, device=dev)
            t["y_hat"] = torch.empty((N, s, s, c.latent_depth), dtype=torch.float32, device=dev)
            t["z_hat"] = torch.empty((N, s4, s4, c.hyperprior_depth), dtype=torch.float32, device=dev)
            t["bpp"] = torch.empty((N,), dtype=torch.float32, device=dev)
            t["rate_sums"] = torch.empty((2,), dtype=torch.float64, device=dev)
            if extra:
                t["y"] = torch.empty((N, s, s, c.latent_depth), dtype=torch.float32, device=dev)
                t["z"] = torch.empty((N, s4, s4, c.hyperprior_depth), dtype=torch.float32, device=dev)
                t["mu"] = torch.empty((N, s, s, c.latent_depth), dtype=torch.float32, device=dev)
                t["sigma"] = torch.empty((N, s, s, c.latent_depth), dtype=torch.float32, device=dev)
        if encoder:
            t["ids_shuffle"] = torch.empty((N, c.num_patches), dtype=torch.int64, device=dev)
            t["ids_restore"] = torch.empty((N, c.num_patches), dtype=torch.int64, device=dev)
            t["ids_keep"] = torch.empty((N, c.num_keep_patches), dtype=torch.int64, device=dev)
            if extra or not rate:
                t["x_remain"] = torch.empty((N, c.num_keep_patches, c.encoder_embed_dim), dtype=torch.float32, device=dev)
        o = _native.TmaeOutputs()
        for k, v in t.items():
            setattr(o, k, v.data_ptr())
        return t, o

    @staticmethod
    def _nchw(t):          # channels-last storage viewed with the reference's [N,C,h,w] shape
        return t.permute(0, 3, 1, 2)

    def _pack_result(self, t):
        out = {
            "likelihoods": {"y": self._nchw(t["y_likelihoods"]), "z": self._nchw(t["z_likelihoods"])},
            "latents": {"y_sym": self._nchw(t["y_symbols"]), "z_sym": self._nchw(t["z_symbols"]),
                        "y_hat": self._nchw(t["y_hat"]), "z_hat": self._nchw(t["z_hat"])},
            "bpp": t["bpp"], "rate_sums": t["rate_sums"],
        }
        for k in ("ids_restore", "ids_keep", "ids_shuffle", "x_remain"):
            if k in t:
                out[k] = t[k]
        for k in ("y", "z", "mu", "sigma"):
            if k in t:
                out[k] = self._nchw(t[k])
        return out

    # ------------------------------------------------------------------ the path
    @torch.no_grad()
    def forward(self, imgs: torch.Tensor, total_scores: torch.Tensor, need_recon: bool = False):
        """MCM.forward (MCM.py:714-803), rate half."""
        if need_recon:
            raise NotImplementedError("reconstruction (g_s, MAE decoder, losses) is outside the compression forward "
                                      "path of this library (SURVEY 8f-2)")
        self._ensure_handle()
        dev = self._handle_device
        self._check_inputs(imgs, total_scores)
        imgs = imgs.to(device=dev, dtype=torch.float32).contiguous()
        total_scores = total_scores.to(device=dev, dtype=torch.float32).contiguous()
        N = imgs.shape[0]
        lib = _native.load()
        with torch.cuda.device(dev):
            t, o = self._alloc_outputs(N, dev, extra=self.extra_outputs)
            st = torch.cuda.current_stream(dev).cuda_stream
            _native.check(lib.tmae_forward(self._handle, C.c_void_p(imgs.data_ptr()), C.c_void_p(total_scores.data_ptr()),
                                           N, C.byref(o), C.c_void_p(st)), self._handle, RuntimeError)
        return self._pack_result(t)

    @torch.no_grad()
    def forward_encoder(self, imgs, total_scores) -> Tuple[torch.Tensor, torch.Tensor]:
        """MCM.forward_encoder (MCM.py:590-634) -> (x_remain [N,K,C], ids_restore [N,L]).
        (The reference returns ids_restore on the CPU; here it stays on the device - no host sync.)"""
        self._ensure_handle()
        dev = self._handle_device
        self._check_inputs(imgs, total_scores)
        imgs = imgs.to(device=dev, dtype=torch.float32).contiguous()
        total_scores = total_scores.to(device=dev, dtype=torch.float32).contiguous()
        N = imgs.shape[0]
        lib = _native.load()
        with torch.cuda.device(dev):
            t, o = self._alloc_outputs(N, dev, encoder=True, rate=False)
            st = torch.cuda.current_stream(dev).cuda_stream
            _native.check(lib.tmae_forward_encoder(self._handle, C.c_void_p(imgs.data_ptr()),
                                                   C.c_void_p(total_scores.data_ptr()), N, C.byref(o), C.c_void_p(st)),
                          self._handle, RuntimeError)
        self._last_encoder = t
        return t["x_remain"], t["ids_restore"]

    @torch.no_grad()
    def forward_from_latent(self, y_nchw: torch.Tensor):
        """Teacher-forced rate half (MCM.py:739-787) from y = g_a(...) [N,Cy,s,s]."""
        self._ensure_handle()
        dev = self._handle_device
        y = y_nchw.to(device=dev, dtype=torch.float32).permute(0, 2, 3, 1).contiguous()
        N = y.shape[0]
        lib = _native.load()
        with torch.cuda.device(dev):
            t, o = self._alloc_outputs(N, dev, encoder=False, rate=True, extra=True)
            st = torch.cuda.current_stream(dev).cuda_stream
            _native.check(lib.tmae_forward_from_latent(self._handle, C.c_void_p(y.data_ptr()), N, C.byref(o),
                                                       C.c_void_p(st)), self._handle, RuntimeError)
        return self._pack_result(t)

    @torch.no_grad()
    def forward_host(self, imgs_cpu: torch.Tensor, scores_cpu: torch.Tensor, bpp_out: Optional[torch.Tensor] = None,
                     stream: Optional[torch.cuda.Stream] = None):
        """End-to-end call with HOST buffers (pinned for full speed): H2D copies, forward and the D2H read of the
        per-image bpp are all enqueued on `stream`; returns the (pinned) CPU bpp tensor - valid after the stream
        is synchronised."""
        self._ensure_handle()
        dev = self._handle_device
        if imgs_cpu.device.type != "cpu" or scores_cpu.device.type != "cpu":
            raise ValueError("forward_host takes CPU tensors")
        self._check_inputs(imgs_cpu, scores_cpu)
        imgs_cpu = imgs_cpu.contiguous()
        scores_cpu = scores_cpu.contiguous()
        N = imgs_cpu.shape[0]
        if bpp_out is None:
            bpp_out = torch.empty((N,), dtype=torch.float32).pin_memory()
        lib = _native.load()
        with torch.cuda.device(dev):
            st = (stream or torch.cuda.current_stream(dev)).cuda_stream
            _native.check(lib.tmae_forward_host(self._handle, C.c_void_p(imgs_cpu.data_ptr()),
                                                C.c_void_p(scores_cpu.data_ptr()), N, C.c_void_p(bpp_out.data_ptr()),
                                                None, None, C.c_void_p(st)), self._handle, RuntimeError)
        return bpp_out

    @torch.no_grad()
    def get_ids_shuffle(self, total_scores: torch.Tensor) -> torch.Tensor:
        """MCM.get_ids_shuffle (MCM.py:364-423) on the GPU: int64 [N, L] (device tensor)."""
        K = self.num_keep_patches
        if K > total_scores.shape[1]:
            raise ValueError("Number of patches should not be greater than the length of scores")
        dev = total_scores.device if total_scores.is_cuda else self._device()
        sc = total_scores.to(device=dev, dtype=torch.float32).contiguous()
        N, L = sc.shape
        out = torch.empty((N, L), dtype=torch.int64, device=dev)
        lib = _native.load()
        with torch.cuda.device(dev):
            st = torch.cuda.current_stream(dev).cuda_stream
            _native.check(lib.tmae_mask_select(C.c_void_p(sc.data_ptr()), N, L, K, self.softmax_isa,
                                               C.c_void_p(out.data_ptr()), None, None, C.c_void_p(st)))
        return out

    def random_masking(self, x, total_scores):
        """MCM.random_masking (MCM.py:548-588): gather of the kept tokens with GPU-computed ids."""
        ids_shuffle = self.get_ids_shuffle(total_scores)
        ids_restore = torch.argsort(ids_shuffle, dim=1)
        ids_keep = ids_shuffle[:, : self.num_keep_patches].to(x.device)
        x_remain = torch.gather(x, 1, ids_keep.unsqueeze(-1).repeat(1, 1, x.shape[-1]))
        return x_remain, ids_restore

    # ------------------------------------------------------------------ small host-side pieces of the interface
    def aux_loss(self) -> torch.Tensor:
        """compressai CompressionModel.aux_loss() == EntropyBottleneck.loss(): sum |logits(quantiles) - target|
        (called by utils/engine.py:194).  Pure function of 11.7 k parameters - evaluated with torch ops."""
        w = self._weights
        q = w["entropy_bottleneck.quantiles"]
        logits = q
        for i in range(5):
            logits = torch.matmul(F.softplus(w[f"entropy_bottleneck._matrix{i}"]), logits) + w[f"entropy_bottleneck._bias{i}"]
            if i < 4:
                logits = logits + torch.tanh(w[f"entropy_bottleneck._factor{i}"]) * torch.tanh(logits)
        import math
        t = math.log(2 / 1e-9 - 1)
        target = torch.tensor([-t, 0.0, t], device=q.device, dtype=q.dtype)
        return torch.abs(logits - target).sum()

    def update(self, *args, **kwargs):
        raise NotImplementedError("CDF-table construction for rANS coding is outside this path (SURVEY 8f-1)")

    def compress(self, *args, **kwargs):
        raise NotImplementedError("bitstream emission is outside this path (north_star; SURVEY 8f-1)")

    def decompress(self, *args, **kwargs):
        raise NotImplementedError("bitstream decoding is outside this path (SURVEY 8f-1)")

    # ------------------------------------------------------------------ profiling hooks (bench.py)
    def profile(self, enable: bool, by_run: bool = False):
        """Device timing of the next forwards: per launch, or per run of consecutive same-family launches (by_run)."""
        self._ensure_handle()
        _native.check(_native.load().tmae_profile_enable(self._handle, (2 if by_run else 1) if enable else 0), self._handle)

    def profile_read(self):
        arr = (_native.TmaeProfileEntry * 16)()
        n = C.c_int(0)
        _native.check(_native.load().tmae_profile_read(self._handle, arr, 16, C.byref(n)), self._handle, RuntimeError)
        return [{"name": arr[i].name.decode(), "launches": arr[i].launches, "ms": arr[i].ms, "flops": arr[i].flops,
                 "bytes": arr[i].bytes} for i in range(n.value)]

    def profile_read_steps(self):
        arr = (_native.TmaeProfileStep * 512)()
        n = C.c_int(0)
        _native.check(_native.load().tmae_profile_read_steps(self._handle, arr, 512, C.byref(n)), self._handle, RuntimeError)
        return [{"name": arr[i].name.decode(), "ms": arr[i].ms, "flops": arr[i].flops, "ctas": arr[i].ctas,
                 "block_n": arr[i].block_n} for i in range(n.value)]

    def launch_count(self, N: int) -> int:
        self._ensure_handle()
        return int(_native.load().tmae_launch_count(self._handle, N))

    def reserve(self, N: int):
        self._ensure_handle()
        with torch.cuda.device(self._handle_device):
            _native.check(_native.load().tmae_reserve(self._handle, N), self._handle, RuntimeError)
