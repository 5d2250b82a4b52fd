"""`HuffmanCoding` with the reference's interface (utils/huffman.py), backed by the host-side coder in libtmae_b200.so.

testing.py:73-76 codes `ids_restore` with it and adds `len(encoded_text) / num_pixels` to the reported bpp (:89).  The bit
string is the reference's, bit for bit (same heap tie-breaking, same left = "0" / right = "1" walk); `compress` / `decompress`
/ `encode` / `decode` / `codes` / `reverse_mapping` keep their meaning.  Unlike the reference object, `compress` starts from
a clean state on every call (the reference creates a new object per image, testing.py:74).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _native


class HuffmanCoding:
    def __init__(self):
        self._lib = _native.load()
        h = C.c_void_p()
        _native.check(self._lib.tmae_huffman_create(C.byref(h)))
        self._h = h
        self.codes = {}
        self.reverse_mapping = {}

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.tmae_huffman_destroy(self._h)
            self._h = None

    def _check(self, rc):
        if rc != _native.TMAE_OK:
            msg = self._lib.tmae_huffman_last_error(self._h)
            raise (MemoryError if rc == _native.TMAE_ENOMEM else ValueError)(msg.decode() if msg else f"error {rc}")

    def _load_codes(self):
        n = self._lib.tmae_huffman_num_symbols(self._h, None, 0)
        vals = (C.c_int64 * max(n, 1))()
        self._lib.tmae_huffman_num_symbols(self._h, vals, n)
        buf = C.create_string_buffer(max(n, 2) + 8)
        self.codes, self.reverse_mapping = {}, {}
        for v in vals[:n]:
            self._check(self._lib.tmae_huffman_code(self._h, v, buf, len(buf)))
            code = buf.value.decode()
            self.codes[int(v)] = code
            self.reverse_mapping[code] = int(v)

    def compress(self, tensor: torch.Tensor):
        """utils/huffman.py:141-157 -> (encoded_text, ori_shape, device)."""
        flat = tensor.detach().reshape(-1).to("cpu", torch.int64).contiguous()
        nbits = C.c_int64()
        self._check(self._lib.tmae_huffman_compress(self._h, C.c_void_p(flat.data_ptr()), flat.numel(), C.byref(nbits)))
        self._load_codes()
        buf = C.create_string_buffer(nbits.value + 1)
        self._check(self._lib.tmae_huffman_bits(self._h, C.cast(buf, C.c_void_p), nbits.value, 1))
        return buf.raw[:nbits.value].decode("ascii"), tensor.shape, tensor.device

    def compress_packed(self, tensor: torch.Tensor):
        """Same code, bits packed MSB-first into bytes (what a real container would store) -> (bytes, n_bits, ori_shape, device)."""
        flat = tensor.detach().reshape(-1).to("cpu", torch.int64).contiguous()
        nbits = C.c_int64()
        self._check(self._lib.tmae_huffman_compress(self._h, C.c_void_p(flat.data_ptr()), flat.numel(), C.byref(nbits)))
        self._load_codes()
        nbytes = (nbits.value + 7) // 8
        buf = C.create_string_buffer(nbytes + 1)
        self._check(self._lib.tmae_huffman_bits(self._h, C.cast(buf, C.c_void_p), nbytes, 0))
        return buf.raw[:nbytes], nbits.value, tensor.shape, tensor.device

    def encode(self, tensor: torch.Tensor) -> str:
        """utils/huffman.py:104-118 with the current code."""
        return "".join(self.codes[int(v)] for v in tensor.reshape(-1).tolist())

    def decode(self, encoded_text) -> torch.Tensor:
        """utils/huffman.py:120-139: str of '0'/'1' (or (bytes, n_bits) from compress_packed) -> 1-D int64 tensor."""
        if isinstance(encoded_text, tuple):
            data, nbits, as_chars = encoded_text[0], int(encoded_text[1]), 0
        else:
            data, nbits, as_chars = encoded_text.encode("ascii"), len(encoded_text), 1
        cap = max(nbits, 1)                      # every value costs >= 1 bit unless the alphabet has one symbol (code "")
        out = torch.empty(cap, dtype=torch.int64)
        n = C.c_int64()
        self._check(self._lib.tmae_huffman_decompress(self._h, data, nbits, as_chars, C.c_void_p(out.data_ptr()), cap, C.byref(n)))
        return out[:n.value].clone()

    def decompress(self, encoded_text, ori_shape, device) -> torch.Tensor:
        """utils/huffman.py:159-171."""
        return self.decode(encoded_text).to(device).view(ori_shape)


def side_info_bits(ids_restore: torch.Tensor) -> int:
    """len(compressed_ids_keep) of testing.py:75: the bits the reference adds to the bpp for `ids_restore` (:89)."""
    return len(HuffmanCoding().compress(ids_restore)[0])
