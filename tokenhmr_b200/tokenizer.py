"""Tokenizer encoder: drop-in for the reference's `EncodeTokens`
(tokenization/models/vanilla_pose_vqvae.py:304-346): 6D body pose -> pose-token indices.

    enc = EncodeTokens(cfg, net)             # net: ckpt['net'] of tokenizer.pth ('encoder.*', 'quantizer.codebook'),
    code_idx = enc(pose6d)                   #      or the flat 'tokenizer.'-prefixed naming used in this repo
                                             # pose6d (B,21,6) CUDA/CPU fp32 -> (B*160,) int64 on the GPU

All arithmetic runs in libtokenhmr_b200.so (`thmr_tok_encode`) in split precision (fp32-grade: the output is an index and
has to equal the fp32 reference's): implicit-GEMM Conv1d layers on the tcgen05 kernel over split-fp16 operands and the
split-precision distance GEMM with the arg-min in its epilogue.  Python repacks the weights once and allocates
tensors; there is no CPU path.
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from . import _lib
from ._lib import check, lib
from .config import TokenHMRConfig
from .weights import split_weight

_CIN0 = 64   # kEncCin0 in csrc/tok_encoder.cuh


class EncodeTokens(nn.Module):
    def __init__(self, cfg: TokenHMRConfig, net: Dict[str, torch.Tensor], device: str | torch.device = "cuda:0"):
        super().__init__()
        self.cfg = cfg
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.ThmrError("EncodeTokens needs a CUDA device (there is no CPU fallback)")
        lib()
        if any(k.startswith("tokenizer.") for k in net):
            net = {k[len("tokenizer."):]: v for k, v in net.items() if k.startswith("tokenizer.")}
        self._keep: List[torch.Tensor] = []
        dev = self.device

        def conv(prefix: str, pad_cin: Optional[int] = None, gathered: bool = False) -> _lib.TokConv:
            w = net[prefix + ".weight"].detach().float()                   # [Cout, Cin, k]
            cout, cin, k = w.shape
            if pad_cin is not None:
                wp = torch.zeros(cout, pad_cin, k)
                wp[:, :cin] = w
                w, cin = wp, pad_cin
            # tap-major, split precision (csrc/strict.cuh): per tap [hi | hi | lo] of w * 2^8; the stride-2 conv's four
            # taps are gathered into ONE operand row on the device, so its weight is split as a single 4*cin-wide tap
            wt = w.permute(0, 2, 1).reshape(cout, k * cin).to(dev)
            wt = split_weight(wt, taps=(1 if gathered else k))
            b = net[prefix + ".bias"].detach().to(dev, torch.float32).contiguous()
            self._keep += [wt, b]
            return _lib.TokConv(wt.data_ptr(), b.data_ptr())

        d = _lib.TokEncoderDesc()
        d.joints, d.in_dim = cfg.tok_joints, 6
        d.width, d.depth, d.dilation_rate = cfg.tok_width, cfg.tok_depth, cfg.tok_dilation_rate
        d.size_mul, d.code_dim, d.nb_code = cfg.tok_size_mul, cfg.code_dim, cfg.nb_code
        e = "encoder.encoder"
        d.conv_in = conv(f"{e}.0", pad_cin=_CIN0)
        d.conv_up[0] = conv(f"{e}.3")
        idx = 5
        for u in range(1, cfg.tok_size_mul):
            d.conv_up[u] = conv(f"{e}.{idx + 1}")
            idx += 3
        d.conv_down = conv(f"{e}.{idx}.0", gathered=True)
        for k in range(cfg.tok_depth):
            d.res_conv1[k] = conv(f"{e}.{idx}.1.model.{k}.conv1")
            d.res_conv2[k] = conv(f"{e}.{idx}.1.model.{k}.conv2")
        d.conv_out = conv(f"{e}.{idx + 1}")
        cb = net["quantizer.codebook"].detach().to(dev, torch.float32).contiguous()
        self._keep.append(cb)
        self.codebook = cb
        d.codebook = cb.data_ptr()
        self._desc = d
        h = ctypes.c_void_p()
        check(lib().thmr_tok_encoder_create(ctypes.byref(d), ctypes.byref(h)))
        self._h = h
        self.num_tokens = lib().thmr_tok_encoder_num_tokens(h)
        self._ws: Optional[torch.Tensor] = None

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().thmr_tok_encoder_destroy(self._h)
                self._h = None
        except Exception:
            pass

    @torch.no_grad()
    def forward(self, x: torch.Tensor, return_latent: bool = False):
        """EncodeTokens.forward: x (B, 21, 6) -> code_idx (B*T,) int64 (and the (B*T, code_dim) latent on request)."""
        if x.dim() != 3 or x.shape[1] != self.cfg.tok_joints or x.shape[2] != 6:
            raise _lib.ThmrError(f"pose must be (B,{self.cfg.tok_joints},6), got {tuple(x.shape)}")
        B = x.shape[0]
        with torch.cuda.device(self.device):
            x = x.to(self.device, torch.float32).contiguous()
            need = lib().thmr_tok_encoder_workspace_bytes(self._h, B)
            if self._ws is None or self._ws.numel() < need:
                self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
            idx = torch.empty(B * self.num_tokens, dtype=torch.int64, device=self.device)
            lat = torch.empty(B * self.num_tokens, self.cfg.code_dim, dtype=torch.float32, device=self.device) \
                if return_latent else None
            check(lib().thmr_tok_encode(self._h, x.data_ptr(), B, idx.data_ptr(), lat.data_ptr() if lat is not None else None,
                                        self._ws.data_ptr(), torch.cuda.current_stream().cuda_stream))
        return (idx, lat) if return_latent else idx
