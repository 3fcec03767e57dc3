"""Native MLP forward (ldpc_mlp_* of include/ldpc_b200.h): Linear + tanh chains on the tensor
cores with fp32-equivalent accuracy (exact binary16 plane splitting, csrc/mlp.cu).  Replaces the
ATen addmm/tanh calls under the reference's LLR estimators (pytorch/nn/llr.py:46-73)."""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _native as N


MODES = {"auto": 0, "per_layer": 1, "chain": 2, "chain_pairs": 3}       # LDPC_MLP_* of include/ldpc_b200.h


class NativeMLP:
    """weights[l]: [out_l, in_l] float32 (nn.Linear.weight), biases[l]: [out_l] or None,
    activations[l]: True -> tanh after layer l (default: every layer but the last).
    mode: "auto" (default: the single-launch L2-resident chain where the shape allows it - on cta_group::2 CTA pairs if the
    device can hold every cluster, else on single SMs), "per_layer" (one launch per layer), "chain" (single SMs; raises if the
    network cannot run it) or "chain_pairs" (4 % faster than "chain", 1.4x its DRAM traffic); the results are bit-identical."""

    def __init__(self, weights, biases=None, activations=None, splits=2, chunk_rows=0, device=None, mode="auto"):
        N.require_cuda()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        ws = [np.ascontiguousarray(_np(w), dtype=np.float32) for w in weights]
        nl = len(ws)
        bs = [None] * nl if biases is None else [None if b is None else np.ascontiguousarray(_np(b), dtype=np.float32) for b in biases]
        dims = [ws[0].shape[1]] + [w.shape[0] for w in ws]
        for l in range(nl):
            if ws[l].shape[1] != dims[l]:
                raise ValueError(f"layer {l}: weight is {ws[l].shape}, expected [*, {dims[l]}]")
            if bs[l] is not None and bs[l].shape != (dims[l + 1],):
                raise ValueError(f"layer {l}: bias must be [{dims[l + 1]}]")
        self.dims, self.splits = dims, int(splits)
        dims_a = (ctypes.c_int32 * (nl + 1))(*dims)
        w_a = (ctypes.c_void_p * nl)(*[w.ctypes.data for w in ws])
        b_a = (ctypes.c_void_p * nl)(*[None if b is None else b.ctypes.data for b in bs])
        act_a = None
        if activations is not None:
            act_a = (ctypes.c_int32 * nl)(*[1 if a else 0 for a in activations])
        h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            N.check(N.lib().ldpc_mlp_create(nl, dims_a, w_a, b_a, act_a, self.splits, int(chunk_rows), ctypes.byref(h)))
        self._h = h
        self.mode = "auto"
        if mode != "auto":
            self.set_mode(mode)

    def set_mode(self, mode):
        if mode not in MODES:
            raise ValueError(f"mode must be one of {sorted(MODES)}")
        N.check(N.lib().ldpc_mlp_set_mode(self._h, MODES[mode]))
        self.mode = mode

    def __call__(self, x, stream=None):
        """x: CUDA float32 [B, dims[0]] -> CUDA float32 [B, dims[-1]]."""
        if not x.is_cuda:
            raise ValueError("x must be a CUDA tensor (there is no CPU fallback)")
        if x.device != self.device:
            raise ValueError(f"x is on {x.device}, this handle's weights are on {self.device}")
        if x.dim() != 2 or x.shape[1] != self.dims[0]:
            raise ValueError(f"x must be [B,{self.dims[0]}], got {tuple(x.shape)}")
        x = x.to(dtype=torch.float32).contiguous()
        y = torch.empty((x.shape[0], self.dims[-1]), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            s = torch.cuda.current_stream(x.device).cuda_stream if stream is None else stream
            N.check(N.lib().ldpc_mlp_forward(self._h, x.data_ptr(), x.shape[0], y.data_ptr(), ctypes.c_void_p(s)))
        return y

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                N.lib().ldpc_mlp_destroy(h)
            except Exception:
                pass
            self._h = None


def _np(a):
    return a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
