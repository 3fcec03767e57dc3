"""LdpcCode - a compiled parity-check matrix + the decode calls, over the C ABI.

Mirrors what BeliefPropagation.__init__/forward do in the reference (bp/bp.py:19-51) but
keeps the graph as sparse device tables inside an opaque native handle.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _native as N
from .codes import EdgeTables, detect_qc

_DTYPES = {torch.float32: N.F32, torch.float64: N.F64, torch.float16: N.F16}
_NP_DTYPES = {np.dtype(np.float32): N.F32, np.dtype(np.float64): N.F64, np.dtype(np.float16): N.F16}


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _update_id(update):
    if isinstance(update, str):
        try:
            return N.UPDATE_IDS[update.lower()]
        except KeyError:
            raise ValueError(f"unknown update rule {update!r}") from None
    return int(update)


class LdpcCode:
    """H compiled to device edge tables (replaces generate_masks, bp/masking.py:12-147)."""

    def __init__(self, H, qc_Z=0, qc_proto=None, device=None):
        N.require_cuda()
        H = np.asarray(H)
        if H.ndim != 2:
            raise ValueError("H must be 2-D")
        self.tables = EdgeTables.from_H(H)
        self.m, self.n, self.E = self.tables.m, self.tables.n, self.tables.E
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        if qc_Z and qc_proto is None:
            qc_proto = detect_qc(H, qc_Z)
            if qc_proto is None:
                qc_Z = 0
        proto = None if qc_proto is None else np.ascontiguousarray(qc_proto, dtype=np.int16)
        h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            N.check(N.lib().ldpc_code_create(
                self.tables.chk_ptr.ctypes.data, self.tables.chk_var.ctypes.data, self.m, self.n,
                int(qc_Z or 0), None if proto is None else proto.ctypes.data, ctypes.byref(h)))
        self._h = h
        info = N.CodeInfo()
        N.check(N.lib().ldpc_code_info(self._h, ctypes.byref(info)))
        self.max_dc, self.max_dv, self.kernel, self.qc_Z = info.max_dc, info.max_dv, info.kernel, info.qc_Z

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                N.lib().ldpc_code_destroy(h)
            except Exception:
                pass
            self._h = None

    def set_kernel(self, kernel):
        kid = {"generic": N.KERNEL_GENERIC, "qc": N.KERNEL_QC, "tiny": N.KERNEL_TINY, "qc_rt": N.KERNEL_QC_RT}.get(kernel, kernel)
        N.check(N.lib().ldpc_code_set_kernel(self._h, int(kid)))
        self.kernel = int(kid)

    def set_precision(self, precision):
        """'f32' (default, the reference's dtype) or 'f16' (two codewords per thread in half2;
        min-sum / normalized min-sum on code-specialised kernels only)."""
        pid = {"f32": N.PREC_F32, "f16": N.PREC_F16X2, "f16x2": N.PREC_F16X2}.get(precision, precision)
        N.check(N.lib().ldpc_code_set_precision(self._h, int(pid)))
        self.precision = int(pid)

    @property
    def packed_bytes(self):
        return (self.n + 7) // 8

    def decode(self, llr, iterations, clamp_value, update="sp", param=1.0, x0=None,
               want=("prob", "hard"), stream=None, early_exit=False):
        """llr: CUDA tensor [B,n] (f32/f64/f16), log(P1/P0).  Returns a dict of CUDA tensors
        for the names in `want`: prob, llr_post, hard, hard_packed, syndrome, x, iters_used.
        early_exit=True freezes a codeword once its hard decision satisfies every check."""
        if not llr.is_cuda:
            raise ValueError("llr must be a CUDA tensor (no CPU fallback); use decode_host for numpy input")
        if llr.device != self.device:
            raise ValueError(f"llr is on {llr.device}, this code's edge tables are on {self.device}")
        if llr.dim() != 2 or llr.shape[1] != self.n:
            raise ValueError(f"llr must be [B,{self.n}], got {tuple(llr.shape)}")
        if llr.dtype not in _DTYPES:
            llr = llr.float()
        llr = llr.contiguous()
        B = llr.shape[0]
        dev = llr.device
        out = {}
        mk = lambda shape, dt: torch.empty(shape, dtype=dt, device=dev)
        prob = mk((B, self.n), torch.float32) if "prob" in want else None
        post = mk((B, self.n), torch.float32) if "llr_post" in want else None
        hard = mk((B, self.n), torch.uint8) if "hard" in want else None
        packed = mk((B, self.packed_bytes), torch.uint8) if "hard_packed" in want else None
        synd = mk((B,), torch.int32) if "syndrome" in want else None
        xo = mk((B, self.E), torch.float32) if "x" in want else None
        used = mk((B,), torch.int32) if "iters_used" in want else None
        if x0 is not None:
            if tuple(x0.shape) != (B, self.E):
                raise ValueError(f"x0 must be [B,{self.E}]")
            x0 = x0.to(device=dev, dtype=torch.float32).contiguous()
        with torch.cuda.device(dev):
            s = torch.cuda.current_stream(dev).cuda_stream if stream is None else stream
            if early_exit or used is not None:
                dp = N.DecodeParams(ctypes.sizeof(N.DecodeParams), _DTYPES[llr.dtype], llr.data_ptr(), B, int(iterations),
                                    _update_id(update), float(clamp_value), float(param),
                                    *[None if v is None else v.data_ptr() for v in (x0, prob, post, hard, packed, synd, xo)],
                                    1 if early_exit else 0, 0, None if used is None else used.data_ptr())
                N.check(N.lib().ldpc_decode_ex(self._h, ctypes.byref(dp), ctypes.c_void_p(s)))
            else:
                N.check(N.lib().ldpc_decode(
                    self._h, _ptr(llr), _DTYPES[llr.dtype], B, int(iterations), _update_id(update),
                    float(clamp_value), float(param), _ptr(x0), _ptr(prob), _ptr(post), _ptr(hard),
                    _ptr(packed), _ptr(synd), _ptr(xo), ctypes.c_void_p(s)))
        for k, v in (("prob", prob), ("llr_post", post), ("hard", hard), ("hard_packed", packed),
                     ("syndrome", synd), ("x", xo), ("iters_used", used)):
            if v is not None:
                out[k] = v
        return out

    def sparse_weights(self, state, iterations):
        """Dense parameters of a reference BeliefPropagation state_dict (bp/bp.py:26-39: layers.{i}.0.input_weight
        [E,E], layers.{i}.0.llr_weight [1,n], final_layer.0.input_weight [n,E], final_layer.0.llr_weight [1,n]) ->
        the device tables of ldpc_decode_weighted.  Returns None when every weight is 1 (the unweighted kernels apply)."""
        T = self.tables
        A = lambda k: (state[k].detach().cpu().numpy() if isinstance(state[k], torch.Tensor) else np.asarray(state[k])).astype(np.float32)
        E, n, mdv = self.E, self.n, int(self.max_dv)
        dv = np.diff(T.var_ptr)
        vm_var = np.repeat(np.arange(n), dv)                       # variable of a variable-major edge
        pos = np.arange(E) - T.var_ptr[vm_var]                     # its position k inside the variable
        w_edge = np.ones((iterations, E, mdv), np.float32)
        w_llr = np.ones((iterations, n), np.float32)
        for i in range(iterations):
            W = A(f"layers.{i}.0.input_weight")
            w_llr[i] = A(f"layers.{i}.0.llr_weight").reshape(-1)
            for j in range(mdv):
                sel = np.nonzero((dv[vm_var] > j) & (pos != j))[0]        # out edges whose variable has a j-th edge
                w_edge[i, sel, j] = W[sel, T.cm_of_vm[T.var_ptr[vm_var[sel]] + j]]
        wf_edge = A("final_layer.0.input_weight")[vm_var, T.cm_of_vm]
        wf_llr = A("final_layer.0.llr_weight").reshape(-1)
        mask = np.ones((E, mdv), bool)
        for j in range(mdv):
            mask[:, j] = (dv[vm_var] > j) & (pos != j)
        if np.all(w_edge[:, mask] == 1) and np.all(w_llr == 1) and np.all(wf_edge == 1) and np.all(wf_llr == 1):
            return None
        dev = self.device
        return dict(w_edge=torch.as_tensor(w_edge).to(dev), w_llr=torch.as_tensor(w_llr).to(dev),
                    wf_edge=torch.as_tensor(np.ascontiguousarray(wf_edge)).to(dev), wf_llr=torch.as_tensor(wf_llr).to(dev),
                    iterations=int(iterations), stride=mdv)

    def decode_weighted(self, llr, weights, clamp_value, update="sp", param=1.0, want=("prob", "hard"), stream=None):
        """decode() with the reference's trainable weights (sparse_weights); runs on the generic kernel."""
        if not llr.is_cuda:
            raise ValueError("llr must be a CUDA tensor (no CPU fallback)")
        if llr.device != self.device:
            raise ValueError(f"llr is on {llr.device}, this code's edge tables are on {self.device}")
        if llr.dim() != 2 or llr.shape[1] != self.n:
            raise ValueError(f"llr must be [B,{self.n}], got {tuple(llr.shape)}")
        if llr.dtype not in _DTYPES:
            llr = llr.float()
        llr = llr.contiguous()
        B, dev = llr.shape[0], llr.device
        mk = lambda shape, dt: torch.empty(shape, dtype=dt, device=dev)
        prob = mk((B, self.n), torch.float32) if "prob" in want else None
        post = mk((B, self.n), torch.float32) if "llr_post" in want else None
        hard = mk((B, self.n), torch.uint8) if "hard" in want else None
        packed = mk((B, self.packed_bytes), torch.uint8) if "hard_packed" in want else None
        synd = mk((B,), torch.int32) if "syndrome" in want else None
        xo = mk((B, self.E), torch.float32) if "x" in want else None
        w = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in weights.items()}
        with torch.cuda.device(dev):
            s = torch.cuda.current_stream(dev).cuda_stream if stream is None else stream
            N.check(N.lib().ldpc_decode_weighted(
                self._h, _ptr(llr), _DTYPES[llr.dtype], B, int(w["iterations"]), _update_id(update), float(clamp_value), float(param),
                _ptr(w["w_edge"]), _ptr(w["w_llr"]), _ptr(w["wf_edge"]), _ptr(w["wf_llr"]), int(w["stride"]),
                _ptr(prob), _ptr(post), _ptr(hard), _ptr(packed), _ptr(synd), _ptr(xo), ctypes.c_void_p(s)))
        return {k: v for k, v in (("prob", prob), ("llr_post", post), ("hard", hard), ("hard_packed", packed),
                                  ("syndrome", synd), ("x", xo)) if v is not None}

    def count_errors(self, hard, ref_bits, k, llr=None, counters=None):
        """Exact link metrics (evaluate_quantized_snr.py:169-188) accumulated into an int64[5]
        CUDA tensor: uncoded errs, info errs, frame errs, bits, frames."""
        dev = hard.device
        if counters is None:
            counters = torch.zeros(5, dtype=torch.int64, device=dev)
        hard = hard.contiguous(); ref_bits = ref_bits.to(torch.uint8).contiguous()
        B, n = hard.shape
        dt = N.F32
        if llr is not None:
            llr = llr.contiguous(); dt = _DTYPES[llr.dtype]
        with torch.cuda.device(dev):
            N.check(N.lib().ldpc_count_errors(_ptr(llr), dt, _ptr(hard), _ptr(ref_bits), B, n, int(k),
                                              _ptr(counters), ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
        return counters


def decode_host(code: LdpcCode, llrs: np.ndarray, iterations, clamp_value, update="sp", param=1.0,
                want=("hard",), chunk=0):
    """numpy in / numpy out through ldpc_decode_host (chunked H2D -> decode -> D2H pipeline).
    llrs: [N,n] float64/float32/float16, C-contiguous."""
    llrs = np.ascontiguousarray(llrs)
    if llrs.dtype not in _NP_DTYPES:
        llrs = llrs.astype(np.float32)
    Nn, n = llrs.shape
    if n != code.n:
        raise ValueError(f"llrs must be [N,{code.n}]")
    hard = np.empty((Nn, n), np.uint8) if "hard" in want else None
    packed = np.empty((Nn, code.packed_bytes), np.uint8) if "hard_packed" in want else None
    post = np.empty((Nn, n), np.float32) if "llr_post" in want else None
    synd = np.empty((Nn,), np.int32) if "syndrome" in want else None
    P = lambda a: a.ctypes.data if a is not None else None
    with torch.cuda.device(code.device):
        N.check(N.lib().ldpc_decode_host(code._h, llrs.ctypes.data, _NP_DTYPES[llrs.dtype], Nn, int(iterations),
                                         _update_id(update), float(clamp_value), float(param),
                                         P(hard), P(packed), P(post), P(synd), int(chunk)))
    out = {}
    for k, v in (("hard", hard), ("hard_packed", packed), ("llr_post", post), ("syndrome", synd)):
        if v is not None:
            out[k] = v
    return out
