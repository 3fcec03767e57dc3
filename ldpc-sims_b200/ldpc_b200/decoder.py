"""LdpcCode - a compiled parity-check matrix + the decode calls, over the C ABI.

Mirrors what BeliefPropagation.__init__/forward do in the reference (bp/bp.py:19-51) but
keeps the graph as sparse device tables inside an opaque native handle.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np
import torch

from . import _native as N
from .codes import EdgeTables, auto_qc_block_size, detect_qc

_DTYPES = {torch.float32: N.F32, torch.float64: N.F64, torch.float16: N.F16, torch.int8: N.I8}
_NP_DTYPES = {np.dtype(np.float32): N.F32, np.dtype(np.float64): N.F64, np.dtype(np.float16): N.F16, np.dtype(np.int8): N.I8}


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

    def __init__(self, H, qc_Z="auto", qc_proto=None, device=None, specialize=False):
        """qc_Z: block size of a quasi-cyclic H (with qc_proto, or detected from H), 0 = treat H as unstructured,
        "auto" (default) = look for a block-circulant structure with Z >= 24 (codes.auto_qc_block_size).
        specialize=True: a quasi-cyclic code outside the built-in IEEE 802.11n family is compiled at run time into the
        code-specialised kernel (ldpc_b200.jit: one nvcc invocation per new prototype, cached) instead of running on the
        run-time-table kernel; raises jit.SpecializeError if the prototype does not fit the mapping or nvcc is missing."""
        N.require_cuda()
        H = np.asarray(H)
        if H.ndim != 2:
            raise ValueError("H must be 2-D")
        if isinstance(qc_Z, str):
            if qc_Z != "auto":
                raise ValueError("qc_Z must be an integer or 'auto'")
            qc_Z = auto_qc_block_size(H) if qc_proto is None else 0
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
        if specialize and proto is not None and qc_Z:
            from . import jit
            jit.specialize_qc(proto, int(qc_Z))               # registers the plug-in: ldpc_code_create below finds it
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
        kid = {"generic": N.KERNEL_GENERIC, "qc": N.KERNEL_QC, "tiny": N.KERNEL_TINY, "qc_rt": N.KERNEL_QC_RT,
               "qc_tma": N.KERNEL_QC_TMA}.get(kernel, kernel)
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
        """Dense parameters of a reference BeliefPropagation state_dict -> the device tables of ldpc_decode_weighted
        (codes.sparse_weights_from_reference_state).  Returns None when every weight is 1 (the unweighted kernels apply)."""
        from .codes import sparse_weights_from_reference_state, _weight_slots
        w = sparse_weights_from_reference_state(self.tables, state, iterations)
        used = _weight_slots(self.tables)[2]
        if np.all(w["w_edge"][:, used] == 1) and np.all(w["w_llr"] == 1) and np.all(w["wf_edge"] == 1) and np.all(w["wf_llr"] == 1):
            return None
        dev = self.device
        return dict(w_edge=torch.as_tensor(w["w_edge"]).to(dev), w_llr=torch.as_tensor(w["w_llr"]).to(dev),
                    wf_edge=torch.as_tensor(w["wf_edge"]).to(dev), wf_llr=torch.as_tensor(w["wf_llr"]).to(dev),
                    iterations=int(iterations), stride=int(self.max_dv))

    def _train_args(self, llr, weights):
        if not llr.is_cuda or llr.device != self.device:
            raise ValueError(f"llr must be a CUDA tensor on {self.device} (no CPU fallback)")
        if llr.dim() != 2 or llr.shape[1] != self.n:
            raise ValueError(f"llr must be [B,{self.n}], got {tuple(llr.shape)}")
        iters, stride = int(weights["iterations"]), int(weights["stride"])
        shapes = dict(w_edge=(iters, self.E, stride), w_llr=(iters, self.n), wf_edge=(self.E,), wf_llr=(self.n,))
        w = {}
        for k, shp in shapes.items():
            t = weights[k]
            if tuple(t.shape) != shp or t.device != self.device:
                raise ValueError(f"{k} must be {shp} on {self.device}, got {tuple(t.shape)} on {t.device}")
            w[k] = t.detach().float().contiguous()
        return llr.detach().float().contiguous(), w, iters, stride

    def train_forward(self, llr, weights, clamp_value, x0=None):
        """Weighted sum-product forward that keeps the tape for train_backward: -> (prob [B,n], tape).  prob is
        bit-identical to decode_weighted(update='sp')."""
        llr, w, iters, stride = self._train_args(llr, weights)
        B = llr.shape[0]
        prob = torch.empty((B, self.n), dtype=torch.float32, device=self.device)
        tape = torch.empty(((iters + 1), self.E, B), dtype=torch.float32, device=self.device)
        if x0 is not None:
            x0 = x0.detach().float().contiguous()
        with torch.cuda.device(self.device):
            N.check(N.lib().ldpc_bp_train_forward(
                self._h, _ptr(llr), B, iters, float(clamp_value), _ptr(w["w_edge"]), _ptr(w["w_llr"]), _ptr(w["wf_edge"]), _ptr(w["wf_llr"]),
                stride, _ptr(x0), _ptr(prob), _ptr(tape), ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))
        return prob, tape

    def train_backward(self, llr, weights, clamp_value, tape, grad_prob):
        """-> dict(grad_llr [B,n], w_edge, w_llr, wf_edge, wf_llr): the gradient of sum(prob * grad_prob), weight
        gradients summed over the batch (ldpc_bp_train_backward)."""
        llr, w, iters, stride = self._train_args(llr, weights)
        B = llr.shape[0]
        grad_prob = grad_prob.detach().float().contiguous()
        if tuple(grad_prob.shape) != (B, self.n) or tuple(tape.shape) != (iters + 1, self.E, B):
            raise ValueError("grad_prob must be [B,n] and tape the one train_forward returned for this batch")
        mk = lambda *shape: torch.empty(shape, dtype=torch.float32, device=self.device)
        g = dict(grad_llr=mk(B, self.n), w_edge=mk(iters, self.E, stride), w_llr=mk(iters, self.n), wf_edge=mk(self.E), wf_llr=mk(self.n))
        ws = mk(2 * self.E * max(B, 1))
        with torch.cuda.device(self.device):
            N.check(N.lib().ldpc_bp_train_backward(
                self._h, _ptr(llr), B, iters, float(clamp_value), _ptr(w["w_edge"]), _ptr(w["w_llr"]), _ptr(w["wf_edge"]), _ptr(w["wf_llr"]),
                stride, _ptr(tape), _ptr(grad_prob), _ptr(g["grad_llr"]), _ptr(g["w_edge"]), _ptr(g["w_llr"]), _ptr(g["wf_edge"]),
                _ptr(g["wf_llr"]), _ptr(ws), ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))
        return g

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


def decode_bits_host(code: LdpcCode, llrs: np.ndarray, iterations, clamp_value, out: np.ndarray, update="sp", param=1.0, chunk=16384, threads=0):
    """llrs [N,n] (ordinary numpy memory, float64/float32/float16/int8) -> out [N,n] filled with {0,1} (float64, float32 or
    uint8) through ldpc_decode_bits_host: threaded cast/copy into pinned staging, packed bits back, threaded expansion."""
    llrs = np.ascontiguousarray(llrs)
    if llrs.dtype not in _NP_DTYPES:
        llrs = llrs.astype(np.float32)
    Nn, n = llrs.shape
    if n != code.n or out.shape != (Nn, n) or not out.flags.c_contiguous:
        raise ValueError(f"llrs and out must be C-contiguous [N,{code.n}]")
    kinds = {np.dtype(np.float64): N.F64, np.dtype(np.float32): N.F32, np.dtype(np.uint8): N.I8}
    if out.dtype not in kinds:
        raise ValueError("out must be float64, float32 or uint8")
    if threads <= 0:
        # the cast / expand threads are host-memory-bound and scale to ~16; under torchrun the ranks of a node share its
        # cores, so each takes its share (16 threads x 8 ranks on 32 cores measured SLOWER than one rank alone)
        try:
            cores = len(os.sched_getaffinity(0))
        except AttributeError:
            cores = os.cpu_count() or 1
        ranks = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1))
        threads = max(1, min(16, cores // ranks))
    with torch.cuda.device(code.device):
        N.check(N.lib().ldpc_decode_bits_host(code._h, llrs.ctypes.data, _NP_DTYPES[llrs.dtype], Nn, int(iterations), _update_id(update),
                                              float(clamp_value), float(param), out.ctypes.data, kinds[out.dtype], int(chunk), int(threads)))
    return out
