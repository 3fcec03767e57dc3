"""CPU ORACLE - TEST INFRASTRUCTURE ONLY.  ctypes binding of oracle/ldpc_oracle.c
(the plain-C restatement; see its header for the reference file:line map).  Used by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
UPDATE_IDS = {"sp": 0, "minsum": 1, "nms": 2, "oms": 3}


def build(force=False):
    so = os.path.join(_HERE, "libldpc_oracle.so")
    src = os.path.join(_HERE, "ldpc_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.oracle_decode.restype = ctypes.c_int
        _LIB.oracle_max_threads.restype = ctypes.c_int
    return _LIB


class CGraph:
    """Edge tables in the layout ldpc_oracle.c wants (restates masking.py:85-95)."""

    def __init__(self, H):
        Hb = np.asarray(H) != 0
        self.m, self.n = Hb.shape
        rows, cols = np.nonzero(Hb)
        self.E = int(rows.size)
        order = np.lexsort((rows, cols))            # vm -> cm
        vm_of_cm = np.empty(self.E, np.int64); vm_of_cm[order] = np.arange(self.E)
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        self.chk_ptr = i32(np.concatenate([[0], np.cumsum(Hb.sum(1))]))
        self.var_ptr = i32(np.concatenate([[0], np.cumsum(Hb.sum(0))]))
        self.chk_var = i32(cols)
        self.cm_of_vm = i32(order)
        self.vm_of_cm = i32(vm_of_cm)


def decode(graph: CGraph, llr, iters, clamp, update="sp", param=1.0, x0=None, nthreads=0,
           want=("t", "prob", "hard", "syndrome")):
    llr = np.ascontiguousarray(llr, dtype=np.float32)
    B, n = llr.shape
    assert n == graph.n
    t = np.empty((B, n), np.float32) if "t" in want else None
    prob = np.empty((B, n), np.float32) if "prob" in want else None
    hard = np.empty((B, n), np.uint8)
    synd = np.empty(B, np.int32) if "syndrome" in want else None
    x = np.empty((B, graph.E), np.float32) if "x" in want else None
    if x0 is not None:
        x0 = np.ascontiguousarray(x0, dtype=np.float32)
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p) if a is not None else None
    rc = lib().oracle_decode(graph.m, graph.n, graph.E, P(graph.chk_ptr), P(graph.chk_var),
                             P(graph.vm_of_cm), P(graph.var_ptr), P(graph.cm_of_vm), P(llr),
                             ctypes.c_int64(B), int(iters), UPDATE_IDS[update] if isinstance(update, str) else int(update),
                             ctypes.c_float(clamp), ctypes.c_float(param), P(x0), P(t), P(prob), P(hard),
                             P(synd), P(x), int(nthreads))
    if rc != 0:
        raise RuntimeError(f"oracle_decode failed rc={rc}")
    return dict(t=t, prob=prob, hard=hard, syndrome=synd, x=x)


def max_threads():
    return int(lib().oracle_max_threads())
