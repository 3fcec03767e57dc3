"""ctypes binding of libldpc_b200.so (the C ABI in include/ldpc_b200.h).

The library is built in-tree by ``python __graft_entry__.py`` / ``make -C ldpc-sims_b200/csrc``.
There is deliberately NO fallback: if the shared object is missing, or a compute call is
made without a CUDA device, this raises - the product path never routes through a CPU
implementation.
"""
from __future__ import annotations

import ctypes
import os

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(_PKG, "lib", "libldpc_b200.so")

OK, EINVAL, ECUDA, ENOMEM, EUNSUPPORTED = 0, -1, -2, -3, -4
UPDATE_SP, UPDATE_MINSUM, UPDATE_NMS, UPDATE_OMS = 0, 1, 2, 3
F32, F64, F16, I8 = 0, 1, 2, 3
KERNEL_GENERIC, KERNEL_QC, KERNEL_TINY, KERNEL_QC_RT, KERNEL_QC_TMA = 0, 1, 2, 3, 4
PREC_F32, PREC_F16X2 = 0, 1
ABI_VERSION = 1

UPDATE_IDS = {"sp": 0, "tanh": 0, "sum-product": 0, "sumproduct": 0,
              "minsum": 1, "min-sum": 1, "ms": 1, "nms": 2, "oms": 3}


class LdpcError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libldpc_b200 error {code}: {msg}")
        self.code = code


class DecodeParams(ctypes.Structure):
    _fields_ = [("struct_size", ctypes.c_int32), ("llr_dtype", ctypes.c_int32), ("llr", ctypes.c_void_p),
                ("B", ctypes.c_int64), ("iters", ctypes.c_int32), ("update", ctypes.c_int32),
                ("clamp_value", ctypes.c_float), ("param", ctypes.c_float), ("x0", ctypes.c_void_p),
                ("prob", ctypes.c_void_p), ("llr_post", ctypes.c_void_p), ("hard", ctypes.c_void_p),
                ("hard_packed", ctypes.c_void_p), ("syndrome", ctypes.c_void_p), ("x_out", ctypes.c_void_p),
                ("early_exit", ctypes.c_int32), ("reserved", ctypes.c_int32), ("iters_used", ctypes.c_void_p)]


class CodeInfo(ctypes.Structure):
    _fields_ = [(k, ctypes.c_int32) for k in ("m", "n", "E", "max_dc", "max_dv", "kernel", "qc_Z", "reserved")]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build the CUDA extension first "
            "(python -c 'import __graft_entry__ as g; g.build()' or make -C ldpc-sims_b200/csrc). "
            "There is no CPU fallback.")
    L = ctypes.CDLL(LIB_PATH)
    vp, i32, i64, f32 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float
    L.ldpc_abi_version.restype = ctypes.c_int
    L.ldpc_last_error.restype = ctypes.c_char_p
    L.ldpc_device_count.restype = ctypes.c_int
    L.ldpc_code_create.restype = ctypes.c_int
    L.ldpc_code_create.argtypes = [vp, vp, i32, i32, i32, vp, ctypes.POINTER(vp)]
    L.ldpc_code_destroy.restype = None
    L.ldpc_code_destroy.argtypes = [vp]
    L.ldpc_code_info.restype = ctypes.c_int
    L.ldpc_code_info.argtypes = [vp, ctypes.POINTER(CodeInfo)]
    L.ldpc_code_plan_info.restype = ctypes.c_int
    L.ldpc_code_plan_info.argtypes = [vp, vp]
    L.ldpc_qc_register_plugin.restype = ctypes.c_int
    L.ldpc_qc_register_plugin.argtypes = [ctypes.c_char_p]
    L.ldpc_code_set_precision.restype = ctypes.c_int
    L.ldpc_code_set_precision.argtypes = [vp, i32]
    L.ldpc_code_set_kernel.restype = ctypes.c_int
    L.ldpc_code_set_kernel.argtypes = [vp, i32]
    L.ldpc_decode.restype = ctypes.c_int
    L.ldpc_decode.argtypes = [vp, vp, i32, i64, i32, i32, f32, f32, vp, vp, vp, vp, vp, vp, vp, vp]
    L.ldpc_decode_weighted.restype = ctypes.c_int
    L.ldpc_decode_weighted.argtypes = [vp, vp, i32, i64, i32, i32, f32, f32, vp, vp, vp, vp, i32, vp, vp, vp, vp, vp, vp, vp]
    L.ldpc_bp_train_forward.restype = ctypes.c_int
    L.ldpc_bp_train_forward.argtypes = [vp, vp, i64, i32, f32, vp, vp, vp, vp, i32, vp, vp, vp, vp]
    L.ldpc_bp_train_backward.restype = ctypes.c_int
    L.ldpc_bp_train_backward.argtypes = [vp, vp, i64, i32, f32, vp, vp, vp, vp, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.ldpc_decode_ex.restype = ctypes.c_int
    L.ldpc_decode_ex.argtypes = [vp, ctypes.POINTER(DecodeParams), vp]
    L.ldpc_decode_host.restype = ctypes.c_int
    L.ldpc_decode_host.argtypes = [vp, vp, i32, i64, i32, i32, f32, f32, vp, vp, vp, vp, i64]
    L.ldpc_decode_bits_host.restype = ctypes.c_int
    L.ldpc_decode_bits_host.argtypes = [vp, vp, i32, i64, i32, i32, f32, f32, vp, i32, i64, i32]
    L.ldpc_count_errors.restype = ctypes.c_int
    L.ldpc_count_errors.argtypes = [vp, i32, vp, vp, i64, i32, i32, vp, vp]
    u64, sz = ctypes.c_uint64, ctypes.c_size_t
    f64 = ctypes.c_double
    L.ldpc_encode_bits.restype = ctypes.c_int
    L.ldpc_encode_bits.argtypes = [vp, vp, i32, i32, i64, vp, vp]
    L.ldpc_modulate_bits.restype = ctypes.c_int
    L.ldpc_modulate_bits.argtypes = [vp, i64, i32, vp, vp]
    L.ldpc_ofdm_transmit.restype = ctypes.c_int
    L.ldpc_ofdm_transmit.argtypes = [vp, i64, i32, i32, vp, f64, u64, vp, vp, vp]
    L.ldpc_quantize.restype = ctypes.c_int
    L.ldpc_quantize.argtypes = [vp, i64, i32, f64, f64, vp, vp]
    L.ldpc_ofdm_demodulate.restype = ctypes.c_int
    L.ldpc_ofdm_demodulate.argtypes = [vp, i64, i32, i32, f64, vp, vp, vp]
    L.ldpc_code_set_generator.restype = ctypes.c_int
    L.ldpc_code_set_generator.argtypes = [vp, vp, i32]
    L.ldpc_sim_run.restype = ctypes.c_int
    L.ldpc_sim_run.argtypes = [vp, vp, vp, sz, vp, vp]
    L.ldpc_sim_generate.restype = ctypes.c_int
    L.ldpc_sim_generate.argtypes = [vp, vp, vp, vp, vp]
    L.ldpc_sim_generate_ex.restype = ctypes.c_int
    L.ldpc_sim_generate_ex.argtypes = [vp, vp, vp, vp, vp, vp]
    L.ldpc_sim_frontend.restype = ctypes.c_int
    L.ldpc_sim_frontend.argtypes = [vp, vp, vp, vp, vp, vp, vp]
    L.ldpc_decode_count.restype = ctypes.c_int
    L.ldpc_decode_count.argtypes = [vp, vp, i32, i64, i32, i32, f32, f32, vp, i32, vp, vp]
    L.ldpc_mlp_create.restype = ctypes.c_int
    L.ldpc_mlp_create.argtypes = [i32, vp, vp, vp, vp, i32, i64, ctypes.POINTER(vp)]
    L.ldpc_mlp_forward.restype = ctypes.c_int
    L.ldpc_mlp_forward.argtypes = [vp, vp, i64, vp, vp]
    L.ldpc_mlp_destroy.restype = None
    L.ldpc_mlp_destroy.argtypes = [vp]
    L.ldpc_mlp_set_mode.restype = ctypes.c_int
    L.ldpc_mlp_set_mode.argtypes = [vp, i32]
    if L.ldpc_abi_version() != ABI_VERSION:
        raise ImportError(f"libldpc_b200.so ABI {L.ldpc_abi_version()} != binding {ABI_VERSION}; rebuild")
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise LdpcError(rc, lib().ldpc_last_error().decode("utf-8", "replace"))


def require_cuda():
    if lib().ldpc_device_count() <= 0:
        raise LdpcError(ECUDA, "no CUDA device visible - libldpc_b200 has no CPU fallback")
