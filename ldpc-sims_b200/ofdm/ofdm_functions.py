"""Drop-in for the reference's ofdm/ofdm_functions.py: same function names, positional
arguments and return arities, computed by libldpc_b200.so on the GPU.

    decode_bits(llrs, H, bp_iterations, batch_size, clamp_value)      ofdm_functions.py:131-163

Like the reference module this one re-exports BeliefPropagation (``from bp.bp import *``,
ofdm_functions.py:6).  Keyword-only extras never change positional behaviour.
"""
import hashlib

import numpy as np
import torch  # noqa: F401  (reference scripts rely on `from ofdm.ofdm_functions import *` exporting torch/np/nn)
import torch.nn as nn  # noqa: F401

from bp.bp import *  # noqa: F401,F403
from ldpc_b200 import _native
from ldpc_b200.decoder import LdpcCode, decode_host

_CODE_CACHE = {}


def _code_for(H, qc_Z=0):
    _native.require_cuda()                      # fail loudly: there is no CPU fallback
    Hb = np.ascontiguousarray((np.asarray(H) != 0).astype(np.uint8))
    key = (Hb.shape, hashlib.sha1(Hb.tobytes()).hexdigest(), int(qc_Z), torch.cuda.current_device())
    if key not in _CODE_CACHE:
        _CODE_CACHE[key] = LdpcCode(Hb, qc_Z=qc_Z)
    return _CODE_CACHE[key]


# batch_size must be divisible! (reference comment, ofdm_functions.py:130)
def decode_bits(llrs, H, bp_iterations, batch_size, clamp_value, *, update="sp", param=1.0, qc_Z=0,
                out_dtype=np.float64):
    """llrs [N,n] float64 log(P1/P0) -> [N,n] array of {0,1} (float64 like the reference).

    Reproduces the reference's batching semantics: only the first (N // batch_size) *
    batch_size rows are decoded, the ragged tail stays zero (ofdm_functions.py:133-135).
    One native handle per distinct H is cached instead of rebuilding the model per call.
    """
    llrs = np.asarray(llrs)
    output_bits = np.zeros(llrs.shape, dtype=out_dtype)
    used = (llrs.shape[0] // int(batch_size)) * int(batch_size)
    if used == 0:
        return output_bits
    code = _code_for(H, qc_Z)
    out = decode_host(code, llrs[:used], bp_iterations, clamp_value, update=update, param=param, want=("hard",))
    output_bits[:used] = out["hard"]
    return output_bits


decoder = decode_bits              # old spelling: `from decoder import decoder` (evaluate.py:9,117)
