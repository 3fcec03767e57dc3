"""Drop-in for the reference's ofdm/ofdm_functions.py: same function names, positional
arguments and return arities, computed by libldpc_b200.so on the GPU.

    create_bits / encode_bits / modulate_bits                         ofdm_functions.py:8-22
    transmit_symbols / quantizer / demodulate_signal                  ofdm_functions.py:25-78
    weighted_mse / compute_ber / DFT / DFTreal                        ofdm_functions.py:80-107
    gen_data / gen_qdata                                              ofdm_functions.py:109-128
    decode_bits(llrs, H, bp_iterations, batch_size, clamp_value)      ofdm_functions.py:131-163

Random draws (create_bits, the AWGN in transmit_symbols) stay on numpy's global generator in
the reference's order, so seeded reference experiments reproduce; encode, modulate, the OFDM
transforms, quantizer, demapper and the decoder run in libldpc_b200.so kernels.

Like the reference module this one re-exports BeliefPropagation (``from bp.bp import *``,
ofdm_functions.py:6).  Keyword-only extras never change positional behaviour.
"""
import hashlib

import numpy as np
import torch  # noqa: F401  (reference scripts rely on `from ofdm.ofdm_functions import *` exporting torch/np/nn)
import torch.nn as nn  # noqa: F401

from bp.bp import *  # noqa: F401,F403
from ldpc_b200 import _native
from ldpc_b200.decoder import LdpcCode, decode_bits_host, decode_host
from ldpc_b200.linksim import encode_bits, modulate_bits, transmit_symbols, quantizer, demodulate_signal  # noqa: F401

_CODE_CACHE = {}


def _code_for(H, qc_Z="auto"):
    _native.require_cuda()                      # fail loudly: there is no CPU fallback
    Hb = np.ascontiguousarray((np.asarray(H) != 0).astype(np.uint8))
    key = (Hb.shape, hashlib.sha1(Hb.tobytes()).hexdigest(), qc_Z, torch.cuda.current_device())
    if key not in _CODE_CACHE:
        _CODE_CACHE[key] = LdpcCode(Hb, qc_Z=qc_Z)
    return _CODE_CACHE[key]


# batch_size must be divisible! (reference comment, ofdm_functions.py:130)
def decode_bits(llrs, H, bp_iterations, batch_size, clamp_value, *, update="sp", param=1.0, qc_Z="auto",
                out_dtype=np.float64):
    """llrs [N,n] float64 log(P1/P0) -> [N,n] array of {0,1} (float64 like the reference).

    Reproduces the reference's batching semantics: only the first (N // batch_size) *
    batch_size rows are decoded, the ragged tail stays zero (ofdm_functions.py:133-135).
    One native handle per distinct H is cached instead of rebuilding the model per call.
    """
    llrs = np.asarray(llrs)
    used = (llrs.shape[0] // int(batch_size)) * int(batch_size)
    if used == 0:
        return np.zeros(llrs.shape, dtype=out_dtype)
    code = _code_for(H, qc_Z)
    if np.dtype(out_dtype) in (np.dtype(np.float64), np.dtype(np.float32), np.dtype(np.uint8)):
        output_bits = np.empty(llrs.shape, dtype=out_dtype)          # every decoded row is written by the native call
        output_bits[used:] = 0
        decode_bits_host(code, llrs[:used], bp_iterations, clamp_value, output_bits[:used], update=update, param=param)
        return output_bits
    output_bits = np.zeros(llrs.shape, dtype=out_dtype)
    out = decode_host(code, llrs[:used], bp_iterations, clamp_value, update=update, param=param, want=("hard",))
    output_bits[:used] = out["hard"]
    return output_bits


def create_bits(num_bits):                                   # ofdm_functions.py:8-9
    return np.random.randint(2, size=num_bits).reshape((1, -1))


def weighted_mse(llr_est, llr, epsilon):                     # ofdm_functions.py:80-81
    return torch.mean((llr_est - llr) ** 2 / (torch.abs(llr) + epsilon))


def compute_ber(bits_est, bits):                             # ofdm_functions.py:83-84
    return np.sum(np.abs(bits_est - bits)) / bits.size


def DFT(N):                                                  # ofdm_functions.py:86-93 (unitary)
    x = np.arange(N).reshape(-1, 1)
    y = np.arange(N).reshape(1, -1)
    return np.exp(-1j * 2 * np.pi * x * y / N) / np.sqrt(N)


def DFTreal(N):                                              # ofdm_functions.py:95-107
    W = DFT(N)
    Wr = np.zeros((2 * N, 2 * N), dtype=float)
    Wr[0::2, 0::2] = W.real
    Wr[0::2, 1::2] = -W.imag
    Wr[1::2, 0::2] = W.imag
    Wr[1::2, 1::2] = W.real
    return Wr


def gen_data(tx_symbols, snrdb, ofdm_size):                  # ofdm_functions.py:109-116
    """-> (rx_signal, rx_symbols, rx_llrs, tx_signal), the reference's 4-tuple."""
    snr = np.power(10, snrdb / 10)
    rx_signal, tx_signal = transmit_symbols(tx_symbols, ofdm_size, snr)
    rx_llrs, rx_symbols = demodulate_signal(rx_signal, ofdm_size, snr)
    return rx_signal, rx_symbols, rx_llrs, tx_signal


def gen_qdata(rx_signal, snrdb, qbits, clip_ratio, ofdm_size):   # ofdm_functions.py:118-128
    """-> (qrx_signal, qrx_symbols, qrx_llrs); AGC clip = std(rx_signal) * clip_ratio over the
    whole array (a data-dependent global statistic)."""
    snr = np.power(10, snrdb / 10)
    sigma_rx = np.max(np.std(rx_signal))
    agc_clip = sigma_rx * clip_ratio
    qrx_signal = quantizer(rx_signal, qbits, agc_clip)
    qrx_llrs, qrx_symbols = demodulate_signal(qrx_signal, ofdm_size, snr)
    return qrx_signal, qrx_symbols, qrx_llrs


decoder = decode_bits              # old spelling: `from decoder import decoder` (evaluate.py:9,117)
