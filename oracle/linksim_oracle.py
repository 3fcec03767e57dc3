"""CPU ORACLE - TEST INFRASTRUCTURE ONLY (never imported by the product path).

NumPy-2-safe restatement of the reference's float64 link simulator
(``pytorch/ofdm/ofdm_functions.py:8-128``) and of the AGC-scaled quantizer front end and
error metrics the evaluate scripts inline (``pytorch/evaluate_quantized_snr.py:96-133``
and ``:169-188``).  Same numpy operations in the same order, so results are bit-identical
to the reference run with the ``np.complex = complex; np.float = float`` shim
(oracle/make_golden.py checks this and writes tests/golden/frontend.npz).

Random draws use numpy's global legacy generator in the reference's order, so seeding
``np.random.seed`` reproduces the reference stream; ``transmit_symbols`` also accepts an
explicit ``noise`` array so GPU kernels can be checked on identical inputs.
"""
from __future__ import annotations

import functools
import numpy as np


def create_bits(num_bits):                                   # ofdm_functions.py:8-9
    return np.random.randint(2, size=num_bits).reshape((1, -1))


def encode_bits(bits, generator_matrix):                     # ofdm_functions.py:11-15
    bits = bits.reshape((-1, generator_matrix.shape[1])).T
    cbits = np.mod(np.matmul(generator_matrix, bits), 2)
    return cbits.T.reshape((1, -1))


def modulate_bits(bits):                                     # ofdm_functions.py:17-22
    b = -2 * bits.reshape((-1, 2)) + 1
    symbols = (1 / np.sqrt(2)) * b[:, 0] + (1j / np.sqrt(2)) * b[:, 1]
    return symbols.reshape((1, -1))


@functools.lru_cache(maxsize=None)
def _dft_cached(N):
    W = np.zeros((N, N), dtype=complex)
    for x in range(N):                                       # ofdm_functions.py:86-93
        for y in range(N):
            W[x, y] = np.exp(-1j * 2 * np.pi * x * y / N) / np.sqrt(N)
    W.setflags(write=False)
    return W


def DFT(N):
    return _dft_cached(int(N)).copy()


def DFTreal(N):                                              # ofdm_functions.py:95-107
    W = _dft_cached(int(N))
    Wr = np.zeros((2 * N, 2 * N), dtype=float)
    Wr[0::2, 0::2] = W.real
    Wr[0::2, 1::2] = -W.imag
    Wr[1::2, 0::2] = W.imag
    Wr[1::2, 1::2] = W.real
    return Wr


def transmit_symbols(symbols, ofdm_size, snr, noise=None):   # ofdm_functions.py:25-35
    symbols = symbols.reshape((-1, ofdm_size)).T
    ofdm_symbols = np.matmul(_dft_cached(int(ofdm_size)).conj().T, symbols)
    if noise is None:
        noise = (np.random.normal(0, 1 / np.sqrt(snr), ofdm_symbols.shape) +
                 1j * np.random.normal(0, 1 / np.sqrt(snr), ofdm_symbols.shape)) / np.sqrt(2)
    received = ofdm_symbols + noise
    return received.T.reshape((1, -1)), ofdm_symbols.T.reshape((1, -1))


def quantizer(inputs, num_bits, clip_value):                 # ofdm_functions.py:37-51
    """Mid-tread quantizer INCLUDING the reference's clip quirk: the +-1 in the clip
    bounds is in signal units, not index units."""
    num_levels = np.power(2, num_bits)
    step = 2 * clip_value / (num_levels - 1)
    idx_real = np.floor(inputs.real / step + .5)
    idx_imag = np.floor(inputs.imag / step + .5)
    lo, hi = -(num_levels / 2) * step + 1, (num_levels / 2) * step - 1
    out = np.zeros(inputs.shape, dtype=complex)
    out.real = np.clip(step * idx_real, lo, hi)
    out.imag = np.clip(step * idx_imag, lo, hi)
    return out


def qpsk_llrs(received_symbols, snr_est):                    # ofdm_functions.py:69-76
    noise_power = .5 * (1 / snr_est)
    a = 1 / np.sqrt(2)
    llr_bit0 = (np.power(received_symbols.real - a, 2) - np.power(received_symbols.real + a, 2)) / (2 * noise_power)
    llr_bit1 = (np.power(received_symbols.imag - a, 2) - np.power(received_symbols.imag + a, 2)) / (2 * noise_power)
    llrs = np.concatenate((llr_bit0.T.reshape((-1, 1)), llr_bit1.T.reshape((-1, 1))), axis=1)
    return llrs.reshape((1, -1))


def demodulate_signal(symbols, ofdm_size, snr_est):          # ofdm_functions.py:63-78
    symbols = symbols.reshape((-1, ofdm_size)).T
    received_symbols = np.matmul(_dft_cached(int(ofdm_size)), symbols)
    return qpsk_llrs(received_symbols, snr_est), received_symbols.T.reshape((1, -1))


def compute_ber(bits_est, bits):                             # ofdm_functions.py:83-84
    return np.sum(np.abs(bits_est - bits)) / bits.size


def gen_data(tx_symbols, snrdb, ofdm_size, noise=None):      # ofdm_functions.py:109-116
    snr = np.power(10, snrdb / 10)
    rx_signal, tx_signal = transmit_symbols(tx_symbols, ofdm_size, snr, noise=noise)
    rx_llrs, rx_symbols = demodulate_signal(rx_signal, ofdm_size, snr)
    return rx_signal, rx_symbols, rx_llrs, tx_signal


def gen_qdata(rx_signal, snrdb, qbits, clip_ratio, ofdm_size):   # ofdm_functions.py:118-128
    snr = np.power(10, snrdb / 10)
    sigma_rx = np.max(np.std(rx_signal))
    agc_clip = sigma_rx * clip_ratio
    qrx_signal = quantizer(rx_signal, qbits, agc_clip)
    qrx_llrs, qrx_symbols = demodulate_signal(qrx_signal, ofdm_size, snr)
    return qrx_signal, qrx_symbols, qrx_llrs


def agc_quantized_frontend(rx_signal, snrdb_val, qbits, clip_ratio, ofdm_size, agc_clip=10):
    """evaluate_quantized_snr.py:96-133 - scale to a fixed clip level with the script's
    'sigma_rx' (a variance used as an amplitude, reference behaviour), quantize, scale
    back, de-OFDM, LLR.  Returns (qrx_llrs [1,2L], qrx_signal_rescaled [ofdm, L/ofdm])."""
    snr_single = np.power(10, snrdb_val / 10)
    received_symbols = rx_signal.reshape((-1, ofdm_size)).T
    sigma_rx = .5 * (1 + 1 / snr_single)
    factor = agc_clip / sigma_rx * clip_ratio
    rx_signal_scaled = (factor * received_symbols).T.reshape((1, -1))
    qrx_signal = quantizer(rx_signal_scaled, qbits, agc_clip)
    qrx_signal_rescaled = qrx_signal.reshape((-1, ofdm_size)).T / factor
    deofdm = np.matmul(_dft_cached(int(ofdm_size)), qrx_signal_rescaled)
    return qpsk_llrs(deofdm, snr_single), qrx_signal_rescaled


def framed_link_llrs(enc_bits, snrdb, ofdm_size, qbits=0, clip_ratio=1.0, agc_clip=10, noise=None, channel="awgn", compander=False,
                     fading=None):
    """The reference chain for a code whose length is NOT 2 * ofdm_size (the reference itself only works for
    n = 2 * ofdm_size, SURVEY hard part 6): every codeword is framed into ceil((n/2)/N) OFDM symbols, null
    subcarriers after its last QPSK symbol; each OFDM symbol then goes through exactly the reference's per-symbol
    steps (ofdm_functions.py:17-35,63-78; with qbits > 0 the AGC-scaled quantizer of
    evaluate_quantized_snr.py:96-133).  With n = 2 * ofdm_size this IS gen_data / agc_quantized_frontend.
    channel='rayleigh' (not in the reference; defined here as in include/ldpc_b200.h): one CN(0,1) gain h per OFDM symbol,
    r = h x + n, the receiver divides by h and demaps with the noise power 0.5 / (snr |h|^2).  compander: the ADC input
    goes through agc_clip * tanh(. / agc_clip) first.
    enc_bits [B, n] of 0/1; returns (llrs float64 [B, n], rx_signal [1, B*S*N])."""
    enc_bits = np.asarray(enc_bits)
    B, n = enc_bits.shape
    nsym = n // 2
    S = (nsym + ofdm_size - 1) // ofdm_size
    sym = np.zeros((B, S * ofdm_size), dtype=complex)
    sym[:, :nsym] = modulate_bits(enc_bits.reshape((1, -1))).reshape(B, nsym)
    snr = np.power(10, snrdb / 10)
    if channel == "awgn" and not compander:
        rx_signal, _ = transmit_symbols(sym.reshape((1, -1)), ofdm_size, snr, noise=noise)
        if qbits > 0:
            llr, _ = agc_quantized_frontend(rx_signal, snrdb, qbits, clip_ratio, ofdm_size, agc_clip=agc_clip)
        else:
            llr, _ = demodulate_signal(rx_signal, ofdm_size, snr)
        return llr.reshape(B, S * ofdm_size * 2)[:, :n].copy(), rx_signal
    W = _dft_cached(int(ofdm_size))
    cols = sym.reshape((-1, ofdm_size)).T                          # [N, B*S]: columns = OFDM symbols
    x = np.matmul(W.conj().T, cols)
    if channel == "rayleigh":
        h = fading if fading is not None else (np.random.normal(0, 1, cols.shape[1]) + 1j * np.random.normal(0, 1, cols.shape[1])) / np.sqrt(2)
    elif channel == "awgn":
        h = np.ones(cols.shape[1], dtype=complex)
    else:
        raise ValueError(channel)
    if noise is None:
        noise = (np.random.normal(0, 1 / np.sqrt(snr), x.shape) + 1j * np.random.normal(0, 1 / np.sqrt(snr), x.shape)) / np.sqrt(2)
    r = x * h[None, :] + noise
    if qbits > 0:
        factor = agc_clip / (.5 * (1 + 1 / snr)) * clip_ratio
        v = factor * r
        if compander:
            v = agc_clip * np.tanh(v.real / agc_clip) + 1j * agc_clip * np.tanh(v.imag / agc_clip)
        r = quantizer(v, qbits, agc_clip) / factor
    elif compander:
        raise ValueError("the compander belongs to the ADC model: qbits > 0")
    z = r * np.conj(h)[None, :] / (np.abs(h) ** 2)[None, :]
    R = np.matmul(W, z)
    llr = qpsk_llrs(R, snr * (np.abs(h) ** 2)[None, :])
    return llr.reshape(B, S * ofdm_size * 2)[:, :n].copy(), r.T.reshape((1, -1))


def error_metrics(llrs, decoded_bits, enc_bits, k):
    """evaluate_quantized_snr.py:169-188 as exact integer counters.
    llrs, decoded_bits, enc_bits: [N,n].  Returns dict of int64 counts:
    uncoded_errs (all n bits, llr==0 -> bit 0), info_errs (first k columns),
    frame_errs (any of n bits wrong), bits = N*n, info_bits = N*k, frames = N."""
    llrs = np.asarray(llrs); decoded_bits = np.asarray(decoded_bits); enc_bits = np.asarray(enc_bits)
    cbits = (np.sign(llrs) + 1) // 2
    N, n = enc_bits.shape
    return dict(
        uncoded_errs=int(np.sum(np.abs(cbits - enc_bits))),
        info_errs=int(np.sum(np.abs(decoded_bits[:, 0:k] - enc_bits[:, 0:k]))),
        frame_errs=int(np.sum(np.sign(np.sum(np.abs(decoded_bits - enc_bits), axis=1)))),
        bits=int(N * n), info_bits=int(N * k), frames=int(N))
