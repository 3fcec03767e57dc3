"""CPU ORACLE - TEST INFRASTRUCTURE ONLY (never imported by the product path).

NumPy float32 restatement of the reference's MLP demappers (``pytorch/nn/llr.py``):

* ``LLRestimator_withSNR`` (``nn/llr.py:54-73``): ``[B, 2N+1] -> tanh(hidden1) -> tanh(hidden2) ->
  tanh(hidden3) -> final``, hidden width 16N, input = N real parts, N imaginary parts of the
  (quantized) time-domain OFDM symbol and the linear SNR (``evaluate_quantized_snr.py:135-140``).
* ``LLRestimator`` (``nn/llr.py:7-52``): ``fft_layer`` (bias-free 2N x 2N) then hidden3..5 and final.

``nn.Linear`` is ``x @ W.T + b`` (ATen addmm, fp32).  Summation order inside sgemm is not
specified, so the pin against the reference (tests/golden/nn_demapper.npz, minted by
oracle/make_golden_nn.py from the reference's own checkpoint) is a tolerance pin: 1e-5
relative to the output scale.
"""
from __future__ import annotations

import numpy as np

WITHSNR_LAYERS = ("hidden1", "hidden2", "hidden3", "final")          # nn/llr.py:62-66
PLAIN_LAYERS = ("fft_layer", "hidden3", "hidden4", "hidden5", "final")   # nn/llr.py:46-52 (forward)
PLAIN_ACTS = (False, True, True, True, False)                             # no tanh after fft_layer (nn/llr.py:46-47)


def strip_module_prefix(state):
    """Checkpoints were saved from nn.DataParallel (evaluate_quantized_snr.py:57,68): 'module.' prefix."""
    return {(k[7:] if k.startswith("module.") else k): np.asarray(v, dtype=np.float32) for k, v in state.items()}


def mlp_forward(state, x, layers=WITHSNR_LAYERS, acts=None):
    """fp32 forward; acts[i] -> tanh after layer i (default: every layer except the last, nn/llr.py:68-73)."""
    if acts is None:
        acts = tuple(i + 1 < len(layers) for i in range(len(layers)))
    s = strip_module_prefix(state)
    h = np.asarray(x, dtype=np.float32)
    for i, name in enumerate(layers):
        W = s[name + ".weight"]
        h = h @ W.T
        if name + ".bias" in s:
            h = h + s[name + ".bias"]
        h = h.astype(np.float32)
        if acts[i]:
            h = np.tanh(h).astype(np.float32)
    return h


def nn_input_samples(qrx_signal_rescaled, snr_linear):
    """evaluate_quantized_snr.py:135-140: [N, S] complex time samples (one OFDM symbol per column)
    -> [S, 2N+1] rows (Re[0..N), Im[0..N), snr)."""
    a = np.concatenate((qrx_signal_rescaled.real.T, qrx_signal_rescaled.imag.T), axis=1)
    a = a.reshape(-1, 2 * qrx_signal_rescaled.shape[0])
    return np.concatenate((a, snr_linear * np.ones((a.shape[0], 1))), axis=1)
