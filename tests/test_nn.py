"""MLP demapper (reference nn/llr.py:54-73, evaluate_quantized_snr.py:135-173).

CPU: the numpy oracle against the golden vectors minted from the reference's own checkpoint
(tests/golden/nn_demapper.npz, oracle/make_golden_nn.py).  GPU: the native tensor-core kernel
through the drop-in nn.llr module against the same vectors, then through the decoder.
Tolerance (fp32 GEMM chains do not pin a summation order): 1e-5 of the output scale for the
fp32-equivalent mode (3 bf16 planes), 2e-4 for 2 planes.
"""
import ctypes
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = np.load(os.path.join(ROOT, "tests", "golden", "nn_demapper.npz"))
STATE = {k[2:]: GOLD[k] for k in GOLD.files if k.startswith("w_")}


def _scale_err(a, b):
    return float(np.max(np.abs(a.astype(np.float64) - b)) / np.max(np.abs(b)))


def test_oracle_matches_reference_golden():
    import nn_oracle as NO
    for tag in GOLD["names"]:
        y = NO.mlp_forward(STATE, GOLD[f"{tag}_x"])
        assert y.dtype == np.float32 and y.shape == GOLD[f"{tag}_llr"].shape
        assert _scale_err(y, GOLD[f"{tag}_llr"]) < 1e-5, tag


def test_golden_decode_chain_with_oracles():
    """reference LLR estimates -> BP oracle reproduces the reference's decoded bits (decode_bits)."""
    import bp_oracle as O
    from ldpc_b200.codes import peg_64_32
    H, _ = peg_64_32()
    for tag in GOLD["names"]:
        snrdb, iters, clamp = GOLD[f"{tag}_meta"]
        bits = O.decode_bits(GOLD[f"{tag}_llr"].astype(np.float64), H, int(iters), 256, float(clamp))
        assert np.array_equal(np.packbits(bits.astype(np.uint8), axis=1), GOLD[f"{tag}_bits"]), tag


def test_mlp_abi_validation_without_gpu():
    from ldpc_b200 import _native as N
    L = N.lib()
    h = ctypes.c_void_p()
    assert L.ldpc_mlp_create(0, None, None, None, None, 3, 0, ctypes.byref(h)) == N.EINVAL
    dims = (ctypes.c_int32 * 2)(65, 512)
    w = np.zeros((512, 65), np.float32)
    wa = (ctypes.c_void_p * 1)(w.ctypes.data)
    assert L.ldpc_mlp_create(1, dims, wa, None, None, 7, 0, ctypes.byref(h)) == N.EINVAL
    if L.ldpc_device_count() == 0:                       # no CPU fallback
        assert L.ldpc_mlp_create(1, dims, wa, None, None, 3, 0, ctypes.byref(h)) == N.ECUDA
    assert L.ldpc_mlp_forward(None, None, 1, None, None) == N.EINVAL


def _model(splits):
    import torch
    from nn.llr import LLRestimator_withSNR
    m = torch.nn.DataParallel(LLRestimator_withSNR(32, splits=splits))      # evaluate_quantized_snr.py:57
    m.load_state_dict({k: torch.tensor(v) for k, v in STATE.items()})
    return m.eval().cuda()


@pytest.mark.gpu
@pytest.mark.parametrize("splits,tol", [(3, 1e-5), (2, 1e-4)])
def test_native_mlp_matches_reference_golden(splits, tol):
    import torch
    m = _model(splits)
    for tag in GOLD["names"]:
        x = torch.tensor(GOLD[f"{tag}_x"], dtype=torch.float, device="cuda")
        y = m(x).cpu().numpy()
        assert _scale_err(y, GOLD[f"{tag}_llr"]) < tol, (tag, _scale_err(y, GOLD[f"{tag}_llr"]))


@pytest.mark.gpu
def test_native_mlp_matches_oracle_ragged_and_chunked():
    """Row counts that are not multiples of the 128-row tile / of the chunk; CPU input tensor."""
    import torch
    import nn_oracle as NO
    from ldpc_b200.mlp import NativeMLP
    rng = np.random.RandomState(3)
    names = ("hidden1", "hidden2", "hidden3", "final")
    net = NativeMLP([STATE[f"module.{n}.weight"] for n in names], [STATE[f"module.{n}.bias"] for n in names], chunk_rows=256)
    for B in (1, 127, 129, 700):
        x = np.concatenate([rng.randn(B, 64).astype(np.float32) * 0.7, np.full((B, 1), 31.6, np.float32)], axis=1)
        y = net(torch.tensor(x).cuda()).cpu().numpy()
        ref = NO.mlp_forward(STATE, x)
        assert y.shape == ref.shape
        assert _scale_err(y, ref) < 1e-5, B
    m = _model(3)
    x = torch.tensor(GOLD["snr15_x"][:100])
    assert _scale_err(m(x).numpy(), GOLD["snr15_llr"][:100]) < 1e-5 and not m(x).is_cuda


@pytest.mark.gpu
def test_nn_demapper_then_decoder_matches_reference_bits():
    """evaluate_quantized_snr.py:150-173: llr_est = LLRest(x); bits_nn = decode_bits(llr_est, H, ...)."""
    import torch
    from ofdm.ofdm_functions import decode_bits
    from bp.parity import H
    m = _model(3)
    for tag in GOLD["names"]:
        snrdb, iters, clamp = GOLD[f"{tag}_meta"]
        x = torch.tensor(GOLD[f"{tag}_x"], dtype=torch.float, device="cuda")
        llr_est = m(x).cpu().numpy().astype(np.float64)
        bits = decode_bits(llr_est, H, int(iters), 256, float(clamp))
        ref = np.unpackbits(GOLD[f"{tag}_bits"], axis=1)[:, :64]
        # identical LLRs up to 1e-5 of scale: at most a handful of marginal bits may differ
        assert np.mean(bits != ref) < 2e-4, (tag, np.mean(bits != ref))


@pytest.mark.gpu
def test_llrestimator_plain_chain_matches_oracle():
    """LLRestimator (fft_layer without bias or tanh, then hidden3..5, final; nn/llr.py:46-52)."""
    import torch
    import nn_oracle as NO
    from nn.llr import LLRestimator
    torch.manual_seed(0)
    m = LLRestimator(32, 10.0).eval()
    x = torch.randn(300, 64)
    y = m(x.cuda()).cpu().numpy()
    state = {k: v.detach().numpy() for k, v in m.state_dict().items()}
    ref = NO.mlp_forward(state, x.numpy(), layers=NO.PLAIN_LAYERS)
    assert _scale_err(y, ref) < 1e-5
