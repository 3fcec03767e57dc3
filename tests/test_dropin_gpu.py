"""The reference-facing Python API (bp.bp.BeliefPropagation, ofdm.ofdm_functions.decode_bits)
called exactly as the reference's scripts call it, against the reference's golden vectors."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_belief_propagation_module_like_reference(golden_dir):
    from bp.bp import BeliefPropagation
    from bp.parity import H
    g = np.load(os.path.join(golden_dir, "bp_default_code.npz"))
    device = torch.device("cuda")
    for name in ("gauss0", "link4dB", "edge"):
        iters, clamp = int(g[f"{name}_iters"]), float(g[f"{name}_clamp"])
        bp_model = BeliefPropagation(H, iters)
        bp_model.eval()
        bp_model.to(device)
        llr = torch.tensor(g[f"{name}_llr"], dtype=torch.float, device=device)
        x = torch.zeros(llr.shape[0], bp_model.layer_size(), dtype=torch.float, device=device)
        y_est = bp_model(x, llr, clamp)
        assert y_est.shape == llr.shape and y_est.dtype == torch.float32 and y_est.is_cuda
        bits = np.round(y_est.cpu().detach().numpy())
        assert np.array_equal(np.packbits(bits.astype(np.uint8), axis=1), g[f"{name}_hard"])
        dp = np.abs(y_est[:64].cpu().numpy() - g[f"{name}_prob"])
        assert np.mean(dp <= 1e-5) >= 0.999 and dp.max() <= 5e-3
    assert bp_model.layer_size() == 96


def test_decode_bits_like_reference(golden_dir):
    from bp.parity import H
    from ofdm.ofdm_functions import decode_bits
    g = np.load(os.path.join(golden_dir, "bp_default_code.npz"))
    for snr in (0, 4, 8, 12):
        llrs = g[f"link{snr}dB_llr"].astype(np.float64)
        bits = decode_bits(llrs, H, 3, 256, 20)
        assert bits.dtype == np.float64 and bits.shape == llrs.shape
        assert np.array_equal(np.packbits(bits.astype(np.uint8), axis=1), g[f"link{snr}dB_hard"])
    rag = decode_bits(g["link4dB_llr"].astype(np.float64), H, 3, 300, 20)
    assert np.array_equal(np.packbits(rag.astype(np.uint8), axis=1), g["link4dB_ragged300"])
    assert not rag[300:].any()
    assert not decode_bits(np.ones((5, 64)), H, 3, 8, 20).any()          # batch larger than N: nothing decoded


def test_legacy_constructor_and_cpu_tensors():
    from bp.bp import BeliefPropagation
    from bp.masking import genMasks
    from bp.parity import H
    mask_c, mask_v, mask_v_final, llr_expander = genMasks(H)
    m5 = BeliefPropagation(mask_v, mask_c, mask_v_final, llr_expander, 3)
    m2 = BeliefPropagation(H, 3)
    llr = torch.randn(32, 64) * 3
    x = torch.zeros(32, 96)
    a, b = m5(x, llr, 20), m2(x, llr, 20)
    assert not a.is_cuda and torch.equal(a, b)
