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


@pytest.mark.gpu
def test_decode_bits_staged_pipeline_formats():
    """ldpc_decode_bits_host (behind decode_bits): float64 / float32 / int8 LLR arrays in ordinary memory, float64 /
    float32 / uint8 {0,1} out, several chunks, ragged tail left zero - all equal to the device path."""
    import torch
    from ldpc_b200.codes import ieee80211n_1944_r12
    from ldpc_b200.decoder import LdpcCode, decode_bits_host
    from ofdm.ofdm_functions import decode_bits
    qc = ieee80211n_1944_r12()
    rng = np.random.RandomState(4)
    N = 300
    llr = rng.randn(N, qc.n) * 2.5 + 2.0
    code = LdpcCode(qc.H, qc_Z=81, qc_proto=qc.proto)
    ref = code.decode(torch.as_tensor(llr.astype(np.float32)).cuda(), 5, 20.0, update="minsum", want=("hard",))["hard"].cpu().numpy()
    for in_dt in (np.float64, np.float32):
        for out_dt in (np.float64, np.float32, np.uint8):
            out = np.full((N, qc.n), 7, out_dt)
            decode_bits_host(code, llr.astype(in_dt), 5, 20.0, out, update="minsum", chunk=64, threads=3)
            assert np.array_equal(out, ref.astype(out_dt)), (in_dt, out_dt)
    q = np.clip(np.round(llr * 4), -127, 127).astype(np.int8)
    refq = code.decode(torch.as_tensor(q).cuda(), 5, 20.0, update="minsum", want=("hard",))["hard"].cpu().numpy()
    out = np.empty((N, qc.n), np.uint8)
    assert np.array_equal(decode_bits_host(code, q, 5, 20.0, out, update="minsum", chunk=128), refq)
    got = decode_bits(llr, qc.H, 5, 128, 20.0, update="minsum", qc_Z=81)           # 300 // 128 * 128 = 256 rows decoded
    assert got.dtype == np.float64 and np.array_equal(got[:256], ref[:256].astype(np.float64)) and not got[256:].any()
    with pytest.raises(ValueError):
        decode_bits_host(code, llr, 5, 20.0, np.empty((N, qc.n), np.int32))


@pytest.mark.gpu
def test_bare_H_selects_the_qc_kernels():
    """BeliefPropagation(H, iterations) / decode_bits(llrs, H, ...) carry no structure hint in the reference: the n=1944
    code must still land on the compiled QC kernel, a quasi-cyclic H without a specialisation on the run-time QC kernel."""
    from bp.bp import BeliefPropagation
    from ldpc_b200.codes import expand_qc, ieee80211n_1944_r12
    from ldpc_b200.decoder import LdpcCode
    qc = ieee80211n_1944_r12()
    m = BeliefPropagation(qc.H, 5).eval()
    llr = torch.randn(8, qc.n, device="cuda") * 3 + 2
    p = m(None, llr, 20)
    assert m._code(llr.device).kernel == 1
    ref = LdpcCode(qc.H, qc_Z=0).decode(llr, 5, 20, want=("prob",))["prob"]
    assert torch.equal(p, ref)                                    # same bits as the generic kernel
    rng = np.random.RandomState(1)
    proto = rng.randint(-1, 32, size=(5, 10)).astype(np.int16)
    assert LdpcCode(expand_qc(proto, 32)).kernel == 3
