"""MLP demapper (reference nn/llr.py:54-73, evaluate_quantized_snr.py:135-173).

CPU: the numpy oracle against the golden vectors minted from the reference's own checkpoint
(tests/golden/nn_demapper.npz, oracle/make_golden_nn.py).  GPU: the native tensor-core kernel
through the drop-in nn.llr module against the same vectors, then through the decoder.
Tolerance (fp32 GEMM chains do not pin a summation order): 1e-5 of the output scale for the
fp32-equivalent mode (2 binary16 planes, the default) and for 3 planes.
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
@pytest.mark.parametrize("splits,tol", [(2, 1e-5), (3, 1e-5), (1, 5e-3)])
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
    m = _model(2)
    x = torch.tensor(GOLD["snr15_x"][:100])
    y = m.module(x)                                      # CPU tensor in -> CPU tensor out (DataParallel itself returns CUDA)
    assert not y.is_cuda and _scale_err(y.numpy(), GOLD["snr15_llr"][:100]) < 1e-5


@pytest.mark.gpu
def test_mlp_single_launch_chain_equals_per_layer_launches():
    """LDPC_MLP_CHAIN (one cooperative launch per chunk, activations handed over through L2 between groups of 4 SMs) against
    LDPC_MLP_PER_LAYER: same MMA order per tile and same epilogue arithmetic, so the results are bit-identical - on the
    reference's two network shapes (nn/llr.py:46-73; a 64-wide first layer without bias or tanh rotates over the group), a
    two-layer net, ragged row counts, several chunks per call and enough rows that every group wraps its ring many times."""
    import torch
    from ldpc_b200.mlp import NativeMLP
    rng = np.random.RandomState(11)
    shapes = (([65, 512, 512, 512, 64], None, True), ([64, 64, 512, 512, 512, 64], [0, 1, 1, 1, 0], False), ([64, 128, 64], None, True),
              ([33, 256, 384, 64], None, True))
    for dims, acts, bias in shapes:
        W = [(rng.rand(dims[l + 1], dims[l]).astype(np.float32) * 2 - 1) / np.sqrt(dims[l]) for l in range(len(dims) - 1)]
        Bs = [(rng.rand(d).astype(np.float32) - 0.5) for d in dims[1:]] if bias else None
        for chunk in (0, 384):
            a = NativeMLP(W, Bs, acts, chunk_rows=chunk, mode="chain")
            b = NativeMLP(W, Bs, acts, chunk_rows=chunk, mode="per_layer")
            c = NativeMLP(W, Bs, acts, chunk_rows=chunk, mode="chain_pairs")   # the same chain on cta_group::2 pairs of SMs
            for B in ((1, 127, 129, 257, 1000, 150001) if chunk == 0 else (1, 385, 5000)):
                x = torch.tensor(rng.randn(B, dims[0]).astype(np.float32) * 0.7).cuda()
                ya, yb, yc = a(x), b(x), c(x)
                assert torch.equal(ya, yb), (dims, chunk, B, float((ya - yb).abs().max()))
                assert torch.equal(yc, yb), (dims, chunk, B, float((yc - yb).abs().max()))
                assert torch.equal(a(x), ya) and torch.equal(c(x), yc)   # reproducible run to run (counters restart per launch)
    # shapes the chain cannot run are refused in "chain" mode and served per layer in "auto"
    W = [(rng.rand(1024, 64).astype(np.float32) - 0.5) / 8, (rng.rand(64, 1024).astype(np.float32) - 0.5) / 32]
    with pytest.raises(RuntimeError, match="single-launch chain"):
        NativeMLP(W, mode="chain")
    with pytest.raises(RuntimeError, match="single-launch chain"):
        NativeMLP([w[:64] for w in W[:1]] + [W[1][:, :64]], splits=3, mode="chain")
    assert NativeMLP(W)(torch.zeros(5, 64, device="cuda")).shape == (5, 64)


@pytest.mark.gpu
def test_nn_demapper_then_decoder_matches_reference_bits():
    """evaluate_quantized_snr.py:150-173: llr_est = LLRest(x); bits_nn = decode_bits(llr_est, H, ...)."""
    import torch
    from ofdm.ofdm_functions import decode_bits
    from bp.parity import H
    m = _model(2)
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
    ref = NO.mlp_forward(state, x.numpy(), layers=NO.PLAIN_LAYERS, acts=NO.PLAIN_ACTS)
    assert _scale_err(y, ref) < 1e-5


def _default_code():
    from ldpc_b200.codes import peg_64_32
    from ldpc_b200.decoder import LdpcCode
    from ldpc_b200.linksim import attach_generator
    H, G = peg_64_32()
    return attach_generator(LdpcCode(H), G), H


@pytest.mark.gpu
def test_frontend_samples_are_the_time_signal_behind_the_llrs():
    """The sample rows the MLP sees (evaluate_quantized_snr.py:135-140) and the conventional LLRs of the
    same launch describe the same received signal: DFT(samples) -> QPSK LLR reproduces the kernel's LLRs,
    the scaled samples sit on the quantizer grid, the last column is the linear SNR."""
    import torch
    from ldpc_b200.linksim import LinkConfig, sim_generate
    code, _ = _default_code()
    for N, qbits in ((32, 3), (32, 0)):
        cfg = LinkConfig(snr_db=15.0, ofdm_size=N, qbits=qbits, agc_mode=1, seed=77)
        cwp, llr, smp = sim_generate(code, cfg, 5, 300, want_samples=True)
        cwp2, llr2 = sim_generate(code, cfg, 5, 300)
        assert torch.equal(cwp, cwp2) and torch.equal(llr, llr2)          # asking for samples changes nothing else
        s = smp.cpu().numpy().astype(np.float64)
        snr = 10 ** 1.5
        assert np.allclose(s[:, 2 * N], snr, rtol=1e-6)
        t = s[:, :N] + 1j * s[:, N:2 * N]
        k = np.arange(N)
        W = np.exp(-2j * np.pi * np.outer(k, k) / N) / np.sqrt(N)          # ofdm_functions.py:86-93
        R = t @ W.T
        ref = np.stack([-2 * np.sqrt(2) * snr * R.real, -2 * np.sqrt(2) * snr * R.imag], axis=2).reshape(300, 2 * N)
        got = llr.cpu().numpy()
        assert np.max(np.abs(got - ref)) <= 2e-4 * np.max(np.abs(ref))
        if qbits:
            factor = 10.0 / (0.5 * (1 + 1 / snr))
            step = 2 * 10.0 / (2 ** qbits - 1)
            lev = s[:, :2 * N] * factor / step
            edge = 2 ** qbits / 2 - 1 / step          # the reference clips at (L/2) * step - 1 in VALUE units (ofdm_functions.py:45)
            off = np.minimum(np.abs(lev - np.round(lev)), np.abs(np.abs(lev) - edge))
            assert np.max(off) < 1e-3


@pytest.mark.gpu
def test_nn_link_counters_are_exact():
    """sim_run_nn (front end -> MLP -> decoder -> fused counters) against the same chain evaluated step by
    step with separate calls and numpy integer arithmetic (evaluate_quantized_snr.py:169-188)."""
    import torch
    from ldpc_b200.linksim import LinkConfig, sim_generate, sim_run_nn
    code, H = _default_code()
    m = _model(2).module
    cfg = LinkConfig(snr_db=10.0, ofdm_size=32, qbits=3, agc_mode=1, iters=10, update="sp", clamp_value=100.0, seed=5)
    cnt = 3000
    c = sim_run_nn(code, cfg, m, 0, cnt, chunk=1024).cpu().numpy()
    cwp, _, smp = sim_generate(code, cfg, 0, cnt, want_samples=True)
    llr_est = m(smp)
    out = code.decode(llr_est, 10, 100.0, update="sp", want=("hard",))
    hard = out["hard"].cpu().numpy()
    enc = np.unpackbits(cwp.cpu().numpy(), axis=1)[:, :64]
    L = llr_est.cpu().numpy()
    unc = int(np.sum(((np.sign(L) + 1) // 2).astype(np.uint8) != enc))
    inf = int(np.sum(hard[:, :32] != enc[:, :32]))
    fe = int(np.sum(np.any(hard != enc, axis=1)))
    assert c.tolist() == [unc, inf, fe, cnt * 64, cnt]
    # the learned demapper is a sane LLR estimator at this operating point (reference: coded BER 4e-4 at 10 dB)
    assert inf / (cnt * 32) < 5e-3


@pytest.mark.gpu
def test_evaluate_full_writes_every_key_plots_py_reads_and_resumes(tmp_path):
    """evaluate_quantized_snr.py:192-212 / plots.py:11-27: the full result set on one noise realisation,
    checkpointed per SNR point; a restart reuses the finished points and gives identical counters."""
    import torch
    from ldpc_b200.linksim import LinkConfig, evaluate_full, evaluate_point, results_dict, sim_run
    code, _ = _default_code()
    m = _model(2).module
    cfgs = [LinkConfig(snr_db=s, ofdm_size=32, qbits=3, agc_mode=1, iters=10, update="sp", clamp_value=100.0, seed=21) for s in (5.0, 10.0, 15.0)]
    state = str(tmp_path / "state")
    c1, w1 = evaluate_full(code, cfgs, m, 4096, state_path=state, chunk=1500)
    assert os.path.exists(state + ".rank0of1.npz")
    res = results_dict([5.0, 10.0, 15.0], c1, w1, 64, 32)
    for key in ("snrdb", "uncoded_ber", "coded_ber", "coded_bler", "uncoded_ber_nn", "coded_ber_nn", "coded_bler_nn",
                "uncoded_ber_quantized", "coded_ber_quantized", "coded_bler_quantized", "wmse_nn", "wmse_quantized"):
        assert key in res and len(res[key]) == 3, key
    # resume: nothing is recomputed (poison the demapper), same numbers
    c2, w2 = evaluate_full(code, cfgs, lambda x: 1 / 0, 4096, state_path=state, chunk=1500)
    assert np.array_equal(c1, c2) and np.array_equal(w1, w2)
    # the traditional and quantized rows are the fused simulator's counters for the same seed
    for i, cfg in enumerate(cfgs):
        assert c1[i, 2].tolist() == sim_run(code, cfg, 0, 4096).cpu().numpy().tolist()
    # sanity of the physics: quantization hurts, the learned demapper recovers most of it at 15 dB
    assert res["coded_ber"][2] <= res["coded_ber_nn"][2] <= 5e-3 and res["wmse_nn"][2] < res["wmse_quantized"][2]
    # sharding invariance: two half-ranges add up to the whole
    a, _ = evaluate_point(code, cfgs[1], m, 0, 2048)
    b, _ = evaluate_point(code, cfgs[1], m, 2048, 2048)
    assert np.array_equal((a + b).cpu().numpy(), c1[1])


@pytest.mark.gpu
def test_data_parallel_replicas_reuse_native_handles():
    """evaluate_quantized_snr.py:53-57 wraps the demapper in nn.DataParallel: on one GPU it forwards to the module, on
    several it re-broadcasts the parameters every call - the native handle must be reused (value checksum), not rebuilt."""
    import torch
    m = _model(2)
    x = torch.tensor(GOLD["snr15_x"], dtype=torch.float, device="cuda")
    y1 = m(x)
    handles = {k: v[1] for k, v in m.module._native.items()}
    y2 = m(x)
    assert torch.equal(y1, y2)
    assert all(m.module._native[k][1] is h for k, h in handles.items())
    # a weight update invalidates the handle
    with torch.no_grad():
        m.module.final.bias.add_(1.0)
    y3 = m(x)
    assert torch.allclose(y3, y1 + 1.0, atol=1e-4)
