"""Front end (K2) parity: the drop-in ofdm.ofdm_functions against the reference's golden
vectors (tests/golden/frontend.npz, minted from the unmodified reference) and the oracle.
Bars: bits / QPSK symbols / quantizer outputs bit-exact; float64 transforms and LLRs within
1e-10 absolute (the reference's DFT matrix itself carries ~1e-13 argument-reduction error for
large x*y, ofdm_functions.py:86-93); seeded numpy noise reproduces the reference stream."""
import os

import numpy as np
import pytest
import torch

import linksim_oracle as LO
from ldpc_b200.codes import peg_64_32

pytestmark = pytest.mark.gpu
TOL = 1e-10


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "frontend.npz"))


def test_encode_and_modulate_bit_exact(g):
    import ofdm.ofdm_functions as F
    _, G = peg_64_32()
    enc = F.encode_bits(g["bits"].astype(np.int64), G)
    assert enc.shape == (1, g["enc"].size) and enc.dtype == np.float64
    assert np.array_equal(enc.astype(np.uint8), g["enc"])
    tx = F.modulate_bits(enc)
    assert tx.dtype == np.complex128 and np.array_equal(tx, g["tx_symbols"])


@pytest.mark.parametrize("snr", [4, 15])
def test_gen_data_reproduces_seeded_reference(g, snr):
    import ofdm.ofdm_functions as F
    tag = f"snr{snr}"
    np.random.seed(21)
    rx_signal, rx_symbols, rx_llrs, tx_signal = F.gen_data(g["tx_symbols"], float(snr), 32)
    for got, key in ((rx_signal, "_rx_signal"), (rx_symbols, "_rx_symbols"), (rx_llrs, "_rx_llrs"), (tx_signal, "_tx_signal")):
        ref = g[tag + key]
        assert got.shape == ref.shape and got.dtype == ref.dtype
        assert np.abs(got - ref).max() < TOL * max(1.0, np.abs(ref).max()), key
    # hard decisions of the channel LLRs identical to the reference's
    assert np.array_equal(np.sign(rx_llrs), np.sign(g[tag + "_rx_llrs"]))


@pytest.mark.parametrize("snr", [4, 15])
def test_quantizer_and_gen_qdata(g, snr):
    import ofdm.ofdm_functions as F
    tag = f"snr{snr}"
    rx = g[tag + "_rx_signal"]
    for qbits, clip in ((1, 1.2), (3, 1.18), (3, 10.0), (5, 2.0), (6, 0.4)):
        q = F.quantizer(rx, qbits, clip)
        assert np.array_equal(q, g[f"{tag}_quant_b{qbits}_c{clip}"]), (qbits, clip)   # bit-exact levels
    for qbits, clipdb in ((1, 0.0), (3, 0.0), (3, 5.0), (5, 10.0)):
        qs, qsym, qllr = F.gen_qdata(rx, float(snr), qbits, np.power(10, clipdb / 10), 32)
        assert np.array_equal(qs, g[f"{tag}_qdata_b{qbits}_c{int(clipdb)}_signal"])
        ref = g[f"{tag}_qdata_b{qbits}_c{int(clipdb)}_llrs"]
        assert np.abs(qllr - ref).max() < TOL * max(1.0, np.abs(ref).max())
    # quantizer accepts the array-valued qbits the reference passes (quantized_snr.py:101)
    assert np.array_equal(F.quantizer(rx, np.array([3]), 10.0), g[f"{tag}_quant_b3_c10.0"])


@pytest.mark.parametrize("N", [32, 64, 128, 256])
def test_ofdm_sizes_against_oracle(N):
    from ldpc_b200.linksim import demodulate_signal, transmit_symbols
    rng = np.random.RandomState(N)
    L = N * 37
    sym = ((1 - 2 * rng.randint(0, 2, L)) + 1j * (1 - 2 * rng.randint(0, 2, L))) / np.sqrt(2)
    noise = (rng.randn(L) + 1j * rng.randn(L)) * 0.1
    rx, tx = transmit_symbols(sym.reshape(1, -1), N, 10.0, noise=noise)
    orx, otx = LO.transmit_symbols(sym.reshape(1, -1), N, 10.0, noise=noise.reshape(-1, N).T)
    assert np.abs(tx - otx).max() < TOL and np.abs(rx - orx).max() < TOL
    llr, rs = demodulate_signal(rx, N, 10.0)
    ollr, ors = LO.demodulate_signal(orx, N, 10.0)
    assert np.abs(rs - ors).max() < TOL and np.abs(llr - ollr).max() < 1e-8
    # complex64 path agrees with the float64 one to single precision
    rx32, _ = transmit_symbols(sym.astype(np.complex64).reshape(1, -1), N, 10.0, noise=noise.astype(np.complex64))
    assert rx32.dtype == np.complex64 and np.abs(rx32 - rx).max() < 2e-6


def test_device_rng_noise_statistics():
    from ldpc_b200.linksim import transmit_symbols
    L = 64 * 4096
    sym = np.zeros((1, L), np.complex128)
    snr = 10 ** 0.4
    rx, tx = transmit_symbols(sym, 64, snr, device_rng=True, seed=5)
    assert np.abs(tx).max() == 0
    assert abs(rx.real.var() - 0.5 / snr) < 0.01 * 0.5 / snr and abs(rx.imag.var() - 0.5 / snr) < 0.01 * 0.5 / snr
    assert abs(rx.real.mean()) < 3e-3 and abs(np.mean(rx.real * rx.imag)) < 3e-3
    k4 = np.mean(rx.real ** 4) / rx.real.var() ** 2
    assert abs(k4 - 3.0) < 0.05                                            # gaussian kurtosis
    rx2, _ = transmit_symbols(sym, 64, snr, device_rng=True, seed=5)
    assert np.array_equal(rx, rx2)                                         # counter-based: reproducible


@pytest.mark.parametrize("snr", [4, 15])
def test_agc_quantizer_frontend_identical_inputs(g, snr):
    """SURVEY 8(a) row a14, evaluate_quantized_snr.py:96-133, on the reference's OWN received samples
    (frontend.npz snr*_rx_signal / snr*_noise) against its own LLRs (snr*_agc_llrs):
    (i) the float64 drop-in functions (device quantizer + device DFT / LLR) to 1e-10;
    (ii) the float32 simulator chain (the code the fused single-launch kernel runs) fed the same transmitted bits
         and the same noise realisation: ADC levels identical, LLRs to single precision."""
    import ofdm.ofdm_functions as F
    from ldpc_b200.decoder import LdpcCode
    from ldpc_b200.linksim import LinkConfig, sim_frontend
    tag = f"snr{snr}"
    ref = g[tag + "_agc_llrs"]
    rx = g[tag + "_rx_signal"]
    snr_lin = np.power(10, snr / 10)
    # (i) the script's lines with the drop-in functions
    factor = 10 / (.5 * (1 + 1 / snr_lin)) * 1.0
    scaled = (factor * rx.reshape((-1, 32)).T).T.reshape((1, -1))
    q = F.quantizer(scaled, 3, 10)
    resc = q.reshape((-1, 32)).T / factor
    llr64, _ = F.demodulate_signal(resc.T.reshape((1, -1)), 32, snr_lin)
    assert llr64.shape == ref.shape
    assert np.abs(llr64 - ref).max() < TOL * max(1.0, np.abs(ref).max())
    # (ii) the simulator's float32 chain on the same bits and the same noise
    H, _ = peg_64_32()
    code = LdpcCode(H)
    enc = g["enc"].reshape(-1, 64)
    cwp = torch.as_tensor(np.packbits(enc, axis=1)).cuda()
    noise = torch.as_tensor(g[tag + "_noise"].reshape(-1, 1, 32).astype(np.complex64)).cuda()
    cfg = LinkConfig(snr_db=float(snr), ofdm_size=32, qbits=3, agc_mode=1, agc_clip=10.0, clip_ratio=1.0)
    llr32, smp = sim_frontend(code, cfg, cwp, noise, want_samples=True)
    llr32 = llr32.cpu().numpy().reshape(1, -1).astype(np.float64)
    smp = smp.cpu().numpy().astype(np.float64)
    _, qresc = LO.agc_quantized_frontend(rx, float(snr), 3, 1.0, 32, agc_clip=10)        # [32, n_ofdm] complex
    step = 2 * 10 / (2 ** 3 - 1)
    lev_ref = np.concatenate([qresc.real.T, qresc.imag.T], axis=1) * factor / step        # [n_ofdm, 64] ADC output in steps
    lev_gpu = smp[:, :64] * factor / step
    assert np.abs(lev_gpu - lev_ref).max() < 1e-4, "ADC output levels differ"          # same level for every sample
    assert np.allclose(smp[:, 64], snr_lin, rtol=1e-6)
    err = np.abs(llr32 - ref).max() / np.abs(ref).max()
    assert err < 3e-6, err
    # without a noise array the same entry point uses the simulator's Philox stream: statistics only
    llr_rng = sim_frontend(code, cfg, cwp).cpu().numpy()
    hard_ok = np.mean((llr_rng > 0) == (enc > 0))
    assert hard_ok > (0.9 if snr == 4 else 0.97)
