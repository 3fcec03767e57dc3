"""The oracle against the reference's golden vectors (tests/golden, minted by
oracle/make_golden.py from the unmodified reference) - CPU only."""
import json
import os

import numpy as np
import pytest

import bp_oracle as O
import c_oracle as C
import linksim_oracle as LO
from ldpc_b200.codes import ieee80211n_1944_r12, peg_64_32


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "bp_default_code.npz"))


def test_numpy_oracle_bit_exact_on_default_code(gold):
    H, _ = peg_64_32()
    g = O.Graph(H)
    for name in gold["names"]:
        llr, iters, clamp = gold[f"{name}_llr"], int(gold[f"{name}_iters"]), float(gold[f"{name}_clamp"])
        o = O.bp_decode(H, llr, iters, clamp, graph=g)
        assert np.array_equal(o["t"], gold[f"{name}_t"]), name
        assert np.array_equal(o["prob"][:64], gold[f"{name}_prob"]), name
        assert np.array_equal(o["x"][:64], gold[f"{name}_x"]), name
        assert np.array_equal(np.packbits(o["hard"], axis=1), gold[f"{name}_hard"]), name
        assert np.array_equal(o["syndrome"], gold[f"{name}_syndrome"]), name


def test_oracle_decode_bits_ragged_tail(gold):
    H, _ = peg_64_32()
    llr = gold["link4dB_llr"].astype(np.float64)
    out = O.decode_bits(llr, H, 3, 300, 20)
    assert out.dtype == np.float64 and out.shape == llr.shape
    assert not out[300:].any()                                  # 512 // 300 = 1 batch; tail stays zero
    # f64 -> f32 cast of an f32-representable value is exact, so the fixture matches
    assert np.array_equal(np.packbits(out.astype(np.uint8), axis=1), gold["link4dB_ragged300"])


def test_tie_rule_zero_llr(gold):
    H, _ = peg_64_32()
    o = O.bp_decode(H, np.zeros((4, 64), np.float32), 5, 20)
    assert not o["hard"].any() and np.all(o["prob"] == 0.5)     # prob == 0.5 rounds to 0 (np.round)


def test_c_oracle_minsum_bit_exact_and_sp_close():
    for H in (peg_64_32()[0], ieee80211n_1944_r12().H):
        cg, og = C.CGraph(H), O.Graph(H)
        rng = np.random.RandomState(3)
        B = 96 if H.shape[1] == 64 else 24
        llr = (rng.randn(B, H.shape[1]) * 3 + 1.5).astype(np.float32)
        for upd, par in (("minsum", 1.0), ("nms", 0.8125), ("oms", 0.3)):
            a = O.bp_decode(H, llr, 6, 20, update=upd, alpha=par, beta=par, graph=og)
            c = C.decode(cg, llr, 6, 20, upd, par, want=("t", "prob", "hard", "syndrome", "x"))
            for k in ("t", "x", "hard", "syndrome"):
                assert np.array_equal(a[k], c[k]), (upd, k)
        a = O.bp_decode(H, llr, 6, 20, graph=og)
        c = C.decode(cg, llr, 6, 20, "sp")
        assert np.array_equal(a["hard"], c["hard"])
        rel = np.abs(a["t"] - c["t"]) / np.maximum(np.abs(a["t"]), 1e-6)
        assert np.mean(rel > 1e-4) < 2e-3


def test_oracle_vs_dense_reference_n1944(golden_dir):
    """Dense reference at n=1944 (B=8): association order is implementation-defined there,
    so LLR parity is a tolerance: >= 99.9 % of marginals within 1e-4 relative."""
    g = np.load(os.path.join(golden_dir, "bp_wifi1944_dense.npz"))
    code = ieee80211n_1944_r12()
    o = O.bp_decode(code.H, g["llr"], int(g["iters"]), float(g["clamp"]))
    assert np.array_equal(np.packbits(o["hard"], axis=1), g["hard"])
    rel = np.abs(o["t"] - g["t"]) / np.maximum(np.abs(g["t"]), 1e-30)
    assert np.mean(rel > 1e-4) < 1e-3
    assert np.array_equal(o["t"], g["oracle_t"])               # the oracle itself is deterministic


def test_frontend_oracle_against_reference_fixtures(golden_dir):
    g = np.load(os.path.join(golden_dir, "frontend.npz"))
    _, G = peg_64_32()
    bits = g["bits"].astype(np.int64)
    enc = LO.encode_bits(bits, G)
    assert np.array_equal(enc.astype(np.uint8), g["enc"])
    tx = LO.modulate_bits(enc)
    assert np.array_equal(tx, g["tx_symbols"])
    for snr in (4, 15):
        tag = f"snr{snr}"
        np.random.seed(21)
        rx_signal, rx_symbols, rx_llrs, tx_signal = LO.gen_data(tx, float(snr), 32)
        assert np.array_equal(rx_signal, g[tag + "_rx_signal"]) and np.array_equal(rx_llrs, g[tag + "_rx_llrs"])
        assert np.array_equal(rx_symbols, g[tag + "_rx_symbols"]) and np.array_equal(tx_signal, g[tag + "_tx_signal"])
        for qbits, clip in ((1, 1.2), (3, 1.18), (3, 10.0), (5, 2.0), (6, 0.4)):
            assert np.array_equal(LO.quantizer(rx_signal, qbits, clip), g[f"{tag}_quant_b{qbits}_c{clip}"])
        for qbits, clipdb in ((1, 0.0), (3, 0.0), (3, 5.0), (5, 10.0)):
            q = LO.gen_qdata(rx_signal, float(snr), qbits, np.power(10, clipdb / 10), 32)
            assert np.array_equal(q[0], g[f"{tag}_qdata_b{qbits}_c{int(clipdb)}_signal"])
            assert np.array_equal(q[2], g[f"{tag}_qdata_b{qbits}_c{int(clipdb)}_llrs"])
        qllr, _ = LO.agc_quantized_frontend(rx_signal, float(snr), 3, 1.0, 32, agc_clip=10)
        assert np.array_equal(qllr, g[tag + "_agc_llrs"])


def test_published_ber_fixture(golden_dir):
    with open(os.path.join(golden_dir, "published_ber.json")) as f:
        d = json.load(f)
    assert d["snrdb"] == [float(i) for i in range(11)]
    assert abs(d["coded_ber"][4] - 4.46e-3) < 1e-4 and abs(d["coded_bler"][4] - 0.128) < 1e-3


def test_oracle_reproduces_published_ber_point():
    """Statistical pin: 4 dB point of the shipped pickle (3 iterations, clamp 20)."""
    H, G = peg_64_32()
    np.random.seed(1234)
    N = 4096
    bits = LO.create_bits(N * 32)
    enc = LO.encode_bits(bits, G)
    tx = LO.modulate_bits(enc)
    _, _, llrs, _ = LO.gen_data(tx, 4.0, 32)
    L = llrs.reshape(-1, 64); E = enc.reshape(-1, 64)
    dec = O.decode_bits(L, H, 3, 1024, 20)
    m = LO.error_metrics(L, dec, E, 32)
    ber = m["info_errs"] / m["info_bits"]; bler = m["frame_errs"] / m["frames"]; unc = m["uncoded_errs"] / m["bits"]
    # binomial 4-sigma bands around the published values
    assert abs(unc - 5.647e-2) < 4 * np.sqrt(5.647e-2 / (N * 64))
    assert abs(bler - 0.1276) < 4 * np.sqrt(0.1276 * 0.8724 / N)
    assert abs(ber - 4.46e-3) < 1.5e-3
