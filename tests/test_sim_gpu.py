"""Fused link simulation (K2a + K2b + K1 + K3): structure, sharding invariance (exact integer
equality), and Monte-Carlo agreement with the oracle / the reference's published BER pickle."""
import json
import os

import numpy as np
import pytest
import torch

import bp_oracle as O
import linksim_oracle as LO
from ldpc_b200.codes import ieee80211n_1944_r12, peg_64_32

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dsim():
    from ldpc_b200.decoder import LdpcCode
    from ldpc_b200.linksim import attach_generator
    H, G = peg_64_32()
    return attach_generator(LdpcCode(H), G)


@pytest.fixture(scope="module")
def wsim():
    from ldpc_b200.decoder import LdpcCode
    from ldpc_b200.linksim import attach_generator
    qc = ieee80211n_1944_r12()
    return attach_generator(LdpcCode(qc.H, qc_Z=81, qc_proto=qc.proto))


def test_generated_codewords_and_noiseless_llrs(wsim):
    from ldpc_b200.linksim import LinkConfig, sim_generate
    qc = ieee80211n_1944_r12()
    for N in (64, 128, 256, 32):
        cwp, llr = sim_generate(wsim, LinkConfig(snr_db=60.0, ofdm_size=N, seed=3), 1000, 256)
        cw = np.unpackbits(cwp.cpu().numpy(), axis=1)[:, :qc.n]
        assert not ((qc.H.astype(np.int64) @ cw.T.astype(np.int64)) % 2).any()       # valid codewords
        assert 0.45 < cw[:, :qc.k].mean() < 0.55
        L = llr.cpu().numpy()
        snr = 10.0 ** 6
        # noiseless limit: llr = -2*sqrt(2)*snr*Re(R) = 2*snr*(2b-1)
        assert np.allclose(L / (2 * snr), 2.0 * cw - 1.0, atol=2e-2), N
    # same global index -> same codeword, whatever the OFDM size or the call it came from
    a, _ = sim_generate(wsim, LinkConfig(snr_db=3.0, ofdm_size=64, seed=3), 1000, 8)
    b, _ = sim_generate(wsim, LinkConfig(snr_db=9.0, ofdm_size=128, seed=3), 1004, 4)
    assert torch.equal(a[4:], b)


def test_sim_equals_generate_plus_decode_plus_count(wsim):
    from ldpc_b200.linksim import LinkConfig, sim_generate, sim_run
    qc = ieee80211n_1944_r12()
    cfg = LinkConfig(snr_db=1.5, ofdm_size=64, iters=10, update="minsum", clamp_value=20.0, seed=11)
    cnt = sim_run(wsim, cfg, 5000, 3000).cpu().numpy()
    cwp, llr = sim_generate(wsim, cfg, 5000, 3000)
    out = wsim.decode(llr, 10, 20.0, update="minsum", want=("hard",))
    cw = torch.as_tensor(np.unpackbits(cwp.cpu().numpy(), axis=1)[:, :qc.n]).cuda()
    ref = wsim.count_errors(out["hard"], cw, qc.k, llr=llr).cpu().numpy()
    assert cnt.tolist() == ref.tolist()
    assert cnt[3] == 3000 * qc.n and cnt[4] == 3000 and 0 < cnt[0] < cnt[3] // 4


def test_sharding_invariance_exact(dsim):
    """Counters depend only on (seed, global codeword index): 1 shard == 3 ragged shards,
    whatever the workspace chunking."""
    from ldpc_b200.linksim import LinkConfig, shard_range, sim_run
    cfg = LinkConfig(snr_db=4.0, ofdm_size=32, qbits=3, agc_mode=1, iters=5, update="sp", clamp_value=20.0, seed=99)
    total = 20000
    one = sim_run(dsim, cfg, 0, total).cpu().numpy()
    acc = torch.zeros(5, dtype=torch.int64, device="cuda")
    ws = torch.empty(1024 * (4 * 64 + 16), dtype=torch.uint8, device="cuda")      # forces 1024-codeword chunks
    for r in range(3):
        first, count = shard_range(total, r, 3)
        sim_run(dsim, cfg, first, count, acc, ws)
    assert acc.cpu().numpy().tolist() == one.tolist()
    assert shard_range(10, 0, 4) == (0, 3) and shard_range(10, 3, 4) == (8, 2)


def test_default_code_matches_published_ber_and_oracle(dsim, golden_dir):
    """Config 1: (64,32) code, QPSK/OFDM-32/AWGN, 3 iterations, clamp 20 - the reference's
    shipped BER pickle is the acceptance band (binomial 4 sigma)."""
    from ldpc_b200.linksim import LinkConfig, rates, sim_run
    pub = json.load(open(os.path.join(golden_dir, "published_ber.json")))
    Ncw = 1 << 17
    for snr in (0, 2, 4, 6):
        c = sim_run(dsim, LinkConfig(snr_db=float(snr), ofdm_size=32, iters=3, update="sp", clamp_value=20.0, seed=7), 0, Ncw).cpu().numpy()
        r = rates(c, 64, 32)
        i = pub["snrdb"].index(float(snr))
        pub_n = 1 << 15
        for key, nbits in (("uncoded_ber", 64), ("coded_ber", 32), ("coded_bler", 1)):
            p = pub[key][i]
            # independent-trial sigma of both estimates; info-bit errors are bursty per frame -> x4 slack
            sig = np.sqrt(p * (1 - p) * (1.0 / (Ncw * nbits) + 1.0 / (pub_n * nbits))) * (4 if key == "coded_ber" else 1)
            assert abs(float(r[key]) - p) < 4 * sig + 1e-6, (snr, key, float(r[key]), p)


def test_quantized_link_matches_oracle_statistics(dsim):
    """Config 2: 3-bit ADC with the script AGC (evaluate_quantized_snr.py:96-133), 15 dB,
    10 iterations, clamp 100: GPU simulation vs the oracle chain on fresh random data."""
    from ldpc_b200.linksim import LinkConfig, rates, sim_run
    H, G = peg_64_32()
    np.random.seed(4321)
    Ncw = 1 << 14
    enc = LO.encode_bits(LO.create_bits(Ncw * 32), G)
    rx_signal, _, _, _ = LO.gen_data(LO.modulate_bits(enc), 15.0, 32)
    qllr, _ = LO.agc_quantized_frontend(rx_signal, 15.0, 3, 1.0, 32, agc_clip=10)
    L = qllr.reshape(-1, 64); E = enc.reshape(-1, 64)
    dec = O.decode_bits(L, H, 10, 1024, 100)
    m = LO.error_metrics(L, dec, E, 32)
    o_unc, o_bler = m["uncoded_errs"] / m["bits"], m["frame_errs"] / m["frames"]
    Ng = 1 << 18
    c = sim_run(dsim, LinkConfig(snr_db=15.0, ofdm_size=32, qbits=3, agc_mode=1, agc_clip=10.0, clip_ratio=1.0,
                                 iters=10, update="sp", clamp_value=100.0, seed=5), 0, Ng).cpu().numpy()
    r = rates(c, 64, 32)
    assert abs(float(r["uncoded_ber"]) - o_unc) < 5 * np.sqrt(o_unc / (Ncw * 64)) + 5 * np.sqrt(o_unc / (Ng * 64))
    assert abs(float(r["coded_bler"]) - o_bler) < 5 * np.sqrt(o_bler / Ncw) + 5 * np.sqrt(o_bler / Ng)
    assert 0.008 < float(r["uncoded_ber"]) < 0.02                      # survey probe: 1.35e-2


def _wilson(k, n, z=3.9):
    """Wilson score interval of a binomial proportion (z = 3.9: 1e-4 two-sided)."""
    p = k / n
    d = 1 + z * z / n
    c = (p + z * z / (2 * n)) / d
    h = z * np.sqrt(p * (1 - p) / n + z * z / (4 * n * n)) / d
    return c - h, c + h


@pytest.mark.parametrize("qbits", [0, 3])
def test_wifi1944_link_matches_oracle_chain(wsim, qbits):
    """BASELINE config 4 at the oracle's size: n=1944 over OFDM-64 (16 symbols per codeword, 52 null subcarriers
    in the last one), min-sum x10, 2^14 codewords per point at three SNRs.  The GPU counters (Philox noise) and
    the oracle chain (numpy noise -> oracle/linksim_oracle.framed_link_llrs -> C oracle decoder) are independent
    Monte-Carlo estimates of the same curve: uncoded BER, coded BER and BLER must agree inside binomial
    confidence intervals."""
    import c_oracle as C
    from ldpc_b200.codes import ieee80211n_1944_r12
    from ldpc_b200.linksim import LinkConfig, sim_run
    qc = ieee80211n_1944_r12()
    g = C.CGraph(qc.H)
    Ncw = 1 << 14
    rng = np.random.RandomState(2024 + qbits)
    enc = qc.encode(rng.randint(0, 2, (Ncw, qc.k)).astype(np.uint8)).astype(np.float64)
    for snr_db in ((1.75, 2.25, 2.75) if qbits == 0 else (3.0, 3.75, 4.5)):
        np.random.seed(int(snr_db * 100) + qbits)
        llr, _ = LO.framed_link_llrs(enc, snr_db, 64, qbits=qbits)
        dec = C.decode(g, llr.astype(np.float32), 10, 20.0, "minsum", want=("hard",))["hard"]
        m = LO.error_metrics(llr, dec, enc, qc.k)
        c = sim_run(wsim, LinkConfig(snr_db=snr_db, ofdm_size=64, qbits=qbits, agc_mode=1, iters=10, update="minsum",
                                     clamp_value=20.0, seed=31 + qbits), 0, Ncw).cpu().numpy()
        assert c[3] == m["bits"] and c[4] == m["frames"]
        # uncoded bit errors are independent per bit
        lo, hi = _wilson(m["uncoded_errs"], m["bits"])
        lo2, hi2 = _wilson(int(c[0]), int(c[3]))
        assert lo <= hi2 and lo2 <= hi, (snr_db, "uncoded", m["uncoded_errs"], int(c[0]))
        # frame errors are independent per frame
        lo, hi = _wilson(m["frame_errs"], m["frames"])
        lo2, hi2 = _wilson(int(c[2]), int(c[4]))
        assert lo <= hi2 and lo2 <= hi, (snr_db, "bler", m["frame_errs"], int(c[2]))
        # information-bit errors come in bursts of one frame: compare the mean burst size given the frame counts
        if m["frame_errs"] >= 50 and c[2] >= 50:
            b_o, b_g = m["info_errs"] / m["frame_errs"], c[1] / c[2]
            assert abs(b_o - b_g) < 0.25 * max(b_o, b_g), (snr_db, "errors per bad frame", b_o, b_g)
        else:
            assert abs(m["info_errs"] - int(c[1])) < 60 * 50


def test_sweep_single_rank(wsim):
    from ldpc_b200.linksim import LinkConfig, rates, sweep
    cfgs = [LinkConfig(snr_db=s, ofdm_size=64, iters=10, update="minsum", clamp_value=20.0, seed=1) for s in (0.0, 2.0, 4.0)]
    c = sweep(wsim, cfgs, 4096)
    assert c.shape == (3, 5) and (c[:, 4] == 4096).all()
    r = rates(c, 1944, 972)
    assert r["coded_bler"][0] >= r["coded_bler"][1] >= r["coded_bler"][2]
    assert r["uncoded_ber"][0] > r["uncoded_ber"][2] > 0


@pytest.mark.parametrize("N,update,qbits", [(64, "minsum", 0), (64, "sp", 3), (64, "nms", 0), (32, "minsum", 3), (128, "minsum", 0), (256, "minsum", 2)])
def test_single_launch_simulator_equals_three_launch_chain(wsim, N, update, qbits):
    """The fused kernel (bits -> encode -> OFDM -> AWGN -> ADC -> LLR -> BP -> counters in ONE launch)
    must reproduce the three-launch chain's integer counters exactly: same Philox draws, same
    encoder output, same float bits (both paths compile the link chain without FMA contraction)."""
    from ldpc_b200.linksim import LinkConfig, sim_run
    kw = dict(snr_db=2.5, ofdm_size=N, qbits=qbits, agc_mode=1, iters=6, update=update, clamp_value=20.0,
              param=0.8125 if update == "nms" else 1.0, seed=77)
    fused = sim_run(wsim, LinkConfig(**kw), 12345, 1000).cpu().numpy()          # 1000: ragged vs the 3-codeword tile
    chain = sim_run(wsim, LinkConfig(force_unfused=True, **kw), 12345, 1000).cpu().numpy()
    assert fused.tolist() == chain.tolist()
    assert fused[4] == 1000 and fused[0] > 0


@pytest.mark.parametrize("n,rate", [(648, "1/2"), (648, "5/6"), (1296, "3/4"), (1944, "2/3"), (1944, "5/6")])
def test_wifi_family_single_launch_simulator(n, rate):
    """The other compiled 802.11n codes in the single-launch simulator (device dual-diagonal encoder of that prototype,
    OFDM-64, min-sum): counters equal to the three-launch chain, transmitted words are codewords of H."""
    from ldpc_b200.codes import ieee80211n
    from ldpc_b200.decoder import LdpcCode
    from ldpc_b200.linksim import LinkConfig, attach_generator, sim_generate, sim_run
    qc = ieee80211n(n, rate)
    code = attach_generator(LdpcCode(qc.H, qc_Z=qc.Z, qc_proto=qc.proto))
    assert code.kernel == 1
    R = qc.k / qc.n
    kw = dict(snr_db=10 * np.log10(2 * R) + 1.5 + 2.5 * R, ofdm_size=64, qbits=0, iters=8, update="minsum", clamp_value=20.0, seed=5 + n)
    fused = sim_run(code, LinkConfig(**kw), 777, 1001).cpu().numpy()
    chain = sim_run(code, LinkConfig(force_unfused=True, **kw), 777, 1001).cpu().numpy()
    assert fused.tolist() == chain.tolist()
    assert fused[4] == 1001 and fused[3] == 1001 * n and 0 < fused[0] < fused[3] // 4 and fused[2] < 1001
    cwp, _ = sim_generate(code, LinkConfig(**kw), 777, 64)
    cw = np.unpackbits(cwp.cpu().numpy(), axis=1)[:, :n]
    assert not ((qc.H.astype(np.int64) @ cw.T.astype(np.int64)) % 2).any()
    u = cw[:, :qc.k]
    assert np.array_equal(qc.encode(u), cw)                      # the device encoder is the library's linear-time encoder


def test_rayleigh_fading_channel_and_tanh_compander(wsim):
    """North-star extras that the reference does not have (it is AWGN + uniform quantizer only): flat Rayleigh block fading
    with a coherent receiver, and a tanh compander in front of the ADC.  Defined in include/ldpc_b200.h and restated in
    oracle/linksim_oracle.framed_link_llrs.  (i) uncoded BER against the closed form 0.5 (1 - sqrt(g / (1 + g))), g = snr / 2;
    (ii) the single-launch simulator and the three-launch chain give identical counters with both options on;
    (iii) coded curves against the oracle chain (numpy fading + noise, C oracle decoder) inside binomial intervals."""
    import c_oracle as C
    from ldpc_b200.codes import ieee80211n_1944_r12
    from ldpc_b200.linksim import LinkConfig, sim_run
    qc = ieee80211n_1944_r12()
    for snr_db in (5.0, 12.0):
        c = sim_run(wsim, LinkConfig(snr_db=snr_db, ofdm_size=64, iters=1, update="minsum", clamp_value=20.0, seed=9, channel="rayleigh"),
                    0, 4096).cpu().numpy()
        g = 10 ** (snr_db / 10) / 2
        pb = 0.5 * (1 - np.sqrt(g / (1 + g)))
        # 4096 codewords x 16 OFDM symbols = 65 536 independent fades: the BER estimate has the fades' variance, not the bits'
        assert abs(c[0] / c[3] - pb) < 6 * np.sqrt(pb / (4096 * 16)) + 2e-3, (snr_db, c[0] / c[3], pb)
    kw = dict(snr_db=9.0, ofdm_size=64, qbits=3, agc_mode=1, iters=6, update="minsum", clamp_value=20.0, seed=13, channel="rayleigh", compander=True)
    fused = sim_run(wsim, LinkConfig(**kw), 100, 1000).cpu().numpy()
    chain = sim_run(wsim, LinkConfig(force_unfused=True, **kw), 100, 1000).cpu().numpy()
    assert fused.tolist() == chain.tolist() and fused[0] > 0
    awgn = sim_run(wsim, LinkConfig(**{**kw, "channel": "awgn", "compander": False}), 100, 1000).cpu().numpy()
    assert awgn[0] != fused[0]                                          # the options do something
    # (iii) against the oracle chain
    gph = C.CGraph(qc.H)
    Ncw = 1 << 13
    rng = np.random.RandomState(77)
    enc = qc.encode(rng.randint(0, 2, (Ncw, qc.k)).astype(np.uint8)).astype(np.float64)
    for snr_db, qbits, comp in ((7.0, 0, False), (9.0, 3, True)):                      # BLER ~ 6 % and ~ 12 % (oracle probe)
        np.random.seed(int(snr_db))
        llr, _ = LO.framed_link_llrs(enc, snr_db, 64, qbits=qbits, channel="rayleigh", compander=comp)
        dec = C.decode(gph, llr.astype(np.float32), 10, 20.0, "minsum", want=("hard",))["hard"]
        m = LO.error_metrics(llr, dec, enc, qc.k)
        c = sim_run(wsim, LinkConfig(snr_db=snr_db, ofdm_size=64, qbits=qbits, agc_mode=1, iters=10, update="minsum", clamp_value=20.0,
                                     seed=21, channel="rayleigh", compander=comp), 0, Ncw).cpu().numpy()
        # uncoded errors are correlated inside an OFDM symbol (one fade): interval from the number of symbols
        p_o, p_g = m["uncoded_errs"] / m["bits"], c[0] / c[3]
        assert abs(p_o - p_g) < 6 * np.sqrt(2 * max(p_o, p_g) / (Ncw * 16)) + 1e-3, (snr_db, p_o, p_g)
        lo, hi = _wilson(m["frame_errs"], m["frames"])
        lo2, hi2 = _wilson(int(c[2]), int(c[4]))
        assert lo <= hi2 and lo2 <= hi, (snr_db, "bler", m["frame_errs"], int(c[2]))
