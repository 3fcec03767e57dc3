#!/usr/bin/env python3
"""Mint golden vectors from the UNMODIFIED reference and pin the oracle against it.

Runs only in the build container (needs /root/reference; the GPU box does not have it).
The reference is imported in place with the three shims of SURVEY.md appendix D; nothing
is copied.  Writes small fixtures to tests/golden/ and asserts, while doing so, that
oracle/bp_oracle.py and oracle/linksim_oracle.py reproduce the reference bit for bit on
the default (64,32) code and within tolerance on the dense n=1944 run.

    python oracle/make_golden.py            # ~3 min, ~8 GB RSS for the dense n=1944 model
"""
import json
import os
import pickle
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/pytorch"
np.complex = complex            # ofdm_functions.py:47,87 use the removed aliases
np.float = float
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, REF)
# The product package is NOT put on sys.path: its bp/ ofdm/ nn/ drop-ins are regular packages and would win over the
# reference's __init__-less directories of the same names whatever the path order.  The one module needed from it
# (the n=1944 prototype and encoder, plain numpy) is loaded by file name.

from bp.bp import BeliefPropagation                      # noqa: E402  (the reference)
from bp.parity import H as H_REF, G as G_REF             # noqa: E402
import ofdm.ofdm_functions as REFOF                      # noqa: E402
import bp_oracle as O                                    # noqa: E402
import linksim_oracle as LO                              # noqa: E402
import importlib.util                                    # noqa: E402
_spec = importlib.util.spec_from_file_location("_b200_codes", os.path.join(ROOT, "ldpc-sims_b200", "ldpc_b200", "codes.py"))
_codes = importlib.util.module_from_spec(_spec)
sys.modules["_b200_codes"] = _codes
_spec.loader.exec_module(_codes)
ieee80211n_1944_r12 = _codes.ieee80211n_1944_r12
for _m in (sys.modules["bp.bp"], sys.modules["bp.parity"], REFOF):      # the modules under test really are the reference's
    assert os.path.abspath(_m.__file__).startswith(REF + os.sep), _m.__file__

GOLD = os.path.join(ROOT, "tests", "golden")
os.makedirs(GOLD, exist_ok=True)
torch.set_num_threads(os.cpu_count())


def ref_forward(model, llr, clamp):
    """Reference forward + the pre-sigmoid marginal t (final_layer[0], bp.py:36-37)."""
    with torch.no_grad():
        L = torch.tensor(llr, dtype=torch.float)
        x = torch.zeros(L.shape[0], model.layer_size(), dtype=torch.float)
        prob = model(x, L, clamp).numpy()
        xx = x
        for layer in model.layers:
            xx = layer([xx, -L]).clamp(-clamp, clamp)
        t = model.final_layer[0]([xx, -L]).numpy()
    return prob, t, xx.numpy()


def default_code_cases():
    """(64,32) code, dense reference, bit-exact pin."""
    out = {}
    cases = []
    B = 512
    # synthetic gaussian LLRs at three scales
    for ci, (seed, scale, iters, clamp) in enumerate(
            [(11, 1.0, 3, 20), (12, 4.0, 5, 10), (13, 10.0, 10, 100), (14, 4.0, 10, 3), (15, 1.0, 1, 20)]):
        llr = (np.random.RandomState(seed).randn(B, 64) * scale).astype(np.float32)
        cases.append((f"gauss{ci}", llr, iters, clamp, dict(seed=seed, scale=scale)))
    # link-simulator LLRs through the reference's own gen_data (float64 -> f32 at the boundary)
    for snrdb in (0.0, 4.0, 8.0, 12.0):
        np.random.seed(int(100 + snrdb))
        bits = REFOF.create_bits(B * 32)
        enc = REFOF.encode_bits(bits, G_REF)
        tx = REFOF.modulate_bits(enc)
        _, _, rx_llrs, _ = REFOF.gen_data(tx, snrdb, 32)
        llr64 = rx_llrs.reshape(-1, 64)
        cases.append((f"link{int(snrdb)}dB", llr64.astype(np.float32), 3, 20,
                      dict(snrdb=snrdb, enc=np.packbits(enc.reshape(-1, 64).astype(np.uint8), axis=1))))
        # decode_bits itself (batching wrapper): batch 256, and a ragged batch of 300
        ref_bits = REFOF.decode_bits(llr64, H_REF, 3, 256, 20)
        ora_bits = O.decode_bits(llr64, H_REF, 3, 256, 20)
        assert np.array_equal(ref_bits, ora_bits)
        if snrdb == 4.0:
            ref_rag = REFOF.decode_bits(llr64, H_REF, 3, 300, 20)
            assert np.array_equal(ref_rag, O.decode_bits(llr64, H_REF, 3, 300, 20))
            out["link4dB_ragged300"] = np.packbits(ref_rag.astype(np.uint8), axis=1)
    # edge cases: all-zero LLRs (tie rule), huge LLRs (saturation), mixed
    edge = np.zeros((64, 64), np.float32)
    edge[16:32] = 1e4 * np.sign(np.random.RandomState(5).randn(16, 64)).astype(np.float32)
    edge[32:48] = np.random.RandomState(6).randn(16, 64).astype(np.float32) * 1e-6
    edge[48:] = np.random.RandomState(7).randn(16, 64).astype(np.float32) * 30
    cases.append(("edge", edge, 5, 20, {}))

    models = {}
    for name, llr, iters, clamp, extra in cases:
        if iters not in models:
            models[iters] = BeliefPropagation(H_REF, iters).eval()
        prob, t, x = ref_forward(models[iters], llr, clamp)
        o = O.bp_decode(H_REF, llr, iters, clamp)
        assert np.array_equal(o["x"], x) and np.array_equal(o["t"], t) and np.array_equal(o["prob"], prob), name
        hard = np.round(prob).astype(np.uint8)
        assert np.array_equal(o["hard"], hard)
        out[name + "_llr"] = llr
        out[name + "_iters"] = np.int64(iters)
        out[name + "_clamp"] = np.float64(clamp)
        out[name + "_t"] = t
        out[name + "_prob"] = prob[:64]
        out[name + "_x"] = x[:64]
        out[name + "_hard"] = np.packbits(hard, axis=1)
        out[name + "_syndrome"] = o["syndrome"]
        for k, v in extra.items():
            out[name + "_" + k] = v
        print(f"  default-code case {name:12s} iters={iters} clamp={clamp}: oracle == reference (x, t, prob, hard)")
    out["names"] = np.array([c[0] for c in cases])
    np.savez_compressed(os.path.join(GOLD, "bp_default_code.npz"), **out)
    # B = 4096 from the dense reference.  To keep the fixture small only the bit-exact outputs are kept for
    # all 4096 codewords (packed hard bits, syndrome weights) plus the marginals of the first 512; gaussian LLRs are
    # regenerated from their seed by the test, link LLRs (the reference's own gen_data draw) are stored.
    big = {}
    B4 = 4096
    big_cases = []
    for ci, (seed, scale, iters, clamp) in enumerate([(21, 4.0, 5, 10), (22, 10.0, 10, 100), (23, 2.0, 3, 20)]):
        llr = (np.random.RandomState(seed).randn(B4, 64) * scale).astype(np.float32)
        big_cases.append((f"gauss{ci}", llr, iters, clamp, dict(seed=np.int64(seed), scale=np.float64(scale))))
    for snrdb in (4.0, 8.0):
        np.random.seed(int(200 + snrdb))
        enc = REFOF.encode_bits(REFOF.create_bits(B4 * 32), G_REF)
        _, _, rx_llrs, _ = REFOF.gen_data(REFOF.modulate_bits(enc), snrdb, 32)
        big_cases.append((f"link{int(snrdb)}dB", rx_llrs.reshape(-1, 64).astype(np.float32), 3, 20, dict(store_llr=True)))
    for name, llr, iters, clamp, extra in big_cases:
        if iters not in models:
            models[iters] = BeliefPropagation(H_REF, iters).eval()
        prob, t, x = ref_forward(models[iters], llr, clamp)
        o = O.bp_decode(H_REF, llr, iters, clamp)
        assert np.array_equal(o["x"], x) and np.array_equal(o["t"], t) and np.array_equal(o["prob"], prob), name
        hard = np.round(prob).astype(np.uint8)
        assert np.array_equal(o["hard"], hard)
        # decode_bits (the batching wrapper) on the same LLRs agrees with the module's own rounding
        assert np.array_equal(REFOF.decode_bits(llr.astype(np.float64), H_REF, iters, 1024, clamp).astype(np.uint8), hard)
        big[name + "_iters"] = np.int64(iters); big[name + "_clamp"] = np.float64(clamp)
        big[name + "_hard"] = np.packbits(hard, axis=1); big[name + "_syndrome"] = o["syndrome"].astype(np.int16)
        big[name + "_t512"] = t[:512]
        if extra.pop("store_llr", False):
            big[name + "_llr"] = llr
        for k, v in extra.items():
            big[name + "_" + k] = v
        print(f"  default-code B=4096 case {name:10s} iters={iters} clamp={clamp}: oracle == reference (x, t, prob, hard)")
    big["names"] = np.array([c[0] for c in big_cases])
    np.savez_compressed(os.path.join(GOLD, "bp_default_code_4096.npz"), **big)


def wifi_dense_case():
    """n=1944 dense reference at tiny batch: tolerance pin (association order differs)."""
    code = ieee80211n_1944_r12()
    Hw = code.H.astype(np.int64)
    B, iters, clamp = 16, 10, 20                               # B = 16 from the dense reference
    rng = np.random.RandomState(2024)
    u = rng.randint(0, 2, size=(B, code.k)).astype(np.uint8)
    c = code.encode(u)
    ebn0 = 10 ** (2.0 / 10)
    sigma = np.sqrt(1.0 / (2 * 0.5 * ebn0))
    y = (1.0 - 2.0 * c) + sigma * rng.randn(B, code.n)
    llr = (-2.0 * y / sigma ** 2).astype(np.float32)          # log P1/P0
    t0 = time.time()
    model = BeliefPropagation(Hw, iters).eval()
    print(f"  dense n=1944 model built in {time.time() - t0:.1f}s")
    t0 = time.time()
    prob, t, x = ref_forward(model, llr, clamp)
    print(f"  dense n=1944 forward x2 in {time.time() - t0:.1f}s")
    del model
    o = O.bp_decode(Hw, llr, iters, clamp)
    hard = np.round(prob).astype(np.uint8)
    rel = np.abs(o["t"] - t) / np.maximum(np.abs(t), 1e-30)
    print(f"  oracle vs dense: hard equal={np.array_equal(o['hard'], hard)}  "
          f"max rel dt={rel.max():.3e}  frac(rel>1e-4)={np.mean(rel > 1e-4):.3e}  max abs dt={np.abs(o['t'] - t).max():.3e}")
    assert np.array_equal(o["hard"], hard)
    np.savez_compressed(os.path.join(GOLD, "bp_wifi1944_dense.npz"), llr=llr, t=t, prob=prob,
                        hard=np.packbits(hard, axis=1), info=u, codeword=np.packbits(c, axis=1),
                        iters=np.int64(iters), clamp=np.float64(clamp), oracle_t=o["t"],
                        x_first=x[:2])


def frontend_cases():
    out = {}
    N = 96                                      # codewords = OFDM symbols (n = 2*ofdm = 64)
    for N_dft in (32, 64, 128, 256):
        assert np.array_equal(REFOF.DFT(N_dft), LO.DFT(N_dft))
    assert np.array_equal(REFOF.DFTreal(32), LO.DFTreal(32))
    np.random.seed(7)
    bits = REFOF.create_bits(N * 32)
    np.random.seed(7)
    assert np.array_equal(bits, LO.create_bits(N * 32))
    enc = REFOF.encode_bits(bits, G_REF)
    assert np.array_equal(enc, LO.encode_bits(bits, G_REF))
    tx = REFOF.modulate_bits(enc)
    assert np.array_equal(tx, LO.modulate_bits(enc))
    out["bits"] = bits.astype(np.uint8); out["enc"] = enc.astype(np.uint8); out["tx_symbols"] = tx
    for snrdb in (4.0, 15.0):
        tag = f"snr{int(snrdb)}"
        np.random.seed(21)
        rx_signal, rx_symbols, rx_llrs, tx_signal = REFOF.gen_data(tx, snrdb, 32)
        np.random.seed(21)
        a, b, c, d = LO.gen_data(tx, snrdb, 32)
        assert all(np.array_equal(p, q) for p, q in zip((rx_signal, rx_symbols, rx_llrs, tx_signal), (a, b, c, d)))
        out[tag + "_rx_signal"] = rx_signal; out[tag + "_rx_symbols"] = rx_symbols
        out[tag + "_rx_llrs"] = rx_llrs; out[tag + "_tx_signal"] = tx_signal
        out[tag + "_noise"] = (rx_signal - tx_signal)      # informational
        for qbits, clip in ((1, 1.2), (3, 1.18), (3, 10.0), (5, 2.0), (6, 0.4)):
            q = REFOF.quantizer(rx_signal, qbits, clip)
            assert np.array_equal(q, LO.quantizer(rx_signal, qbits, clip))
            out[f"{tag}_quant_b{qbits}_c{clip}"] = q
        for qbits, clipdb in ((1, 0.0), (3, 0.0), (3, 5.0), (5, 10.0)):
            cr = np.power(10, clipdb / 10)
            r = REFOF.gen_qdata(rx_signal, snrdb, qbits, cr, 32)
            o = LO.gen_qdata(rx_signal, snrdb, qbits, cr, 32)
            assert all(np.array_equal(p, q) for p, q in zip(r, o))
            out[f"{tag}_qdata_b{qbits}_c{int(clipdb)}_signal"] = r[0]
            out[f"{tag}_qdata_b{qbits}_c{int(clipdb)}_llrs"] = r[2]
        # the inline AGC front end of evaluate_quantized_snr.py:96-133, run from the script's own lines
        qllr, qresc = LO.agc_quantized_frontend(rx_signal, snrdb, 3, 1.0, 32, agc_clip=10)
        snr_single = np.power(10, snrdb / 10)
        factor = 10 / (.5 * (1 + 1 / snr_single)) * 1.0
        scaled = (factor * rx_signal.reshape((-1, 32)).T).T.reshape((1, -1))
        qref = REFOF.quantizer(scaled, 3, 10).reshape((-1, 32)).T / factor
        assert np.array_equal(qref, qresc)
        llr_ref, _ = REFOF.demodulate_signal(qref.T.reshape((1, -1)), 32, snr_single)
        assert np.array_equal(llr_ref, qllr)
        out[tag + "_agc_llrs"] = qllr
        # metrics (evaluate_quantized_snr.py:169-188)
        L = rx_llrs.reshape(-1, 64); E = enc.reshape(-1, 64)
        dec = REFOF.decode_bits(L, H_REF, 3, 32, 20)
        cb = (np.sign(L) + 1) // 2
        met = LO.error_metrics(L, dec, E, 32)
        assert met["uncoded_errs"] == int(np.sum(np.abs(cb - E)))
        assert abs(met["info_errs"] / met["info_bits"] - np.mean(np.abs(dec[:, 0:32] - E[:, 0:32]))) < 1e-15
        assert abs(met["frame_errs"] / met["frames"] - np.mean(np.sign(np.sum(np.abs(dec - E), axis=1)))) < 1e-15
        out[tag + "_metrics"] = np.array([met["uncoded_errs"], met["info_errs"], met["frame_errs"]], dtype=np.int64)
        print(f"  front end @ {snrdb} dB: oracle == reference; metrics {met}")
    np.savez_compressed(os.path.join(GOLD, "frontend.npz"), **out)


def published_ber():
    p = os.path.join(REF, "outputs/ber/20191203-191640_tx=20191203-162534_quantized.pkl")
    with open(p, "rb") as f:
        d = pickle.load(f)
    js = {k: np.asarray(v).astype(float).tolist() for k, v in d.items()}
    js["_source"] = "pytorch/outputs/ber/20191203-191640_tx=20191203-162534_quantized.pkl (3 BP iterations, clamp 20, 2^15 codewords/point)"
    with open(os.path.join(GOLD, "published_ber.json"), "w") as f:
        json.dump(js, f, indent=1)
    print("  published BER pickle ->", list(js.keys()))


if __name__ == "__main__":
    t0 = time.time()
    print("default code (bit-exact pin):"); default_code_cases()
    print("front end (bit-exact pin):"); frontend_cases()
    print("published BER:"); published_ber()
    print("n=1944 dense reference (tolerance pin):"); wifi_dense_case()
    print(f"done in {time.time() - t0:.0f}s; fixtures in {GOLD}")
