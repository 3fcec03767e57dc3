#!/usr/bin/env python3
"""Mint tests/golden/nn_demapper.npz from the UNMODIFIED reference (build container only).

Follows evaluate_quantized_snr.py:41-173 with the script's own constants (qbits 3, agc_clip 10,
clip_ratio 1, the checkpoint it loads) at a small sample count: reference front end -> AGC
quantizer -> input_samples -> reference LLRestimator_withSNR (CPU, fp32) -> reference decode_bits.
Stores the checkpoint's weights (fp32, the only copy that travels to the GPU box), the inputs,
the reference's LLR estimates and decoded bits, and pins oracle/nn_oracle.py against them.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/pytorch"
np.complex = complex
np.float = float
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

from bp.parity import H, G                                # noqa: E402  (the reference)
import ofdm.ofdm_functions as F                           # noqa: E402
from nn.llr import LLRestimator, LLRestimator_withSNR     # noqa: E402
import nn_oracle as NO                                    # noqa: E402

CKPT = "outputs/model/20191214-172134_qbits=3_clipdb=0_snrlow=5_snrhigh=15_lr=0.1.pth"   # evaluate_quantized_snr.py:25,67


def main():
    ofdm_size, qbits, agc_clip, clip_ratio = 32, 3, 10, 1.0
    ck = torch.load(os.path.join(REF, CKPT), map_location="cpu", weights_only=False)
    model = torch.nn.DataParallel(LLRestimator_withSNR(ofdm_size))
    model.load_state_dict(ck["model_state_dict"])
    model.eval()
    out = {"w_" + k: v.numpy() for k, v in ck["model_state_dict"].items()}
    names = []
    for snrdb, nsym, iters, clamp in ((15.0, 512, 10, 100), (5.0, 256, 10, 100), (10.0, 256, 5, 20)):
        np.random.seed(int(1000 + snrdb))
        bits = F.create_bits(nsym * ofdm_size)
        enc = F.encode_bits(bits, G)
        tx = F.modulate_bits(enc)
        snr = np.power(10, snrdb / 10)
        rx_signal, _, rx_llrs, _ = F.gen_data(tx, snrdb, ofdm_size)
        rs = rx_signal.reshape((-1, ofdm_size)).T
        sigma_rx = .5 * (1 + 1 / snr)                        # evaluate_quantized_snr.py:103 (a variance, used as amplitude)
        factor = agc_clip / sigma_rx * clip_ratio
        q = F.quantizer((factor * rs).T.reshape((1, -1)), qbits, agc_clip)
        qr = q.reshape((-1, ofdm_size)).T / factor
        x = NO.nn_input_samples(qr, snr)
        with torch.no_grad():
            y = model.module(torch.tensor(x, dtype=torch.float)).numpy()
        yo = NO.mlp_forward(ck["model_state_dict"], x)
        err = np.max(np.abs(yo - y)) / np.max(np.abs(y))
        assert err < 1e-5, err
        dec = F.decode_bits(y.astype(np.float64), H, iters, 256, clamp)
        tag = f"snr{int(snrdb)}"
        names.append(tag)
        out[tag + "_x"] = x.astype(np.float32)
        out[tag + "_llr"] = y.astype(np.float32)
        out[tag + "_bits"] = np.packbits(dec.astype(np.uint8), axis=1)
        out[tag + "_enc"] = np.packbits(enc.reshape(-1, 2 * ofdm_size).astype(np.uint8), axis=1)
        out[tag + "_rx_llr"] = rx_llrs.reshape(-1, 2 * ofdm_size).astype(np.float32)
        out[tag + "_meta"] = np.array([snrdb, iters, clamp], np.float64)
        ber = np.mean(np.abs(dec[:, :32] - enc.reshape(-1, 64)[:, :32]))
        print(f"  {tag}: {nsym} OFDM symbols, oracle-vs-reference max err {err:.2e} of scale, coded BER (NN LLRs) {ber:.4f}")
    # the plain LLRestimator chain (fft_layer without bias / tanh, nn/llr.py:46-52): pinned here at mint time
    # on a seeded random initialisation (no fixture: the weights would be another 2.4 MB)
    torch.manual_seed(0)
    plain = LLRestimator(ofdm_size, 10.0).eval()
    xp = torch.randn(300, 2 * ofdm_size)
    with torch.no_grad():
        yp = plain(xp).numpy()
    yo = NO.mlp_forward({k: v.detach().numpy() for k, v in plain.state_dict().items()}, xp.numpy(), layers=NO.PLAIN_LAYERS, acts=NO.PLAIN_ACTS)
    err = np.max(np.abs(yo - yp)) / np.max(np.abs(yp))
    assert err < 1e-5, err
    print(f"  plain LLRestimator chain: oracle-vs-reference max err {err:.2e} of scale")
    out["names"] = np.array(names)
    p = os.path.join(ROOT, "tests", "golden", "nn_demapper.npz")
    np.savez_compressed(p, **out)
    print("wrote", p, os.path.getsize(p), "bytes")


if __name__ == "__main__":
    main()
