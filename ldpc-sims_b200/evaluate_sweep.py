#!/usr/bin/env python3
"""BER/FER-vs-SNR sweep on the GPU(s): the per-SNR loop of the reference's evaluate scripts
(evaluate_quantized_snr.py:91-214) as one fused simulation per point, codewords sharded by
batch over the ranks, ONE all-reduce of the int64 counter matrix at the end.

    python ldpc-sims_b200/evaluate_sweep.py --code wifi --snr 0:6:1 --codewords 524288
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 ldpc-sims_b200/evaluate_sweep.py ...

Writes the same pickle keys the reference's plots.py reads (snrdb, uncoded_ber, coded_ber,
coded_bler and, with --qbits, the *_quantized variants).

    python ldpc-sims_b200/evaluate_sweep.py --code default --full --qbits 3 --snr 5:15:1 --update sp --clamp 100 \
        --nn-checkpoint outputs/model/....pth --out outputs/ber/run.pkl --resume /tmp/run_state

--full reproduces the WHOLE result set of evaluate_quantized_snr.py:91-214 (traditional, *_nn, *_quantized, wmse_*)
on one noise realisation per codeword - every key plots.py:11-27 loads; --resume checkpoints each finished SNR point
per rank and skips it on restart."""
import argparse
import json
import os
import pickle
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ldpc_b200.codes import ieee80211n_1944_r12, peg_64_32            # noqa: E402
from ldpc_b200.decoder import LdpcCode                                # noqa: E402
from ldpc_b200.linksim import LinkConfig, attach_generator, rates, sweep   # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--code", default="wifi", choices=["wifi", "default"])
    ap.add_argument("--snr", default="0:6:1", help="lo:hi:step in dB (per-subcarrier Es/N0, ofdm_functions.py:110)")
    ap.add_argument("--codewords", type=int, default=1 << 19, help="codewords per SNR point (all ranks together)")
    ap.add_argument("--ofdm", type=int, default=0, help="OFDM size (default: 32 for the default code, 64 for wifi)")
    ap.add_argument("--qbits", type=int, default=0)
    ap.add_argument("--agc-mode", type=int, default=1)
    ap.add_argument("--clip-ratio", type=float, default=1.0)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--update", default="minsum")
    ap.add_argument("--clamp", type=float, default=20.0)
    ap.add_argument("--param", type=float, default=1.0)
    ap.add_argument("--seed", type=int, default=1234)
    ap.add_argument("--out", default="")
    ap.add_argument("--full", action="store_true", help="traditional + quantized (+ NN) links per point, all plots.py keys")
    ap.add_argument("--nn-checkpoint", default="", help=".pth saved by the reference (LLRestimator_withSNR) or .npz with w_module.* arrays")
    ap.add_argument("--resume", default="", help="state file prefix for resumable --full sweeps")
    a = ap.parse_args()

    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"          # keep stdout to the single JSON line
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if a.code == "wifi":
        qc = ieee80211n_1944_r12()
        code = attach_generator(LdpcCode(qc.H, qc_Z=81, qc_proto=qc.proto))
        n, k, ofdm = qc.n, qc.k, a.ofdm or 64
    else:
        H, G = peg_64_32()
        code = attach_generator(LdpcCode(H), G)
        n, k, ofdm = 64, 32, a.ofdm or 32
    lo, hi, st = (float(v) for v in a.snr.split(":"))
    snrdb = np.arange(lo, hi + 1e-9, st)
    cfgs = [LinkConfig(snr_db=float(s), ofdm_size=ofdm, qbits=a.qbits, agc_mode=a.agc_mode, clip_ratio=a.clip_ratio,
                       iters=a.iters, update=a.update, clamp_value=a.clamp, param=a.param, seed=a.seed) for s in snrdb]
    if a.full:
        from ldpc_b200.linksim import evaluate_full, results_dict
        demapper = None
        if a.nn_checkpoint:
            from nn.llr import LLRestimator_withSNR
            if a.nn_checkpoint.endswith(".npz"):
                g = np.load(a.nn_checkpoint)
                state = {key[2:]: torch.tensor(g[key]) for key in g.files if key.startswith("w_")}
            else:
                state = torch.load(a.nn_checkpoint, map_location="cpu", weights_only=False)["model_state_dict"]
            model = torch.nn.DataParallel(LLRestimator_withSNR(ofdm))
            model.load_state_dict(state)
            demapper = model.module.eval()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        counters, wm = evaluate_full(code, cfgs, demapper, a.codewords, rank=rank, world=world, state_path=a.resume or None)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if rank == 0:
            res = results_dict(snrdb, counters, wm, n, k)
            if a.out:
                with open(a.out, "wb") as f:
                    pickle.dump(res, f)
            print(json.dumps({"world": world, "seconds": dt, "counters": counters.tolist(),
                              **{key: np.asarray(val).tolist() for key, val in res.items()}}))
        if world > 1:
            dist.destroy_process_group()
        return
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    counters = sweep(code, cfgs, a.codewords, rank=rank, world=world)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if rank == 0:
        r = rates(counters, n, k)
        suffix = "_quantized" if a.qbits else ""
        res = {"snrdb": snrdb, **{key + suffix: val for key, val in r.items()}}
        if a.out:
            with open(a.out, "wb") as f:
                pickle.dump(res, f)
        print(json.dumps({"snrdb": snrdb.tolist(), "counters": counters.tolist(), "world": world, "seconds": dt,
                          "codewords_per_s": float(a.codewords * len(snrdb) / dt),
                          **{key + suffix: val.tolist() for key, val in r.items()}}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
