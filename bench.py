#!/usr/bin/env python3
"""bench.py - decoded information Gbit/s of the B200-native LDPC decoder on the headline
workload of BASELINE.json: IEEE 802.11n n=1944 R=1/2, 10 fixed min-sum iterations,
1M-codeword batch per GPU (configs[2]); LLRs resident in HBM for `value`, host buffers for
`e2e`.

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # CPU restatement of the reference on host cores

One "step" = one pass of the decoder over one batch of synthetic LLRs.  Prints ONE JSON line.
"""
import argparse
import ctypes
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "ldpc-sims_b200"))

N_CODE, K_CODE, E_CODE = 1944, 972, 6966
METRIC = "decoded info Gbps/GPU at n=1944 r=1/2, 10 iters"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--codewords", type=int, default=1_000_000, help="codewords per GPU per step")
    ap.add_argument("--e2e-codewords", type=int, default=131072, help="codewords per e2e step (host buffers)")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--update", default="minsum", choices=["minsum", "sp", "nms", "oms"])
    ap.add_argument("--clamp", type=float, default=20.0)
    ap.add_argument("--ebn0", type=float, default=2.0)
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target duration of the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-f16", action="store_true")
    ap.add_argument("--no-sp", action="store_true", help="skip the sum-product measurement")
    ap.add_argument("--no-nn", action="store_true", help="skip the NN-demapper link (BASELINE.json configs[4])")
    ap.add_argument("--no-train", action="store_true", help="skip the weighted-BP training step (SURVEY 8f rank 2)")
    ap.add_argument("--no-numa-bind", action="store_true", help="do not pin the rank to the NUMA node of its GPU")
    ap.add_argument("--e2e-chunk", type=int, default=16384, help="codewords per H2D / decode / D2H chunk of the host pipeline")
    ap.add_argument("--no-family", action="store_true", help="skip the per-code table of the 802.11n family")
    ap.add_argument("--no-sweep", action="store_true", help="skip the sharded BER/FER sweep with its all-reduce (BASELINE.json configs[3])")
    ap.add_argument("--sweep-codewords", type=int, default=1 << 19, help="codewords per SNR point of the sweep, TOTAL over all GPUs (strong scaling)")
    ap.add_argument("--nn-symbols", type=int, default=1 << 20, help="OFDM symbols per GPU of the NN-demapper measurement")
    return ap.parse_args()


def workload_name(a):
    return f"802.11n n=1944 r=1/2 Z=81, {a.update} x{a.iters} flooding, clamp {a.clamp:g}, fp32, AWGN Eb/N0={a.ebn0:g} dB"


def make_llr_numpy(ncw, ebn0_db, seed):
    """BPSK/AWGN LLRs (log P1/P0) of random codewords, numpy (for the CPU arm)."""
    import numpy as np
    from ldpc_b200.codes import ieee80211n_1944_r12
    qc = ieee80211n_1944_r12()
    rng = np.random.RandomState(seed)
    u = rng.randint(0, 2, (min(ncw, 256), qc.k)).astype(np.uint8)
    c = qc.encode(u)
    c = np.tile(c, ((ncw + c.shape[0] - 1) // c.shape[0], 1))[:ncw]
    sigma = (1.0 / (2 * 0.5 * 10 ** (ebn0_db / 10))) ** 0.5
    y = (1.0 - 2.0 * c) + sigma * rng.standard_normal((ncw, qc.n)).astype(np.float32)
    return (-2.0 * y / sigma ** 2).astype(np.float32)


# ------------------------------------------------------------------------------------------
# CPU arm: the oracle port (plain C restatement of the reference algorithm) on host cores
# ------------------------------------------------------------------------------------------
def cpu_sample(a, seconds):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import c_oracle as C
    from ldpc_b200.codes import ieee80211n_1944_r12
    g = C.CGraph(ieee80211n_1944_r12().H)
    cores = C.max_threads()
    llr = make_llr_numpy(64 * cores, a.ebn0, 99)
    t0 = time.perf_counter()
    C.decode(g, llr, a.iters, a.clamp, a.update, 1.0, want=("hard",))
    rate = llr.shape[0] / (time.perf_counter() - t0)                 # calibration
    ncw = int(max(64 * cores, min(rate * seconds, 4_000_000)))
    llr = make_llr_numpy(ncw, a.ebn0, 100)
    return C, g, llr, cores


def run_reference(a):
    """--impl reference: the reference is pure Python and is not on the GPU box, so its CPU
    implementation is represented by the oracle port (oracle/ldpc_oracle.c, all host threads)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    per_step = max(1.0, min(a.cpu_seconds, 60.0 / max(1, a.steps + a.warmup)))
    C, g, llr, cores = cpu_sample(a, per_step)
    ncw = llr.shape[0]
    for _ in range(a.warmup):
        C.decode(g, llr, a.iters, a.clamp, a.update, 1.0, want=("hard",))
    t0 = time.perf_counter()
    for _ in range(a.steps):
        C.decode(g, llr, a.iters, a.clamp, a.update, 1.0, want=("hard",))
    dt = time.perf_counter() - t0
    gbps = ncw * a.steps * K_CODE / dt / 1e9
    sample = f"{ncw} codewords/step of the same workload, C port of the reference algorithm, {cores} pthreads"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": gbps, "unit": "Gbit/s", "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": dt / a.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a), "codewords_per_step": ncw},
        "cpu_baseline": {"value": gbps, "unit": "Gbit/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": gbps, "unit": "Gbit/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------
# clocks sampler (NVML)
# ------------------------------------------------------------------------------------------
class Clocks:
    def __init__(self, index):
        self.samples, self.reasons, self.stop = [], set(), False
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None
        self.th = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40,
                 "sw_thermal_slowdown": 0x20, "hw_power_brake": 0x80}
        while not self.stop:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.02)

    def __enter__(self):
        if self.nv:
            self.th.start()
        return self

    def __exit__(self, *exc):
        self.stop = True
        if self.nv:
            self.th.join(timeout=1)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def bind_to_gpu_numa(index):
    """Pin this rank's host threads (and therefore its first-touch pinned allocations) to the NUMA node its GPU hangs off:
    with 8 ranks on a two-socket box a staging buffer on the far socket crosses the inter-socket link on every H2D copy.
    Returns a small record for the JSON line; never fails the run."""
    rec = {"bound": False}
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dom, rest = bus.split(":", 1)
        sysfs = f"/sys/bus/pci/devices/{dom[-4:].lower()}:{rest.lower()}/numa_node"
        node = int(open(sysfs).read().strip())
        rec["pci"] = bus
        rec["numa_node"] = node
        if node >= 0:
            cpus = set()
            for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
            allowed = os.sched_getaffinity(0)
            use = cpus & allowed
            if use:
                os.sched_setaffinity(0, use)
                rec["bound"] = True
                rec["cpus"] = len(use)
    except Exception as e:                                  # no NVML / no sysfs (containers): run unbound
        rec["error"] = str(e)[:80]
    return rec


def bench_nn(a, dev, world, barrier, peaks):
    """evaluate_quantized_snr.py:91-188 with the reference's checkpoint (weights from tests/golden/nn_demapper.npz):
    front end -> received time samples -> MLP (tcgen05, fp32-equivalent) -> sum-product decoder -> counters."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from ldpc_b200.codes import peg_64_32
    from ldpc_b200.decoder import LdpcCode
    from ldpc_b200.linksim import LinkConfig, attach_generator, sim_run_nn
    from ldpc_b200.mlp import NativeMLP
    g = np.load(os.path.join(ROOT, "tests", "golden", "nn_demapper.npz"))
    names = ("hidden1", "hidden2", "hidden3", "final")
    W = [g[f"w_module.{n}.weight"] for n in names]
    net = NativeMLP(W, [g[f"w_module.{n}.bias"] for n in names], splits=2, device=dev)
    H, G = peg_64_32()
    code = attach_generator(LdpcCode(H, device=dev), G)
    cfg = LinkConfig(snr_db=15.0, ofdm_size=32, qbits=3, agc_mode=1, iters=10, update="sp", clamp_value=100.0, seed=99)
    S = a.nn_symbols
    x = torch.randn(S, 65, device=dev) * 0.7
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(fn, reps):
        fn()
        barrier()
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    ms_mlp = timed(lambda: net(x), 3)                   # default schedule: the single-launch L2-resident chain on CTA pairs
    net.set_mode("chain")
    ms_mlp_chain = timed(lambda: net(x), 3)             # the same chain on single SMs (cta_group::1)
    net.set_mode("per_layer")
    ms_mlp_pl = timed(lambda: net(x), 3)                # round-1 schedule: one launch per layer, activations through HBM
    net.set_mode("auto")
    counters = torch.zeros(5, dtype=torch.int64, device=dev)
    ms_link = timed(lambda: sim_run_nn(code, cfg, net, 0, S, counters), 2)
    # the reference's default code alone (BASELINE.json configs[0]/[1]): register-resident kernel, LLRs resident in HBM
    from ldpc_b200.linksim import sim_generate, decode_count
    cwp, llr_d = sim_generate(code, cfg, 0, S)
    cfg_ms = LinkConfig(snr_db=15.0, ofdm_size=32, qbits=3, agc_mode=1, iters=10, update="minsum", clamp_value=100.0, seed=99)
    c2 = torch.zeros(5, dtype=torch.int64, device=dev)
    ms_sp = timed(lambda: decode_count(code, llr_d, cwp, cfg, c2), 3)
    ms_ms = timed(lambda: decode_count(code, llr_d, cwp, cfg_ms, c2), 3)
    flops = 2.0 * sum(w.shape[0] * w.shape[1] for w in W) * S
    tf_peak = float(peaks.get("bf16_tflops", 2250.0))
    c = counters.cpu().numpy().astype(np.float64)
    return {"workload": "LLRestimator_withSNR(32) 65-512-512-512-64 tanh (reference checkpoint) + (64,32) code, sum-product x10, 15 dB, 3-bit ADC",
            "ofdm_symbols_per_gpu": S, "link_ms": ms_link, "link_symbols_per_s": S * world / (ms_link * 1e-3),
            "link_info_gbps": S * world * 32 / (ms_link * 1e-3) / 1e9, "mlp_ms": ms_mlp, "mlp_ms_single_sm_chain": ms_mlp_chain, "mlp_ms_per_layer_launches": ms_mlp_pl,
            "mlp_schedule": "one cooperative launch per 303 104-row chunk: groups of 4 SMs (two cta_group::2 pairs) carry 256-row blocks through all "
                            "layers, activation planes handed over through L2 (DRAM traffic 3.8 KB/row on pairs against "
                            "10.6 KB/row of the per-layer launches; ncu: profiles/r02_mlp_experiments.md)",
            "mlp_fp32_equivalent_tflops": flops / (ms_mlp * 1e-3) / 1e12,
            "roofline": {"bound": "tensor", "achieved": 3 * flops / (ms_mlp * 1e-3) / 1e12, "peak": tf_peak, "unit": "TFLOP/s",
                         "frac": 3 * flops / (ms_mlp * 1e-3) / 1e12 / tf_peak, "traffic": None,
                         "note": "16-bit tcgen05.mma flops issued: 3 plane pairs per fp32 product (2 exact binary16 planes per operand)"},
            "coded_ber_nn": c[1] / max(c[4] * 32, 1), "uncoded_ber_nn": c[0] / max(c[3], 1), "gpu_launches_per_chunk": 5,
            "default_code_decode": {"codewords_per_gpu": S, "kernel": ("generic", "qc", "tiny", "qc_rt")[code.kernel],
                                    "sum_product_x10_info_gbps": S * world * 32 / (ms_sp * 1e-3) / 1e9, "sum_product_ms": ms_sp,
                                    "min_sum_x10_info_gbps": S * world * 32 / (ms_ms * 1e-3) / 1e9, "min_sum_ms": ms_ms,
                                    "note": "(64,32) code of bp/parity.py, one thread per codeword, decode + fused counters, LLRs resident in HBM"}}


def bench_sweep(a, dev, world, rank, barrier):
    """BASELINE.json configs[3] (evaluate_quantized_snr.py:91-214 as one sharded job): 7 SNR points x 2^19 codewords of the
    n=1944 code (1.02e9 coded bits per point) over 64-point OFDM with the 3-bit ADC front end, min-sum x10.  The TOTAL is
    fixed (strong scaling): rank r simulates codewords shard_range(total, r, world) of every point with the single-launch
    simulator (random bits -> ... -> counters in one kernel), then ONE NCCL all-reduce of the int64 [7,5] counter matrix -
    inside the timed region.  The reduced counters must equal a 1-GPU run of the whole sweep (recomputed on rank 0 outside
    the timed region): the draws depend on (seed, global codeword index) only."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from ldpc_b200.codes import ieee80211n_1944_r12
    from ldpc_b200.decoder import LdpcCode
    from ldpc_b200.linksim import LinkConfig, attach_generator, rates, shard_range, sim_run
    qc = ieee80211n_1944_r12()
    code = attach_generator(LdpcCode(qc.H, qc_Z=81, qc_proto=qc.proto, device=dev))
    snrs = [float(s) for s in range(7)]
    cfgs = [LinkConfig(snr_db=s, ofdm_size=64, qbits=3, agc_mode=1, agc_clip=10.0, clip_ratio=1.0, iters=10, update="minsum",
                       clamp_value=20.0, seed=2026) for s in snrs]
    total = a.sweep_codewords
    first, count = shard_range(total, rank, world)
    counters = torch.zeros(len(cfgs), 5, dtype=torch.int64, device=dev)
    ws = torch.empty(16, dtype=torch.uint8, device=dev)            # the single-launch simulator needs no workspace
    stream = torch.cuda.current_stream(dev)

    def one_sweep():
        counters.zero_()
        for i, cfg in enumerate(cfgs):
            if count:
                sim_run(code, cfg, first, count, counters[i], ws)
        if world > 1:
            dist.all_reduce(counters, op=dist.ReduceOp.SUM)
    one_sweep()                                                    # warm-up (kernel load, NCCL channel set-up)
    one_sweep()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    e0.record(stream)
    for _ in range(reps):
        one_sweep()
    e1.record(stream)
    barrier()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    got = counters.cpu().numpy().copy()
    # the all-reduce alone, ranks aligned by a barrier first (its own latency, not the wait for the slowest rank)
    ar_ms = None
    if world > 1:
        tmp = counters.clone()
        lat = []
        for _ in range(5):
            barrier()
            e0.record(stream)
            dist.all_reduce(tmp, op=dist.ReduceOp.SUM)
            e1.record(stream)
            torch.cuda.synchronize()
            lat.append(e0.elapsed_time(e1))
        tt = torch.tensor([sorted(lat)[len(lat) // 2]], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ar_ms = float(tt.item())
    # exactness: the whole sweep on ONE GPU gives the same integers
    same = None
    if rank == 0:
        ref = torch.zeros(len(cfgs), 5, dtype=torch.int64, device=dev)
        for i, cfg in enumerate(cfgs):
            sim_run(code, cfg, 0, total, ref[i], ws)
        same = bool(np.array_equal(ref.cpu().numpy(), got))
        assert same, "sharded sweep counters differ from the 1-GPU counters"
    r = rates(got, qc.n, qc.k)
    bits = float(total) * qc.n * len(cfgs)
    return {"workload": "802.11n n=1944 r=1/2 over OFDM-64 + 3-bit ADC (script AGC), min-sum x10, Es/N0 0..6 dB step 1, single-launch simulator",
            "snr_db": snrs, "codewords_per_point_total": total, "coded_bits_per_point": total * qc.n, "scaling": "strong",
            "seconds": ms * 1e-3, "ms": ms, "coded_gbit_per_s": bits / (ms * 1e-3) / 1e9, "gpu_launches": len(cfgs),
            "collective": "one NCCL all_reduce(SUM) of int64[7,5] inside the timed region" if world > 1 else "none (1 GPU)",
            "allreduce_ms_ranks_aligned": ar_ms, "counters_equal_1gpu_run": same,
            "uncoded_ber": [float(v) for v in r["uncoded_ber"]], "coded_ber": [float(v) for v in r["coded_ber"]],
            "coded_bler": [float(v) for v in r["coded_bler"]], "counters": got.tolist()}


def bench_quantized_link(a, dev, world, rank, barrier):
    """BASELINE.json configs[1] (evaluate_quantized_snr.py:14-23,91-188 at one SNR): the reference's default (64,32) code over
    32-point OFDM, 3-bit ADC with the script's AGC (clip 10), 15 dB, sum-product x10, clamp 100 - 2^24 codewords in total,
    sharded by batch, one all-reduce of the 5 counters.  The survey's CPU run of the reference script at 2^14 codewords gave
    uncoded 1.35e-2, coded 1.9e-4, BLER 2.1e-2."""
    import torch
    import torch.distributed as dist
    from ldpc_b200.codes import peg_64_32
    from ldpc_b200.decoder import LdpcCode
    from ldpc_b200.linksim import LinkConfig, attach_generator, rates, shard_range, sim_run
    H, G = peg_64_32()
    code = attach_generator(LdpcCode(H, device=dev), G)
    cfg = LinkConfig(snr_db=15.0, ofdm_size=32, qbits=3, agc_mode=1, agc_clip=10.0, clip_ratio=1.0, iters=10, update="sp",
                     clamp_value=100.0, seed=15)
    total = 1 << 24
    first, count = shard_range(total, rank, world)
    c = torch.zeros(5, dtype=torch.int64, device=dev)
    sim_run(code, cfg, first, min(count, 1 << 20), c)                # warm-up
    c.zero_()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sim_run(code, cfg, first, count, c)
    if world > 1:
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    r = rates(c.cpu().numpy(), 64, 32)
    return {"workload": "default (64,32) code, OFDM-32, 3-bit ADC (script AGC, clip 10), 15 dB, sum-product x10, clamp 100", "codewords_total": total,
            "ms": ms, "codewords_per_s": total / (ms * 1e-3), "uncoded_ber": float(r["uncoded_ber"]), "coded_ber": float(r["coded_ber"]),
            "coded_bler": float(r["coded_bler"]), "reference_script_probe_2e14_codewords": {"uncoded_ber": 1.35e-2, "coded_ber": 1.87e-4, "coded_bler": 2.09e-2}}


def bench_family(a, dev, world, barrier, headline_updates_per_s):
    """every IEEE 802.11n prototype on its code-compiled kernel - decoded information rate and directed
    edge-updates per second (the roofline quantity) next to the headline code's, same decoder settings, LLRs in HBM."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from ldpc_b200.codes import WIFI_LENGTHS, WIFI_RATES, ieee80211n
    from ldpc_b200.decoder import LdpcCode
    from ldpc_b200 import _native as N
    lib = N.lib()
    stream = torch.cuda.current_stream(dev)
    out = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for n in WIFI_LENGTHS:
        for rate in WIFI_RATES:
            qc = ieee80211n(n, rate)
            code = LdpcCode(qc.H, qc_Z=qc.Z, qc_proto=qc.proto, device=dev)
            B = int(2.0e8 // n)                                   # ~0.8 GB of f32 LLRs per launch: larger than L2
            R = qc.k / qc.n
            sigma = (1.0 / (2 * R * 10 ** (0.1 * (1.5 + 2.5 * R)))) ** 0.5
            cw = torch.as_tensor(qc.encode(np.random.RandomState(n).randint(0, 2, (64, qc.k)).astype(np.uint8))).to(dev)
            llr = torch.empty(B, n, dtype=torch.float32, device=dev)
            for s0 in range(0, B, 32768):
                s1 = min(B, s0 + 32768)
                y = (1.0 - 2.0 * cw[torch.arange(s0, s1, device=dev) % 64].float()) + sigma * torch.randn(s1 - s0, n, device=dev)
                llr[s0:s1] = -2.0 * y / sigma ** 2
            post = torch.empty(B, n, dtype=torch.float32, device=dev)
            packed = torch.empty(B, code.packed_bytes, dtype=torch.uint8, device=dev)

            def step():
                N.check(lib.ldpc_decode(code._h, llr.data_ptr(), N.F32, B, a.iters, N.UPDATE_IDS[a.update], a.clamp, 1.0, None, None,
                                        post.data_ptr(), None, packed.data_ptr(), None, None, ctypes.c_void_p(stream.cuda_stream)))
            step(); step()
            barrier()
            e0.record(stream)
            for _ in range(3):
                step()
            e1.record(stream)
            barrier()
            t = torch.tensor([e0.elapsed_time(e1) / 3], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            E = int(qc.H.sum())
            ups = B / (ms * 1e-3) * 2 * E * a.iters
            out.append({"code": qc.name, "Z": qc.Z, "k": qc.k, "edges": E, "kernel": ("generic", "qc", "tiny", "qc_rt", "qc_tma")[code.kernel],
                        "codewords_per_gpu": B, "ms": ms, "info_gbps": B * world * qc.k / (ms * 1e-3) / 1e9,
                        "edge_updates_per_s": ups * world,
                        "edge_update_rate_vs_headline": ups / headline_updates_per_s if headline_updates_per_s else None})
            del llr, post, packed, code
    return out


def bench_train(dev, world, barrier):
    """ofdm/ofdm_nn.py:281-343 in miniature: BCE through the weighted decoder on the default code, loss.backward() on
    the native sparse backward, at the reference's minibatch (512) and at a GPU-sized batch."""
    import torch
    import torch.distributed as dist
    from bp.bp import BeliefPropagation
    from bp.parity import H
    res = {}
    for tag, B, iters in (("minibatch_512_iters_3", 512, 3), ("batch_65536_iters_5", 65536, 5)):
        m = BeliefPropagation(H, iters).to(dev)
        llr = (torch.randn(B, 64, device=dev) * 2).requires_grad_(True)
        y = (torch.rand(B, 64, device=dev) > 0.5).float()

        def step():
            m.zero_grad()
            llr.grad = None
            loss = torch.nn.functional.binary_cross_entropy(m(None, llr, 20.0).clamp(1e-6, 1 - 1e-6), y)
            loss.backward()
        for _ in range(3):
            step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            step()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1) / 10], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        res[tag] = {"step_ms": ms, "codewords_per_s": B * world / (ms * 1e-3)}
        # the same step as ONE CUDA graph (forward with tape + BCE + sparse backward + SGD update captured once, replayed):
        # at the reference's minibatch the eager step is dominated by autograd / launch overhead outside the two kernels
        if B > 16384:
            continue                                         # large batches are kernel-bound: nothing to gain from a graph
        try:
            opt = torch.optim.SGD(m.parameters(), lr=1e-3)
            static_llr = llr.detach().clone()

            def gstep():
                opt.zero_grad(set_to_none=True)
                loss = torch.nn.functional.binary_cross_entropy(m(None, static_llr, 20.0).clamp(1e-6, 1 - 1e-6), y)
                loss.backward()
                opt.step()
                return loss
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for _ in range(3):
                    gstep()
            torch.cuda.current_stream(dev).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            opt.zero_grad(set_to_none=True)
            with torch.cuda.graph(graph):
                static_loss = gstep()
            for _ in range(3):
                graph.replay()
            barrier()
            e0.record()
            for _ in range(20):
                graph.replay()
            e1.record()
            barrier()
            tg = torch.tensor([e0.elapsed_time(e1) / 20], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(tg, op=dist.ReduceOp.MAX)
            res[tag]["cuda_graph_step_ms"] = float(tg.item())
            res[tag]["cuda_graph_loss_finite"] = bool(torch.isfinite(static_loss).item())
        except Exception as ex:                              # capture is an optimisation, never a requirement
            res[tag]["cuda_graph_step_ms"] = None
            res[tag]["cuda_graph_error"] = str(ex)[:160]
    res["note"] = ("default (64,32) code, sum-product, forward with tape + BCE + backward (ldpc_bp_train_forward/backward, 2 kernel launches "
                   "per step + torch loss ops); cuda_graph_step_ms = the whole SGD step (zero_grad, forward, loss, backward, optimizer update) captured "
                   "once in a CUDA graph and replayed; the reference's check-node backward alone materialises [B,E,E,E] = 1.8 GB at B = 512")
    return res


def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
        return
    import numpy as np
    import torch
    import torch.distributed as dist
    from ldpc_b200.codes import ieee80211n_1944_r12
    from ldpc_b200.decoder import LdpcCode
    from ldpc_b200 import _native as N

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # NCCL_DEBUG is left exactly as the launcher set it: NCCL writes its log lines to stdout, this script prints ONE
    # line that starts with '{' - a consumer takes that line (torchrun interleaves the ranks' output anyway).
    numa = bind_to_gpu_numa(local) if not a.no_numa_bind else {"bound": False}
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    qc = ieee80211n_1944_r12()
    code = LdpcCode(qc.H, qc_Z=81, qc_proto=qc.proto)
    B = a.codewords
    # ---- synthetic LLRs generated on the device (per-rank Philox stream) -----------------------
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    rng = np.random.RandomState(1234)
    cw = torch.as_tensor(qc.encode(rng.randint(0, 2, (256, qc.k)).astype(np.uint8))).to(dev)
    sigma = (1.0 / (2 * 0.5 * 10 ** (a.ebn0 / 10))) ** 0.5
    llr = torch.empty(B, qc.n, dtype=torch.float32, device=dev)
    chunk = 65536
    for s in range(0, B, chunk):
        e = min(B, s + chunk)
        bits = cw[torch.arange(s, e, device=dev) % 256].float()
        y = (1.0 - 2.0 * bits) + sigma * torch.randn(e - s, qc.n, device=dev, generator=gen)
        llr[s:e] = -2.0 * y / sigma ** 2
    del bits, y
    post = torch.empty(B, qc.n, dtype=torch.float32, device=dev)
    packed = torch.empty(B, code.packed_bytes, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream(dev)
    upd = N.UPDATE_IDS[a.update]
    lib = N.lib()

    def step():
        N.check(lib.ldpc_decode(code._h, llr.data_ptr(), N.F32, B, a.iters, upd, a.clamp, 1.0, None, None,
                                post.data_ptr(), None, packed.data_ptr(), None, None,
                                ctypes.c_void_p(stream.cuda_stream)))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(a.warmup, 1)):
        step()
    barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps + 1)]
    with Clocks(local) as clk:
        barrier()
        ev[0].record(stream)
        for i in range(a.steps):
            step()
            ev[i + 1].record(stream)
        barrier()
    total_ms = ev[0].elapsed_time(ev[-1])
    kern_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(a.steps)]
    t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / a.steps
    cw_per_s = B * world / (ms_per_step * 1e-3)
    value = cw_per_s * K_CODE / 1e9

    # ---- roofline of the dominant (only) kernel: measured live with CUDA events ------------------
    alg_bytes = B * (N_CODE * 4 + N_CODE * 4 + code.packed_bytes)         # LLR in + posterior out + packed bits
    k_ms = sum(kern_ms) / len(kern_ms)
    peaks, peak_src = {}, "fallback (B200_PROFILING.md: 6650 GB/s)"
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9
    clocks = clk.summary()
    sm_mhz = clocks["sm_mhz"] or peaks.get("sm_max_mhz", 1965.0)
    upd_per_cw = 2 * E_CODE * a.iters
    upd_s = B / (k_ms * 1e-3) * upd_per_cw
    plan = (ctypes.c_int32 * 4)()
    N.check(lib.ldpc_code_plan_info(code._h, plan))
    n_loc, n_sm, thr_cta, cw_cta = [int(v) for v in plan]
    # shared-memory ceiling of the storage format actually used: 2 LDS + 2 STS of 4 B per edge of the
    # n_sm blocks exchanged through shared memory (the other n_loc blocks and the LLRs are registers)
    smem_bytes_per_update = 4 * n_sm * 81 * 4 / (2 * E_CODE) if n_sm else (4 * E_CODE + N_CODE) * 4 / (2 * E_CODE)
    smem_peak = 148 * 128 * sm_mhz * 1e6 / smem_bytes_per_update
    # issue ceiling of the instruction stream: 705 SASS instructions per thread-iteration (371 variable phase + 334 check
    # phase, profiles/r02_sass_loop.txt), thr_cta/cw_cta thread slots per codeword, 2*E updates per iteration; ALU-pipe
    # ceiling: 210 FMNMX.XORSIGN + 66 ISETP/SEL per thread-iteration on the half-rate ALU pipe (measured 0.5
    # warp-instr/clk/SMSP, profiles/r01_pipe_rates.txt)
    lane_instr_per_update = 705 * (thr_cta / max(cw_cta, 1)) / (2 * E_CODE) if n_sm else None
    issue_peak = 148 * 4 * 32 * sm_mhz * 1e6 / lane_instr_per_update if lane_instr_per_update else None
    alu_peak = 148 * 4 * 16 * sm_mhz * 1e6 / (276 * (thr_cta / max(cw_cta, 1)) / (2 * E_CODE)) if n_sm else None
    out = {
        "metric": METRIC, "value": value, "unit": "Gbit/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a), "codewords_per_gpu_per_step": B,
                   "outputs": "posterior LLR f32 + packed hard bits", "kernel": ("generic", "qc", "tiny", "qc_rt")[code.kernel],
                   "l2_policy": f"inputs larger than L2 ({B * N_CODE * 4 / 1e9:.2f} GB of LLRs per step)"},
        "codewords_per_s": cw_per_s, "edge_updates_per_s": upd_s * world,
        "gpu_launches": a.steps,
        "clocks": clocks,
        # the BINDING resource of the dominant kernel: shared-memory bandwidth / issue slots (north_star: "edge-updates/s as
        # the roofline quantity"); HBM is at < 10 % and is reported beside it as roofline_hbm
        "roofline": {"bound": "issue" if (issue_peak and issue_peak < smem_peak) else "smem",
                     "achieved": upd_s, "peak": min(smem_peak, issue_peak or smem_peak),
                     "unit": "edge-updates/s", "frac": upd_s / min(smem_peak, issue_peak or smem_peak),
                     "smem_peak": smem_peak, "issue_peak": issue_peak, "alu_pipe_peak": alu_peak,
                     "plan": {"register_blocks": n_loc, "smem_blocks": n_sm, "threads_per_cta": thr_cta, "codewords_per_cta": cw_cta},
                     "traffic": None,
                     "peak_source": f"smem: 148 SMs x 128 B/clk x {sm_mhz:.0f} MHz / {smem_bytes_per_update:.2f} B per directed edge-update; "
                                    f"issue: 148 SMs x 4 SMSP x 32 lanes x {sm_mhz:.0f} MHz / {lane_instr_per_update or 0:.2f} lane-instr per update; "
                                    "ALU pipe (FMNMX at half rate) in alu_pipe_peak"},
        "roofline_hbm": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                         "traffic": (15756.0 * B / 1e9) if (code.kernel == 1 and a.update != "sp") else None, "traffic_unit": "GB per launch",
                         "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum = 15 756 B per codeword in the ncu --set full capture of this "
                                           "launch shape (profiles/r01_decode_qc_ncu_summary.txt), scaled by the batch; not re-measured per run",
                         "algorithmic_bytes_per_launch_gb": alg_bytes / 1e9, "peak_source": peak_src,
                         "note": "not the binding resource: the decoder keeps its messages on the SM"},
    }

    # ---- the f16x2 variant of the same kernel (two codewords per thread), reported beside the fp32 headline ----
    if a.update in ("minsum", "nms") and not a.no_f16:
        code.set_precision("f16")
        for _ in range(2):
            step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(a.steps):
            step()
        e1.record(stream)
        barrier()
        tt = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms16 = float(tt.item()) / a.steps
        out["f16x2"] = {"value": B * world / (ms16 * 1e-3) * K_CODE / 1e9, "unit": "Gbit/s", "ms_per_step": ms16, "dtype": "f16",
                        "note": "same workload and outputs, messages and LLRs in binary16, two codewords per thread (LDPC_PREC_F16X2); "
                                "bit-exact against oracle/bp_oracle.py::bp_decode_f16; not the headline value"}
        code.set_precision("f32")
        step()
        barrier()

    # ---- BASELINE.json configs[2] names both update rules: the reference's own rule (tanh sum-product), same kernel family ----
    if a.update == "minsum" and not a.no_sp:
        Bs = min(B, 262144)

        def sp_step():
            N.check(lib.ldpc_decode(code._h, llr.data_ptr(), N.F32, Bs, a.iters, N.UPDATE_IDS["sp"], a.clamp, 1.0, None, None,
                                    post.data_ptr(), None, packed.data_ptr(), None, None, ctypes.c_void_p(stream.cuda_stream)))
        sp_step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(2):
            sp_step()
        e1.record(stream)
        barrier()
        tt = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_sp = float(tt.item()) / 2
        out["sum_product"] = {"value": Bs * world / (ms_sp * 1e-3) * K_CODE / 1e9, "unit": "Gbit/s", "ms_per_step": ms_sp,
                              "codewords_per_gpu_per_step": Bs, "dtype": "f32",
                              "note": "the reference's update rule (tanh product, clamp, log((1+p)/(1-p)), bp/bp_cv.py:38-50) on the same kernel; "
                                      "libm tanhf/logf and a correctly rounded division per edge (~5 150 SASS instructions per thread-iteration vs 734)"}
        step()
        barrier()

    # ---- BASELINE.json configs[4]: NN demapper + quantized OFDM + BP decoder on the default (64,32) code ------
    if not a.no_nn:
        out["nn_demapper"] = bench_nn(a, dev, world, barrier, peaks)

    if not a.no_train:
        out["bp_training"] = bench_train(dev, world, barrier)

    # ---- BASELINE.json configs[3]: sharded BER/FER sweep, fixed total, one all-reduce of the counters in the timed region ----
    if not a.no_sweep:
        out["sweep"] = bench_sweep(a, dev, world, rank, barrier)

    # ---- BASELINE.json configs[1]: the quantized link on the reference's own code at one SNR -------------------------------
    if not a.no_sweep:
        out["quantized_link"] = bench_quantized_link(a, dev, world, rank, barrier)

    # ---- the whole IEEE 802.11n family on compiled kernels ----------------------------------------------
    if not a.no_family:
        out["wifi_family"] = bench_family(a, dev, world, barrier, upd_s)

    # ---- e2e: the C-ABI host call (decode_bits path): pinned host LLRs in, packed bits out ---------
    if not a.no_e2e:
        Be = min(a.e2e_codewords, B)
        h_llr = torch.empty(Be, qc.n, dtype=torch.float32).pin_memory()
        h_llr.copy_(llr[:Be])
        h_packed = torch.empty(Be, code.packed_bytes, dtype=torch.uint8).pin_memory()

        def e2e_step():
            N.check(lib.ldpc_decode_host(code._h, h_llr.data_ptr(), N.F32, Be, a.iters, upd, a.clamp, 1.0,
                                         None, h_packed.data_ptr(), None, None, a.e2e_chunk))
        # raw pinned H2D bandwidth of this box, for context
        d_tmp = torch.empty_like(h_llr, device=dev)
        barrier()                                            # all ranks copy at the same time: the rate the box sustains under load
        t0 = time.perf_counter()
        for _ in range(3):
            d_tmp.copy_(h_llr, non_blocking=True)
        torch.cuda.synchronize()
        h2d_gbs = 3 * h_llr.numel() * 4 / (time.perf_counter() - t0) / 1e9
        if world > 1:
            hb = torch.tensor([h2d_gbs], device=dev, dtype=torch.float64)
            dist.all_reduce(hb, op=dist.ReduceOp.MIN)
            h2d_gbs = float(hb.item())
        del d_tmp
        for _ in range(2):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            e2e_step()
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = Be * world * a.steps * K_CODE / float(dt.item()) / 1e9
        out["e2e"] = {"value": e2e, "unit": "Gbit/s", "h2d_bytes_per_step": Be * qc.n * 4,
                      "d2h_bytes_per_step": Be * code.packed_bytes, "codewords_per_step": Be, "pinned_h2d_gbs": h2d_gbs,
                      "pinned_h2d_note": "slowest rank, all ranks copying concurrently", "chunk_codewords": a.e2e_chunk, "numa": numa,
                      "pcie_ceiling_gbps": h2d_gbs * 1e9 / (qc.n * 4) * K_CODE / 1e9 * world,
                      "api": "ldpc_decode_host (C ABI behind ofdm_functions.decode_bits), pinned f32 LLRs in, packed bits out"}
        assert torch.equal(h_packed.to(dev), packed[:Be]), "e2e result differs from the device path"
        # the same call with receiver-quantised int8 LLRs (LDPC_I8: a quarter of the PCIe bytes); NOT the headline - the
        # reference's input is f32 - but it shows what bounds e2e.  Checked against the device path on the same values.
        h_q = (llr[:Be] * (127.0 / 32.0)).round().clamp(-127, 127).to(torch.int8).cpu().pin_memory()

        def e2e_i8_step():
            N.check(lib.ldpc_decode_host(code._h, h_q.data_ptr(), N.I8, Be, a.iters, upd, a.clamp, 1.0,
                                         None, h_packed.data_ptr(), None, None, 16384))
        for _ in range(2):
            e2e_i8_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            e2e_i8_step()
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        ref_q = code.decode(h_q.to(dev).float(), a.iters, a.clamp, update=a.update, want=("hard_packed",))["hard_packed"]
        assert torch.equal(h_packed.to(dev), ref_q), "int8 e2e result differs from the device path on the same values"
        # the reference's decode_bits contract itself: ordinary (pageable) float64 ndarray in, float64 {0,1} ndarray out
        from ldpc_b200.decoder import decode_bits_host
        Bd = min(Be, 65536)
        np_llr = h_llr[:Bd].numpy().astype(np.float64)
        np_out = np.empty((Bd, qc.n), np.float64)
        decode_bits_host(code, np_llr, a.iters, a.clamp, np_out, update=a.update)
        barrier()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            decode_bits_host(code, np_llr, a.iters, a.clamp, np_out, update=a.update)
        barrier()
        dtb = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dtb, op=dist.ReduceOp.MAX)
        assert np.array_equal(np.packbits(np_out.astype(np.uint8), axis=1), packed[:Bd].cpu().numpy()), "decode_bits result differs from the device path"
        out["e2e_decode_bits_f64"] = {"value": Bd * world * a.steps * K_CODE / float(dtb.item()) / 1e9, "unit": "Gbit/s", "codewords_per_step": Bd,
                                      "host_bytes_per_codeword": qc.n * 16,
                                      "note": "ldpc_decode_bits_host behind ofdm_functions.decode_bits: pageable float64 ndarray in, float64 {0,1} ndarray out "
                                              "(the reference's own formats, 31 KB of host memory per codeword); host threads cast/expand around the GPU"}
        out["e2e_int8_llr"] = {"value": Be * world * a.steps * K_CODE / float(dt.item()) / 1e9, "unit": "Gbit/s", "h2d_bytes_per_step": Be * qc.n,
                               "d2h_bytes_per_step": Be * code.packed_bytes,
                               "note": "same call, LLRs quantised to int8 on the host side of the receiver (llr * 127/32, clamp 20 unchanged => not the same "
                                       "decisions as f32; a feature for callers whose demapper emits fixed-point LLRs)"}

    # ---- CPU baseline beside it (rank 0, N=1 only) ---------------------------------------------------
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        C, g, cl, cores = cpu_sample(a, a.cpu_seconds)
        t0 = time.perf_counter()
        ref = C.decode(g, cl, a.iters, a.clamp, a.update, 1.0, want=("hard",))
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": cl.shape[0] * K_CODE / dt / 1e9, "unit": "Gbit/s", "cores": cores, "kind": "port",
                               "sample": f"{cl.shape[0]} codewords of the same workload, oracle/ldpc_oracle.c, {cores} pthreads, {dt:.1f} s"}
        got = code.decode(torch.as_tensor(cl[:4096]).to(dev), a.iters, a.clamp, update=a.update, want=("hard",))["hard"]
        out["cpu_baseline"]["hard_bits_equal_on_4096"] = bool(np.array_equal(got.cpu().numpy(), ref["hard"][:4096]))
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
