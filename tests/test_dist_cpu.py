"""N > 1 host logic on CPU: world_size-2 gloo run of the sweep's shard + all-reduce path with
the GPU simulation replaced by a deterministic per-codeword stub, and the Philox
known-answer vectors (Random123 kat_vectors)."""
import ctypes
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _stub_counters(first, count, snr_db):
    idx = np.arange(first, first + count, dtype=np.int64)
    h = (idx * 2654435761 + int(snr_db * 10)) % 97
    return np.array([h.sum(), (h % 5).sum(), (h % 2).sum(), count * 64, count], dtype=np.int64)


def _worker(rank, world, port, total, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "ldpc-sims_b200"))
    import ldpc_b200.linksim as LS

    class FakeCode:
        device = torch.device("cpu"); n = 64; packed_bytes = 8

    def fake_sim_run(code, cfg, first, count, counters, ws):
        counters += torch.from_numpy(_stub_counters(first, count, cfg.snr_db))
    LS.sim_run = fake_sim_run
    cfgs = [LS.LinkConfig(snr_db=s) for s in (0.0, 3.0)]
    torch.empty_like_orig = torch.empty
    real_empty = torch.empty
    torch.empty = lambda *a, **k: real_empty(*a, **{**k, "device": "cpu"}) if "device" in k else real_empty(*a, **k)
    res = LS.sweep(FakeCode(), cfgs, total, rank=rank, world=world)
    torch.empty = real_empty
    if rank == 0:
        np.save(out, res)
    dist.destroy_process_group()


def test_sweep_shards_and_allreduces_world2(tmp_path):
    total, port = 1001, 29591
    out = str(tmp_path / "res.npy")
    mp.spawn(_worker, args=(2, port, total, out), nprocs=2, join=True)
    res = np.load(out)
    want = np.stack([_stub_counters(0, total, s) for s in (0.0, 3.0)])
    assert np.array_equal(res, want)                       # 2 ranks == 1 rank, exactly


def test_philox_known_answers():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "ldpc-sims_b200"))
    from ldpc_b200 import _native as N
    L = N.lib()

    def ph(ctr, key):
        c = (ctypes.c_uint32 * 4)(*ctr); k = (ctypes.c_uint32 * 2)(*key); o = (ctypes.c_uint32 * 4)()
        L.ldpc_philox4x32(c, k, o)
        return list(o)
    assert ph([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert ph([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert ph([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
