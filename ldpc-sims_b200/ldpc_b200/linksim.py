"""Link-simulator front end and fused Monte-Carlo simulation over the C ABI.

numpy-facing wrappers (float64 / complex128 like the reference's host code,
ofdm/ofdm_functions.py:8-128) stage through torch CUDA tensors; all arithmetic runs in the
hand-written kernels of csrc/frontend.cu and csrc/sim.cu.
"""
from __future__ import annotations

import ctypes
import os
import dataclasses

import numpy as np
import torch

from . import _native as N
from .codes import systematic_generator

_RD = {np.dtype(np.complex128): (N.F64, torch.complex128, torch.float64),
       np.dtype(np.complex64): (N.F32, torch.complex64, torch.float32)}


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def pack_generator_rows(G: np.ndarray) -> np.ndarray:
    """[rows, k] 0/1 -> u32 [rows, ceil(k/32)], bit j of word w = G[r][32 w + j]."""
    G = (np.asarray(G) != 0).astype(np.uint8)
    rows, k = G.shape
    kw = (k + 31) // 32
    pad = np.zeros((rows, kw * 32), np.uint8)
    pad[:, :k] = G
    by = np.packbits(pad.reshape(rows, kw, 4, 8), axis=-1, bitorder="little").reshape(rows, kw, 4)
    return np.ascontiguousarray(by).view("<u4").reshape(rows, kw).copy()


def encode_bits(bits, generator_matrix):
    """(1, N*k) bits -> (1, N*n) float64 codeword bits, c = G u mod 2 (ofdm_functions.py:11-15)."""
    N.require_cuda()
    G = np.asarray(generator_matrix)
    n, k = G.shape
    b = np.asarray(bits).reshape(-1, k).astype(np.uint8)
    out = torch.empty(b.shape[0], n, dtype=torch.uint8, device="cuda")
    Gp = _dev(pack_generator_rows(G).view(np.int32))
    bd = _dev(b)
    N.check(N.lib().ldpc_encode_bits(bd.data_ptr(), Gp.data_ptr(), n, k, b.shape[0], out.data_ptr(), _stream()))
    return out.cpu().numpy().astype(np.float64).reshape(1, -1)


def modulate_bits(bits):
    """bit pairs -> QPSK symbols (1, L) complex128 (ofdm_functions.py:17-22)."""
    N.require_cuda()
    b = np.asarray(bits).reshape(-1).astype(np.uint8)
    L = b.size // 2
    out = torch.empty(L, dtype=torch.complex128, device="cuda")
    bd = _dev(b)
    N.check(N.lib().ldpc_modulate_bits(bd.data_ptr(), L, N.F64, out.data_ptr(), _stream()))
    return out.cpu().numpy().reshape(1, -1)


def _complex_in(a):
    a = np.ascontiguousarray(np.asarray(a).reshape(-1))
    if a.dtype not in _RD:
        a = a.astype(np.complex128)
    return a, _RD[a.dtype]


def transmit_symbols(symbols, ofdm_size, snr, noise=None, device_rng=False, seed=0):
    """-> (received (1,L), clean OFDM time signal (1,L)) (ofdm_functions.py:25-35).

    By default the noise is drawn on the HOST with np.random.normal in the reference's order
    (real part first, then imaginary), so a seeded reference experiment reproduces; the IDFT
    and the addition run on the GPU.  device_rng=True draws it from Philox on the device."""
    N.require_cuda()
    s, (rd, ct, _) = _complex_in(symbols)
    L = s.size
    if L % ofdm_size:
        raise ValueError("number of symbols must be a multiple of ofdm_size")
    n_ofdm = L // ofdm_size
    nd = None
    if noise is None and not device_rng:
        shape = (ofdm_size, n_ofdm)                         # the reference draws on the transposed layout
        noise = (np.random.normal(0, 1 / np.sqrt(snr), shape) + 1j * np.random.normal(0, 1 / np.sqrt(snr), shape)) / np.sqrt(2)
        noise = noise.T
    if noise is not None:
        nd = _dev(np.asarray(noise).reshape(-1).astype(s.dtype))
    sd = _dev(s)
    rx = torch.empty(L, dtype=ct, device="cuda")
    tx = torch.empty(L, dtype=ct, device="cuda")
    N.check(N.lib().ldpc_ofdm_transmit(sd.data_ptr(), n_ofdm, int(ofdm_size), rd, None if nd is None else nd.data_ptr(),
                                       float(snr), int(seed), rx.data_ptr(), tx.data_ptr(), _stream()))
    return rx.cpu().numpy().reshape(1, -1), tx.cpu().numpy().reshape(1, -1)


def quantizer(inputs, num_bits, clip_value):
    """Mid-tread quantizer with the reference's clip behaviour (ofdm_functions.py:37-51)."""
    N.require_cuda()
    x = np.asarray(inputs)
    shape = x.shape
    s, (rd, ct, _) = _complex_in(x)
    num_levels = float(np.power(2, np.asarray(num_bits).reshape(-1)[0]))
    sd = _dev(s)
    out = torch.empty_like(sd)
    N.check(N.lib().ldpc_quantize(sd.data_ptr(), 2 * s.size, rd, num_levels, float(np.asarray(clip_value).reshape(-1)[0]),
                                  out.data_ptr(), _stream()))
    return out.cpu().numpy().reshape(shape)


def demodulate_signal(symbols, ofdm_size, snr_est):
    """-> (llrs (1,2L) float64 log P1/P0, de-OFDM'd symbols (1,L)) (ofdm_functions.py:63-78)."""
    N.require_cuda()
    s, (rd, ct, rt) = _complex_in(symbols)
    L = s.size
    if L % ofdm_size:
        raise ValueError("number of samples must be a multiple of ofdm_size")
    sd = _dev(s)
    llr = torch.empty(2 * L, dtype=rt, device="cuda")
    sym = torch.empty(L, dtype=ct, device="cuda")
    N.check(N.lib().ldpc_ofdm_demodulate(sd.data_ptr(), L // ofdm_size, int(ofdm_size), rd, float(snr_est),
                                         llr.data_ptr(), sym.data_ptr(), _stream()))
    return llr.cpu().numpy().reshape(1, -1), sym.cpu().numpy().reshape(1, -1)


# ------------------------------------------------------------------------------------------
# fused Monte-Carlo simulation
# ------------------------------------------------------------------------------------------
class SimParams(ctypes.Structure):
    _fields_ = [("struct_size", ctypes.c_int32), ("ofdm_size", ctypes.c_int32), ("qbits", ctypes.c_int32),
                ("agc_mode", ctypes.c_int32), ("agc_clip", ctypes.c_float), ("clip_ratio", ctypes.c_float),
                ("snr_db", ctypes.c_float), ("iters", ctypes.c_int32), ("update", ctypes.c_int32),
                ("clamp_value", ctypes.c_float), ("param", ctypes.c_float), ("reserved", ctypes.c_int32),
                ("seed", ctypes.c_uint64), ("first_codeword", ctypes.c_int64), ("n_codewords", ctypes.c_int64)]


@dataclasses.dataclass
class LinkConfig:
    """One operating point of the link (the constants at the top of the evaluate scripts,
    e.g. evaluate_quantized_snr.py:14-25)."""
    snr_db: float
    ofdm_size: int = 32
    qbits: int = 0
    agc_mode: int = 1
    agc_clip: float = 10.0
    clip_ratio: float = 1.0
    iters: int = 10
    update: str = "sp"
    clamp_value: float = 20.0
    param: float = 1.0
    seed: int = 1234
    force_unfused: bool = False      # A/B: three-launch chain instead of the single-launch kernel
    channel: str = "awgn"            # "awgn" (the reference) or "rayleigh": flat block fading per OFDM symbol, coherent receiver
    compander: bool = False          # tanh compander clip * tanh(x / clip) in front of the uniform ADC

    def to_struct(self, first, count):
        from .decoder import _update_id
        if self.channel not in ("awgn", "rayleigh"):
            raise ValueError("channel must be 'awgn' or 'rayleigh'")
        return SimParams(ctypes.sizeof(SimParams), self.ofdm_size, self.qbits, self.agc_mode, self.agc_clip,
                         self.clip_ratio, self.snr_db, self.iters, _update_id(self.update), self.clamp_value,
                         self.param, (1 if self.force_unfused else 0) | (2 if self.channel == "rayleigh" else 0) | (4 if self.compander else 0),
                         self.seed, first, count)


COUNTER_NAMES = ("uncoded_bit_errors", "info_bit_errors", "frame_errors", "bits", "frames")


def attach_generator(code, G=None, k=None):
    """Give a LdpcCode its systematic encoder.  G: [n,k] (rows k.. are the parity rows); if
    omitted it is derived from H by GF(2) elimination (H = [A | B], B invertible)."""
    if G is None:
        H = np.zeros((code.m, code.n), np.uint8)
        H[np.repeat(np.arange(code.m), np.diff(code.tables.chk_ptr)), code.tables.chk_var] = 1
        G = systematic_generator(H)
    G = (np.asarray(G) != 0).astype(np.uint8)
    n, kk = G.shape
    k = kk if k is None else k
    if n != code.n or not np.array_equal(G[:k], np.eye(k, dtype=np.uint8)):
        raise ValueError("G must be [n,k] systematic with the information bits first")
    rows = pack_generator_rows(G[k:])
    with torch.cuda.device(code.device):
        N.check(N.lib().ldpc_code_set_generator(code._h, rows.ctypes.data, int(k)))
    code.k = int(k)
    return code


def sim_generate(code, cfg: LinkConfig, first, count, want_samples=False):
    """K2 alone: (codewords packed MSB-first [count, ceil(n/8)] u8, llr f32 [count, n]) CUDA tensors;
    with want_samples also the MLP demappers' input rows f32 [count * symbols_per_codeword, 2N+1]
    (Re, Im of the received time samples and the linear SNR, evaluate_quantized_snr.py:135-140)."""
    dev = code.device
    cwp = torch.empty(count, code.packed_bytes, dtype=torch.uint8, device=dev)
    llr = torch.empty(count, code.n, dtype=torch.float32, device=dev)
    sp = cfg.to_struct(first, count)
    smp = None
    if want_samples:
        per_cw = (code.n // 2 + cfg.ofdm_size - 1) // cfg.ofdm_size
        smp = torch.zeros(count * per_cw, 2 * cfg.ofdm_size + 1, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        N.check(N.lib().ldpc_sim_generate_ex(code._h, ctypes.byref(sp), cwp.data_ptr(), llr.data_ptr(),
                                             None if smp is None else smp.data_ptr(), _stream()))
    return (cwp, llr, smp) if want_samples else (cwp, llr)


def sim_frontend(code, cfg: LinkConfig, cw_packed, noise=None, want_samples=False):
    """The simulator's link chain (QPSK, IFFT, + noise, AGC / ADC / rescale, FFT, exact LLR - float32) on
    CALLER-SUPPLIED codewords (packed MSB-first [count, ceil(n/8)] u8 CUDA tensor) and, optionally, caller-supplied
    noise (complex64 CUDA tensor [count, symbols_per_codeword, ofdm_size], already scaled): the identical-input
    form of evaluate_quantized_snr.py:96-133.  Returns llr f32 [count, n] (and the MLP input rows)."""
    dev = code.device
    count = cw_packed.shape[0]
    per_cw = (code.n // 2 + cfg.ofdm_size - 1) // cfg.ofdm_size
    llr = torch.empty(count, code.n, dtype=torch.float32, device=dev)
    smp = torch.zeros(count * per_cw, 2 * cfg.ofdm_size + 1, dtype=torch.float32, device=dev) if want_samples else None
    nz = None
    if noise is not None:
        nz = torch.view_as_real(noise.to(torch.complex64).contiguous()).contiguous()
        if nz.numel() != count * per_cw * cfg.ofdm_size * 2:
            raise ValueError("noise must be [count, symbols_per_codeword, ofdm_size] complex")
    sp = cfg.to_struct(0, count)
    with torch.cuda.device(dev):
        N.check(N.lib().ldpc_sim_frontend(code._h, ctypes.byref(sp), cw_packed.contiguous().data_ptr(),
                                          None if nz is None else nz.data_ptr(), llr.data_ptr(),
                                          None if smp is None else smp.data_ptr(), _stream()))
    return (llr, smp) if want_samples else llr


def decode_count(code, llr, ref_packed, cfg: LinkConfig, counters=None):
    """Decode llr [B,n] and add the exact link metrics against the transmitted codewords (packed) into
    `counters` (int64[5] CUDA tensor): one launch, ldpc_decode_count."""
    from .decoder import _update_id, _DTYPES
    dev = llr.device
    if counters is None:
        counters = torch.zeros(5, dtype=torch.int64, device=dev)
    if llr.dtype not in _DTYPES:
        llr = llr.float()
    llr = llr.contiguous()
    with torch.cuda.device(dev):
        N.check(N.lib().ldpc_decode_count(code._h, llr.data_ptr(), _DTYPES[llr.dtype], llr.shape[0], int(cfg.iters),
                                          _update_id(cfg.update), float(cfg.clamp_value), float(cfg.param),
                                          ref_packed.data_ptr(), int(code.k), counters.data_ptr(), _stream()))
    return counters


def sim_run_nn(code, cfg: LinkConfig, demapper, first, count, counters=None, chunk=1 << 20):
    """The NN-demapper link (evaluate_quantized_snr.py:91-188, the *_nn results): front end -> received
    time samples -> MLP LLR estimates -> BP decoder -> exact counters, all on the GPU.  `demapper`
    maps CUDA f32 [S, 2N+1] -> [S, 2N] (ldpc_b200.mlp.NativeMLP or the drop-in nn.llr modules).
    Needs n to be a multiple of 2 * ofdm_size (the reference uses n = 2 * ofdm_size)."""
    if code.n % (2 * cfg.ofdm_size):
        raise ValueError("the NN demapper path needs n to be a multiple of 2 * ofdm_size")
    dev = code.device
    if counters is None:
        counters = torch.zeros(5, dtype=torch.int64, device=dev)
    done = 0
    while done < count:
        cnt = min(chunk, count - done)
        cwp, _, smp = sim_generate(code, cfg, first + done, cnt, want_samples=True)
        llr_est = demapper(smp).reshape(cnt, code.n)
        decode_count(code, llr_est, cwp, cfg, counters)
        done += cnt
    return counters


def sim_run(code, cfg: LinkConfig, first, count, counters=None, workspace=None):
    """Fused link simulation of codewords [first, first+count); adds into `counters`
    (int64[5] CUDA tensor, COUNTER_NAMES) and returns it."""
    dev = code.device
    if counters is None:
        counters = torch.zeros(5, dtype=torch.int64, device=dev)
    if workspace is None:
        per_cw = 4 * code.n + ((code.packed_bytes + 15) & ~15)
        chunk = int(min(max(count, 1024), max(16384, (256 << 20) // per_cw)))     # up to 256 MB of LLR tile per chunk
        chunk = (chunk + 1023) & ~1023
        workspace = torch.empty(chunk * per_cw, dtype=torch.uint8, device=dev)
    sp = cfg.to_struct(first, count)
    with torch.cuda.device(dev):
        N.check(N.lib().ldpc_sim_run(code._h, ctypes.byref(sp), workspace.data_ptr(), ctypes.c_size_t(workspace.numel()),
                                     counters.data_ptr(), _stream()))
    return counters


def shard_range(total, rank, world):
    """Contiguous codeword range of `rank` out of `world` (batch sharding)."""
    base, rem = divmod(int(total), int(world))
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def sweep(code, cfg_per_snr, codewords_per_point, rank=0, world=1, group=None):
    """BER/FER sweep: every rank simulates its shard of every SNR point, then ONE all-reduce
    (sum) of the int64 [S,5] counter matrix.  Returns a numpy int64 [S,5] (identical on all ranks)."""
    dev = code.device
    S = len(cfg_per_snr)
    counters = torch.zeros(S, 5, dtype=torch.int64, device=dev)
    first, count = shard_range(codewords_per_point, rank, world)
    ws = None
    for i, cfg in enumerate(cfg_per_snr):
        if count:
            if ws is None:
                per_cw = 4 * code.n + ((code.packed_bytes + 15) & ~15)
                ws = torch.empty(((min(max(count, 1024), max(16384, (256 << 20) // per_cw)) + 1023) & ~1023) * per_cw, dtype=torch.uint8, device=dev)
            sim_run(code, cfg, first, count, counters[i], ws)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(counters, op=dist.ReduceOp.SUM, group=group)
    return counters.cpu().numpy()


def weighted_mse_sums(est, ref, epsilon=10e-4):
    """Sum and count of (est - ref)^2 / (|ref| + eps) (ofdm_functions.py:80-81 / evaluate_quantized_snr.py:163-165)."""
    d = (est.double() - ref.double()) ** 2 / (ref.double().abs() + epsilon)
    return torch.stack([d.sum(), torch.tensor(float(d.numel()), dtype=torch.float64, device=d.device)])


def evaluate_point(code, cfg_q: LinkConfig, demapper, first, count, chunk=1 << 16):
    """One SNR point of evaluate_quantized_snr.py:91-188 on the SAME noise realisation (the simulator's draws
    depend on (seed, codeword index) only): unquantized link, quantized link with the conventional LLRs and -
    if `demapper` is given - quantized link with the MLP's LLR estimates.  Returns (int64 [3,5] counters in the
    order traditional / nn / quantized, float64 [2,2] weighted-MSE (sum, count) of nn / quantized LLRs)."""
    dev = code.device
    counters = torch.zeros(3, 5, dtype=torch.int64, device=dev)
    wm = torch.zeros(2, 2, dtype=torch.float64, device=dev)
    cfg_u = dataclasses.replace(cfg_q, qbits=0)
    done = 0
    while done < count:
        cnt = min(chunk, count - done)
        cwp, llr_u = sim_generate(code, cfg_u, first + done, cnt)
        decode_count(code, llr_u, cwp, cfg_u, counters[0])
        if cfg_q.qbits > 0:
            cwq, llr_q, smp = sim_generate(code, cfg_q, first + done, cnt, want_samples=demapper is not None)
            decode_count(code, llr_q, cwq, cfg_q, counters[2])
            wm[1] += weighted_mse_sums(llr_q, llr_u)
            if demapper is not None:
                llr_n = demapper(smp).reshape(cnt, code.n)
                decode_count(code, llr_n, cwq, cfg_q, counters[1])
                wm[0] += weighted_mse_sums(llr_n, llr_u)
        done += cnt
    return counters, wm


def evaluate_full(code, cfgs_q, demapper, codewords_per_point, rank=0, world=1, group=None, state_path=None, chunk=1 << 16):
    """The whole evaluate_quantized_snr.py result set (every key plots.py:11-27 reads) for a list of SNR points,
    sharded by batch over the ranks, reduced with ONE all-reduce per dtype at the end.  With `state_path` every
    finished point of this rank is checkpointed (npz) and skipped on restart: multi-hour sweeps resume."""
    dev = code.device
    S = len(cfgs_q)
    counters = torch.zeros(S, 3, 5, dtype=torch.int64, device=dev)
    wm = torch.zeros(S, 2, 2, dtype=torch.float64, device=dev)
    done = np.zeros(S, dtype=bool)
    path = None if state_path is None else f"{state_path}.rank{rank}of{world}.npz"
    import zlib
    sig = np.array([codewords_per_point] + [zlib.crc32(repr(dataclasses.astuple(c)).encode()) for c in cfgs_q], dtype=np.int64)
    if path and os.path.exists(path):
        st = np.load(path)
        if np.array_equal(st["sig"], sig):
            done = st["done"].copy()
            counters.copy_(torch.as_tensor(st["counters"]))
            wm.copy_(torch.as_tensor(st["wm"]))
    first, count = shard_range(codewords_per_point, rank, world)
    for i, cfg in enumerate(cfgs_q):
        if done[i] or count == 0:
            continue
        c, w = evaluate_point(code, cfg, demapper, first, count, chunk)
        counters[i], wm[i] = c, w
        done[i] = True
        if path:
            torch.cuda.synchronize()
            tmp = path + ".tmp.npz"
            np.savez(tmp, sig=sig, done=done, counters=counters.cpu().numpy(), wm=wm.cpu().numpy())
            os.replace(tmp, path)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(counters, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(wm, op=dist.ReduceOp.SUM, group=group)
    return counters.cpu().numpy(), wm.cpu().numpy()


def results_dict(snrdb, counters, wm, n, k):
    """Counters of evaluate_full -> the reference's result pickle (evaluate_quantized_snr.py:192-212)."""
    out = {"snrdb": np.asarray(snrdb)}
    for j, suffix in enumerate(("", "_nn", "_quantized")):
        r = rates(counters[:, j], n, k)
        for key, val in r.items():
            out[key + suffix] = val
    out["wmse_nn"] = wm[:, 0, 0] / np.maximum(wm[:, 0, 1], 1)
    out["wmse_quantized"] = wm[:, 1, 0] / np.maximum(wm[:, 1, 1], 1)
    return out


def rates(counters, n, k):
    """int64 [..,5] -> dict of the reference's result arrays (evaluate_quantized_snr.py:178-180)."""
    c = np.asarray(counters, dtype=np.float64)
    frames = np.maximum(c[..., 4], 1)
    return dict(uncoded_ber=c[..., 0] / np.maximum(c[..., 3], 1), coded_ber=c[..., 1] / (frames * k),
                coded_bler=c[..., 2] / frames)
