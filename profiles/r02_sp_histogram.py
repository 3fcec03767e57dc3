import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("ldpc-sims_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import numpy as np, torch
import c_oracle as C
from ldpc_b200.codes import ieee80211n_1944_r12
from ldpc_b200.decoder import LdpcCode
qc = ieee80211n_1944_r12(); g = C.CGraph(qc.H)
code = LdpcCode(qc.H, qc_Z=81, qc_proto=qc.proto)
rng = np.random.RandomState(1234); B = 4096
c = qc.encode(rng.randint(0, 2, (B, qc.k)).astype(np.uint8))
sigma = (1.0 / (2 * 0.5 * 10 ** 0.2)) ** 0.5
llr = (-2.0 * ((1.0 - 2.0 * c) + sigma * rng.randn(B, qc.n)) / sigma ** 2).astype(np.float32)
cols = np.nonzero(qc.H)[1]
edges = np.array([0, 1e-7, 1e-6, 1e-5, 1e-4, 1e-3, 1e-2, 1e-1, np.inf])
print("# sum-product marginals, GPU (decode_qc_kernel, CUDA libm) vs oracle/ldpc_oracle.c (glibc libm): 802.11n n=1944 r=1/2, first 4096 codewords")
print("# of the headline workload (BPSK/AWGN Eb/N0 = 2 dB), clamp 20.  e = |dt| / max(|t|, 16.64); histogram bins", edges.tolist())
for iters in (1, 2, 3, 5, 10):
    sp = C.decode(g, llr, iters, 20.0, "sp", want=("t", "hard", "x", "syndrome"))
    o = code.decode(torch.as_tensor(llr).cuda(), iters, 20.0, update="sp", want=("llr_post", "hard", "syndrome"))
    t = o["llr_post"].cpu().numpy().astype(np.float64) / -2.0
    ab = np.abs(t - sp["t"]); e = ab / np.maximum(np.abs(sp["t"]), 16.64); rel = ab / np.maximum(np.abs(sp["t"]), 1e-30)
    amax = np.zeros((B, qc.n)); np.maximum.at(amax, (slice(None), cols), np.abs(sp["x"]))
    bad = e > 1e-4
    dv = qc.H.sum(0)
    print(f"iterations {iters:2d}: hard bits equal {np.array_equal(o['hard'].cpu().numpy(), sp['hard'])}, syndrome weights equal {np.array_equal(o['syndrome'].cpu().numpy(), sp['syndrome'])}, "
          f"within 1e-4 relative {np.mean((rel <= 1e-4) | (ab <= 1e-6)):.6f}, histogram {np.histogram(e, bins=edges)[0].tolist()}, worst e {e.max():.3e} (abs {ab.max():.3e})"
          + (f"; outliers: {int(bad.sum())}, smallest largest-incoming-|x| among them {amax[bad].min():.2f}, column degrees {sorted(set(dv[np.nonzero(bad)[1]].tolist()))}" if bad.any() else "; no outlier"))
