import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ldpc-sims_b200"))
import numpy as np, torch
from ldpc_b200.codes import expand_qc
from ldpc_b200.decoder import LdpcCode
# a 12 x 24, Z = 96 prototype with the 802.11n r=1/2 degree profile but random shifts: a code the library has never seen
rng = np.random.RandomState(5)
from ldpc_b200.codes import ieee80211n_1944_r12
base = ieee80211n_1944_r12().proto
Z = 96
proto = np.where(base >= 0, rng.randint(0, Z, base.shape), -1).astype(np.int16)
H = expand_qc(proto, Z)
llr = torch.randn(131072, H.shape[1], device="cuda") * 2 + 1.5
res = {}
for name, kw in (("qc_rt (run-time tables)", {}), ("jit (compiled at run time)", dict(specialize=True))):
    t0 = time.time()
    code = LdpcCode(H, qc_Z=Z, qc_proto=proto, **kw)
    t_create = time.time() - t0
    out = code.decode(llr, 10, 20.0, update="minsum", want=("hard_packed", "llr_post"))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        out = code.decode(llr, 10, 20.0, update="minsum", want=("hard_packed", "llr_post"))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    res[name] = out
    print(f"{name}: kernel id {code.kernel}, create {t_create:.1f} s, {ms:.2f} ms per 131072 codewords = {131072 * (H.shape[1] - H.shape[0]) / ms / 1e6:.2f} Gbit/s information")
a, b = res.values()
print("bit-identical:", torch.equal(a["hard_packed"], b["hard_packed"]) and torch.equal(a["llr_post"], b["llr_post"]))
