"""Trainable-weight ("neural") BP forward, SURVEY.md section 8f rank 2 (reference bp/bp_vc.py:16-32,101-107;
bp/bp.py:26-51).  Golden vectors: the reference model itself with seeded random weights
(oracle/make_golden_weighted.py)."""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = np.load(os.path.join(ROOT, "tests", "golden", "bp_weighted.npz"))
STATE = {k[2:]: G[k] for k in G.files if k.startswith("s_")}


def test_weighted_oracle_matches_reference_golden():
    import bp_oracle as O
    from ldpc_b200.codes import peg_64_32
    H = peg_64_32()[0]
    w = O.weights_from_reference_state(O.Graph(H), STATE, int(G["iters"]))
    o = O.bp_decode(H, G["llr"], int(G["iters"]), float(G["clamp"]), weights=w)
    assert np.abs(o["prob"] - G["prob"]).max() < 5e-7
    assert np.array_equal(np.packbits(o["hard"], axis=1), G["hard"])
    # all-ones weights reproduce the unweighted definition on this code (dv <= 2)
    ones = {k: np.ones_like(v) for k, v in w.items()}
    a = O.bp_decode(H, G["llr"], 3, 20.0, weights=ones)
    b = O.bp_decode(H, G["llr"], 3, 20.0)
    assert np.array_equal(a["t"], b["t"]) and np.array_equal(a["x"], b["x"])


@pytest.mark.gpu
def test_drop_in_module_loads_reference_weights():
    """model.load_state_dict(reference_state) then model(x, llr, clamp) - bp/bp.py:43-51 with trained weights."""
    import torch
    from bp.bp import BeliefPropagation
    from bp.parity import H
    m = BeliefPropagation(H, int(G["iters"]))
    m.load_state_dict({("module." + k): torch.tensor(v) for k, v in STATE.items()})
    llr = torch.tensor(G["llr"]).cuda()
    prob = m(torch.zeros(llr.shape[0], m.layer_size(), device="cuda"), llr, float(G["clamp"])).cpu().detach().numpy()   # as ofdm_functions.py:161
    dp = np.abs(prob - G["prob"])                       # CUDA libm vs ATen tanh/log: same bar as the unweighted SP tests
    assert np.mean(dp <= 1e-5) >= 0.999 and dp.max() <= 5e-3
    assert np.array_equal(np.packbits(np.round(prob).astype(np.uint8), axis=1), G["hard"])
    # a state_dict whose weights are all ones keeps the fast unweighted kernels
    ones = {k: torch.ones(v.shape) for k, v in STATE.items()}
    m1 = BeliefPropagation(H, int(G["iters"])).eval()
    m1.load_state_dict(ones)
    p1 = m1(None, llr, 20.0).cpu().numpy()
    p0 = BeliefPropagation(H, int(G["iters"])).eval()(None, llr, 20.0).cpu().numpy()
    assert m1._all_ones() and not m.eval()._all_ones() and np.array_equal(p1, p0)
    # training mode (tape kernel) and inference mode (weighted decoder) give the same bits
    assert np.array_equal(m.eval()(None, llr, float(G["clamp"])).cpu().numpy(), prob)
    with pytest.raises(RuntimeError, match="layers.4.0.input_weight"):        # a checkpoint of another iteration count: strict loading names the layers it lacks
        BeliefPropagation(H, 5).load_state_dict({k: torch.tensor(v) for k, v in STATE.items()})


@pytest.mark.gpu
@pytest.mark.parametrize("update", ["minsum", "sp"])
def test_weighted_kernel_against_oracle(update):
    """Random weights on the default code and on the 802.11n code (dv up to 11): min-sum bit-exact, sum-product
    within the libm tolerance of the unweighted tests."""
    import torch
    import bp_oracle as O
    from ldpc_b200.codes import peg_64_32, ieee80211n_1944_r12
    from ldpc_b200.decoder import LdpcCode
    rng = np.random.RandomState(17)
    for name, H, B, iters in (("default", peg_64_32()[0], 333, 4), ("wifi", ieee80211n_1944_r12().H, 40, 3)):
        g = O.Graph(H)
        code = LdpcCode(H)
        mdv = int(g.dv.max())
        w = dict(w_edge=(0.5 + rng.rand(iters, g.E, mdv)).astype(np.float32), w_llr=(0.5 + rng.rand(iters, g.n)).astype(np.float32),
                 wf_edge=(0.5 + rng.rand(g.E)).astype(np.float32), wf_llr=(0.5 + rng.rand(g.n)).astype(np.float32))
        llr = (rng.randn(B, g.n) * 3).astype(np.float32)
        o = O.bp_decode(H, llr, iters, 20.0, update=update, weights=w)
        dw = {k: torch.as_tensor(v).cuda() for k, v in w.items()}
        dw.update(iterations=iters, stride=mdv)
        out = code.decode_weighted(torch.as_tensor(llr).cuda(), dw, 20.0, update=update, want=("llr_post", "hard", "syndrome", "x"))
        t = out["llr_post"].cpu().numpy() / -2.0
        if update == "minsum":
            assert np.array_equal(t, o["t"]) and np.array_equal(out["x"].cpu().numpy(), o["x"]), name
            assert np.array_equal(out["hard"].cpu().numpy(), o["hard"]) and np.array_equal(out["syndrome"].cpu().numpy(), o["syndrome"])
        else:
            d = np.abs(t - o["t"])
            assert np.mean(d <= 1e-4 * np.maximum(np.abs(o["t"]), 1e-3)) >= 0.999, name
            assert np.mean(out["hard"].cpu().numpy() != o["hard"]) < 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("update", ["minsum", "sp"])
def test_weighted_tiny_kernel_equals_generic(update):
    """The register-resident (64,32) kernel applies the same weighted node arithmetic: bit-identical to the generic kernel."""
    import torch
    from ldpc_b200.codes import peg_64_32
    from ldpc_b200.decoder import LdpcCode
    H = peg_64_32()[0]
    tiny, gen = LdpcCode(H), LdpcCode(H)
    gen.set_kernel("generic")
    rng = np.random.RandomState(5)
    iters, mdv, E, n = 6, 2, 96, 64
    dw = dict(w_edge=torch.as_tensor((0.5 + rng.rand(iters, E, mdv)).astype(np.float32)).cuda(),
              w_llr=torch.as_tensor((0.5 + rng.rand(iters, n)).astype(np.float32)).cuda(),
              wf_edge=torch.as_tensor((0.5 + rng.rand(E)).astype(np.float32)).cuda(),
              wf_llr=torch.as_tensor((0.5 + rng.rand(n)).astype(np.float32)).cuda(), iterations=iters, stride=mdv)
    llr = torch.as_tensor((rng.randn(1003, n) * 3).astype(np.float32)).cuda()
    want = ("prob", "llr_post", "hard", "hard_packed", "syndrome")
    a = tiny.decode_weighted(llr, dw, 20.0, update=update, want=want)
    b = gen.decode_weighted(llr, dw, 20.0, update=update, want=want)
    for k in want:
        assert torch.equal(a[k], b[k]), (update, k)
