"""Training path of the weighted decoder, SURVEY.md section 8f rank 2 (reference autograd Functions bp/bp_vc.py:34-58,
bp/bp_cv.py:57-91; training loop ofdm/ofdm_nn.py:257-396).  Golden gradients: the reference model's own .backward()
on CPU (oracle/make_golden_grad.py); oracle: float64 autograd of the restated forward (bp_oracle.bp_weighted_grad)."""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = np.load(os.path.join(ROOT, "tests", "golden", "bp_grad.npz"))
STATE = {k[2:]: G[k] for k in G.files if k.startswith("s_")}
KEYS = ("g_w_edge", "g_w_llr", "g_wf_edge", "g_wf_llr")


def _rel(a, b):
    return float(np.abs(np.asarray(a, np.float64) - b).max() / max(np.abs(b).max(), 1e-30))


def test_gradient_oracle_matches_reference_backward():
    """Unsaturated regime: the reference's hand-written backward is the derivative of its forward there."""
    import bp_oracle as O
    from ldpc_b200.codes import peg_64_32
    H = peg_64_32()[0]
    g = O.Graph(H)
    w = O.weights_from_reference_state(g, STATE, int(G["iters"]))
    o = O.bp_weighted_grad(H, G["llr"], int(G["iters"]), float(G["clamp"]), w, G["grad_prob"], graph=g)
    assert np.abs(o["prob"] - G["prob"]).max() < 1e-6
    assert _rel(G["grad_llr"], o["grad_llr"]) < 2e-4
    for k in KEYS:
        assert _rel(G[k], o[k]) < 2e-4, k


def test_sparse_dense_weight_round_trip():
    from ldpc_b200.codes import EdgeTables, peg_64_32, sparse_weights_from_reference_state, reference_state_from_sparse_weights
    import bp_oracle as O
    H = peg_64_32()[0]
    T = EdgeTables.from_H(H)
    w = sparse_weights_from_reference_state(T, STATE, int(G["iters"]))
    ow = O.weights_from_reference_state(O.Graph(H), STATE, int(G["iters"]))
    for k in ("w_edge", "w_llr", "wf_edge", "wf_llr"):
        assert np.array_equal(w[k], ow[k]), k
    back = reference_state_from_sparse_weights(T, w)
    for k, v in STATE.items():
        assert np.array_equal(back[k], v), k                  # the reference keeps weights masked: zeros elsewhere


def _module(iters):
    import torch
    from bp.bp import BeliefPropagation
    from bp.parity import H
    m = BeliefPropagation(H, iters)
    m.load_state_dict({k: torch.tensor(v) for k, v in STATE.items()})
    return m.cuda()


@pytest.mark.gpu
def test_backward_matches_reference_golden():
    """loss.backward() through the drop-in module = the gradients the reference left on its dense parameters."""
    import torch
    m = _module(int(G["iters"]))
    llr = torch.tensor(G["llr"]).cuda().requires_grad_(True)
    prob = m(torch.zeros(llr.shape[0], m.layer_size(), device="cuda"), llr, float(G["clamp"]))
    assert np.abs(prob.detach().cpu().numpy() - G["prob"]).max() < 1e-5
    loss = torch.nn.functional.binary_cross_entropy(prob, torch.tensor(G["target"]).float().cuda())
    loss.backward()
    assert _rel(llr.grad.cpu().numpy(), G["grad_llr"].astype(np.float64)) < 1e-4
    for k, p in (("g_w_edge", m.w_edge), ("g_w_llr", m.w_llr), ("g_wf_edge", m.wf_edge), ("g_wf_llr", m.wf_llr)):
        assert _rel(p.grad.cpu().numpy(), G[k].astype(np.float64)) < 1e-4, k
    # forward of the tape kernel == the inference kernel, bit for bit
    with torch.no_grad():
        assert torch.equal(m(None, llr.detach(), float(G["clamp"])), prob.detach())


@pytest.mark.gpu
def test_backward_with_saturated_clamp():
    """clamp_value = 3 (joint_evaluate.py:23 uses it): many messages sit ON the clamp, far from the boundary in the
    well-conditioned part of 2 atanh, and must pass no gradient (torch.clamp semantics = the derivative of the
    forward).  Rows in which fp32 and the float64 oracle fall on different sides of a clamp boundary may differ as a
    whole; the bulk must agree."""
    import torch
    import bp_oracle as O
    from ldpc_b200.codes import peg_64_32, EdgeTables, sparse_weights_from_reference_state
    from ldpc_b200.decoder import LdpcCode
    H = peg_64_32()[0]
    iters, clamp = int(G["iters"]), 3.0
    code = LdpcCode(H)
    w = sparse_weights_from_reference_state(EdgeTables.from_H(H), STATE, iters)
    dw = {k: torch.as_tensor(v).cuda() for k, v in w.items()}
    dw.update(iterations=iters, stride=w["w_edge"].shape[2])
    llr_np, gp_np = (2.5 * G["llr"]).astype(np.float32), G["grad_prob"]
    o = O.bp_weighted_grad(H, llr_np, iters, clamp, w, gp_np)
    llr, gp = torch.tensor(llr_np).cuda(), torch.tensor(gp_np).cuda()
    prob, tape = code.train_forward(llr, dw, clamp)
    assert float((tape[1:].abs() == clamp).float().mean()) > 0.05          # the case is about saturation
    g = code.train_backward(llr, dw, clamp, tape, gp)
    d = np.abs(g["grad_llr"].cpu().numpy() - o["grad_llr"])
    rows_ok = d.max(axis=1) <= 1e-3 * np.abs(o["grad_llr"]).max()
    assert rows_ok.mean() >= 0.95, rows_ok.mean()
    assert _rel(g["w_llr"].cpu().numpy(), o["g_w_llr"]) < 0.05


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["default", "wifi"])
def test_backward_random_weights_against_oracle(name):
    """Random weights and ragged batch sizes, default code and the 802.11n code (dv up to 11, dc 8)."""
    import torch
    import bp_oracle as O
    from ldpc_b200.codes import peg_64_32, ieee80211n_1944_r12
    from ldpc_b200.decoder import LdpcCode
    rng = np.random.RandomState(23)
    H, B, iters, scale = (peg_64_32()[0], 77, 4, 1.0) if name == "default" else (ieee80211n_1944_r12().H, 9, 2, 0.5)
    g = O.Graph(H)
    code = LdpcCode(H)
    mdv = int(g.dv.max())
    w = dict(w_edge=(0.5 + rng.rand(iters, g.E, mdv)).astype(np.float32), w_llr=(0.5 + rng.rand(iters, g.n)).astype(np.float32),
             wf_edge=(0.5 + rng.rand(g.E)).astype(np.float32), wf_llr=(0.5 + rng.rand(g.n)).astype(np.float32))
    llr = (rng.randn(B, g.n) * scale).astype(np.float32)
    gp = rng.randn(B, g.n).astype(np.float32)
    o = O.bp_weighted_grad(H, llr, iters, 20.0, w, gp, graph=g)
    dw = {k: torch.as_tensor(v).cuda() for k, v in w.items()}
    dw.update(iterations=iters, stride=mdv)
    prob, tape = code.train_forward(torch.as_tensor(llr).cuda(), dw, 20.0)
    assert np.abs(prob.cpu().numpy() - o["prob"]).max() < 2e-5
    ref = code.decode_weighted(torch.as_tensor(llr).cuda(), dw, 20.0, update="sp", want=("prob",))["prob"]
    assert torch.equal(prob, ref)
    out = code.train_backward(torch.as_tensor(llr).cuda(), dw, 20.0, tape, torch.as_tensor(gp).cuda())
    assert _rel(out["grad_llr"].cpu().numpy(), o["grad_llr"]) < 2e-4
    for k in ("w_edge", "w_llr", "wf_edge", "wf_llr"):
        assert _rel(out[k].cpu().numpy(), o["g_" + k]) < 2e-4, k


@pytest.mark.gpu
def test_joint_training_step_reduces_loss():
    """The reference's joint loop in miniature (ofdm/ofdm_nn.py:281-337): BCE through the decoder, SGD on its weights
    AND on a layer in front of it (the gradient reaches the demapper through grad_llr)."""
    import torch
    from bp.bp import BeliefPropagation
    from bp.parity import H, G as GEN
    torch.manual_seed(0)
    rng = np.random.RandomState(1)
    bits = rng.randint(0, 2, (512, 32))
    cw = (bits @ np.asarray(GEN).T % 2).astype(np.float32) if np.asarray(GEN).shape[0] == 64 else (bits @ np.asarray(GEN) % 2).astype(np.float32)
    y = torch.tensor(cw).cuda()
    obs = ((2 * y - 1) * 1.2 + torch.randn_like(y) * 1.5)          # log(P1/P0)-like observations, deliberately mis-scaled
    bp = BeliefPropagation(H, 3).cuda()
    scale = torch.nn.Parameter(torch.tensor(0.3, device="cuda"))
    opt = torch.optim.SGD([{"params": bp.parameters()}, {"params": [scale]}], lr=0.5)
    losses = []
    for _ in range(25):
        opt.zero_grad()
        prob = bp(None, obs * scale, 20.0)
        loss = torch.nn.functional.binary_cross_entropy(prob.clamp(1e-6, 1 - 1e-6), y)
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
    assert losses[-1] < 0.9 * losses[0], losses
    assert float(scale.detach()) != 0.3 and not bp.eval()._all_ones()
    # the trained module exports a checkpoint in the reference's layout and reads it back
    st = bp.reference_state_dict()
    bp2 = BeliefPropagation(H, 3)
    bp2.load_state_dict(st)
    for k in ("w_edge", "w_llr", "wf_edge", "wf_llr"):
        assert torch.equal(getattr(bp, k).detach().cpu(), getattr(bp2, k).detach()), k


@pytest.mark.gpu
def test_train_joint_drop_in(tmp_path):
    """train_joint with the reference's argument list on a small quantized-link data set: the loss falls, the test BER
    does not get worse than the untrained model's, the checkpoint has the reference's keys."""
    import torch
    from bp.parity import H, G as GEN
    from ofdm.ofdm_functions import create_bits, encode_bits, modulate_bits, gen_data, gen_qdata
    from ofdm.ofdm_nn import train_joint
    np.random.seed(0)
    torch.manual_seed(0)
    ofdm_size, snrdb = 32, 5.0

    def make(n_cw):
        bits = create_bits(n_cw * 32)
        enc = encode_bits(bits, GEN)
        tx = modulate_bits(enc)
        rx_signal = gen_data(tx, snrdb, ofdm_size)[0]
        q = gen_qdata(rx_signal, snrdb, 3, 1.0, ofdm_size)[0]
        x = np.concatenate((q.real.T, q.imag.T), axis=1).reshape(-1, 2 * ofdm_size)
        return x, enc.reshape(-1, 2 * ofdm_size)

    x, y = make(2048)
    xt, yt = make(512)
    name, model = train_joint(x, y, xt, yt, H, 3, 20, "test", snrdb, 0.01, 3, 0, ofdm_size, 3, 1024, model_dir=str(tmp_path),
                              minibatch_size=256, verbose=False, return_model=True)
    ck = torch.load(os.path.join(str(tmp_path), name), weights_only=False)
    assert ck["loss"].shape == (3,) and ck["loss"][-1] < ck["loss"][0]
    keys = set(ck["model_state_dict"].keys())
    assert {"module.LLRest.final.weight", "module.BP.w_edge", "module.BP.wf_llr"} <= keys
    assert not model.module.BP.eval()._all_ones()


@pytest.mark.gpu
def test_training_edge_cases():
    """Zero iterations (marginal of the channel LLR only), an empty batch, a ragged batch that leaves idle lanes in the
    last warp, and the argument checks of the training entry points."""
    import torch
    import bp_oracle as O
    from bp.bp import BeliefPropagation
    from ldpc_b200.codes import peg_64_32
    from ldpc_b200.decoder import LdpcCode
    H = peg_64_32()[0]
    g = O.Graph(H)
    code = LdpcCode(H)
    rng = np.random.RandomState(5)
    mdv = int(g.dv.max())
    for iters, B in ((0, 5), (2, 1), (1, 33)):
        w = dict(w_edge=(0.5 + rng.rand(iters, g.E, mdv)).astype(np.float32), w_llr=(0.5 + rng.rand(iters, g.n)).astype(np.float32),
                 wf_edge=(0.5 + rng.rand(g.E)).astype(np.float32), wf_llr=(0.5 + rng.rand(g.n)).astype(np.float32))
        llr, gp = rng.randn(B, g.n).astype(np.float32), rng.randn(B, g.n).astype(np.float32)
        o = O.bp_weighted_grad(H, llr, iters, 20.0, w, gp, graph=g)
        dw = {k: torch.as_tensor(v).cuda() for k, v in w.items()}
        dw.update(iterations=iters, stride=mdv)
        prob, tape = code.train_forward(torch.as_tensor(llr).cuda(), dw, 20.0)
        out = code.train_backward(torch.as_tensor(llr).cuda(), dw, 20.0, tape, torch.as_tensor(gp).cuda())
        assert np.abs(prob.cpu().numpy() - o["prob"]).max() < 2e-5
        assert _rel(out["grad_llr"].cpu().numpy(), o["grad_llr"]) < 2e-4
        assert _rel(out["wf_edge"].cpu().numpy(), o["g_wf_edge"]) < 2e-4
        if iters:
            assert _rel(out["w_edge"].cpu().numpy(), o["g_w_edge"]) < 2e-4
    # empty batch: outputs exist, weight gradients are zero
    dw0 = {k: torch.ones(s, device="cuda") for k, s in (("w_edge", (1, g.E, mdv)), ("w_llr", (1, g.n)), ("wf_edge", (g.E,)), ("wf_llr", (g.n,)))}
    dw0.update(iterations=1, stride=mdv)
    prob, tape = code.train_forward(torch.zeros(0, g.n, device="cuda"), dw0, 20.0)
    out = code.train_backward(torch.zeros(0, g.n, device="cuda"), dw0, 20.0, tape, torch.zeros(0, g.n, device="cuda"))
    assert prob.shape == (0, g.n) and float(out["w_edge"].abs().sum()) == 0.0
    with pytest.raises(ValueError):
        code.train_forward(torch.zeros(4, g.n + 1, device="cuda"), dw0, 20.0)
    with pytest.raises(ValueError):
        code.train_forward(torch.zeros(4, g.n), dw0, 20.0)                       # CPU tensor: no CPU fallback
    with pytest.raises(ValueError):
        code.train_backward(torch.zeros(4, g.n, device="cuda"), dw0, 20.0, torch.zeros(2, g.E, 3, device="cuda"), torch.zeros(4, g.n, device="cuda"))
    m = BeliefPropagation(H, 2, update="minsum").cuda()
    with pytest.raises(ValueError):
        m(None, torch.zeros(4, 64, device="cuda"), 20.0)                         # training mode is sum-product only
    assert m.eval()(None, torch.zeros(4, 64, device="cuda"), 20.0).shape == (4, 64)


@pytest.mark.gpu
def test_both_training_variants_and_isolated_variable():
    """The warp-per-codeword pair (batches up to 16384) and the thread-per-codeword pair (larger) give the same result;
    H here has a variable without any check (a degree the compile-time-degree switch has no case for)."""
    import torch
    import bp_oracle as O
    from ldpc_b200.decoder import LdpcCode
    rng = np.random.RandomState(3)
    H = (rng.rand(12, 24) < 0.2).astype(np.uint8)
    H[:, 5] = 0                                               # isolated variable
    H[np.arange(12), np.arange(12)] = 1                       # no empty check
    H[:, 7] = 0; H[3, 7] = 1                                  # a degree-1 variable
    g = O.Graph(H)
    code = LdpcCode(H)
    mdv, iters = int(g.dv.max()), 2
    w = dict(w_edge=(0.5 + rng.rand(iters, g.E, mdv)).astype(np.float32), w_llr=(0.5 + rng.rand(iters, g.n)).astype(np.float32),
             wf_edge=(0.5 + rng.rand(g.E)).astype(np.float32), wf_llr=(0.5 + rng.rand(g.n)).astype(np.float32))
    dw = {k: torch.as_tensor(v).cuda() for k, v in w.items()}
    dw.update(iterations=iters, stride=mdv)
    small = 50
    llr = (rng.randn(small, g.n)).astype(np.float32)
    gp = rng.randn(small, g.n).astype(np.float32)
    o = O.bp_weighted_grad(H, llr, iters, 20.0, w, gp, graph=g)
    big = 16384 + 64                                           # same rows tiled past the variant threshold
    reps = -(-big // small)
    llr_b, gp_b = np.tile(llr, (reps, 1))[:big], np.tile(gp, (reps, 1))[:big]
    res = {}
    for tag, L, Gp in (("warp", llr, gp), ("thread", llr_b, gp_b)):
        prob, tape = code.train_forward(torch.as_tensor(L).cuda(), dw, 20.0)
        out = code.train_backward(torch.as_tensor(L).cuda(), dw, 20.0, tape, torch.as_tensor(Gp).cuda())
        res[tag] = (prob.cpu().numpy(), out["grad_llr"].cpu().numpy())
        assert np.abs(res[tag][0][:small] - o["prob"]).max() < 2e-5, tag
        assert _rel(res[tag][1][:small], o["grad_llr"]) < 2e-4, tag
        if tag == "warp":
            for k in ("w_edge", "w_llr", "wf_edge", "wf_llr"):
                assert _rel(out[k].cpu().numpy(), o["g_" + k]) < 2e-4, k
    assert np.array_equal(res["warp"][0], res["thread"][0][:small])           # same forward arithmetic, bit for bit


def test_joint_module_and_state_dict_on_cpu():
    """Host logic only (no GPU): Joint's three constructor forms, the sparse parameters, loading a reference
    state_dict (with the DataParallel / Joint prefixes) and exporting it again in the reference's dense layout."""
    import torch
    from bp.bp import BeliefPropagation
    from bp.masking import generate_masks
    from bp.parity import H
    from nn.joint import Joint
    iters = int(G["iters"])
    j = Joint(32, 3.16, H, iters)
    assert j.layer_size() == 96 and Joint(H, iters).BP.iterations == iters
    mask_c, mask_v, mask_v_final, llr_expander = generate_masks(H)
    assert Joint(32, 3.16, mask_v, mask_c, mask_v_final, llr_expander, iters).BP.layer_size() == 96
    names = {k for k, _ in j.named_parameters()}
    assert {"BP.w_edge", "BP.w_llr", "BP.wf_edge", "BP.wf_llr", "LLRest.final.weight"} <= names
    assert all(p.requires_grad for p in j.BP.parameters()) and j.BP._all_ones()
    bp = BeliefPropagation(H, iters)
    bp.load_state_dict({"module.BP." + k: torch.tensor(v) for k, v in STATE.items()})
    assert not bp._all_ones() and tuple(bp.w_edge.shape) == (iters, 96, 2)
    back = bp.reference_state_dict()
    for k, v in STATE.items():
        assert np.array_equal(back[k].numpy(), v), k
    bp2 = BeliefPropagation(H, iters)
    bp2.load_state_dict(bp.state_dict())                      # its own (sparse) format
    assert all(torch.equal(a, b) for a, b in zip(bp.parameters(), bp2.parameters()))
    with pytest.raises(TypeError):
        Joint(32, 3.16, H)
