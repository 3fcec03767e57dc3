#!/usr/bin/env python3
"""Mint tests/golden/bp_grad.npz from the UNMODIFIED reference (build container only) and pin
oracle/bp_oracle.py::bp_weighted_grad against it.

The reference trains BeliefPropagation's weights (and the demapper in front of it) with BCE through its hand-written
autograd Functions (bp/bp_vc.py:34-58, bp/bp_cv.py:57-91; loop ofdm/ofdm_nn.py:257-396).  Here the reference model
(CPU, fp32, seeded random weights) runs forward and its own .backward() for a BCE loss; the gradients it leaves on
llr, input_weight, llr_weight are stored in the sparse layout of ldpc_bp_train_backward.  Inputs are scaled so that
no product saturates (|p| < 1 - 1e-7 and |message| < clamp everywhere): the regime in which the reference's backward
and the true derivative coincide.  A second, saturating batch only PRINTS how far the reference's backward is from the
derivative of its own forward there (informational; the CUDA path follows the derivative).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
np.complex = complex
np.float = float
sys.path.insert(0, "/root/reference/pytorch")
sys.path.insert(0, os.path.join(ROOT, "oracle"))

from bp.bp import BeliefPropagation                      # noqa: E402  (the reference)
from bp.parity import H                                   # noqa: E402
import bp_oracle as O                                     # noqa: E402


def reference_grads(m, llr, target, clamp):
    m.zero_grad()
    L = llr.clone().requires_grad_(True)
    prob = m(torch.zeros(llr.shape[0], m.layer_size()), L, clamp)
    gp = torch.autograd.grad(torch.nn.functional.binary_cross_entropy(prob, target), prob, retain_graph=True)[0]
    prob.backward(gp)
    return prob.detach(), gp.detach(), L.grad.detach(), {k: p.grad.detach().clone() for k, p in m.named_parameters() if p.grad is not None}


def sparse(g, grads, iters):
    st = {k: v for k, v in grads.items()}
    w = O.weights_from_reference_state(g, st, iters)
    # weights_from_reference_state fills unused entries with 1: zero them (they carry no gradient)
    mdv = int(g.dv.max())
    vm_var = np.repeat(np.arange(g.n), g.dv)
    pos = np.arange(g.E) - g.var_ptr[vm_var]
    for j in range(mdv):
        w["w_edge"][:, ~((g.dv[vm_var] > j) & (pos != j)), j] = 0
    return w


def main():
    torch.manual_seed(5)
    iters, clamp = 3, 20.0
    m = BeliefPropagation(H, iters)
    params = dict(m.named_parameters())
    with torch.no_grad():
        for name, p in params.items():
            if name.endswith("input_weight"):
                p.copy_(params[name.replace("input_weight", "mask")] * (0.5 + torch.rand_like(p)))
            elif name.endswith("llr_weight"):
                p.copy_(0.5 + torch.rand_like(p))
    g = O.Graph(H)
    state = {k: v.detach().clone() for k, v in m.state_dict().items()}
    w = O.weights_from_reference_state(g, state, iters)
    B = 96
    target = torch.randint(0, 2, (B, 64)).float()
    out = {}
    for tag, scale in (("soft", 0.8), ("sat", 6.0)):
        llr = ((2 * target - 1) * scale + torch.randn(B, 64) * np.sqrt(2 * scale)).float()
        prob, gp, gl, gw = reference_grads(m, llr, target, clamp)
        ref = sparse(g, gw, iters)
        o = O.bp_weighted_grad(H, llr.numpy(), iters, clamp, w, gp.numpy(), graph=g)
        errs = {}
        for k, a, b in (("grad_llr", gl.numpy(), o["grad_llr"]), ("g_w_edge", ref["w_edge"], o["g_w_edge"]), ("g_w_llr", ref["w_llr"], o["g_w_llr"]),
                        ("g_wf_edge", ref["wf_edge"], o["g_wf_edge"]), ("g_wf_llr", ref["wf_llr"], o["g_wf_llr"])):
            errs[k] = float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
        print(tag, "max |reference backward - oracle| / max|oracle|:", {k: f"{v:.1e}" for k, v in errs.items()},
              "max |dP|", float(np.abs(prob.numpy() - o["prob"]).max()))
        if tag == "soft":
            assert max(errs.values()) < 2e-4, errs          # reference is fp32, oracle float64
            keep = [f"layers.{i}.0.{k}" for i in range(iters) for k in ("input_weight", "llr_weight")] + \
                   ["final_layer.0.input_weight", "final_layer.0.llr_weight"]
            out.update({"s_" + k: state[k].numpy() for k in keep})
            out.update(llr=llr.numpy(), target=target.numpy().astype(np.uint8), prob=prob.numpy(), grad_prob=gp.numpy(), grad_llr=gl.numpy(),
                       g_w_edge=ref["w_edge"], g_w_llr=ref["w_llr"], g_wf_edge=ref["wf_edge"], g_wf_llr=ref["wf_llr"],
                       iters=np.int64(iters), clamp=np.float64(clamp))
    p = os.path.join(ROOT, "tests", "golden", "bp_grad.npz")
    np.savez_compressed(p, **out)
    print(f"wrote {p} ({os.path.getsize(p)} bytes)")


if __name__ == "__main__":
    main()
