#!/usr/bin/env python3
"""Mint tests/golden/bp_weighted.npz from the UNMODIFIED reference (build container only).

The reference's BeliefPropagation carries trainable weights (bp/bp_vc.py:101-107: input_weight [E,E] masked,
llr_weight [1,n]; one set per iteration layer and one for the final layer, bp/bp.py:26-39).  No trained BP
checkpoint ships with the reference, so the weights are a SEEDED random perturbation of the masks; the reference
model itself (CPU, fp32) produces the outputs.  Also pins oracle/bp_oracle.py's weighted mode against them.
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


def main():
    torch.manual_seed(3)
    iters, clamp = 3, 20.0
    m = BeliefPropagation(H, iters).eval()
    params = dict(m.named_parameters())
    with torch.no_grad():
        for name, p in params.items():
            if name.endswith("input_weight"):
                p.copy_(params[name.replace("input_weight", "mask")] * (0.5 + torch.rand_like(p)))
            elif name.endswith("llr_weight"):
                p.copy_(0.5 + torch.rand_like(p))
    llr = (torch.randn(256, 64) * 3).float()
    with torch.no_grad():
        prob = m(torch.zeros(256, m.layer_size()), llr, clamp).numpy()
    state = m.state_dict()
    g = O.Graph(H)
    w = O.weights_from_reference_state(g, state, iters)
    o = O.bp_decode(H, llr.numpy(), iters, clamp, weights=w)
    err = float(np.abs(o["prob"] - prob).max())
    hard = np.round(prob).astype(np.uint8)
    assert err < 5e-7 and np.array_equal(o["hard"], hard), err
    keep = [f"layers.{i}.0.{k}" for i in range(iters) for k in ("input_weight", "llr_weight")] + \
           ["final_layer.0.input_weight", "final_layer.0.llr_weight"]
    out = {"s_" + k: state[k].numpy() for k in keep}
    out.update(llr=llr.numpy(), prob=prob, hard=np.packbits(hard, axis=1), iters=np.int64(iters), clamp=np.float64(clamp))
    p = os.path.join(ROOT, "tests", "golden", "bp_weighted.npz")
    np.savez_compressed(p, **out)
    print(f"weighted BP: oracle vs reference max |dP| = {err:.2e}, hard bits equal; wrote {p} ({os.path.getsize(p)} bytes)")


if __name__ == "__main__":
    main()
