"""Host logic of the drop-in modules that needs no GPU: checkpoint loading through parent modules (the reference's
`nn.DataParallel(Joint(...)).load_state_dict(ckpt['model_state_dict'])`, joint_evaluate.py:62-67) and the bookkeeping
the native-handle caches rely on when nn.DataParallel replicates a module."""
import numpy as np
import torch
import torch.nn as nn

from bp.bp import BeliefPropagation, _weight_slots
from bp.parity import H
from nn.joint import Joint
from nn.llr import LLRestimator_withSNR


def _random_bp(iters, seed):
    m = BeliefPropagation(H, iters)
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for p in m.parameters():
            p.mul_(0.5 + torch.rand(p.shape, generator=g))
    return m


def _used(m):
    return torch.as_tensor(_weight_slots(m._tables)[2])


def _same_weights(a, b):
    u = _used(a)
    ok = torch.equal(a.w_edge.detach()[:, u], b.w_edge.detach()[:, u])
    return ok and all(torch.equal(getattr(a, k).detach(), getattr(b, k).detach()) for k in ("w_llr", "wf_edge", "wf_llr"))


def test_reference_state_loads_through_joint_and_dataparallel_prefixes():
    src = _random_bp(2, 1)
    ref = src.reference_state_dict()                       # the reference's dense layout and key names
    ref["layers.0.0.mask"] = torch.zeros(4)                # the reference also stores its mask buffers: must be ignored
    j = Joint(32, 1.0, H, 2)
    sd = {k: v.clone() for k, v in j.state_dict().items() if not k.startswith("BP.")}
    sd.update({"BP." + k: v for k, v in ref.items()})
    res = j.load_state_dict(sd)                            # strict: no missing / unexpected keys
    assert not res.missing_keys and not res.unexpected_keys
    assert _same_weights(j.BP, src)
    # DataParallel prefix on top (keys 'module.BP.layers...')
    class Wrap(nn.Module):
        def __init__(self, mod):
            super().__init__()
            self.module = mod
    w = Wrap(Joint(32, 1.0, H, 2))
    res = w.load_state_dict({"module." + k: v for k, v in sd.items()})
    assert not res.missing_keys and not res.unexpected_keys
    assert _same_weights(w.module.BP, src)
    # direct load with the stale prefixes the reference's own key-rewriting leaves behind
    d = BeliefPropagation(H, 2)
    d.load_state_dict({"module.BP." + k: v for k, v in ref.items()})
    assert _same_weights(d, src)
    # the module's own (sparse) state_dict still round-trips
    e = BeliefPropagation(H, 2)
    e.load_state_dict(src.state_dict())
    assert _same_weights(e, src)


def test_reference_state_for_another_iteration_count_is_rejected():
    ref = _random_bp(2, 2).reference_state_dict()
    try:
        BeliefPropagation(H, 3).load_state_dict(ref)
    except RuntimeError as e:
        assert "layers.2.0.input_weight" in str(e)
    else:
        raise AssertionError("a 2-iteration reference state must not load into a 3-iteration module")
    res = BeliefPropagation(H, 3).load_state_dict(ref, strict=False)      # non-strict: reported, weights untouched
    assert {"w_edge", "w_llr", "wf_edge", "wf_llr", "layers.2.0.input_weight"} <= set(res.missing_keys)


def test_native_chain_reads_replica_tensors():
    """nn.DataParallel replicas have EMPTY parameters() (replicate() stores plain tensors): the handle cache must key on
    the tensors the forward uses, and the training gate on their requires_grad."""
    m = LLRestimator_withSNR(32)
    rep = m._replicate_for_data_parallel()                  # what torch.nn.parallel.replicate does, module by module:
    for name, child in m._modules.items():                  # replicas of the children with EMPTY _parameters ...
        rc = child._replicate_for_data_parallel()
        rep._modules[name] = rc
        for pn, p in child._parameters.items():             # ... and the broadcast copies as plain tensor attributes
            if p is not None:
                setattr(rc, pn, p.detach().clone().requires_grad_(p.requires_grad))
    assert len(list(rep.parameters())) == 0
    ts = rep._tensors()
    assert len(ts) == 8 and all(isinstance(t, torch.Tensor) for t in ts)
    c0 = m._checksum(m._tensors())
    assert rep._checksum(ts) == c0                          # same values -> same key on the replica
    with torch.no_grad():
        m.hidden2.weight.data.add_(1.0)                     # a `.data` edit: _version is not bumped
    assert m._checksum(m._tensors()) != c0                  # ... but the key changes
    assert any(t.requires_grad for t in ts)
