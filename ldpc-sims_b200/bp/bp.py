"""Drop-in BeliefPropagation (reference bp/bp.py:19-62) on the B200-native decoder.

Same constructor and call signature:

    model = BeliefPropagation(H, iterations)          # nn.Module, .eval(), .to(device), DataParallel-able
    prob  = model(x, llr, clamp_value)                # P(bit=1) [B,n] float32
    E     = model.layer_size()

Differences, all additive: keyword-only extras (update='sp'|'minsum'|'nms'|'oms', param,
warm_start, return_llr / return_hard / return_syndrome) and the stale 5-argument
constructor BeliefPropagation(mask_vc, mask_cv, mask_v_final, llr_expander, iterations)
(ber_test.py:46).  ``x`` (initial C->V messages, check-major) is zeros in every reference
caller (ofdm_functions.py:157); its values are only read when warm_start=True.
"""
import numpy as np
import torch
import torch.nn as nn

from ldpc_b200.codes import (EdgeTables, _weight_slots, reference_state_from_sparse_weights,
                              sparse_weights_from_reference_state)
from ldpc_b200.decoder import LdpcCode
from .masking import generate_masks, masks_to_H  # noqa: F401

__all__ = ["BeliefPropagation", "pyd", "generate_masks"]


def pyd(tensor):                                   # bp/bp.py:16-17
    return tensor.detach().cpu().numpy()


class _WeightedBP(torch.autograd.Function):
    """prob = BP(llr; weights) with the native forward-with-tape / sparse backward pair (ldpc_bp_train_forward /
    ldpc_bp_train_backward) in place of the reference's dense autograd Functions (bp/bp_vc.py:6-58, bp/bp_cv.py:6-96)."""

    @staticmethod
    def forward(ctx, llr, w_edge, w_llr, wf_edge, wf_llr, code, clamp_value, x0):
        w = dict(w_edge=w_edge, w_llr=w_llr, wf_edge=wf_edge, wf_llr=wf_llr, iterations=w_edge.shape[0], stride=w_edge.shape[2])
        prob, tape = code.train_forward(llr, w, clamp_value, x0)
        ctx.code, ctx.clamp_value = code, clamp_value
        ctx.save_for_backward(llr, w_edge, w_llr, wf_edge, wf_llr, tape)
        return prob

    @staticmethod
    def backward(ctx, grad_prob):
        llr, w_edge, w_llr, wf_edge, wf_llr, tape = ctx.saved_tensors
        w = dict(w_edge=w_edge, w_llr=w_llr, wf_edge=wf_edge, wf_llr=wf_llr, iterations=w_edge.shape[0], stride=w_edge.shape[2])
        g = ctx.code.train_backward(llr, w, ctx.clamp_value, tape, grad_prob)
        return g["grad_llr"], g["w_edge"], g["w_llr"], g["wf_edge"], g["wf_llr"], None, None, None


_PARAMS = ("w_edge", "w_llr", "wf_edge", "wf_llr")


class BeliefPropagation(nn.Module):
    """The reference's trainable weights (bp/bp_vc.py:101-107, one set per iteration layer and one for the final layer,
    bp/bp.py:26-39) are nn.Parameters here too, initialised to ones and requiring gradients, but SPARSE:
    w_edge [iterations, E, max_dv], w_llr [iterations, n], wf_edge [E], wf_llr [n] (layout: ldpc_decode_weighted).
    In training mode with autograd enabled the forward keeps a tape and .backward() runs the native sparse backward;
    otherwise (model.eval() / torch.no_grad(), as ofdm_functions.py:147 does) the inference kernels run - the
    unweighted fast paths while every weight is still 1."""

    def __init__(self, H, iterations, *legacy, update="sp", param=1.0, warm_start=False, qc_Z="auto"):
        super().__init__()
        if len(legacy) == 3:                       # (mask_vc, mask_cv, mask_v_final, llr_expander, iterations)
            mask_v, mask_c, mask_v_final, llr_expander, iterations = H, iterations, legacy[0], legacy[1], legacy[2]
            H = masks_to_H(_np(mask_c), _np(mask_v), _np(mask_v_final), _np(llr_expander))
        elif legacy:
            raise TypeError("BeliefPropagation(H, iterations) or the legacy 5-argument form")
        self._H = (np.asarray(_np(H)) != 0).astype(np.uint8)
        self.iterations = int(iterations)
        self.update, self.param, self.warm_start, self._qc_Z = update, float(param), bool(warm_start), qc_Z
        self._tables = EdgeTables.from_H(self._H)
        self.layer_size_val = int(self._tables.E)
        E, n, mdv = self._tables.E, self._tables.n, self._tables.max_dv
        self.w_edge = nn.Parameter(torch.ones(self.iterations, E, mdv))
        self.w_llr = nn.Parameter(torch.ones(self.iterations, n))
        self.wf_edge = nn.Parameter(torch.ones(E))
        self.wf_llr = nn.Parameter(torch.ones(n))
        self._codes = {}                           # device index -> LdpcCode (one native handle per GPU)
        self._trivial = (None, True)               # (parameter versions, every weight == 1)

    def _code(self, device):
        key = device.index if device.index is not None else torch.cuda.current_device()
        if key not in self._codes:
            self._codes[key] = LdpcCode(self._H, qc_Z=self._qc_Z, device=torch.device("cuda", key))
        return self._codes[key]

    def _all_ones(self):
        ps = [getattr(self, k) for k in _PARAMS]
        ver = tuple((p.data_ptr(), p._version) for p in ps)
        if self._trivial[0] != ver:
            used = torch.as_tensor(_weight_slots(self._tables)[2], device=self.w_edge.device)
            ones = bool((self.w_edge.detach()[:, used] == 1).all()) and all(bool((p.detach() == 1).all()) for p in ps[1:])
            self._trivial = (ver, ones)
        return self._trivial[1]

    def forward(self, x, llr, clamp_value, *, return_llr=False, return_hard=False, return_syndrome=False):
        src = llr.device
        dev = src if src.type == "cuda" else torch.device("cuda", torch.cuda.current_device())
        if llr.dim() != 2:
            raise ValueError("llr must be [B,n]")
        if x is not None and tuple(x.shape) != (llr.shape[0], self.layer_size_val):
            raise ValueError(f"x must be [B,{self.layer_size_val}] (check-major C->V messages)")
        code = self._code(dev)
        ps = [getattr(self, k) for k in _PARAMS]
        if self.training and torch.is_grad_enabled() and (llr.requires_grad or any(p.requires_grad for p in ps)):
            if self.update not in ("sp", "tanh", "sum-product"):
                raise ValueError("the training path is sum-product (the reference's rule); call .eval() for min-sum decoding")
            if return_llr or return_hard or return_syndrome:
                raise ValueError("return_llr / return_hard / return_syndrome are inference outputs: call under torch.no_grad() or .eval()")
            x0 = x.detach().to(dev) if (self.warm_start and x is not None) else None
            prob = _WeightedBP.apply(llr.to(dev).float(), *[p.to(dev) for p in ps], code, float(clamp_value), x0)
            return prob.to(src)
        want = ["prob"]
        if return_llr: want.append("llr_post")
        if return_hard: want.append("hard")
        if return_syndrome: want.append("syndrome")
        if not self._all_ones():                   # trained weights: weighted kernels
            if self.warm_start and x is not None:
                raise NotImplementedError("warm_start with trained weights is only available in training mode "
                                          "(ldpc_decode_weighted takes no initial messages)")
            w = {k: p.detach().to(dev) for k, p in zip(_PARAMS, ps)}
            w.update(iterations=self.iterations, stride=int(self.w_edge.shape[2]))
            out = code.decode_weighted(llr.detach().to(dev), w, clamp_value, update=self.update, param=self.param, want=tuple(want))
        else:
            out = code.decode(llr.detach().to(dev), self.iterations, clamp_value, update=self.update, param=self.param,
                              x0=(x.detach().to(dev) if (self.warm_start and x is not None) else None), want=tuple(want))
        res = [out[k].to(src) for k in want]
        return res[0] if len(res) == 1 else tuple(res)

    def layer_size(self):
        return self.layer_size_val

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        """nn.Module loads a PARENT's state_dict by recursing through this hook (never through a child's
        load_state_dict), so the conversion of a REFERENCE BeliefPropagation state (bp/bp.py:26-39: per-iteration dense
        `layers.{i}.0.input_weight` [E,E] / `llr_weight` [1,n] and `final_layer.0.*`, plus the mask buffers) lives here: the
        reference pattern `nn.DataParallel(Joint(...)).load_state_dict(ckpt['model_state_dict'])` (joint_evaluate.py:62-67,
        keys `module.BP.layers...`) then loads into the sparse tables.  The dense tensors are converted, the masks dropped."""
        ref_keys = [k for k in state_dict if k.startswith(prefix + "layers.") or k.startswith(prefix + "final_layer.")]
        if ref_keys:
            st = {k[len(prefix):]: state_dict[k] for k in ref_keys}
            need = [f"layers.{i}.0.{w}" for i in range(self.iterations) for w in ("input_weight", "llr_weight")]
            need += ["final_layer.0.input_weight", "final_layer.0.llr_weight"]
            lacking = [k for k in need if k not in st]
            if lacking:                                    # e.g. a checkpoint trained with another iteration count: reported as
                missing_keys.extend(prefix + k for k in lacking)   # missing keys (strict loading raises, non-strict returns them)
            else:
                w = sparse_weights_from_reference_state(self._tables, st, self.iterations)
                for k in _PARAMS:
                    state_dict[prefix + k] = torch.as_tensor(w[k])
            for k in ref_keys:                             # consumed (or rejected above): never "unexpected"
                del state_dict[k]
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs)

    def load_state_dict(self, state_dict, strict=True, **kw):
        """Accepts this module's own state_dict (the four sparse tables) or the state_dict of a REFERENCE
        BeliefPropagation (bp/bp.py:26-39), optionally with the 'module.' / 'BP.' prefixes of DataParallel /
        nn/joint.py when it is loaded DIRECTLY into this module; nested loading (through Joint / DataParallel) goes
        through _load_from_state_dict with the parent's own prefixes."""
        st = {}
        for k, v in state_dict.items():
            for pre in ("module.", "BP."):
                if k.startswith(pre):
                    k = k[len(pre):]
            st[k] = v
        return super().load_state_dict(st, strict=strict, **kw)

    def invalidate(self):
        """Forget the cached "every weight is 1" decision (use after editing the weights through `.data`, which bumps
        neither data_ptr nor _version)."""
        self._trivial = (None, True)

    def reference_state_dict(self):
        """The trainable tensors in the REFERENCE's dense layout and key names (layers.{i}.0.input_weight [E,E], ...),
        for checkpoints the reference's scripts can load (masks omitted: load with strict=False there)."""
        w = {k: getattr(self, k).detach().cpu().numpy() for k in _PARAMS}
        return {k: torch.as_tensor(v) for k, v in reference_state_from_sparse_weights(self._tables, w).items()}

    def __getstate__(self):                        # DataParallel.replicate / deepcopy: never copy native handles
        d = self.__dict__.copy()
        d["_codes"] = {}
        return d


def _np(a):
    return a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
