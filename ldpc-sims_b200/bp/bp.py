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

from ldpc_b200.decoder import LdpcCode
from .masking import generate_masks, masks_to_H  # noqa: F401

__all__ = ["BeliefPropagation", "pyd", "generate_masks"]


def pyd(tensor):                                   # bp/bp.py:16-17
    return tensor.detach().cpu().numpy()


class BeliefPropagation(nn.Module):
    def __init__(self, H, iterations, *legacy, update="sp", param=1.0, warm_start=False, qc_Z=0):
        super().__init__()
        if len(legacy) == 3:                       # (mask_vc, mask_cv, mask_v_final, llr_expander, iterations)
            mask_v, mask_c, mask_v_final, llr_expander, iterations = H, iterations, legacy[0], legacy[1], legacy[2]
            H = masks_to_H(_np(mask_c), _np(mask_v), _np(mask_v_final), _np(llr_expander))
        elif legacy:
            raise TypeError("BeliefPropagation(H, iterations) or the legacy 5-argument form")
        self._H = (np.asarray(_np(H)) != 0).astype(np.uint8)
        self.iterations = int(iterations)
        self.update, self.param, self.warm_start, self._qc_Z = update, float(param), bool(warm_start), int(qc_Z)
        self.layer_size_val = int(self._H.sum())
        self._codes = {}                           # device index -> LdpcCode (one native handle per GPU)
        self._ref_state = None                     # a reference state_dict with trained weights (load_state_dict)
        self._weights = {}                         # device index -> sparse weight tables

    def _code(self, device):
        key = device.index if device.index is not None else torch.cuda.current_device()
        if key not in self._codes:
            self._codes[key] = LdpcCode(self._H, qc_Z=self._qc_Z, device=torch.device("cuda", key))
        return self._codes[key]

    def forward(self, x, llr, clamp_value, *, return_llr=False, return_hard=False, return_syndrome=False):
        src = llr.device
        dev = src if src.type == "cuda" else torch.device("cuda", torch.cuda.current_device())
        if llr.dim() != 2:
            raise ValueError("llr must be [B,n]")
        if x is not None and tuple(x.shape) != (llr.shape[0], self.layer_size_val):
            raise ValueError(f"x must be [B,{self.layer_size_val}] (check-major C->V messages)")
        code = self._code(dev)
        want = ["prob"]
        if return_llr: want.append("llr_post")
        if return_hard: want.append("hard")
        if return_syndrome: want.append("syndrome")
        if self._ref_state is not None:            # trained weights (bp_vc.py:101-107): weighted generic kernel
            key = dev.index if dev.index is not None else torch.cuda.current_device()
            if key not in self._weights:
                self._weights[key] = code.sparse_weights(self._ref_state, self.iterations)
            if self._weights[key] is None:
                self._ref_state = None             # every weight is 1: nothing to apply
        if self._ref_state is not None:
            out = code.decode_weighted(llr.detach().to(dev), self._weights[key], clamp_value, update=self.update,
                                       param=self.param, want=tuple(want))
        else:
            out = code.decode(llr.detach().to(dev), self.iterations, clamp_value, update=self.update, param=self.param,
                              x0=(x.detach().to(dev) if (self.warm_start and x is not None) else None), want=tuple(want))
        res = [out[k].to(src) for k in want]
        return res[0] if len(res) == 1 else tuple(res)

    def layer_size(self):
        return self.layer_size_val

    def load_state_dict(self, state_dict, strict=True):
        """Accepts the state_dict of a REFERENCE BeliefPropagation (bp/bp.py:26-39, optionally with the 'module.' /
        'BP.' prefixes of DataParallel / nn/joint.py): its trainable input_weight / llr_weight tensors are converted to
        sparse tables and applied by the weighted kernel; the dense masks are not kept."""
        st = {}
        for k, v in state_dict.items():
            for pre in ("module.", "BP."):
                if k.startswith(pre):
                    k = k[len(pre):]
            st[k] = v
        need = [f"layers.{i}.0.{w}" for i in range(self.iterations) for w in ("input_weight", "llr_weight")]
        need += ["final_layer.0.input_weight", "final_layer.0.llr_weight"]
        missing = [k for k in need if k not in st]
        if missing and strict:
            raise KeyError(f"not a reference BeliefPropagation state_dict for {self.iterations} iterations: missing {missing[:3]}...")
        if not missing:
            self._ref_state = {k: st[k] for k in need}
            self._weights = {}
        return self

    def __getstate__(self):                        # DataParallel.replicate / deepcopy: never copy native handles
        d = self.__dict__.copy()
        d["_codes"] = {}
        d["_weights"] = {}
        return d


def _np(a):
    return a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
