"""Drop-in Joint model (reference nn/joint.py:12-30, nn/joint_connected.py:13-26): the NN demapper feeding the
belief-propagation decoder, trained end to end (ofdm/ofdm_nn.py:257-396).

    Joint(ofdm_size, snr, H, iterations)                                             # this package
    Joint(ofdm_size, snr, mask_vc, mask_cv, mask_v_final, llr_expander, iterations)  # as ofdm_nn.py:278 calls it
    Joint(H, iterations, ofdm_size=32, snr=1.0)                                      # nn/joint.py:13 (its ofdm_size / snr are undefined globals there)
"""
import torch.nn as nn

from bp.bp import BeliefPropagation
from .llr import LLRestimator

__all__ = ["Joint"]


class Joint(nn.Module):
    def __init__(self, *args, ofdm_size=32, snr=1.0):
        super().__init__()
        if len(args) == 2:
            H, iterations = args
            bp_args = (H, iterations)
        elif len(args) == 4:
            ofdm_size, snr, H, iterations = args
            bp_args = (H, iterations)
        elif len(args) == 7:
            ofdm_size, snr = args[:2]
            bp_args = args[2:]
        else:
            raise TypeError("Joint(ofdm_size, snr, H, iterations) or the reference's 2- / 7-argument forms")
        self.LLRest = LLRestimator(ofdm_size, snr)
        self.BP = BeliefPropagation(*bp_args)
        self.layer_size_val = self.BP.layer_size()

    def forward(self, signal, x, clamp_value):
        return self.BP(x, self.LLRest(signal), clamp_value)

    def layer_size(self):
        return self.layer_size_val
