"""Host side of the B200-native LDPC belief-propagation / OFDM link-simulation hot path.

Everything here is plumbing around libldpc_b200.so (hand-written sm_100a CUDA behind the
C ABI in include/ldpc_b200.h); torch is used for device memory, streams and
torch.distributed only.
"""
from .codes import EdgeTables, QCCode, peg_64_32, ieee80211n_1944_r12, detect_qc, systematic_generator  # noqa: F401
from .decoder import LdpcCode, decode_host  # noqa: F401
