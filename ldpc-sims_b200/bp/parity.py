"""Default code, same names/shapes/dtypes as the reference's bp/parity.py:7-47
(H int64 [32,64] PEG code, P = H[:, :32], G float64 [64,32] systematic generator)."""
from ldpc_b200.codes import peg_64_32

H, G = peg_64_32()
P = H[:, 0:32]
block_size = 64
rate = 1 / 2
