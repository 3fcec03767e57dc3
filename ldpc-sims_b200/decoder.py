"""Stale flat-module spelling kept for old scripts: `from decoder import decoder`
(evaluate.py:9,117; evaluate_quantized_grid.py:9,144; joint_test.py:10,136)."""
from ofdm.ofdm_functions import decode_bits as decoder, decode_bits  # noqa: F401
