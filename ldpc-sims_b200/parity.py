"""Flat alias of bp/parity.py (ber_test.py:5-11 imports `parity`)."""
from bp.parity import H, G, P, block_size, rate  # noqa: F401
