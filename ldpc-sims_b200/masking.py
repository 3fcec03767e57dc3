"""Flat alias of bp/masking.py (ber_test.py:9 `from masking import genMasks`)."""
from bp.masking import generate_masks, genMasks, masks_to_H  # noqa: F401
