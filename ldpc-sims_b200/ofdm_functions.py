"""Flat alias of ofdm/ofdm_functions.py (evaluate_quantized.py:9 `from ofdm_functions import *`)."""
from ofdm.ofdm_functions import *  # noqa: F401,F403
from ofdm.ofdm_functions import decode_bits, decoder, gen_data, gen_qdata  # noqa: F401
