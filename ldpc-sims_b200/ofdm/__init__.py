"""Drop-in for the reference's ``pytorch/ofdm`` package (link simulator + decode_bits)."""
