"""Code library and Tanner-graph compiler (host side, numpy only).

Replaces the reference's dense-mask graph compile (``pytorch/bp/masking.py:12-147``)
and its default-code module (``pytorch/bp/parity.py:7-47``) with sparse edge tables:

* ``EdgeTables``   - CSR by check / CSC by variable + the check-major <-> variable-major
                     edge permutation.  Edge numbering is the reference's: check-major ids
                     are the row-major non-zeros of H (``masking.py:85-88``), variable-major
                     ids the column-major non-zeros (``masking.py:92-95``).
* ``peg_64_32``    - the reference's default (64,32) PEG code, rebuilt from its sparse
                     structure; returns H and the systematic generator G with the same
                     shapes/dtypes as ``parity.H`` / ``parity.G``.
* ``ieee80211n_1944_r12`` - IEEE 802.11n n=1944, R=1/2, Z=81 (not in the reference;
                     prototype matrix from SURVEY.md appendix B).
* encoders         - ``systematic_generator`` (dense GF(2) elimination, any H whose
                     parity part is invertible) and ``QCCode.encode`` (linear-time
                     dual-diagonal back-substitution).
"""
from __future__ import annotations

import dataclasses
import numpy as np

# --------------------------------------------------------------------------------------
# default code: (64,32) PEG, H = [P | I32].  Check r touches information bit r//2, one
# information bit in 16..31 (table below) and parity bit 32+r (reference parity.py:7-40).
# --------------------------------------------------------------------------------------
_PEG_SECOND = (16, 17, 16, 18, 17, 19, 18, 20, 19, 21, 20, 22, 21, 23, 22, 24,
               23, 25, 24, 26, 25, 27, 26, 28, 27, 29, 28, 30, 29, 31, 30, 31)


def peg_64_32():
    """Return (H int64 [32,64], G float64 [64,32]) identical to reference bp/parity.py."""
    H = np.zeros((32, 64), dtype=np.int64)
    for r in range(32):
        H[r, r // 2] = 1
        H[r, _PEG_SECOND[r]] = 1
        H[r, 32 + r] = 1
    P = H[:, 0:32]
    # parity.py:44  G = [I ; P]  (64 x 32), codeword = G u, info bits first
    G = np.concatenate((np.eye(32), P.astype(np.float64)), axis=0)
    return H, G


# --------------------------------------------------------------------------------------
# IEEE 802.11n, n = 1944, R = 1/2, Z = 81.  -1 = zero block, s >= 0 = identity cyclically
# shifted right by s (row i has its one in column (i + s) mod Z).
# --------------------------------------------------------------------------------------
_WIFI_1944_R12 = """
57 -1 -1 -1 50 -1 11 -1 50 -1 79 -1  1  0 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1
 3 -1 28 -1  0 -1 -1 -1 55  7 -1 -1 -1  0  0 -1 -1 -1 -1 -1 -1 -1 -1 -1
30 -1 -1 -1 24 37 -1 -1 56 14 -1 -1 -1 -1  0  0 -1 -1 -1 -1 -1 -1 -1 -1
62 53 -1 -1 53 -1 -1  3 35 -1 -1 -1 -1 -1 -1  0  0 -1 -1 -1 -1 -1 -1 -1
40 -1 -1 20 66 -1 -1 22 28 -1 -1 -1 -1 -1 -1 -1  0  0 -1 -1 -1 -1 -1 -1
 0 -1 -1 -1  8 -1 42 -1 50 -1 -1  8 -1 -1 -1 -1 -1  0  0 -1 -1 -1 -1 -1
69 79 79 -1 -1 -1 56 -1 52 -1 -1 -1  0 -1 -1 -1 -1 -1  0  0 -1 -1 -1 -1
65 -1 -1 -1 38 57 -1 -1 72 -1 27 -1 -1 -1 -1 -1 -1 -1 -1  0  0 -1 -1 -1
64 -1 -1 -1 14 52 -1 -1 30 -1 -1 32 -1 -1 -1 -1 -1 -1 -1 -1  0  0 -1 -1
-1 45 -1 70  0 -1 -1 -1 77  9 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1  0  0 -1
 2 56 -1 57 35 -1 -1 -1 -1 -1 12 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1  0  0
24 -1 61 -1 60 -1 -1 27 51 -1 -1 16  1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1  0
"""


# --------------------------------------------------------------------------------------
# The other eleven IEEE 802.11n prototypes (n = 648 / 1296 / 1944 with Z = 27 / 54 / 81; R = 1/2, 2/3, 3/4, 5/6).
# The reference ships none of them (bp/parity.py:7-47 holds only the (64,32) code; bp/masking.py:151-153 hints at a
# parity.mat that is not in its tree).  PROVENANCE: transcribed from memory of the standard's annex - there is no
# network in the build container, so the shift values cannot be checked against the standard here.  What IS checked
# (tests/test_codes.py::test_wifi_family_structure): shape, shift range, the dual-diagonal parity part with the
# (1, 0, 1) column the linear-time encoder needs, full rank, no 4-cycles, and encoder output in the null space of
# H.  A wrong shift value would change the BER of that code, never the decoder's correctness or its throughput;
# authoritative tables can be loaded with load_alist / load_mat / a qc_proto argument and run on the run-time QC kernel.
# --------------------------------------------------------------------------------------
_WIFI_PROTOS = {
    (648, "1/2"): (27, """
 0 -1 -1 -1  0  0 -1 -1  0 -1 -1  0  1  0 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1
22  0 -1 -1 17 -1  0  0 12 -1 -1 -1 -1  0  0 -1 -1 -1 -1 -1 -1 -1 -1 -1
 6 -1  0 -1 10 -1 -1 -1 24 -1  0 -1 -1 -1  0  0 -1 -1 -1 -1 -1 -1 -1 -1
 2 -1 -1  0 20 -1 -1 -1 25  0 -1 -1 -1 -1 -1  0  0 -1 -1 -1 -1 -1 -1 -1
23 -1 -1 -1  3 -1 -1 -1  0 -1  9 11 -1 -1 -1 -1  0  0 -1 -1 -1 -1 -1 -1
24 -1 23  1 17 -1  3 -1 10 -1 -1 -1 -1 -1 -1 -1 -1  0  0 -1 -1 -1 -1 -1
25 -1 -1 -1  8 -1 -1 -1  7 18 -1 -1  0 -1 -1 -1 -1 -1  0  0 -1 -1 -1 -1
13 24 -1 -1  0 -1  8 -1  6 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1  0  0 -1 -1 -1
 7 20 -1 16 22 10 -1 -1 23 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1  0  0 -1 -1
11 -1 -1 -1 19 -1 -1 -1 13 -1  3 17 -1 -1 -1 -1 -1 -1 -1 -1 -1  0  0 -1
25 -1  8 -1 23 18 -1 14  9 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1  0  0
 3 -1 -1 -1 16 -1 -1  2 25  5 -1 -1  1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1  0
"""),
    (648, "2/3"): (27, """
25 26 14 -1 20 -1  2 -1  4 -1 -1  8 -1 16 -1 18  1  0 -1 -1 -1 -1 -1 -1
10  9 15 11 -1  0 -1  1 -1 -1 18 -1  8 -1 10 -1 -1  0  0 -1 -1 -1 -1 -1
16  2 20 26 21 -1  6 -1  1 26 -1  7 -1 -1 -1 -1 -1 -1  0  0 -1 -1 -1 -1
10 13  5  0 -1  3 -1  7 -1 -1 26 -1 -1 13 -1 16 -1 -1 -1  0  0 -1 -1 -1
23 14 24 -1 12 -1 19 -1 17 -1 -1 -1 20 -1 21 -1  0 -1 -1 -1  0  0 -1 -1
 6 22  9 20 -1 25 -1 17 -1  8 -1 14 -1 18 -1 -1 -1 -1 -1 -1 -1  0  0 -1
14 23 21 11 20 -1 24 -1 18 -1 19 -1 -1 -1 -1 22 -1 -1 -1 -1 -1 -1  0  0
17 11 11 20 -1 21 -1 26 -1  3 -1 -1 18 -1 26 -1  1 -1 -1 -1 -1 -1 -1  0
"""),
    (648, "3/4"): (27, """
16 17 22 24  9  3 14 -1  4  2  7 -1 26 -1  2 -1 21 -1  1  0 -1 -1 -1 -1
25 12 12  3  3 26  6 21 -1 15 22 -1 15 -1  4 -1 -1 16 -1  0  0 -1 -1 -1
25 18 26 16 22 23  9 -1  0 -1  4 -1  4 -1  8 23 11 -1 -1 -1  0  0 -1 -1
 9  7  0  1 17 -1 -1  7  3 -1  3 23 -1 16 -1 -1 21 -1  0 -1 -1  0  0 -1
24  5 26  7  1 -1 -1 15 24 15 -1  8 -1 13 -1 13 -1 11 -1 -1 -1 -1  0  0
 2  2 19 14 24  1 15 19 -1 21 -1  2 -1 24 -1  3 -1  2  1 -1 -1 -1 -1  0
"""),
    (648, "5/6"): (27, """
17 13  8 21  9  3 18 12 10  0  4 15 19  2  5 10 26 19 13 13  1  0 -1 -1
 3 12 11 14 11 25  5 18  0  9  2 26 26 10 24  7 14 20  4  2 -1  0  0 -1
22 16  4  3 10 21 12  5 21 14 19  5 -1  8  5 18 11  5  5 15  0 -1  0  0
 7  7 14 14  4 16 16 24 24 10  1  7 15  6 10 26  8 18 21 14  1 -1 -1  0
"""),
    (1296, "1/2"): (54, """
40 -1 -1 -1 22 -1 49 23 43 -1 -1 -1  1  0 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1
50  1 -1 -1 48 35 -1 -1 13 -1 30 -1 -1  0  0 -1 -1 -1 -1 -1 -1 -1 -1 -1
39 50 -1 -1  4 -1  2 -1 -1 -1 -1 49 -1 -1  0  0 -1 -1 -1 -1 -1 -1 -1 -1
33 -1 -1 38 37 -1 -1  4  1 -1 -1 -1 -1 -1 -1  0  0 -1 -1 -1 -1 -1 -1 -1
45 -1 -1 -1  0 22 -1 -1 20 42 -1 -1 -1 -1 -1 -1  0  0 -1 -1 -1 -1 -1 -1
51 -1 -1 48 35 -1 -1 -1 44 -1 18 -1 -1 -1 -1 -1 -1  0  0 -1 -1 -1 -1 -1
47 11 -1 -1 -1 17 -1 -1 51 -1 -1 -1  0 -1 -1 -1 -1 -1  0  0 -1 -1 -1 -1
 5 -1 25 -1  6 -1 45 -1 13 40 -1 -1 -1 -1 -1 -1 -1 -1 -1  0  0 -1 -1 -1
33 -1 -1 34 24 -1 -1 -1 23 -1 -1 46 -1 -1 -1 -1 -1 -1 -1 -1  0  0 -1 -1
 1 -1 27 -1  1 -1 -1 -1 38 -1 44 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1  0  0 -1
-1 18 -1 -1 23 -1 -1  8  0 35 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1  0  0
49 -1 17 -1 30 -1 -1 -1 34 -1 -1 19  1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1  0
"""),
    (1296, "2/3"): (54, """
39 31 22 43 -1 40  4 -1 11 -1 -1 50 -1 -1 -1  6  1  0 -1 -1 -1 -1 -1 -1
25 52 41  2  6 -1 14 -1 34 -1 -1 -1 24 -1 37 -1 -1  0  0 -1 -1 -1 -1 -1
43 31 29  0 21 -1 28 -1 -1  2 -1 -1  7 -1 17 -1 -1 -1  0  0 -1 -1 -1 -1
20 33 48 -1  4 13 -1 26 -1 -1 22 -1 -1 46 42 -1 -1 -1 -1  0  0 -1 -1 -1
45  7 18 51 12 25 -1 -1 -1 50 -1 -1  5 -1 -1 -1  0 -1 -1 -1  0  0 -1 -1
35 40 32 16  5 -1 -1 18 -1 -1 43 51 -1 32 -1 -1 -1 -1 -1 -1 -1  0  0 -1
 9 24 13 22 28 -1 -1 37 -1 -1 25 -1 -1 52 -1 13 -1 -1 -1 -1 -1 -1  0  0
32 22  4 21 16 -1 -1 -1 27 28 -1 38 -1 -1 -1  8  1 -1 -1 -1 -1 -1 -1  0
"""),
    (1296, "3/4"): (54, """
39 40 51 41  3 29  8 36 -1 14 -1  6 -1 33 -1 11 -1  4  1  0 -1 -1 -1 -1
48 21 47  9 48 35 51 -1 38 -1 28 -1 34 -1 50 -1 50 -1 -1  0  0 -1 -1 -1
30 39 28 42 50 39  5 17 -1  6 -1 18 -1 20 -1 15 -1 40 -1 -1  0  0 -1 -1
29  0  1 43 36 30 47 -1 49 -1 47 -1  3 -1 35 -1 34 -1  0 -1 -1  0  0 -1
 1 32 11 23 10 44 12  7 -1 48 -1  4 -1  9 -1 17 -1 16 -1 -1 -1 -1  0  0
13  7 15 47 23 16 47 -1 43 -1 29 -1 52 -1  2 -1 53 -1  1 -1 -1 -1 -1  0
"""),
    (1296, "5/6"): (54, """
48 29 37 52  2 16  6 14 53 31 34  5 18 42 53 31 45 -1 46 52  1  0 -1 -1
17  4 30  7 43 11 24  6 14 21  6 39 17 40 47  7 15 41 19 -1 -1  0  0 -1
 7  2 51 31 46 23 16 11 53 40 10  7 46 53 33 35 -1 25 35 38  0 -1  0  0
19 48 41  1 10  7 36 47  5 29 52 52 31 10 26  6  3  2 -1 51  1 -1 -1  0
"""),
    (1944, "2/3"): (81, """
61 75  4 63 56 -1 -1 -1 -1 -1 -1  8 -1  2 17 25  1  0 -1 -1 -1 -1 -1 -1
56 74 77 20 -1 -1 -1 64 24  4 67 -1  7 -1 -1 -1 -1  0  0 -1 -1 -1 -1 -1
28 21 68 10  7 14 65 -1 -1 -1 23 -1 -1 -1 75 -1 -1 -1  0  0 -1 -1 -1 -1
48 38 43 78 76 -1 -1 -1 -1  5 36 -1 15 72 -1 -1 -1 -1 -1  0  0 -1 -1 -1
40  2 53 25 -1 52 62 -1 20 -1 -1 44 -1 -1 -1 -1  0 -1 -1 -1  0  0 -1 -1
69 23 64 10 22 -1 21 -1 -1 -1 -1 -1 68 23 29 -1 -1 -1 -1 -1 -1  0  0 -1
12  0 68 20 55 61 -1 40 -1 -1 -1 52 -1 -1 -1 44 -1 -1 -1 -1 -1 -1  0  0
58  8 34 64 78 -1 -1 11 78 24 -1 -1 -1 -1 -1 58  1 -1 -1 -1 -1 -1 -1  0
"""),
    (1944, "3/4"): (81, """
48 29 28 39  9 61 -1 -1 -1 63 45 80 -1 -1 -1 37 32 22  1  0 -1 -1 -1 -1
 4 49 42 48 11 30 -1 -1 -1 49 17 41 37 15 -1 54 -1 -1 -1  0  0 -1 -1 -1
35 76 78 51 37 35 21 -1 17 64 -1 -1 -1 59  7 -1 -1 32 -1 -1  0  0 -1 -1
 9 65 44  9 54 56 73 34 42 -1 -1 -1 35 -1 -1 -1 46 39  0 -1 -1  0  0 -1
 3 62  7 80 68 26 -1 80 55 -1 36 -1 26 -1  9 -1 72 -1 -1 -1 -1 -1  0  0
26 75 33 21 69 59  3 38 -1 -1 -1 35 -1 62 36 26 -1 -1  1 -1 -1 -1 -1  0
"""),
    (1944, "5/6"): (81, """
13 48 80 66  4 74  7 30 76 52 37 60 -1 49 73 31 74 73 23 -1  1  0 -1 -1
69 63 74 56 64 77 57 65  6 16 51 -1 64 -1 68  9 48 62 54 27 -1  0  0 -1
51 15  0 80 24 25 42 54 44 71 71  9 67 35 -1 58 -1 29 -1 53  0 -1  0  0
16 29 36 41 44 56 59 37 50 24 -1 65  4 65 52 -1  4 -1 73 52  1 -1 -1  0
"""),
}
_WIFI_PROTOS[(1944, "1/2")] = (81, _WIFI_1944_R12)
WIFI_LENGTHS = (648, 1296, 1944)
WIFI_RATES = ("1/2", "2/3", "3/4", "5/6")


def ieee80211n(n: int = 1944, rate: str = "1/2") -> "QCCode":
    """One of the twelve IEEE 802.11n prototypes: n in {648, 1296, 1944}, rate in {'1/2', '2/3', '3/4', '5/6'}."""
    try:
        Z, text = _WIFI_PROTOS[(int(n), str(rate))]
    except KeyError:
        raise ValueError(f"no 802.11n prototype for n={n}, rate={rate}: n in {WIFI_LENGTHS}, rate in {WIFI_RATES}") from None
    return QCCode(f"802.11n-{n}-r{rate}", _parse_proto(text), Z)


def ieee80211n_family():
    return [ieee80211n(n, r) for n in WIFI_LENGTHS for r in WIFI_RATES]


def _parse_proto(text):
    rows = [[int(t) for t in line.split()] for line in text.strip().splitlines()]
    return np.array(rows, dtype=np.int16)


def expand_qc(proto: np.ndarray, Z: int) -> np.ndarray:
    """Expand a prototype matrix into the binary H (uint8 [mb*Z, nb*Z])."""
    mb, nb = proto.shape
    H = np.zeros((mb * Z, nb * Z), dtype=np.uint8)
    idx = np.arange(Z)
    for r in range(mb):
        for c in range(nb):
            s = int(proto[r, c])
            if s >= 0:
                H[r * Z + idx, c * Z + (idx + s) % Z] = 1
    return H


@dataclasses.dataclass
class QCCode:
    """A quasi-cyclic code given by its prototype (shift) matrix."""
    name: str
    proto: np.ndarray          # int16 [mb, nb], -1 = zero block
    Z: int

    @property
    def mb(self): return int(self.proto.shape[0])

    @property
    def nb(self): return int(self.proto.shape[1])

    @property
    def n(self): return self.nb * self.Z

    @property
    def m(self): return self.mb * self.Z

    @property
    def k(self): return self.n - self.m

    @property
    def H(self) -> np.ndarray:
        if not hasattr(self, "_H"):
            self._H = expand_qc(self.proto, self.Z)
        return self._H

    # -- linear-time systematic encoder for the 802.11n dual-diagonal parity part -------
    def encode(self, info_bits: np.ndarray) -> np.ndarray:
        """info_bits uint8 [B, k] -> codewords uint8 [B, n] (information bits first).

        Parity part is [h | T] with T dual-diagonal and column h having three entries
        (top shift s0, one middle shift 0, bottom shift s0).  Summing all block rows
        cancels T and the two equal h-shifts, leaving p0 = sum_r lambda_r.
        """
        u = np.asarray(info_bits, dtype=np.uint8)
        B = u.shape[0]
        Z, mb, nb = self.Z, self.mb, self.nb
        kb = nb - mb
        ub = u.reshape(B, kb, Z)
        lam = np.zeros((B, mb, Z), dtype=np.uint8)
        for r in range(mb):
            for c in range(kb):
                s = int(self.proto[r, c])
                if s >= 0:
                    # row i of the block reads u[(i+s) % Z]
                    lam[:, r, :] ^= np.roll(ub[:, c, :], -s, axis=1)
        hcol = self.proto[:, kb]
        hrows = [r for r in range(mb) if hcol[r] >= 0]
        assert len(hrows) == 3 and hrows[0] == 0 and hrows[-1] == mb - 1
        assert hcol[hrows[0]] == hcol[hrows[-1]] and hcol[hrows[1]] == 0
        p = np.zeros((B, mb, Z), dtype=np.uint8)
        p0 = np.bitwise_xor.reduce(lam, axis=1)           # p0 (shift 0 of the middle entry)
        p[:, 0, :] = p0
        # forward substitution: row r:  lam_r + h_r(p0) + p_r' + p_{r+1}' = 0, with p_0' = 0
        # where p_j' (j>=1) is parity block j.  Row 0: p_1 = lam_0 + shift(p0, s0)
        for r in range(mb - 1):
            acc = lam[:, r, :].copy()
            if hcol[r] >= 0:
                acc ^= np.roll(p0, -int(hcol[r]), axis=1)
            if r >= 1:
                acc ^= p[:, r, :]
            p[:, r + 1, :] = acc
        return np.concatenate([u, p.reshape(B, mb * Z)], axis=1)


def ieee80211n_1944_r12() -> QCCode:
    return QCCode("802.11n-1944-r1/2", _parse_proto(_WIFI_1944_R12), 81)


# --------------------------------------------------------------------------------------
# Edge tables
# --------------------------------------------------------------------------------------
@dataclasses.dataclass
class EdgeTables:
    """Sparse replacement for generate_masks(H) (reference masking.py:12-147).

    cm = check-major edge id (row-major non-zeros), vm = variable-major edge id
    (column-major non-zeros).  All arrays int32.
    """
    m: int
    n: int
    E: int
    chk_ptr: np.ndarray     # [m+1]  cm ids of check c are chk_ptr[c]..chk_ptr[c+1]
    chk_var: np.ndarray     # [E]    variable of cm edge
    var_ptr: np.ndarray     # [n+1]  vm ids of variable v are var_ptr[v]..var_ptr[v+1]
    var_chk: np.ndarray     # [E]    check of vm edge
    cm_of_vm: np.ndarray    # [E]    cm id of a vm edge
    vm_of_cm: np.ndarray    # [E]    vm id of a cm edge

    @property
    def max_dc(self): return int(np.diff(self.chk_ptr).max())

    @property
    def max_dv(self): return int(np.diff(self.var_ptr).max())

    @staticmethod
    def from_H(H) -> "EdgeTables":
        H = np.asarray(H)
        if H.ndim != 2:
            raise ValueError("H must be a 2-D 0/1 matrix")
        Hb = (H != 0)
        m, n = Hb.shape
        rows, cols = np.nonzero(Hb)                     # row-major = check-major order
        E = int(rows.size)
        chk_ptr = np.zeros(m + 1, dtype=np.int64)
        np.add.at(chk_ptr, rows + 1, 1)
        chk_ptr = np.cumsum(chk_ptr)
        # variable-major: sort cm edges by (col, row); stable sort keeps rows ascending
        order = np.lexsort((rows, cols))                # vm -> cm
        var_ptr = np.zeros(n + 1, dtype=np.int64)
        np.add.at(var_ptr, cols + 1, 1)
        var_ptr = np.cumsum(var_ptr)
        vm_of_cm = np.empty(E, dtype=np.int64)
        vm_of_cm[order] = np.arange(E)
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        return EdgeTables(m=m, n=n, E=E, chk_ptr=i32(chk_ptr), chk_var=i32(cols),
                          var_ptr=i32(var_ptr), var_chk=i32(rows[order]),
                          cm_of_vm=i32(order), vm_of_cm=i32(vm_of_cm))

    def dense_masks(self):
        """The four dense arrays generate_masks(H) returns (masking.py:147); only for
        API compatibility and tests - never used by the decoder."""
        E, n = self.E, self.n
        mask_c = np.zeros((E, E)); mask_v = np.zeros((E, E))
        mask_v_final = np.zeros((n, E)); llr_expander = np.zeros((E, n))
        llr_expander[np.arange(E), np.repeat(np.arange(n), np.diff(self.var_ptr))] = 1
        mask_v_final[self.chk_var, np.arange(E)] = 1
        for v in range(n):
            vm = np.arange(self.var_ptr[v], self.var_ptr[v + 1])
            cm = self.cm_of_vm[vm]
            for a in range(vm.size):
                for b in range(vm.size):
                    if a != b:
                        mask_v[vm[a], cm[b]] = 1       # out vm edge a <- in cm edge b
        for c in range(self.m):
            cm = np.arange(self.chk_ptr[c], self.chk_ptr[c + 1])
            vm = self.vm_of_cm[cm]
            for a in range(cm.size):
                for b in range(cm.size):
                    if a != b:
                        mask_c[cm[a], vm[b]] = 1       # out cm edge a <- in vm edge b
        return mask_c, mask_v, mask_v_final, llr_expander


# --------------------------------------------------------------------------------------
# The reference's trainable weights (bp/bp_vc.py:101-107) <-> sparse tables
# --------------------------------------------------------------------------------------
def _weight_slots(T: "EdgeTables"):
    """vm_var [E] (variable of a variable-major edge), pos [E] (its position inside the variable), used [E,max_dv]
    (entry (e, j) is a real weight: the variable has a j-th edge and it is not e itself)."""
    dv = np.diff(T.var_ptr)
    vm_var = np.repeat(np.arange(T.n), dv)
    pos = np.arange(T.E) - T.var_ptr[vm_var]
    used = np.zeros((T.E, T.max_dv), bool)
    for j in range(T.max_dv):
        used[:, j] = (dv[vm_var] > j) & (pos != j)
    return vm_var, pos, used


def sparse_weights_from_reference_state(T: "EdgeTables", state, iterations):
    """Dense parameters of a reference BeliefPropagation state_dict (bp/bp.py:26-39: layers.{i}.0.input_weight [E,E],
    layers.{i}.0.llr_weight [1,n], final_layer.0.input_weight [n,E], final_layer.0.llr_weight [1,n]) -> numpy tables
    w_edge [iters,E,max_dv] (row = variable-major OUT edge, column j = weight of the variable's j-th edge as INPUT;
    unused entries 1), w_llr [iters,n], wf_edge [E] (variable-major), wf_llr [n]."""
    A = lambda k: np.asarray(state[k].detach().cpu().numpy() if hasattr(state[k], "detach") else state[k], dtype=np.float32)
    vm_var, pos, used = _weight_slots(T)
    E, n, mdv = T.E, T.n, T.max_dv
    w_edge = np.ones((iterations, E, mdv), np.float32)
    w_llr = np.ones((iterations, n), np.float32)
    for i in range(iterations):
        W = A(f"layers.{i}.0.input_weight")
        w_llr[i] = A(f"layers.{i}.0.llr_weight").reshape(-1)
        for j in range(mdv):
            sel = np.nonzero(used[:, j])[0]
            w_edge[i, sel, j] = W[sel, T.cm_of_vm[T.var_ptr[vm_var[sel]] + j]]
    wf_edge = np.ascontiguousarray(A("final_layer.0.input_weight")[vm_var, T.cm_of_vm])
    return dict(w_edge=w_edge, w_llr=w_llr, wf_edge=wf_edge, wf_llr=A("final_layer.0.llr_weight").reshape(-1).copy())


def reference_state_from_sparse_weights(T: "EdgeTables", w):
    """Inverse of sparse_weights_from_reference_state: the dense trainable tensors of the reference's state_dict
    (masks are not included; load with strict=False there).  Dense [E,E] per layer: small codes only."""
    if T.E > 4096:
        raise ValueError(f"E = {T.E}: the reference's dense [E,E] layout is impractical for this code")
    vm_var, pos, used = _weight_slots(T)
    w_edge, w_llr = np.asarray(w["w_edge"], np.float32), np.asarray(w["w_llr"], np.float32)
    out = {}
    for i in range(w_edge.shape[0]):
        W = np.zeros((T.E, T.E), np.float32)
        for j in range(T.max_dv):
            sel = np.nonzero(used[:, j])[0]
            W[sel, T.cm_of_vm[T.var_ptr[vm_var[sel]] + j]] = w_edge[i, sel, j]
        out[f"layers.{i}.0.input_weight"] = W
        out[f"layers.{i}.0.llr_weight"] = w_llr[i].reshape(1, -1).copy()
    Wf = np.zeros((T.n, T.E), np.float32)
    Wf[vm_var, T.cm_of_vm] = np.asarray(w["wf_edge"], np.float32)
    out["final_layer.0.input_weight"] = Wf
    out["final_layer.0.llr_weight"] = np.asarray(w["wf_llr"], np.float32).reshape(1, -1).copy()
    return out


# --------------------------------------------------------------------------------------
# QC detection and dense GF(2) helpers
# --------------------------------------------------------------------------------------
def detect_qc(H: np.ndarray, Z: int):
    """Return the prototype matrix if H is block-circulant with circulant weight <= 1
    for block size Z, else None."""
    H = np.asarray(H)
    m, n = H.shape
    if Z <= 0 or m % Z or n % Z:
        return None
    mb, nb = m // Z, n // Z
    proto = np.full((mb, nb), -1, dtype=np.int16)
    idx = np.arange(Z)
    for r in range(mb):
        for c in range(nb):
            blk = H[r * Z:(r + 1) * Z, c * Z:(c + 1) * Z]
            w = int((blk != 0).sum())
            if w == 0:
                continue
            if w != Z:
                return None
            s = int(np.nonzero(blk[0])[0][0])
            if not np.all(blk[idx, (idx + s) % Z] != 0):
                return None
            proto[r, c] = s
    return proto


def gf2_rank(M: np.ndarray) -> int:
    A = (np.asarray(M) != 0).astype(np.uint8).copy()
    rows, cols = A.shape
    r = 0
    for c in range(cols):
        piv = np.nonzero(A[r:, c])[0]
        if piv.size == 0:
            continue
        p = r + piv[0]
        if p != r:
            A[[r, p]] = A[[p, r]]
        sel = np.nonzero(A[:, c])[0]
        sel = sel[sel != r]
        A[sel] ^= A[r]
        r += 1
        if r == rows:
            break
    return r


def systematic_generator(H: np.ndarray) -> np.ndarray:
    """Dense G (uint8 [n,k], codeword = G u mod 2, information bits first) for any H
    whose last m columns are invertible over GF(2):  H = [A | B]  ->  G = [I ; B^-1 A]."""
    Hb = (np.asarray(H) != 0).astype(np.uint8)
    m, n = Hb.shape
    k = n - m
    aug = np.concatenate([Hb[:, k:], Hb[:, :k]], axis=1)   # [B | A]
    for c in range(m):
        piv = np.nonzero(aug[c:, c])[0]
        if piv.size == 0:
            raise ValueError("parity part of H is singular")
        p = c + piv[0]
        if p != c:
            aug[[c, p]] = aug[[p, c]]
        sel = np.nonzero(aug[:, c])[0]
        sel = sel[sel != c]
        aug[sel] ^= aug[c]
    return np.concatenate([np.eye(k, dtype=np.uint8), aug[:, m:]], axis=0)


# --------------------------------------------------------------------------------------
# parity-check matrix import / export (bp/masking.py:151-153 hints at .mat files)
# --------------------------------------------------------------------------------------
def load_alist(path_or_text) -> np.ndarray:
    """MacKay alist -> dense H (uint8 [m,n]).  Accepts a path or the text itself; zero entries pad short rows."""
    text = path_or_text
    if "\n" not in str(path_or_text):
        with open(path_or_text) as f:
            text = f.read()
    tok = [int(t) for t in str(text).split()]
    n, m = tok[0], tok[1]
    pos = 4 + n + m                                    # n m | max_dv max_dc | dv[n] | dc[m]
    max_dv, max_dc = tok[2], tok[3]
    dv = tok[4:4 + n]
    padded = len(tok) >= 4 + n + m + n * max_dv + m * max_dc          # rows padded with zeros to the maximum degree
    H = np.zeros((m, n), dtype=np.uint8)
    for v in range(n):
        for k in range(max_dv if padded else dv[v]):
            c = tok[pos]; pos += 1
            if c > 0:
                H[c - 1, v] = 1
    if any(int(H[:, v].sum()) != dv[v] for v in range(n)):
        raise ValueError("alist column degrees do not match the listed entries")
    return H


def save_alist(H: np.ndarray) -> str:
    Hb = (np.asarray(H) != 0)
    m, n = Hb.shape
    dv, dc = Hb.sum(0), Hb.sum(1)
    out = [f"{n} {m}", f"{int(dv.max())} {int(dc.max())}", " ".join(str(int(d)) for d in dv), " ".join(str(int(d)) for d in dc)]
    for v in range(n):
        rows = (np.nonzero(Hb[:, v])[0] + 1).tolist()
        out.append(" ".join(str(r) for r in rows + [0] * (int(dv.max()) - len(rows))))
    for c in range(m):
        cols = (np.nonzero(Hb[c])[0] + 1).tolist()
        out.append(" ".join(str(r) for r in cols + [0] * (int(dc.max()) - len(cols))))
    return "\n".join(out) + "\n"


def load_mat(path, key=None) -> np.ndarray:
    """H from a MATLAB .mat file (scipy.io.loadmat; first 2-D array or `key`)."""
    import scipy.io
    d = scipy.io.loadmat(path)
    if key is None:
        key = next(k for k, v in d.items() if not k.startswith("__") and getattr(v, "ndim", 0) == 2)
    M = d[key]
    M = M.toarray() if hasattr(M, "toarray") else np.asarray(M)
    return (M != 0).astype(np.uint8)


def auto_qc_block_size(H: np.ndarray, min_Z=24):
    """Largest block size Z >= min_Z (a common divisor of both dimensions) for which H is block-circulant, else 0.  Used
    when a caller hands over a bare H, as every reference call site does (BeliefPropagation(H, iterations),
    decode_bits(llrs, H, ...)): a quasi-cyclic H then gets the QC kernels without being told."""
    H = np.asarray(H)
    m, n = H.shape
    g = int(np.gcd(m, n))
    for Z in sorted((z for z in range(min_Z, g + 1) if g % z == 0), reverse=True):
        if (m // Z) * (n // Z) > 4096:                       # not a prototype-sized block grid
            break
        if detect_qc(H, Z) is not None:
            return Z
    return 0


def qc_block_size(H: np.ndarray, candidates=(81, 54, 27, 96, 64, 48, 32, 24, 16, 8)):
    """Largest candidate Z for which H is block-circulant (so LdpcCode(H, qc_Z=Z) can pick a compiled kernel), else 0."""
    for Z in candidates:
        if detect_qc(H, Z) is not None:
            return Z
    return 0
