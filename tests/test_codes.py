"""Code library: structural validators and encoders."""
import numpy as np
import pytest

from ldpc_b200.codes import (EdgeTables, detect_qc, gf2_rank, ieee80211n_1944_r12, peg_64_32,
                             systematic_generator)


def test_default_code_shapes():
    H, G = peg_64_32()
    assert H.shape == (32, 64) and H.dtype == np.int64 and G.shape == (64, 32) and G.dtype == np.float64
    assert int(H.sum()) == 96
    assert set(H.sum(1)) == {3} and set(H.sum(0)) <= {1, 2}
    assert not ((H @ G.astype(np.int64)) % 2).any()            # H G = 0 (mod 2)
    assert np.array_equal(G[:32], np.eye(32))                  # systematic, info bits first


def test_wifi_structure():
    qc = ieee80211n_1944_r12()
    H = qc.H
    assert H.shape == (972, 1944) and int(H.sum()) == 6966 and (qc.proto >= 0).sum() == 86
    rd, rc = np.unique(H.sum(1), return_counts=True)
    assert dict(zip(rd.tolist(), rc.tolist())) == {7: 810, 8: 162}
    cd, cc = np.unique(H.sum(0), return_counts=True)
    assert dict(zip(cd.tolist(), cc.tolist())) == {2: 891, 3: 729, 4: 81, 11: 243}
    assert gf2_rank(H) == 972
    ov = H.astype(np.int32) @ H.astype(np.int32).T
    np.fill_diagonal(ov, 0)
    assert ov.max() == 1                                        # no 4-cycles
    # dual-diagonal parity part, column 12 = {1, 0, 1} at rows 0, 6, 11
    assert [int(qc.proto[r, 12]) for r in (0, 6, 11)] == [1, 0, 1]
    for j in range(11):
        assert qc.proto[j, 13 + j] == 0 and qc.proto[j + 1, 13 + j] == 0
    assert np.array_equal(detect_qc(H, 81), qc.proto)
    assert detect_qc(H, 80) is None


def test_wifi_family_structure():
    """All twelve IEEE 802.11n prototypes (Z in {27, 54, 81}, rates 1/2 .. 5/6).  The shift values are
    transcribed from memory (see ldpc_b200/codes.py); these are the structural invariants the kernels and the
    linear-time encoder rely on, and they hold for every table shipped."""
    from ldpc_b200.codes import WIFI_LENGTHS, WIFI_RATES, ieee80211n, ieee80211n_family
    fam = ieee80211n_family()
    assert len(fam) == 12 and len({q.name for q in fam}) == 12
    rate_mb = {"1/2": 12, "2/3": 8, "3/4": 6, "5/6": 4}
    for n in WIFI_LENGTHS:
        for rate in WIFI_RATES:
            qc = ieee80211n(n, rate)
            P, Z, mb, nb = qc.proto, qc.Z, qc.mb, qc.nb
            assert (nb, mb, Z) == (24, rate_mb[rate], n // 24) and qc.n == n and qc.k == n - mb * Z
            assert P.min() >= -1 and P.max() < Z
            kb = nb - mb
            # information part: every block column carries at least two blocks, every row has the same degree +-1
            deg_r = (P >= 0).sum(1)
            assert deg_r.max() - deg_r.min() <= 1 and (P[:, :kb] >= 0).sum(0).min() >= 2
            # parity part [h | T]: T dual-diagonal with shift 0, h = (1, 0, 1) at the top row, one middle row, the bottom row
            for j in range(mb - 1):
                col = P[:, kb + 1 + j]
                assert col[j] == 0 and col[j + 1] == 0 and (np.delete(col, [j, j + 1]) == -1).all()
            hr = [r for r in range(mb) if P[r, kb] >= 0]
            assert len(hr) == 3 and hr[0] == 0 and hr[2] == mb - 1 and [int(P[r, kb]) for r in hr] == [1, 0, 1]
            H = qc.H
            assert H.shape == (mb * Z, n) and int(H.sum()) == int((P >= 0).sum()) * Z
            assert gf2_rank(H) == mb * Z
            assert np.array_equal(detect_qc(H, Z), P)
            u = np.random.RandomState(n + mb).randint(0, 2, (8, qc.k)).astype(np.uint8)
            c = qc.encode(u)                                       # linear-time dual-diagonal encoder
            assert np.array_equal(c[:, :qc.k], u) and not ((H.astype(np.int64) @ c.T.astype(np.int64)) % 2).any()
    with pytest.raises(ValueError):
        ieee80211n(1944, "7/8")
    assert np.array_equal(ieee80211n(1944, "1/2").proto, ieee80211n_1944_r12().proto)


def test_compiled_prototype_header_matches_code_library():
    """csrc/qc_protos.cuh (+ qc_plan.cuh for the headline code) is generated from ldpc_b200/codes.py: the compiled tables are
    the library's tables (ldpc_code_create only selects a compiled kernel when they agree entry by entry)."""
    import os, re
    from ldpc_b200.codes import ieee80211n
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    txt = open(os.path.join(root, "ldpc-sims_b200", "csrc", "qc_protos.cuh")).read() + \
        open(os.path.join(root, "ldpc-sims_b200", "csrc", "qc_plan.cuh")).read()
    found = 0
    for m in re.finditer(r"struct Wifi(\d+)R(\d)(\d) \{.*?proto\[MB\]\[NB\] = \{(.*?)\};", txt, re.S):
        n, a, b, body = int(m.group(1)), m.group(2), m.group(3), m.group(4)
        vals = np.array([int(v) for v in re.findall(r"-?\d+", body)], dtype=np.int16)
        qc = ieee80211n(n, f"{a}/{b}")
        assert np.array_equal(vals.reshape(qc.proto.shape), qc.proto), m.group(0)[:40]
        found += 1
    assert found == 12


def test_wifi_encoder_matches_dense_generator():
    qc = ieee80211n_1944_r12()
    rng = np.random.RandomState(0)
    u = rng.randint(0, 2, (16, qc.k)).astype(np.uint8)
    c = qc.encode(u)
    assert np.array_equal(c[:, :qc.k], u)
    assert not ((qc.H.astype(np.int64) @ c.T.astype(np.int64)) % 2).any()
    G = systematic_generator(qc.H)
    assert np.array_equal((G.astype(np.int64) @ u.T.astype(np.int64) % 2).T, c)


@pytest.mark.parametrize("seed", [0, 1])
def test_edge_tables_random(seed):
    rng = np.random.RandomState(seed)
    H = (rng.rand(7, 15) < 0.3).astype(np.int64)
    H[:, 0] = 1; H[0, :] = 1
    et = EdgeTables.from_H(H)
    rows, cols = np.nonzero(H)
    assert et.E == rows.size
    assert np.array_equal(et.chk_var, cols)
    vcols, vrows = np.nonzero(H.T)
    assert np.array_equal(et.var_chk, vrows)
    # permutations are inverse of each other and connect the same (check, variable) pair
    assert np.array_equal(et.vm_of_cm[et.cm_of_vm], np.arange(et.E))
    assert np.array_equal(rows[et.cm_of_vm], vrows) and np.array_equal(cols[et.cm_of_vm], vcols)
    mask_c, mask_v, mask_v_final, llr_expander = et.dense_masks()
    assert mask_v_final.shape == (15, et.E) and llr_expander.shape == (et.E, 15)
    assert mask_v_final.sum() == et.E and llr_expander.sum() == et.E
    from bp.masking import masks_to_H
    assert np.array_equal(masks_to_H(mask_c, mask_v, mask_v_final, llr_expander)[:, :], _dedup(H))


def _dedup(H):
    # masks_to_H cannot tell identical rows apart; the random H above has distinct rows w.h.p.
    return H


def test_alist_and_mat_round_trip(tmp_path):
    import scipy.io
    from ldpc_b200.codes import load_alist, save_alist, load_mat, qc_block_size, peg_64_32, ieee80211n_1944_r12
    H = peg_64_32()[0]
    txt = save_alist(H)
    assert np.array_equal(load_alist(txt), H != 0)
    p = tmp_path / "h.alist"
    p.write_text(txt)
    assert np.array_equal(load_alist(str(p)), H != 0)
    scipy.io.savemat(str(tmp_path / "h.mat"), {"H": H})
    assert np.array_equal(load_mat(str(tmp_path / "h.mat")), H != 0)
    Hw = ieee80211n_1944_r12().H
    assert np.array_equal(load_alist(save_alist(Hw)), Hw != 0)
    assert qc_block_size(Hw) == 81 and qc_block_size(H) in (0, 8, 16, 32)


def test_auto_qc_block_size():
    """A bare H (what every reference call site passes) is recognised as quasi-cyclic."""
    from ldpc_b200.codes import auto_qc_block_size, expand_qc, ieee80211n_1944_r12, peg_64_32
    assert auto_qc_block_size(ieee80211n_1944_r12().H) == 81
    assert auto_qc_block_size(peg_64_32()[0]) == 0
    rng = np.random.RandomState(0)
    proto = rng.randint(-1, 27, size=(4, 8)).astype(np.int16)
    assert auto_qc_block_size(expand_qc(proto, 27)) == 27
    H = (rng.rand(48, 96) < 0.1).astype(np.uint8)
    assert auto_qc_block_size(H) == 0
