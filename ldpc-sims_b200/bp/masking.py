"""generate_masks(H) kept for API compatibility (reference bp/masking.py:12-147).

The decoder never uses dense masks - it runs on the sparse edge tables in
ldpc_b200.codes.EdgeTables.  ``masks_to_H`` inverts the four masks for the stale 5-argument
BeliefPropagation constructor (ber_test.py:46, joint_connected.py:19)."""
import numpy as np

from ldpc_b200.codes import EdgeTables


def generate_masks(H):
    """-> (mask_c [E,E], mask_v [E,E], mask_v_final [n,E], llr_expander [E,n]) float64."""
    return EdgeTables.from_H(np.asarray(H)).dense_masks()


genMasks = generate_masks          # old spelling (ber_test.py:9, joint_test.py:14, ofdm_nn.py:268)


def masks_to_H(mask_c, mask_v, mask_v_final, llr_expander):
    """Rebuild H from the four dense masks.  Check-major edge e touches variable
    argmax(mask_v_final[:, e]); the other edges of its check are the variable-major edges b
    with mask_c[e, b] == 1, whose variables come from llr_expander."""
    mask_c = np.asarray(mask_c); mask_v_final = np.asarray(mask_v_final); llr_expander = np.asarray(llr_expander)
    n, E = mask_v_final.shape
    var_of_cm = np.argmax(mask_v_final, axis=0)
    var_of_vm = np.argmax(llr_expander, axis=1)
    rows, seen = [], {}
    for e in range(E):
        members = frozenset([int(var_of_cm[e])] + [int(var_of_vm[b]) for b in np.nonzero(mask_c[e])[0]])
        if members not in seen:
            seen[members] = len(rows)
            rows.append(sorted(members))
    H = np.zeros((len(rows), n), dtype=np.int64)
    for r, vs in enumerate(rows):
        H[r, vs] = 1
    if int(H.sum()) != E:
        raise ValueError("masks do not describe a parity-check matrix (duplicate checks are not recoverable)")
    return H
