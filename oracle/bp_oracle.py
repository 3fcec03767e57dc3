"""CPU ORACLE - TEST INFRASTRUCTURE ONLY (never imported by the product path).

Sparse restatement of the reference's dense-masked belief-propagation decoder:

    pytorch/bp/masking.py:75-138   edge numbering (check-major / variable-major)
    pytorch/bp/bp_vc.py:16-32      V->C   0.5 * (llr' + sum of the other C->V messages)
    pytorch/bp/bp.py:29            tanh
    pytorch/bp/bp_cv.py:22-55      C->V   log((1+p)/(1-p)), p = clamp(prod of others, +-(1-1e-7))
    pytorch/bp/bp.py:43-51         iteration loop, outer clamp, final marginal, 1 - sigmoid
    pytorch/ofdm/ofdm_functions.py:131-163   decode_bits batching wrapper / np.round

Arithmetic is fp32 throughout.  Transcendentals (tanh, log, sigmoid) go through the same
PyTorch CPU ATen kernels the reference itself calls, so on the reference's default code
(every sum / product has at most two non-trivial operands) this restatement is
bit-identical to the dense reference.  For higher degrees the dense reference's
association order is whatever sgemm / prod pick; this file FIXES it:

    sum of others   S_k = P_k + Q_k,  P_k = ((x_0 + x_1) + ...) + x_{k-1}   (forward)
                                      Q_k = x_{k+1} + (x_{k+2} + (... + x_{d-1}))  (backward)
                    an empty side contributes nothing; both empty -> +0.0f
    prod of others  same two-sweep association with *
    marginal        ((x_0 + x_1) + ...) + x_{d-1}, then  0.5 * (llr' + that)

Min-sum / normalized / offset min-sum do not exist in the reference; they are DEFINED
here: V->C y = llr' + S_k (no 1/2, no tanh); C->V magnitude =
min over the others of |y| (times alpha, or minus beta floored at 0), sign = xor of the
others' IEEE sign bits; same outer clamp; same marginal 0.5*(llr' + sum) so prob / hard /
posterior keep their meaning.  Only add / min / abs / sign ops => bit-exact CPU<->GPU.

Pinned against the dense reference by oracle/make_golden.py (fixtures in tests/golden/).
"""
from __future__ import annotations

import numpy as np
import torch

F32 = np.float32
P_CLAMP = np.float32(1.0 - 1e-7)      # rounds to 0.99999988 in fp32 (bp_cv.py:44-47)

UPDATE_SP, UPDATE_MINSUM, UPDATE_NMS, UPDATE_OMS = 0, 1, 2, 3
_UPDATE_IDS = {"sp": 0, "tanh": 0, "sum-product": 0, "minsum": 1, "min-sum": 1, "nms": 2, "oms": 3}


class Graph:
    """Edge enumeration exactly as masking.py does it."""

    def __init__(self, H):
        H = np.asarray(H) != 0
        self.m, self.n = H.shape
        rows, cols = np.nonzero(H)                      # check-major ids (masking.py:85-88)
        self.E = rows.size
        self.cm_row, self.cm_col = rows, cols
        vcols, vrows = np.nonzero(H.T)                  # variable-major ids (masking.py:92-95)
        self.vm_row, self.vm_col = vrows, vcols
        cm_id = -np.ones(H.shape, dtype=np.int64); cm_id[rows, cols] = np.arange(self.E)
        vm_id = -np.ones(H.shape, dtype=np.int64); vm_id[vrows, vcols] = np.arange(self.E)
        self.cm_of_vm = cm_id[vrows, vcols]
        self.vm_of_cm = vm_id[rows, cols]
        self.dv = H.sum(0).astype(np.int64)
        self.dc = H.sum(1).astype(np.int64)
        self.var_ptr = np.concatenate([[0], np.cumsum(self.dv)])
        self.chk_ptr = np.concatenate([[0], np.cumsum(self.dc)])
        # degree groups: for every degree d, the nodes having it and their edge ids [nodes, d]
        self.var_groups = []
        for d in np.unique(self.dv):
            vs = np.nonzero(self.dv == d)[0]
            vm = self.var_ptr[vs][:, None] + np.arange(d)[None, :]      # vm ids, ascending check
            self.var_groups.append((int(d), vs, vm, self.cm_of_vm[vm] if d else vm))
        self.chk_groups = []
        for d in np.unique(self.dc):
            cs = np.nonzero(self.dc == d)[0]
            cm = self.chk_ptr[cs][:, None] + np.arange(d)[None, :]      # cm ids, ascending variable
            self.chk_groups.append((int(d), cs, cm, self.vm_of_cm[cm] if d else cm))


def _others(vals, op):
    """vals: list of d arrays. Returns list of d arrays: op over the others with the
    documented two-sweep association (None where there are no others)."""
    d = len(vals)
    pre = [None] * d
    acc = None
    for k in range(d):
        pre[k] = acc
        acc = vals[k] if acc is None else op(acc, vals[k])
    suf = [None] * d
    acc = None
    for k in range(d - 1, -1, -1):
        suf[k] = acc
        acc = vals[k] if acc is None else op(vals[k], acc)
    out = []
    for k in range(d):
        if pre[k] is None:
            out.append(suf[k])
        elif suf[k] is None:
            out.append(pre[k])
        else:
            out.append(op(pre[k], suf[k]))
    return out


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def _tanh(a):
    return torch.tanh(_t(a)).numpy()


def _two_atanh(p):
    tp = _t(p)
    return torch.div(1 + tp, 1 - tp).log_().numpy()        # bp_cv.py:50


def weights_from_reference_state(g, state, iterations):
    """Dense parameters of a reference BeliefPropagation (bp/bp.py:26-39; keys layers.{i}.0.input_weight [E,E],
    layers.{i}.0.llr_weight [1,n], final_layer.0.input_weight [n,E], final_layer.0.llr_weight [1,n]) -> the sparse
    tables of the weighted decoder: w_edge [iters][E, max_dv] (row = variable-major OUT edge, column j = weight
    of the variable's j-th edge as INPUT, bp_vc.py:19: output_vm = input_cm @ (mask * weight)^T),
    w_llr [iters][n], wf_edge [E] (variable-major), wf_llr [n]."""
    mdv = int(g.dv.max())
    E, n = g.E, g.n
    w_edge = np.ones((iterations, E, mdv), F32)
    w_llr = np.ones((iterations, n), F32)
    A = lambda k: np.asarray(state[k].detach().cpu().numpy() if hasattr(state[k], "detach") else state[k], dtype=F32)
    for i in range(iterations):
        W = A(f"layers.{i}.0.input_weight")
        w_llr[i] = A(f"layers.{i}.0.llr_weight").reshape(-1)
        for v in range(n):
            b, d = int(g.var_ptr[v]), int(g.dv[v])
            for k in range(d):
                for j in range(d):
                    if j != k:
                        w_edge[i, b + k, j] = W[b + k, g.cm_of_vm[b + j]]
    Wf = A("final_layer.0.input_weight")
    wf_edge = np.ones(E, F32)
    for v in range(n):
        b, d = int(g.var_ptr[v]), int(g.dv[v])
        for k in range(d):
            wf_edge[b + k] = Wf[v, g.cm_of_vm[b + k]]
    return dict(w_edge=w_edge, w_llr=w_llr, wf_edge=wf_edge, wf_llr=A("final_layer.0.llr_weight").reshape(-1).astype(F32))


def bp_decode(H, llr, iterations, clamp_value, update="sp", x0=None, alpha=1.0, beta=0.0,
              graph=None, trace=False, early_exit=False, weights=None):
    """Decode a batch.  llr: [B,n] (log P1/P0, the callers' convention, ofdm_functions.py:72).

    Returns dict with
      x      [B,E] f32  final C->V messages, check-major
      t      [B,n] f32  pre-sigmoid marginal (bp.py:36-37); posterior log(P1/P0) = -2 t
      prob   [B,n] f32  P(bit=1) = 1 - sigmoid(t)           (bp.py:51)
      hard   [B,n] u8   np.round(prob)  (ties -> 0)         (ofdm_functions.py:161)
      syndrome [B] i32  number of unsatisfied checks of `hard`
    weights (weights_from_reference_state): the reference's trainable weights (bp_vc.py:16-32):
      V->C  0.5 * ( fl(w_llr[v] * llr') + ((w_0 x_0) + (w_1 x_1)) + ... over the OTHER edges, ascending ),
      marginal 0.5 * ( fl(wf_llr[v] * llr') + ((wf_0 x_0) + (wf_1 x_1)) + ... ); every product and sum rounded
      to fp32 separately.  With all-ones weights this equals the unweighted definition for dv <= 3.
    early_exit (NOT in the reference, bp.py:46-47 runs a fixed count): after every iteration but
    the last, rows whose hard decision satisfies every check are frozen (their messages stop
    changing); out["iters_used"] [B] i32 is the number of iterations each row ran.
    """
    if early_exit:
        return _bp_decode_early_exit(H, llr, iterations, clamp_value, update, alpha, beta, graph)
    g = graph or Graph(H)
    upd = _UPDATE_IDS[update] if isinstance(update, str) else int(update)
    llr = np.ascontiguousarray(llr, dtype=F32)
    B = llr.shape[0]
    Lp = (-llr).astype(F32)                                 # bp.py:47  layer([x, -llr])
    x = np.zeros((B, g.E), F32) if x0 is None else np.array(x0, dtype=F32)
    y = np.zeros((B, g.E), F32)
    cl = F32(clamp_value)
    half = F32(0.5)
    traces = []
    for it in range(int(iterations)):
        # ---- V -> C (variable-major output) ----
        for d, vs, vm, cm in g.var_groups:
            if d == 0:
                continue
            xin = [x[:, cm[:, k]] for k in range(d)]
            if weights is not None:
                wl = (weights["w_llr"][it][vs][None, :] * Lp[:, vs]).astype(F32)
                for k in range(d):
                    acc = None
                    for j in range(d):
                        if j == k:
                            continue
                        term = (weights["w_edge"][it][vm[:, k], j][None, :] * xin[j]).astype(F32)
                        acc = term if acc is None else (acc + term).astype(F32)
                    if acc is None:
                        acc = np.zeros((B, vs.size), F32)
                    s = (wl + acc).astype(F32)
                    y[:, vm[:, k]] = _tanh(half * s) if upd == UPDATE_SP else s
                continue
            if d == 1:
                S = [np.zeros((B, vs.size), F32)]
            else:
                S = _others(xin, np.add)
            for k in range(d):
                s = (Lp[:, vs] + S[k]).astype(F32)
                y[:, vm[:, k]] = _tanh(half * s) if upd == UPDATE_SP else s
        # ---- C -> V (check-major output) ----
        for d, cs, cm, vm in g.chk_groups:
            if d == 0:
                continue
            yin = [y[:, vm[:, j]] for j in range(d)]
            if upd == UPDATE_SP:
                if d == 1:
                    P = [np.ones((B, cs.size), F32)]         # empty product (mask all ones)
                else:
                    P = _others(yin, np.multiply)
                for j in range(d):
                    p = np.clip(P[j], -P_CLAMP, P_CLAMP).astype(F32)
                    x[:, cm[:, j]] = np.clip(_two_atanh(p), -cl, cl)
            else:
                mag = [np.abs(v) for v in yin]
                sgn = [np.signbit(v) for v in yin]
                if d == 1:
                    M = [np.full((B, cs.size), np.inf, F32)]
                    Sg = [np.zeros((B, cs.size), bool)]
                else:
                    M = _others(mag, np.minimum)
                    Sg = _others(sgn, np.logical_xor)
                for j in range(d):
                    mj = M[j]
                    if upd == UPDATE_NMS:
                        mj = (F32(alpha) * mj).astype(F32)
                    elif upd == UPDATE_OMS:
                        mj = np.maximum((mj - F32(beta)).astype(F32), F32(0))
                    mj = np.minimum(mj, cl)
                    x[:, cm[:, j]] = np.where(Sg[j], -mj, mj)
        if trace:
            traces.append(x.copy())
    # ---- marginal (bp.py:36-39,51) ----
    t = np.zeros((B, g.n), F32)
    for d, vs, vm, cm in g.var_groups:
        acc = None
        for k in range(d):
            term = x[:, cm[:, k]]
            if weights is not None:
                term = (weights["wf_edge"][vm[:, k]][None, :] * term).astype(F32)
            acc = term if acc is None else (acc + term).astype(F32)
        if acc is None:
            acc = np.zeros((B, vs.size), F32)
        lp = Lp[:, vs] if weights is None else (weights["wf_llr"][vs][None, :] * Lp[:, vs]).astype(F32)
        t[:, vs] = half * (lp + acc)
    prob = (-1 * torch.sigmoid(_t(t)) + 1).numpy()          # bp.py:51
    hard = np.round(prob).astype(np.uint8)                  # ofdm_functions.py:161
    synd = syndrome_weight(g, hard)
    out = dict(x=x, t=t, prob=prob, hard=hard, syndrome=synd)
    if trace:
        out["trace"] = traces
    return out


def _bp_decode_early_exit(H, llr, iterations, clamp_value, update, alpha, beta, graph):
    g = graph or Graph(H)
    llr = np.ascontiguousarray(llr, dtype=F32)
    B = llr.shape[0]
    x = np.zeros((B, g.E), F32)
    used = np.full(B, int(iterations), np.int32)
    running = np.arange(B)
    for it in range(1, int(iterations) + 1):
        if running.size == 0:
            break
        o = bp_decode(H, llr[running], 1, clamp_value, update=update, x0=x[running], alpha=alpha, beta=beta, graph=g)
        x[running] = o["x"]
        if it < iterations:
            done = o["syndrome"] == 0
            used[running[done]] = it
            running = running[~done]
    out = bp_decode(H, llr, 0, clamp_value, update=update, x0=x, alpha=alpha, beta=beta, graph=g)   # marginal of the frozen messages
    out["iters_used"] = used
    return out


def syndrome_weight(g, hard):
    par = np.zeros((hard.shape[0], g.m), dtype=np.uint8)
    np.bitwise_xor.at(par, (slice(None), g.cm_row), hard[:, g.cm_col])
    return par.sum(1).astype(np.int32)


def decode_bits(llrs, H, bp_iterations, batch_size, clamp_value, **kw):
    """ofdm_functions.py:131-163 - batching wrapper: float64 0/1 output, ragged tail
    (N % batch_size rows) left at zero because the reference silently skips it."""
    llrs = np.asarray(llrs)
    out = np.zeros(llrs.shape)
    g = Graph(H)
    for b in range(llrs.shape[0] // batch_size):
        s, e = b * batch_size, (b + 1) * batch_size
        out[s:e] = bp_decode(H, llrs[s:e].astype(F32), bp_iterations, clamp_value, graph=g, **kw)["hard"]
    return out


# ------------------------------------------------------------------------------------------
# "min-sum f16": the definition of the f16x2 fast path (csrc/decode_qc_h2.cu).  NOT in the
# reference (no min-sum, no half precision there: bp/bp.py:27-31) - this function IS the spec.
# Every operation is an IEEE binary16 operation with round-to-nearest-even, emulated exactly:
# operands are binary16 values held in float64 (sum / product of two binary16 values is exact
# in float64), rounded ONCE by float64 -> float16.
# ------------------------------------------------------------------------------------------
def _r16(x):
    return np.asarray(x, dtype=np.float64).astype(np.float16).astype(np.float64)


def _boxmin(a, b):
    """sign(a) sign(b) min(|a|,|b|) with IEEE sign bits (min.xorsign.abs)."""
    m = np.minimum(np.abs(a), np.abs(b))
    return np.where(np.logical_xor(np.signbit(a), np.signbit(b)), -m, m)


def bp_decode_f16(H, llr, iterations, clamp_value, update="minsum", alpha=1.0, graph=None):
    """Returns dict t [B,n] f32 (binary16 values), hard u8, syndrome i32, prob f32."""
    g = graph or Graph(H)
    upd = _UPDATE_IDS[update] if isinstance(update, str) else int(update)
    assert upd in (UPDATE_MINSUM, UPDATE_NMS)
    llr = np.ascontiguousarray(llr, dtype=F32)
    B = llr.shape[0]
    Lh = _r16(np.clip(llr, -32768.0, 32768.0))
    Lp = -Lh
    c_h = float(_r16(np.float32(clamp_value)))
    a_h = float(_r16(np.float32(alpha)))
    x = np.zeros((B, g.E), np.float64)
    y = np.zeros((B, g.E), np.float64)
    add = lambda p, q: _r16(p + q)
    for _ in range(int(iterations)):
        for d, vs, vm, cm in g.var_groups:
            if d == 0:
                continue
            xin = [x[:, cm[:, k]] for k in range(d)]
            S = [np.zeros((B, vs.size))] if d == 1 else _others(xin, add)
            for k in range(d):
                y[:, vm[:, k]] = add(Lp[:, vs], S[k])
        for d, cs, cm, vm in g.chk_groups:
            if d == 0:
                continue
            yin = [y[:, vm[:, j]] for j in range(d)]
            if upd == UPDATE_NMS:
                yin = [_r16(a_h * v) for v in yin]
            if d == 1:
                O_ = [np.full((B, cs.size), np.inf)]
            else:
                O_ = _others(yin, _boxmin)
            for j in range(d):
                x[:, cm[:, j]] = _boxmin(O_[j], c_h)
    t = np.zeros((B, g.n), np.float64)
    for d, vs, vm, cm in g.var_groups:
        acc = None
        for k in range(d):
            acc = x[:, cm[:, k]] if acc is None else add(acc, x[:, cm[:, k]])
        if acc is None:
            acc = np.zeros((B, vs.size))
        t[:, vs] = _r16(0.5 * add(Lp[:, vs], acc))
    t32 = t.astype(F32)
    prob = (-1 * torch.sigmoid(_t(t32)) + 1).numpy()
    hard = np.round(prob).astype(np.uint8)
    return dict(t=t32, prob=prob, hard=hard, syndrome=syndrome_weight(g, hard))


def bp_weighted_grad(H, llr, iterations, clamp_value, weights, grad_prob, graph=None, dtype=torch.float64):
    """Gradient oracle of the weighted decoder: the forward of bp_decode(weights=...) restated with differentiable torch
    ops (bp/bp_vc.py:16-32, nn.Tanh bp/bp.py:29, bp/bp_cv.py:22-55, clamp bp/bp.py:47, final layer bp/bp.py:36-39,51)
    in float64, differentiated by autograd.  Every clamp is torch.clamp (gradient passes inside and on the bound), so
    a saturated product has zero gradient; the reference's hand-written CV backward (bp_cv.py:57-91) keeps
    2/(1-q^2) there - the two agree wherever nothing saturates, which is where make_golden_grad.py pins this function
    against the reference's own .backward().
    Returns dict(prob, grad_llr [B,n], g_w_edge [iters,E,max_dv], g_w_llr [iters,n], g_wf_edge [E], g_wf_llr [n])."""
    g = graph or Graph(H)
    E, n = g.E, g.n
    mdv, mdc = int(g.dv.max()), int(g.dc.max())
    vm_var = np.repeat(np.arange(n), g.dv)
    pos = np.arange(E) - g.var_ptr[vm_var]
    vidx = np.zeros((E, mdv), np.int64); vmask = np.zeros((E, mdv), bool)
    for j in range(mdv):
        has = g.dv[vm_var] > j
        vidx[has, j] = g.cm_of_vm[g.var_ptr[vm_var[has]] + j]
        vmask[:, j] = has & (pos != j)
    cm_chk = np.repeat(np.arange(g.m), g.dc)
    cpos = np.arange(E) - g.chk_ptr[cm_chk]
    cidx = np.zeros((E, mdc), np.int64); cmask = np.zeros((E, mdc), bool)
    for j in range(mdc):
        has = g.dc[cm_chk] > j
        cidx[has, j] = g.chk_ptr[cm_chk[has]] + j
        cmask[:, j] = has & (cpos != j)
    T = lambda a: torch.as_tensor(np.asarray(a))
    P = lambda a: torch.tensor(np.asarray(a), dtype=dtype, requires_grad=True)
    L = P(llr)
    w_edge, w_llr, wf_edge, wf_llr = P(weights["w_edge"]), P(weights["w_llr"]), P(weights["wf_edge"]), P(weights["wf_llr"])
    vidx_t, vmask_t, cidx_t, cmask_t = T(vidx), T(vmask).to(dtype), T(cidx), T(cmask)
    vm_var_t, cm_of_vm_t, vm_of_cm_t = T(vm_var), T(g.cm_of_vm), T(g.vm_of_cm)
    lim = float(P_CLAMP)
    x = torch.zeros(L.shape[0], E, dtype=dtype)
    Lp = -L
    for it in range(iterations):
        acc = (w_edge[it, :, :mdv] * vmask_t * x[:, vidx_t]).sum(-1)                     # [B, E] variable-major
        u_vm = torch.tanh(0.5 * (w_llr[it][vm_var_t] * Lp[:, vm_var_t] + acc))
        u_cm = u_vm[:, vm_of_cm_t]
        p = torch.where(cmask_t, u_cm[:, cidx_t], torch.ones((), dtype=dtype)).prod(-1)
        p = torch.clamp(p, -lim, lim)
        x = torch.clamp(torch.log((1 + p) / (1 - p)), -float(clamp_value), float(clamp_value))
    t_edges = wf_edge * x[:, cm_of_vm_t]                                                  # [B, E] variable-major
    t = 0.5 * (wf_llr * Lp + torch.zeros(L.shape[0], n, dtype=dtype).index_add(1, vm_var_t, t_edges))
    prob = 1 - torch.sigmoid(t)
    prob.backward(torch.as_tensor(np.asarray(grad_prob), dtype=dtype))
    G = lambda p: (torch.zeros_like(p) if p.grad is None else p.grad).numpy()            # zero iterations: unused tables
    gw = G(w_edge).copy()
    gw[:, ~vmask] = 0                                                                     # unused (k,k) / padding entries
    return dict(prob=prob.detach().numpy(), grad_llr=L.grad.numpy(), g_w_edge=gw, g_w_llr=G(w_llr),
                g_wf_edge=wf_edge.grad.numpy(), g_wf_llr=wf_llr.grad.numpy())
