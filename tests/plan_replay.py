"""Replays the device sweep's streaming plan on the CPU with numpy (test infrastructure).

The plan (mast-multiphysics_b200/csrc/plan.h) is pure integer bookkeeping: state slots, pair-block slots with
orientation / first-touch flags, per-node ranks, flush rows.  This replay executes it literally -- shared-memory
arena poisoned with NaN, slots overwritten exactly when the device would overwrite them -- with the element
arithmetic taken from the oracle's FE tables, and assembles the node-block CSR the device kernel would write.
A wrong slot lifetime, orientation or rank shows up as NaN / a mismatch against the oracle's Jacobian.
"""
import ctypes as C

import numpy as np

import oracle_lib as O


def _fe(prob, e):
    nn, nq = 8, 8
    coords = np.ascontiguousarray(prob.xyz[prob.conn[e]])
    phi = np.empty((nn, nq)); dphi = np.empty((nn, nq, 3)); jxw = np.empty(nq); xq = np.empty((nq, 3))
    O.lib().oracle_fe_tables(8, coords.ctypes.data_as(C.c_void_p), phi.ctypes.data_as(C.c_void_p),
                             dphi.ctypes.data_as(C.c_void_p), jxw.ctypes.data_as(C.c_void_p),
                             xq.ctypes.data_as(C.c_void_p))
    return phi, dphi, jxw


def replay(prob, plan, ctx, X=None, late_flush=True):
    """Returns (J node-block CSR values in device layout, R or None, per-node flush count)."""
    nv = prob.nv
    E0, nu = prob.mat["E0"], prob.mat["nu"]
    lam1 = nu / (1. - nu - 2. * nu * nu)
    mu1 = 1. / 2. / (1. + nu)
    rp, ci = ctx.get_pattern()
    rows = ctx.get_row_dofs()
    row_of_node = {int(d // nv): k // nv for k, d in enumerate(rows) if d % nv == 0}
    nbrs = {}
    for n, r in row_of_node.items():
        cols = ci[rp[r * nv]:rp[r * nv + 1]]
        nbrs[n] = (cols[::nv] // nv).astype(np.int64)
    elem_of = {tuple(int(x) for x in c): e for e, c in enumerate(prob.conn)}
    cm = np.zeros(prob.n_nodes, np.int64)
    if prob.constrained is not None:
        cc = prob.constrained.reshape(prob.n_nodes, nv)
        for v in range(nv):
            cm |= cc[:, v].astype(np.int64) << v
    diag = prob.diag_elem if prob.diag_elem is not None else np.zeros(prob.n_elem, np.uint8)
    nadj = np.zeros(prob.n_nodes, np.int64); nadjd = np.zeros(prob.n_nodes, np.int64)
    for e, c in enumerate(prob.conn):
        (nadjd if diag[e] else nadj)[c] += 1
    J = np.full(len(ci), np.nan)
    R = None if X is None else np.full(prob.n_dof, np.nan)
    flushed = np.zeros(prob.n_nodes, np.int64)
    visited_by_node = {}
    tri = lambda a, b: a * (15 - a) // 2 + b

    T, S, V = plan["tiles"], plan["steps"], plan["visits"]
    FR, FE, LD = plan["flush_rows"], plan["flush_ent"], plan["loads"]
    for t in T:
        arena = np.full((plan["arena_blocks"], 3, 3), np.nan)
        blk_key = [None] * plan["arena_blocks"]
        slot_gid = np.full(plan["max_slots"], -1, np.int64)
        slot_owned = np.zeros(plan["max_slots"], bool)
        cnt = np.zeros(plan["max_slots"], np.int64)
        steps = S[t["step_begin"]:t["step_begin"] + t["n_steps"]]

        def load(sd):
            for lr in LD[sd["load_begin"]:sd["load_begin"] + sd["n_loads"]]:
                slot_gid[lr["slot"]] = lr["gid"]; slot_owned[lr["slot"]] = bool(lr["owned"]); cnt[lr["slot"]] = 0

        def visits(sd):
            # grid [worker][position]; accumulation order (position, worker); empty positions have ln[0] == 0xFFFF
            K = int(sd["n_pos"])
            grid = V[sd["visit_begin"]:sd["visit_begin"] + plan["n_workers"] * K].reshape(plan["n_workers"], K)
            seq = [grid[w, k] for k in range(K) for w in range(plan["n_workers"]) if grid[w, k]["ln"][0] != 0xFFFF]
            assert len(seq) == sd["n_visits"]
            for k in range(K):   # visits at the same position share no owned node
                seen = set()
                for w in range(plan["n_workers"]):
                    v = grid[w, k]
                    if v["ln"][0] == 0xFFFF:
                        continue
                    mine = {int(v["ln"][a]) for a in range(8) if v["rank"][a] != 0xFF}
                    assert not (mine & seen), "visits at one position share an owned node"
                    seen |= mine
            for v in seq:
                gids = slot_gid[v["ln"]]
                e = elem_of[tuple(int(g) for g in gids)]
                assert not diag[e]
                phi, dphi, jxw = _fe(prob, e)
                w = jxw * E0
                Sab = np.einsum("q,aqi,bqj->abij", w, dphi, dphi)
                for a in range(8):
                    own = v["rank"][a] != 0xFF
                    assert own == slot_owned[v["ln"][a]]
                    if own:
                        assert cnt[v["ln"][a]] == v["rank"][a], "accumulation rank out of order"
                for a in range(8):
                    for b in range(a, 8):
                        ent = int(v["pair"][tri(a, b)])
                        own_any = v["rank"][a] != 0xFF or v["rank"][b] != 0xFF
                        if ent == 0xFFFF:
                            assert not own_any
                            continue
                        assert own_any
                        bs = ent & 0xFFF
                        contrib = Sab[b, a] if ent & 0x1000 else Sab[a, b]
                        key = frozenset((int(gids[a]), int(gids[b])))
                        if ent & 0x2000:
                            arena[bs] = contrib; blk_key[bs] = key
                        else:
                            assert blk_key[bs] == key, "pair block reused while live"
                            arena[bs] += contrib
                for a in range(8):
                    if v["rank"][a] != 0xFF:
                        cnt[v["ln"][a]] += 1
                        visited_by_node.setdefault(int(gids[a]), set()).add(e)

        def flush(sd):
            rows_ = FR[sd["flush_begin"]:sd["flush_begin"] + sd["n_flush"]]
            ents = FE[sd["fent_begin"]:sd["fent_begin"] + sd["n_fent"]]
            for fr in rows_:
                n = int(fr["gid"])
                assert slot_gid[fr["slot"]] == n and slot_owned[fr["slot"]]
                r = row_of_node[n]
                nb = nbrs[n]
                assert fr["nnb"] == len(nb) and fr["dst_block"] * nv * nv == rp[r * nv]
                assert fr["nadj"] == nadj[n] and fr["nadjd"] == nadjd[n]
                flushed[n] += 1
                rs = nv * len(nb)
                f = np.zeros(3)
                for c, q in enumerate(nb):
                    ent = int(ents[fr["ent_rel"] + c])
                    bs, tr, self_, ns = ent & 0xFFF, bool(ent & 0x1000), bool(ent & 0x2000), ent >> 16
                    assert self_ == (q == n)
                    if bs != 0xFFF:
                        assert blk_key[bs] == frozenset((n, int(q))), "flush reads a block of another pair"
                        Sb = arena[bs].copy()
                        if self_:
                            Sb[1, 0], Sb[2, 0], Sb[2, 1] = Sb[0, 1], Sb[0, 2], Sb[1, 2]
                        assert ns != 0xFFFF and slot_gid[ns] == q
                    else:
                        Sb = np.zeros((3, 3))
                    l1, l2 = (mu1, lam1) if tr else (lam1, mu1)
                    K = l1 * Sb + l2 * Sb.T + mu1 * np.trace(Sb) * np.eye(3)
                    cmm = 0
                    if ns != 0xFFFF:
                        assert slot_gid[ns] == q
                        cmm = cm[q]
                        if X is not None:
                            f += K @ X[q * nv:q * nv + 3]
                    for i in range(3):
                        for j in range(3):
                            k = K[i, j]
                            if ((cm[n] >> i) | (cmm >> j)) & 1:
                                k = 0.
                            if self_ and i == j:
                                k = float(nadj[n] + nadjd[n]) if (cm[n] >> i) & 1 else k + 1e-24 * nadjd[n]
                            J[rp[r * nv] + i * rs + nv * c + j] = k
                if R is not None:
                    for i in range(3):
                        R[n * nv + i] = 0. if (cm[n] >> i) & 1 else f[i]
        # device timeline (sweep_hex8.cu): the flush of step s runs concurrently with the visits of step s+1 and is
        # followed by the staging of step s+2.  Both extreme interleavings are legal and must give the same result.
        load(steps[0])
        if len(steps) > 1:
            load(steps[1])
        if late_flush:
            # visits running maximally ahead of the flush; state of step s+2 staged at the top of flush iteration s
            for si, sd in enumerate(steps):
                if si >= 1:
                    if si + 1 < len(steps):
                        load(steps[si + 1])
                visits(sd)
                if si >= 1:
                    flush(steps[si - 1])
            flush(steps[-1])
        else:
            # flush right behind the visits; state staged as early as the device may stage it
            for si, sd in enumerate(steps):
                if si + 2 < len(steps):
                    load(steps[si + 2])
                visits(sd)
                flush(sd)
    # every (non-diagonal) element adjacent to an owned node was visited on behalf of that node
    for n in row_of_node:
        want = {e for e in range(prob.n_elem) if n in prob.conn[e] and not diag[e]} if prob.n_elem < 400 else None
        if want is not None:
            assert visited_by_node.get(n, set()) == want
    return J, R, flushed
