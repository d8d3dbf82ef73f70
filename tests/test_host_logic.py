"""CPU-side checks of the product's host logic: the C-ABI library loads and exports every symbol the
header declares, the sparsity pattern it builds is bit-identical to the oracle's (libMesh rule), the
assembly plan is consistent, and every sweep fails loudly without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

import mast_b200
import oracle_lib as O
from mast_b200 import mesh

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(built):
    lib = mast_b200.load_library()
    hdr = open(os.path.join(ROOT, "include", "mast_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(mastb200_[a-z0-9_]+)\s*\(", hdr)))
    assert declared, "no declarations found"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/mast_b200.h but not exported"
    assert sorted(mast_b200.EXPORTED_SYMBOLS) == declared
    assert b"sm_100a" in lib.mastb200_version()


@pytest.mark.parametrize("nv", [3, 6])
@pytest.mark.parametrize("shape", [(5, 3, 4), (1, 1, 1), (9, 8, 7)])
def test_pattern_bit_exact_vs_oracle(built, nv, shape):
    xyz, conn = mesh.build_cube(*shape, jitter=0.2)
    p = O.Problem(O.HEX8, xyz, conn, nv)
    ctx = p.make_context(-1)
    rp, ci = ctx.get_pattern()
    orp, oci = p.pattern()
    assert rp.dtype == np.int64 and ci.dtype == np.int32
    assert np.array_equal(rp, orp) and np.array_equal(ci, oci)
    assert ctx.nnz == (2 * shape[0] + 1 if False else len(oci))
    if shape == (5, 3, 4) and nv == 3:  # closed form for a structured block: (3nx+1-... ) see BASELINE.md
        assert ctx.nnz == (3 * 5 + 1) * (3 * 3 + 1) * (3 * 4 + 1) * 9


def test_pattern_with_permuted_unstructured_numbering(built):
    xyz, conn = mesh.build_cube(4, 4, 3, jitter=0.2)
    rng = np.random.default_rng(0)
    perm = rng.permutation(len(xyz))            # new id of old node
    xyz2 = np.empty_like(xyz); xyz2[perm] = xyz
    conn2 = perm[conn].astype(np.int32)[rng.permutation(len(conn))]
    p = O.Problem(O.HEX8, xyz2, conn2, 3)
    ctx = p.make_context(-1)
    rp, ci = ctx.get_pattern()
    orp, oci = p.pattern()
    assert np.array_equal(rp, orp) and np.array_equal(ci, oci)


def test_owned_subset_rows(built):
    # rank-local view: only owned nodes produce rows; ghosts are columns only
    xyz, conn = mesh.build_cube(4, 3, 3)
    owned = (xyz[:, 2] <= 0.5).astype(np.uint8)
    ctx = mast_b200.Context(-1)
    ctx.set_mesh(8, xyz, conn); ctx.set_dof_map(3, None, owned)
    n_rows, nnz = ctx.build_pattern()
    assert n_rows == 3 * owned.sum()
    rows = ctx.get_row_dofs()
    assert np.array_equal(rows, (3 * np.flatnonzero(owned)[:, None] + np.arange(3)).ravel())
    p = O.Problem(O.HEX8, xyz, conn, 3)
    orp, oci = p.pattern()
    rp, ci = ctx.get_pattern()
    for k, r in enumerate(rows[::7]):
        k = k * 7
        assert np.array_equal(ci[rp[k]:rp[k + 1]], oci[orp[r]:orp[r + 1]])


def test_plan_statistics(built):
    xyz, conn = mesh.build_cube(24, 24, 24)
    ctx = mast_b200.Context(-1)
    ctx.set_mesh(8, xyz, conn); ctx.set_dof_map(3)
    ctx.build_pattern()
    s = ctx.stats()
    assert s["n_owned_nodes"] == 25 ** 3 and s["n_elem"] == 24 ** 3
    assert s["n_visits"] >= s["n_elem"] and s["n_visits"] < 2 * s["n_elem"]
    assert s["smem_bytes"] <= 232448
    for aux in (0, 1):
        P = ctx.debug_plan(aux)
        assert P["tiles"]["n_owned"].sum() == 25 ** 3          # every owned node belongs to exactly one strip
        assert P["steps"]["n_flush"].sum() == 25 ** 3          # ... and is written exactly once
        assert np.all(P["steps"]["fent_begin"] % 4 == 0)
        real = P["visits"][P["visits"]["ln"][:, 0] != 0xFFFF]          # the grid has empty positions
        assert len(real) == P["tiles"]["n_visits"].sum() == P["steps"]["n_visits"].sum()
        assert len(P["visits"]) == (P["steps"]["n_pos"] * P["n_workers"]).sum()
        assert len(real) >= 0.7 * len(P["visits"])                     # balanced: few idle grid positions
        used = real["pair"][real["pair"] != 0xFFFF] & 0xFFF
        assert used.max() < P["arena_blocks"] and real["ln"].max() < P["max_slots"]
        assert np.all(np.diff(P["tiles"]["n_visits"]) <= 0)   # longest strips first for the dynamic scheduler


def _replay_case(prob, X):
    import plan_replay
    ctx = prob.make_context(-1)
    rp, ci = ctx.get_pattern()
    Ro, Jo = prob.residual_and_jacobian(X)
    rows = ctx.get_row_dofs()
    orp, _ = prob.pattern()
    Jref = np.concatenate([Jo[orp[r]:orp[r + 1]] for r in rows])
    for aux, late in ((0, True), (0, False), (1, True)):
        J, R, flushed = plan_replay.replay(prob, ctx.debug_plan(aux), ctx, X, late_flush=late)
        own = np.zeros(prob.n_nodes, bool); own[rows[::prob.nv] // prob.nv] = True
        assert np.array_equal(flushed, own.astype(np.int64))
        assert not np.isnan(J).any()
        assert np.abs(J - Jref).max() <= 1e-12 * np.abs(Jref).max()
        od = np.repeat(own, prob.nv)
        assert np.abs(R[od] - Ro[od]).max() <= 1e-12 * np.abs(Ro).max()
    return ctx


def test_plan_replay_structured_and_jittered(built):
    # the streaming plan executed literally on the CPU reproduces the oracle's Jacobian and internal force
    for shape, jit in (((7, 5, 11), 0.0), ((12, 9, 4), 0.2), ((1, 1, 1), 0.0), ((2, 1, 30), 0.2)):
        xyz, conn = mesh.build_cube(*shape, jitter=jit)
        con = mesh.dirichlet_flags(len(xyz), 3, mesh.boundary_nodes(*shape, 4))
        prob = O.Problem(O.HEX8, xyz, conn, 3, constrained=con)
        _replay_case(prob, O.state(prob.n_dof, 1))


def test_plan_replay_unstructured_diag_and_ghosts(built):
    xyz, conn = mesh.build_cube(6, 5, 4, jitter=0.2)
    rng = np.random.default_rng(0)
    perm = rng.permutation(len(xyz))
    xyz2 = np.empty_like(xyz); xyz2[perm] = xyz
    conn2 = perm[conn].astype(np.int32)[rng.permutation(len(conn))]
    con = (rng.uniform(size=len(xyz) * 3) < 0.1).astype(np.uint8)
    diag = np.zeros(len(conn2), np.uint8); diag[[0, 7, 8, 30]] = 1
    prob = O.Problem(O.HEX8, xyz2, conn2, 3, constrained=con, diag_elem=diag)
    _replay_case(prob, O.state(prob.n_dof, 1))
    # rank-local view: rows of owned nodes only
    import plan_replay
    owned = (xyz2[:, 2] < 0.55).astype(np.uint8)
    ctx = mast_b200.Context(-1)
    ctx.set_mesh(8, prob.xyz, prob.conn, diag); ctx.set_dof_map(3, con, owned)
    ctx.build_pattern()
    X = O.state(prob.n_dof, 1)
    J, R, flushed = plan_replay.replay(prob, ctx.debug_plan(0), ctx, X)
    assert np.array_equal(flushed, owned.astype(np.int64))
    Ro, Jo = prob.residual_and_jacobian(X)
    orp, _ = prob.pattern()
    Jref = np.concatenate([Jo[orp[r]:orp[r + 1]] for r in ctx.get_row_dofs()])
    assert np.abs(J - Jref).max() <= 1e-12 * np.abs(Jref).max()


def test_plan_small_arena_forces_splits(built, monkeypatch):
    # anisotropic spacing + a tiny target cross-section: strips get split until they fit, result unchanged
    monkeypatch.setenv("MASTB200_PLAN_COLS", "400")
    xyz, conn = mesh.build_cube(5, 16, 16, 1.0, 0.1, 3.0, jitter=0.2)
    prob = O.Problem(O.HEX8, xyz, conn, 3)
    _replay_case(prob, O.state(prob.n_dof, 1))


def test_no_cpu_fallback(built):
    p = O.cantilever(2, 2, 2)
    ctx = p.make_context(-1)
    X = O.state(p.n_dof)
    R = np.zeros(p.n_dof)
    with pytest.raises(mast_b200.MastB200Error) as ei:
        ctx.residual_and_jacobian(X, R, None)
    assert ei.value.code == 2 and "no CPU fallback" in str(ei.value)
    with pytest.raises(mast_b200.MastB200Error):
        ctx.compliance(X)
    # every compute entry point added later refuses the same way (planning-only context = no device)
    for call in (lambda: ctx.sensitivity_sweep(X, X, np.zeros(p.n_nodes)),
                 lambda: ctx.stress(X, von_mises=np.zeros(p.n_elem * 8)),
                 lambda: ctx.stress_functional(X, 2.0, 0.0, 1.0),
                 lambda: ctx.stress_functional_derivative(X, 2.0, 0.0, 1.0, np.ones(2), np.zeros(p.n_dof)),
                 lambda: ctx.stress_functional_density_derivative(X, 2.0, 0.0, 1.0, np.ones(2), np.zeros(p.n_nodes)),
                 lambda: ctx.compliance_derivative(X, np.zeros(p.n_dof)),
                 lambda: ctx.res_l2_norm()):
        with pytest.raises(mast_b200.MastB200Error) as ei:
            call()
        assert ei.value.code == 2
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(mast_b200.MastB200Error) as ei:
            mast_b200.Context(0)
        assert ei.value.code == 2


def test_bad_arguments(built):
    ctx = mast_b200.Context(-1)
    with pytest.raises(mast_b200.MastB200Error):
        ctx.set_dof_map(3)  # before set_mesh
    xyz, conn = mesh.build_cube(2, 2, 2)
    bad = conn.copy(); bad[0, 0] = 10 ** 6
    ctx.set_mesh(8, xyz, bad); ctx.set_dof_map(3)
    with pytest.raises(mast_b200.MastB200Error):
        ctx.build_pattern()
    with pytest.raises(mast_b200.MastB200Error):
        ctx.set_mesh(5, xyz, conn)
