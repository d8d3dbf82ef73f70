"""Parity of the implicit-connectivity HEX8 sweeps (sweep_hex8_struct.cu, and the opt-in column march of
sweep_hex8_march.cu: build_cube numbering recognised by the host) against the CPU oracle, through the C-ABI, and
agreement with the plan-driven sweep on the same inputs.  Every test runs once per kernel."""
import numpy as np
import pytest

import oracle_lib as O
from submesh import close_entries
from mast_b200 import mesh

pytestmark = pytest.mark.gpu
RTOL = 1e-12


@pytest.fixture(autouse=True, params=["lockstep", "march", "slots"])
def struct_kernel(request, monkeypatch):
    monkeypatch.setenv("MASTB200_STRUCT_KERNEL", request.param)
    return request.param


def _is_struct(ctx, kernel, nv=3, fused=False):
    """stats()['structured']: 1 = lock-step kernel, 2 = column march, 3 = slot kernel (the last two serve the 3-dof system
    when opted in; sensitivities fused into the launch stay with the lock-step kernel)."""
    want = 1
    if nv == 3 and kernel == "march":
        want = 2
    if nv == 3 and kernel == "slots" and not fused:
        want = 3
    return ctx.stats()["structured"] == want


def _close(a, b, rtol=RTOL):
    """Entry-wise: |a - b| <= rtol |b| + 64 eps (rtol / 1e-12) max|b| -- relative 1e-12 for every entry, with the floor an
    entry that cancels to ~0 cannot beat (eps times the magnitude of the <= 64 terms it is the sum of)."""
    close_entries(a, b, rtol=rtol, terms=64 * rtol / 1e-12)


def _problem(nx, ny, nz, jitter=0.2, constrained=True, seed=0):
    xyz, conn = mesh.build_cube(nx, ny, nz, 0.3, 0.2, 0.25, 8, jitter=jitter)
    con = None
    if constrained:
        con = mesh.dirichlet_flags(len(xyz), 3, mesh.boundary_nodes(nx, ny, nz, 4))
        rng = np.random.default_rng(seed)
        con = np.maximum(con, (rng.uniform(size=len(con)) < 0.05).astype(np.uint8))  # scattered single dofs too
    se, sid = mesh.boundary_sides(nx, ny, nz, 3)
    return O.Problem(O.HEX8, xyz, conn, 3, constrained=con, side_elem=se, side_id=sid, side_pressure=np.full(len(se), 5e7))


@pytest.mark.parametrize("shape", [(1, 1, 1), (2, 3, 1), (3, 1, 2), (7, 8, 9), (8, 8, 8), (9, 15, 5), (14, 6, 30), (20, 10, 10)])
def test_struct_residual_and_jacobian(built, shape, struct_kernel):
    prob = _problem(*shape)
    X = O.state(prob.n_dof, 1)
    ctx = prob.make_context(0)
    rp, ci = ctx.get_pattern()
    orp, oci = prob.pattern()
    assert np.array_equal(rp, orp) and np.array_equal(ci, oci)
    R = np.full(prob.n_dof, np.nan); J = np.full(ctx.nnz, np.nan)
    ctx.residual_and_jacobian(X, R, J)
    assert _is_struct(ctx, struct_kernel)
    Ro, Jo = prob.residual_and_jacobian(X)
    _close(R, Ro); _close(J, Jo)
    con = prob.constrained.astype(bool)
    assert np.all(R[con] == 0.0)
    # each alone, and bitwise reproducible
    R2 = np.full(prob.n_dof, np.nan); J2 = np.full(ctx.nnz, np.nan)
    ctx.residual_and_jacobian(X, R2, None)
    ctx.residual_and_jacobian(X, None, J2)
    assert np.array_equal(R2, R) and np.array_equal(J2, J)


def test_struct_matches_plan_driven_sweep(built, monkeypatch, struct_kernel):
    prob = _problem(13, 9, 11)
    X = O.state(prob.n_dof, 1)
    ctx = prob.make_context(0)
    R = np.empty(prob.n_dof); J = np.empty(ctx.nnz)
    ctx.residual_and_jacobian(X, R, J)
    assert _is_struct(ctx, struct_kernel)
    monkeypatch.setenv("MASTB200_NO_STRUCT", "1")
    ctx2 = prob.make_context(0)
    R2 = np.empty(prob.n_dof); J2 = np.empty(ctx2.nnz)
    ctx2.residual_and_jacobian(X, R2, J2)
    assert ctx2.stats()["structured"] == 0
    _close(R, R2); _close(J, J2)
    # a permuted numbering of the same mesh is not recognised and takes the plan-driven path
    monkeypatch.delenv("MASTB200_NO_STRUCT")
    perm = np.random.default_rng(0).permutation(len(prob.xyz))
    xyz2 = np.empty_like(prob.xyz); xyz2[perm] = prob.xyz
    p3 = O.Problem(O.HEX8, xyz2, perm[prob.conn].astype(np.int32), 3)
    ctx3 = p3.make_context(0)
    ctx3.residual_and_jacobian(O.state(p3.n_dof, 1), np.empty(p3.n_dof), np.empty(ctx3.nnz))
    assert ctx3.stats()["structured"] == 0


def test_struct_owned_z_slab_with_ghost_planes(built, struct_kernel):
    # a rank's local mesh in the z-slab partition: owned node planes [2, 6) of a 9-plane box, ghost planes around
    import mast_b200
    nx, ny, nz = 9, 8, 8
    prob = _problem(nx, ny, nz)
    X = O.state(prob.n_dof, 1)
    npxy = (nx + 1) * (ny + 1)
    owned = np.zeros(prob.n_nodes, np.uint8); owned[2 * npxy:6 * npxy] = 1
    ctx = mast_b200.Context(0)
    ctx.set_mesh(8, prob.xyz, prob.conn)
    ctx.set_dof_map(3, prob.constrained, owned)
    ctx.set_material(**prob.mat)
    ctx.build_pattern()
    R = np.full(prob.n_dof, np.nan); J = np.full(ctx.nnz, np.nan)
    ctx.residual_and_jacobian(X, R, J)
    assert _is_struct(ctx, struct_kernel)
    Ro, Jo = prob.residual_and_jacobian(X)
    rows = ctx.get_row_dofs()
    orp, _ = prob.pattern()
    Jo_rows = np.concatenate([Jo[orp[r]:orp[r + 1]] for r in rows])
    _close(J, Jo_rows)
    # the oracle's residual includes the pressure load; this context has none, so compare the internal part
    prob0 = O.Problem(O.HEX8, prob.xyz, prob.conn, 3, constrained=prob.constrained)
    R0, _ = prob0.residual_and_jacobian(X, want_J=False)
    _close(R[rows], R0[rows])
    ghost = np.ones(prob.n_dof, bool); ghost[rows] = False
    assert np.all(R[ghost] == 0.0)


def test_struct_large_mesh_properties(built, struct_kernel):
    # 96 x 70 x 50 jittered: symmetry of the assembled operator, rigid-body null space, K u == R (linear element)
    import scipy.sparse as sp
    nx, ny, nz = 96, 70, 50
    xyz, conn = mesh.build_cube(nx, ny, nz, 1.0, 0.8, 0.6, 8, jitter=0.2)
    prob = O.Problem(O.HEX8, xyz, conn, 3)
    ctx = prob.make_context(0)
    rp, ci = ctx.get_pattern()
    X = O.state(prob.n_dof, 1)
    R = np.empty(prob.n_dof); J = np.empty(ctx.nnz)
    ctx.residual_and_jacobian(X, R, J)
    assert _is_struct(ctx, struct_kernel)
    A = sp.csr_matrix((J, ci, rp), shape=(prob.n_dof, prob.n_dof))
    assert abs(A - A.T).max() <= 1e-12 * np.abs(J).max()
    assert np.abs(A @ X - R).max() <= 1e-12 * np.abs(R).max()
    rb = np.zeros((len(xyz), 3)); rb[:, 0] = 1.0 - 0.3 * xyz[:, 1]; rb[:, 1] = 0.3 * xyz[:, 0]  # translation + rotation
    assert np.abs(A @ rb.ravel()).max() <= 1e-9 * np.abs(J).max()


def test_struct_simp_modulus_and_fused_sensitivities(built, monkeypatch, struct_kernel):
    # SIMP modulus in the implicit-connectivity sweep; G requested together with R/J comes from the element-parallel
    # kernels right after it; same numbers as the plan-driven fused sweep
    prob = O.cantilever(11, 9, 8, jitter=0.2, simp=True)
    X = O.state(prob.n_dof, 1); lam = O.state(prob.n_dof, 2)
    n = prob.n_nodes
    ctx = prob.make_context(0)
    G = np.full(n, np.nan); R = np.full(prob.n_dof, np.nan); J = np.full(ctx.nnz, np.nan)
    ctx.sensitivity_sweep(X, lam, G, R, J)
    assert _is_struct(ctx, struct_kernel, fused=True)
    Ro, Jo = prob.residual_and_jacobian(X)
    _close(R, Ro); _close(J, Jo)
    g_ref = prob.adjoint_sensitivity(X, lam, np.arange(n + 1), np.arange(n), np.ones(n))
    _close(G, g_ref, 1e-11)
    assert np.isclose(ctx.res_l2_norm(), np.linalg.norm(R), rtol=1e-13)
    monkeypatch.setenv("MASTB200_NO_STRUCT", "1")
    ctx2 = prob.make_context(0)
    G2 = np.full(n, np.nan); R2 = np.full(prob.n_dof, np.nan); J2 = np.full(ctx2.nnz, np.nan)
    ctx2.sensitivity_sweep(X, lam, G2, R2, J2)
    assert ctx2.stats()["structured"] == 0
    _close(G2, G, 1e-11); _close(R2, R); _close(J2, J)


@pytest.mark.parametrize("shape", [(1, 1, 1), (3, 2, 2), (9, 8, 7), (16, 9, 12)])
def test_struct_nv6_reference_layout(built, shape, monkeypatch, struct_kernel):
    # MAST's own structural system has six variables per node: zero rotational rows/columns kept in the pattern,
    # 1e-20*max(diag K_e) summed on the rotational diagonal (solid_element_3d.cpp:413-417); SIMP modulus too
    nx, ny, nz = shape
    for simp in (False, True):
        prob = O.cantilever(nx, ny, nz, nv=6, jitter=0.2, simp=simp)
        X = O.state(prob.n_dof, 1)
        ctx = prob.make_context(0)
        rp, ci = ctx.get_pattern()
        orp, oci = prob.pattern()
        assert np.array_equal(rp, orp) and np.array_equal(ci, oci)
        R = np.full(prob.n_dof, np.nan); J = np.full(ctx.nnz, np.nan)
        ctx.residual_and_jacobian(X, R, J)
        assert _is_struct(ctx, struct_kernel, 6)
        Ro, Jo = prob.residual_and_jacobian(X)
        _close(R, Ro); _close(J, Jo)
        # the rotational diagonal is far below the tolerance of the whole-matrix comparison: check it on its own
        rows = np.repeat(np.arange(prob.n_dof), np.diff(rp))
        rot = (rows == ci) & (rows % 6 >= 3) & ~prob.constrained.astype(bool)[rows]
        assert rot.sum() > 0 and np.all(Jo[rot] > 0)
        assert np.abs(J[rot] - Jo[rot]).max() <= 1e-11 * np.abs(Jo[rot]).max()


def test_struct_random_shapes_slabs_and_layouts(built, struct_kernel):
    # random box shapes around the tile sizes (7 owned nodes / 8 elements per tile edge, z-segments), random owned
    # z-slabs, both dof layouts, SIMP or constant modulus: pattern bit-exact, values vs the oracle
    import mast_b200
    rng = np.random.default_rng(2026)
    for trial in range(10):
        nx, ny = int(rng.integers(1, 19)), int(rng.integers(1, 19))
        nz = int(rng.integers(1, 40))
        nv = int(rng.choice([3, 6]))
        simp = bool(rng.integers(0, 2))
        xyz, conn = mesh.build_cube(nx, ny, nz, 0.3 + 0.05 * nx, 0.2 + 0.05 * ny, 0.1 + 0.05 * nz, 8, jitter=0.2)
        con = (rng.uniform(size=len(xyz) * nv) < 0.08).astype(np.uint8)
        if nv == 6:
            con.reshape(-1, 6)[:, 3:] = (rng.uniform(size=(len(xyz), 3)) < 0.03)
        rho = 0.2 + 0.6 * rng.uniform(size=len(xyz)) if simp else None
        prob = O.Problem(O.HEX8, xyz, conn, nv, constrained=con, rho=rho, rho_min=1e-8 if simp else 0.0,
                         penalty=4.0 if simp else 1.0)
        X = O.state(prob.n_dof, 1)
        npxy = (nx + 1) * (ny + 1)
        z0 = int(rng.integers(0, nz + 1)); z1 = int(rng.integers(z0 + 1, nz + 2))
        owned = None
        if trial % 2:
            owned = np.zeros(prob.n_nodes, np.uint8); owned[z0 * npxy:z1 * npxy] = 1
        ctx = mast_b200.Context(0)
        ctx.set_mesh(8, prob.xyz, prob.conn)
        ctx.set_dof_map(nv, prob.constrained, owned)
        ctx.set_material(**prob.mat)
        if simp:
            ctx.set_density(prob.rho)
        ctx.build_pattern()
        R = np.full(prob.n_dof, np.nan); J = np.full(ctx.nnz, np.nan)
        ctx.residual_and_jacobian(X, R, J)
        assert _is_struct(ctx, struct_kernel, nv), (nx, ny, nz, nv)
        Ro, Jo = prob.residual_and_jacobian(X)
        rows = ctx.get_row_dofs()
        orp, oci = prob.pattern()
        rp, ci = ctx.get_pattern()
        assert np.array_equal(np.diff(rp), np.diff(orp)[rows])
        assert np.array_equal(ci, np.concatenate([oci[orp[r]:orp[r + 1]] for r in rows]))
        Jo_rows = np.concatenate([Jo[orp[r]:orp[r + 1]] for r in rows])
        _close(J, Jo_rows); _close(R[rows], Ro[rows])
        ghost = np.ones(prob.n_dof, bool); ghost[rows] = False
        assert np.all(R[ghost] == 0.0)


def test_struct_green_lagrange_tangent_in_the_slot_kernel(built, monkeypatch, struct_kernel):
    # NONLINEAR_STRAIN on build_cube meshes (3-dof system): tangent from the slot kernel's Green-Lagrange variant, residual
    # from the matrix-free force kernels; random shapes, owned z-slabs, SIMP, scattered Dirichlet dofs; against the oracle and
    # against the row-owner gather (MASTB200_GL_STRUCT=0)
    if struct_kernel != "slots":
        pytest.skip("one run is enough: the Green-Lagrange path does not depend on the linear kernel's choice")
    import mast_b200
    rng = np.random.default_rng(77)
    for trial, shape in enumerate([(1, 1, 1), (2, 3, 1), (8, 8, 8), (9, 15, 5), (13, 6, 17), (7, 9, 30)]):
        nx, ny, nz = shape
        simp = bool(trial % 2)
        xyz, conn = mesh.build_cube(nx, ny, nz, 0.3 + 0.05 * nx, 0.2 + 0.05 * ny, 0.1 + 0.05 * nz, 8, jitter=0.2)
        con = (rng.uniform(size=len(xyz) * 3) < 0.08).astype(np.uint8)
        rho = 0.2 + 0.6 * rng.uniform(size=len(xyz)) if simp else None
        se, sid = mesh.boundary_sides(nx, ny, nz, 3)
        prob = O.Problem(O.HEX8, xyz, conn, 3, constrained=con, rho=rho, rho_min=1e-8 if simp else 0.0,
                         penalty=4.0 if simp else 1.0, nonlinear_strain=1, side_elem=se, side_id=sid,
                         side_pressure=np.full(len(se), 5e7))
        X = O.state(prob.n_dof, 1, scale=4e-3)
        npxy = (nx + 1) * (ny + 1)
        owned = None
        if trial >= 3:
            z0 = int(rng.integers(0, nz + 1)); z1 = int(rng.integers(z0 + 1, nz + 2))
            owned = np.zeros(prob.n_nodes, np.uint8); owned[z0 * npxy:z1 * npxy] = 1
        def run():
            ctx = mast_b200.Context(0)
            ctx.set_mesh(8, prob.xyz, prob.conn)
            ctx.set_dof_map(3, prob.constrained, owned)
            ctx.set_material(**prob.mat)
            if simp:
                ctx.set_density(prob.rho)
            if owned is None:
                ctx.set_side_pressure(prob.side_elem, prob.side_id, prob.side_pressure)
            ctx.build_pattern()
            R = np.full(prob.n_dof, np.nan); J = np.full(ctx.nnz, np.nan)
            ctx.residual_and_jacobian(X, R, J)
            return ctx, R, J
        ctx, R, J = run()
        assert ctx.stats()["structured"] == 3, shape
        Ro, Jo = prob.residual_and_jacobian(X)
        if owned is not None:  # the context above has no pressure load: compare the internal part
            Ro, _ = O.Problem(O.HEX8, xyz, conn, 3, constrained=con, rho=rho, rho_min=1e-8 if simp else 0.0,
                              penalty=4.0 if simp else 1.0, nonlinear_strain=1).residual_and_jacobian(X, want_J=False)
        rows = ctx.get_row_dofs()
        orp, _ = prob.pattern()
        Jo_rows = np.concatenate([Jo[orp[r]:orp[r + 1]] for r in rows])
        _close(J, Jo_rows); _close(R[rows], Ro[rows])
        J2 = np.full(ctx.nnz, np.nan)
        ctx.residual_and_jacobian(X, None, J2)
        assert np.array_equal(J2, J)  # bitwise reproducible, tangent alone
        monkeypatch.setenv("MASTB200_GL_STRUCT", "0")
        ctx_g, Rg, Jg = run()
        monkeypatch.delenv("MASTB200_GL_STRUCT")
        assert ctx_g.stats()["structured"] == 0
        _close(J, Jg); _close(R[rows], Rg[rows])


def test_struct_thermoelastic_load(built, monkeypatch, struct_kernel):
    # TEMPERATURE volume load with linear strain (solid_element_3d.cpp:950-1089): the structured kernel serves R / J, the load's
    # force vector comes from the matrix-free force kernels; nodal and constant temperature, SIMP, nv = 3 / 6
    rng = np.random.default_rng(4)
    for nv, simp, const_T in ((3, True, False), (6, False, True), (3, False, False)):
        kw = dict(alpha=1.5e-5, thermal=1, ref_temperature=300.0)
        if const_T:
            kw["temperature_const"] = 345.0
        else:
            kw["temperature"] = 300 + 50 * rng.uniform(0, 1, 12 * 8 * 10)
        prob = O.cantilever(11, 7, 9, nv=nv, jitter=0.2, simp=simp, **kw)
        X = O.state(prob.n_dof, 1)
        ctx = prob.make_context(0)
        R = np.full(prob.n_dof, np.nan); J = np.full(ctx.nnz, np.nan)
        ctx.residual_and_jacobian(X, R, J)
        assert _is_struct(ctx, struct_kernel, nv)
        Ro, Jo = prob.residual_and_jacobian(X)
        _close(R, Ro); _close(J, Jo)
        R2 = np.full(prob.n_dof, np.nan)
        ctx.residual_and_jacobian(X, R2, None)
        assert np.array_equal(R2, R)


def test_struct_density_sensitivities_alone(built, monkeypatch, struct_kernel):
    # G without R / J on build_cube meshes: the structured sensitivity sweep (sweep_hex8_sens_struct.cu) against the oracle's
    # adjoint sensitivities and against the element-parallel kernels; compliance (adj == X) and a separate adjoint, scattered
    # Dirichlet dofs, owned z-slabs, shapes around the tile sizes
    if struct_kernel != "lockstep":
        pytest.skip("one run is enough: the sensitivity sweep does not depend on the R / J kernel's choice")
    import mast_b200
    rng = np.random.default_rng(31)
    for trial, shape in enumerate([(1, 1, 1), (3, 2, 2), (8, 8, 8), (9, 15, 5), (14, 7, 19), (7, 8, 30)]):
        nx, ny, nz = shape
        xyz, conn = mesh.build_cube(nx, ny, nz, 0.3 + 0.05 * nx, 0.2 + 0.05 * ny, 0.1 + 0.05 * nz, 8, jitter=0.2)
        con = (rng.uniform(size=len(xyz) * 3) < 0.08).astype(np.uint8)
        rho = 0.2 + 0.6 * rng.uniform(size=len(xyz))
        prob = O.Problem(O.HEX8, xyz, conn, 3, constrained=con, rho=rho, rho_min=1e-8, penalty=3.0 + (trial % 2))
        X = O.state(prob.n_dof, 1); lam = X if trial % 3 == 0 else O.state(prob.n_dof, 2)
        n = prob.n_nodes
        npxy = (nx + 1) * (ny + 1)
        owned = None
        if trial >= 3:
            z0 = int(rng.integers(0, nz + 1)); z1 = int(rng.integers(z0 + 1, nz + 2))
            owned = np.zeros(n, np.uint8); owned[z0 * npxy:z1 * npxy] = 1
        def run():
            ctx = mast_b200.Context(0)
            ctx.set_mesh(8, prob.xyz, prob.conn)
            ctx.set_dof_map(3, prob.constrained, owned)
            ctx.set_material(**prob.mat)
            ctx.set_density(prob.rho)
            ctx.build_pattern()
            G = np.full(n, np.nan)
            ctx.sensitivity_sweep(X, lam, G)
            return ctx, G
        ctx, G = run()
        assert ctx.stats()["kernel_launches"] == 1, shape  # one launch: the structured sweep
        g_ref = prob.adjoint_sensitivity(X, lam, np.arange(n + 1), np.arange(n), np.ones(n))
        own = np.ones(n, bool) if owned is None else owned.astype(bool)
        _close(G[own], g_ref[own], 1e-11)
        assert np.all(G[~own] == 0.0)
        _, G2 = run()
        assert np.array_equal(G, G2)  # bitwise reproducible
        monkeypatch.setenv("MASTB200_SENS_STRUCT", "0")
        ctx_e, Ge = run()
        monkeypatch.delenv("MASTB200_SENS_STRUCT")
        assert ctx_e.stats()["kernel_launches"] == 2
        _close(G[own], Ge[own], 1e-11)


@pytest.mark.parametrize("nv", [3, 6])
def test_struct_store_paths_give_the_same_bits(built, monkeypatch, struct_kernel, nv):
    # The CSR stores of the structured kernels take a wide path when the caller's J is 32-byte aligned (16-byte + 8-byte
    # stores of a row piece; TMA bulk stores of the rotational rows at nv = 6) and plain 8-byte stores otherwise; the
    # lock-step kernel clears its accumulators by TMA bulk copies unless MASTB200_STRUCT_NO_TMA_CLEAR is set.  Every
    # combination must give the same bits as the library-owned (aligned) buffer.
    import torch
    xyz, conn = mesh.build_cube(9, 8, 6, 0.3, 0.2, 0.25, 8, jitter=0.2)
    con = mesh.dirichlet_flags(len(xyz), nv, mesh.boundary_nodes(9, 8, 6, 4))
    prob = O.Problem(O.HEX8, xyz, conn, nv, constrained=con)
    X = O.state(prob.n_dof, 3)
    ctx = prob.make_context(0)
    R0 = np.empty(prob.n_dof); J0 = np.empty(ctx.nnz)
    ctx.residual_and_jacobian(X, R0, J0)
    assert _is_struct(ctx, struct_kernel, nv)
    dX = torch.from_numpy(X).cuda()
    for off in (0, 1, 2):  # doubles: 32-byte aligned, 8-byte aligned, 16-byte aligned
        buf = torch.full((ctx.nnz + 4,), float("nan"), dtype=torch.float64, device="cuda")
        dJ = buf[off:off + ctx.nnz]
        dR = torch.empty(prob.n_dof, dtype=torch.float64, device="cuda")
        ctx.residual_and_jacobian(dX, dR, dJ)
        torch.cuda.synchronize()
        assert np.array_equal(dJ.cpu().numpy(), J0) and np.array_equal(dR.cpu().numpy(), R0), off
        assert torch.isnan(buf[:off]).all() and torch.isnan(buf[off + ctx.nnz:]).all()  # nothing written outside J
    monkeypatch.setenv("MASTB200_STRUCT_NO_TMA_CLEAR", "1")
    ctx2 = prob.make_context(0)
    R1 = np.empty(prob.n_dof); J1 = np.empty(ctx2.nnz)
    ctx2.residual_and_jacobian(X, R1, J1)
    assert np.array_equal(J1, J0) and np.array_equal(R1, R0)


def test_struct_tall_slab_uneven_host_segments(built, monkeypatch, struct_kernel):
    # Host-buffer sweeps of a slab with >= 160 node planes run in UNEVEN z-segments (short ones at both ends, their copies
    # of X / R overlapped with the kernels); the result must equal the oracle's and, bit for bit, the one-launch sweep with
    # device buffers and the even split (MASTB200_SEG_UNIFORM=1).
    import torch
    if struct_kernel == "march":
        pytest.skip("the march serves device buffers only")
    prob = _problem(5, 4, 170)
    X = O.state(prob.n_dof, 2)
    ctx = prob.make_context(0)
    R = np.full(prob.n_dof, np.nan); J = np.full(ctx.nnz, np.nan)
    ctx.residual_and_jacobian(X, R, J)
    assert _is_struct(ctx, struct_kernel)
    Ro, Jo = prob.residual_and_jacobian(X)
    _close(R, Ro); _close(J, Jo)
    dX = torch.from_numpy(X).cuda()
    dR = torch.empty(prob.n_dof, dtype=torch.float64, device="cuda")
    dJ = torch.empty(ctx.nnz, dtype=torch.float64, device="cuda")
    ctx.residual_and_jacobian(dX, dR, dJ)
    torch.cuda.synchronize()
    assert np.array_equal(dR.cpu().numpy(), R) and np.array_equal(dJ.cpu().numpy(), J)
    monkeypatch.setenv("MASTB200_SEG_UNIFORM", "1")
    ctx2 = prob.make_context(0)
    R2 = np.full(prob.n_dof, np.nan); J2 = np.full(ctx2.nnz, np.nan)
    ctx2.residual_and_jacobian(X, R2, J2)
    assert np.array_equal(R2, R) and np.array_equal(J2, J)
