"""Parity of the CUDA path against the CPU oracle, through the C-ABI (ctypes), on a real B200.

Bar (BASELINE.json north_star): sparsity pattern bit-exact; values within 1e-12 relative, entry by entry, with the
floor 64 eps max|array| for entries that cancel (tests/submesh.py: close_entries); the summation order on the device is
fixed by the plan and differs from the oracle's element order, hence not bit-identical values.
"""
import numpy as np
import pytest

import oracle_lib as O
from submesh import close_entries
from mast_b200 import mesh

pytestmark = pytest.mark.gpu

RTOL = 1e-12


def _close(a, b, rtol=RTOL):
    """Entry-wise: |a - b| <= rtol |b| + 64 eps (rtol / 1e-12) max|b| -- relative 1e-12 for every entry, with the floor an
    entry that cancels to ~0 cannot beat (eps times the magnitude of the <= 64 terms it is the sum of)."""
    close_entries(a, b, rtol=rtol, terms=64 * rtol / 1e-12)


def _check_RJ(prob, X, want_R=True, want_J=True):
    ctx = prob.make_context(0)
    rp, ci = ctx.get_pattern()
    orp, oci = prob.pattern()
    assert np.array_equal(rp, orp) and np.array_equal(ci, oci)  # bit-exact pattern
    R = np.full(prob.n_dof, np.nan) if want_R else None
    J = np.full(ctx.nnz, np.nan) if want_J else None
    ctx.residual_and_jacobian(X, R, J)
    Ro, Jo = prob.residual_and_jacobian(X, want_R, want_J)
    if want_R:
        _close(R, Ro)
    if want_J:
        _close(J, Jo)
        # explicit zeros / exact values where the reference has them
        assert np.array_equal(J == 0.0, Jo == 0.0) or np.abs(J[(J == 0.0) != (Jo == 0.0)]).max() <= RTOL * np.abs(Jo).max()
    return ctx, R, J


@pytest.mark.parametrize("nv", [3, 6])
@pytest.mark.parametrize("jitter", [0.0, 0.2])
def test_c1_cantilever_residual_and_jacobian(built, nv, jitter):
    # BASELINE config C1: 20x10x10 HEX8, clamped x=0, pressure on y=Ly
    prob = O.cantilever(20, 10, 10, nv=nv, jitter=jitter)
    X = O.state(prob.n_dof, 1)
    ctx, R, J = _check_RJ(prob, X)
    assert np.isclose(ctx.res_l2_norm(), np.linalg.norm(R), rtol=1e-13)
    con = prob.constrained.astype(bool)
    assert np.all(R[con] == 0.0)


def test_jacobian_only_and_residual_only(built):
    prob = O.cantilever(7, 5, 3, jitter=0.2)
    X = O.state(prob.n_dof, 1)
    _check_RJ(prob, X, want_R=False)
    _check_RJ(prob, X, want_J=False)


def test_simp_density_and_thermal_load(built):
    prob = O.cantilever(8, 5, 4, jitter=0.2, simp=True, alpha=1.5e-5, thermal=1,
                        temperature=300 + 50 * np.random.default_rng(4).uniform(0, 1, 9 * 6 * 5),
                        ref_temperature=300.0)
    X = O.state(prob.n_dof, 1)
    _check_RJ(prob, X)
    prob2 = O.cantilever(5, 4, 4, alpha=1.5e-5, thermal=1, temperature_const=350.0, ref_temperature=300.0, nv=6)
    _check_RJ(prob2, O.state(prob2.n_dof, 1))


def test_unstructured_numbering(built):
    xyz, conn = mesh.build_cube(6, 5, 4, jitter=0.2)
    rng = np.random.default_rng(0)
    perm = rng.permutation(len(xyz))
    xyz2 = np.empty_like(xyz); xyz2[perm] = xyz
    conn2 = perm[conn].astype(np.int32)[rng.permutation(len(conn))]
    con = (rng.uniform(size=len(xyz) * 3) < 0.1).astype(np.uint8)  # scattered Dirichlet dofs
    prob = O.Problem(O.HEX8, xyz2, conn2, 3, constrained=con)
    _check_RJ(prob, O.state(prob.n_dof, 1))


def test_diagonal_elem_subdomain(built):
    # AssemblyBase::diagonal_elem_subdomain_id elements add 1e-24*I instead of physics
    xyz, conn = mesh.build_cube(5, 4, 3)
    diag = np.zeros(len(conn), np.uint8); diag[[0, 7, 8, 30]] = 1
    con = mesh.dirichlet_flags(len(xyz), 3, mesh.boundary_nodes(5, 4, 3, 4))
    prob = O.Problem(O.HEX8, xyz, conn, 3, constrained=con, diag_elem=diag)
    ctx, R, J = _check_RJ(prob, O.state(prob.n_dof, 1))


def test_device_pointers_and_determinism(built):
    import torch
    prob = O.cantilever(12, 9, 7, jitter=0.2, simp=True)
    X = O.state(prob.n_dof, 1)
    ctx = prob.make_context(0)
    dX = torch.from_numpy(X).cuda()
    dR = torch.empty(prob.n_dof, dtype=torch.float64, device="cuda")
    dJ = torch.empty(ctx.nnz, dtype=torch.float64, device="cuda")
    ctx.residual_and_jacobian(dX, dR, dJ)
    torch.cuda.synchronize()
    R1, J1 = dR.cpu().numpy().copy(), dJ.cpu().numpy().copy()
    Ro, Jo = prob.residual_and_jacobian(X)
    _close(R1, Ro); _close(J1, Jo)
    for _ in range(3):  # bitwise reproducible: no atomics, plan-fixed summation order
        dR.zero_(); dJ.zero_()
        ctx.residual_and_jacobian(dX, dR, dJ)
        torch.cuda.synchronize()
        assert np.array_equal(dR.cpu().numpy(), R1) and np.array_equal(dJ.cpu().numpy(), J1)
    # host-buffer path gives the same bits
    R2 = np.empty(prob.n_dof); J2 = np.empty(ctx.nnz)
    ctx.residual_and_jacobian(X, R2, J2)
    assert np.array_equal(R2, R1) and np.array_equal(J2, J1)


def test_stream_ordered_mode_gives_the_same_bits(built):
    # mastb200_set_async: with device buffers the call only enqueues on the context stream; same values, and host buffers
    # still come back complete
    import torch
    prob = O.cantilever(10, 7, 6, jitter=0.2)
    X = O.state(prob.n_dof, 1)
    ctx = prob.make_context(0)
    R0 = np.empty(prob.n_dof); J0 = np.empty(ctx.nnz)
    ctx.residual_and_jacobian(X, R0, J0)
    ctx.set_async(True)
    dX = torch.from_numpy(X).cuda()
    dR = torch.zeros(prob.n_dof, dtype=torch.float64, device="cuda")
    dJ = torch.zeros(ctx.nnz, dtype=torch.float64, device="cuda")
    stream = torch.cuda.ExternalStream(ctx.stream)
    for _ in range(3):
        ctx.residual_and_jacobian(dX, dR, dJ)
    stream.synchronize()
    assert np.array_equal(dR.cpu().numpy(), R0) and np.array_equal(dJ.cpu().numpy(), J0)
    assert ctx.stats()["last_main_kernel_ms"] > 0
    R1 = np.full(prob.n_dof, np.nan); J1 = np.full(ctx.nnz, np.nan)
    ctx.residual_and_jacobian(X, R1, J1)   # host buffers: synchronous as before
    assert np.array_equal(R1, R0) and np.array_equal(J1, J0)
    ctx.set_async(False)


def test_density_sensitivities(built):
    # one fused sweep == n_dv sensitivity_assemble sweeps + dots of the reference
    prob = O.cantilever(6, 4, 4, jitter=0.2, simp=True)
    X = O.state(prob.n_dof, 1); lam = O.state(prob.n_dof, 2)
    ctx = prob.make_context(0)
    n = prob.n_nodes
    G = np.full(n, np.nan); R = np.empty(prob.n_dof); J = np.empty(ctx.nnz)
    ctx.sensitivity_sweep(X, lam, G, R, J)
    g_ref = prob.adjoint_sensitivity(X, lam, np.arange(n + 1), np.arange(n), np.ones(n))
    _close(G, g_ref, 1e-11)
    Ro, Jo = prob.residual_and_jacobian(X)
    _close(R, Ro); _close(J, Jo)
    G2 = np.empty(n)
    ctx.sensitivity_sweep(X, lam, G2)  # sensitivities alone: the element-parallel kernels (another summation order)
    _close(G2, g_ref, 1e-11)
    # filter-like DV map: each DV spreads over a few nodes
    rng = np.random.default_rng(5)
    n_dv = 40
    cnt = rng.integers(1, 9, n_dv)
    dv_ptr = np.concatenate([[0], np.cumsum(cnt)])
    dv_node = rng.integers(0, n, dv_ptr[-1]).astype(np.int32)
    dv_w = rng.uniform(0.1, 1.0, dv_ptr[-1])
    sens = ctx.apply_dv_map_transpose(G, dv_ptr, dv_node, dv_w)
    s_ref = prob.adjoint_sensitivity(X, lam, dv_ptr, dv_node, dv_w)
    _close(sens, s_ref, 1e-11)


def test_compliance(built):
    prob = O.cantilever(9, 5, 4, jitter=0.2)
    X = O.state(prob.n_dof, 1)
    ctx = prob.make_context(0)
    assert np.isclose(ctx.compliance(X), prob.compliance(X), rtol=1e-12)
    dq = np.full(prob.n_dof, np.nan)
    ctx.compliance_derivative(X, dq)
    _close(dq, prob.compliance_derivative(X))


def test_owned_rows_with_ghosts(built):
    # rank-local assembly: rows of owned nodes only, ghost nodes contribute values of X
    import mast_b200
    prob = O.cantilever(8, 6, 6, jitter=0.2)
    X = O.state(prob.n_dof, 1)
    owned = (prob.xyz[:, 2] < 0.08).astype(np.uint8)
    ctx = mast_b200.Context(0)
    ctx.set_mesh(8, prob.xyz, prob.conn); ctx.set_dof_map(3, prob.constrained, owned)
    ctx.set_material(**prob.mat)
    ctx.set_side_pressure(prob.side_elem, prob.side_id, prob.side_pressure)
    n_rows, nnz = ctx.build_pattern()
    R = np.full(prob.n_dof, np.nan); J = np.full(nnz, np.nan)
    ctx.residual_and_jacobian(X, R, J)
    Ro, Jo = prob.residual_and_jacobian(X)
    orp, _ = prob.pattern()
    rows = ctx.get_row_dofs()
    rp, _ = ctx.get_pattern()
    Jref = np.concatenate([Jo[orp[r]:orp[r + 1]] for r in rows])
    _close(J, Jref)
    own_dofs = np.repeat(owned.astype(bool), 3)
    _close(R[own_dofs], Ro[own_dofs])
    # ghost rows are zeroed, also where a ghost node carries a surface load (mast_b200.h)
    assert np.all(R[~own_dofs] == 0.0)


def test_large_mesh_properties(built):
    # size-independent properties at a size the oracle would take minutes for (64^3, jittered)
    import torch
    n = 64
    xyz, conn = mesh.build_cube(n, n, n, jitter=0.2)
    prob = O.Problem(O.HEX8, xyz, conn, 3)
    ctx = prob.make_context(0)
    rp, ci = ctx.get_pattern()
    assert ctx.nnz == (3 * n + 1) ** 3 * 9
    # rigid translation + rotation produce zero residual; J t = 0 for rigid modes (linearity)
    t = np.array([0.3, -0.2, 0.5])
    w = np.array([0.01, 0.02, -0.015])
    Xr = (t + np.cross(w, xyz)).ravel()  # infinitesimal rigid motion: zero linear strain
    dJ = torch.empty(ctx.nnz, dtype=torch.float64, device="cuda")
    dR = torch.empty(prob.n_dof, dtype=torch.float64, device="cuda")
    ctx.residual_and_jacobian(torch.from_numpy(Xr).cuda(), dR, dJ)
    J = dJ.cpu().numpy(); R = dR.cpu().numpy()
    scale = np.abs(J).max() * np.abs(Xr).max()
    assert np.abs(R).max() <= 1e-11 * scale
    import scipy.sparse as sp
    A = sp.csr_matrix((J, ci, rp), shape=(prob.n_dof, prob.n_dof))
    assert abs(A - A.T).max() <= 1e-12 * np.abs(J).max()  # symmetric
    assert np.abs(A @ Xr).max() <= 1e-11 * scale           # rigid modes in the null space
    X = O.state(prob.n_dof, 1)
    ctx.residual_and_jacobian(torch.from_numpy(X).cuda(), dR, None)
    _close(dR.cpu().numpy(), A @ X, 1e-11)                 # linear elasticity: R(X) = J X without loads
    # spot-check a slab of elements against the oracle through element matrices
    sub = O.Problem(O.HEX8, xyz, conn[:n * n], 3)  # first element layer, same node numbering
    Ro, Jo = sub.residual_and_jacobian(X, want_R=False)
    orp, oci = sub.pattern()
    interior = [3 * (i + j * (n + 1)) for i in range(1, 6) for j in range(1, 6)]  # nodes on z=0: only layer 0 touches them
    for r in interior:
        assert np.array_equal(ci[rp[r]:rp[r + 1]], oci[orp[r]:orp[r + 1]])
        _close(J[rp[r]:rp[r + 1]], Jo[orp[r]:orp[r + 1]])


# ---- QUAD4 heat conduction (BASELINE config C2) ---------------------------------------------------------------------
def _heat_problem(nx, ny, k_of_T=1, jitter=0.0, diag=False, source=0.0):
    xyz, conn = mesh.build_square(nx, ny, 1.0, 1.0)
    if jitter:
        rng = np.random.Generator(np.random.MT19937(12345))
        h = np.array([1.0 / nx, 1.0 / ny, 0.0])
        interior = (xyz[:, 0] > 0) & (xyz[:, 0] < 1) & (xyz[:, 1] > 0) & (xyz[:, 1] < 1)
        xyz[interior] += (jitter * h * rng.uniform(-1, 1, xyz.shape))[interior]
    con = np.isclose(xyz[:, 0], 0.0).astype(np.uint8)  # fixed temperature on the edge x = 0
    de = None
    if diag:
        de = np.zeros(len(conn), np.uint8); de[[1, nx + 2, len(conn) - 1]] = 1
    return O.Problem(O.QUAD4, xyz, conn, 1, constrained=con, k0=237.0, k_beta=1e-3, k_of_T=k_of_T, thickness=0.06,
                     heat_source=source, diag_elem=de)


def _temperature(n):
    return 300 + 50 * np.random.Generator(np.random.MT19937(4)).uniform(0, 1, n)


@pytest.mark.parametrize("k_of_T", [0, 1])
def test_quad4_heat_residual_and_jacobian(built, k_of_T):
    prob = _heat_problem(37, 23, k_of_T=k_of_T, jitter=0.2, source=1.0e4)
    T = _temperature(prob.n_dof)
    ctx, R, J = _check_RJ(prob, T)
    assert np.isclose(ctx.res_l2_norm(), np.linalg.norm(R), rtol=1e-13)
    _check_RJ(prob, T, want_R=False)
    _check_RJ(prob, T, want_J=False)


@pytest.mark.parametrize("shape", [(1, 1), (2, 5), (30, 6), (31, 7), (32, 8), (63, 15), (70, 3)])
def test_quad4_heat_implicit_connectivity_kernel(built, shape, monkeypatch):
    # build_square numbering is recognised (every node owned, no diagonal elements) and takes quad4_heat_struct_kernel:
    # parity with the oracle, and agreement with the gather kernel on the same inputs
    prob = _heat_problem(*shape, k_of_T=1, jitter=0.2 if min(shape) > 1 else 0.0, source=2.0e3)
    T = _temperature(prob.n_dof)
    ctx, R, J = _check_RJ(prob, T)
    assert ctx.stats()["structured"] == 1
    monkeypatch.setenv("MASTB200_NO_STRUCT", "1")
    ctx2, R2, J2 = _check_RJ(prob, T)
    assert ctx2.stats()["structured"] == 0
    _close(R, R2); _close(J, J2)


def test_quad4_heat_diag_elems_and_unstructured_numbering(built):
    prob = _heat_problem(9, 7, diag=True)
    ctx, _, _ = _check_RJ(prob, _temperature(prob.n_dof))
    assert ctx.stats()["structured"] == 0
    rng = np.random.default_rng(0)
    xyz, conn = mesh.build_square(11, 8)
    perm = rng.permutation(len(xyz))
    xyz2 = np.empty_like(xyz); xyz2[perm] = xyz
    conn2 = perm[conn].astype(np.int32)[rng.permutation(len(conn))]
    con = (rng.uniform(size=len(xyz)) < 0.1).astype(np.uint8)
    p2 = O.Problem(O.QUAD4, xyz2, conn2, 1, constrained=con, k0=237.0, k_beta=1e-3, k_of_T=1, thickness=0.06)
    _check_RJ(p2, _temperature(p2.n_dof))


def test_quad4_heat_c2_full_size_properties(built):
    # C2: 1024^2 QUAD4, k(T); nnz = 3073^2; for constant k the matrix is symmetric and K.1 = 0 on free rows
    import scipy.sparse as sp
    n = 1024
    xyz, conn = mesh.build_square(n, n)
    import mast_b200
    ctx = mast_b200.Context(0)
    ctx.set_mesh(mast_b200.QUAD4, xyz, conn)
    ctx.set_dof_map(1)
    ctx.set_material(k0=237.0, k_beta=1e-3, k_of_T=0, thickness=1.0)
    ctx.build_pattern()
    assert ctx.nnz == 3073 ** 2
    T = _temperature(len(xyz))
    R = np.empty(len(xyz)); J = np.empty(ctx.nnz)
    ctx.residual_and_jacobian(T, R, J)
    rp, ci = ctx.get_pattern()
    A = sp.csr_matrix((J, ci, rp), shape=(len(xyz),) * 2)
    assert abs(A - A.T).max() <= 1e-12 * np.abs(J).max()
    assert np.abs(A @ np.ones(len(xyz))).max() <= 1e-11 * np.abs(J).max()
    assert np.abs(R - A @ T).max() <= 1e-11 * np.abs(R).max()  # linear conduction: R = K T
    ctx.set_material(k0=237.0, k_beta=1e-3, k_of_T=1, thickness=1.0)
    J2 = np.empty(ctx.nnz)
    ctx.residual_and_jacobian(T, None, J2)
    assert np.abs(J2 - J).max() > 1e-3 * np.abs(J).max()  # the dk/dT tangent is there


# ---- HEX27 quadratic elasticity (BASELINE config C5) ---------------------------------------------------------------
def _hex27_problem(nx, ny, nz, nv=3, jitter=0.0, **kw):
    xyz, conn = mesh.build_cube(nx, ny, nz, 0.3, 0.15, 0.15, 27, jitter=jitter)
    con = mesh.dirichlet_flags(len(xyz), nv, mesh.boundary_nodes(nx, ny, nz, 4, 27))
    return O.Problem(O.HEX27, xyz, conn, nv, constrained=con, **kw)


@pytest.mark.parametrize("nv", [3, 6])
def test_hex27_residual_and_jacobian(built, nv):
    prob = _hex27_problem(5, 3, 3, nv=nv, jitter=0.1)
    X = O.state(prob.n_dof, 1)
    ctx, R, J = _check_RJ(prob, X)
    assert np.isclose(ctx.res_l2_norm(), np.linalg.norm(R), rtol=1e-13)
    _check_RJ(prob, X, want_R=False)
    _check_RJ(prob, X, want_J=False)


def test_hex27_thermoelastic_simp_and_diag_elems(built):
    nx, ny, nz = 4, 3, 2
    n_nodes = (2 * nx + 1) * (2 * ny + 1) * (2 * nz + 1)
    rho = 0.2 + 0.6 * np.random.Generator(np.random.MT19937(3)).uniform(0, 1, n_nodes)
    diag = np.zeros(nx * ny * nz, np.uint8); diag[[0, 5, 23]] = 1
    prob = _hex27_problem(nx, ny, nz, jitter=0.1, alpha=1.5e-5, thermal=1, temperature=_temperature(n_nodes),
                          ref_temperature=300.0, rho=rho, rho_min=1e-8, penalty=4.0, diag_elem=diag)
    _check_RJ(prob, O.state(prob.n_dof, 1))
    prob2 = _hex27_problem(3, 2, 2, nv=6, alpha=1.5e-5, thermal=1, temperature_const=350.0, ref_temperature=300.0)
    _check_RJ(prob2, O.state(prob2.n_dof, 1))


def test_hex27_determinism_and_unstructured_numbering(built):
    xyz, conn = mesh.build_cube(4, 3, 3, elem_type=27, jitter=0.1)
    rng = np.random.default_rng(0)
    perm = rng.permutation(len(xyz))
    xyz2 = np.empty_like(xyz); xyz2[perm] = xyz
    conn2 = perm[conn].astype(np.int32)[rng.permutation(len(conn))]
    con = (rng.uniform(size=len(xyz) * 3) < 0.1).astype(np.uint8)
    prob = O.Problem(O.HEX27, xyz2, conn2, 3, constrained=con)
    X = O.state(prob.n_dof, 1)
    ctx, R, J = _check_RJ(prob, X)
    R2 = np.empty_like(R); J2 = np.empty_like(J)
    ctx.residual_and_jacobian(X, R2, J2)
    assert np.array_equal(R, R2) and np.array_equal(J, J2)  # colour order fixes the summation order


# ---- Green-Lagrange strain (HEX8 row-owner gather) ------------------------------------------------------------------
@pytest.mark.parametrize("nv", [3, 6])
def test_hex8_green_lagrange_residual_and_jacobian(built, nv):
    # StructuralElement3D::internal_residual with NONLINEAR_STRAIN (solid_element_3d.cpp:299-409)
    prob = O.cantilever(7, 5, 4, nv=nv, jitter=0.2, nonlinear_strain=1)
    X = O.state(prob.n_dof, 1, scale=2e-3)  # displacement gradients of order 0.1: the nonlinear terms matter
    ctx, R, J = _check_RJ(prob, X)
    # they do matter: the linear element gives a visibly different tangent
    lin = O.cantilever(7, 5, 4, nv=nv, jitter=0.2)
    _, Jl = lin.residual_and_jacobian(X, want_R=False)
    assert np.abs(J - Jl).max() > 1e-3 * np.abs(Jl).max()


def test_hex8_green_lagrange_thermal_simp_diag_unstructured(built):
    xyz, conn = mesh.build_cube(6, 4, 4, jitter=0.2)
    rng = np.random.default_rng(0)
    perm = rng.permutation(len(xyz))
    xyz2 = np.empty_like(xyz); xyz2[perm] = xyz
    conn2 = perm[conn].astype(np.int32)[rng.permutation(len(conn))]
    con = (rng.uniform(size=len(xyz) * 3) < 0.1).astype(np.uint8)
    diag = np.zeros(len(conn), np.uint8); diag[[3, 11]] = 1
    prob = O.Problem(O.HEX8, xyz2, conn2, 3, constrained=con, diag_elem=diag, nonlinear_strain=1,
                     rho=0.2 + 0.6 * rng.uniform(size=len(xyz)), rho_min=1e-8, penalty=4.0, alpha=1.5e-5, thermal=1,
                     temperature=300 + 50 * rng.uniform(size=len(xyz)), ref_temperature=300.0)
    _check_RJ(prob, O.state(prob.n_dof, 1, scale=2e-3))
    # linear strain through the same gather kernel is not reachable from the API (the streaming sweep serves it);
    # constant temperature + nv = 6
    prob2 = O.cantilever(5, 4, 3, nv=6, jitter=0.2, nonlinear_strain=1, alpha=1.5e-5, thermal=1, temperature_const=340.0,
                         ref_temperature=300.0)
    _check_RJ(prob2, O.state(prob2.n_dof, 1, scale=2e-3))


# ---- Green-Lagrange strain on HEX27 (row-owner gather, sweep_hex27_gather.cu) ---------------------------------------------
@pytest.mark.parametrize("nv", [3, 6])
def test_hex27_green_lagrange_residual_and_jacobian(built, nv):
    # StructuralElement3D::internal_residual with NONLINEAR_STRAIN on the tri-quadratic element (solid_element_3d.cpp:299-409)
    prob = _hex27_problem(4, 3, 2, nv=nv, jitter=0.1, nonlinear_strain=1)
    X = O.state(prob.n_dof, 1, scale=2e-3)
    ctx, R, J = _check_RJ(prob, X)
    lin = _hex27_problem(4, 3, 2, nv=nv, jitter=0.1)
    _, Jl = lin.residual_and_jacobian(X, want_R=False)
    assert np.abs(J - Jl).max() > 1e-3 * np.abs(Jl).max()  # the nonlinear terms matter at this state
    _check_RJ(prob, X, want_R=False)
    _check_RJ(prob, X, want_J=False)


def test_hex27_green_lagrange_thermal_simp_diag(built):
    # thermoelastic load with its geometric term (solid_element_3d.cpp:1039-1084), SIMP modulus, diagonal elements
    nx, ny, nz = 3, 3, 2
    n_nodes = (2 * nx + 1) * (2 * ny + 1) * (2 * nz + 1)
    rng = np.random.default_rng(5)
    diag = np.zeros(nx * ny * nz, np.uint8); diag[[2, 7]] = 1
    prob = _hex27_problem(nx, ny, nz, jitter=0.1, nonlinear_strain=1, alpha=1.5e-5, thermal=1,
                          temperature=_temperature(n_nodes), ref_temperature=300.0,
                          rho=0.2 + 0.6 * rng.uniform(size=n_nodes), rho_min=1e-8, penalty=4.0, diag_elem=diag)
    _check_RJ(prob, O.state(prob.n_dof, 1, scale=2e-3))


def test_green_lagrange_density_sensitivities_hex8_hex27(built):
    # internal_residual_sensitivity with NONLINEAR_STRAIN (solid_element_3d.cpp:566-788): adj . dR/d(rho_n) with the
    # Green-Lagrange strain of the state and its variation along the adjoint, alone and together with R / J
    rng = np.random.default_rng(9)
    for nv in (3, 6):
        prob = O.cantilever(6, 4, 3, nv=nv, jitter=0.2, simp=True, nonlinear_strain=1)
        X = O.state(prob.n_dof, 1, scale=2e-3); lam = O.state(prob.n_dof, 2, scale=2e-3)
        n = prob.n_nodes
        ctx = prob.make_context(0)
        g_ref = prob.adjoint_sensitivity(X, lam, np.arange(n + 1), np.arange(n), np.ones(n))
        g_lin = O.cantilever(6, 4, 3, nv=nv, jitter=0.2, simp=True).adjoint_sensitivity(X, lam, np.arange(n + 1), np.arange(n), np.ones(n))
        assert np.abs(g_ref - g_lin).max() > 1e-3 * np.abs(g_lin).max()  # the nonlinear terms matter at this state
        G = np.full(n, np.nan); R = np.empty(prob.n_dof); J = np.empty(ctx.nnz)
        ctx.sensitivity_sweep(X, lam, G, R, J)
        _close(G, g_ref, 1e-11)
        Ro, Jo = prob.residual_and_jacobian(X)
        _close(R, Ro); _close(J, Jo)
        G2 = np.full(n, np.nan)
        ctx.sensitivity_sweep(X, lam, G2)
        assert np.array_equal(G, G2)
    p27 = _hex27_problem(3, 2, 2, jitter=0.1, nonlinear_strain=1, rho=0.2 + 0.6 * rng.uniform(size=7 * 5 * 5), rho_min=1e-8,
                         penalty=4.0)
    X = O.state(p27.n_dof, 1, scale=2e-3); lam = O.state(p27.n_dof, 2, scale=2e-3)
    n = p27.n_nodes
    c27 = p27.make_context(0)
    g_ref = p27.adjoint_sensitivity(X, lam, np.arange(n + 1), np.arange(n), np.ones(n))
    G = np.full(n, np.nan); R = np.empty(p27.n_dof); J = np.empty(c27.nnz)
    c27.sensitivity_sweep(X, lam, G, R, J)
    _close(G, g_ref, 1e-11)
    Ro, Jo = p27.residual_and_jacobian(X)
    _close(R, Ro); _close(J, Jo)


# ---- stress output (StressStrainOutputBase) ------------------------------------------------------------------------
def _stress_case(prob, X, p_norm=2.0, rho_exp=0.0, sigma0=None):
    ctx = prob.make_context(0)
    st_o, en_o, vm_o, jw_o = prob.stress(X)
    n = len(vm_o)
    st = np.empty((n, 6)); en = np.empty((n, 6)); vm = np.empty(n); jw = np.empty(n)
    ctx.stress(X, st, en, vm, jw)
    _close(st, st_o); _close(en, en_o); _close(vm, vm_o); _close(jw, jw_o)
    if sigma0 is None:
        sigma0 = float(np.median(vm_o))
    q_o, dq_o = prob.stress_functional(X, p_norm, rho_exp, sigma0, want_derivative=True)
    q, sums = ctx.stress_functional(X, p_norm, rho_exp, sigma0)
    assert abs(q - q_o) <= 1e-12 * abs(q_o)
    dq = np.full(prob.n_dof, np.nan)
    ctx.stress_functional_derivative(X, p_norm, rho_exp, sigma0, sums, dq)
    _close(dq, dq_o, 1e-11)
    return ctx, sigma0, sums


def test_hex8_stress_output_and_aggregate(built):
    prob = O.cantilever(8, 5, 4, jitter=0.2, simp=True)
    X = O.state(prob.n_dof, 1)
    ctx, sigma0, sums = _stress_case(prob, X, p_norm=2.0, rho_exp=0.0)
    _stress_case(prob, X, p_norm=8.0, rho_exp=2.0)
    # participating elements: the sums of two complementary masks add up to the whole
    m = (np.arange(prob.n_elem) % 3 == 0).astype(np.uint8)
    _, s1 = ctx.stress_functional(X, 2.0, 0.0, sigma0, elem_mask=m)
    _, s2 = ctx.stress_functional(X, 2.0, 0.0, sigma0, elem_mask=1 - m)
    assert np.allclose(s1 + s2, sums, rtol=1e-13)
    # nv = 6 and Green-Lagrange strain
    prob6 = O.cantilever(5, 4, 3, nv=6, jitter=0.2, nonlinear_strain=1)
    _stress_case(prob6, O.state(prob6.n_dof, 1, scale=2e-3), p_norm=4.0, rho_exp=1.0)


def test_hex27_stress_output_and_aggregate(built):
    # BASELINE config C5's output: stress aggregate + its adjoint right-hand side on HEX27
    prob = _hex27_problem(3, 2, 2, jitter=0.1)
    _stress_case(prob, O.state(prob.n_dof, 1), p_norm=6.0, rho_exp=1.0)


# ---- QUAD4 Mindlin plate (StructuralElement2D) -----------------------------------------------------------------------
def _plate_problem(nx, ny, jitter=0.0, offset=0.03, plane_strain=0, diag=False, scramble=False, **kw):
    xyz, conn = mesh.build_square(nx, ny, 1.5, 1.0)
    rng = np.random.default_rng(11)
    if jitter:
        inner = (xyz[:, 0] > 0) & (xyz[:, 0] < 1.5) & (xyz[:, 1] > 0) & (xyz[:, 1] < 1.0)
        xyz[inner, :2] += jitter * np.array([1.5 / nx, 1.0 / ny]) * rng.uniform(-1, 1, (inner.sum(), 2))
    con = mesh.dirichlet_flags(len(xyz), 6, np.flatnonzero(xyz[:, 0] == 0.0))
    con2 = con.reshape(-1, 6); con2[np.flatnonzero(xyz[:, 1] == 1.0)[::2], 2] = 1  # a few w-only supports
    con = con2.ravel()
    d = None
    if diag:
        d = np.zeros(len(conn), np.uint8); d[[1, 5]] = 1
    if scramble:
        perm = rng.permutation(len(xyz))
        xyz2 = np.empty_like(xyz); xyz2[perm] = xyz
        c6 = np.empty_like(con2); c6[perm] = con2
        xyz, con = xyz2, c6.ravel()
        conn = perm[conn].astype(np.int32)[rng.permutation(len(conn))]
    return O.Problem(O.QUAD4, xyz, conn, 6, constrained=con, E0=72e9, nu=0.33, thickness=0.06, offset=offset,
                     kappa=5.0 / 6.0, plane_strain=plane_strain, diag_elem=d, **kw)


@pytest.mark.parametrize("offset,plane_strain", [(0.0, 0), (0.03, 0), (0.03, 1)])
def test_quad4_mindlin_plate_residual_and_jacobian(built, offset, plane_strain):
    prob = _plate_problem(9, 6, jitter=0.2, offset=offset, plane_strain=plane_strain)
    X = O.state(prob.n_dof, 1)
    ctx, R, J = _check_RJ(prob, X)
    assert np.isclose(ctx.res_l2_norm(), np.linalg.norm(R), rtol=1e-13)
    _check_RJ(prob, X, want_R=False)
    _check_RJ(prob, X, want_J=False)


def test_quad4_mindlin_plate_diag_elems_unstructured_and_size(built):
    prob = _plate_problem(7, 5, jitter=0.2, diag=True, scramble=True)
    _check_RJ(prob, O.state(prob.n_dof, 1))
    # a larger mesh: symmetry and the rigid-body null space of the assembled operator (no constraints)
    xyz, conn = mesh.build_square(160, 120, 1.5, 1.0)
    prob = O.Problem(O.QUAD4, xyz, conn, 6, E0=72e9, nu=0.33, thickness=0.06, offset=0.0, kappa=5.0 / 6.0)
    ctx = prob.make_context(0)
    rp, ci = ctx.get_pattern()
    J = np.empty(ctx.nnz); R = np.empty(prob.n_dof)
    U = np.zeros((len(xyz), 6)); U[:, 2] = 1.0 + 0.3 * xyz[:, 0] - 0.2 * xyz[:, 1]; U[:, 3] = -0.2; U[:, 4] = -0.3
    ctx.residual_and_jacobian(U.ravel(), R, J)           # rigid rotation + translation: no internal force
    import scipy.sparse as sp
    A = sp.csr_matrix((J, ci, rp), shape=(prob.n_dof, prob.n_dof))
    assert abs(A - A.T).max() <= 1e-12 * np.abs(J).max()
    assert np.abs(R).max() <= 1e-9 * np.abs(J).max()


def test_quad4_mindlin_plate_out_of_plane_and_folded(built):
    # elements out of the xy-plane: local frame per element, T K T^T (structural_element_base.cpp:1565-1674)
    prob = _plate_problem(8, 5, jitter=0.2)
    A = np.random.default_rng(2).normal(size=(3, 3))
    Q, _ = np.linalg.qr(A)
    rot = O.Problem(O.QUAD4, prob.xyz @ Q.T + 0.4, prob.conn, 6, constrained=prob.constrained, E0=72e9, nu=0.33,
                    thickness=0.06, offset=0.03, kappa=5.0 / 6.0)
    X = O.state(rot.n_dof, 1)
    _check_RJ(rot, X)
    _check_RJ(rot, X, want_R=False)
    _check_RJ(rot, X, want_J=False)
    # a folded plate: the left half stays in the plane z = 0.1 (no local frame), the right half climbs; diag elements too
    prob = _plate_problem(8, 5, jitter=0.2, diag=True)
    xyz = prob.xyz.copy()
    xyz[:, 2] = 0.1 + 0.6 * np.maximum(xyz[:, 0] - 0.75, 0.0)
    fold = O.Problem(O.QUAD4, xyz, prob.conn, 6, constrained=prob.constrained, E0=72e9, nu=0.33, thickness=0.06,
                     offset=0.0, kappa=5.0 / 6.0, diag_elem=prob.diag_elem)
    _check_RJ(fold, O.state(fold.n_dof, 2))


def _plate_state(prob, seed, amp_uv, amp_w, amp_th):
    rng = np.random.default_rng(seed)
    U = np.zeros((prob.n_nodes, 6))
    U[:, :2] = amp_uv * rng.uniform(-1, 1, (prob.n_nodes, 2))
    U[:, 2] = amp_w * rng.uniform(-1, 1, prob.n_nodes)
    U[:, 3:5] = amp_th * rng.uniform(-1, 1, (prob.n_nodes, 2))
    return U.ravel()


@pytest.mark.parametrize("offset", [0.0, 0.03])
def test_quad4_plate_von_karman_residual_and_jacobian(built, offset):
    # NONLINEAR_STRAIN on the plate: Green-Lagrange membrane strain + von Karman strain (structural_element_2d.cpp:57-100,
    # 141-250) with the tangent terms of :1886-1998, at a state where they matter (deflections of a few thicknesses)
    prob = _plate_problem(8, 5, jitter=0.2, offset=offset, nonlinear_strain=1)
    X = _plate_state(prob, 4, 0.01, 0.08, 0.05)
    ctx, R, J = _check_RJ(prob, X)
    lin = _plate_problem(8, 5, jitter=0.2, offset=offset)
    Rl, Jl = lin.residual_and_jacobian(X)
    assert np.abs(J - Jl).max() > 1e-2 * np.abs(Jl).max() and np.abs(R - Rl).max() > 1e-2 * np.abs(Rl).max()
    _check_RJ(prob, X, want_R=False)
    _check_RJ(prob, X, want_J=False)


def test_quad4_plate_von_karman_local_frames_diag_scrambled(built):
    # out of the xy-plane (local frames rotate the state in and the block out), diagonal elements, scrambled numbering
    prob = _plate_problem(7, 5, jitter=0.2, diag=True, scramble=True, nonlinear_strain=1)
    _check_RJ(prob, _plate_state(prob, 5, 0.01, 0.08, 0.05))
    flat = _plate_problem(6, 4, jitter=0.2, nonlinear_strain=1)
    Q, _ = np.linalg.qr(np.random.default_rng(2).normal(size=(3, 3)))
    rot = O.Problem(O.QUAD4, flat.xyz @ Q.T + 0.4, flat.conn, 6, constrained=flat.constrained, E0=72e9, nu=0.33,
                    thickness=0.06, offset=0.03, kappa=5.0 / 6.0, nonlinear_strain=1)
    _check_RJ(rot, _plate_state(rot, 6, 0.01, 0.05, 0.05))


def test_density_sensitivities_alone_hex8_hex27_and_old_sweep(built, monkeypatch):
    # G without R/J: element-parallel kernels for HEX8 and HEX27; compliance case adj == X; the streaming sweep
    # (MASTB200_SENS_SWEEP=1) gives the same numbers
    rng = np.random.default_rng(3)
    prob = O.cantilever(9, 6, 5, jitter=0.2, simp=True, nv=6)
    X = O.state(prob.n_dof, 1)
    n = prob.n_nodes
    ctx = prob.make_context(0)
    g_ref = prob.adjoint_sensitivity(X, X, np.arange(n + 1), np.arange(n), np.ones(n))
    G = np.full(n, np.nan)
    ctx.sensitivity_sweep(X, X, G)
    _close(G, g_ref, 1e-11)
    monkeypatch.setenv("MASTB200_SENS_SWEEP", "1")
    G_old = np.full(n, np.nan)
    ctx.sensitivity_sweep(X, X, G_old)
    _close(G_old, g_ref, 1e-11)
    monkeypatch.delenv("MASTB200_SENS_SWEEP")
    p27 = _hex27_problem(3, 2, 2, jitter=0.1, rho=0.2 + 0.6 * rng.uniform(size=7 * 5 * 5), rho_min=1e-8, penalty=4.0)
    X = O.state(p27.n_dof, 1); lam = O.state(p27.n_dof, 2)
    n = p27.n_nodes
    c27 = p27.make_context(0)
    g_ref = p27.adjoint_sensitivity(X, lam, np.arange(n + 1), np.arange(n), np.ones(n))
    G = np.full(n, np.nan); R = np.empty(p27.n_dof); J = np.empty(c27.nnz)
    c27.sensitivity_sweep(X, lam, G, R, J)   # with R and J of the same state
    _close(G, g_ref, 1e-11)
    Ro, Jo = p27.residual_and_jacobian(X)
    _close(R, Ro); _close(J, Jo)


def test_stress_aggregate_density_derivative_hex8_hex27(built):
    # dq/dp of the von Mises aggregate for density DVs at fixed state (the partial-sensitivity term of the adjoint method)
    rng = np.random.default_rng(9)
    for prob in (O.cantilever(7, 5, 4, jitter=0.2, simp=True),
                 _hex27_problem(3, 2, 2, jitter=0.1, rho=0.2 + 0.6 * rng.uniform(size=7 * 5 * 5), rho_min=1e-8, penalty=4.0)):
        X = O.state(prob.n_dof, 1)
        n = prob.n_nodes
        ctx = prob.make_context(0)
        st, en, vm, jw = prob.stress(X)
        sigma0 = float(np.median(vm))
        q, sums = ctx.stress_functional(X, 6.0, 1.5, sigma0)
        Gq = np.full(n, np.nan)
        ctx.stress_functional_density_derivative(X, 6.0, 1.5, sigma0, sums, Gq)
        n_dv = 25
        cnt = rng.integers(1, 6, n_dv)
        dv_ptr = np.concatenate([[0], np.cumsum(cnt)])
        dv_node = rng.integers(0, n, dv_ptr[-1]).astype(np.int32)
        dv_w = rng.uniform(0.1, 1.0, dv_ptr[-1])
        sens = ctx.apply_dv_map_transpose(Gq, dv_ptr, dv_node, dv_w)
        ref = prob.stress_functional_density_sensitivity(X, 6.0, 1.5, sigma0, dv_ptr, dv_node, dv_w)
        _close(sens, ref, 1e-11)


def test_hex27_surface_pressure_and_compliance(built):
    # SURFACE_PRESSURE on QUAD9 faces of HEX27 elements (3x3 Gauss), compliance value and derivative
    nx, ny, nz = 4, 3, 2
    xyz, conn = mesh.build_cube(nx, ny, nz, 0.3, 0.15, 0.15, 27, jitter=0.1)
    con = mesh.dirichlet_flags(len(xyz), 3, mesh.boundary_nodes(nx, ny, nz, 4, 27))
    se, sid = mesh.boundary_sides(nx, ny, nz, 3)
    se2, sid2 = mesh.boundary_sides(nx, ny, nz, 5)
    prob = O.Problem(O.HEX27, xyz, conn, 3, constrained=con, side_elem=np.concatenate([se, se2]),
                     side_id=np.concatenate([sid, sid2]),
                     side_pressure=np.concatenate([np.full(len(se), 5e7), np.full(len(se2), -2e7)]))
    X = O.state(prob.n_dof, 1)
    ctx, R, J = _check_RJ(prob, X)
    assert np.isclose(ctx.compliance(X), prob.compliance(X), rtol=1e-12)
    dq = np.full(prob.n_dof, np.nan)
    ctx.compliance_derivative(X, dq)
    _close(dq, prob.compliance_derivative(X))


def test_quad4_heat_flux_and_convection_side_loads(built):
    # HEAT_FLUX on the bottom edge (boundary 0 of build_square), CONVECTION_HEAT_FLUX on the right edge (boundary 1)
    nx, ny = 14, 9
    xyz, conn = mesh.build_square(nx, ny, 1.4, 0.9)
    rng = np.random.default_rng(2)
    inner = (xyz[:, 0] > 0) & (xyz[:, 0] < 1.4) & (xyz[:, 1] > 0) & (xyz[:, 1] < 0.9)
    xyz[inner, :2] += 0.02 * rng.uniform(-1, 1, (inner.sum(), 2))
    con = np.isclose(xyz[:, 0], 0.0).astype(np.uint8)
    bottom = np.arange(nx, dtype=np.int32)                       # elements j = 0, side 0
    right = (np.arange(ny, dtype=np.int32) * nx + nx - 1)        # elements i = nx-1, side 1
    se = np.concatenate([bottom, right, right[:3]])              # three elements carry both kinds on the same side
    sid = np.concatenate([np.zeros(nx, np.int32), np.ones(ny, np.int32), np.ones(3, np.int32)])
    flux = np.concatenate([np.full(nx, -2.5e3), np.zeros(ny), np.full(3, 4.0e2)])
    hc = np.concatenate([np.zeros(nx), np.full(ny, 35.0), np.zeros(3)])
    Tamb = np.concatenate([np.zeros(nx), np.full(ny, 290.0), np.zeros(3)])
    for k_of_T in (0, 1):
        prob = O.Problem(O.QUAD4, xyz, conn, 1, constrained=con, k0=237.0, k_beta=1e-3, k_of_T=k_of_T, thickness=0.06,
                         heat_source=1.0e4, side_elem=se, side_id=sid, side_pressure=flux, side_conv_coeff=hc, side_T_amb=Tamb)
        T = _temperature(prob.n_dof)
        ctx, R, J = _check_RJ(prob, T)
        plain = O.Problem(O.QUAD4, xyz, conn, 1, constrained=con, k0=237.0, k_beta=1e-3, k_of_T=k_of_T, thickness=0.06,
                          heat_source=1.0e4)
        R0, J0 = plain.residual_and_jacobian(T)
        assert np.abs(R - R0).max() > 1e-3 * np.abs(R0).max() and np.abs(J - J0).max() > 0  # the loads are there


def test_quad4_heat_surface_radiation_side_loads(built):
    # SURFACE_RADIATION_HEAT_FLUX (heat_conduction_elem_base.cpp:1404-1458) on the top edge (side 2), together with
    # convection on the right edge and a flux entry that shares a side with a radiation entry
    nx, ny = 12, 7
    xyz, conn = mesh.build_square(nx, ny, 1.2, 0.7)
    rng = np.random.default_rng(4)
    inner = (xyz[:, 0] > 0) & (xyz[:, 0] < 1.2) & (xyz[:, 1] > 0) & (xyz[:, 1] < 0.7)
    xyz[inner, :2] += 0.02 * rng.uniform(-1, 1, (inner.sum(), 2))
    con = np.isclose(xyz[:, 0], 0.0).astype(np.uint8)
    top = (np.arange(nx, dtype=np.int32) + (ny - 1) * nx)        # elements j = ny-1, side 2
    right = (np.arange(ny, dtype=np.int32) * nx + nx - 1)        # elements i = nx-1, side 1
    se = np.concatenate([top, right, top[:4]])
    sid = np.concatenate([np.full(nx, 2, np.int32), np.ones(ny, np.int32), np.full(4, 2, np.int32)])
    flux = np.concatenate([np.zeros(nx), np.zeros(ny), np.full(4, 3.0e2)])
    hc = np.concatenate([np.zeros(nx), np.full(ny, 35.0), np.zeros(4)])
    Tamb = np.concatenate([np.full(nx, 3.0), np.full(ny, 290.0), np.zeros(4)])
    eps = np.concatenate([0.6 + 0.3 * rng.uniform(size=nx), np.zeros(ny), np.zeros(4)])
    kw = dict(constrained=con, k0=237.0, k_beta=1e-3, k_of_T=1, thickness=0.06, heat_source=1.0e4)
    prob = O.Problem(O.QUAD4, xyz, conn, 1, side_elem=se, side_id=sid, side_pressure=flux, side_conv_coeff=hc,
                     side_T_amb=Tamb, side_emissivity=eps, rad_sb_const=5.670374419e-8, rad_T_ref_zero=273.15, **kw)
    T = _temperature(prob.n_dof)
    ctx, R, J = _check_RJ(prob, T)
    norad = O.Problem(O.QUAD4, xyz, conn, 1, side_elem=se, side_id=sid, side_pressure=flux, side_conv_coeff=hc,
                      side_T_amb=Tamb, **kw)
    R0, J0 = norad.residual_and_jacobian(T)
    assert np.abs(R - R0).max() > 1e-6 * np.abs(R0).max() and np.abs(J - J0).max() > 0  # the radiation is there


@pytest.mark.parametrize("k_of_T", [0, 1])
@pytest.mark.parametrize("struct", [True, False])
def test_quad4_heat_sensitivity_assemble_conductivity_and_thickness(built, k_of_T, struct, monkeypatch):
    # sensitivity_assemble of the heat-conduction system for the conductivity k0 and the section thickness
    # (heat_conduction_elem_base.cpp:552-617 + the load sensitivities), with a source and flux / convection / radiation sides
    if not struct:
        monkeypatch.setenv("MASTB200_NO_STRUCT", "1")
    nx, ny = 33, 17
    xyz, conn = mesh.build_square(nx, ny, 1.2, 0.7)
    rng = np.random.default_rng(4)
    inner = (xyz[:, 0] > 0) & (xyz[:, 0] < 1.2) & (xyz[:, 1] > 0) & (xyz[:, 1] < 0.7)
    xyz[inner, :2] += 0.01 * rng.uniform(-1, 1, (inner.sum(), 2))
    con = np.isclose(xyz[:, 0], 0.0).astype(np.uint8)
    top = (np.arange(nx, dtype=np.int32) + (ny - 1) * nx)
    right = (np.arange(ny, dtype=np.int32) * nx + nx - 1)
    se = np.concatenate([top, right, top[:4]])
    sid = np.concatenate([np.full(nx, 2, np.int32), np.ones(ny, np.int32), np.full(4, 2, np.int32)])
    flux = np.concatenate([np.zeros(nx), np.zeros(ny), np.full(4, 3.0e2)])
    hc = np.concatenate([np.zeros(nx), np.full(ny, 35.0), np.zeros(4)])
    Tamb = np.concatenate([np.full(nx, 3.0), np.full(ny, 290.0), np.zeros(4)])
    eps = np.concatenate([0.6 + 0.3 * rng.uniform(size=nx), np.zeros(ny), np.zeros(4)])
    prob = O.Problem(O.QUAD4, xyz, conn, 1, constrained=con, k0=237.0, k_beta=1e-3, k_of_T=k_of_T, thickness=0.06,
                     heat_source=1.0e4, side_elem=se, side_id=sid, side_pressure=flux, side_conv_coeff=hc, side_T_amb=Tamb,
                     side_emissivity=eps, rad_sb_const=5.670374419e-8, rad_T_ref_zero=273.15)
    T = _temperature(prob.n_dof)
    ctx = prob.make_context(0)
    for param, kind in ((3, 1), (4, 2)):   # MASTB200_PARAM_K0, MASTB200_PARAM_THICKNESS
        out = np.full(prob.n_dof, np.nan)
        ctx.sensitivity_assemble(T, out, param=param)
        _close(out, prob.sensitivity_assemble_param(T, kind))
    # the card is back in place afterwards
    R = np.full(prob.n_dof, np.nan)
    ctx.residual_and_jacobian(T, R, None)
    _close(R, prob.residual_and_jacobian(T)[0])
    with pytest.raises(Exception):
        ctx.sensitivity_assemble(T, out, param=1)


# ---- round 2: sensitivity_assemble as a vector, linearized product, point loads, thermoelastic compliance -----------
def _solid(elem_type, simp=True, jitter=0.2, loads=True, **kw):
    nx, ny, nz = (5, 4, 3) if elem_type == 8 else (3, 2, 2)
    xyz, conn = mesh.build_cube(nx, ny, nz, 0.3, 0.2, 0.15, elem_type, jitter=jitter)
    con = mesh.dirichlet_flags(len(xyz), 3, mesh.boundary_nodes(nx, ny, nz, 4, elem_type))
    rho = None
    if simp:
        rho = 0.2 + 0.6 * np.random.Generator(np.random.MT19937(3)).uniform(0.0, 1.0, len(xyz))
        kw.setdefault("rho_min", 1e-8); kw.setdefault("penalty", 3.0)
    se = sid = sp = None
    if loads:
        se, sid = mesh.boundary_sides(nx, ny, nz, 3)
        sp = np.full(len(se), 5e7)
    return O.Problem(elem_type, xyz, conn, 3, constrained=con, rho=rho, side_elem=se, side_id=sid, side_pressure=sp, **kw)


@pytest.mark.parametrize("elem_type", [8, 27])
def test_sensitivity_assemble_vector(built, elem_type):
    """NonlinearImplicitAssembly::sensitivity_assemble as the reference defines it: the global dR/dp vector of ONE
    parameter (a filtered density DV, the card's E0, the card's nu), against the oracle's element loop."""
    prob = _solid(elem_type)
    X = O.state(prob.n_dof, 1)
    ctx = prob.make_context(0)
    rng = np.random.default_rng(11)
    dv_node = rng.choice(prob.n_nodes, 9, replace=False).astype(np.int32)
    dv_w = rng.uniform(0.1, 1.0, 9)
    out = np.full(prob.n_dof, np.nan)
    ctx.sensitivity_assemble(X, out, 0, dv_node, dv_w)
    _close(out, prob.sensitivity_assemble(X, dv_node, dv_w), 1e-11)
    assert np.all(out[prob.constrained.astype(bool)] == 0.0)
    for kind in (1, 2):
        ctx.sensitivity_assemble(X, out, kind)
        _close(out, prob.sensitivity_assemble_param(X, kind), 1e-11)
    # adjoint identity of assembly_base.cpp:578: lambda . dR/dp equals the fused sweep's W^T G
    lam = O.state(prob.n_dof, 2)
    ctx.sensitivity_assemble(X, out, 0, dv_node, dv_w)
    G = np.empty(prob.n_nodes)
    ctx.sensitivity_sweep(X, lam, G)
    s = ctx.apply_dv_map_transpose(G, [0, 9], dv_node, dv_w)[0]
    lam_c = np.where(prob.constrained.astype(bool), 0.0, lam)
    assert abs(lam_c @ out - s) <= 1e-10 * abs(s)


@pytest.mark.parametrize("elem_type", [8, 27])
def test_sensitivity_assemble_vector_green_lagrange(built, elem_type):
    """The same three parameters with NONLINEAR_STRAIN: dR/dp = sum_q JxW B_a(F)^T (dD/dp E_GL), F = I + grad u."""
    prob = _solid(elem_type, nonlinear_strain=1)
    X = O.state(prob.n_dof, 1, scale=2e-3)
    ctx = prob.make_context(0)
    rng = np.random.default_rng(11)
    dv_node = rng.choice(prob.n_nodes, 9, replace=False).astype(np.int32)
    dv_w = rng.uniform(0.1, 1.0, 9)
    out = np.full(prob.n_dof, np.nan)
    ctx.sensitivity_assemble(X, out, 0, dv_node, dv_w)
    ref = prob.sensitivity_assemble(X, dv_node, dv_w)
    _close(out, ref, 1e-11)
    lin = _solid(elem_type).sensitivity_assemble(X, dv_node, dv_w)
    assert np.abs(ref - lin).max() > 1e-3 * np.abs(lin).max()
    for kind in (1, 2):
        ctx.sensitivity_assemble(X, out, kind)
        _close(out, prob.sensitivity_assemble_param(X, kind), 1e-11)


@pytest.mark.parametrize("elem_type", [8, 27])
def test_linearized_jacobian_solution_product(built, elem_type):
    prob = _solid(elem_type, loads=False)
    X = O.state(prob.n_dof, 1); dX = O.state(prob.n_dof, 7)
    ctx = prob.make_context(0)
    out = np.full(prob.n_dof, np.nan)
    ctx.linearized_jacobian_solution_product(X, dX, out)
    _close(out, prob.linearized_jacobian_solution_product(X, dX), 1e-11)


@pytest.mark.parametrize("elem_type", [8, 27])
def test_linearized_jacobian_solution_product_green_lagrange(built, elem_type):
    # K_T(X) dX with NONLINEAR_STRAIN: material tangent + initial-stress stiffness applied to a vector, matrix-free
    prob = _solid(elem_type, loads=False, nonlinear_strain=1)
    X = O.state(prob.n_dof, 1, scale=2e-3); dX = O.state(prob.n_dof, 7)
    ctx = prob.make_context(0)
    out = np.full(prob.n_dof, np.nan)
    ctx.linearized_jacobian_solution_product(X, dX, out)
    ref = prob.linearized_jacobian_solution_product(X, dX)
    _close(out, ref, 1e-11)
    lin = _solid(elem_type, loads=False).linearized_jacobian_solution_product(X, dX)
    assert np.abs(ref - lin).max() > 1e-3 * np.abs(lin).max()
    # and it is the assembled tangent times dX
    import scipy.sparse as sp
    J = np.empty(ctx.nnz)
    ctx.residual_and_jacobian(X, None, J)
    rp, ci = ctx.get_pattern()
    A = sp.csr_matrix((J, ci, rp), shape=(prob.n_dof, prob.n_dof))
    free = ~prob.constrained.astype(bool)
    dXc = np.where(free, dX, 0.0)   # the element matrices are constrained: columns of constrained dofs are zero
    outc = np.full(prob.n_dof, np.nan)
    ctx.linearized_jacobian_solution_product(X, dXc, outc)
    assert np.abs((A @ dXc)[free] - outc[free]).max() <= 1e-10 * np.abs(outc).max()


def test_point_loads(built):
    prob = _solid(8, simp=False)
    X = O.state(prob.n_dof, 1)
    ctx = prob.make_context(0)
    rng = np.random.default_rng(3)
    nodes = np.concatenate([rng.choice(prob.n_nodes, 12, replace=False), [5, 5, 0]]).astype(np.int32)  # repeats, a clamped node
    loads = rng.uniform(-1e5, 1e5, (len(nodes), 3))
    ctx.set_point_loads(nodes, loads)
    R = np.full(prob.n_dof, np.nan)
    ctx.residual_and_jacobian(X, R, None)
    Ro, _ = prob.residual_and_jacobian(X, want_J=False)
    prob.add_point_loads(Ro, nodes, loads)
    _close(R, Ro)
    ctx.set_point_loads(np.zeros(0, np.int32), np.zeros((0, 3)))
    ctx.residual_and_jacobian(X, R, None)
    _close(R, prob.residual_and_jacobian(X, want_J=False)[0])


@pytest.mark.parametrize("elem_type", [8, 27])
def test_compliance_with_thermoelastic_load(built, elem_type):
    """ComplianceOutput with the TEMPERATURE volume load in f_ext (C5's load case): value, dq/dX, direct sensitivity."""
    T = None
    prob = _solid(elem_type, loads=False, alpha=1.5e-5, thermal=1, temperature_const=350.0, ref_temperature=300.0)
    X = O.state(prob.n_dof, 1)
    ctx = prob.make_context(0)
    q = ctx.compliance(X)
    assert np.isclose(q, prob.compliance(X), rtol=1e-11)
    dq = np.full(prob.n_dof, np.nan)
    ctx.compliance_derivative(X, dq)
    _close(dq, prob.compliance_derivative(X), 1e-11)
    dXdp = O.state(prob.n_dof, 5)
    assert np.isclose(ctx.compliance_direct_sensitivity(X, dXdp), prob.compliance(dXdp), rtol=1e-11)
    # nodal temperature field + pressure: the value adds both parts
    rngT = np.random.Generator(np.random.MT19937(4))
    T = 300 + 50 * rngT.uniform(0, 1, prob.n_nodes)
    prob2 = _solid(elem_type, loads=True, alpha=1.5e-5, thermal=1, temperature=T, ref_temperature=300.0)
    ctx2 = prob2.make_context(0)
    assert np.isclose(ctx2.compliance(X), prob2.compliance(X), rtol=1e-11)


def test_owned_ranks_sum_to_the_serial_compliance(built):
    """Two 'ranks' (two contexts owning complementary node sets of one mesh, ghost elements carrying loads across the
    cut): ghost rows stay zero and the ranks' compliances add up to the serial value (each loaded node counted once)."""
    import mast_b200
    prob = O.cantilever(8, 6, 6, jitter=0.2)
    X = O.state(prob.n_dof, 1)
    q_ref = prob.compliance(X)
    total = 0.0
    for part in (0, 1):
        owned = ((prob.xyz[:, 2] < 0.08) == (part == 0)).astype(np.uint8)
        ctx = mast_b200.Context(0)
        ctx.set_mesh(8, prob.xyz, prob.conn); ctx.set_dof_map(3, prob.constrained, owned)
        ctx.set_material(**prob.mat)
        ctx.set_side_pressure(prob.side_elem, prob.side_id, prob.side_pressure)
        ctx.build_pattern()
        R = np.full(prob.n_dof, np.nan)
        ctx.residual_and_jacobian(X, R, None)
        assert np.all(R[~np.repeat(owned.astype(bool), 3)] == 0.0)
        total += ctx.compliance(X)
    assert np.isclose(total, q_ref, rtol=1e-12)
