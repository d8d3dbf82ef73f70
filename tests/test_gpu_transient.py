"""Transient and modal variants of the sweep: mass / capacitance matrices, the Newmark element combination
of TransientAssembly::residual_and_jacobian and the A / B matrices of the eigenproblem assembly, against the oracle."""
import numpy as np
import pytest

import oracle_lib as O
from mast_b200 import mesh
from submesh import close_entries

pytestmark = pytest.mark.gpu


def _solid(elem_type, nv=3):
    nx, ny, nz = (6, 4, 3) if elem_type == 8 else (3, 2, 2)
    xyz, conn = mesh.build_cube(nx, ny, nz, 0.3, 0.2, 0.15, elem_type, jitter=0.2)
    con = mesh.dirichlet_flags(len(xyz), nv, mesh.boundary_nodes(nx, ny, nz, 4, elem_type))
    con = np.maximum(con, (np.random.default_rng(0).uniform(size=len(con)) < 0.04).astype(np.uint8))
    se, sid = mesh.boundary_sides(nx, ny, nz, 3)
    return O.Problem(elem_type, xyz, conn, nv, constrained=con, side_elem=se, side_id=sid, side_pressure=np.full(len(se), 5e7))


@pytest.mark.parametrize("elem_type,nv", [(8, 3), (8, 6), (27, 3)])
@pytest.mark.parametrize("lumped", [False, True])
def test_mass_matrix_and_newmark_step(built, elem_type, nv, lumped):
    prob = _solid(elem_type, nv)
    ctx = prob.make_context(0)
    rho_m = 2700.0
    M = np.full(ctx.nnz, np.nan)
    ctx.mass_matrix(rho_m, M, lumped)
    Mo = prob.mass_matrix(rho_m, lumped)
    close_entries(M, Mo, rtol=1e-12, what="mass matrix")
    assert np.array_equal(M == 0.0, Mo == 0.0)
    # Newmark: x_ddot = (x - x0) / beta / dt^2 - ... (the solver's business); the assembly sees X, the acceleration and 1/(beta dt^2)
    X = O.state(prob.n_dof, 1); acc = O.state(prob.n_dof, 9, scale=50.0)
    if nv == 6:
        X.reshape(-1, 6)[:, 3:] = 0.0
    beta, dt = 0.25, 1e-4
    coeff = 1.0 / beta / dt / dt
    R = np.full(prob.n_dof, np.nan); J = np.full(ctx.nnz, np.nan)
    ctx.transient_residual_and_jacobian(X, acc, rho_m, coeff, R, J, lumped)
    Ro, Jo = prob.transient_residual_and_jacobian(X, acc, rho_m, coeff, lumped)
    close_entries(J, Jo, rtol=1e-12, terms=8 * prob.nn, what="transient Jacobian")
    close_entries(R, Ro, rtol=1e-12, terms=8 * prob.nn, what="transient residual")
    assert np.all(R[prob.constrained.astype(bool)] == 0.0)


def test_heat_capacitance_and_first_order_step(built):
    xyz, conn = mesh.build_square(19, 13, 1.0, 0.7)
    rng = np.random.Generator(np.random.MT19937(12345))
    interior = (xyz[:, 0] > 0) & (xyz[:, 0] < 1) & (xyz[:, 1] > 0) & (xyz[:, 1] < 0.7)
    xyz[interior] += (0.2 * np.array([1 / 19, 0.7 / 13, 0.0]) * rng.uniform(-1, 1, xyz.shape))[interior]
    con = np.isclose(xyz[:, 0], 0.0).astype(np.uint8)
    prob = O.Problem(O.QUAD4, xyz, conn, 1, constrained=con, k0=237.0, k_beta=1e-3, k_of_T=1, thickness=0.06, heat_source=1e4)
    ctx = prob.make_context(0)
    T = 300 + 50 * np.random.Generator(np.random.MT19937(4)).uniform(0, 1, prob.n_dof)
    Tdot = np.random.default_rng(1).uniform(-5, 5, prob.n_dof)
    rho_cp = 2700.0 * 900.0
    for lumped in (False, True):
        C = np.full(ctx.nnz, np.nan)
        ctx.mass_matrix(rho_cp, C, lumped)
        close_entries(C, prob.mass_matrix(rho_cp, lumped), rtol=1e-12, what="capacitance matrix")
        R = np.full(prob.n_dof, np.nan); J = np.full(ctx.nnz, np.nan)
        ctx.transient_residual_and_jacobian(T, Tdot, rho_cp, 1.0 / 0.5 / 1e-2, R, J, lumped)
        Ro, Jo = prob.transient_residual_and_jacobian(T, Tdot, rho_cp, 1.0 / 0.5 / 1e-2, lumped)
        close_entries(J, Jo, rtol=1e-12, what="first-order Jacobian"); close_entries(R, Ro, rtol=1e-12, what="first-order residual")


def test_eigenproblem_matrices_and_a_mode(built):
    """A = K, B = M in the device pattern; the lowest mode of the clamped block from scipy on the device matrices equals
    the one on the oracle's matrices (the modal analysis of src/base/eigenproblem_assembly.cpp)."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    import torch
    xyz, conn = mesh.build_cube(8, 3, 3, 0.4, 0.05, 0.05, 8)
    con = mesh.dirichlet_flags(len(xyz), 3, mesh.boundary_nodes(8, 3, 3, 4))
    prob = O.Problem(O.HEX8, xyz, conn, 3, constrained=con)
    ctx = prob.make_context(0)
    X0 = np.zeros(prob.n_dof)
    dA = torch.empty(ctx.nnz, dtype=torch.float64, device="cuda"); dB = torch.empty_like(dA)
    ctx.eigenproblem_assemble(X0, 2700.0, dA, dB)
    A, B = dA.cpu().numpy(), dB.cpu().numpy()
    _, Ko = prob.residual_and_jacobian(X0, want_R=False)
    Mo = prob.mass_matrix(2700.0)
    close_entries(A, Ko, what="A"); close_entries(B, Mo, what="B")
    rp, ci = ctx.get_pattern()
    free = ~prob.constrained.astype(bool)
    def lowest(Kv, Mv):
        K = sp.csr_matrix((Kv, ci, rp), shape=(prob.n_dof, prob.n_dof))[free][:, free]
        M = sp.csr_matrix((Mv, ci, rp), shape=(prob.n_dof, prob.n_dof))[free][:, free]
        return np.sort(spla.eigsh(K.tocsc(), k=3, M=M.tocsc(), sigma=0.0, which="LM")[0])
    w_dev, w_ref = lowest(A, B), lowest(Ko, Mo)
    assert np.allclose(w_dev, w_ref, rtol=1e-9)
    # first bending frequency of the cantilever within a few percent of Euler-Bernoulli (coarse mesh, trilinear elements lock a bit)
    E, rho, L, h = 72e9, 2700.0, 0.4, 0.05
    f_beam = (1.875104 ** 2) * np.sqrt(E * h ** 4 / 12 / (rho * h * h * L ** 4)) / (2 * np.pi)
    f_fe = np.sqrt(w_dev[0]) / (2 * np.pi)
    assert 0.9 * f_beam < f_fe < 1.6 * f_beam
