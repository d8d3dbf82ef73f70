"""The device linear solve on the assembled CSR (mastb200_solve: block-Jacobi PCG) and the rigid-body near null space,
against scipy on the oracle's matrix."""
import numpy as np
import pytest

import oracle_lib as O
from mast_b200 import mesh

pytestmark = pytest.mark.gpu


def _problem(elem_type, shape):
    nx, ny, nz = shape
    xyz, conn = mesh.build_cube(nx, ny, nz, 0.3, 0.2, 0.15, elem_type, jitter=0.2)
    con = mesh.dirichlet_flags(len(xyz), 3, mesh.boundary_nodes(nx, ny, nz, 4, elem_type))
    se, sid = mesh.boundary_sides(nx, ny, nz, 3)
    return O.Problem(elem_type, xyz, conn, 3, constrained=con, side_elem=se, side_id=sid, side_pressure=np.full(len(se), 5e7))


@pytest.mark.parametrize("elem_type,shape", [(8, (12, 8, 6)), (27, (4, 3, 3))])
def test_solve_matches_scipy(built, elem_type, shape):
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    import torch
    prob = _problem(elem_type, shape)
    ctx = prob.make_context(0)
    X0 = np.zeros(prob.n_dof)
    dJ = torch.empty(ctx.nnz, dtype=torch.float64, device="cuda")
    R = np.empty(prob.n_dof)
    ctx.residual_and_jacobian(X0, R, dJ)          # J stays on the device; Newton step: J dx = -R
    rhs = -R
    dx = np.full(prob.n_dof, np.nan)
    info = ctx.solve(dJ, rhs, dx, rtol=1e-10, max_it=20000)
    assert info["converged"] == 1 and info["rel_residual"] <= 1e-10
    # reference: the oracle's matrix, direct solve (and scipy's CG to the reference tolerance of the verdict: 1e-8)
    Ro, Jo = prob.residual_and_jacobian(X0)
    rp, ci = prob.pattern()
    A = sp.csr_matrix((Jo, ci, rp), shape=(prob.n_dof, prob.n_dof))
    ref = spla.spsolve(A.tocsc(), -Ro)
    assert np.abs(dx - ref).max() <= 1e-7 * np.abs(ref).max()
    cg, flag = spla.cg(A, -Ro, rtol=1e-10, maxiter=20000, M=None)
    assert flag == 0 and np.abs(dx - cg).max() <= 1e-6 * np.abs(ref).max()
    assert np.linalg.norm(A @ dx + Ro) <= 1e-8 * np.linalg.norm(Ro)
    # host J and device vectors work too, and the solve is bitwise reproducible
    dx2 = torch.empty(prob.n_dof, dtype=torch.float64, device="cuda")
    info2 = ctx.solve(dJ.cpu().numpy(), torch.from_numpy(rhs).cuda(), dx2, rtol=1e-10, max_it=20000)
    assert info2["iterations"] == info["iterations"] and np.array_equal(dx2.cpu().numpy(), dx)


def test_near_null_space_is_in_the_kernel_of_the_free_body(built):
    import scipy.sparse as sp
    xyz, conn = mesh.build_cube(5, 4, 3, 0.3, 0.2, 0.15, 8, jitter=0.2)
    prob = O.Problem(O.HEX8, xyz, conn, 3)        # no constraints: six rigid-body modes
    ctx = prob.make_context(0)
    modes = np.full((6, prob.n_dof), np.nan)
    ctx.near_null_space(modes)
    _, Jo = prob.residual_and_jacobian(np.zeros(prob.n_dof))
    rp, ci = prob.pattern()
    A = sp.csr_matrix((Jo, ci, rp), shape=(prob.n_dof, prob.n_dof))
    th = 1e-5
    for m in range(6):
        v = modes[m]
        # rotations by th are rigid up to O(th^2): the strain of the mode is second order
        tol = 1e-11 if m < 3 else 5 * th
        assert np.abs(A @ v).max() <= tol * np.abs(Jo).max() * np.abs(v).max()
    assert np.allclose(modes[0].reshape(-1, 3), [1, 0, 0]) and np.allclose(modes[2].reshape(-1, 3), [0, 0, 1])
    # the reference's formula for the rotation about z (structural_near_null_vector_space.cpp:88-97)
    c, s = np.cos(th), np.sin(th)
    ref = np.stack([(c * xyz[:, 0] - s * xyz[:, 1]) - xyz[:, 0], (s * xyz[:, 0] + c * xyz[:, 1]) - xyz[:, 1], np.zeros(len(xyz))], 1)
    assert np.allclose(modes[5].reshape(-1, 3), ref, rtol=0, atol=1e-15)  # one ulp of cos(th) times |x|
