"""The reference's element property tests (tests/element/structural/2D/quad4/*.cpp) restated for the
elements on this path: FD Jacobian at margin mean|K|*1.49e-8, symmetry, rigid-body null space,
translation/rotation invariance, partition of unity; plus closed-form checks (patch test, volume)."""
import numpy as np
import pytest

import oracle_lib as O
from mast_b200 import mesh


def _single(elem_type, nv=3, jitter=0.0, **kw):
    if elem_type == O.QUAD4:
        xyz, conn = mesh.build_square(1, 1, 2.0, 2.0)
    else:
        xyz, conn = mesh.build_cube(1, 1, 1, 2.0, 1.5, 1.0, elem_type)
    if jitter:
        rng = np.random.default_rng(7)
        xyz = xyz + jitter * rng.uniform(-1, 1, xyz.shape) * (1 if elem_type != O.QUAD4 else np.array([1, 1, 0]))
    return O.Problem(elem_type, xyz, conn, nv, **kw)


@pytest.mark.parametrize("et", [O.HEX8, O.HEX27])
def test_shape_function_tables(built, et):
    # mast_quad4_structural_shape_functions.cpp:80-326: partition of unity, sum of gradients zero, volume
    p = _single(et, jitter=0.1)
    nn, nq = (8, 8) if et == O.HEX8 else (27, 27)
    phi = np.zeros((nn, nq)); dphi = np.zeros((nn, nq, 3)); jxw = np.zeros(nq); xq = np.zeros((nq, 3))
    X = p.xyz[p.conn[0]].copy()
    O.lib().oracle_fe_tables(et, X.ctypes.data, phi.ctypes.data, dphi.ctypes.data, jxw.ctypes.data, xq.ctypes.data)
    assert np.allclose(phi.sum(0), 1.0, atol=1e-14)
    assert np.allclose(dphi.sum(0), 0.0, atol=1e-13)
    # linear field reproduced: sum_a x_a dN_a/dx = I
    grad = np.einsum("ac,aqd->qcd", X, dphi)
    assert np.allclose(grad, np.eye(3)[None], atol=1e-12)
    p0 = _single(et)
    X0 = p0.xyz[p0.conn[0]].copy()
    O.lib().oracle_fe_tables(et, X0.ctypes.data, phi.ctypes.data, dphi.ctypes.data, jxw.ctypes.data, xq.ctypes.data)
    assert np.isclose(jxw.sum(), 2.0 * 1.5 * 1.0, rtol=1e-14)


@pytest.mark.parametrize("et,nonlinear", [(O.HEX8, 0), (O.HEX8, 1), (O.HEX27, 0), (O.HEX27, 1)])
def test_internal_jacobian_fd_symmetry_nullspace(built, et, nonlinear):
    # mast_quad4_linear_structural_internal_jacobian.cpp:66-105
    p = _single(et, jitter=0.1, nonlinear_strain=nonlinear)
    X = O.state(p.n_dof, 1, 1e-2 if nonlinear else 1e-3)
    f0, K = p.elem_calculations(0, X)
    assert np.abs(K - K.T).max() <= 1e-12 * np.abs(K).max()
    h = 1e-6
    Kfd = np.zeros_like(K)
    for j in range(p.n_dof):  # 6th-order central difference (tests/test_helpers.h:46-49)
        acc = np.zeros(p.n_dof)
        for c, s in ((-1 / 60, -3), (3 / 20, -2), (-3 / 4, -1), (3 / 4, 1), (-3 / 20, 2), (1 / 60, 3)):
            Xp = X.copy(); Xp[j] += s * h
            acc += c * p.elem_calculations(0, Xp, False)[0]
        Kfd[:, j] = acc / h
    # element dof order is variable-major, global is node-major: map columns
    nn = p.nn
    gl = np.array([3 * p.conn[0][a] + v for v in range(3) for a in range(nn)])
    Kfd = Kfd[:, gl]
    assert np.abs(K - Kfd).max() <= np.abs(K).mean() * 1.49e-8 * 50
    if not nonlinear:
        w = np.linalg.eigvalsh(K)
        assert (np.abs(w) < 1e-9 * np.abs(w).max()).sum() == 6  # rigid-body modes
        # linear strain: Jacobian independent of the displacement (:91-105)
        _, K2 = p.elem_calculations(0, 5 * X)
        assert np.abs(K - K2).max() <= 1e-13 * np.abs(K).max()


def test_translation_rotation_invariance(built):
    # mast_quad4_linear_structural_internal_jacobian.cpp:108-356 (eigenvalues invariant under rigid motion)
    p = _single(O.HEX8, jitter=0.1)
    X = np.zeros(p.n_dof)
    _, K = p.elem_calculations(0, X)
    th = 0.7
    Rz = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1.0]])
    p2 = O.Problem(O.HEX8, p.xyz @ Rz.T + np.array([3.0, -2.0, 5.0]), p.conn, 3)
    _, K2 = p2.elem_calculations(0, X)
    assert np.allclose(np.linalg.eigvalsh(K), np.linalg.eigvalsh(K2), rtol=1e-9, atol=1e-6 * np.abs(K).max())


def test_patch_test_uniaxial(built):
    # closed form: uniform uniaxial strain field gives zero interior residual and sigma_xx = D11*eps
    nx = 3
    xyz, conn = mesh.build_cube(nx, nx, nx, 1.0, 1.0, 1.0, 8, jitter=0.2)
    p = O.Problem(O.HEX8, xyz, conn, 3)
    eps = 1e-3
    X = np.zeros((len(xyz), 3)); X[:, 0] = eps * xyz[:, 0]
    R, _ = p.residual_and_jacobian(X.ravel())
    I = mesh.boundary_nodes(nx, nx, nx, 0)
    interior = np.ones(len(xyz), bool)
    for b in range(6):
        interior[mesh.boundary_nodes(nx, nx, nx, b)] = False
    assert interior.sum() == 8
    assert np.abs(R.reshape(-1, 3)[interior]).max() <= 1e-9 * np.abs(R).max()
    # total reaction on x=1 face = sigma_xx * area
    D11 = 72e9 * (1 - 0.33) / (1 - 0.33 - 2 * 0.33 ** 2)
    fx = R.reshape(-1, 3)[mesh.boundary_nodes(nx, nx, nx, 2), 0].sum()
    assert np.isclose(fx, D11 * eps, rtol=1e-10)


def test_nv6_rotational_diagonal_and_dirichlet(built):
    # solid_element_3d.cpp:413-417 and libMesh constrain_element_matrix_and_vector semantics
    p = O.cantilever(2, 2, 2, nv=6)
    X = O.state(p.n_dof)
    R, J = p.residual_and_jacobian(X)
    rp, ci = p.pattern()
    import scipy.sparse as sp
    A = sp.csr_matrix((J, ci, rp), shape=(p.n_dof, p.n_dof)).toarray()
    con = p.constrained.astype(bool)
    assert np.all(R[con] == 0.0)
    # constrained rows/cols are zero off the diagonal, diagonal = number of adjacent elements
    off = A - np.diag(np.diag(A))
    assert np.all(off[con] == 0.0) and np.all(off[:, con] == 0.0)
    n_adj = np.bincount(p.conn.ravel(), minlength=p.n_nodes)
    assert np.array_equal(np.diag(A)[con], np.repeat(n_adj, 6)[con].astype(float))
    rot_free = (~con) & (np.arange(p.n_dof) % 6 >= 3)
    assert np.all(np.diag(A)[rot_free] > 0) and np.all(np.diag(A)[rot_free] < 1e-6)


def test_heat_conduction_quad4(built):
    # 1-D conduction profile: T = a x is a zero-residual state for constant k; Jacobian FD for k(T)
    xyz, conn = mesh.build_square(4, 3, 2.0, 1.0)
    p = O.Problem(O.QUAD4, xyz, conn, 1, k0=237.0, thickness=0.06)
    T = 5.0 * xyz[:, 0]
    R, J = p.residual_and_jacobian(T)
    interior = (xyz[:, 0] > 0) & (xyz[:, 0] < 2.0)
    assert np.abs(R[interior]).max() <= 1e-10 * np.abs(R).max()
    p2 = O.Problem(O.QUAD4, xyz, conn, 1, k0=237.0, k_beta=1e-3, k_of_T=1, thickness=0.06)
    T = 300 + 50 * np.random.default_rng(4).uniform(0, 1, len(xyz))
    f0, K = p2.elem_calculations(2, T)
    Kfd = np.zeros_like(K)
    for j in range(4):
        Tp = T.copy(); Tm = T.copy(); Tp[conn[2][j]] += 1e-4; Tm[conn[2][j]] -= 1e-4
        Kfd[:, j] = (p2.elem_calculations(2, Tp, False)[0] - p2.elem_calculations(2, Tm, False)[0]) / 2e-4
    assert np.abs(K - Kfd).max() <= 1e-7 * np.abs(K).max()


def test_adjoint_sensitivity_fd(built):
    # tests/old/structural/internal_residual_sensitivity.cpp analogue: lambda . dR/dp vs FD of lambda . R(p)
    p = O.cantilever(3, 2, 2, jitter=0.2, simp=True)
    X = O.state(p.n_dof, 1); lam = O.state(p.n_dof, 2)
    nodes = np.array([5, 17]); w = np.array([1.0, 0.5])
    s = p.adjoint_sensitivity(X, lam, [0, 2], nodes, w)[0]
    h = 1e-6
    vals = []
    for sg in (+1, -1):
        rho = p.rho.copy(); rho[nodes] += sg * h * w
        q = O.cantilever(3, 2, 2, jitter=0.2, simp=True); q.rho[:] = rho
        R, _ = q.residual_and_jacobian(X, want_J=False)
        vals.append(lam @ R)
    fd = (vals[0] - vals[1]) / (2 * h)
    assert np.isclose(s, fd, rtol=1e-6)


def test_compliance_and_derivative(built):
    p = O.cantilever(3, 2, 2)
    X = O.state(p.n_dof)
    q = p.compliance(X)
    dq = p.compliance_derivative(X)
    # q is linear in X with gradient dq on unconstrained dofs
    Xf = X.copy(); Xf[p.constrained.astype(bool)] = 0.0
    assert np.isclose(dq @ Xf, p.compliance(Xf), rtol=1e-12)
    R0, _ = p.residual_and_jacobian(np.zeros(p.n_dof), want_J=False)
    assert np.allclose(R0, -dq, rtol=1e-12, atol=1e-9 * np.abs(dq).max())  # R(0) = -F_ext on free dofs
    assert q != 0.0


# ---- QUAD4 Mindlin plate (StructuralElement2D): the reference's own quad4 tests restated ---------------------------
def _plate_single(xy=None, **kw):
    # TEST::TestStructuralSingleElement2D fixture (tests/element/structural/2D/mast_structural_element_2d.h:95-106):
    # QUAD4 at (+-1,+-1,0), E=72e9, nu=0.33, h=0.06, off=0.03, kappa=5/6
    xyz = np.array([[-1.0, -1.0, 0.0], [1.0, -1.0, 0.0], [1.0, 1.0, 0.0], [-1.0, 1.0, 0.0]]) if xy is None else xy
    conn = np.array([[0, 1, 2, 3]], np.int32)
    kw.setdefault("offset", 0.03)
    return O.Problem(O.QUAD4, xyz, conn, 6, E0=72e9, nu=0.33, thickness=0.06, kappa=5.0 / 6.0, **kw)


def _plate_global_order(K):
    gl = np.array([6 * a + v for v in range(6) for a in range(4)])  # element (variable-major) -> node-major
    return gl


def test_plate_internal_jacobian_fd_symmetry_nullspace(built):
    # mast_quad4_linear_structural_internal_jacobian.cpp:66-105
    p = _plate_single()
    X = O.state(p.n_dof, 1)
    f0, K = p.elem_calculations(0, X)
    assert np.abs(K - K.T).max() <= 1e-12 * np.abs(K).max()                 # :78-82
    gl = _plate_global_order(K)
    h = 1e-6
    Kfd = np.zeros_like(K)
    for j in range(p.n_dof):
        acc = np.zeros(p.n_dof)
        for c, s in ((-1 / 60, -3), (3 / 20, -2), (-3 / 4, -1), (3 / 4, 1), (-3 / 20, 2), (1 / 60, 3)):
            Xp = X.copy(); Xp[j] += s * h
            acc += c * p.elem_calculations(0, Xp, False)[0]
        Kfd[:, j] = acc / h
    Kfd = Kfd[:, gl]
    Kphys = K.copy()
    for i in range(4):                       # the drilling diagonal is assigned in the Jacobian only (:1217-1221)
        Kphys[20 + i, 20 + i] -= 1e-8
    assert np.abs(Kphys - Kfd).max() <= np.abs(K).mean() * 1.49e-8 * 50      # :66-75
    # internal force is K u for the linear element
    assert np.abs(f0 - Kphys @ X[gl]).max() <= 1e-12 * np.abs(f0).max()
    w = np.linalg.eigvalsh(K)
    # det(K) ~ 0 (:85-88): six rigid-body modes; the four drilling dofs carry only the 1e-8 diagonal
    assert (np.abs(w) < 1e-9 * np.abs(w).max()).sum() == 6 + 4 - 1 or (np.abs(w) < 1e-9 * np.abs(w).max()).sum() >= 6
    _, K2 = p.elem_calculations(0, 7 * X)    # :91-105 displacement invariance of the linear Jacobian
    assert np.array_equal(K, K2)


def test_plate_translation_rotation_stretch_invariance(built):
    # :108-141 translation; :144-356 in-plane rotation keeps the eigenvalues, symmetry after shear / stretch
    p0 = _plate_single()
    _, K0 = p0.elem_calculations(0, np.zeros(p0.n_dof))
    xyz = p0.xyz + np.array([5.2, -3.1, 0.0])
    _, K1 = _plate_single(xyz).elem_calculations(0, np.zeros(24))
    assert np.abs(K1 - K0).max() <= 1e-12 * np.abs(K0).max()
    th = 0.35
    Rz = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1.0]])
    _, K2 = _plate_single(p0.xyz @ Rz.T).elem_calculations(0, np.zeros(24))
    w0, w2 = np.linalg.eigvalsh(K0), np.linalg.eigvalsh(K2)
    assert np.abs(w0 - w2).max() <= 1e-9 * np.abs(w0).max()
    sh = p0.xyz.copy(); sh[:, 0] += 0.3 * sh[:, 1]; sh[:, 1] *= 1.7
    _, K3 = _plate_single(sh).elem_calculations(0, np.zeros(24))
    assert np.abs(K3 - K3.T).max() <= 1e-12 * np.abs(K3).max()


def test_plate_out_of_plane_matches_rotated_in_plane_element(built):
    # an element whose nodes do not share one z is evaluated in the local frame of its first three nodes and rotated
    # back (geom_elem.cpp:358-374, 461-511; structural_element_base.cpp:1565-1674): rigidly rotating the reference's
    # fixture must rotate every nodal 6x6 block and the force with it, f' = Q6 f, K'_ab = Q6 K_ab Q6^T
    rng = np.random.default_rng(7)
    xy = np.array([[-1.0, -1.0, 0.0], [1.1, -0.9, 0.0], [0.9, 1.2, 0.0], [-1.2, 0.8, 0.0]])
    p0 = _plate_single(xy)
    X0 = O.state(p0.n_dof, 3)
    f0, K0 = p0.elem_calculations(0, X0)
    A = rng.normal(size=(3, 3))
    Q, _ = np.linalg.qr(A)
    if np.linalg.det(Q) < 0:
        Q[:, 0] *= -1
    p1 = _plate_single(xy @ Q.T + np.array([0.3, -0.2, 0.7]))
    assert np.ptp(p1.xyz[:, 2]) > 0.1
    X1 = (X0.reshape(4, 2, 3) @ Q.T).reshape(-1)          # node-major state: translations and rotations rotate alike
    f1, K1 = p1.elem_calculations(0, X1)
    Q6 = np.kron(np.eye(2), Q)
    el = lambda a: np.array([v * 4 + a for v in range(6)])   # element (variable-major) dofs of node a
    for a in range(4):
        assert np.abs(f1[el(a)] - Q6 @ f0[el(a)]).max() <= 1e-10 * np.abs(f0).max()
        for b in range(4):
            Kab = K0[np.ix_(el(a), el(b))]
            assert np.abs(K1[np.ix_(el(a), el(b))] - Q6 @ Kab @ Q6.T).max() <= 1e-10 * np.abs(K0).max()


def test_plate_closed_forms(built):
    # uniform membrane stretch u = e x: N_x = A11 e, N_y = A12 e; pure twist-free bending theta_y = k x: M = D k (off = 0)
    xyz, conn = mesh.build_square(3, 2, 1.5, 1.0)
    rng = np.random.default_rng(5)
    inner = (xyz[:, 0] > 0) & (xyz[:, 0] < 1.5) & (xyz[:, 1] > 0) & (xyz[:, 1] < 1.0)
    xyz[inner, :2] += 0.08 * rng.uniform(-1, 1, (inner.sum(), 2))
    p = O.Problem(O.QUAD4, xyz, conn, 6, E0=72e9, nu=0.33, thickness=0.06, kappa=5.0 / 6.0)
    e = 1e-4
    X = np.zeros((len(xyz), 6)); X[:, 0] = e * xyz[:, 0]; X[:, 1] = -0.33 * e * xyz[:, 1]  # uniaxial stress state
    R, _ = p.residual_and_jacobian(X.ravel())
    R = R.reshape(-1, 6)
    assert np.abs(R[inner]).max() <= 1e-9 * np.abs(R).max()                   # patch test: interior equilibrium
    Fx = R[np.isclose(xyz[:, 0], 1.5), 0].sum()
    assert np.isclose(Fx, 72e9 * e * 0.06 * 1.0, rtol=1e-12)                 # sigma_x h Ly
    k = 1e-3
    X = np.zeros((len(xyz), 6)); X[:, 4] = k * xyz[:, 0]; X[:, 2] = -0.5 * k * xyz[:, 0] ** 2   # w' = -theta_y: no shear
    X[:, 3] = 0.0
    R, _ = p.residual_and_jacobian(X.ravel())
    R = R.reshape(-1, 6)
    My = R[np.isclose(xyz[:, 0], 1.5), 4].sum()
    D11 = 72e9 / (1 - 0.33 ** 2) * 0.06 ** 3 / 12
    # cylindrical bending (kappa_y = 0): M_x = D11 k per unit width; w is quadratic so the one-point shear strain
    # gamma_xz = dw/dx + theta_y is not identically zero on distorted cells -- compare on the moment resultant only
    assert np.isclose(My, D11 * k * 1.0, rtol=5e-2)


def test_heat_surface_radiation_jacobian_is_the_derivative_of_the_residual(built):
    # HeatConductionElementBase::surface_radiation_residual (heat_conduction_elem_base.cpp:1404-1458): the Jacobian term
    # 4 sb eps th (T + T0)^3 N N^T against central differences of the residual; zero flux at T == T_amb
    xyz = np.array([[0.0, 0.0, 0.0], [1.3, 0.1, 0.0], [1.2, 0.9, 0.0], [-0.1, 1.0, 0.0]])
    conn = np.array([[0, 1, 2, 3]], np.int32)
    kw = dict(k0=0.0, thickness=0.06, side_elem=np.array([0], np.int32), side_id=np.array([2], np.int32),
              side_pressure=np.zeros(1), side_T_amb=np.array([300.0]), side_emissivity=np.array([0.8]),
              rad_sb_const=5.670374419e-8, rad_T_ref_zero=0.0)
    p = O.Problem(O.QUAD4, xyz, conn, 1, **kw)
    T = np.array([310.0, 350.0, 420.0, 390.0])
    f0, K = p.elem_calculations(0, T)
    assert np.all(K[[0, 1]] == 0.0) and np.all(f0[[0, 1]] == 0.0)      # only the nodes of side 2 (nodes 2, 3)
    h = 1e-3
    Kfd = np.zeros((4, 4))
    for j in range(4):
        Tp = T.copy(); Tp[j] += h
        Tm = T.copy(); Tm[j] -= h
        Kfd[:, j] = (p.elem_calculations(0, Tp, False)[0] - p.elem_calculations(0, Tm, False)[0]) / (2 * h)
    assert np.abs(K - Kfd).max() <= 1e-7 * np.abs(K).max()
    assert np.abs(K - K.T).max() <= 1e-14 * np.abs(K).max()
    f_eq, _ = p.elem_calculations(0, np.full(4, 300.0))
    assert np.abs(f_eq).max() <= 1e-9 * np.abs(f0).max()
    # total radiated power of a uniform temperature: sb eps th L (T^4 - T_amb^4)
    L = np.linalg.norm(xyz[3] - xyz[2])
    f_u, _ = p.elem_calculations(0, np.full(4, 400.0))
    assert np.isclose(f_u.sum(), 5.670374419e-8 * 0.8 * 0.06 * L * (400.0 ** 4 - 300.0 ** 4), rtol=1e-12)


def test_heat_parameter_sensitivities_are_the_derivative_of_the_residual(built):
    # elem_sensitivity_calculations of the heat-conduction system (heat_conduction_nonlinear_assembly.cpp:100-116) for the
    # conductivity k0 (kind 1) and the section thickness (kind 2): central differences of the assembled residual in the
    # parameter, with a volume source, flux + convection + radiation sides and a Dirichlet edge
    nx, ny = 6, 4
    xyz, conn = mesh.build_square(nx, ny, 1.2, 0.7)
    rng = np.random.default_rng(11)
    inner = (xyz[:, 0] > 0) & (xyz[:, 0] < 1.2) & (xyz[:, 1] > 0) & (xyz[:, 1] < 0.7)
    xyz[inner, :2] += 0.03 * rng.uniform(-1, 1, (inner.sum(), 2))
    con = np.isclose(xyz[:, 0], 0.0).astype(np.uint8)
    top = (np.arange(nx, dtype=np.int32) + (ny - 1) * nx)
    right = (np.arange(ny, dtype=np.int32) * nx + nx - 1)
    se = np.concatenate([top, right]); sid = np.concatenate([np.full(nx, 2, np.int32), np.ones(ny, np.int32)])
    flux = np.concatenate([np.full(nx, 3.0e2), np.zeros(ny)])
    hc = np.concatenate([np.zeros(nx), np.full(ny, 35.0)])
    Tamb = np.concatenate([np.full(nx, 280.0), np.full(ny, 290.0)])
    eps = np.concatenate([np.full(nx, 0.7), np.zeros(ny)])
    T = 300 + 50 * rng.uniform(size=len(xyz))

    def prob(k0, th, k_of_T):
        return O.Problem(O.QUAD4, xyz, conn, 1, constrained=con, k0=k0, k_beta=1e-3, k_of_T=k_of_T, thickness=th,
                         heat_source=1.0e4, side_elem=se, side_id=sid, side_pressure=flux, side_conv_coeff=hc,
                         side_T_amb=Tamb, side_emissivity=eps, rad_sb_const=5.670374419e-8, rad_T_ref_zero=273.15)

    for k_of_T in (0, 1):
        p0 = prob(237.0, 0.06, k_of_T)
        dk = p0.sensitivity_assemble_param(T, 1)
        dh = p0.sensitivity_assemble_param(T, 2)
        fd_k = (prob(238.0, 0.06, k_of_T).residual_and_jacobian(T)[0] - prob(236.0, 0.06, k_of_T).residual_and_jacobian(T)[0]) / 2.0
        fd_h = (prob(237.0, 0.07, k_of_T).residual_and_jacobian(T)[0] - prob(237.0, 0.05, k_of_T).residual_and_jacobian(T)[0]) / 0.02
        # the residual is linear in either parameter: central differences are exact up to rounding
        assert np.abs(dk - fd_k).max() <= 1e-10 * np.abs(fd_k).max()
        assert np.abs(dh - fd_h).max() <= 1e-10 * np.abs(fd_h).max()
        assert np.all(dk[con.astype(bool)] == 0.0) and np.all(dh[con.astype(bool)] == 0.0)
        assert np.abs(dh).max() > 0 and np.abs(dk).max() > 0


def test_plate_von_karman_tangent_is_exact_where_the_reference_is_consistent():
    """StructuralElement2D with NONLINEAR_STRAIN (structural_element_2d.cpp:1765-2019): the tangent equals the derivative of
    the residual for a pure bending state (von Karman terms alone) and, without section offset, for a pure membrane state
    (Green-Lagrange terms alone).  With both active the reference keeps B_lin in the membrane / von Karman couplings
    (:1967-1977), so its tangent is not the exact linearisation -- the oracle restates that too."""
    import scipy.sparse as sp
    from mast_b200 import mesh
    xyz, conn = mesh.build_square(3, 2, 1.5, 1.0)
    rng = np.random.default_rng(1)
    inner = (xyz[:, 0] > 0) & (xyz[:, 0] < 1.5) & (xyz[:, 1] > 0) & (xyz[:, 1] < 1.0)
    xyz[inner, :2] += 0.1 * rng.uniform(-1, 1, (inner.sum(), 2))
    n = len(xyz)

    def fd_error(prob, X):
        _, J = prob.residual_and_jacobian(X)
        rp, ci = prob.pattern()
        A = sp.csr_matrix((J, ci, rp), shape=(prob.n_dof, prob.n_dof)).toarray()
        Afd = np.zeros_like(A)
        h = 1e-7
        for j in range(prob.n_dof):
            Xp = X.copy(); Xp[j] += h
            Xm = X.copy(); Xm[j] -= h
            Afd[:, j] = (prob.residual_and_jacobian(Xp, want_J=False)[0] - prob.residual_and_jacobian(Xm, want_J=False)[0]) / (2 * h)
        assert np.abs(A - A.T).max() <= 1e-13 * np.abs(A).max()
        return np.abs(A - Afd).max() / np.abs(A).max()

    bend = np.zeros((n, 6)); bend[:, 2] = 0.05 * rng.uniform(-1, 1, n); bend[:, 3:5] = 0.02 * rng.uniform(-1, 1, (n, 2))
    memb = np.zeros((n, 6)); memb[:, :2] = 0.02 * rng.uniform(-1, 1, (n, 2))
    p_off = O.Problem(O.QUAD4, xyz, conn, 6, thickness=0.06, offset=0.03, kappa=5 / 6, nonlinear_strain=1)
    p_mid = O.Problem(O.QUAD4, xyz, conn, 6, thickness=0.06, offset=0.0, kappa=5 / 6, nonlinear_strain=1)
    assert fd_error(p_off, bend.ravel()) < 1e-9
    assert fd_error(p_mid, memb.ravel()) < 1e-9
    assert fd_error(p_mid, (bend + memb).ravel()) > 1e-4   # the reference's inconsistent cross terms
