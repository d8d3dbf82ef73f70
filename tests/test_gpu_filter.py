"""The density filter and the material volume on the device (mastb200_filter_*, mastb200_volume) against the oracle's
restatement of MAST::FilterBase (linear radius search) and of example_6's volume."""
import numpy as np
import pytest

import oracle_lib as O
from mast_b200 import mesh
from submesh import close_entries

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("elem_type,shape,radius", [(8, (9, 6, 5), 0.071), (27, (3, 3, 2), 0.06), (8, (4, 4, 4), 1e-4)])
def test_filter_matches_the_reference_search(built, elem_type, shape, radius):
    xyz, conn = mesh.build_cube(*shape, 0.3, 0.2, 0.15, elem_type, jitter=0.2)
    prob = O.Problem(elem_type, xyz, conn, 3)
    ctx = prob.make_context(0)
    nnz = ctx.filter_init(radius)
    rp, col, w = ctx.filter_get()
    orp, ocol, ow = prob.filter_weights(radius)
    assert nnz == len(ocol) and np.array_equal(rp, orp) and np.array_equal(col, ocol)   # the neighbour sets, in order
    close_entries(w, ow, rtol=1e-14, terms=4, what="filter weights")
    rng = np.random.default_rng(2)
    d = rng.uniform(0.0, 1.0, prob.n_nodes)
    rho = np.full(prob.n_nodes, np.nan)
    ctx.filter_apply(d, rho)
    close_entries(rho, prob.filtered_values((orp, ocol, ow), d), rtol=1e-13, what="filtered values")
    # transpose: <W d, G> == <d, W^T G>, and against the explicit matrix
    import scipy.sparse as sp
    W = sp.csr_matrix((ow, ocol, orp), shape=(prob.n_nodes, prob.n_nodes))
    G = rng.uniform(-1.0, 1.0, prob.n_nodes)
    grad = np.full(prob.n_nodes, np.nan)
    ctx.filter_apply_transpose(G, grad)
    close_entries(grad, W.T @ G, rtol=1e-12, what="transposed filter")
    # some nodes are not design variables: identity rows
    dv = (rng.uniform(size=prob.n_nodes) < 0.7).astype(np.uint8)
    ctx.filter_init(radius, dv)
    ctx.filter_apply(d, rho)
    close_entries(rho, prob.filtered_values((orp, ocol, ow), d, dv), rtol=1e-13, what="filtered values, partial DV set")
    Wd = sp.lil_matrix(W)
    for n in np.flatnonzero(dv == 0):
        Wd.rows[n] = [int(n)]; Wd.data[n] = [1.0]
    ctx.filter_apply_transpose(G, grad)
    close_entries(grad, Wd.tocsr().T @ G, rtol=1e-12, what="transposed filter, partial DV set")


@pytest.mark.parametrize("elem_type,shape", [(8, (7, 5, 4)), (27, (3, 2, 2))])
def test_volume_and_gradient(built, elem_type, shape):
    import torch
    xyz, conn = mesh.build_cube(*shape, 0.3, 0.2, 0.15, elem_type, jitter=0.2)
    prob = O.Problem(elem_type, xyz, conn, 3)
    ctx = prob.make_context(0)
    rho = 0.2 + 0.6 * np.random.default_rng(3).uniform(size=prob.n_nodes)
    g = np.full(prob.n_nodes, np.nan)
    v = ctx.volume(rho, g)
    vo, go = prob.volume(rho, want_gradient=True)
    assert np.isclose(v, vo, rtol=1e-13)
    close_entries(g, go, rtol=1e-13, what="volume gradient")
    assert np.isclose(ctx.volume(torch.from_numpy(rho).cuda()), vo, rtol=1e-13)
    assert np.isclose(ctx.volume(np.ones(prob.n_nodes)), 0.3 * 0.2 * 0.15, rtol=1e-12)


def test_simp_iteration_stays_on_the_device(built):
    """One design iteration with every array on the device: d -> rho = W d -> sweep (R, J, G) -> grad = W^T G, volume."""
    import torch
    prob = O.cantilever(10, 6, 5, jitter=0.2, simp=True)
    ctx = prob.make_context(0)
    n = prob.n_nodes
    ctx.filter_init(0.05)
    d = torch.from_numpy(0.2 + 0.6 * np.random.default_rng(5).uniform(size=n)).cuda()
    rho = torch.empty(n, dtype=torch.float64, device="cuda")
    ctx.filter_apply(d, rho)
    ctx.set_density(rho)
    X = torch.from_numpy(O.state(prob.n_dof, 1)).cuda()
    G = torch.empty(n, dtype=torch.float64, device="cuda"); grad = torch.empty_like(G)
    R = torch.empty(prob.n_dof, dtype=torch.float64, device="cuda"); J = torch.empty(ctx.nnz, dtype=torch.float64, device="cuda")
    ctx.sensitivity_sweep(X, X, G, R, J)
    ctx.filter_apply_transpose(G, grad)
    vol = ctx.volume(rho)
    # the same chain through the oracle
    filt = prob.filter_weights(0.05)
    rho_o = prob.filtered_values(filt, d.cpu().numpy())
    p2 = O.Problem(O.HEX8, prob.xyz, prob.conn, 3, constrained=prob.constrained, rho=rho_o, rho_min=1e-8, penalty=4.0,
                   side_elem=prob.side_elem, side_id=prob.side_id, side_pressure=prob.side_pressure)
    Xh = X.cpu().numpy()
    g_ref = p2.adjoint_sensitivity(Xh, Xh, np.arange(n + 1), np.arange(n), np.ones(n))
    import scipy.sparse as sp
    W = sp.csr_matrix((filt[2], filt[1], filt[0]), shape=(n, n))
    close_entries(grad.cpu().numpy(), W.T @ g_ref, rtol=1e-11, terms=512, what="design gradient")
    assert np.isclose(vol, p2.volume(rho_o), rtol=1e-12)
