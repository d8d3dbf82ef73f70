"""Sub-mesh parity checker: compares sampled CSR rows / residual entries of a LARGE device assembly with the CPU oracle.

The oracle cannot assemble an 8-16 M element mesh in test time, but a row of J depends only on the elements around its
node.  A box of elements is cut out of the big mesh (same node order, so the compact numbering is monotone in the global
one), handed to the oracle as a small problem with the same nodal data (coordinates, state, density, temperature,
constraint flags, loaded sides), and the rows of the nodes whose adjacent elements all lie inside the box are compared
entry by entry with the device's rows: column indices bit-exact, values entry-wise.  Test infrastructure (tests/, bench.py's
cpu_baseline leg) only.
"""
import numpy as np

import oracle_lib as O

EPS = np.finfo(np.float64).eps


def close_entries(a, b, rtol=1e-12, terms=64, scale=None, what=""):
    """Entry-wise: |a - b| <= rtol |b| + terms * eps * scale, scale = max |b| unless given (the magnitude of the summed
    terms; an entry that cancels to ~0 cannot be better than eps times the terms it is the sum of)."""
    a = np.asarray(a); b = np.asarray(b)
    sc = np.abs(b).max() if scale is None else scale
    tol = rtol * np.abs(b) + terms * EPS * sc
    bad = np.abs(a - b) > tol
    if bad.any():
        k = int(np.argmax(np.abs(a - b) - tol))
        raise AssertionError(f"{what}: {int(bad.sum())} of {a.size} entries differ; worst at {k}: {a.flat[k]!r} vs {b.flat[k]!r} "
                             f"(tol {tol.flat[k]:.3e})")
    return float(np.max(np.abs(a - b) / (np.abs(b) + terms * EPS * sc))) if a.size else 0.0


class BoxMesh:
    """Structured (build_cube numbered) HEX8 / HEX27 mesh description."""

    def __init__(self, nx, ny, nz, elem_type=8):
        self.nx, self.ny, self.nz, self.et = nx, ny, nz, elem_type
        self.m = 1 if elem_type == 8 else 2
        self.npx, self.npy, self.npz = self.m * nx + 1, self.m * ny + 1, self.m * nz + 1

    def node(self, i, j, k):
        return i + j * self.npx + k * self.npx * self.npy


def sub_problem(bm, box, xyz, conn, nv, constrained=None, rho=None, temperature=None, pressure_top=None, **kw):
    """Oracle problem of the element box [i0,i1) x [j0,j1) x [k0,k1) and the map local node -> global node, plus the
    local nodes whose every adjacent element of the big mesh lies inside the box."""
    (i0, i1), (j0, j1), (k0, k1) = box
    ii, jj, kk = np.meshgrid(np.arange(i0, i1), np.arange(j0, j1), np.arange(k0, k1), indexing="ij")
    elems = np.sort((ii + jj * bm.nx + kk * bm.nx * bm.ny).ravel())
    gconn = conn[elems]
    gnodes = np.unique(gconn)                      # ascending: local numbering is monotone in the global one
    lconn = np.searchsorted(gnodes, gconn).astype(np.int32)
    m = bm.m
    gi = gnodes % bm.npx; gj = (gnodes // bm.npx) % bm.npy; gk = gnodes // (bm.npx * bm.npy)

    def inner(g, lo, hi, n_el):  # node coordinate g (in node units) is interior to the element range, or on the mesh boundary
        lo_ok = (g > m * lo) | ((lo == 0) & (g == 0))
        hi_ok = (g < m * hi) | ((hi == n_el) & (g == m * n_el))
        return lo_ok & hi_ok
    complete = inner(gi, i0, i1, bm.nx) & inner(gj, j0, j1, bm.ny) & inner(gk, k0, k1, bm.nz)
    con = None if constrained is None else np.ascontiguousarray(constrained.reshape(-1, nv)[gnodes].ravel())
    se = sid = sp = None
    if pressure_top is not None:  # uniform pressure on the face y = Ly (libMesh side 3 of the top element layer)
        ej = (elems // bm.nx) % bm.ny
        se = np.flatnonzero(ej == bm.ny - 1).astype(np.int32)
        sid = np.full(len(se), 3, np.int32); sp = np.full(len(se), float(pressure_top))
        if len(se) == 0:
            se = sid = sp = None
    prob = O.Problem(bm.et, xyz[gnodes], lconn, nv, constrained=con,
                     rho=None if rho is None else rho[gnodes],
                     temperature=None if temperature is None else temperature[gnodes],
                     side_elem=se, side_id=sid, side_pressure=sp, **kw)
    return prob, gnodes, np.flatnonzero(complete)


def default_boxes(bm, w=4):
    """Element boxes at the start, in the middle and at the end of the mesh (the end box holds the last CSR rows)."""
    def rng(n, pos):
        w_ = min(w, n)
        a = {0: 0, 1: max((n - w_) // 2, 0), 2: n - w_}[pos]
        return (a, a + w_)
    return [tuple(rng(n, pos) for n in (bm.nx, bm.ny, bm.nz)) for pos in (0, 1, 2)]


def check_rows(ctx, bm, box, xyz, conn, nv, X, J=None, R=None, G=None, adj=None, node_lo=0, owned_range=None, **prob_kw):
    """Compares the device rows of the box's complete nodes with the oracle.  J / R / G: device outputs as torch CUDA
    tensors or numpy arrays (any may be None); node_lo: first global node of a rank-local numbering; owned_range: global
    node range [lo, hi) whose rows this context owns (None: all).  Returns a dict of the largest relative errors."""
    prob, gnodes, complete = sub_problem(bm, box, xyz, conn, nv, **prob_kw)
    if owned_range is not None:
        complete = complete[(gnodes[complete] >= owned_range[0]) & (gnodes[complete] < owned_range[1])]
    out = {"nodes": int(len(complete))}
    if len(complete) == 0:
        return out
    Xl = np.ascontiguousarray(np.asarray(X).reshape(-1, nv)[gnodes].ravel())
    orp, oci = prob.pattern()
    Ro, Jo = prob.residual_and_jacobian(Xl, want_R=R is not None, want_J=J is not None)

    def fetch(t, idx):
        if hasattr(t, "is_cuda"):
            import torch
            return t[torch.as_tensor(idx, device=t.device)].cpu().numpy()
        return np.asarray(t)[idx]
    first_owned = 0 if owned_range is None else owned_range[0]
    ldofs = (complete[:, None] * nv + np.arange(nv)[None, :]).ravel()           # oracle rows
    gdofs = (gnodes[complete][:, None] * nv + np.arange(nv)[None, :]).ravel()   # global dofs
    if J is not None:
        rows = gdofs - first_owned * nv                                          # CSR rows of this context
        begin, length, cols = ctx.get_pattern_rows(rows)
        idx = np.concatenate([np.arange(b, b + l) for b, l in zip(begin, length)])
        vals = fetch(J, idx)
        worst = 0.0; k = 0
        for r, l in zip(ldofs, length):
            oc = oci[orp[r]:orp[r + 1]]
            gc = (gnodes[oc // nv] - node_lo) * nv + oc % nv   # oracle columns in the context's local dof numbering
            assert l == len(oc) and np.array_equal(cols[k:k + l], gc), f"sparsity of row {r} differs"
            ref = Jo[orp[r]:orp[r + 1]]
            # an entry sums (adjacent elements) x (quadrature points) terms: 8 x 8 for HEX8, 8 x 27 for HEX27
            worst = max(worst, close_entries(vals[k:k + l], ref, terms=8 * prob.nn, what=f"J row of dof {r}"))
            k += l
        out["J"] = worst; out["rows"] = int(len(rows)); out["max_value_index"] = int((begin + length).max())
    if R is not None:
        # floor of an R entry: eps times the sum of the magnitudes of its terms K_ij u_j (oracle J unavailable when
        # only R is requested, so the scale is the largest |R| of the sample times the condition of the sum)
        got = fetch(R, gdofs - node_lo * nv)
        ref = Ro[ldofs]
        if J is not None:
            sc = np.array([np.abs(Jo[orp[r]:orp[r + 1]] * Xl[oci[orp[r]:orp[r + 1]]]).sum() for r in ldofs])
            tol = 1e-12 * np.abs(ref) + 8 * prob.nn * EPS * np.maximum(sc, np.abs(ref).max())
            assert np.all(np.abs(got - ref) <= tol), f"R differs: {np.abs(got - ref).max():.3e}"
            out["R"] = float(np.max(np.abs(got - ref) / tol))
        else:
            out["R"] = close_entries(got, ref, what="R")
    if G is not None:
        n_loc = prob.n_nodes
        g_ref = prob.adjoint_sensitivity(Xl, Xl if adj is None else np.ascontiguousarray(np.asarray(adj).reshape(-1, nv)[gnodes].ravel()),
                                         np.arange(n_loc + 1), np.arange(n_loc), np.ones(n_loc))
        got = fetch(G, gnodes[complete] - node_lo)
        out["G"] = close_entries(got, g_ref[complete], rtol=1e-11, terms=512, what="G")
    return out
