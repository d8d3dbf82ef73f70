"""Element-ownership partition of a structured HEX8 block over the GPUs of one box.

The mesh is cut into z-slabs of node planes.  A rank OWNS the nodes of its planes -- its rows of R and J, a contiguous
block of dofs exactly like libMesh's ``first_dof/end_dof`` blocks
(/root/reference/src/base/nonlinear_implicit_assembly.cpp:201-203) -- and sweeps every element touching an owned node,
i.e. its own layers plus one layer of ghost elements per cut, recomputed redundantly so that no matrix rows travel
(<= 2 n^2 of n^3/N elements).  The only exchange on the data path is the pre-sweep halo of the ghost node planes of X
(and of the adjoint / density for sensitivity sweeps): the device analogue of ``AssemblyBase::build_localized_vector``
(src/base/assembly_base.cpp:209-228).  Local node numbering is global numbering minus ``node_lo``.
"""
import numpy as np

from . import mesh


def slab_planes(n_planes, rank, world):
    """Owned node planes [p0, p1) of ``rank``."""
    return n_planes * rank // world, n_planes * (rank + 1) // world


def slab_problem(n, rank, world, jitter, nv=3):
    """Rank-local view of the n^3 cantilever (clamped at x=0, pressure on y=1): local xyz / conn / owned flags /
    constrained flags / loaded sides, and the halo bookkeeping ``info``."""
    xyz, conn = mesh.build_cube(n, n, n, 1.0, 1.0, 1.0, 8, jitter=jitter)
    npl = (n + 1) * (n + 1)
    p0, p1 = slab_planes(n + 1, rank, world)
    l0, l1 = max(p0 - 1, 0), min(p1 - 1, n - 1)   # element layers touching owned planes
    q0, q1 = l0, l1 + 1                            # local node planes [q0, q1]
    node_lo = q0 * npl
    lxyz = xyz[node_lo:(q1 + 1) * npl]
    lconn = conn[l0 * n * n:(l1 + 1) * n * n] - node_lo
    plane = np.arange(len(lxyz)) // npl + q0
    owned = ((plane >= p0) & (plane < p1)).astype(np.uint8)
    con = np.zeros((len(lxyz), nv), np.uint8)
    con[np.isclose(lxyz[:, 0], 0.0)] = 1
    j = (np.arange(len(lconn)) // n) % n           # pressure on y = 1 (libMesh side 3) of the local elements
    se = np.flatnonzero(j == n - 1).astype(np.int32)
    e0, e1 = min(p0, n), min(p1, n)                # element layers counted for this rank (each element once)
    info = dict(p0=p0, p1=p1, q0=q0, q1=q1, npl=npl, node_lo=node_lo, n_owned_elem=(e1 - e0) * n * n,
                elem_lo=l0 * n * n)
    return lxyz, lconn.astype(np.int32), owned, con.ravel(), se, info


def halo_ops(dist, X, info, rank, world, width):
    """P2P operations that refresh the ghost planes of the local vector ``X`` (a torch tensor with ``width`` values per
    node: nv for X / adjoint, 1 for density).  Each rank sends its first / last owned plane and receives the plane
    below / above."""
    ops = []
    w = info["npl"] * width
    p0, p1, q0 = info["p0"], info["p1"], info["q0"]
    if rank > 0:
        ops.append(dist.P2POp(dist.isend, X[(p0 - q0) * w:(p0 - q0 + 1) * w], rank - 1))
        ops.append(dist.P2POp(dist.irecv, X[(p0 - 1 - q0) * w:(p0 - q0) * w], rank - 1))
    if rank < world - 1:
        ops.append(dist.P2POp(dist.isend, X[(p1 - 1 - q0) * w:(p1 - q0) * w], rank + 1))
        ops.append(dist.P2POp(dist.irecv, X[(p1 - q0) * w:(p1 - q0 + 1) * w], rank + 1))
    return ops


def exchange_halo(dist, X, info, rank, world, width):
    if world == 1:
        return
    for r in dist.batch_isend_irecv(halo_ops(dist, X, info, rank, world, width)):
        r.wait()


def slab_exchange_lists(info, rank, world):
    """Ghost exchange of the slab partition in the library's terms (``Context.set_ghost_exchange``): per neighbouring rank
    the local node ids of the owned plane it keeps as ghosts (send) and of the ghost plane it owns (recv)."""
    npl, p0, p1, q0 = info["npl"], info["p0"], info["p1"], info["q0"]
    plane = lambda z: np.arange((z - q0) * npl, (z - q0 + 1) * npl, dtype=np.int32)  # noqa: E731
    ranks, send, recv = [], [], []
    if rank > 0:
        ranks.append(rank - 1); send.append(plane(p0)); recv.append(plane(p0 - 1))
    if rank < world - 1:
        ranks.append(rank + 1); send.append(plane(p1 - 1)); recv.append(plane(p1))
    return ranks, send, recv
