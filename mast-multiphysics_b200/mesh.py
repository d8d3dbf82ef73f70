"""Structured synthetic meshes with libMesh ``MeshTools::Generation`` numbering.

``build_cube(nx,ny,nz,...)``: node id ``i + j(nx+1) + k(nx+1)(ny+1)`` (HEX27: ``2n+1`` nodes per
direction), element id ``i + j nx + k nx ny``, libMesh HEX8/HEX27 node ordering, boundary ids
0: z-, 1: y-, 2: x+, 3: y+, 4: x-, 5: z+ (side number == boundary id for build_cube).
Host-side test/bench scaffolding only; it plays the role of the caller's libMesh mesh.
"""
import numpy as np

_H8 = [(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)]
_H27 = [(0, 0, 0), (2, 0, 0), (2, 2, 0), (0, 2, 0), (0, 0, 2), (2, 0, 2), (2, 2, 2), (0, 2, 2),
        (1, 0, 0), (2, 1, 0), (1, 2, 0), (0, 1, 0), (0, 0, 1), (2, 0, 1), (2, 2, 1), (0, 2, 1),
        (1, 0, 2), (2, 1, 2), (1, 2, 2), (0, 1, 2), (1, 1, 0), (1, 0, 1), (2, 1, 1), (1, 2, 1),
        (0, 1, 1), (1, 1, 2), (1, 1, 1)]


def build_cube(nx, ny, nz, lx=1.0, ly=1.0, lz=1.0, elem_type=8, jitter=0.0, seed=12345):
    """Returns xyz [n_nodes,3] float64, conn [n_elem,nn] int32."""
    m = 1 if elem_type == 8 else 2
    npx, npy, npz = m * nx + 1, m * ny + 1, m * nz + 1
    x = np.linspace(0.0, lx, npx); y = np.linspace(0.0, ly, npy); z = np.linspace(0.0, lz, npz)
    Z, Y, X = np.meshgrid(z, y, x, indexing="ij")
    xyz = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1).astype(np.float64)
    if jitter > 0.0:  # perturb interior nodes so that kernels cannot rely on uniform geometry
        rng = np.random.Generator(np.random.MT19937(seed))
        h = np.array([lx / (npx - 1), ly / (npy - 1), lz / (npz - 1)])
        d = jitter * h * rng.uniform(-1.0, 1.0, size=xyz.shape)
        K, J, I = np.meshgrid(np.arange(npz), np.arange(npy), np.arange(npx), indexing="ij")
        interior = ((I > 0) & (I < npx - 1) & (J > 0) & (J < npy - 1) & (K > 0) & (K < npz - 1)).ravel()
        xyz[interior] += d[interior]
    k, j, i = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    i = i.ravel(); j = j.ravel(); k = k.ravel()
    offs = _H8 if elem_type == 8 else _H27
    conn = np.empty((nx * ny * nz, len(offs)), np.int32)
    for a, (di, dj, dk) in enumerate(offs):
        conn[:, a] = (m * i + di) + (m * j + dj) * npx + (m * k + dk) * npx * npy
    return xyz, conn


def build_square(nx, ny, lx=1.0, ly=1.0):
    """QUAD4 mesh in the xy-plane: node id i + j(nx+1), element id i + j nx."""
    x = np.linspace(0.0, lx, nx + 1); y = np.linspace(0.0, ly, ny + 1)
    Y, X = np.meshgrid(y, x, indexing="ij")
    xyz = np.stack([X.ravel(), Y.ravel(), np.zeros(X.size)], axis=1).astype(np.float64)
    j, i = np.meshgrid(np.arange(ny), np.arange(nx), indexing="ij")
    i = i.ravel(); j = j.ravel()
    n0 = i + j * (nx + 1)
    conn = np.stack([n0, n0 + 1, n0 + nx + 2, n0 + nx + 1], axis=1).astype(np.int32)
    return xyz, conn


def boundary_sides(nx, ny, nz, bid):
    """(element ids, side ids) of build_cube boundary ``bid``."""
    k, j, i = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    sel = {0: k == 0, 1: j == 0, 2: i == nx - 1, 3: j == ny - 1, 4: i == 0, 5: k == nz - 1}[bid]
    e = (i + j * nx + k * nx * ny)[sel].astype(np.int32)
    return e, np.full(e.shape, bid, np.int32)


def boundary_nodes(nx, ny, nz, bid, elem_type=8):
    m = 1 if elem_type == 8 else 2
    npx, npy, npz = m * nx + 1, m * ny + 1, m * nz + 1
    K, J, I = np.meshgrid(np.arange(npz), np.arange(npy), np.arange(npx), indexing="ij")
    sel = {0: K == 0, 1: J == 0, 2: I == npx - 1, 3: J == npy - 1, 4: I == 0, 5: K == npz - 1}[bid]
    return (I + J * npx + K * npx * npy)[sel].astype(np.int64)


def dirichlet_flags(n_nodes, nv, nodes, vars_=None):
    c = np.zeros((n_nodes, nv), np.uint8)
    c[np.asarray(nodes)[:, None], np.arange(nv) if vars_ is None else np.asarray(vars_)] = 1
    return c.ravel()
