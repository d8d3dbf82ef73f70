"""Multi-GPU host logic on CPU: two gloo ranks partition a HEX8 block into z-slabs (mast_b200.partition), build the
rank-local sparsity of their owned rows with planning-only contexts, exchange the ghost planes of X, and rank 0 checks
(i) the rank patterns tile the global pattern of the oracle bit-exactly, (ii) every element is counted once and ghost
elements are exactly one layer per cut, (iii) after the halo exchange each rank's localized X equals the global vector
on all its local nodes -- what the device sweep needs before it can produce its rows without any row exchange."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, nv, out):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import mast_b200
    from mast_b200 import partition
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lxyz, lconn, owned, con, se, info = partition.slab_problem(n, rank, world, 0.2, nv)
        ctx = mast_b200.Context(-1)
        ctx.set_mesh(mast_b200.HEX8, lxyz, lconn)
        ctx.set_dof_map(nv, con, owned)
        ctx.build_pattern()
        rp, ci = ctx.get_pattern()
        rows = ctx.get_row_dofs() + info["node_lo"] * nv       # global dof of each local row
        cols = ci.astype(np.int64) + info["node_lo"] * nv
        # localized solution: owned planes hold the global values, ghost planes start as NaN and arrive by halo exchange
        rng = np.random.Generator(np.random.MT19937(1))
        Xg = 1e-3 * rng.uniform(-1.0, 1.0, (n + 1) ** 3 * nv)
        lo = info["node_lo"] * nv
        Xl = Xg[lo:lo + len(lxyz) * nv].copy()
        ghost = np.repeat(owned == 0, nv)
        Xl[ghost] = np.nan
        X = torch.from_numpy(Xl)
        partition.exchange_halo(dist, X, info, rank, world, nv)
        ok_halo = bool(np.array_equal(X.numpy(), Xg[lo:lo + len(lxyz) * nv]))
        res = dict(rank=rank, rp=rp, rows=rows, cols=cols, ok_halo=ok_halo, n_owned_elem=info["n_owned_elem"],
                   n_local_elem=len(lconn), n_ghost_nodes=int((owned == 0).sum()), npl=info["npl"])
        gathered = [None] * world
        dist.all_gather_object(gathered, res)
        # scalar reduction used for ||R|| and the compliance
        t = torch.tensor([float(info["n_owned_elem"])], dtype=torch.float64)
        dist.all_reduce(t)
        if rank == 0:
            torch.save(dict(parts=gathered, n_elem_sum=float(t.item())), out)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 6), (3, 7)])
def test_slab_partition_tiles_the_global_problem(built, tmp_path, world, n):
    import oracle_lib as O
    nv = 3
    out = str(tmp_path / "parts.pt")
    mp.spawn(_worker, args=(world, _free_port(), n, nv, out), nprocs=world, join=True)
    data = torch.load(out, weights_only=False)
    parts = sorted(data["parts"], key=lambda r: r["rank"])
    assert data["n_elem_sum"] == n ** 3 and sum(p["n_owned_elem"] for p in parts) == n ** 3
    prob = O.cantilever(n, n, n, 1.0, 1.0, 1.0, nv=nv, jitter=0.2)
    orp, oci = prob.pattern()
    seen = np.zeros(prob.n_dof, bool)
    for p in parts:
        assert p["ok_halo"], f"rank {p['rank']}: localized X differs from the global vector after the halo exchange"
        for r, g in enumerate(p["rows"]):
            assert not seen[g]
            seen[g] = True
            a, b = p["rp"][r], p["rp"][r + 1]
            assert np.array_equal(p["cols"][a:b], oci[orp[g]:orp[g + 1]])   # same columns, same order
    assert seen.all()                                                        # every global row has exactly one owner
    # ghost elements: one layer per cut side; ghost nodes: one plane per cut side
    cuts = [int(r > 0) + int(r < world - 1) for r in range(world)]
    for p, c in zip(parts, cuts):
        assert p["n_local_elem"] - p["n_owned_elem"] in (c * n * n, (c - 1) * n * n)  # the last rank owns no element of the top plane
        assert p["n_ghost_nodes"] == c * p["npl"]
