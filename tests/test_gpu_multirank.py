"""Two NCCL ranks on two GPUs of one box (mp.spawn): the ranks' concatenated CSR rows / residuals equal the oracle's
assembly of the whole mesh, the in-library ghost exchange (mastb200_set_ghost_exchange, auto_localize) replaces the
caller's halo, the summed compliance equals the serial one and the distributed solve equals the serial solve.
Skipped on boxes with fewer than two GPUs (the driver's 1-GPU test box)."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, n, port, outdir):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import mast_b200
    from mast_b200 import partition
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    nv = 3
    lxyz, lconn, owned, con, se, info = partition.slab_problem(n, rank, world, 0.2, nv)
    ctx = mast_b200.Context(rank)
    ctx.set_mesh(mast_b200.HEX8, lxyz, lconn)
    ctx.set_dof_map(nv, con, owned if not owned.all() else None)
    ctx.set_material(E0=72e9, nu=0.33)
    ctx.set_side_pressure(se, np.full(len(se), 3, np.int32), np.full(len(se), 5e7))
    ctx.build_pattern()
    # the library's own communicator: rank 0 makes the id, torch broadcasts the bytes
    uid = torch.from_numpy(mast_b200.Context.comm_unique_id() if rank == 0 else np.zeros(128, np.uint8)).cuda()
    dist.broadcast(uid, 0)
    ctx.comm_init(rank, world, uid.cpu().numpy())
    ranks, send, recv = partition.slab_exchange_lists(info, rank, world)
    ctx.set_ghost_exchange(ranks, send, recv, auto_localize=True)
    npl = info["npl"]
    rng = np.random.Generator(np.random.MT19937(1))
    Xg = 1e-3 * rng.uniform(-1.0, 1.0, (n + 1) ** 3 * nv)
    Xl = Xg[info["q0"] * npl * nv:(info["q1"] + 1) * npl * nv].copy()
    own_dofs = np.repeat(owned.astype(bool), nv)
    Xl[~own_dofs] = np.nan                      # ghosts must come from the neighbour
    dX = torch.from_numpy(Xl).cuda()
    dR = torch.empty(len(Xl), dtype=torch.float64, device="cuda")
    dJ = torch.empty(ctx.nnz, dtype=torch.float64, device="cuda")
    ctx.residual_and_jacobian(dX, dR, dJ)       # localizes X itself
    assert not torch.isnan(dX).any()
    q = ctx.all_reduce_sum([ctx.compliance(dX)])[0]
    nrm2 = ctx.all_reduce_sum([ctx.res_l2_norm() ** 2])[0]
    # distributed Newton step from X = 0
    Z = torch.zeros(len(Xl), dtype=torch.float64, device="cuda")
    ctx.residual_and_jacobian(Z, dR, dJ)
    dx = torch.zeros(len(Xl), dtype=torch.float64, device="cuda")
    sinfo = ctx.solve(dJ, -dR, dx, rtol=1e-10, max_it=20000)
    ctx.residual_and_jacobian(dX, dR, dJ)
    rows = ctx.get_row_dofs()
    np.savez(os.path.join(outdir, f"rank{rank}.npz"), J=dJ.cpu().numpy(), R=dR.cpu().numpy()[rows], rows=rows + info["node_lo"] * nv,
             q=q, nrm2=nrm2, dx=dx.cpu().numpy()[rows], iters=sinfo["iterations"], conv=sinfo["converged"])
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_match_the_oracle(built, tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    import oracle_lib as O
    n, world = 12, 2
    mp.spawn(_worker, args=(world, n, 29611, str(tmp_path)), nprocs=world, join=True)
    prob = O.cantilever(n, n, n, 1.0, 1.0, 1.0, nv=3, jitter=0.2)
    Xg = 1e-3 * np.random.Generator(np.random.MT19937(1)).uniform(-1.0, 1.0, prob.n_dof)
    Ro, Jo = prob.residual_and_jacobian(Xg)
    rp, _ = prob.pattern()
    parts = [np.load(os.path.join(tmp_path, f"rank{r}.npz")) for r in range(world)]
    rows = np.concatenate([p["rows"] for p in parts])
    assert np.array_equal(rows, np.arange(prob.n_dof))       # the ranks' row blocks tile the system in order
    J = np.concatenate([p["J"] for p in parts]); R = np.concatenate([p["R"] for p in parts])
    from submesh import close_entries
    close_entries(J, Jo, what="J over two ranks"); close_entries(R, Ro, what="R over two ranks")
    assert np.isclose(parts[0]["q"], prob.compliance(Xg), rtol=1e-12) and parts[0]["q"] == parts[1]["q"]
    assert np.isclose(parts[0]["nrm2"], Ro @ Ro, rtol=1e-12)
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    R0, J0 = prob.residual_and_jacobian(np.zeros(prob.n_dof))
    A = sp.csr_matrix((J0, prob.pattern()[1], rp), shape=(prob.n_dof, prob.n_dof))
    ref = spla.spsolve(A.tocsc(), -R0)
    dx = np.concatenate([p["dx"] for p in parts])
    assert all(int(p["conv"]) == 1 for p in parts) and np.abs(dx - ref).max() <= 1e-7 * np.abs(ref).max()
