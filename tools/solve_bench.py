"""Solver-tuning helper: assembles the n^3 HEX8 Jacobian on the device and times a bounded PCG (ms per iteration, SpMV GB/s)."""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, nargs="+", default=[100, 200])
ap.add_argument("--iters", type=int, default=100)
a = ap.parse_args()
for n in a.n:
    lxyz, lconn, owned, con, se, info = bench.slab_problem(n, 0, 1, 0.2)
    ctx = bench.make_context(lxyz, lconn, owned, con, se, 0)
    X = torch.from_numpy(1e-3 * np.random.default_rng(1).uniform(-1, 1, len(lxyz) * 3)).cuda()
    R = torch.empty(len(lxyz) * 3, dtype=torch.float64, device="cuda")
    J = torch.empty(ctx.nnz, dtype=torch.float64, device="cuda")
    dx = torch.zeros_like(R)
    ctx.residual_and_jacobian(X, R, J)
    for rep in range(2):
        si = ctx.solve(J, -R, dx, rtol=1e-30, max_it=a.iters)
    per = si["solve_ms"] / max(1, si["iterations"])
    print(f"n={n} iters={si['iterations']} ms/iter={per:.3f} spmv_GBps={ctx.nnz * 8 / per / 1e6:.0f} rel_res={si['rel_residual']:.3e}", flush=True)
    del ctx, J
    torch.cuda.empty_cache()
