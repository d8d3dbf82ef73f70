"""Small driver for ncu: one 64^3 (default) HEX8 R+J sweep with device-resident buffers."""
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
ap.add_argument("--n", type=int, default=64)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--nv", type=int, default=3)
a = ap.parse_args()
lxyz, lconn, owned, con, se, info = bench.slab_problem(a.n, 0, 1, 0.2, a.nv)
ctx = bench.make_context(lxyz, lconn, owned, con, se, 0, a.nv)
X = torch.from_numpy(1e-3 * np.random.default_rng(1).uniform(-1, 1, len(lxyz) * a.nv)).cuda()
R = torch.empty(len(lxyz) * a.nv, dtype=torch.float64, device="cuda")
J = torch.empty(ctx.nnz, dtype=torch.float64, device="cuda")
for _ in range(a.reps):
    ctx.residual_and_jacobian(X, R, J)
torch.cuda.synchronize()
print(ctx.stats())
