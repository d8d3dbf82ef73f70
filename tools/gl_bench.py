"""HEX8 Green-Lagrange R+J on an n^3 build_cube mesh: slot-kernel tangent + matrix-free residual vs the row-owner gather."""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mast_b200  # noqa: E402
from mast_b200 import mesh  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, nargs="+", default=[100])
ap.add_argument("--reps", type=int, default=4)
a = ap.parse_args()
for n in a.n:
    xyz, conn = mesh.build_cube(n, n, n, 1.0, 1.0, 1.0, 8, jitter=0.2)
    ctx = mast_b200.Context(0)
    ctx.set_mesh(mast_b200.HEX8, xyz, conn)
    ctx.set_dof_map(3, mesh.dirichlet_flags(len(xyz), 3, mesh.boundary_nodes(n, n, n, 4)))
    ctx.set_material(E0=72e9, nu=0.33, nonlinear_strain=1)
    ctx.build_pattern()
    X = torch.from_numpy(2e-3 * np.random.default_rng(1).uniform(-1, 1, len(xyz) * 3)).cuda()
    R = torch.empty(len(xyz) * 3, dtype=torch.float64, device="cuda")
    J = torch.empty(ctx.nnz, dtype=torch.float64, device="cuda")
    sw, km = [], []
    for _ in range(a.reps):
        ctx.residual_and_jacobian(X, R, J)
        s = ctx.stats()
        sw.append(s["last_sweep_ms"]); km.append(s["last_main_kernel_ms"])
    print(f"n={n} GL_STRUCT={os.environ.get('MASTB200_GL_STRUCT', '1')} structured={s['structured']} sweep_ms={min(sw[1:]):.3f} "
          f"tangent_kernel_ms={min(km[1:]):.3f} Melem/s={n ** 3 / min(sw[1:]) / 1e3:.1f}", flush=True)
    del ctx, J
    torch.cuda.empty_cache()
