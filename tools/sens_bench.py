"""Density sensitivities alone (M2) on an n^3 build_cube mesh: structured sweep vs the element-parallel kernels."""
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
ap.add_argument("--n", type=int, nargs="+", default=[200])
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--adjoint", action="store_true", help="a separate adjoint vector instead of the compliance case adj == X")
a = ap.parse_args()
for n in a.n:
    xyz, conn = mesh.build_cube(n, n, n, 1.0, 1.0, 1.0, 8, jitter=0.2)
    ctx = mast_b200.Context(0)
    ctx.set_mesh(mast_b200.HEX8, xyz, conn)
    ctx.set_dof_map(3, mesh.dirichlet_flags(len(xyz), 3, mesh.boundary_nodes(n, n, n, 4)))
    ctx.set_material(E0=72e9, nu=0.33, rho_min=1e-8, penalty=4.0)
    rng = np.random.default_rng(3)
    ctx.set_density(0.2 + 0.6 * rng.uniform(size=len(xyz)))
    ctx.build_pattern()
    X = torch.from_numpy(1e-3 * rng.uniform(-1, 1, len(xyz) * 3)).cuda()
    L = torch.from_numpy(1e-3 * rng.uniform(-1, 1, len(xyz) * 3)).cuda() if a.adjoint else X
    G = torch.empty(len(xyz), dtype=torch.float64, device="cuda")
    ms = []
    for _ in range(a.reps):
        ctx.sensitivity_sweep(X, L, G)
        ms.append(ctx.stats()["last_sweep_ms"])
    best = min(ms[1:])
    print(f"n={n} SENS_STRUCT={os.environ.get('MASTB200_SENS_STRUCT', '1')} adjoint={'separate' if a.adjoint else 'state'} "
          f"ms={best:.3f} Gelem/s={n ** 3 / best / 1e6:.3f} fp64_frac={n ** 3 * 5100 / best / 1e9 / 36.7:.3f} |G|max={float(G.abs().max()):.4e}",
          flush=True)
    del ctx
    torch.cuda.empty_cache()
