"""Kernel-tuning helper: times the device-resident R+J sweep (CUDA events inside the library) on an n^3 mesh."""
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
ap.add_argument("--reps", type=int, default=5)
a = ap.parse_args()
for n in a.n:
    lxyz, lconn, owned, con, se, info = bench.slab_problem(n, 0, 1, 0.2)
    ctx = bench.make_context(lxyz, lconn, owned, con, se, 0)
    X = torch.from_numpy(1e-3 * np.random.default_rng(1).uniform(-1, 1, len(lxyz) * 3)).cuda()
    R = torch.empty(len(lxyz) * 3, dtype=torch.float64, device="cuda")
    J = torch.empty(ctx.nnz, dtype=torch.float64, device="cuda")
    ms = []
    for _ in range(a.reps):
        ctx.residual_and_jacobian(X, R, J)
        ms.append(ctx.stats()["last_main_kernel_ms"])
    s = ctx.stats()
    best = min(ms[1:])
    print(f"n={n} lib={os.environ.get('MASTB200_LIB', 'default')} kernel_ms={best:.3f} Melem/s={n ** 3 / best / 1e3:.1f} "
          f"visits/elem={s['n_visits'] / s['n_elem']:.3f} tiles={s['n_tiles']} |J|max={float(J.abs().max()):.4e}", flush=True)
    del ctx, J
    torch.cuda.empty_cache()
