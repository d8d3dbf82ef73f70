"""HEX27 sweep: time with and without the Jacobian scatter (R only = geometry + DMMA contraction + f = K u)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import mast_b200
from mast_b200 import mesh
nx, ny, nz = 100, 50, 50
xyz, conn = mesh.build_cube(nx, ny, nz, 2.0, 1.0, 1.0, 27, jitter=0.1)
ctx = mast_b200.Context(0)
ctx.set_mesh(mast_b200.HEX27, xyz, conn)
ctx.set_dof_map(3, mesh.dirichlet_flags(len(xyz), 3, mesh.boundary_nodes(nx, ny, nz, 4, 27)))
ctx.set_material(E0=72e9, nu=0.33)
ctx.build_pattern()
X = torch.from_numpy(1e-3 * np.random.default_rng(1).uniform(-1, 1, len(xyz) * 3)).cuda()
R = torch.empty(len(xyz) * 3, dtype=torch.float64, device="cuda")
J = torch.empty(ctx.nnz, dtype=torch.float64, device="cuda")
for name, args in (("R+J", (X, R, J)), ("R only", (X, R, None)), ("J only", (X, None, J))):
    ms = []
    for _ in range(3):
        ctx.residual_and_jacobian(*args); ms.append(ctx.stats()["last_main_kernel_ms"])
    print(name, round(min(ms[1:]), 3), "ms", flush=True)
