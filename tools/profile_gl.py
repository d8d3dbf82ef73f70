"""Driver for ncu: HEX8 Green-Lagrange tangent (slot kernel) on an n^3 mesh."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import mast_b200
from mast_b200 import mesh
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
xyz, conn = mesh.build_cube(n, n, n, 1.0, 1.0, 1.0, 8, jitter=0.2)
ctx = mast_b200.Context(0)
ctx.set_mesh(mast_b200.HEX8, xyz, conn)
ctx.set_dof_map(3, mesh.dirichlet_flags(len(xyz), 3, mesh.boundary_nodes(n, n, n, 4)))
ctx.set_material(E0=72e9, nu=0.33, nonlinear_strain=1)
ctx.build_pattern()
X = torch.from_numpy(2e-3 * np.random.default_rng(1).uniform(-1, 1, len(xyz) * 3)).cuda()
J = torch.empty(ctx.nnz, dtype=torch.float64, device="cuda")
for _ in range(3):
    ctx.residual_and_jacobian(X, None, J)
torch.cuda.synchronize()
print(ctx.stats())
