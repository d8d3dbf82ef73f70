"""Driver for ncu: density sensitivities alone (structured sensitivity sweep) on an n^3 mesh."""
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
ctx.set_material(E0=72e9, nu=0.33, rho_min=1e-8, penalty=4.0)
rng = np.random.default_rng(3)
ctx.set_density(0.2 + 0.6 * rng.uniform(size=len(xyz)))
ctx.build_pattern()
X = torch.from_numpy(1e-3 * rng.uniform(-1, 1, len(xyz) * 3)).cuda()
G = torch.empty(len(xyz), dtype=torch.float64, device="cuda")
for _ in range(3):
    ctx.sensitivity_sweep(X, X, G)
torch.cuda.synchronize()
print(ctx.stats())
