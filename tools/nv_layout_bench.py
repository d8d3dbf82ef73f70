import sys, os, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch
import mast_b200
from mast_b200 import mesh
for nv in (3, 6):
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    xyz, conn = mesh.build_cube(n, n, n, 1.0, 1.0, 1.0, 8, jitter=0.2)
    ctx = mast_b200.Context(0)
    ctx.set_mesh(mast_b200.HEX8, xyz, conn)
    ctx.set_dof_map(nv, mesh.dirichlet_flags(len(xyz), nv, mesh.boundary_nodes(n, n, n, 4)))
    ctx.set_material(E0=72e9, nu=0.33)
    t0 = time.time(); ctx.build_pattern(); tb = time.time() - t0
    X = torch.from_numpy(1e-3 * np.random.default_rng(1).uniform(-1, 1, len(xyz) * nv)).cuda()
    R = torch.empty(len(xyz) * nv, dtype=torch.float64, device="cuda")
    J = torch.empty(ctx.nnz, dtype=torch.float64, device="cuda")
    ms = []
    for _ in range(4):
        ctx.residual_and_jacobian(X, R, J); ms.append(ctx.stats()["last_main_kernel_ms"])
    st = ctx.stats()
    print(nv, "structured", st["structured"], "ms", min(ms[1:]), "Melem/s", n**3 / min(ms[1:]) / 1e3, "GB/s", ctx.nnz * 8 / min(ms[1:]) / 1e6, "plan_s", round(tb, 1), flush=True)
    del ctx, J
    torch.cuda.empty_cache()
