"""Times the parity-test configurations of BASELINE.json other than the bench.py headline (C3) on one B200:
C1 (20x10x10 HEX8, nv=6 reference layout), C2 (1024^2 QUAD4 heat, k(T)), C4 (256^3 HEX8 SIMP: compliance gradient fused,
with and without J), C5 (HEX27 + thermoelastic load).  Device-resident buffers, CUDA-event kernel times from the library.
One JSON line per configuration; results are copied into profiles/ by hand.

    python tools/bench_configs.py [--only c2 c5] [--c5 200 100 100] [--c4 256]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mast_b200  # noqa: E402
from mast_b200 import mesh  # noqa: E402

HBM = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def mt(seed, n, lo=0.0, hi=1.0):
    return lo + (hi - lo) * np.random.Generator(np.random.MT19937(seed)).uniform(0.0, 1.0, n)


def timed(ctx, fn, reps, key="last_main_kernel_ms"):
    ms = []
    for _ in range(reps):
        fn()
        ms.append(ctx.stats()[key])
    return float(np.min(ms[1:])), ms


def report(name, n_elem, ms, alg_bytes, flops=None, **extra):
    out = {"config": name, "elements": int(n_elem), "kernel_ms": ms, "elements_per_s": n_elem / ms * 1e3,
           "algorithmic_GBps": alg_bytes / ms / 1e6, "hbm_frac": alg_bytes / ms / 1e6 / HBM}
    if flops:
        out["executed_tflops"] = flops / ms / 1e9
    out.update(extra)
    print(json.dumps(out), flush=True)


def c1(a):
    xyz, conn = mesh.build_cube(20, 10, 10, 0.3, 0.15, 0.15, 8)
    nv = 6
    ctx = mast_b200.Context(0)
    ctx.set_mesh(mast_b200.HEX8, xyz, conn)
    ctx.set_dof_map(nv, mesh.dirichlet_flags(len(xyz), nv, mesh.boundary_nodes(20, 10, 10, 4)))
    ctx.set_material(E0=72e9, nu=0.33)
    se, sid = mesh.boundary_sides(20, 10, 10, 3)
    ctx.set_side_pressure(se, sid, np.full(len(se), 5e7))
    ctx.build_pattern()
    X = torch.from_numpy(1e-3 * mt(1, len(xyz) * nv, -1, 1)).cuda()
    R = torch.empty(len(xyz) * nv, dtype=torch.float64, device="cuda")
    J = torch.empty(ctx.nnz, dtype=torch.float64, device="cuda")
    ms, _ = timed(ctx, lambda: ctx.residual_and_jacobian(X, R, J), a.reps)
    report("C1 20x10x10 HEX8 nv=6 R+J", len(conn), ms, ctx.nnz * 8 + len(xyz) * (nv * 16 + 24) + len(conn) * 32)


def c2(a):
    n = 1024
    xyz, conn = mesh.build_square(n, n)
    ctx = mast_b200.Context(0)
    ctx.set_mesh(mast_b200.QUAD4, xyz, conn)
    ctx.set_dof_map(1, np.isclose(xyz[:, 0], 0.0).astype(np.uint8))
    ctx.set_material(k0=237.0, k_beta=1e-3, k_of_T=1, thickness=1.0)
    ctx.build_pattern()
    T = torch.from_numpy(mt(4, len(xyz), 300, 350)).cuda()
    R = torch.empty(len(xyz), dtype=torch.float64, device="cuda")
    J = torch.empty(ctx.nnz, dtype=torch.float64, device="cuda")
    ms, _ = timed(ctx, lambda: ctx.residual_and_jacobian(T, R, J), a.reps)
    report("C2 1024^2 QUAD4 heat k(T) R+J", len(conn), ms, ctx.nnz * 8 + len(xyz) * (8 + 8 + 16) + len(conn) * 16,
           note="126 MB working set: L2-resident")


def c4(a):
    n = a.c4
    xyz, conn = mesh.build_cube(n, n, n, 1.0, 1.0, 1.0, 8, jitter=a.jitter)
    nv = 3
    ctx = mast_b200.Context(0)
    ctx.set_mesh(mast_b200.HEX8, xyz, conn)
    ctx.set_dof_map(nv, mesh.dirichlet_flags(len(xyz), nv, mesh.boundary_nodes(n, n, n, 4)))
    ctx.set_material(E0=72e9, nu=0.33, rho_min=1e-8, penalty=4.0)
    se, sid = mesh.boundary_sides(n, n, n, 3)
    ctx.set_side_pressure(se, sid, np.full(len(se), 5e7))
    ctx.set_density(mt(3, len(xyz), 0.2, 0.8))
    t0 = time.time()
    ctx.build_pattern()
    plan_s = time.time() - t0
    n_nodes = len(xyz)
    X = torch.from_numpy(1e-3 * mt(1, n_nodes * nv, -1, 1)).cuda()
    G = torch.empty(n_nodes, dtype=torch.float64, device="cuda")
    # sensitivity-only sweep (compliance: adjoint = X)
    ms, _ = timed(ctx, lambda: ctx.sensitivity_sweep(X, X, G), a.reps)
    report(f"C4 {n}^3 HEX8 SIMP adjoint density sensitivities only (M2)", len(conn), ms,
           n_nodes * (24 + 24 + 8 + 24 + 8) + len(conn) * 32, plan_build_s=round(plan_s, 1),
           visits_per_elem=ctx.stats()["n_visits"] / len(conn))
    q = ctx.compliance(X)
    R = torch.empty(n_nodes * nv, dtype=torch.float64, device="cuda")
    try:
        J = torch.empty(ctx.nnz, dtype=torch.float64, device="cuda")
    except torch.OutOfMemoryError:
        print(json.dumps({"config": f"C4 {n}^3 fused", "skipped": "CSR values do not fit"}))
        return
    ms, _ = timed(ctx, lambda: ctx.sensitivity_sweep(X, X, G, R, J), a.reps)
    report(f"C4 {n}^3 HEX8 SIMP R+J+G fused", len(conn), ms,
           ctx.nnz * 8 + n_nodes * (24 * 4 + 8 + 8) + len(conn) * 32, compliance=q)


def c5(a):
    nx, ny, nz = a.c5
    t0 = time.time()
    xyz, conn = mesh.build_cube(nx, ny, nz, 2.0, 1.0, 1.0, 27, jitter=a.jitter27)
    nv = 3
    ctx = mast_b200.Context(0)
    ctx.set_mesh(mast_b200.HEX27, xyz, conn)
    ctx.set_dof_map(nv, mesh.dirichlet_flags(len(xyz), nv, mesh.boundary_nodes(nx, ny, nz, 4, 27)))
    ctx.set_material(E0=72e9, nu=0.33, alpha=1.5e-5)
    ctx.set_temperature(1, None, 350.0, 300.0)
    ctx.build_pattern()
    plan_s = time.time() - t0
    n_nodes = len(xyz)
    X = torch.from_numpy(1e-3 * mt(1, n_nodes * nv, -1, 1)).cuda()
    R = torch.empty(n_nodes * nv, dtype=torch.float64, device="cuda")
    J = torch.empty(ctx.nnz, dtype=torch.float64, device="cuda")
    ms, all_ms = timed(ctx, lambda: ctx.residual_and_jacobian(X, R, J), a.reps)
    s = ctx.stats()
    report(f"C5 {nx}x{ny}x{nz} HEX27 thermoelastic R+J", len(conn), ms,
           ctx.nnz * 8 + n_nodes * (24 * 3) + len(conn) * 27 * 4, flops=len(conn) * 1008 * 512.0,
           nnz=int(ctx.nnz), sweep_ms_incl_memset=s["last_sweep_ms"], setup_s=round(plan_s, 1),
           fp64_peak_tflops=36.7)


def gl(a):
    # HEX8 Green-Lagrange strain, 100^3
    n = a.gl
    xyz, conn = mesh.build_cube(n, n, n, 1.0, 1.0, 1.0, 8, jitter=a.jitter)
    nv = 3
    ctx = mast_b200.Context(0)
    ctx.set_mesh(mast_b200.HEX8, xyz, conn)
    ctx.set_dof_map(nv, mesh.dirichlet_flags(len(xyz), nv, mesh.boundary_nodes(n, n, n, 4)))
    ctx.set_material(E0=72e9, nu=0.33, nonlinear_strain=1)
    ctx.build_pattern()
    X = torch.from_numpy(1e-3 * mt(1, len(xyz) * nv, -1, 1)).cuda()
    R = torch.empty(len(xyz) * nv, dtype=torch.float64, device="cuda")
    J = torch.empty(ctx.nnz, dtype=torch.float64, device="cuda")
    ms, _ = timed(ctx, lambda: ctx.residual_and_jacobian(X, R, J), a.reps, "last_sweep_ms")  # tangent + residual kernels
    report(f"HEX8 {n}^3 Green-Lagrange R+J (slot-kernel tangent + nodal force kernels; MASTB200_GL_STRUCT=0: gather)", len(conn), ms,
           ctx.nnz * 8 + len(xyz) * (nv * 16 + 24) + len(conn) * 32)


def plate(a):
    # QUAD4 Mindlin plate, 1024^2, nv = 6
    n = a.plate
    xyz, conn = mesh.build_square(n, n)
    ctx = mast_b200.Context(0)
    ctx.set_mesh(mast_b200.QUAD4, xyz, conn)
    ctx.set_dof_map(6, mesh.dirichlet_flags(len(xyz), 6, np.flatnonzero(xyz[:, 0] == 0.0)))
    ctx.set_material(E0=72e9, nu=0.33, thickness=0.06, offset=0.0, kappa=5.0 / 6.0)
    ctx.build_pattern()
    X = torch.from_numpy(1e-3 * mt(1, len(xyz) * 6, -1, 1)).cuda()
    R = torch.empty(len(xyz) * 6, dtype=torch.float64, device="cuda")
    J = torch.empty(ctx.nnz, dtype=torch.float64, device="cuda")
    ms, _ = timed(ctx, lambda: ctx.residual_and_jacobian(X, R, J), a.reps)
    report(f"QUAD4 {n}^2 Mindlin plate nv=6 R+J", len(conn), ms, ctx.nnz * 8 + len(xyz) * (6 * 16 + 24) + len(conn) * 16)


def stress(a):
    # C5's output: von Mises p-norm aggregate and its state derivative (the adjoint right-hand side) on HEX27
    nx, ny, nz = a.c5s
    xyz, conn = mesh.build_cube(nx, ny, nz, 2.0, 1.0, 1.0, 27, jitter=a.jitter27)
    nv = 3
    ctx = mast_b200.Context(0)
    ctx.set_mesh(mast_b200.HEX27, xyz, conn)
    ctx.set_dof_map(nv, mesh.dirichlet_flags(len(xyz), nv, mesh.boundary_nodes(nx, ny, nz, 4, 27)))
    ctx.set_material(E0=72e9, nu=0.33)
    ctx.build_pattern()
    n_nodes = len(xyz)
    X = torch.from_numpy(1e-3 * mt(1, n_nodes * nv, -1, 1)).cuda()
    dq = torch.empty(n_nodes * nv, dtype=torch.float64, device="cuda")
    q, sums = ctx.stress_functional(X, 8.0, 2.0, 1.0e8)
    t = []
    for _ in range(a.reps):
        ctx.stress_functional(X, 8.0, 2.0, 1.0e8)
        t.append(ctx.stats()["last_sweep_ms"])
    ms = float(np.min(t[1:]))
    report(f"C5-output {nx}x{ny}x{nz} HEX27 von Mises p-norm aggregate", len(conn), ms,
           n_nodes * 48 + len(conn) * 27 * 4, q=q)
    t = []
    for _ in range(a.reps):
        ctx.stress_functional_derivative(X, 8.0, 2.0, 1.0e8, sums, dq)
        t.append(ctx.stats()["last_sweep_ms"])
    ms = float(np.min(t[1:]))
    report(f"C5-output {nx}x{ny}x{nz} HEX27 aggregate state derivative (adjoint rhs)", len(conn), ms,
           n_nodes * 72 + len(conn) * 27 * (4 + 2 * 72))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", nargs="*", default=["c1", "c2", "c4", "c5"])
    ap.add_argument("--reps", type=int, default=4)
    ap.add_argument("--c4", type=int, default=256)
    ap.add_argument("--c5", type=int, nargs=3, default=[200, 100, 100])
    ap.add_argument("--gl", type=int, default=100)
    ap.add_argument("--plate", type=int, default=1024)
    ap.add_argument("--c5s", type=int, nargs=3, default=[100, 50, 50])
    ap.add_argument("--jitter", type=float, default=0.2)
    ap.add_argument("--jitter27", type=float, default=0.1)
    a = ap.parse_args()
    for name in a.only:
        {"c1": c1, "c2": c2, "c4": c4, "c5": c5, "gl": gl, "plate": plate, "stress": stress}[name](a)
        torch.cuda.empty_cache()
