"""Parity AT the BASELINE.json sizes: sampled CSR rows / residual / sensitivity entries of the full-size device assembly
against the CPU oracle (tests/submesh.py), with element boxes at the start, middle and END of the mesh so that the value
offsets beyond 2^31 are exercised (C3 nv=6: nnz 7.8e9, C4: 4.09e9, C5: 9.2e9).  Through the C-ABI, device-resident
outputs (the J of C3 nv=6 alone is 62.5 GB)."""
import gc

import numpy as np
import pytest

import oracle_lib as O
import submesh as SM
from mast_b200 import mesh

pytestmark = pytest.mark.gpu


def _torch():
    import torch
    return torch


def _free():
    gc.collect()
    _torch().cuda.empty_cache()


def _cantilever_ctx(n, nv, xyz, conn, rho=None, **mat):
    import mast_b200
    ctx = mast_b200.Context(0)
    ctx.set_mesh(mast_b200.HEX8, xyz, conn)
    con = np.zeros((len(xyz), nv), np.uint8)
    con[np.isclose(xyz[:, 0], 0.0)] = 1
    ctx.set_dof_map(nv, con.ravel())
    ctx.set_material(E0=72e9, nu=0.33, alpha=1.5e-5, **mat)
    j = (np.arange(len(conn)) // n) % n
    se = np.flatnonzero(j == n - 1).astype(np.int32)
    ctx.set_side_pressure(se, np.full(len(se), 3, np.int32), np.full(len(se), 5e7))
    if rho is not None:
        ctx.set_density(rho)
    ctx.build_pattern()
    return ctx, con.ravel()


@pytest.mark.parametrize("nv", [3, 6])
def test_c3_200cubed_rows_match_oracle(built, nv):
    torch = _torch()
    n = 200
    xyz, conn = mesh.build_cube(n, n, n, 1.0, 1.0, 1.0, 8, jitter=0.2)
    ctx, con = _cantilever_ctx(n, nv, xyz, conn)
    assert ctx.nnz == (3 * n + 1) ** 3 * nv * nv
    X = 1e-3 * np.random.Generator(np.random.MT19937(1)).uniform(-1.0, 1.0, len(xyz) * nv)
    if nv == 6:
        X.reshape(-1, 6)[:, 3:] = 0.0
    dX = torch.from_numpy(X).cuda()
    dR = torch.empty(len(X), dtype=torch.float64, device="cuda")
    dJ = torch.empty(ctx.nnz, dtype=torch.float64, device="cuda")
    ctx.residual_and_jacobian(dX, dR, dJ)
    assert ctx.stats()["structured"] == 1
    bm = SM.BoxMesh(n, n, n)
    seen = 0
    for box in SM.default_boxes(bm):
        r = SM.check_rows(ctx, bm, box, xyz, conn, nv, X, J=dJ, R=dR, constrained=con, pressure_top=5e7)
        assert r["nodes"] >= 27 and r["J"] <= 1.0
        seen = max(seen, r["max_value_index"])
    assert seen == ctx.nnz  # the last sampled row ends at the last CSR value
    if nv == 6:
        assert seen > 2 ** 32
    del dJ, dR, dX, ctx
    _free()


def test_c4_256cubed_simp_fused_rows_match_oracle(built):
    """C4: 256^3 SIMP, R + J + adjoint density sensitivities of one state."""
    torch = _torch()
    n = 256
    xyz, conn = mesh.build_cube(n, n, n, 1.0, 1.0, 1.0, 8, jitter=0.2)
    rho = 0.2 + 0.6 * np.random.Generator(np.random.MT19937(3)).uniform(0.0, 1.0, len(xyz))
    drho = torch.from_numpy(rho).cuda()
    ctx, con = _cantilever_ctx(n, 3, xyz, conn, rho=drho, rho_min=1e-8, penalty=4.0)
    assert ctx.nnz == (3 * n + 1) ** 3 * 9 and ctx.nnz > 2 ** 31
    X = 1e-3 * np.random.Generator(np.random.MT19937(1)).uniform(-1.0, 1.0, len(xyz) * 3)
    dX = torch.from_numpy(X).cuda()
    dR = torch.empty(len(X), dtype=torch.float64, device="cuda")
    dG = torch.empty(len(xyz), dtype=torch.float64, device="cuda")
    dJ = torch.empty(ctx.nnz, dtype=torch.float64, device="cuda")
    ctx.sensitivity_sweep(dX, dX, dG, dR, dJ)
    bm = SM.BoxMesh(n, n, n)
    seen = 0
    for box in SM.default_boxes(bm):
        r = SM.check_rows(ctx, bm, box, xyz, conn, 3, X, J=dJ, R=dR, G=dG, constrained=con, pressure_top=5e7, rho=rho,
                          rho_min=1e-8, penalty=4.0)
        assert r["nodes"] >= 27
        seen = max(seen, r["max_value_index"])
    assert seen == ctx.nnz
    del dJ, dR, dG, dX, ctx
    _free()


def test_c5_hex27_2m_thermoelastic_rows_match_oracle(built):
    """C5: 200 x 100 x 100 HEX27 (2 M elements, nnz 9.2e9), thermoelastic load."""
    torch = _torch()
    import mast_b200
    nx, ny, nz = 200, 100, 100
    xyz, conn = mesh.build_cube(nx, ny, nz, 2.0, 1.0, 1.0, 27, jitter=0.2)
    con = np.zeros((len(xyz), 3), np.uint8)
    con[np.isclose(xyz[:, 0], 0.0)] = 1
    con = con.ravel()
    ctx = mast_b200.Context(0)
    ctx.set_mesh(mast_b200.HEX27, xyz, conn)
    ctx.set_dof_map(3, con)
    ctx.set_material(E0=72e9, nu=0.33, alpha=1.5e-5)
    ctx.set_temperature(1, None, 350.0, 300.0)
    ctx.build_pattern()
    assert ctx.nnz == (8 * nx + 1) * (8 * ny + 1) * (8 * nz + 1) * 9
    X = 1e-3 * np.random.Generator(np.random.MT19937(1)).uniform(-1.0, 1.0, len(xyz) * 3)
    dX = torch.from_numpy(X).cuda()
    dR = torch.empty(len(X), dtype=torch.float64, device="cuda")
    dJ = torch.empty(ctx.nnz, dtype=torch.float64, device="cuda")
    ctx.residual_and_jacobian(dX, dR, dJ)
    bm = SM.BoxMesh(nx, ny, nz, 27)
    seen = 0
    for box in SM.default_boxes(bm, w=2):
        r = SM.check_rows(ctx, bm, box, xyz, conn, 3, X, J=dJ, R=dR, constrained=con, alpha=1.5e-5, thermal=1,
                          temperature_const=350.0, ref_temperature=300.0)
        assert r["nodes"] >= 27  # 2 x 2 x 2 HEX27 elements: 3^3 interior nodes
        seen = max(seen, r["max_value_index"])
    assert seen == ctx.nnz and seen > 2 ** 33
    del dJ, dR, dX, ctx
    _free()
