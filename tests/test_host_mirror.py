"""The host-side C++ mirror of MAST's assembly classes (mast-multiphysics_b200/host/): built over the C-ABI and the
libMesh shim, driven by tests/host/assembly_driver.cpp with the call sequence of a MAST example, compared with the oracle.

CPU part: the mirror compiles and links, and without a GPU the first sweep raises libMesh::LogicError (no fallback).
GPU part (-m gpu): C1 cantilever in the reference's 6-variable layout, SIMP compliance + adjoint sensitivities through
calculate_output_adjoint_sensitivity_multiple_parameters_no_direct, QUAD4 heat with k(T), HEX27 thermoelastic."""
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as O
from mast_b200 import mesh

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "tests", "host", "assembly_driver")


def _run(case, d):
    return subprocess.run([DRIVER, case, str(d)], capture_output=True, text=True)


def _scalars(d):
    return {k: float(v) for k, v in (l.split() for l in open(os.path.join(d, "scalars.txt")))}


def _close(a, b, rtol=1e-12):
    assert np.abs(a - b).max() <= rtol * np.abs(b).max()


def test_mirror_builds_and_fails_loudly_without_gpu(built, tmp_path):
    import torch
    assert os.path.exists(DRIVER)
    assert os.path.exists(os.path.join(ROOT, "mast-multiphysics_b200", "libmast_b200_host.so"))
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu tests")
    np.zeros(2541 * 6).tofile(tmp_path / "X.bin")
    r = _run("c1", tmp_path)
    assert r.returncode == 3 and "libMesh::LogicError" in r.stderr and "no CPU fallback" in r.stderr


def _compare_common(prob, d, X):
    rp = np.fromfile(os.path.join(d, "row_ptr.bin"), np.int64); ci = np.fromfile(os.path.join(d, "col_idx.bin"), np.int32)
    orp, oci = prob.pattern()
    assert np.array_equal(rp, orp) and np.array_equal(ci, oci)
    Ro, Jo = prob.residual_and_jacobian(X)
    R = np.fromfile(os.path.join(d, "R.bin")); J = np.fromfile(os.path.join(d, "J.bin"))
    _close(R, Ro); _close(J, Jo)
    s = _scalars(d)
    assert np.isclose(s["res_l2_norm"], np.linalg.norm(Ro), rtol=1e-12)
    assert s["elem_calculations_throws"] == 1 and s["kernel_launches"] > 0
    return s


@pytest.mark.gpu
def test_mirror_c1_cantilever_reference_layout(built, tmp_path):
    prob = O.cantilever(20, 10, 10, nv=6)
    X = O.state(prob.n_dof, 1); dX = O.state(prob.n_dof, 7)
    X.tofile(tmp_path / "X.bin"); dX.tofile(tmp_path / "dX.bin")
    r = _run("c1", tmp_path)
    assert r.returncode == 0, r.stderr
    s = _compare_common(prob, tmp_path, X)
    assert np.isclose(s["compliance"], prob.compliance(X), rtol=1e-12)
    _close(np.fromfile(tmp_path / "dq_dX.bin"), prob.compliance_derivative(X))
    # sensitivity_assemble for the card's Parameters (true) and a foreign one (false => caller may finite-difference)
    assert s["sens_E_ok"] == 1 and s["sens_nu_ok"] == 1 and s["sens_other_ok"] == 0
    _close(np.fromfile(tmp_path / "dRdE.bin"), prob.sensitivity_assemble_param(X, 1), 1e-11)
    _close(np.fromfile(tmp_path / "dRdnu.bin"), prob.sensitivity_assemble_param(X, 2), 1e-11)
    _close(np.fromfile(tmp_path / "JdX.bin"), prob.linearized_jacobian_solution_product(X, dX), 1e-11)
    assert np.isclose(s["compliance_direct"], prob.compliance(dX), rtol=1e-12)
    # second sweep: halved pressure (load stepping) and point loads
    p2 = O.cantilever(20, 10, 10, nv=6)
    p2.side_pressure[:] = 2.5e7
    R2, _ = p2.residual_and_jacobian(X, want_J=False)
    nodes = np.array([0, 1300, 2540], np.int32)
    p2.add_point_loads(R2, nodes, np.tile([1.0e5, -2.0e5, 3.0e5, 0., 0., 0.], (3, 1)))
    _close(np.fromfile(tmp_path / "R_step.bin"), R2)
    # StressStrainOutputBase through calculate_output / calculate_output_derivative
    q, dq = prob.stress_functional(X, 6.0, 1.5, 2.0e8, want_derivative=True)
    assert np.isclose(s["stress_functional"], q, rtol=1e-12)
    _close(np.fromfile(tmp_path / "ds_dX.bin"), dq, 1e-11)


@pytest.mark.gpu
def test_mirror_simp_adjoint_sensitivities(built, tmp_path):
    prob = O.cantilever(20, 10, 10, nv=6, simp=True)
    X = O.state(prob.n_dof, 1); adj = O.state(prob.n_dof, 2)
    X.tofile(tmp_path / "X.bin"); adj.tofile(tmp_path / "adj.bin"); prob.rho.tofile(tmp_path / "rho.bin")
    r = _run("simp", tmp_path)
    assert r.returncode == 0, r.stderr
    s = _compare_common(prob, tmp_path, X)
    ref = prob.adjoint_sensitivity(X, adj, [0, 3, 5, 6], [100, 101, 121, 1200, 1201, 2540], [0.5, 0.25, 0.25, 0.7, 0.3, 1.0])
    got = np.array([s["sens0"], s["sens1"], s["sens2"]])
    assert np.abs(got - ref).max() <= 1e-11 * np.abs(ref).max()
    assert np.isclose(s["sens1_single"], ref[1], rtol=1e-10)
    # the dR/dp vector itself, and the stress aggregate's total sensitivity = adjoint part + partial part
    assert s["sens_dv_ok"] == 1
    _close(np.fromfile(tmp_path / "dRdp1.bin"), prob.sensitivity_assemble(X, [1200, 1201], [0.7, 0.3]), 1e-11)
    part = prob.stress_functional_density_sensitivity(X, 6.0, 1.5, 2.0e8, [0, 2], [1200, 1201], [0.7, 0.3])[0]
    assert np.isclose(s["stress_adjoint_only"], ref[1], rtol=1e-10)
    assert np.isclose(s["stress_total_sens"], ref[1] + part, rtol=1e-10)


@pytest.mark.gpu
def test_mirror_quad4_heat(built, tmp_path):
    xyz, conn = mesh.build_square(24, 16)
    con = np.isclose(xyz[:, 0], 0.0).astype(np.uint8)
    prob = O.Problem(O.QUAD4, xyz, conn, 1, constrained=con, k0=237.0, k_beta=1e-3, k_of_T=1, thickness=0.06)
    T = 300 + 50 * np.random.Generator(np.random.MT19937(4)).uniform(0, 1, prob.n_dof)
    T.tofile(tmp_path / "X.bin")
    r = _run("heat", tmp_path)
    assert r.returncode == 0, r.stderr
    _compare_common(prob, tmp_path, T)


@pytest.mark.gpu
def test_mirror_hex27_thermoelastic(built, tmp_path):
    xyz, conn = mesh.build_cube(4, 3, 3, 0.3, 0.15, 0.15, 27)
    con = mesh.dirichlet_flags(len(xyz), 6, mesh.boundary_nodes(4, 3, 3, 4, 27))
    prob = O.Problem(O.HEX27, xyz, conn, 6, constrained=con, alpha=1.5e-5, thermal=1, temperature_const=350.0,
                     ref_temperature=300.0)
    X = O.state(prob.n_dof, 1)
    X.tofile(tmp_path / "X.bin")
    r = _run("hex27", tmp_path)
    assert r.returncode == 0, r.stderr
    _compare_common(prob, tmp_path, X)


@pytest.mark.gpu
def test_mirror_quad4_mindlin_plate(built, tmp_path):
    # StructuralNonlinearAssemblyElemOperations on a QUAD4 mesh with a solid 2D section: StructuralElement2D on the device
    xyz, conn = mesh.build_square(24, 16)
    con = mesh.dirichlet_flags(len(xyz), 6, np.flatnonzero(np.isclose(xyz[:, 0], 0.0)))
    prob = O.Problem(O.QUAD4, xyz, conn, 6, constrained=con, E0=72e9, nu=0.33, thickness=0.06, offset=0.03, kappa=5.0 / 6.0)
    X = O.state(prob.n_dof, 1)
    X.tofile(tmp_path / "X.bin")
    r = _run("plate", tmp_path)
    assert r.returncode == 0, r.stderr
    _compare_common(prob, tmp_path, X)
