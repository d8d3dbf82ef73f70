"""ctypes wrapper of the CPU oracle (oracle/liboracle.so).  Test infrastructure only."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEX8, HEX27, QUAD4 = 8, 27, 4


class OracleProblem(C.Structure):
    _fields_ = [
        ("elem_type", C.c_int32), ("n_elem", C.c_int64), ("n_nodes", C.c_int64), ("xyz", C.c_void_p),
        ("conn", C.c_void_p), ("diag_elem", C.c_void_p), ("nv", C.c_int32), ("constrained", C.c_void_p),
        ("E0", C.c_double), ("nu", C.c_double), ("alpha", C.c_double), ("rho", C.c_void_p),
        ("rho_min", C.c_double), ("penalty", C.c_double), ("nonlinear_strain", C.c_int32),
        ("n_side_loads", C.c_int64), ("side_elem", C.c_void_p), ("side_id", C.c_void_p),
        ("side_pressure", C.c_void_p), ("thermal", C.c_int32), ("temperature", C.c_void_p),
        ("temperature_const", C.c_double), ("ref_temperature", C.c_double), ("k0", C.c_double),
        ("k_beta", C.c_double), ("thickness", C.c_double), ("k_of_T", C.c_int32), ("heat_source", C.c_double),
        ("offset", C.c_double), ("kappa", C.c_double), ("plane_strain", C.c_int32),
        ("side_conv_coeff", C.c_void_p), ("side_T_amb", C.c_void_p),
        ("side_emissivity", C.c_void_p), ("rad_sb_const", C.c_double), ("rad_T_ref_zero", C.c_double),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(ROOT, "oracle", "liboracle.so")
        L = C.CDLL(path)
        vp, i64, dbl = C.c_void_p, C.c_int64, C.c_double
        P = C.POINTER(OracleProblem)
        L.oracle_build_pattern.argtypes = [P, vp, vp]; L.oracle_build_pattern.restype = i64
        L.oracle_residual_and_jacobian.argtypes = [P, vp, vp, vp, vp, vp]
        L.oracle_elem_calculations.argtypes = [P, i64, vp, C.c_int, vp, vp]
        L.oracle_sensitivity_assemble.argtypes = [P, vp, i64, i64, vp, vp, vp]
        L.oracle_adjoint_sensitivity.argtypes = [P, vp, vp, i64, vp, vp, vp, vp]
        L.oracle_mass_matrix.argtypes = [P, dbl, C.c_int, vp, vp, vp]
        L.oracle_transient_residual_and_jacobian.argtypes = [P, vp, vp, dbl, C.c_int, dbl, vp, vp, vp, vp]
        L.oracle_sensitivity_assemble_param.argtypes = [P, vp, C.c_int, vp]
        L.oracle_linearized_jacobian_solution_product.argtypes = [P, vp, vp, vp]
        L.oracle_add_point_loads.argtypes = [P, i64, vp, vp, vp]
        L.oracle_compliance.argtypes = [P, vp]; L.oracle_compliance.restype = dbl
        L.oracle_compliance_derivative.argtypes = [P, vp, vp]
        L.oracle_stress.argtypes = [P, vp, vp, vp, vp, vp]
        L.oracle_stress_functional.argtypes = [P, vp, dbl, dbl, dbl, vp]; L.oracle_stress_functional.restype = dbl
        L.oracle_stress_functional_density_sensitivity.argtypes = [P, vp, dbl, dbl, dbl, i64, vp, vp, vp, vp]
        L.oracle_filter_weights.argtypes = [i64, vp, dbl, vp, vp, vp]; L.oracle_filter_weights.restype = i64
        L.oracle_filtered_values.argtypes = [i64, vp, vp, vp, vp, vp, vp]
        L.oracle_volume.argtypes = [P, vp, vp]; L.oracle_volume.restype = dbl
        L.oracle_stiffness_3d.argtypes = [dbl, dbl, vp]
        L.oracle_stiffness_3d_derivative.argtypes = [dbl, dbl, dbl, dbl, vp]
        L.oracle_thermal_expansion_A_3d.argtypes = [dbl, dbl, dbl, vp]
        L.oracle_plate_section_matrices.argtypes = [dbl, dbl, dbl, dbl, dbl, C.c_int, vp, vp, vp, vp]
        L.oracle_fe_tables.argtypes = [C.c_int32, vp, vp, vp, vp, vp]
        L.oracle_time_residual_and_jacobian.argtypes = [P, vp, vp, vp, vp, vp, C.c_int]
        L.oracle_time_residual_and_jacobian.restype = dbl
        L.oracle_time_residual_and_jacobian_ex.argtypes = [P, vp, vp, vp, vp, vp, C.c_int, C.POINTER(C.c_int)]
        L.oracle_time_residual_and_jacobian_ex.restype = dbl
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Problem:
    """Holds numpy arrays alive and exposes the OracleProblem struct."""

    def __init__(self, elem_type, xyz, conn, nv, constrained=None, E0=72e9, nu=0.33, alpha=0.0, rho=None,
                 rho_min=0.0, penalty=1.0, nonlinear_strain=0, side_elem=None, side_id=None, side_pressure=None,
                 thermal=0, temperature=None, temperature_const=0.0, ref_temperature=0.0, k0=0.0, k_beta=0.0,
                 thickness=1.0, k_of_T=0, heat_source=0.0, diag_elem=None, offset=0.0, kappa=5.0 / 6.0, plane_strain=0, side_conv_coeff=None, side_T_amb=None,
                 side_emissivity=None, rad_sb_const=0.0, rad_T_ref_zero=0.0):
        self.elem_type = elem_type
        self.xyz = np.ascontiguousarray(xyz, np.float64)
        self.conn = np.ascontiguousarray(conn, np.int32)
        self.nv = nv
        self.n_nodes, self.n_elem, self.nn = self.xyz.shape[0], self.conn.shape[0], self.conn.shape[1]
        self.n_dof = self.n_nodes * nv
        self.constrained = None if constrained is None else np.ascontiguousarray(constrained, np.uint8)
        self.rho = None if rho is None else np.ascontiguousarray(rho, np.float64)
        self.side_elem = None if side_elem is None else np.ascontiguousarray(side_elem, np.int32)
        self.side_id = None if side_id is None else np.ascontiguousarray(side_id, np.int32)
        self.side_pressure = None if side_pressure is None else np.ascontiguousarray(side_pressure, np.float64)
        self.temperature = None if temperature is None else np.ascontiguousarray(temperature, np.float64)
        self.diag_elem = None if diag_elem is None else np.ascontiguousarray(diag_elem, np.uint8)
        self.mat = dict(E0=E0, nu=nu, alpha=alpha, rho_min=rho_min, penalty=penalty, nonlinear_strain=nonlinear_strain,
                        k0=k0, k_beta=k_beta, thickness=thickness, k_of_T=k_of_T, heat_source=heat_source,
                        offset=offset, kappa=kappa, plane_strain=plane_strain)
        self.thermal, self.temperature_const, self.ref_temperature = thermal, temperature_const, ref_temperature
        s = OracleProblem()
        s.elem_type = elem_type; s.n_elem = self.n_elem; s.n_nodes = self.n_nodes
        s.xyz = _p(self.xyz); s.conn = _p(self.conn); s.diag_elem = _p(self.diag_elem)
        s.nv = nv; s.constrained = _p(self.constrained)
        s.E0 = E0; s.nu = nu; s.alpha = alpha; s.rho = _p(self.rho); s.rho_min = rho_min; s.penalty = penalty
        s.nonlinear_strain = nonlinear_strain
        s.n_side_loads = 0 if self.side_elem is None else len(self.side_elem)
        s.side_elem = _p(self.side_elem); s.side_id = _p(self.side_id); s.side_pressure = _p(self.side_pressure)
        s.thermal = thermal; s.temperature = _p(self.temperature); s.temperature_const = temperature_const
        s.ref_temperature = ref_temperature
        s.k0 = k0; s.k_beta = k_beta; s.thickness = thickness; s.k_of_T = k_of_T; s.heat_source = heat_source
        s.offset = offset; s.kappa = kappa; s.plane_strain = plane_strain
        self.side_conv_coeff = None if side_conv_coeff is None else np.ascontiguousarray(side_conv_coeff, np.float64)
        self.side_T_amb = None if side_T_amb is None else np.ascontiguousarray(side_T_amb, np.float64)
        s.side_conv_coeff = _p(self.side_conv_coeff); s.side_T_amb = _p(self.side_T_amb)
        self.side_emissivity = None if side_emissivity is None else np.ascontiguousarray(side_emissivity, np.float64)
        self.rad_sb_const, self.rad_T_ref_zero = rad_sb_const, rad_T_ref_zero
        s.side_emissivity = _p(self.side_emissivity); s.rad_sb_const = rad_sb_const; s.rad_T_ref_zero = rad_T_ref_zero
        self.s = s
        self._pattern = None

    # ---- oracle calls ----
    def pattern(self):
        if self._pattern is None:
            row_ptr = np.zeros(self.n_dof + 1, np.int64)
            nnz = lib().oracle_build_pattern(C.byref(self.s), _p(row_ptr), None)
            col = np.empty(nnz, np.int32)
            lib().oracle_build_pattern(C.byref(self.s), _p(row_ptr), _p(col))
            self._pattern = (row_ptr, col)
        return self._pattern

    def residual_and_jacobian(self, X, want_R=True, want_J=True):
        row_ptr, col = self.pattern()
        R = np.empty(self.n_dof) if want_R else None
        J = np.empty(len(col)) if want_J else None
        rc = lib().oracle_residual_and_jacobian(C.byref(self.s), _p(X), _p(R), _p(J), _p(row_ptr), _p(col))
        assert rc == 0, rc
        return R, J

    def elem_calculations(self, e, X, want_jac=True):
        nd = self.nv * self.nn
        vec = np.zeros(nd); mat = np.zeros((nd, nd))
        rc = lib().oracle_elem_calculations(C.byref(self.s), e, _p(X), int(want_jac), _p(vec), _p(mat))
        assert rc == 0
        return vec, mat

    def adjoint_sensitivity(self, X, adj, dv_ptr, dv_node, dv_weight):
        dv_ptr = np.ascontiguousarray(dv_ptr, np.int64); dv_node = np.ascontiguousarray(dv_node, np.int32)
        dv_weight = np.ascontiguousarray(dv_weight, np.float64)
        sens = np.empty(len(dv_ptr) - 1)
        rc = lib().oracle_adjoint_sensitivity(C.byref(self.s), _p(X), _p(adj), len(dv_ptr) - 1, _p(dv_ptr),
                                              _p(dv_node), _p(dv_weight), _p(sens))
        assert rc == 0, rc
        return sens

    def sensitivity_assemble(self, X, dv_node, dv_weight):
        dv_node = np.ascontiguousarray(dv_node, np.int32); dv_weight = np.ascontiguousarray(dv_weight, np.float64)
        out = np.empty(self.n_dof)
        rc = lib().oracle_sensitivity_assemble(C.byref(self.s), _p(X), 0, len(dv_node), _p(dv_node), _p(dv_weight), _p(out))
        assert rc == 0, rc
        return out

    def sensitivity_assemble_param(self, X, kind):
        """dR/dp for a Parameter of the cards: 3D solids kind 1 = E0, 2 = nu; QUAD4 heat conduction kind 1 = k0, 2 = thickness."""
        out = np.empty(self.n_dof)
        rc = lib().oracle_sensitivity_assemble_param(C.byref(self.s), _p(X), int(kind), _p(out))
        assert rc == 0, rc
        return out

    def linearized_jacobian_solution_product(self, X, dX):
        out = np.empty(self.n_dof)
        rc = lib().oracle_linearized_jacobian_solution_product(C.byref(self.s), _p(X), _p(dX), _p(out))
        assert rc == 0, rc
        return out

    def add_point_loads(self, R, node, load):
        node = np.ascontiguousarray(node, np.int32); load = np.ascontiguousarray(load, np.float64)
        assert lib().oracle_add_point_loads(C.byref(self.s), len(node), _p(node), _p(load), _p(R)) == 0
        return R

    def filter_weights(self, radius):
        rp = np.zeros(self.n_nodes + 1, np.int64)
        nnz = lib().oracle_filter_weights(self.n_nodes, _p(self.xyz), radius, _p(rp), None, None)
        col = np.empty(nnz, np.int32); w = np.empty(nnz)
        lib().oracle_filter_weights(self.n_nodes, _p(self.xyz), radius, _p(rp), _p(col), _p(w))
        return rp, col, w

    def filtered_values(self, filt, d, dv=None):
        rp, col, w = filt
        out = np.empty(self.n_nodes)
        dv = None if dv is None else np.ascontiguousarray(dv, np.uint8)
        lib().oracle_filtered_values(self.n_nodes, _p(rp), _p(col), _p(w), _p(dv), _p(np.ascontiguousarray(d, np.float64)), _p(out))
        return out

    def volume(self, rho, want_gradient=False):
        g = np.empty(self.n_nodes) if want_gradient else None
        v = lib().oracle_volume(C.byref(self.s), _p(np.ascontiguousarray(rho, np.float64)), _p(g))
        return (v, g) if want_gradient else v

    def mass_matrix(self, rho_m, lumped=False):
        row_ptr, col = self.pattern()
        M = np.empty(len(col))
        assert lib().oracle_mass_matrix(C.byref(self.s), rho_m, int(lumped), _p(row_ptr), _p(col), _p(M)) == 0
        return M

    def transient_residual_and_jacobian(self, X, rate, rho_m, jac_coeff, lumped=False):
        row_ptr, col = self.pattern()
        R = np.empty(self.n_dof); J = np.empty(len(col))
        rc = lib().oracle_transient_residual_and_jacobian(C.byref(self.s), _p(X), _p(rate), rho_m, int(lumped), jac_coeff,
                                                          _p(R), _p(J), _p(row_ptr), _p(col))
        assert rc == 0, rc
        return R, J

    def compliance(self, X):
        return lib().oracle_compliance(C.byref(self.s), _p(X))

    def compliance_derivative(self, X):
        out = np.empty(self.n_dof)
        assert lib().oracle_compliance_derivative(C.byref(self.s), _p(X), _p(out)) == 0
        return out

    def stress(self, X):
        nq = self.nn
        n = self.n_elem * nq
        st = np.empty((n, 6)); en = np.empty((n, 6)); vm = np.empty(n); jw = np.empty(n)
        assert lib().oracle_stress(C.byref(self.s), _p(X), _p(st), _p(en), _p(vm), _p(jw)) == 0
        return st, en, vm, jw

    def stress_functional(self, X, p_norm, rho_exp, sigma0, want_derivative=False):
        dq = np.empty(self.n_dof) if want_derivative else None
        q = lib().oracle_stress_functional(C.byref(self.s), _p(X), p_norm, rho_exp, sigma0, _p(dq))
        return (q, dq) if want_derivative else q

    def stress_functional_density_sensitivity(self, X, p_norm, rho_exp, sigma0, dv_ptr, dv_node, dv_weight):
        dv_ptr = np.ascontiguousarray(dv_ptr, np.int64); dv_node = np.ascontiguousarray(dv_node, np.int32)
        dv_weight = np.ascontiguousarray(dv_weight, np.float64)
        sens = np.empty(len(dv_ptr) - 1)
        rc = lib().oracle_stress_functional_density_sensitivity(C.byref(self.s), _p(X), p_norm, rho_exp, sigma0,
                                                                len(dv_ptr) - 1, _p(dv_ptr), _p(dv_node), _p(dv_weight), _p(sens))
        assert rc == 0, rc
        return sens

    def time_residual_and_jacobian(self, X, threads, want_R=True, want_J=True, out=None):
        """Seconds of one sweep with up to ``threads`` threads; ``out`` (a dict) receives the thread count really used
        and the result arrays (reused between calls when passed back in)."""
        row_ptr, col = self.pattern()
        out = {} if out is None else out
        if want_R and "R" not in out:
            out["R"] = np.empty(self.n_dof)
        if want_J and "J" not in out:
            out["J"] = np.empty(len(col))
        used = C.c_int(0)
        t = lib().oracle_time_residual_and_jacobian_ex(C.byref(self.s), _p(X), _p(out.get("R") if want_R else None),
                                                       _p(out.get("J") if want_J else None), _p(row_ptr), _p(col), threads,
                                                       C.byref(used))
        out["threads_used"] = used.value
        return t

    # ---- the same problem on the device library ----
    def make_context(self, device=0):
        import mast_b200
        ctx = mast_b200.Context(device)
        ctx.set_mesh(self.elem_type, self.xyz, self.conn, self.diag_elem)
        ctx.set_dof_map(self.nv, self.constrained)
        ctx.set_material(**self.mat)
        if device >= 0:
            if self.side_elem is not None:
                if self.elem_type == QUAD4 and self.nv == 1:
                    ctx.set_heat_side_loads(self.side_elem, self.side_id, self.side_pressure, self.side_conv_coeff,
                                            self.side_T_amb, self.side_emissivity, self.rad_sb_const, self.rad_T_ref_zero)
                else:
                    ctx.set_side_pressure(self.side_elem, self.side_id, self.side_pressure)
            if self.thermal:
                ctx.set_temperature(1, self.temperature, self.temperature_const, self.ref_temperature)
            if self.rho is not None:
                ctx.set_density(self.rho)
        ctx.build_pattern()
        return ctx


def cantilever(nx=20, ny=10, nz=10, lx=0.3, ly=0.15, lz=0.15, nv=3, jitter=0.0, simp=False, **kw):
    """BASELINE config C1: clamped at x=0 (boundary 4), pressure 5e7 on y=Ly (boundary 3)."""
    from mast_b200 import mesh
    xyz, conn = mesh.build_cube(nx, ny, nz, lx, ly, lz, 8, jitter=jitter)
    con = mesh.dirichlet_flags(len(xyz), nv, mesh.boundary_nodes(nx, ny, nz, 4))
    se, sid = mesh.boundary_sides(nx, ny, nz, 3)
    rho = None
    if simp:
        rng = np.random.Generator(np.random.MT19937(3))
        rho = 0.2 + 0.6 * rng.uniform(0.0, 1.0, len(xyz))
        kw.setdefault("rho_min", 1e-8); kw.setdefault("penalty", 4.0)
    return Problem(HEX8, xyz, conn, nv, constrained=con, side_elem=se, side_id=sid,
                   side_pressure=np.full(len(se), 5e7), rho=rho, **kw)


def state(n_dof, seed=1, scale=1e-3):
    rng = np.random.Generator(np.random.MT19937(seed))
    return scale * rng.uniform(-1.0, 1.0, n_dof)
