"""ctypes binding of the C-ABI in ``include/mast_b200.h`` (``libmast_b200.so``).

Thin by design: argument marshalling only.  All arithmetic happens in the CUDA library; if the
library is missing or no B200 is present the calls raise -- there is no CPU fallback.
"""
import ctypes as C
import os

import numpy as np

__all__ = ["Context", "Material", "MastB200Error", "lib_path", "load_library", "HEX8", "HEX27", "QUAD4",
           "EXPORTED_SYMBOLS"]

HEX8, HEX27, QUAD4 = 8, 27, 4

EXPORTED_SYMBOLS = [
    "mastb200_create", "mastb200_destroy", "mastb200_last_error", "mastb200_version", "mastb200_set_mesh",
    "mastb200_set_dof_map", "mastb200_set_material", "mastb200_set_side_pressure", "mastb200_set_heat_side_loads", "mastb200_set_heat_side_loads_ex", "mastb200_set_temperature",
    "mastb200_set_point_loads", "mastb200_sensitivity_assemble", "mastb200_linearized_jacobian_solution_product",
    "mastb200_compliance_direct_sensitivity", "mastb200_set_density", "mastb200_build_pattern", "mastb200_get_pattern", "mastb200_get_pattern_rows", "mastb200_get_row_dofs",
    "mastb200_residual_and_jacobian", "mastb200_res_l2_norm", "mastb200_sensitivity_sweep",
    "mastb200_apply_dv_map_transpose", "mastb200_compliance", "mastb200_compliance_derivative",
    "mastb200_stress", "mastb200_stress_functional", "mastb200_stress_functional_derivative",
    "mastb200_stress_functional_density_derivative",
    "mastb200_mass_matrix", "mastb200_transient_residual_and_jacobian", "mastb200_eigenproblem_assemble",
    "mastb200_filter_init", "mastb200_filter_get", "mastb200_filter_apply", "mastb200_filter_apply_transpose", "mastb200_volume",
    "mastb200_comm_unique_id", "mastb200_comm_init", "mastb200_comm_attach", "mastb200_set_ghost_exchange",
    "mastb200_build_localized_vector", "mastb200_all_reduce_sum", "mastb200_solve", "mastb200_near_null_space",
    "mastb200_get_stats", "mastb200_stream", "mastb200_set_async", "mastb200_debug_plan_sizes", "mastb200_debug_plan_copy",
]


class MastB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"mast_b200 error {code}: {msg}")
        self.code = code


class Material(C.Structure):
    _fields_ = [("E0", C.c_double), ("nu", C.c_double), ("alpha", C.c_double), ("rho_min", C.c_double),
                ("penalty", C.c_double), ("nonlinear_strain", C.c_int32), ("k0", C.c_double),
                ("k_beta", C.c_double), ("thickness", C.c_double), ("k_of_T", C.c_int32),
                ("heat_source", C.c_double), ("offset", C.c_double), ("kappa", C.c_double),
                ("plane_strain", C.c_int32)]


class SolveInfo(C.Structure):
    _fields_ = [("iterations", C.c_int32), ("converged", C.c_int32), ("rel_residual", C.c_double), ("solve_ms", C.c_double)]


class Stats(C.Structure):
    _fields_ = [("n_tiles", C.c_int64), ("n_visits", C.c_int64), ("n_owned_nodes", C.c_int64),
                ("n_elem", C.c_int64), ("kernel_launches", C.c_int64), ("last_sweep_ms", C.c_double),
                ("last_main_kernel_ms", C.c_double), ("plan_bytes", C.c_int64), ("smem_bytes", C.c_int32),
                ("threads_per_cta", C.c_int32), ("ctas", C.c_int32), ("structured", C.c_int32)]


def lib_path():
    # MASTB200_LIB: alternative build of the same library (kernel-tuning experiments)
    return os.environ.get("MASTB200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libmast_b200.so")


_LIB = None


def load_library():
    """Loads libmast_b200.so (built in-tree by __graft_entry__.build()). Fails loudly if absent."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise MastB200Error(-1, f"{path} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                                "(there is no CPU fallback)")
    try:  # make sure the CUDA runtime torch ships is the one that gets resolved
        import torch  # noqa: F401
    except Exception:  # pragma: no cover
        pass
    lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
    vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    lib.mastb200_create.argtypes = [C.c_int, C.POINTER(vp)]
    lib.mastb200_destroy.argtypes = [vp]
    lib.mastb200_last_error.argtypes = [vp]
    lib.mastb200_last_error.restype = C.c_char_p
    lib.mastb200_version.restype = C.c_char_p
    lib.mastb200_set_mesh.argtypes = [vp, i32, i64, i64, vp, vp, vp]
    lib.mastb200_set_dof_map.argtypes = [vp, i32, vp, vp]
    lib.mastb200_set_material.argtypes = [vp, C.POINTER(Material)]
    lib.mastb200_set_side_pressure.argtypes = [vp, i64, vp, vp, vp]
    lib.mastb200_set_heat_side_loads.argtypes = [vp, i64, vp, vp, vp, vp, vp]
    lib.mastb200_set_heat_side_loads_ex.argtypes = [vp, i64, vp, vp, vp, vp, vp, vp, dbl, dbl]
    lib.mastb200_set_temperature.argtypes = [vp, i32, vp, dbl, dbl]
    lib.mastb200_set_density.argtypes = [vp, vp]
    lib.mastb200_set_point_loads.argtypes = [vp, i64, vp, vp]
    lib.mastb200_sensitivity_assemble.argtypes = [vp, vp, i32, i64, vp, vp, vp]
    lib.mastb200_linearized_jacobian_solution_product.argtypes = [vp, vp, vp, vp]
    lib.mastb200_compliance_direct_sensitivity.argtypes = [vp, vp, vp, C.POINTER(dbl)]
    lib.mastb200_build_pattern.argtypes = [vp, C.POINTER(i64), C.POINTER(i64)]
    lib.mastb200_get_pattern.argtypes = [vp, vp, vp]
    lib.mastb200_get_row_dofs.argtypes = [vp, vp]
    lib.mastb200_get_pattern_rows.argtypes = [vp, i64, vp, vp, vp, vp]
    lib.mastb200_residual_and_jacobian.argtypes = [vp, vp, vp, vp]
    lib.mastb200_res_l2_norm.argtypes = [vp, C.POINTER(dbl)]
    lib.mastb200_sensitivity_sweep.argtypes = [vp, vp, vp, vp, vp, vp]
    lib.mastb200_apply_dv_map_transpose.argtypes = [vp, vp, i64, vp, vp, vp, vp]
    lib.mastb200_compliance.argtypes = [vp, vp, C.POINTER(dbl)]
    lib.mastb200_compliance_derivative.argtypes = [vp, vp, vp]
    lib.mastb200_stress.argtypes = [vp, vp, vp, vp, vp, vp]
    lib.mastb200_stress_functional.argtypes = [vp, vp, dbl, dbl, dbl, vp, vp, C.POINTER(dbl)]
    lib.mastb200_stress_functional_derivative.argtypes = [vp, vp, dbl, dbl, dbl, vp, vp, vp]
    lib.mastb200_stress_functional_density_derivative.argtypes = [vp, vp, dbl, dbl, dbl, vp, vp, vp]
    lib.mastb200_mass_matrix.argtypes = [vp, dbl, i32, vp]
    lib.mastb200_transient_residual_and_jacobian.argtypes = [vp, vp, vp, dbl, i32, dbl, vp, vp]
    lib.mastb200_eigenproblem_assemble.argtypes = [vp, vp, dbl, i32, vp, vp]
    lib.mastb200_filter_init.argtypes = [vp, dbl, vp, C.POINTER(i64)]
    lib.mastb200_filter_get.argtypes = [vp, vp, vp, vp]
    lib.mastb200_filter_apply.argtypes = [vp, vp, vp]
    lib.mastb200_filter_apply_transpose.argtypes = [vp, vp, vp]
    lib.mastb200_volume.argtypes = [vp, vp, C.POINTER(dbl), vp]
    lib.mastb200_comm_unique_id.argtypes = [vp, i32]
    lib.mastb200_comm_init.argtypes = [vp, i32, i32, vp]
    lib.mastb200_comm_attach.argtypes = [vp, vp, i32, i32]
    lib.mastb200_set_ghost_exchange.argtypes = [vp, i32, vp, vp, vp, vp, vp, i32]
    lib.mastb200_build_localized_vector.argtypes = [vp, vp, i32]
    lib.mastb200_all_reduce_sum.argtypes = [vp, vp, i32]
    lib.mastb200_solve.argtypes = [vp, vp, vp, vp, dbl, i32, C.POINTER(SolveInfo)]
    lib.mastb200_near_null_space.argtypes = [vp, vp]
    lib.mastb200_get_stats.argtypes = [vp, C.POINTER(Stats)]
    lib.mastb200_debug_plan_sizes.argtypes = [vp, i32, vp]
    lib.mastb200_debug_plan_copy.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp]
    lib.mastb200_stream.argtypes = [vp]
    lib.mastb200_stream.restype = vp
    lib.mastb200_set_async.argtypes = [vp, i32]
    _LIB = lib
    return lib


def _ptr(a):
    """void* of a numpy array, a torch tensor (host or CUDA), or None."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"]
        return a.ctypes.data_as(C.c_void_p)
    if hasattr(a, "data_ptr"):  # torch tensor
        assert a.is_contiguous()
        return C.c_void_p(a.data_ptr())
    raise TypeError(type(a))


def _np(a, dtype):
    return None if a is None else np.ascontiguousarray(a, dtype=dtype)


class Context:
    """One assembly context per GPU (``device=-1``: host-only planning context, no sweeps)."""

    def __init__(self, device=0):
        self._lib = load_library()
        self._h = C.c_void_p()
        rc = self._lib.mastb200_create(int(device), C.byref(self._h))
        if rc:
            raise MastB200Error(rc, self._lib.mastb200_last_error(None).decode())
        self.nv = 0
        self.n_nodes = 0
        self.n_rows = 0
        self.nnz = 0
        self._keep = {}

    def close(self):
        if self._h:
            self._lib.mastb200_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc:
            raise MastB200Error(rc, self._lib.mastb200_last_error(self._h).decode())

    # ---- problem definition ----
    def set_mesh(self, elem_type, xyz, conn, diag_elem=None):
        xyz = _np(xyz, np.float64); conn = _np(conn, np.int32); diag_elem = _np(diag_elem, np.uint8)
        self.n_nodes = xyz.shape[0]
        self.n_elem = conn.shape[0]
        self._ck(self._lib.mastb200_set_mesh(self._h, elem_type, conn.shape[0], xyz.shape[0], _ptr(xyz), _ptr(conn),
                                             _ptr(diag_elem)))

    def set_dof_map(self, nv, constrained=None, owned=None):
        self.nv = nv
        self._ck(self._lib.mastb200_set_dof_map(self._h, nv, _ptr(_np(constrained, np.uint8)),
                                                _ptr(_np(owned, np.uint8))))

    def set_material(self, **kw):
        m = Material(E0=kw.get("E0", 72e9), nu=kw.get("nu", 0.33), alpha=kw.get("alpha", 0.0),
                     rho_min=kw.get("rho_min", 0.0), penalty=kw.get("penalty", 1.0),
                     nonlinear_strain=int(kw.get("nonlinear_strain", 0)), k0=kw.get("k0", 0.0),
                     k_beta=kw.get("k_beta", 0.0), thickness=kw.get("thickness", 1.0),
                     k_of_T=int(kw.get("k_of_T", 0)), heat_source=kw.get("heat_source", 0.0),
                     offset=kw.get("offset", 0.0), kappa=kw.get("kappa", 5.0 / 6.0),
                     plane_strain=int(kw.get("plane_strain", 0)))
        self._ck(self._lib.mastb200_set_material(self._h, C.byref(m)))

    def set_side_pressure(self, elem, side, pressure):
        elem = _np(elem, np.int32); side = _np(side, np.int32); pressure = _np(pressure, np.float64)
        self._ck(self._lib.mastb200_set_side_pressure(self._h, len(elem), _ptr(elem), _ptr(side), _ptr(pressure)))

    def set_heat_side_loads(self, elem, side, heat_flux=None, conv_coeff=None, T_amb=None, emissivity=None,
                            sb_const=0.0, T_ref_zero=0.0):
        """HEAT_FLUX / CONVECTION_HEAT_FLUX / SURFACE_RADIATION_HEAT_FLUX on (element, side) entries of a QUAD4 mesh."""
        elem = _np(elem, np.int32); side = _np(side, np.int32)
        self._ck(self._lib.mastb200_set_heat_side_loads_ex(
            self._h, len(elem), _ptr(elem), _ptr(side), _ptr(_np(heat_flux, np.float64)),
            _ptr(_np(conv_coeff, np.float64)), _ptr(_np(T_amb, np.float64)), _ptr(_np(emissivity, np.float64)),
            float(sb_const), float(T_ref_zero)))

    def set_temperature(self, enable, T_nodal=None, T_const=0.0, T_ref=0.0):
        self._keep["T"] = T_nodal
        self._ck(self._lib.mastb200_set_temperature(self._h, int(enable), _ptr(T_nodal), T_const, T_ref))

    def set_point_loads(self, node, load):
        node = _np(node, np.int32); load = _np(load, np.float64)
        self._ck(self._lib.mastb200_set_point_loads(self._h, len(node), _ptr(node), _ptr(load)))

    def set_density(self, rho):
        self._keep["rho"] = rho  # a device tensor is used in place: keep it alive
        self._ck(self._lib.mastb200_set_density(self._h, _ptr(rho)))

    def build_pattern(self):
        n_rows, nnz = C.c_int64(), C.c_int64()
        self._ck(self._lib.mastb200_build_pattern(self._h, C.byref(n_rows), C.byref(nnz)))
        self.n_rows, self.nnz = n_rows.value, nnz.value
        return self.n_rows, self.nnz

    def get_pattern(self, with_cols=True):
        row_ptr = np.empty(self.n_rows + 1, np.int64)
        col_idx = np.empty(self.nnz, np.int32) if with_cols else None
        self._ck(self._lib.mastb200_get_pattern(self._h, _ptr(row_ptr), _ptr(col_idx)))
        return row_ptr, col_idx

    def get_pattern_rows(self, rows):
        """(row_begin, row_len, col_idx) of a sample of CSR rows: positions of the rows' values in J and their columns."""
        rows = _np(rows, np.int64)
        begin = np.empty(len(rows), np.int64); length = np.empty(len(rows), np.int32)
        self._ck(self._lib.mastb200_get_pattern_rows(self._h, len(rows), _ptr(rows), _ptr(begin), _ptr(length), None))
        cols = np.empty(int(length.sum()), np.int32)
        self._ck(self._lib.mastb200_get_pattern_rows(self._h, len(rows), _ptr(rows), _ptr(begin), _ptr(length), _ptr(cols)))
        return begin, length, cols

    def get_row_dofs(self):
        out = np.empty(self.n_rows, np.int64)
        self._ck(self._lib.mastb200_get_row_dofs(self._h, _ptr(out)))
        return out

    # ---- sweeps (arrays: numpy = host buffers, torch CUDA tensors = device buffers used in place) ----
    def residual_and_jacobian(self, X, R=None, J=None):
        self._ck(self._lib.mastb200_residual_and_jacobian(self._h, _ptr(X), _ptr(R), _ptr(J)))

    def res_l2_norm(self):
        v = C.c_double()
        self._ck(self._lib.mastb200_res_l2_norm(self._h, C.byref(v)))
        return v.value

    def sensitivity_sweep(self, X, adj, G, R=None, J=None):
        self._ck(self._lib.mastb200_sensitivity_sweep(self._h, _ptr(X), _ptr(adj), _ptr(G), _ptr(R), _ptr(J)))

    def apply_dv_map_transpose(self, G, dv_ptr, dv_node, dv_weight):
        dv_ptr = _np(dv_ptr, np.int64); dv_node = _np(dv_node, np.int32); dv_weight = _np(dv_weight, np.float64)
        sens = np.empty(len(dv_ptr) - 1, np.float64)
        self._ck(self._lib.mastb200_apply_dv_map_transpose(self._h, _ptr(G), len(dv_ptr) - 1, _ptr(dv_ptr),
                                                           _ptr(dv_node), _ptr(dv_weight), _ptr(sens)))
        return sens

    def sensitivity_assemble(self, X, dRdp, param=0, dv_node=None, dv_weight=None):
        """dR/dp of one parameter: param 0 = nodal-density DV (dv_node, dv_weight), 1 = E0, 2 = nu (3D solids);
        3 = conductivity k0, 4 = section thickness (QUAD4 heat conduction)."""
        dv_node = _np(dv_node, np.int32); dv_weight = _np(dv_weight, np.float64)
        n_k = 0 if dv_node is None else len(dv_node)
        self._ck(self._lib.mastb200_sensitivity_assemble(self._h, _ptr(X), int(param), n_k, _ptr(dv_node), _ptr(dv_weight), _ptr(dRdp)))

    def linearized_jacobian_solution_product(self, X, dX, out):
        self._ck(self._lib.mastb200_linearized_jacobian_solution_product(self._h, _ptr(X), _ptr(dX), _ptr(out)))

    def compliance_direct_sensitivity(self, X, dXdp):
        v = C.c_double()
        self._ck(self._lib.mastb200_compliance_direct_sensitivity(self._h, _ptr(X), _ptr(dXdp), C.byref(v)))
        return v.value

    def compliance(self, X):
        v = C.c_double()
        self._ck(self._lib.mastb200_compliance(self._h, _ptr(X), C.byref(v)))
        return v.value

    def compliance_derivative(self, X, dq_dX):
        self._ck(self._lib.mastb200_compliance_derivative(self._h, _ptr(X), _ptr(dq_dX)))

    def stress(self, X, stress=None, strain=None, von_mises=None, jxw=None):
        self._ck(self._lib.mastb200_stress(self._h, _ptr(X), _ptr(stress), _ptr(strain), _ptr(von_mises), _ptr(jxw)))

    def stress_functional(self, X, p_norm, rho_exp, sigma0, elem_mask=None):
        """Returns (q, sums) with sums = [numerator, denominator] of this rank."""
        sums = np.zeros(2)
        q = C.c_double()
        self._ck(self._lib.mastb200_stress_functional(self._h, _ptr(X), p_norm, rho_exp, sigma0,
                                                      _ptr(_np(elem_mask, np.uint8)), _ptr(sums), C.byref(q)))
        return q.value, sums

    def stress_functional_derivative(self, X, p_norm, rho_exp, sigma0, sums, dq_dX, elem_mask=None):
        sums = _np(sums, np.float64)
        self._ck(self._lib.mastb200_stress_functional_derivative(self._h, _ptr(X), p_norm, rho_exp, sigma0,
                                                                 _ptr(_np(elem_mask, np.uint8)), _ptr(sums), _ptr(dq_dX)))

    def stress_functional_density_derivative(self, X, p_norm, rho_exp, sigma0, sums, dq_drho, elem_mask=None):
        sums = _np(sums, np.float64)
        self._ck(self._lib.mastb200_stress_functional_density_derivative(self._h, _ptr(X), p_norm, rho_exp, sigma0,
                                                                         _ptr(_np(elem_mask, np.uint8)), _ptr(sums),
                                                                         _ptr(dq_drho)))

    # ---- transient / modal ----
    def mass_matrix(self, rho_m, M, lumped=False):
        self._ck(self._lib.mastb200_mass_matrix(self._h, float(rho_m), int(bool(lumped)), _ptr(M)))

    def transient_residual_and_jacobian(self, X, rate, rho_m, jac_coeff, R=None, J=None, lumped=False):
        self._ck(self._lib.mastb200_transient_residual_and_jacobian(self._h, _ptr(X), _ptr(rate), float(rho_m), int(bool(lumped)),
                                                                    float(jac_coeff), _ptr(R), _ptr(J)))

    def eigenproblem_assemble(self, X, rho_m, A, B, lumped=False):
        self._ck(self._lib.mastb200_eigenproblem_assemble(self._h, _ptr(X), float(rho_m), int(bool(lumped)), _ptr(A), _ptr(B)))

    # ---- density filter, volume ----
    def filter_init(self, radius, dv_node=None):
        nnz = C.c_int64()
        self._ck(self._lib.mastb200_filter_init(self._h, float(radius), _ptr(_np(dv_node, np.uint8)), C.byref(nnz)))
        self._filter_nnz = nnz.value
        return nnz.value

    def filter_get(self):
        rp = np.empty(self.n_nodes + 1, np.int64); col = np.empty(self._filter_nnz, np.int32); w = np.empty(self._filter_nnz)
        self._ck(self._lib.mastb200_filter_get(self._h, _ptr(rp), _ptr(col), _ptr(w)))
        return rp, col, w

    def filter_apply(self, d, rho):
        self._ck(self._lib.mastb200_filter_apply(self._h, _ptr(d), _ptr(rho)))

    def filter_apply_transpose(self, G, grad):
        self._ck(self._lib.mastb200_filter_apply_transpose(self._h, _ptr(G), _ptr(grad)))

    def volume(self, rho, dV_drho=None):
        v = C.c_double()
        self._ck(self._lib.mastb200_volume(self._h, _ptr(rho), C.byref(v), _ptr(dV_drho)))
        return v.value

    # ---- communicator, linear solve ----
    @staticmethod
    def comm_unique_id():
        """128 bytes rank 0 hands to every rank (any broadcast will do) for ``comm_init``."""
        buf = np.zeros(128, np.uint8)
        rc = load_library().mastb200_comm_unique_id(_ptr(buf), 128)
        if rc:
            raise MastB200Error(rc, load_library().mastb200_last_error(None).decode())
        return buf

    def comm_init(self, rank, world, unique_id):
        unique_id = _np(unique_id, np.uint8)
        self._ck(self._lib.mastb200_comm_init(self._h, int(rank), int(world), _ptr(unique_id)))

    def set_ghost_exchange(self, nbr_rank, send_lists, recv_lists, auto_localize=True):
        """send_lists[k] / recv_lists[k]: local node ids exchanged with rank nbr_rank[k] (same order on both sides)."""
        nbr_rank = _np(nbr_rank, np.int32)
        sp = np.concatenate([[0], np.cumsum([len(a) for a in send_lists])]).astype(np.int64)
        rp = np.concatenate([[0], np.cumsum([len(a) for a in recv_lists])]).astype(np.int64)
        sn = _np(np.concatenate(send_lists) if len(send_lists) else np.zeros(0), np.int32)
        rn = _np(np.concatenate(recv_lists) if len(recv_lists) else np.zeros(0), np.int32)
        self._ck(self._lib.mastb200_set_ghost_exchange(self._h, len(nbr_rank), _ptr(nbr_rank), _ptr(sp), _ptr(sn), _ptr(rp),
                                                       _ptr(rn), int(bool(auto_localize))))

    def build_localized_vector(self, V, width):
        self._ck(self._lib.mastb200_build_localized_vector(self._h, _ptr(V), int(width)))

    def all_reduce_sum(self, vals):
        vals = np.ascontiguousarray(vals, np.float64)
        self._ck(self._lib.mastb200_all_reduce_sum(self._h, _ptr(vals), len(vals)))
        return vals

    def solve(self, J, rhs, x, rtol=1e-8, max_it=10000):
        """Block-Jacobi PCG on the device CSR; returns dict(iterations, converged, rel_residual, solve_ms)."""
        info = SolveInfo()
        self._ck(self._lib.mastb200_solve(self._h, _ptr(J), _ptr(rhs), _ptr(x), float(rtol), int(max_it), C.byref(info)))
        return {k: getattr(info, k) for k, _ in SolveInfo._fields_}

    def near_null_space(self, modes):
        self._ck(self._lib.mastb200_near_null_space(self._h, _ptr(modes)))

    def stats(self):
        s = Stats()
        self._ck(self._lib.mastb200_get_stats(self._h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in Stats._fields_}

    def debug_plan(self, aux=0):
        """Streaming plan of a planning-only context as numpy structured arrays (layouts: csrc/plan.h)."""
        sz = np.zeros(9, np.int64)
        self._ck(self._lib.mastb200_debug_plan_sizes(self._h, int(aux), _ptr(sz)))
        tiles = np.zeros(sz[0], np.dtype([("step_begin", "i4"), ("n_steps", "i4"), ("n_visits", "i4"), ("n_owned", "i4")]))
        steps = np.zeros(sz[1], np.dtype([("visit_begin", "i4"), ("n_visits", "i4"), ("load_begin", "i4"), ("n_loads", "i4"),
                                          ("flush_begin", "i4"), ("n_flush", "i4"), ("fent_begin", "i4"), ("n_fent", "i4"),
                                          ("n_pos", "i4"), ("pad", "i4", 3)]))
        visits = np.zeros(sz[2], np.dtype([("ln", "u2", 8), ("rank", "u1", 8), ("pair", "u2", 36)]))
        frows = np.zeros(sz[3], np.dtype([("dst_block", "i8"), ("gid", "i4"), ("ent_rel", "i4"), ("slot", "u2"),
                                          ("nnb", "u2"), ("nadj", "u2"), ("nadjd", "u2"), ("pad", "i4", 2)]))
        fent = np.zeros(sz[4], np.uint32)
        loads = np.zeros(sz[5], np.dtype([("gid", "i4"), ("slot", "u2"), ("owned", "u2")]))
        assert visits.itemsize == 96 and frows.itemsize == 32 and steps.itemsize == 48 and loads.itemsize == 8
        self._ck(self._lib.mastb200_debug_plan_copy(self._h, int(aux), _ptr(tiles), _ptr(steps), _ptr(visits),
                                                    _ptr(frows), _ptr(fent), _ptr(loads)))
        return dict(tiles=tiles, steps=steps, visits=visits, flush_rows=frows, flush_ent=fent, loads=loads,
                    arena_blocks=int(sz[6]), max_slots=int(sz[7]), n_workers=int(sz[8]))

    @property
    def stream(self):
        return self._lib.mastb200_stream(self._h)

    def set_async(self, on=True):
        """Stream-ordered mode: sweeps whose buffers all live on the device return without synchronising the stream."""
        self._ck(self._lib.mastb200_set_async(self._h, 1 if on else 0))
