// mast_b200_assembly.cpp -- see mast_b200_assembly.h.  Thin adaptors: libMesh objects -> plain arrays -> C-ABI.
#include "mast_b200_assembly.h"

#include <cmath>

#include <cstring>

#include "../../include/mast_b200.h"

namespace libMesh {
#ifndef MAST_B200_USE_LIBMESH
namespace MeshTools {
namespace Generation {

void build_cube(MeshBase& mesh, unsigned nx, unsigned ny, unsigned nz, Real xmin, Real xmax, Real ymin, Real ymax,
                Real zmin, Real zmax, ElemType type) {
  if (type != HEX8 && type != HEX27) libmesh_error_msg("build_cube: HEX8 or HEX27");
  const unsigned m = type == HEX8 ? 1 : 2;
  const unsigned npx = m * nx + 1, npy = m * ny + 1, npz = m * nz + 1;
  mesh.elem_type = type;
  mesh.xyz.resize(size_t(npx) * npy * npz * 3);
  for (unsigned k = 0; k < npz; ++k)
    for (unsigned j = 0; j < npy; ++j)
      for (unsigned i = 0; i < npx; ++i) {
        Real* p = &mesh.xyz[(size_t(i) + size_t(j) * npx + size_t(k) * npx * npy) * 3];
        p[0] = xmin + (xmax - xmin) * i / (npx - 1);
        p[1] = ymin + (ymax - ymin) * j / (npy - 1);
        p[2] = zmin + (zmax - zmin) * k / (npz - 1);
      }
  // reference position of each element node on the local (2m+1)^3 lattice, libMesh node order
  static const int8_t H27[27][3] = {{0, 0, 0}, {2, 0, 0}, {2, 2, 0}, {0, 2, 0}, {0, 0, 2}, {2, 0, 2}, {2, 2, 2}, {0, 2, 2},
                                    {1, 0, 0}, {2, 1, 0}, {1, 2, 0}, {0, 1, 0}, {0, 0, 1}, {2, 0, 1}, {2, 2, 1}, {0, 2, 1},
                                    {1, 0, 2}, {2, 1, 2}, {1, 2, 2}, {0, 1, 2}, {1, 1, 0}, {1, 0, 1}, {2, 1, 1}, {1, 2, 1},
                                    {0, 1, 1}, {1, 1, 2}, {1, 1, 1}};
  const unsigned nn = type == HEX8 ? 8 : 27;
  mesh.conn.resize(size_t(nx) * ny * nz * nn);
  mesh.subdomain.assign(size_t(nx) * ny * nz, 0);
  for (unsigned k = 0; k < nz; ++k)
    for (unsigned j = 0; j < ny; ++j)
      for (unsigned i = 0; i < nx; ++i) {
        const size_t e = size_t(i) + size_t(j) * nx + size_t(k) * nx * ny;
        for (unsigned a = 0; a < nn; ++a) {
          const unsigned di = H27[a][0] * m / 2, dj = H27[a][1] * m / 2, dk = H27[a][2] * m / 2;
          mesh.conn[e * nn + a] = int32_t((m * i + di) + size_t(m * j + dj) * npx + size_t(m * k + dk) * npx * npy);
        }
        if (k == 0) mesh.boundary_info.add_side((dof_id_type)e, 0, 0);
        if (j == 0) mesh.boundary_info.add_side((dof_id_type)e, 1, 1);
        if (i == nx - 1) mesh.boundary_info.add_side((dof_id_type)e, 2, 2);
        if (j == ny - 1) mesh.boundary_info.add_side((dof_id_type)e, 3, 3);
        if (i == 0) mesh.boundary_info.add_side((dof_id_type)e, 4, 4);
        if (k == nz - 1) mesh.boundary_info.add_side((dof_id_type)e, 5, 5);
      }
}

void build_square(MeshBase& mesh, unsigned nx, unsigned ny, Real xmin, Real xmax, Real ymin, Real ymax, ElemType type) {
  if (type != QUAD4) libmesh_error_msg("build_square: QUAD4");
  mesh.elem_type = QUAD4;
  mesh.xyz.resize(size_t(nx + 1) * (ny + 1) * 3);
  for (unsigned j = 0; j <= ny; ++j)
    for (unsigned i = 0; i <= nx; ++i) {
      Real* p = &mesh.xyz[(size_t(i) + size_t(j) * (nx + 1)) * 3];
      p[0] = xmin + (xmax - xmin) * i / nx; p[1] = ymin + (ymax - ymin) * j / ny; p[2] = 0.;
    }
  mesh.conn.resize(size_t(nx) * ny * 4);
  mesh.subdomain.assign(size_t(nx) * ny, 0);
  for (unsigned j = 0; j < ny; ++j)
    for (unsigned i = 0; i < nx; ++i) {
      const size_t e = size_t(i) + size_t(j) * nx;
      const int32_t n0 = int32_t(i + j * (nx + 1));
      const int32_t c[4] = {n0, n0 + 1, n0 + int32_t(nx) + 2, n0 + int32_t(nx) + 1};
      std::memcpy(&mesh.conn[e * 4], c, sizeof c);
      if (j == 0) mesh.boundary_info.add_side((dof_id_type)e, 0, 0);
      if (i == nx - 1) mesh.boundary_info.add_side((dof_id_type)e, 1, 1);
      if (j == ny - 1) mesh.boundary_info.add_side((dof_id_type)e, 2, 2);
      if (i == 0) mesh.boundary_info.add_side((dof_id_type)e, 3, 3);
    }
}

}  // namespace Generation
}  // namespace MeshTools

void System::init_data() {
  solution.reset(new NumericVector<Number>(n_dofs()));
  if (constrained_dofs.size() != n_dofs()) constrained_dofs.assign(n_dofs(), 0);
}

// libMesh DofMap sparsity for one variable group: every dof of a node couples to every dof of every node sharing an
// element with it; columns sorted (what PETSc AIJ is preallocated with).
void ImplicitSystem::init_data() {
  System::init_data();
  const MeshBase& mesh = get_mesh();
  const unsigned nn = mesh.nodes_per_elem(), nv = n_vars();
  const dof_id_type n_nodes = mesh.n_nodes();
  std::vector<std::vector<int32_t>> nbr(n_nodes);
  for (dof_id_type e = 0; e < mesh.n_elem(); ++e)
    for (unsigned a = 0; a < nn; ++a) {
      std::vector<int32_t>& l = nbr[mesh.conn[size_t(e) * nn + a]];
      l.insert(l.end(), &mesh.conn[size_t(e) * nn], &mesh.conn[size_t(e) * nn] + nn);
    }
  std::vector<int64_t> row_ptr(size_t(n_nodes) * nv + 1, 0);
  std::vector<int32_t> col;
  for (dof_id_type n = 0; n < n_nodes; ++n) {
    std::vector<int32_t>& l = nbr[n];
    std::sort(l.begin(), l.end());
    l.erase(std::unique(l.begin(), l.end()), l.end());
    for (unsigned v = 0; v < nv; ++v) {
      for (int32_t m : l)
        for (unsigned w = 0; w < nv; ++w) col.push_back(m * int32_t(nv) + int32_t(w));
      row_ptr[size_t(n) * nv + v + 1] = (int64_t)col.size();
    }
  }
  _rhs.reset(new NumericVector<Number>(n_dofs()));
  _matrix.reset(new SparseMatrix<Number>());
  _matrix->init(row_ptr, col);
  rhs = _rhs.get();
  matrix = _matrix.get();
}
#endif
}  // namespace libMesh

namespace MAST {

// ---- function sets / cards / discipline ---------------------------------------------------------------------------
void FunctionSetBase::add(FunctionBase& f) {
  if (_map.count(f.name())) libmesh_error_msg("function " + f.name() + " already added");
  _map[f.name()] = &f;
}
const MaterialPropertyCardBase& ElementPropertyCardBase::get_material() const {
  if (!_mat) libmesh_error_msg("property card has no material");
  return *_mat;
}

StructuralSystemInitialization::StructuralSystemInitialization(NonlinearSystem& sys, const std::string& prefix,
                                                               const libMesh::FEType& fe)
    : SystemInitialization(sys, prefix) {
  for (const char* v : {"_ux", "_uy", "_uz", "_tx", "_ty", "_tz"}) _vars.push_back(sys.add_variable(prefix + v, fe));
}
HeatConductionSystemInitialization::HeatConductionSystemInitialization(NonlinearSystem& sys, const std::string& prefix,
                                                                       const libMesh::FEType& fe)
    : SystemInitialization(sys, prefix) {
  _vars.push_back(sys.add_variable(prefix + "_T", fe));
}

void PhysicsDisciplineBase::add_side_load(libMesh::boundary_id_type bid, BoundaryConditionBase& load) {
  if (load.type() == DIRICHLET) libmesh_error_msg("use add_dirichlet_bc for Dirichlet conditions");
  for (auto r = _side_bc.equal_range(bid); r.first != r.second; ++r.first)
    if (r.first->second == &load) libmesh_error_msg("load already added to this boundary");
  _side_bc.insert({bid, &load});
}
void PhysicsDisciplineBase::add_dirichlet_bc(libMesh::boundary_id_type bid, DirichletBoundaryCondition& load) {
  if (_dirichlet.count(bid)) libmesh_error_msg("boundary already has a Dirichlet condition");
  _dirichlet[bid] = &load;
}
void PhysicsDisciplineBase::add_volume_load(libMesh::subdomain_id_type sid, BoundaryConditionBase& load) {
  _vol_bc.insert({sid, &load});
}
void PhysicsDisciplineBase::set_property_for_subdomain(libMesh::subdomain_id_type sid, const ElementPropertyCardBase& prop) {
  if (_prop.count(sid)) libmesh_error_msg("subdomain already has a property card");
  _prop[sid] = &prop;
}
const ElementPropertyCardBase& PhysicsDisciplineBase::get_property_card(unsigned int sid) const {
  auto it = _prop.find((libMesh::subdomain_id_type)sid);
  if (it == _prop.end()) libmesh_error_msg("no property card for subdomain " + std::to_string(sid));
  return *it->second;
}
void PhysicsDisciplineBase::init_system_dirichlet_bc(NonlinearSystem& sys) const {
  const libMesh::MeshBase& mesh = sys.get_mesh();
  const unsigned nn = mesh.nodes_per_elem(), nv = sys.n_vars();
  sys.constrained_dofs.assign(size_t(mesh.n_nodes()) * nv, 0);
  // nodes of a side: libMesh side -> element node maps (outward normal ordering)
  static const int H8S[6][4] = {{0, 3, 2, 1}, {0, 1, 5, 4}, {1, 2, 6, 5}, {2, 3, 7, 6}, {3, 0, 4, 7}, {4, 5, 6, 7}};
  static const int H27S[6][9] = {{0, 3, 2, 1, 11, 10, 9, 8, 20},  {0, 1, 5, 4, 8, 13, 16, 12, 21},
                                 {1, 2, 6, 5, 9, 14, 17, 13, 22}, {2, 3, 7, 6, 10, 15, 18, 14, 23},
                                 {3, 0, 4, 7, 11, 12, 19, 15, 24}, {4, 5, 6, 7, 16, 17, 18, 19, 25}};
  static const int Q4S[4][2] = {{0, 1}, {1, 2}, {2, 3}, {3, 0}};
  for (const auto& s : mesh.get_boundary_info().sides()) {
    auto it = _dirichlet.find(s.id);
    if (it == _dirichlet.end()) continue;
    const int nsn = nn == 8 ? 4 : nn == 27 ? 9 : 2;
    for (int k = 0; k < nsn; ++k) {
      const int a = nn == 8 ? H8S[s.side][k] : nn == 27 ? H27S[s.side][k] : Q4S[s.side][k];
      const int32_t node = mesh.conn[size_t(s.elem) * nn + a];
      for (unsigned v : it->second->vars()) sys.constrained_dofs[size_t(node) * nv + v] = 1;
    }
  }
}
void PhysicsDisciplineBase::clear_system_dirichlet_bc(NonlinearSystem& sys) const { sys.constrained_dofs.assign(sys.n_dofs(), 0); }

// ---- elem-ops -----------------------------------------------------------------------------------------------------
AssemblyElemOperations::AssemblyElemOperations()
    : _system(nullptr), _discipline(nullptr), _assembly(nullptr), _physics_elem(nullptr), _skip_comm_sum(false) {}
AssemblyElemOperations::~AssemblyElemOperations() {}
SystemInitialization& AssemblyElemOperations::get_system_initialization() { libmesh_assert(_system); return *_system; }
PhysicsDisciplineBase& AssemblyElemOperations::get_discipline() { libmesh_assert(_discipline); return *_discipline; }
void AssemblyElemOperations::set_discipline_and_system(PhysicsDisciplineBase& discipline, SystemInitialization& system) {
  if (_discipline || _system) libmesh_error_msg("elem-ops: discipline and system already set");  // assembly_elem_operation.cpp:48-58
  _discipline = &discipline; _system = &system;
}
void AssemblyElemOperations::clear_discipline_and_system() { _discipline = nullptr; _system = nullptr; }
void AssemblyElemOperations::set_assembly(AssemblyBase& assembly) {
  if (_assembly) libmesh_error_msg("elem-ops: assembly already set");
  _assembly = &assembly;
}
AssemblyBase& AssemblyElemOperations::get_assembly() { libmesh_assert(_assembly); return *_assembly; }
void AssemblyElemOperations::clear_assembly() { _assembly = nullptr; }
void AssemblyElemOperations::clear_elem() { _physics_elem = nullptr; }
void AssemblyElemOperations::_host_element_path(const char* what) const {
  libmesh_error_msg(std::string(what) + ": per-element host evaluation is not part of the device path "
                    "(the element loop runs on the GPU behind AssemblyBase; there is no CPU fallback)");
}
void AssemblyElemOperations::set_elem_solution(const RealVectorX&) { _host_element_path("set_elem_solution"); }
void AssemblyElemOperations::set_elem_solution_sensitivity(const RealVectorX&) { _host_element_path("set_elem_solution_sensitivity"); }

void DeviceElemOperations::init(const GeomElem&) { _host_element_path("init"); }
void DeviceElemOperations::elem_calculations(bool, RealVectorX&, RealMatrixX&) { _host_element_path("elem_calculations"); }
void DeviceElemOperations::elem_linearized_jacobian_solution_product(RealVectorX&) { _host_element_path("elem_linearized_jacobian_solution_product"); }
void DeviceElemOperations::elem_sensitivity_calculations(const FunctionBase&, RealVectorX&) { _host_element_path("elem_sensitivity_calculations"); }
void DeviceElemOperations::elem_shape_sensitivity_calculations(const FunctionBase&, RealVectorX&) { _host_element_path("elem_shape_sensitivity_calculations"); }
void DeviceElemOperations::elem_topology_sensitivity_calculations(const FunctionBase&, RealVectorX&) { _host_element_path("elem_topology_sensitivity_calculations"); }
void DeviceElemOperations::elem_second_derivative_dot_solution_assembly(RealMatrixX&) { _host_element_path("elem_second_derivative_dot_solution_assembly"); }

void ComplianceOutput::init(const GeomElem&) { _host_element_path("ComplianceOutput::init"); }
void ComplianceOutput::evaluate() { _host_element_path("ComplianceOutput::evaluate"); }
void ComplianceOutput::evaluate_sensitivity(const FunctionBase&) { _host_element_path("ComplianceOutput::evaluate_sensitivity"); }
void ComplianceOutput::output_derivative_for_elem(RealVectorX&) { _host_element_path("ComplianceOutput::output_derivative_for_elem"); }

// ---- AssemblyBase -------------------------------------------------------------------------------------------------
AssemblyBase::AssemblyBase()
    : close_matrix(true), _elem_ops(nullptr), _discipline(nullptr), _system(nullptr), _param_dependence(nullptr),
      _ctx(nullptr), _device_id(0), _pattern_checked(false), _rho_field(nullptr) {}
AssemblyBase::~AssemblyBase() {
  if (_ctx) mastb200_destroy(_ctx);
}
const PhysicsDisciplineBase& AssemblyBase::discipline() const { libmesh_assert(_discipline); return *_discipline; }
PhysicsDisciplineBase& AssemblyBase::discipline() { libmesh_assert(_discipline); return *_discipline; }
AssemblyElemOperations& AssemblyBase::get_elem_ops() { libmesh_assert(_elem_ops); return *_elem_ops; }
const NonlinearSystem& AssemblyBase::system() const { libmesh_assert(_system); return _system->system(); }
NonlinearSystem& AssemblyBase::system() { libmesh_assert(_system); return _system->system(); }
SystemInitialization& AssemblyBase::system_init() { libmesh_assert(_system); return *_system; }

void AssemblyBase::set_discipline_and_system(PhysicsDisciplineBase& discipline, SystemInitialization& system) {
  if (_discipline || _system) libmesh_error_msg("Error: Assembly should be cleared before attaching System.");  // assembly_base.cpp:160-172
  _discipline = &discipline; _system = &system;
}
void AssemblyBase::clear_discipline_and_system() {
  _discipline = nullptr; _system = nullptr;
  if (_ctx) { mastb200_destroy(_ctx); _ctx = nullptr; }  // the device problem belongs to the (discipline, system) pair
  _pattern_checked = false; _rho_field = nullptr;
}
void AssemblyBase::attach_elem_parameter_dependence_object(ElemParameterDependence& dep) {
  if (_param_dependence) libmesh_error_msg("parameter dependence object already attached");
  _param_dependence = &dep;
}
void AssemblyBase::clear_elem_parameter_dependence_object() { _param_dependence = nullptr; }
void AssemblyBase::set_elem_operation_object(AssemblyElemOperations& elem_ops) {
  if (_elem_ops) libmesh_error_msg("Error: Assembly should be cleared before attaching Elem.");  // assembly_base.cpp:186-194
  _elem_ops = &elem_ops;
  elem_ops.set_assembly(*this);
}
void AssemblyBase::clear_elem_operation_object() {
  if (_elem_ops) { _elem_ops->clear_assembly(); _elem_ops = nullptr; }
}
std::unique_ptr<libMesh::NumericVector<Real>> AssemblyBase::build_localized_vector(const libMesh::System&,
                                                                                 const libMesh::NumericVector<Real>& global) const {
  // serial shim: the localized vector is a copy (assembly_base.cpp:209-228 scatters owned + send_list entries)
  std::unique_ptr<libMesh::NumericVector<Real>> v(new libMesh::NumericVector<Real>(global.size()));
  std::memcpy(v->get_array(), global.get_array_read(), sizeof(Real) * global.size());
  return v;
}
void AssemblyBase::_device_error(const char* what) const {
  libmesh_error_msg(std::string(what) + ": " + mastb200_last_error(_ctx));
}
void AssemblyBase::_check(int rc, const char* what) const {
  if (rc != MASTB200_OK) _device_error(what);
}
long long AssemblyBase::device_kernel_launches() const {
  if (!_ctx) return 0;
  mastb200_stats s;
  mastb200_get_stats(_ctx, &s);
  return s.kernel_launches;
}

namespace {
Real constant_of(const FunctionSetBase& set, const char* nm, bool required, Real dflt = 0.) {
  if (!set.contains(nm)) {
    if (required) libmesh_error_msg(std::string("property card has no function named ") + nm);
    return dflt;
  }
  const ConstantFieldFunction* c = set.get_if<ConstantFieldFunction>(nm);
  if (!c) libmesh_error_msg(std::string(nm) + ": only ConstantFieldFunction is on the device path for this property");
  return c->value();
}
}  // namespace

mastb200_context* AssemblyBase::_device_problem(const libMesh::SparseMatrix<Real>* J) {
  if (!_discipline || !_system) libmesh_error_msg("set_discipline_and_system first");
  if (!_elem_ops) libmesh_error_msg("set_elem_operation_object first");
  const bool structural = dynamic_cast<StructuralNonlinearAssemblyElemOperations*>(_elem_ops) != nullptr;
  const bool thermal = dynamic_cast<HeatConductionNonlinearAssemblyElemOperations*>(_elem_ops) != nullptr;
  if (!structural && !thermal && !dynamic_cast<OutputAssemblyElemOperations*>(_elem_ops))
    libmesh_error_msg("this AssemblyElemOperations type has no device sweep (closed set: structural solids, heat conduction)");
  NonlinearSystem& sys = system();
  const libMesh::MeshBase& mesh = sys.get_mesh();
  const unsigned nv = sys.n_vars();
  const ElementPropertyCardBase& card = _discipline->get_property_card(0);
  for (auto sid : mesh.subdomain)
    if (sid != 0 && !diagonal_elem_subdomain_id.count(sid))
      libmesh_error_msg("device path: one property card (subdomain 0) plus diagonal_elem_subdomain_id subdomains");
  const MaterialPropertyCardBase& mat = card.get_material();

  const bool fresh = _ctx == nullptr;
  if (!_ctx) {
    if (mastb200_create(_device_id, &_ctx) != MASTB200_OK) {
      _ctx = nullptr;
      libmesh_error_msg(std::string("mastb200_create: ") + mastb200_last_error(nullptr));
    }
    const int et = mesh.elem_type == libMesh::HEX8 ? MASTB200_HEX8 : mesh.elem_type == libMesh::HEX27 ? MASTB200_HEX27 : MASTB200_QUAD4;
    std::vector<uint8_t> diag;
    if (!diagonal_elem_subdomain_id.empty()) {
      diag.resize(mesh.n_elem());
      for (size_t e = 0; e < diag.size(); ++e) diag[e] = diagonal_elem_subdomain_id.count(mesh.subdomain[e]) ? 1 : 0;
    }
    _check(mastb200_set_mesh(_ctx, et, mesh.n_elem(), mesh.n_nodes(), mesh.xyz.data(), mesh.conn.data(),
                             diag.empty() ? nullptr : diag.data()), "set_mesh");
    _check(mastb200_set_dof_map(_ctx, (int32_t)nv, sys.constrained_dofs.data(), nullptr), "set_dof_map");
    int64_t n_rows = 0, nnz = 0;
    _check(mastb200_build_pattern(_ctx, &n_rows, &nnz), "build_pattern");
  }

  // loads: SURFACE_PRESSURE sides (structural_element_base.cpp:628-701), heat-conduction side loads
  // (heat_conduction_elem_base.cpp:428-483) and point loads (nonlinear_implicit_assembly.cpp:190-247).  Their values are
  // Parameters the caller may change between sweeps (load stepping): gathered every time, re-sent when they differ
  {
    std::vector<int32_t> se, ss, pn;
    std::vector<double> sp, hc, ta, em, pl, sig;
    bool any_heat = false, any_rad = false;
    double sbc = 0., T0 = 0.;
    for (const auto& kv : _discipline->side_loads()) {
      const BoundaryConditionType t = kv.second->type();
      double v = 0., c = 0., a = 0., e = 0.;
      if (t == SURFACE_PRESSURE) v = constant_of(*kv.second, "pressure", true);
      else if (t == HEAT_FLUX) { v = constant_of(*kv.second, "heat_flux", true); any_heat = true; }
      else if (t == CONVECTION_HEAT_FLUX) {
        c = constant_of(*kv.second, "convection_coeff", true); a = constant_of(*kv.second, "ambient_temperature", true); any_heat = true;
      } else if (t == SURFACE_RADIATION_HEAT_FLUX) {
        e = constant_of(*kv.second, "emissivity", true); a = constant_of(*kv.second, "ambient_temperature", true);
        sbc = constant_of(*kv.second, "stefan_bolzmann_constant", true); T0 = constant_of(*kv.second, "reference_zero_temperature", true);
        any_heat = any_rad = true;
      } else libmesh_error_msg("side load type has no device path");
      for (const auto& sd : mesh.get_boundary_info().sides())
        if (sd.id == kv.first) {
          se.push_back((int32_t)sd.elem); ss.push_back(sd.side); sp.push_back(v); hc.push_back(c); ta.push_back(a); em.push_back(e);
        }
    }
    for (const PointLoadCondition* pc : _discipline->point_loads()) {
      const ConstantVectorFieldFunction* f = pc->get_if<ConstantVectorFieldFunction>("load");
      if (!f || f->values.size() != nv) libmesh_error_msg("point load: a ConstantVectorFieldFunction \"load\" with one value per variable");
      for (libMesh::dof_id_type n : pc->get_nodes()) { pn.push_back((int32_t)n); pl.insert(pl.end(), f->values.begin(), f->values.end()); }
    }
    sig = sp; sig.insert(sig.end(), hc.begin(), hc.end()); sig.insert(sig.end(), ta.begin(), ta.end());
    sig.insert(sig.end(), em.begin(), em.end()); sig.insert(sig.end(), pl.begin(), pl.end());
    sig.push_back(double(se.size())); sig.push_back(double(pn.size())); sig.push_back(sbc); sig.push_back(T0);
    if (fresh || sig != _side_values) {
      if (any_heat) {
        if (structural) libmesh_error_msg("heat side loads on a structural system");
        _check(mastb200_set_heat_side_loads_ex(_ctx, (int64_t)se.size(), se.data(), ss.data(), sp.data(), hc.data(), ta.data(),
                                               any_rad ? em.data() : nullptr, sbc, T0), "set_heat_side_loads");
      } else if (!se.empty() || !fresh) {
        _check(mastb200_set_side_pressure(_ctx, (int64_t)se.size(), se.data(), ss.data(), sp.data()), "set_side_pressure");
      }
      if (!pn.empty() || !fresh) _check(mastb200_set_point_loads(_ctx, (int64_t)pn.size(), pn.data(), pl.data()), "set_point_loads");
      _side_values = sig;
    }
  }

  // values that may change between sweeps (parameters, density field): refreshed every time, cheap
  mastb200_material m;
  std::memset(&m, 0, sizeof m);
  m.penalty = 1.; m.thickness = 1.;
  m.nonlinear_strain = card.strain_type() == NONLINEAR_STRAIN ? 1 : 0;
  _rho_field = nullptr;
  if (structural || !thermal) {
    m.nu = constant_of(mat, "nu", true);
    m.alpha = constant_of(mat, "alpha_expansion", false);
    if (const SIMPModulusFieldFunction* simp = mat.get_if<SIMPModulusFieldFunction>("E")) {
      m.E0 = simp->E0; m.rho_min = simp->rho_min; m.penalty = simp->penalty;
      _rho_field = &simp->rho;
      if (_rho_field->values.size() != mesh.n_nodes()) libmesh_error_msg("density field size != number of nodes");
    } else {
      m.E0 = constant_of(mat, "E", true);
    }
    if (mesh.elem_type == libMesh::QUAD4) {  // StructuralElement2D: solid 2D section (solid_2d_section_element_property_card.cpp)
      m.thickness = constant_of(card, "h", true);
      m.offset = constant_of(card, "off", false, 0.);
      m.kappa = constant_of(card, "kappa", false, 5. / 6.);
      const Solid2DSectionElementPropertyCard* c2 = dynamic_cast<const Solid2DSectionElementPropertyCard*>(&card);
      m.plane_strain = (c2 && !c2->plane_stress()) ? 1 : 0;
    }
  } else {
    if (const LinearTemperatureFieldFunction* k = mat.get_if<LinearTemperatureFieldFunction>("k_th")) {
      m.k0 = k->k0; m.k_beta = k->beta; m.k_of_T = 1;
    } else {
      m.k0 = constant_of(mat, "k_th", true);
    }
    m.thickness = constant_of(card, "h", false, 1.);
  }
  bool have_T = false;
  for (const auto& kv : _discipline->volume_loads()) {
    if (kv.second->type() == HEAT_SOURCE) m.heat_source = constant_of(*kv.second, "heat_source", true);
    else if (kv.second->type() == TEMPERATURE) {
      const Real Tref = constant_of(*kv.second, "ref_temperature", true);
      if (const NodalFieldFunction* T = kv.second->get_if<NodalFieldFunction>("temperature"))
        _check(mastb200_set_temperature(_ctx, 1, T->values.data(), 0., Tref), "set_temperature");
      else
        _check(mastb200_set_temperature(_ctx, 1, nullptr, constant_of(*kv.second, "temperature", true), Tref), "set_temperature");
      have_T = true;
    } else libmesh_error_msg("volume load type has no device path (TEMPERATURE, HEAT_SOURCE)");
  }
  if (!have_T) _check(mastb200_set_temperature(_ctx, 0, nullptr, 0., 0.), "set_temperature");
  _check(mastb200_set_material(_ctx, &m), "set_material");
  _check(mastb200_set_density(_ctx, _rho_field ? _rho_field->values.data() : nullptr), "set_density");

  if (J && !_pattern_checked) {
    // the values are written straight into J's CSR storage: its pattern must be the library's, bit for bit
    std::vector<int64_t> rp(size_t(sys.n_dofs()) + 1);
    std::vector<int32_t> ci(J->n_nonzeros());
    int64_t n_rows = 0, nnz = 0;
    (void)n_rows;
    _check(mastb200_get_pattern(_ctx, rp.data(), nullptr), "get_pattern");
    nnz = rp.back();
    if ((size_t)nnz != J->n_nonzeros()) libmesh_error_msg("Jacobian sparsity differs from the device pattern (nnz)");
    _check(mastb200_get_pattern(_ctx, rp.data(), ci.data()), "get_pattern");
    if (rp != J->row_ptr() || ci != J->col_idx()) libmesh_error_msg("Jacobian sparsity differs from the device pattern");
    _pattern_checked = true;
  }
  return _ctx;
}

void StressStrainOutputBase::init(const GeomElem&) { _host_element_path("StressStrainOutputBase::init"); }
void StressStrainOutputBase::evaluate() { _host_element_path("StressStrainOutputBase::evaluate"); }
void StressStrainOutputBase::evaluate_sensitivity(const FunctionBase&) { _host_element_path("StressStrainOutputBase::evaluate_sensitivity"); }
void StressStrainOutputBase::output_derivative_for_elem(RealVectorX&) { _host_element_path("StressStrainOutputBase::output_derivative_for_elem"); }
Real StressStrainOutputBase::output_sensitivity_total(const FunctionBase&) { return _dq_dp; }
// StressStrainOutputBase::functional_for_all_elems (stress_output_base.cpp:851-909); with skip_comm_sum the caller sums
// the two integrals over the ranks first (here: one rank per context, the sums are this rank's)
Real StressStrainOutputBase::output_total() {
  if (!_primal_data_initialized) libmesh_error_msg("StressStrainOutputBase::output_total before calculate_output");
  return _sigma0 * std::pow(_sigma_vm_int / _JxW_val, 1. / _p_norm_stress);
}

// AssemblyBase::calculate_output (assembly_base.cpp:253-333) for ComplianceOutput (compliance_output.cpp:71-102) and
// StressStrainOutputBase (stress_output_base.cpp:318-331, 851-909)
void AssemblyBase::calculate_output(const libMesh::NumericVector<Real>& X, bool, OutputAssemblyElemOperations& output) {
  if (StressStrainOutputBase* so = dynamic_cast<StressStrainOutputBase*>(&output)) {
    if (so->_p_norm_weight != 1.) libmesh_error_msg("StressStrainOutputBase: p_norm_weight != 1 is not on the device path");
    AssemblyElemOperations* saved0 = _elem_ops;
    if (!_elem_ops) _elem_ops = &output;
    mastb200_context* c0 = _device_problem(nullptr);
    _elem_ops = saved0;
    so->zero_for_analysis();
    double sums[2] = {0., 0.};
    _check(mastb200_stress_functional(c0, X.get_array_read(), so->_p_norm_stress, so->_rho, so->_sigma0, nullptr, sums, nullptr),
           "stress_functional");
    so->_sigma_vm_int = sums[0]; so->_JxW_val = sums[1];
    so->_primal_data_initialized = true;
    return;
  }
  ComplianceOutput* c = dynamic_cast<ComplianceOutput*>(&output);
  if (!c) libmesh_error_msg("calculate_output: only ComplianceOutput and StressStrainOutputBase are on the device path");
  AssemblyElemOperations* saved = _elem_ops;
  if (!_elem_ops) _elem_ops = &output;
  mastb200_context* ctx = _device_problem(nullptr);
  _elem_ops = saved;
  c->zero_for_analysis();
  double q = 0.;
  _check(mastb200_compliance(ctx, X.get_array_read(), &q), "compliance");
  c->_compliance = q;
}

// AssemblyBase::calculate_output_derivative (assembly_base.cpp:338-430): dq_dX zeroed, element derivatives added, closed
void AssemblyBase::calculate_output_derivative(const libMesh::NumericVector<Real>& X, bool, OutputAssemblyElemOperations& output,
                                               libMesh::NumericVector<Real>& dq_dX) {
  if (StressStrainOutputBase* so = dynamic_cast<StressStrainOutputBase*>(&output)) {
    // functional_state_derivartive_for_elem (stress_output_base.cpp:1104-1172) needs the primal sums of calculate_output
    if (!so->_primal_data_initialized) libmesh_error_msg("calculate_output_derivative: call calculate_output first");
    AssemblyElemOperations* saved0 = _elem_ops;
    if (!_elem_ops) _elem_ops = &output;
    mastb200_context* c0 = _device_problem(nullptr);
    _elem_ops = saved0;
    const double sums[2] = {so->_sigma_vm_int, so->_JxW_val};
    _check(mastb200_stress_functional_derivative(c0, X.get_array_read(), so->_p_norm_stress, so->_rho, so->_sigma0, nullptr, sums,
                                                 dq_dX.get_array()), "stress_functional_derivative");
    dq_dX.close();
    return;
  }
  if (!dynamic_cast<ComplianceOutput*>(&output)) libmesh_error_msg("calculate_output_derivative: only ComplianceOutput and StressStrainOutputBase are on the device path");
  AssemblyElemOperations* saved = _elem_ops;
  if (!_elem_ops) _elem_ops = &output;
  mastb200_context* ctx = _device_problem(nullptr);
  _elem_ops = saved;
  _check(mastb200_compliance_derivative(ctx, X.get_array_read(), dq_dX.get_array()), "compliance_derivative");
  dq_dX.close();
}

int AssemblyBase::_card_parameter(const FunctionBase& f) const {
  if (!_discipline) return -1;
  const MaterialPropertyCardBase& mat = _discipline->get_property_card(0).get_material();
  if (const ConstantFieldFunction* c = mat.get_if<ConstantFieldFunction>("E"))
    if (&c->parameter() == &f) return MASTB200_PARAM_E;
  if (const ConstantFieldFunction* c = mat.get_if<ConstantFieldFunction>("nu"))
    if (&c->parameter() == &f) return MASTB200_PARAM_NU;
  // heat conduction: the constant conductivity of the material card and the thickness of the 2D section card
  if (const ConstantFieldFunction* c = mat.get_if<ConstantFieldFunction>("k_th"))
    if (&c->parameter() == &f) return MASTB200_PARAM_K0;
  if (const ConstantFieldFunction* c = _discipline->get_property_card(0).get_if<ConstantFieldFunction>("h"))
    if (&c->parameter() == &f) return MASTB200_PARAM_THICKNESS;
  return -1;
}

// AssemblyBase::calculate_output_direct_sensitivity (assembly_base.cpp:434-550): output.evaluate_sensitivity(p) over the
// elements = the partial derivative at fixed state plus (dq/dX) . dXdp
void AssemblyBase::calculate_output_direct_sensitivity(const libMesh::NumericVector<Real>& X, bool,
                                                       const libMesh::NumericVector<Real>* dXdp, bool, const FunctionBase& p,
                                                       OutputAssemblyElemOperations& output) {
  AssemblyElemOperations* saved = _elem_ops;
  if (!_elem_ops) _elem_ops = &output;
  mastb200_context* ctx = _device_problem(nullptr);
  _elem_ops = saved;
  output.zero_for_sensitivity();
  if (ComplianceOutput* c = dynamic_cast<ComplianceOutput*>(&output)) {
    // compliance_output.cpp:106-150: -d(f_ext)/dp . u - f_ext . du/dp; the device loads do not depend on a density DV or a
    // card parameter, a load's own Parameter is not on the device path
    if (!dynamic_cast<const DensityDesignVariable*>(&p) && _card_parameter(p) < 0)
      libmesh_error_msg("calculate_output_direct_sensitivity: parameter " + p.name() + " has no device path");
    double dq = 0.;
    _check(mastb200_compliance_direct_sensitivity(ctx, X.get_array_read(), dXdp ? dXdp->get_array_read() : nullptr, &dq),
           "compliance_direct_sensitivity");
    c->_dcompliance_dp = dq;
    return;
  }
  if (StressStrainOutputBase* so = dynamic_cast<StressStrainOutputBase*>(&output)) {
    if (!so->_primal_data_initialized) libmesh_error_msg("calculate_output_direct_sensitivity: call calculate_output first");
    const DensityDesignVariable* dv = dynamic_cast<const DensityDesignVariable*>(&p);
    if (!dv) libmesh_error_msg("stress output sensitivity: density design variables only");
    const double sums[2] = {so->_sigma_vm_int, so->_JxW_val};
    const size_t nn = system().get_mesh().n_nodes();
    double dq = 0.;
    if (_rho_field) {  // partial derivative at fixed state: W^T (dq / d rho_n)   (stress_output_base.cpp:911-1043)
      std::vector<double> G(nn);
      _check(mastb200_stress_functional_density_derivative(ctx, X.get_array_read(), so->_p_norm_stress, so->_rho, so->_sigma0, nullptr,
                                                           sums, G.data()), "stress_functional_density_derivative");
      const int64_t ptr[2] = {0, (int64_t)dv->nodes.size()};
      _check(mastb200_apply_dv_map_transpose(ctx, G.data(), 1, ptr, dv->nodes.data(), dv->weights.data(), &dq), "apply_dv_map_transpose");
    }
    if (dXdp) {
      std::vector<double> g(system().n_dofs());
      _check(mastb200_stress_functional_derivative(ctx, X.get_array_read(), so->_p_norm_stress, so->_rho, so->_sigma0, nullptr, sums,
                                                   g.data()), "stress_functional_derivative");
      const double* d = dXdp->get_array_read();
      for (size_t i = 0; i < g.size(); ++i) dq += g[i] * d[i];
    }
    so->_dq_dp = dq;
    return;
  }
  libmesh_error_msg("calculate_output_direct_sensitivity: only ComplianceOutput and StressStrainOutputBase are on the device path");
}

Real AssemblyBase::calculate_output_adjoint_sensitivity(const libMesh::NumericVector<Real>& X, bool if_localize_sol,
                                                        const libMesh::NumericVector<Real>& adj_sol, const FunctionBase& p,
                                                        AssemblyElemOperations& elem_ops, OutputAssemblyElemOperations& output,
                                                        const bool include_partial_sens) {
  // dq/dp = adj . dR/dp [+ partial q / partial p]   (assembly_base.cpp:555-590)
  Real dq_dp = 0.;
  if (dynamic_cast<const DensityDesignVariable*>(&p)) {
    std::vector<const FunctionBase*> pv(1, &p);
    std::vector<Real> sens;
    calculate_output_adjoint_sensitivity_multiple_parameters_no_direct(X, if_localize_sol, adj_sol, pv, elem_ops, output, sens);
    dq_dp = sens[0];
  } else {  // a Parameter of the card: the reference's own recipe, sensitivity_assemble + dot
    libMesh::NumericVector<Real> dres(system().n_dofs());
    AssemblyElemOperations* saved = _elem_ops;
    _elem_ops = &elem_ops;
    const bool ok = this->sensitivity_assemble(X, if_localize_sol, p, dres);
    _elem_ops = saved;
    if (!ok) libmesh_error_msg("parameter " + p.name() + ": no device sensitivity path");
    const double *a = adj_sol.get_array_read(), *d = dres.get_array_read();
    for (size_t i = 0; i < dres.size(); ++i) dq_dp += a[i] * d[i];
  }
  if (include_partial_sens) {  // :578-586: the partial sensitivity of the output, with zero solution sensitivity
    calculate_output_direct_sensitivity(X, if_localize_sol, nullptr, false, p, output);
    dq_dp += output.output_sensitivity_total(p);
  }
  return dq_dp;
}

void AssemblyBase::calculate_output_adjoint_sensitivity_multiple_parameters_no_direct(
    const libMesh::NumericVector<Real>& X, bool, const libMesh::NumericVector<Real>& adj_sol,
    const std::vector<const FunctionBase*>& p_vec, AssemblyElemOperations& elem_ops, OutputAssemblyElemOperations&,
    std::vector<Real>& sens) {
  // reference (assembly_base.cpp:594-639): n_dv sensitivity_assemble sweeps + dots.  Here: one fused sweep for the nodal
  // gradient G_n = adj . dR/d rho_n, then sens = W^T G on the device.
  AssemblyElemOperations* saved = _elem_ops;
  _elem_ops = &elem_ops;
  mastb200_context* ctx = _device_problem(nullptr);
  _elem_ops = saved;
  std::vector<int64_t> ptr(1, 0);
  std::vector<int32_t> nodes;
  std::vector<double> w;
  for (const FunctionBase* f : p_vec) {
    const DensityDesignVariable* dv = dynamic_cast<const DensityDesignVariable*>(f);
    if (!dv) libmesh_error_msg("parameter " + f->name() + " is not a DensityDesignVariable: no device sensitivity path");
    nodes.insert(nodes.end(), dv->nodes.begin(), dv->nodes.end());
    w.insert(w.end(), dv->weights.begin(), dv->weights.end());
    ptr.push_back((int64_t)nodes.size());
  }
  std::vector<double> G(system().get_mesh().n_nodes());
  _check(mastb200_sensitivity_sweep(ctx, X.get_array_read(), adj_sol.get_array_read(), G.data(), nullptr, nullptr), "sensitivity_sweep");
  sens.assign(p_vec.size(), 0.);
  _check(mastb200_apply_dv_map_transpose(ctx, G.data(), (int64_t)p_vec.size(), ptr.data(), nodes.data(), w.data(), sens.data()),
         "apply_dv_map_transpose");
}

// ---- NonlinearImplicitAssembly ------------------------------------------------------------------------------------
NonlinearImplicitAssembly::NonlinearImplicitAssembly() : _post_assembly(nullptr), _res_l2_norm(0.) {}
NonlinearImplicitAssembly::~NonlinearImplicitAssembly() {}

void NonlinearImplicitAssembly::residual_and_jacobian(const libMesh::NumericVector<Real>& X, libMesh::NumericVector<Real>* R,
                                                      libMesh::SparseMatrix<Real>* J, libMesh::NonlinearImplicitSystem& S) {
  // nonlinear_implicit_assembly.cpp:64-266: zero R and J, localize X, element loop, close, ||R||
  (void)S;
  mastb200_context* ctx = _device_problem(J);
  _check(mastb200_residual_and_jacobian(ctx, X.get_array_read(), R ? R->get_array() : nullptr, J ? J->values() : nullptr),
         "residual_and_jacobian");
  if (_post_assembly) _post_assembly->post_assembly(X, R, J, S);
  if (R) {
    R->close();
    double nrm = 0.;
    _check(mastb200_res_l2_norm(ctx, &nrm), "res_l2_norm");
    _res_l2_norm = nrm;
  }
  if (J && close_matrix) J->close();
}

// NonlinearImplicitAssembly::sensitivity_assemble (nonlinear_implicit_assembly.cpp:515-686): the global dR/dp vector of one
// parameter.  Device paths: a density design variable (SIMP), the card's E / nu Parameters (3D solids) and the conductivity /
// section thickness Parameters of a QUAD4 heat-conduction system;
// for anything else the reference's answer "false => the caller may finite-difference" (assembly_base.h:241-244).
bool NonlinearImplicitAssembly::sensitivity_assemble(const libMesh::NumericVector<Real>& X, bool, const FunctionBase& f,
                                                     libMesh::NumericVector<Real>& sensitivity_rhs, bool close_vector) {
  mastb200_context* ctx = _device_problem(nullptr);
  int rc;
  if (const DensityDesignVariable* dv = dynamic_cast<const DensityDesignVariable*>(&f)) {
    rc = mastb200_sensitivity_assemble(ctx, X.get_array_read(), MASTB200_PARAM_DENSITY_DV, (int64_t)dv->nodes.size(), dv->nodes.data(),
                                       dv->weights.data(), sensitivity_rhs.get_array());
  } else {
    const int kind = _card_parameter(f);
    if (kind < 0) return false;
    rc = mastb200_sensitivity_assemble(ctx, X.get_array_read(), kind, 0, nullptr, nullptr, sensitivity_rhs.get_array());
  }
  if (rc == MASTB200_ERR_UNSUPPORTED) return false;
  _check(rc, "sensitivity_assemble");
  if (close_vector) sensitivity_rhs.close();
  return true;
}

void NonlinearImplicitAssembly::linearized_jacobian_solution_product(const libMesh::NumericVector<Real>& X,
                                                                     const libMesh::NumericVector<Real>& dX,
                                                                     libMesh::NumericVector<Real>& JdX, libMesh::NonlinearImplicitSystem&) {
  mastb200_context* ctx = _device_problem(nullptr);
  _check(mastb200_linearized_jacobian_solution_product(ctx, X.get_array_read(), dX.get_array_read(), JdX.get_array()),
         "linearized_jacobian_solution_product");
  JdX.close();
}

void NonlinearImplicitAssembly::second_derivative_dot_solution_assembly(const libMesh::NumericVector<Real>&,
                                                                        const libMesh::NumericVector<Real>&,
                                                                        libMesh::SparseMatrix<Real>& d_JdX_dX,
                                                                        libMesh::NonlinearImplicitSystem&) {
  // d(J dX)/dX: with LINEAR_STRAIN the tangent does not depend on the state, the matrix is identically zero; the
  // Green-Lagrange case needs the third derivative of the strain energy and has no device path
  if (_discipline && _discipline->get_property_card(0).strain_type() == NONLINEAR_STRAIN)
    libmesh_error_msg("second_derivative_dot_solution_assembly with NONLINEAR_STRAIN has no device path");
  std::memset(d_JdX_dX.values(), 0, sizeof(Real) * d_JdX_dX.n_nonzeros());
  if (close_matrix) d_JdX_dX.close();
}

}  // namespace MAST
