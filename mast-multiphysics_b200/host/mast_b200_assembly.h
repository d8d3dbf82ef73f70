// mast_b200_assembly.h -- host-side mirror of MAST's assembly / elem-ops / output classes for the device hot path.
//
// Same class names, method names, argument meaning, attach/clear pairing and error behaviour as the reference
// (paths relative to /root/reference):
//   MAST::AssemblyBase                              src/base/assembly_base.h:51-331
//   MAST::NonlinearImplicitAssembly                 src/base/nonlinear_implicit_assembly.h, .cpp:64-266, 515-686
//   MAST::AssemblyElemOperations                    src/base/assembly_elem_operation.h:45-205
//   MAST::NonlinearImplicitAssemblyElemOperations   src/base/nonlinear_implicit_assembly_elem_operations.h:50-106
//   MAST::StructuralNonlinearAssemblyElemOperations src/elasticity/structural_nonlinear_assembly.h:36-155
//   MAST::HeatConductionNonlinearAssemblyElemOperations src/heat_conduction/heat_conduction_nonlinear_assembly.h
//   MAST::OutputAssemblyElemOperations              src/base/output_assembly_elem_operations.h
//   MAST::ComplianceOutput                          src/elasticity/compliance_output.h, .cpp:71-296
//   MAST::PhysicsDisciplineBase / SystemInitialization / property cards / boundary conditions: only the members the
//   example drivers use to describe a problem (examples/structural/example_1/example_1.cpp:52-170).
//
// What is different underneath: the element loop does not run on the host.  The assembly recognises the concrete
// elem-ops (closed set: structural 3D solids, heat conduction) and the closed set of property field functions
// (constant, SIMP of a nodal density, k0(1+beta T)), pulls plain arrays out of the libMesh objects once per mesh and
// calls the C-ABI (include/mast_b200.h) once per sweep.  Anything outside that closed set raises libMesh::LogicError
// (libmesh_error_msg) -- there is no host fallback, the per-element virtuals (elem_calculations, ...) exist for API
// compatibility and throw when called.
#pragma once
#include <map>
#include <memory>
#include <set>
#include <string>
#include <vector>

#include "libmesh_shim.h"

struct mastb200_context;

typedef libMesh::Real Real;
typedef std::vector<Real> RealVectorX;      // Eigen::Matrix<Real, Dynamic, 1> in the reference (mast_data_types.h:32-40)
struct RealMatrixX { int rows = 0, cols = 0; std::vector<Real> a; };

namespace MAST {

// ---- functions / parameters (src/base/function_base.h, parameter.h, constant_field_function.h) -------------------
class FunctionBase {
 public:
  explicit FunctionBase(const std::string& nm, bool is_field = false) : _name(nm), _is_field(is_field) {}
  virtual ~FunctionBase() {}
  const std::string& name() const { return _name; }
 protected:
  std::string _name;
  bool _is_field;
};
class Parameter : public FunctionBase {
 public:
  Parameter(const std::string& nm, Real v) : FunctionBase(nm), _val(v) {}
  Real operator()() const { return _val; }
  void operator=(Real v) { _val = v; }
 private:
  Real _val;
};
template <typename ValType>
class FieldFunction : public FunctionBase {
 public:
  explicit FieldFunction(const std::string& nm) : FunctionBase(nm, true) {}
};
class ConstantFieldFunction : public FieldFunction<Real> {
 public:
  ConstantFieldFunction(const std::string& nm, const Parameter& p) : FieldFunction<Real>(nm), _p(p) {}
  Real value() const { return _p(); }
  const Parameter& parameter() const { return _p; }  // the Parameter sensitivity_assemble is asked about
 private:
  const Parameter& _p;
};
// constant vector field (a point load's "load": one value per system variable)
class ConstantVectorFieldFunction : public FieldFunction<RealVectorX> {
 public:
  ConstantVectorFieldFunction(const std::string& nm, RealVectorX v) : FieldFunction<RealVectorX>(nm), values(std::move(v)) {}
  RealVectorX values;
};
// nodal scalar field (the role of MAST::MeshFieldFunction over the density system, src/base/mesh_field_function.cpp:72-100)
class NodalFieldFunction : public FieldFunction<Real> {
 public:
  NodalFieldFunction(const std::string& nm, std::vector<Real> v) : FieldFunction<Real>(nm), values(std::move(v)) {}
  std::vector<Real> values;
};
// E(x) = E0*(rho_min + (1-rho_min)*rho(x)^penalty): examples/structural/example_6/example_6.cpp:100-137
class SIMPModulusFieldFunction : public FieldFunction<Real> {
 public:
  SIMPModulusFieldFunction(const std::string& nm, Real E0_, Real rho_min_, Real penalty_, NodalFieldFunction& rho_)
      : FieldFunction<Real>(nm), E0(E0_), rho_min(rho_min_), penalty(penalty_), rho(rho_) {}
  Real E0, rho_min, penalty;
  NodalFieldFunction& rho;
};
// k(T) = k0*(1 + beta*T): a temperature-dependent conductivity evaluated through the attached solution function
// (src/heat_conduction/heat_conduction_elem_base.cpp:88-90, 125-141)
class LinearTemperatureFieldFunction : public FieldFunction<Real> {
 public:
  LinearTemperatureFieldFunction(const std::string& nm, Real k0_, Real beta_) : FieldFunction<Real>(nm), k0(k0_), beta(beta_) {}
  Real k0, beta;
};
// a density design variable: rho_n = sum_i W[n,i] d_i; this object is column i of W (src/level_set/filter_base.cpp:59-97)
class DensityDesignVariable : public FunctionBase {
 public:
  DensityDesignVariable(const std::string& nm, std::vector<int32_t> n, std::vector<Real> w)
      : FunctionBase(nm), nodes(std::move(n)), weights(std::move(w)) {}
  std::vector<int32_t> nodes;
  std::vector<Real> weights;
};

class FunctionSetBase {  // src/base/function_set_base.h
 public:
  virtual ~FunctionSetBase() {}
  void add(FunctionBase& f);
  bool contains(const std::string& nm) const { return _map.count(nm) != 0; }
  template <typename T>
  const T* get_if(const std::string& nm) const {
    auto it = _map.find(nm);
    return it == _map.end() ? nullptr : dynamic_cast<const T*>(it->second);
  }
 protected:
  std::map<std::string, FunctionBase*> _map;
};

enum BoundaryConditionType {  // src/base/boundary_condition_base.h
  SURFACE_PRESSURE, TEMPERATURE, HEAT_SOURCE, DIRICHLET, POINT_LOAD, HEAT_FLUX, CONVECTION_HEAT_FLUX, SURFACE_RADIATION_HEAT_FLUX
};
class BoundaryConditionBase : public FunctionSetBase {
 public:
  explicit BoundaryConditionBase(BoundaryConditionType t) : _type(t) {}
  BoundaryConditionType type() const { return _type; }
 private:
  BoundaryConditionType _type;
};
class DirichletBoundaryCondition : public BoundaryConditionBase {  // src/boundary_condition/dirichlet_boundary_condition.h
 public:
  DirichletBoundaryCondition() : BoundaryConditionBase(DIRICHLET) {}
  void init(libMesh::boundary_id_type bid, const std::vector<unsigned int>& constrained_vars) { _bid = bid; _vars = constrained_vars; }
  libMesh::boundary_id_type boundary_id() const { return _bid; }
  const std::vector<unsigned int>& vars() const { return _vars; }
 private:
  libMesh::boundary_id_type _bid = 0;
  std::vector<unsigned int> _vars;
};

// a load on a set of nodes (src/boundary_condition/point_load_condition.h): "load" = ConstantVectorFieldFunction
class PointLoadCondition : public BoundaryConditionBase {
 public:
  explicit PointLoadCondition(BoundaryConditionType t = POINT_LOAD) : BoundaryConditionBase(t) {}
  void add_node(libMesh::dof_id_type n) { _nodes.insert(n); }
  const std::set<libMesh::dof_id_type>& get_nodes() const { return _nodes; }
 private:
  std::set<libMesh::dof_id_type> _nodes;
};

class MaterialPropertyCardBase : public FunctionSetBase {};
class IsotropicMaterialPropertyCard : public MaterialPropertyCardBase {};  // expects "E", "nu" [, "alpha_expansion", "k_th"]
enum StrainType { LINEAR_STRAIN, NONLINEAR_STRAIN };
class ElementPropertyCardBase : public FunctionSetBase {
 public:
  void set_material(MaterialPropertyCardBase& m) { _mat = &m; }
  const MaterialPropertyCardBase& get_material() const;
  void set_strain(StrainType s) { _strain = s; }
  StrainType strain_type() const { return _strain; }
 private:
  MaterialPropertyCardBase* _mat = nullptr;
  StrainType _strain = LINEAR_STRAIN;
};
class IsotropicElementPropertyCard3D : public ElementPropertyCardBase {};
class Solid2DSectionElementPropertyCard : public ElementPropertyCardBase {  // expects "h" [, "off", "kappa"]
 public:
  void set_plane_stress(bool val) { _if_plane_stress = val; }   // src/property_cards/element_property_card_2D.h:73-82
  bool plane_stress() const { return _if_plane_stress; }
 private:
  bool _if_plane_stress = true;
};

// ---- systems ------------------------------------------------------------------------------------------------------
class NonlinearSystem : public libMesh::NonlinearImplicitSystem {  // src/base/nonlinear_system.h
 public:
  using libMesh::NonlinearImplicitSystem::NonlinearImplicitSystem;
};
class SystemInitialization {  // src/base/system_initialization.h
 public:
  SystemInitialization(NonlinearSystem& sys, const std::string& prefix) : _system(sys), _prefix(prefix) {}
  virtual ~SystemInitialization() {}
  NonlinearSystem& system() { return _system; }
  const NonlinearSystem& system() const { return _system; }
  unsigned int n_vars() const { return (unsigned)_vars.size(); }
  const std::vector<unsigned int>& vars() const { return _vars; }
 protected:
  NonlinearSystem& _system;
  std::string _prefix;
  std::vector<unsigned int> _vars;
};
class StructuralSystemInitialization : public SystemInitialization {  // ux,uy,uz,tx,ty,tz (src/elasticity/structural_system_initialization.cpp)
 public:
  StructuralSystemInitialization(NonlinearSystem& sys, const std::string& prefix, const libMesh::FEType& fe);
};
class HeatConductionSystemInitialization : public SystemInitialization {  // one temperature variable
 public:
  HeatConductionSystemInitialization(NonlinearSystem& sys, const std::string& prefix, const libMesh::FEType& fe);
};

class PhysicsDisciplineBase {  // src/base/physics_discipline_base.h
 public:
  explicit PhysicsDisciplineBase(libMesh::EquationSystems& eq) : _eq(eq) {}
  virtual ~PhysicsDisciplineBase() {}
  void add_side_load(libMesh::boundary_id_type bid, BoundaryConditionBase& load);
  void add_dirichlet_bc(libMesh::boundary_id_type bid, DirichletBoundaryCondition& load);
  void add_volume_load(libMesh::subdomain_id_type sid, BoundaryConditionBase& load);
  void add_point_load(PointLoadCondition& load) { _point_loads.insert(&load); }
  const std::set<PointLoadCondition*>& point_loads() const { return _point_loads; }
  void set_property_for_subdomain(libMesh::subdomain_id_type sid, const ElementPropertyCardBase& prop);
  const ElementPropertyCardBase& get_property_card(unsigned int sid) const;
  void init_system_dirichlet_bc(NonlinearSystem& sys) const;
  void clear_system_dirichlet_bc(NonlinearSystem& sys) const;
  const std::multimap<libMesh::boundary_id_type, BoundaryConditionBase*>& side_loads() const { return _side_bc; }
  const std::multimap<libMesh::subdomain_id_type, BoundaryConditionBase*>& volume_loads() const { return _vol_bc; }
  libMesh::EquationSystems& get_equation_systems() { return _eq; }
 private:
  libMesh::EquationSystems& _eq;
  std::multimap<libMesh::boundary_id_type, BoundaryConditionBase*> _side_bc;
  std::multimap<libMesh::subdomain_id_type, BoundaryConditionBase*> _vol_bc;
  std::map<libMesh::boundary_id_type, DirichletBoundaryCondition*> _dirichlet;
  std::set<PointLoadCondition*> _point_loads;
  std::map<libMesh::subdomain_id_type, const ElementPropertyCardBase*> _prop;
};

// ---- elem-ops -----------------------------------------------------------------------------------------------------
class AssemblyBase;
class GeomElem;
class ElementBase;

class AssemblyElemOperations {
 public:
  AssemblyElemOperations();
  virtual ~AssemblyElemOperations();
  SystemInitialization& get_system_initialization();
  PhysicsDisciplineBase& get_discipline();
  virtual void set_discipline_and_system(PhysicsDisciplineBase& discipline, SystemInitialization& system);
  virtual void clear_discipline_and_system();
  virtual void set_assembly(AssemblyBase& assembly);
  virtual AssemblyBase& get_assembly();
  virtual void clear_assembly();
  virtual void init(const GeomElem& elem) = 0;
  virtual void clear_elem();
  void set_skip_comm_sum(bool skip) { _skip_comm_sum = skip; }
  virtual void set_elem_solution(const RealVectorX& sol);
  virtual void set_elem_solution_sensitivity(const RealVectorX& sol);
 protected:
  [[noreturn]] void _host_element_path(const char* what) const;
  SystemInitialization* _system;
  PhysicsDisciplineBase* _discipline;
  AssemblyBase* _assembly;
  ElementBase* _physics_elem;
  bool _skip_comm_sum;
};

class NonlinearImplicitAssemblyElemOperations : public AssemblyElemOperations {
 public:
  virtual void elem_calculations(bool if_jac, RealVectorX& vec, RealMatrixX& mat) = 0;
  virtual void elem_linearized_jacobian_solution_product(RealVectorX& vec) = 0;
  virtual void elem_sensitivity_calculations(const FunctionBase& f, RealVectorX& vec) = 0;
  virtual void elem_shape_sensitivity_calculations(const FunctionBase& f, RealVectorX& vec) = 0;
  virtual void elem_topology_sensitivity_calculations(const FunctionBase& f, RealVectorX& vec) = 0;
  virtual void elem_second_derivative_dot_solution_assembly(RealMatrixX& mat) = 0;
};

// device-backed concrete elem-ops: the per-element virtuals throw (the loop lives on the GPU)
class DeviceElemOperations : public NonlinearImplicitAssemblyElemOperations {
 public:
  void init(const GeomElem& elem) override;
  void elem_calculations(bool if_jac, RealVectorX& vec, RealMatrixX& mat) override;
  void elem_linearized_jacobian_solution_product(RealVectorX& vec) override;
  void elem_sensitivity_calculations(const FunctionBase& f, RealVectorX& vec) override;
  void elem_shape_sensitivity_calculations(const FunctionBase& f, RealVectorX& vec) override;
  void elem_topology_sensitivity_calculations(const FunctionBase& f, RealVectorX& vec) override;
  void elem_second_derivative_dot_solution_assembly(RealMatrixX& mat) override;
};
class StructuralNonlinearAssemblyElemOperations : public DeviceElemOperations {};
class HeatConductionNonlinearAssemblyElemOperations : public DeviceElemOperations {};

class OutputAssemblyElemOperations : public AssemblyElemOperations {
 public:
  virtual void zero_for_analysis() = 0;
  virtual void zero_for_sensitivity() = 0;
  virtual Real output_total() = 0;
  virtual Real output_sensitivity_total(const FunctionBase& p) = 0;
  virtual void evaluate() = 0;
  virtual void evaluate_sensitivity(const FunctionBase& f) = 0;
  virtual void output_derivative_for_elem(RealVectorX& dq_dX) = 0;
  void set_participating_elements_to_all() { _all = true; }
 protected:
  bool _all = true;
};
class ComplianceOutput : public OutputAssemblyElemOperations {
 public:
  void init(const GeomElem& elem) override;
  void zero_for_analysis() override { _compliance = 0.; _dcompliance_dp = 0.; }
  void zero_for_sensitivity() override { _dcompliance_dp = 0.; }
  Real output_total() override { return _compliance; }
  Real output_sensitivity_total(const FunctionBase&) override { return _dcompliance_dp; }
  void evaluate() override;
  void evaluate_sensitivity(const FunctionBase& f) override;
  void output_derivative_for_elem(RealVectorX& dq_dX) override;
 private:
  friend class AssemblyBase;
  Real _compliance = 0., _dcompliance_dp = 0.;
};

// von Mises p-norm / KS aggregate over the quadrature points (src/elasticity/stress_output_base.h:246-252, .cpp:851-909,
// 1104-1172).  Only p_norm_weight == 1 (the reference's default, stress_output_base.cpp:278-281) is on the device path.
class StressStrainOutputBase : public OutputAssemblyElemOperations {
 public:
  void init(const GeomElem& elem) override;
  void set_aggregation_coefficients(Real p1, Real p2, Real rho, Real sigma0) {
    _p_norm_stress = p1; _p_norm_weight = p2; _rho = rho; _sigma0 = sigma0;
  }
  Real get_p_stress_val() const { return _p_norm_stress; }
  void zero_for_analysis() override { _sigma_vm_int = 0.; _JxW_val = 0.; _primal_data_initialized = false; }
  void zero_for_sensitivity() override {}
  Real output_total() override;   // functional_for_all_elems
  Real output_sensitivity_total(const FunctionBase&) override;  // set by calculate_output_direct_sensitivity
  void evaluate() override;
  void evaluate_sensitivity(const FunctionBase& f) override;
  void output_derivative_for_elem(RealVectorX& dq_dX) override;
 private:
  friend class AssemblyBase;
  Real _p_norm_stress = 2., _p_norm_weight = 1., _rho = 1., _sigma0 = 1.;
  Real _sigma_vm_int = 0., _JxW_val = 0.;   // the two sums of functional_for_all_elems (this rank's part)
  Real _dq_dp = 0.;                         // last direct sensitivity
  bool _primal_data_initialized = false;
};

// ---- assemblies ---------------------------------------------------------------------------------------------------
class AssemblyBase : public libMesh::NonlinearImplicitSystem::ComputeResidualandJacobian {
 public:
  AssemblyBase();
  virtual ~AssemblyBase();
  std::set<unsigned int> diagonal_elem_subdomain_id;
  bool close_matrix;

  class ElemParameterDependence {
   public:
    explicit ElemParameterDependence(bool o_flag) : override_flag(o_flag) {}
    virtual ~ElemParameterDependence() {}
    bool override_flag;
  };

  const PhysicsDisciplineBase& discipline() const;
  PhysicsDisciplineBase& discipline();
  AssemblyElemOperations& get_elem_ops();
  virtual void set_discipline_and_system(PhysicsDisciplineBase& discipline, SystemInitialization& system);
  virtual void clear_discipline_and_system();
  void attach_elem_parameter_dependence_object(ElemParameterDependence& dep);
  void clear_elem_parameter_dependence_object();
  virtual void set_elem_operation_object(AssemblyElemOperations& elem_ops);
  virtual void clear_elem_operation_object();
  const NonlinearSystem& system() const;
  NonlinearSystem& system();
  SystemInitialization& system_init();

  void residual_and_jacobian(const libMesh::NumericVector<Real>& X, libMesh::NumericVector<Real>* R,
                             libMesh::SparseMatrix<Real>* J, libMesh::NonlinearImplicitSystem& S) override {
    (void)X; (void)R; (void)J; (void)S;
    libmesh_assert(false);  // implement in the derived class
  }
  virtual bool sensitivity_assemble(const libMesh::NumericVector<Real>& X, bool if_localize_sol, const FunctionBase& f,
                                    libMesh::NumericVector<Real>& sensitivity_rhs, bool close_vector = true) {
    (void)X; (void)if_localize_sol; (void)f; (void)sensitivity_rhs; (void)close_vector;
    libmesh_assert(false);  // implemented in the derived class
    return false;
  }
  virtual void calculate_output(const libMesh::NumericVector<Real>& X, bool if_localize_sol,
                                OutputAssemblyElemOperations& output);
  virtual void calculate_output_derivative(const libMesh::NumericVector<Real>& X, bool if_localize_sol,
                                           OutputAssemblyElemOperations& output, libMesh::NumericVector<Real>& dq_dX);
  // assembly_base.h:282-288: partial (fixed-state) sensitivity of the output plus dq/dX . dXdp when dXdp is given
  virtual void calculate_output_direct_sensitivity(const libMesh::NumericVector<Real>& X, bool if_localize_sol,
                                                   const libMesh::NumericVector<Real>* dXdp, bool if_localize_sol_sens,
                                                   const FunctionBase& p, OutputAssemblyElemOperations& output);
  virtual Real calculate_output_adjoint_sensitivity(const libMesh::NumericVector<Real>& X, bool if_localize_sol,
                                                    const libMesh::NumericVector<Real>& adj_sol, const FunctionBase& p,
                                                    AssemblyElemOperations& elem_ops, OutputAssemblyElemOperations& output,
                                                    const bool include_partial_sens = true);
  virtual void calculate_output_adjoint_sensitivity_multiple_parameters_no_direct(
      const libMesh::NumericVector<Real>& X, bool if_localize_sol, const libMesh::NumericVector<Real>& adj_sol,
      const std::vector<const FunctionBase*>& p_vec, AssemblyElemOperations& elem_ops,
      OutputAssemblyElemOperations& output, std::vector<Real>& sens);
  std::unique_ptr<libMesh::NumericVector<Real>> build_localized_vector(const libMesh::System& sys,
                                                                      const libMesh::NumericVector<Real>& global) const;

  // device the sweeps run on (default 0); must be set before the first sweep
  void set_device(int device) { _device_id = device; }
  // statistics of the last sweep (kernel launches, CUDA-event times): see mastb200_get_stats
  long long device_kernel_launches() const;

 protected:
  // builds (once per mesh) and refreshes (every sweep) the device problem from the discipline; throws on failure
  mastb200_context* _device_problem(const libMesh::SparseMatrix<Real>* J);
  [[noreturn]] void _device_error(const char* what) const;
  void _check(int rc, const char* what) const;

  AssemblyElemOperations* _elem_ops;
  PhysicsDisciplineBase* _discipline;
  SystemInitialization* _system;
  ElemParameterDependence* _param_dependence;
  mastb200_context* _ctx;
  int _device_id;
  bool _pattern_checked;
  const NodalFieldFunction* _rho_field;
  std::vector<double> _side_values;   // side / point load values of the last sweep: loads are re-sent when they change
  // which Parameter of the property card a FunctionBase is: MASTB200_PARAM_E / _NU, or -1
  int _card_parameter(const FunctionBase& f) const;
};

class NonlinearImplicitAssembly : public AssemblyBase {
 public:
  NonlinearImplicitAssembly();
  ~NonlinearImplicitAssembly() override;
  class PostAssemblyOperation {
   public:
    virtual ~PostAssemblyOperation() {}
    virtual void post_assembly(const libMesh::NumericVector<Real>& X, libMesh::NumericVector<Real>* R,
                               libMesh::SparseMatrix<Real>* J, libMesh::NonlinearImplicitSystem& S) = 0;
  };
  void set_post_assembly_operation(PostAssemblyOperation& post) { _post_assembly = &post; }
  Real res_l2_norm() const { return _res_l2_norm; }
  void residual_and_jacobian(const libMesh::NumericVector<Real>& X, libMesh::NumericVector<Real>* R,
                             libMesh::SparseMatrix<Real>* J, libMesh::NonlinearImplicitSystem& S) override;
  bool sensitivity_assemble(const libMesh::NumericVector<Real>& X, bool if_localize_sol, const FunctionBase& f,
                            libMesh::NumericVector<Real>& sensitivity_rhs, bool close_vector = true) override;
  // nonlinear_implicit_assembly.h:115: JdX = J(X) dX without the matrix (matrix-free on the device)
  virtual void linearized_jacobian_solution_product(const libMesh::NumericVector<Real>& X, const libMesh::NumericVector<Real>& dX,
                                                    libMesh::NumericVector<Real>& JdX, libMesh::NonlinearImplicitSystem& S);
  // nonlinear_implicit_assembly.h:125: d(J dX)/dX for the second-order adjoint; identically zero for linear strain
  virtual void second_derivative_dot_solution_assembly(const libMesh::NumericVector<Real>& X, const libMesh::NumericVector<Real>& dX,
                                                       libMesh::SparseMatrix<Real>& d_JdX_dX, libMesh::NonlinearImplicitSystem& S);
 protected:
  PostAssemblyOperation* _post_assembly;
  Real _res_l2_norm;
};

}  // namespace MAST
