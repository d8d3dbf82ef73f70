// libmesh_shim.h -- the handful of libMesh types that MAST's assembly signatures mention, so that the host mirror
// (mast_b200_assembly.h) compiles in an image without libMesh/PETSc.  In a real MAST build define
// MAST_B200_USE_LIBMESH and the genuine headers are used instead; the mirror only relies on the members declared here.
//
// This is NOT a libMesh re-implementation: no FE, no solvers, no parallel vectors -- a serial dense NumericVector,
// a CSR SparseMatrix with a libMesh-style preallocated pattern, a replicated structured mesh with build_cube /
// build_square numbering and boundary ids (MeshTools::Generation), and an EquationSystems that numbers dofs node-major
// for one variable group (what libMesh's DofMap does for MAST's single-FEType systems).
#pragma once
#ifdef MAST_B200_USE_LIBMESH
#include "libmesh/boundary_info.h"
#include "libmesh/dof_map.h"
#include "libmesh/equation_systems.h"
#include "libmesh/mesh_generation.h"
#include "libmesh/nonlinear_implicit_system.h"
#include "libmesh/numeric_vector.h"
#include "libmesh/replicated_mesh.h"
#include "libmesh/sparse_matrix.h"
#else

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <map>
#include <memory>
#include <set>
#include <stdexcept>
#include <string>
#include <vector>

namespace libMesh {

typedef double Real;
typedef Real Number;
typedef uint32_t dof_id_type;
typedef unsigned short subdomain_id_type;
typedef short boundary_id_type;

class LogicError : public std::logic_error {
 public:
  explicit LogicError(const std::string& m) : std::logic_error(m) {}
};
#define libmesh_error_msg(msg) throw libMesh::LogicError(std::string(__FILE__) + ":" + std::to_string(__LINE__) + ": " + (msg))
#define libmesh_error() libmesh_error_msg("libmesh_error()")
#ifdef NDEBUG
#define libmesh_assert(x) ((void)0)
#else
#define libmesh_assert(x) do { if (!(x)) libmesh_error_msg("assertion failed: " #x); } while (0)
#endif
#define libmesh_assert_msg(x, m) libmesh_assert(x)

enum ElemType { QUAD4 = 5, HEX8 = 10, HEX27 = 12 };
enum Order { FIRST = 1, SECOND = 2 };
enum FEFamily { LAGRANGE = 0 };
struct FEType {
  Order order; FEFamily family;
  FEType(Order o = FIRST, FEFamily f = LAGRANGE) : order(o), family(f) {}
};

template <typename T>
class NumericVector {
 public:
  explicit NumericVector(dof_id_type n = 0) : _v(n, T(0)), _closed(true) {}
  void init(dof_id_type n) { _v.assign(n, T(0)); _closed = true; }
  dof_id_type size() const { return (dof_id_type)_v.size(); }
  dof_id_type local_size() const { return size(); }
  dof_id_type first_local_index() const { return 0; }
  dof_id_type last_local_index() const { return size(); }
  void zero() { std::fill(_v.begin(), _v.end(), T(0)); }
  void close() { _closed = true; }
  bool closed() const { return _closed; }
  T operator()(dof_id_type i) const { return _v[i]; }
  void set(dof_id_type i, T v) { _v[i] = v; _closed = false; }
  void add(dof_id_type i, T v) { _v[i] += v; _closed = false; }
  Real l2_norm() const { Real s = 0; for (T x : _v) s += x * x; return std::sqrt(s); }
  T dot(const NumericVector<T>& o) const { T s = 0; for (size_t i = 0; i < _v.size(); ++i) s += _v[i] * o._v[i]; return s; }
  // contiguous local storage (PetscVector: VecGetArray)
  T* get_array() { return _v.data(); }
  const T* get_array_read() const { return _v.data(); }
 private:
  std::vector<T> _v;
  bool _closed;
};

// CSR matrix with the pattern libMesh's DofMap computes and PETSc AIJ preallocates (sorted columns, zeros kept)
template <typename T>
class SparseMatrix {
 public:
  void init(const std::vector<int64_t>& row_ptr, const std::vector<int32_t>& col_idx) {
    _row_ptr = row_ptr; _col = col_idx; _val.assign(col_idx.size(), T(0)); _closed = true;
  }
  dof_id_type m() const { return (dof_id_type)(_row_ptr.empty() ? 0 : _row_ptr.size() - 1); }
  dof_id_type n() const { return m(); }
  void zero() { std::fill(_val.begin(), _val.end(), T(0)); }
  void close() { _closed = true; }
  bool closed() const { return _closed; }
  T operator()(dof_id_type i, dof_id_type j) const {
    auto b = _col.begin() + _row_ptr[i], e = _col.begin() + _row_ptr[i + 1];
    auto it = std::lower_bound(b, e, (int32_t)j);
    return (it != e && *it == (int32_t)j) ? _val[it - _col.begin()] : T(0);
  }
  // raw CSR access (PetscMatrix: MatSeqAIJGetArray / MatGetRowIJ)
  const std::vector<int64_t>& row_ptr() const { return _row_ptr; }
  const std::vector<int32_t>& col_idx() const { return _col; }
  T* values() { _closed = false; return _val.data(); }
  const T* values() const { return _val.data(); }
  size_t n_nonzeros() const { return _val.size(); }
 private:
  std::vector<int64_t> _row_ptr;
  std::vector<int32_t> _col;
  std::vector<T> _val;
  bool _closed = true;
};

// replicated structured mesh: nodes, connectivity in libMesh node order, subdomain ids, boundary (element, side, id)
class BoundaryInfo {
 public:
  struct Side { dof_id_type elem; unsigned short side; boundary_id_type id; };
  void add_side(dof_id_type e, unsigned short s, boundary_id_type id) { _sides.push_back({e, s, id}); }
  const std::vector<Side>& sides() const { return _sides; }
 private:
  std::vector<Side> _sides;
};

class MeshBase {
 public:
  virtual ~MeshBase() {}
  ElemType elem_type = HEX8;
  std::vector<Real> xyz;            // [n_nodes][3]
  std::vector<int32_t> conn;        // [n_elem][nn]
  std::vector<subdomain_id_type> subdomain;
  BoundaryInfo boundary_info;
  unsigned int nodes_per_elem() const { return elem_type == HEX8 ? 8 : elem_type == HEX27 ? 27 : 4; }
  dof_id_type n_nodes() const { return (dof_id_type)(xyz.size() / 3); }
  dof_id_type n_elem() const { return (dof_id_type)(conn.size() / nodes_per_elem()); }
  unsigned int mesh_dimension() const { return elem_type == QUAD4 ? 2 : 3; }
  BoundaryInfo& get_boundary_info() { return boundary_info; }
  const BoundaryInfo& get_boundary_info() const { return boundary_info; }
};
class ReplicatedMesh : public MeshBase {};

namespace MeshTools {
namespace Generation {
// node id i + j*npx + k*npx*npy, element id i + j*nx + k*nx*ny, boundary ids 0:z- 1:y- 2:x+ 3:y+ 4:x- 5:z+
// (side number == boundary id for hexes), as libMesh's build_cube.
void build_cube(MeshBase& mesh, unsigned nx, unsigned ny, unsigned nz, Real xmin, Real xmax, Real ymin, Real ymax,
                Real zmin, Real zmax, ElemType type);
// QUAD4 in the xy-plane: boundary ids 0:y- 1:x+ 2:y+ 3:x-
void build_square(MeshBase& mesh, unsigned nx, unsigned ny, Real xmin, Real xmax, Real ymin, Real ymax, ElemType type);
}  // namespace Generation
}  // namespace MeshTools

class EquationSystems;

class System {
 public:
  System(EquationSystems& es, const std::string& name) : _es(es), _name(name) {}
  virtual ~System() {}
  const std::string& name() const { return _name; }
  EquationSystems& get_equation_systems() { return _es; }
  const EquationSystems& get_equation_systems() const { return _es; }
  MeshBase& get_mesh();
  const MeshBase& get_mesh() const;
  unsigned int add_variable(const std::string& var, const FEType& fe) { _vars.push_back(var); _fe = fe; return (unsigned)_vars.size() - 1; }
  unsigned int n_vars() const { return (unsigned)_vars.size(); }
  dof_id_type n_dofs() const { return get_mesh().n_nodes() * n_vars(); }
  // homogeneous Dirichlet dofs (DofMap constraints built from DirichletBoundary objects)
  std::vector<uint8_t> constrained_dofs;
  std::unique_ptr<NumericVector<Number>> solution;
  virtual void init_data();
 protected:
  EquationSystems& _es;
  std::string _name;
  std::vector<std::string> _vars;
  FEType _fe;
};

class ImplicitSystem : public System {
 public:
  using System::System;
  NumericVector<Number>* rhs = nullptr;
  SparseMatrix<Number>* matrix = nullptr;
  void init_data() override;
 private:
  std::unique_ptr<NumericVector<Number>> _rhs;
  std::unique_ptr<SparseMatrix<Number>> _matrix;
};

class NonlinearImplicitSystem : public ImplicitSystem {
 public:
  using ImplicitSystem::ImplicitSystem;
  // the callback libMesh's nonlinear solver invokes (nonlinear_implicit_system.h)
  class ComputeResidualandJacobian {
   public:
    virtual ~ComputeResidualandJacobian() {}
    virtual void residual_and_jacobian(const NumericVector<Number>& X, NumericVector<Number>* R, SparseMatrix<Number>* J,
                                       NonlinearImplicitSystem& S) = 0;
  };
};

class EquationSystems {
 public:
  explicit EquationSystems(MeshBase& m) : _mesh(m) {}
  template <typename S>
  S& add_system(const std::string& name) {
    S* s = new S(*this, name);
    _systems[name].reset(s);
    return *s;
  }
  MeshBase& get_mesh() { return _mesh; }
  const MeshBase& get_mesh() const { return _mesh; }
  void init() { for (auto& kv : _systems) kv.second->init_data(); }
 private:
  MeshBase& _mesh;
  std::map<std::string, std::unique_ptr<System>> _systems;
};

inline MeshBase& System::get_mesh() { return _es.get_mesh(); }
inline const MeshBase& System::get_mesh() const { return _es.get_mesh(); }

}  // namespace libMesh
#endif  // MAST_B200_USE_LIBMESH
