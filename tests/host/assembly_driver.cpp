// assembly_driver.cpp -- drives the host mirror exactly as a MAST example drives the reference
// (/root/reference/examples/structural/example_1/example_1.cpp:52-170 for the call sequence) and dumps the assembled
// vectors/matrices for the parity tests (tests/test_host_mirror.py compares them with the CPU oracle).
//
//   assembly_driver <case> <dir>     case: c1 | simp | heat | hex27
// Reads <dir>/X.bin (and rho.bin / adj.bin / T.bin where the case needs them), writes R.bin, J.bin, row_ptr.bin,
// col_idx.bin, scalars.txt.
#include <cstdio>
#include <fstream>
#include <iostream>

#include "mast_b200_assembly.h"

static std::vector<double> read_bin(const std::string& path) {
  std::ifstream f(path, std::ios::binary | std::ios::ate);
  if (!f) throw std::runtime_error("cannot open " + path);
  const std::streamsize n = f.tellg();
  f.seekg(0);
  std::vector<double> v(n / 8);
  f.read(reinterpret_cast<char*>(v.data()), n);
  return v;
}
template <typename T>
static void write_bin(const std::string& path, const T* p, size_t n) {
  std::ofstream f(path, std::ios::binary);
  f.write(reinterpret_cast<const char*>(p), std::streamsize(n * sizeof(T)));
}

int main(int argc, char** argv) {
  if (argc < 3) { std::fprintf(stderr, "usage: assembly_driver <c1|simp|heat|hex27|plate> <dir>\n"); return 2; }
  const std::string which = argv[1], dir = argv[2];
  try {
    libMesh::ReplicatedMesh mesh;
    const bool heat = which == "heat", plate = which == "plate";
    if (heat || plate) libMesh::MeshTools::Generation::build_square(mesh, 24, 16, 0., 1., 0., 1., libMesh::QUAD4);
    else if (which == "hex27") libMesh::MeshTools::Generation::build_cube(mesh, 4, 3, 3, 0., 0.3, 0., 0.15, 0., 0.15, libMesh::HEX27);
    else libMesh::MeshTools::Generation::build_cube(mesh, 20, 10, 10, 0., 0.3, 0., 0.15, 0., 0.15, libMesh::HEX8);

    libMesh::EquationSystems equation_systems(mesh);
    MAST::NonlinearSystem& system = equation_systems.add_system<MAST::NonlinearSystem>(heat ? "thermal" : "structural");
    libMesh::FEType fetype(which == "hex27" ? libMesh::SECOND : libMesh::FIRST, libMesh::LAGRANGE);
    std::unique_ptr<MAST::SystemInitialization> sys_init;
    if (heat) sys_init.reset(new MAST::HeatConductionSystemInitialization(system, system.name(), fetype));
    else sys_init.reset(new MAST::StructuralSystemInitialization(system, system.name(), fetype));
    MAST::PhysicsDisciplineBase discipline(equation_systems);

    // clamp the face x = 0 (boundary id 4 of build_cube, 3 of build_square)
    MAST::DirichletBoundaryCondition dirichlet_bc;
    dirichlet_bc.init((heat || plate) ? 3 : 4, sys_init->vars());
    discipline.add_dirichlet_bc((heat || plate) ? 3 : 4, dirichlet_bc);
    equation_systems.init();
    discipline.init_system_dirichlet_bc(system);

    MAST::Parameter E("E", 72.0e9), nu("nu", 0.33), alpha("alpha", 1.5e-5), press("p", 5.0e7), kth("k", 237.), h("h", 0.06);
    MAST::Parameter T0("T0", 300.), Tc("T", 350.);
    MAST::ConstantFieldFunction E_f("E", E), nu_f("nu", nu), alpha_f("alpha_expansion", alpha), press_f("pressure", press);
    MAST::ConstantFieldFunction h_f("h", h), T0_f("ref_temperature", T0), T_f("temperature", Tc);
    MAST::Parameter off("off", 0.03), kappa("kappa", 5. / 6.);
    MAST::ConstantFieldFunction off_f("off", off), kappa_f("kappa", kappa);
    MAST::LinearTemperatureFieldFunction k_f("k_th", 237., 1.0e-3);
    std::unique_ptr<MAST::NodalFieldFunction> rho_f;
    std::unique_ptr<MAST::SIMPModulusFieldFunction> simp_f;

    MAST::IsotropicMaterialPropertyCard material;
    MAST::IsotropicElementPropertyCard3D section3d;
    MAST::Solid2DSectionElementPropertyCard section2d;
    MAST::BoundaryConditionBase pressure_load(MAST::SURFACE_PRESSURE), temperature_load(MAST::TEMPERATURE);
    if (heat) {
      material.add(k_f);
      section2d.add(h_f);
      section2d.set_material(material);
      discipline.set_property_for_subdomain(0, section2d);
    } else if (plate) {  // Mindlin plate: solid 2D section (h, off, kappa) of an isotropic material, plane stress
      material.add(E_f);
      material.add(nu_f);
      section2d.add(h_f); section2d.add(off_f); section2d.add(kappa_f);
      section2d.set_material(material);
      discipline.set_property_for_subdomain(0, section2d);
    } else {
      if (which == "simp") {
        rho_f.reset(new MAST::NodalFieldFunction("rho", read_bin(dir + "/rho.bin")));
        simp_f.reset(new MAST::SIMPModulusFieldFunction("E", 72.0e9, 1.0e-8, 4.0, *rho_f));
        material.add(*simp_f);
      } else {
        material.add(E_f);
      }
      material.add(nu_f);
      material.add(alpha_f);
      section3d.set_material(material);
      discipline.set_property_for_subdomain(0, section3d);
      if (which == "hex27") {  // thermoelastic volume load (C5)
        temperature_load.add(T_f);
        temperature_load.add(T0_f);
        discipline.add_volume_load(0, temperature_load);
      } else {                 // pressure on the face y = Ly (boundary id 3)
        pressure_load.add(press_f);
        discipline.add_side_load(3, pressure_load);
      }
    }

    // a point load the c1 case adds for its second sweep (the discipline keeps the pointer: lives as long as it)
    MAST::PointLoadCondition point_load;
    MAST::ConstantVectorFieldFunction load_f("load", {1.0e5, -2.0e5, 3.0e5, 0., 0., 0.});

    MAST::NonlinearImplicitAssembly assembly;
    MAST::StructuralNonlinearAssemblyElemOperations structural_ops;
    MAST::HeatConductionNonlinearAssemblyElemOperations thermal_ops;
    MAST::NonlinearImplicitAssemblyElemOperations& elem_ops =
        heat ? static_cast<MAST::NonlinearImplicitAssemblyElemOperations&>(thermal_ops) : structural_ops;
    assembly.set_discipline_and_system(discipline, *sys_init);
    elem_ops.set_discipline_and_system(discipline, *sys_init);
    assembly.set_elem_operation_object(elem_ops);

    const std::vector<double> X = read_bin(dir + "/X.bin");
    if (X.size() != system.n_dofs()) throw std::runtime_error("X.bin has the wrong size");
    for (libMesh::dof_id_type i = 0; i < system.n_dofs(); ++i) system.solution->set(i, X[i]);
    system.solution->close();

    assembly.residual_and_jacobian(*system.solution, system.rhs, system.matrix, system);

    write_bin(dir + "/R.bin", system.rhs->get_array_read(), system.n_dofs());
    write_bin(dir + "/J.bin", system.matrix->values(), system.matrix->n_nonzeros());
    write_bin(dir + "/row_ptr.bin", system.matrix->row_ptr().data(), system.matrix->row_ptr().size());
    write_bin(dir + "/col_idx.bin", system.matrix->col_idx().data(), system.matrix->col_idx().size());
    std::ofstream sc(dir + "/scalars.txt");
    sc.precision(17);
    sc << "res_l2_norm " << assembly.res_l2_norm() << "\n";

    if (which == "c1" || which == "simp") {  // ComplianceOutput value and derivative
      MAST::ComplianceOutput compliance;
      compliance.set_discipline_and_system(discipline, *sys_init);
      assembly.calculate_output(*system.solution, false, compliance);
      sc << "compliance " << compliance.output_total() << "\n";
      libMesh::NumericVector<Real> dq_dX(system.n_dofs());
      assembly.calculate_output_derivative(*system.solution, false, compliance, dq_dX);
      write_bin(dir + "/dq_dX.bin", dq_dX.get_array_read(), dq_dX.size());
      if (which == "simp") {
        // three density DVs, each a small patch of nodes with filter-like weights; adjoint read from file
        const std::vector<double> adjv = read_bin(dir + "/adj.bin");
        libMesh::NumericVector<Real> adj(system.n_dofs());
        for (libMesh::dof_id_type i = 0; i < system.n_dofs(); ++i) adj.set(i, adjv[i]);
        adj.close();
        MAST::DensityDesignVariable dv0("dv0", {100, 101, 121}, {0.5, 0.25, 0.25});
        MAST::DensityDesignVariable dv1("dv1", {1200, 1201}, {0.7, 0.3});
        MAST::DensityDesignVariable dv2("dv2", {2540}, {1.0});
        std::vector<const MAST::FunctionBase*> p_vec = {&dv0, &dv1, &dv2};
        std::vector<Real> sens;
        assembly.calculate_output_adjoint_sensitivity_multiple_parameters_no_direct(*system.solution, false, adj, p_vec,
                                                                                   structural_ops, compliance, sens);
        for (size_t i = 0; i < sens.size(); ++i) sc << "sens" << i << " " << sens[i] << "\n";
        sc << "sens1_single " << assembly.calculate_output_adjoint_sensitivity(*system.solution, false, adj, dv1, structural_ops, compliance) << "\n";
      }
      compliance.clear_discipline_and_system();
    }
    if (which == "c1") {
      // sensitivity_assemble for the card's Parameters, the matrix-free J dX, point loads, load stepping of the pressure
      libMesh::NumericVector<Real> dres(system.n_dofs());
      const bool okE = assembly.sensitivity_assemble(*system.solution, false, E, dres);
      write_bin(dir + "/dRdE.bin", dres.get_array_read(), dres.size());
      const bool okN = assembly.sensitivity_assemble(*system.solution, false, nu, dres);
      write_bin(dir + "/dRdnu.bin", dres.get_array_read(), dres.size());
      const bool okH = assembly.sensitivity_assemble(*system.solution, false, h, dres);  // not a parameter of this problem
      sc << "sens_E_ok " << okE << "\nsens_nu_ok " << okN << "\nsens_other_ok " << okH << "\n";
      const std::vector<double> dXv = read_bin(dir + "/dX.bin");
      libMesh::NumericVector<Real> dX(system.n_dofs()), JdX(system.n_dofs());
      for (libMesh::dof_id_type i = 0; i < system.n_dofs(); ++i) dX.set(i, dXv[i]);
      dX.close();
      assembly.linearized_jacobian_solution_product(*system.solution, dX, JdX, system);
      write_bin(dir + "/JdX.bin", JdX.get_array_read(), JdX.size());
      {
        MAST::ComplianceOutput compliance;
        compliance.set_discipline_and_system(discipline, *sys_init);
        assembly.calculate_output_direct_sensitivity(*system.solution, false, &dX, false, E, compliance);
        sc << "compliance_direct " << compliance.output_sensitivity_total(E) << "\n";
        compliance.clear_discipline_and_system();
      }
      point_load.add(load_f);
      point_load.add_node(2540); point_load.add_node(1300); point_load.add_node(0);
      discipline.add_point_load(point_load);
      press = 2.5e7;  // load stepping: the assembly must pick the new value up
      assembly.residual_and_jacobian(*system.solution, system.rhs, nullptr, system);
      write_bin(dir + "/R_step.bin", system.rhs->get_array_read(), system.n_dofs());
    }
    if (which == "simp") {  // the global dR/dp vector of one filtered density design variable
      libMesh::NumericVector<Real> dres(system.n_dofs());
      MAST::DensityDesignVariable dv1("dv1", {1200, 1201}, {0.7, 0.3});
      const bool ok = assembly.sensitivity_assemble(*system.solution, false, dv1, dres);
      sc << "sens_dv_ok " << ok << "\n";
      write_bin(dir + "/dRdp1.bin", dres.get_array_read(), dres.size());
      // stress aggregate: adjoint sensitivity with the partial term (include_partial_sens = true)
      const std::vector<double> adjv = read_bin(dir + "/adj.bin");
      libMesh::NumericVector<Real> adj(system.n_dofs());
      for (libMesh::dof_id_type i = 0; i < system.n_dofs(); ++i) adj.set(i, adjv[i]);
      adj.close();
      MAST::StressStrainOutputBase stress;
      stress.set_discipline_and_system(discipline, *sys_init);
      stress.set_aggregation_coefficients(6., 1., 1.5, 2.0e8);
      assembly.calculate_output(*system.solution, false, stress);
      sc << "stress_total_sens " << assembly.calculate_output_adjoint_sensitivity(*system.solution, false, adj, dv1, structural_ops, stress, true) << "\n";
      sc << "stress_adjoint_only " << assembly.calculate_output_adjoint_sensitivity(*system.solution, false, adj, dv1, structural_ops, stress, false) << "\n";
      stress.clear_discipline_and_system();
    }
    if (which == "c1") {  // StressStrainOutputBase: von Mises p-norm aggregate and its state derivative (adjoint rhs)
      MAST::StressStrainOutputBase stress;
      stress.set_discipline_and_system(discipline, *sys_init);
      stress.set_aggregation_coefficients(6., 1., 1.5, 2.0e8);
      assembly.calculate_output(*system.solution, false, stress);
      sc << "stress_functional " << stress.output_total() << "\n";
      libMesh::NumericVector<Real> ds_dX(system.n_dofs());
      assembly.calculate_output_derivative(*system.solution, false, stress, ds_dX);
      write_bin(dir + "/ds_dX.bin", ds_dX.get_array_read(), ds_dX.size());
      stress.clear_discipline_and_system();
    }
    sc << "kernel_launches " << assembly.device_kernel_launches() << "\n";

    // error behaviour: the per-element host virtuals are not a fallback
    bool threw = false;
    try {
      RealVectorX v; RealMatrixX m;
      structural_ops.elem_calculations(true, v, m);
    } catch (const libMesh::LogicError&) { threw = true; }
    sc << "elem_calculations_throws " << (threw ? 1 : 0) << "\n";

    assembly.clear_elem_operation_object();
    assembly.clear_discipline_and_system();
    elem_ops.clear_discipline_and_system();
  } catch (const libMesh::LogicError& e) {
    std::cerr << "libMesh::LogicError: " << e.what() << std::endl;
    return 3;
  } catch (const std::exception& e) {
    std::cerr << "error: " << e.what() << std::endl;
    return 4;
  }
  return 0;
}
