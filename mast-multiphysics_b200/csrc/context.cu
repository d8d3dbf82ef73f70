// context.cu -- implementation of the C-ABI in include/mast_b200.h.
//
// Owns the device copies of mesh, dof map, plan and state; resolves host-or-device pointers; orders all
// work on one CUDA stream.  There is no CPU compute path: without a usable CUDA device every sweep
// entry point fails with MASTB200_ERR_CUDA.
#include "context.h"

std::string g_create_error;

// device view of an input array: device pointers pass through, host pointers are copied into `stage`
int resolve_in(mastb200_context* ctx, const double* ptr, size_t n, DevBuf<double>& stage, const double** out) {
  if (!ptr) { *out = nullptr; return MASTB200_OK; }
  if (is_device_ptr(ptr)) { *out = ptr; return MASTB200_OK; }
  CU(stage.ensure(n));
  CU(cudaMemcpyAsync(stage.p, ptr, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  *out = stage.p;
  return MASTB200_OK;
}
int resolve_out(mastb200_context* ctx, double* ptr, size_t n, DevBuf<double>& stage, double** out) {
  if (!ptr) { *out = nullptr; return MASTB200_OK; }
  if (is_device_ptr(ptr)) { *out = ptr; return MASTB200_OK; }
  CU(stage.ensure(n));
  *out = stage.p;
  return MASTB200_OK;
}
int finish_out(mastb200_context* ctx, double* ptr, size_t n, const double* dev) {
  if (!ptr || ptr == dev) return MASTB200_OK;
  CU(cudaMemcpyAsync(ptr, dev, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  return MASTB200_OK;
}

int64_t plan_nnz(const mastb200_context* ctx) {
  return ctx->pat.nbr_ptr[ctx->pat.n_owned] * int64_t(ctx->nv) * ctx->nv;
}

// builds (once) the streaming plan for the plain or the aux shared-memory layout and uploads it
int ensure_plan(mastb200_context* ctx, int aux) {
  mastb200_context::DevicePlan& D = ctx->dplan[aux];
  if (D.built) return MASTB200_OK;
  if (ctx->elem_type != MASTB200_HEX8) return MASTB200_OK;  // pattern only: no device sweep for the type yet
  D.smem = hex8_sweep_smem_bytes(aux != 0);
  if ((size_t)ctx->max_optin < D.smem) return ctx->fail(MASTB200_ERR_CUDA, "device shared memory too small");
  D.lim.arena_blocks = hex8_sweep_arena_blocks(aux != 0);
  D.lim.max_slots = kMaxSlots;
  D.lim.flush_buf_bytes = kFlushBufBytes;
  D.lim.sm_count = ctx->sm_count * kCtasPerSm;
  D.lim.n_workers = hex8_sweep_workers();
  StreamPlan P;
  const std::string e = build_stream_plan(ctx->pat, ctx->xyz.data(), ctx->conn.data(),
                                          ctx->diag_elem.empty() ? nullptr : ctx->diag_elem.data(), D.lim, P);
  if (!e.empty()) return ctx->fail(MASTB200_ERR_PLAN, "plan: %s", e.c_str());
  D.n_tiles = (int64_t)P.tiles.size(); D.n_visits = P.n_visits(); D.bytes = P.bytes();
  D.max_live_blocks = P.max_live_blocks; D.max_live_slots = P.max_live_slots;
  D.n_ctas = (int)std::max<int64_t>(1, std::min<int64_t>(D.n_tiles, (int64_t)ctx->sm_count * kCtasPerSm));
  if (ctx->planning_only) {
    D.host = std::move(P);
  } else {
    CU(D.tiles.upload(P.tiles, ctx->stream));
    CU(D.steps.upload(P.steps, ctx->stream));
    CU(D.visits.upload(P.visits, ctx->stream));
    CU(D.frows.upload(P.flush_rows, ctx->stream));
    CU(D.fent.upload(P.flush_ent, ctx->stream));
    CU(D.loads.upload(P.loads, ctx->stream));
    CU(ctx->d_tile_counter.ensure(4));
    CU(cudaStreamSynchronize(ctx->stream));
  }
  D.built = true;
  return MASTB200_OK;
}

void fill_params(mastb200_context* ctx, int aux, SweepParams& p) {
  std::memset(&p, 0, sizeof p);
  mastb200_context::DevicePlan& D = ctx->dplan[aux];
  p.tiles = D.tiles.p; p.n_tiles = (int32_t)D.n_tiles;
  p.steps = D.steps.p; p.visits = D.visits.p; p.flush_rows = D.frows.p; p.flush_ent = D.fent.p; p.loads = D.loads.p;
  p.tile_counter = ctx->d_tile_counter.p;
  p.xyz = ctx->d_xyz.p;
  p.cmask = ctx->has_constraints ? ctx->d_cmask.p : nullptr;
  const mastb200_material& m = ctx->mat;
  p.E0 = m.E0;
  // unit-modulus Lame constants written as the reference writes D (isotropic_material_property_card.cpp:512-528)
  p.lam1 = m.nu / (1. - m.nu - 2. * m.nu * m.nu);
  p.mu1 = 1. / 2. / (1. + m.nu);
  p.rho_min = m.rho_min; p.penalty = m.penalty;
  p.alpha_k = m.alpha * (3. * p.lam1 + 2. * p.mu1);
  p.T_const = ctx->T_const; p.T_ref = ctx->T_ref;
  p.nv = ctx->nv;
  p.thermal = ctx->thermal ? 1 : 0;
  p.rho = ctx->rho_set ? (ctx->rho_ext ? ctx->rho_ext : ctx->d_rho.p) : nullptr;
  p.temperature = (ctx->thermal && ctx->T_nodal) ? (ctx->T_ext ? ctx->T_ext : ctx->d_T.p) : nullptr;
  p.aux = aux;
}

FaceLoadPlan face_plan(const mastb200_context* ctx) {
  FaceLoadPlan fp;
  fp.n_loaded_nodes = (int32_t)ctx->fl_node.size();
  fp.nfn = ctx->elem_type == MASTB200_HEX27 ? 9 : 4;
  fp.node = ctx->d_fl_node.p; fp.ptr = ctx->d_fl_ptr.p; fp.face_nodes = ctx->d_fl_face_nodes.p;
  fp.face_k = ctx->d_fl_k.p; fp.pressure = ctx->d_fl_p.p;
  return fp;
}

// external loads of the residual: SURFACE_PRESSURE sides (side_external_residual) and point loads
int apply_external_loads(mastb200_context* ctx, double* dR, const uint8_t* cmask) {
  if (!dR) return MASTB200_OK;
  if (!ctx->fl_node.empty()) {
    CU(launch_face_loads(face_plan(ctx), ctx->d_xyz.p, nullptr, cmask, ctx->nv, dR, nullptr, nullptr, nullptr, ctx->stream));
    ctx->launches++;
  }
  if (!ctx->pl_node.empty()) {
    CU(launch_point_loads((int)ctx->pl_node.size(), ctx->d_pl_node.p, ctx->d_pl_load.p, cmask, ctx->nv, dR, ctx->stream));
    ctx->launches++;
  }
  return MASTB200_OK;
}

// sweeps of a context with auto_localize refresh the ghost values of their input vector themselves (in place, on the
// context stream): AssemblyBase::build_localized_vector inside residual_and_jacobian (nonlinear_implicit_assembly.cpp:93-95)
int localize_if_needed(mastb200_context* ctx, const double* dV, int width) {
  if (!ctx->auto_localize || ctx->neighbours.empty() || !dV) return MASTB200_OK;
  return comm_exchange(ctx, const_cast<double*>(dV), width);
}

int check_ready(mastb200_context* ctx) {
  if (!ctx) return MASTB200_ERR_ARG;
  if (ctx->planning_only)
    return ctx->fail(MASTB200_ERR_CUDA, "planning-only context (device -1): sweeps need a B200; there is no CPU fallback");
  if (!ctx->have_plan) return ctx->fail(MASTB200_ERR_ARG, "call set_mesh, set_dof_map and build_pattern first");
  if (!ctx->have_mat) return ctx->fail(MASTB200_ERR_ARG, "call set_material first");
  return MASTB200_OK;
}


// HEX27 shape tables (libMesh node order, 3x3x3 Gauss rule, xi fastest): N[q][b], dN[q][b][r], w[q]
void hex27_tables(std::vector<double>& tab) {
  static const int8_t C[27][3] = {
      {-1, -1, -1}, {1, -1, -1}, {1, 1, -1}, {-1, 1, -1}, {-1, -1, 1}, {1, -1, 1}, {1, 1, 1}, {-1, 1, 1},
      {0, -1, -1},  {1, 0, -1},  {0, 1, -1}, {-1, 0, -1}, {-1, -1, 0}, {1, -1, 0}, {1, 1, 0}, {-1, 1, 0},
      {0, -1, 1},   {1, 0, 1},   {0, 1, 1},  {-1, 0, 1},  {0, 0, -1},  {0, -1, 0}, {1, 0, 0}, {0, 1, 0},
      {-1, 0, 0},   {0, 0, 1},   {0, 0, 0}};
  auto l = [](int c, double x) { return c < 0 ? 0.5 * x * (x - 1.) : c > 0 ? 0.5 * x * (x + 1.) : (1. - x * x); };
  auto dl = [](int c, double x) { return c < 0 ? x - 0.5 : c > 0 ? x + 0.5 : -2. * x; };
  const double gp[3] = {-0.77459666924148337704, 0., 0.77459666924148337704};
  const double gw[3] = {5. / 9., 8. / 9., 5. / 9.};
  tab.assign(27 * 27 * 4 + 27, 0.);
  for (int q = 0; q < 27; ++q) {
    const double xi = gp[q % 3], eta = gp[(q / 3) % 3], zeta = gp[q / 9];
    for (int b = 0; b < 27; ++b) {
      const double li = l(C[b][0], xi), lj = l(C[b][1], eta), lk = l(C[b][2], zeta);
      tab[q * 27 + b] = li * lj * lk;
      double* d = &tab[27 * 27 + (q * 27 + b) * 3];
      d[0] = dl(C[b][0], xi) * lj * lk;
      d[1] = li * dl(C[b][1], eta) * lk;
      d[2] = li * lj * dl(C[b][2], zeta);
    }
    tab[27 * 27 * 4 + q] = gw[q % 3] * gw[(q / 3) % 3] * gw[q / 9];
  }
}

// HEX8 shape tables in the same layout (2x2x2 Gauss rule, xi fastest, unit weights)
void hex8_tables(std::vector<double>& tab) {
  static const int8_t C[8][3] = {{-1, -1, -1}, {1, -1, -1}, {1, 1, -1}, {-1, 1, -1}, {-1, -1, 1}, {1, -1, 1}, {1, 1, 1}, {-1, 1, 1}};
  const double gp = 0.57735026918962576451;
  tab.assign(8 * 8 * 4 + 8, 0.);
  for (int q = 0; q < 8; ++q) {
    const double xi = (q & 1) ? gp : -gp, eta = (q & 2) ? gp : -gp, zeta = (q & 4) ? gp : -gp;
    for (int b = 0; b < 8; ++b) {
      const double fx = 1. + C[b][0] * xi, fy = 1. + C[b][1] * eta, fz = 1. + C[b][2] * zeta;
      tab[q * 8 + b] = 0.125 * fx * fy * fz;
      double* d = &tab[8 * 8 + (q * 8 + b) * 3];
      d[0] = 0.125 * C[b][0] * fy * fz;
      d[1] = 0.125 * C[b][1] * fx * fz;
      d[2] = 0.125 * C[b][2] * fx * fy;
    }
    tab[8 * 8 * 4 + q] = 1.;
  }
}

int ensure_generic_plan(mastb200_context* ctx) {
  mastb200_context::GenericPlan& G = ctx->gplan;
  if (G.built || ctx->planning_only) return MASTB200_OK;
  const Pattern& P = ctx->pat;
  CU(G.conn.upload(ctx->conn, ctx->stream));
  CU(G.nbr_ptr.upload(P.nbr_ptr, ctx->stream));
  if (ctx->has_diag) CU(G.diag.upload(ctx->diag_elem, ctx->stream));
  CU(G.n2e_ptr.upload(P.n2e_ptr, ctx->stream));
  CU(G.n2e.upload(P.n2e, ctx->stream));
  if (ctx->elem_type == MASTB200_QUAD4) {
    for (int64_t r = 0; r < P.n_owned; ++r)
      if (P.nbr_ptr[r + 1] - P.nbr_ptr[r] > kGatherMaxNbr)
        return ctx->fail(MASTB200_ERR_PLAN, "plan: a node has more than %d neighbours", kGatherMaxNbr);
    CU(G.nbr.upload(P.nbr, ctx->stream));
    if (ctx->nv == 6) {  // the plate kernel runs one thread per neighbour entry: its row, instead of a search per thread
      std::vector<int32_t> er((size_t)P.nbr_ptr[P.n_owned]);
      for (int64_t r = 0; r < P.n_owned; ++r)
        for (int64_t k = P.nbr_ptr[r]; k < P.nbr_ptr[r + 1]; ++k) er[k] = (int32_t)r;
      CU(G.ent_row.upload(er, ctx->stream));
    }
  } else if (ctx->elem_type == MASTB200_HEX8) {  // gather sweep (Green-Lagrange strain) and stress output
    CU(G.nbr.upload(P.nbr, ctx->stream));
    std::vector<double> tab;
    hex8_tables(tab);
    CU(G.tab.upload(tab, ctx->stream));
  } else {  // HEX27
    std::vector<int32_t> color_elems;
    std::vector<uint8_t> posmap;
    std::string e;
    const char* env = getenv("MASTB200_H27_ORDERED");  // 1: the single-launch ordered sweep (experiment, slower on B200)
    if (!(env && env[0] == '1')) {
      e = build_coloring(P, ctx->conn.data(), G.color_ptr, color_elems);
      if (!e.empty()) return ctx->fail(MASTB200_ERR_PLAN, "plan: %s", e.c_str());
      e = build_posmap(P, ctx->conn.data(), posmap);
      if (!e.empty()) return ctx->fail(MASTB200_ERR_PLAN, "plan: %s", e.c_str());
      // MASTB200_H27_FIRST_WRITER=1 (experiment): the colour launches run in colour order, so the first colour to touch a CSR
      // block can store and the later ones add -- no memset of J, no read of 70 % of its values.  Measured neutral on B200
      // (C5: 197.0 ms against 188.9 + 10.1 ms of memset): a 24-byte store into a 32-byte sector costs the L2 the same fill as
      // the read-modify-write did (profiles/r02e_hex27_first_writer.md), so it stays off by default.
      const char* fwe = getenv("MASTB200_H27_FIRST_WRITER");
      G.first_writer = (fwe && fwe[0] == '1') &&
                       flag_first_writers(P, ctx->conn.data(), ctx->diag_elem.empty() ? nullptr : ctx->diag_elem.data(),
                                          G.color_ptr, color_elems, true, posmap);
    } else {
      std::vector<int32_t> order, pred_ptr, pred;
      bool fw = false;
      int brick = 1024;
      if (const char* b = getenv("MASTB200_H27_BRICK")) brick = std::max(8, atoi(b));
      e = build_ordered_scatter(P, ctx->xyz.data(), ctx->conn.data(), ctx->diag_elem.empty() ? nullptr : ctx->diag_elem.data(),
                                brick, order, pred_ptr, pred, posmap, fw);
      if (!e.empty()) return ctx->fail(MASTB200_ERR_PLAN, "plan: %s", e.c_str());
      G.ordered = true; G.first_writer = fw; G.n_order = (int32_t)order.size(); G.epoch = 0;
      CU(G.order.upload(order, ctx->stream));
      CU(G.pred_ptr.upload(pred_ptr, ctx->stream));
      if (pred.empty()) pred.push_back(0);
      CU(G.pred.upload(pred, ctx->stream));
      CU(G.done.ensure((size_t)ctx->n_elem));
      CU(cudaMemsetAsync(G.done.p, 0, size_t(ctx->n_elem) * sizeof(int32_t), ctx->stream));
    }
    std::vector<double> tab;
    hex27_tables(tab);
    CU(G.color_elems.upload(color_elems, ctx->stream));
    CU(G.posmap.upload(posmap, ctx->stream));
    CU(G.node_row.upload(P.node_row, ctx->stream));
    CU(G.tab.upload(tab, ctx->stream));
  }
  CU(cudaStreamSynchronize(ctx->stream));
  G.built = true;
  return MASTB200_OK;
}

// Recognises libMesh build_square numbering of a QUAD4 heat-conduction mesh (node id i + j(nx+1), element id i + j nx,
// node order (i,j), (i+1,j), (i+1,j+1), (i,j+1)) with every node owned and no diagonal elements: the sweep then runs
// with implicit connectivity (quad4_heat_struct_kernel).  MASTB200_NO_STRUCT=1 disables it.
void detect_quad4_struct(mastb200_context* ctx) {
  if (ctx->q4s_tried) return;
  ctx->q4s_tried = true; ctx->q4s_ok = false;
  if (ctx->elem_type != MASTB200_QUAD4 || ctx->nv != 1 || ctx->has_diag || !ctx->owned.empty() || ctx->planning_only) return;
  if (const char* e = getenv("MASTB200_NO_STRUCT")) if (e[0] == '1') return;
  const int32_t* c = ctx->conn.data();
  const int64_t ne = ctx->n_elem, nn = ctx->n_nodes;
  if (ne < 1 || c[0] != 0 || c[1] != 1) return;
  const int64_t npx = c[3];
  if (npx < 2 || nn % npx) return;
  const int64_t nx = npx - 1, ny = nn / npx - 1;
  if (ny < 1 || nx * ny != ne || nx > 2000000 || ny > 2000000) return;
  for (int64_t e = 0; e < ne; ++e) {
    const int64_t i = e % nx, j = e / nx, n0 = i + j * npx;
    const int32_t* q = c + e * 4;
    if (q[0] != n0 || q[1] != n0 + 1 || q[2] != n0 + npx + 1 || q[3] != n0 + npx) return;
  }
  ctx->q4s_nx = (int32_t)nx; ctx->q4s_ny = (int32_t)ny; ctx->q4s_ok = true;
}

// QUAD4 Mindlin plate (row-block gather; per-element local frames when an element is out of the xy-plane)
int run_plate_sweep(mastb200_context* ctx, const double* X, double* R, double* J) {
  int rc;
  if (ctx->thermal)
    return ctx->fail(MASTB200_ERR_UNSUPPORTED, "TEMPERATURE load on the plate element is not on the device path");
  if (ctx->rho_set)
    return ctx->fail(MASTB200_ERR_UNSUPPORTED, "SIMP density field on the plate element is not on the device path");
  // GeomElem's rule (geom_elem.cpp:358-374): an element whose nodes do not share one z gets a local frame
  int32_t local_frames = 0;
  for (int64_t e = 0; e < ctx->n_elem && !local_frames; ++e)
    for (int k = 1; k < 4; ++k)
      if (ctx->xyz[(int64_t)ctx->conn[e * 4 + k] * 3 + 2] != ctx->xyz[(int64_t)ctx->conn[e * 4] * 3 + 2]) local_frames = 1;
  if ((rc = ensure_generic_plan(ctx))) return rc;
  mastb200_context::GenericPlan& G = ctx->gplan;
  const size_t n_dof = size_t(ctx->n_nodes) * ctx->nv;
  const size_t nnz = (size_t)plan_nnz(ctx);
  const double* dX = nullptr;
  double *dR = nullptr, *dJ = nullptr;
  if ((rc = resolve_in(ctx, X, n_dof, ctx->d_X, &dX))) return rc;
  if ((rc = localize_if_needed(ctx, dX, ctx->nv))) return rc;
  if ((rc = resolve_out(ctx, R, n_dof, ctx->d_R, &dR))) return rc;
  if ((rc = resolve_out(ctx, J, nnz, ctx->d_J, &dJ))) return rc;
  const mastb200_material& m = ctx->mat;
  CU(cudaEventRecord(ctx->ev[0], ctx->stream));
  if (dR && ctx->has_ghosts) CU(cudaMemsetAsync(dR, 0, n_dof * sizeof(double), ctx->stream));
  CU(cudaEventRecord(ctx->ev[1], ctx->stream));
  PlateParams p;
  p.n_owned = ctx->pat.n_owned; p.n_entries = ctx->pat.nbr_ptr[ctx->pat.n_owned];
  p.owned_nodes = ctx->d_owned_nodes.p;
  p.nbr_ptr = G.nbr_ptr.p; p.nbr = G.nbr.p; p.n2e_ptr = G.n2e_ptr.p; p.n2e = G.n2e.p; p.conn = G.conn.p;
  p.ent_row = G.ent_row.p;
  p.diag_elem = ctx->has_diag ? G.diag.p : nullptr;
  p.cmask = ctx->has_constraints ? ctx->d_cmask.p : nullptr;
  p.xyz = ctx->d_xyz.p; p.X = dX; p.J = dJ; p.R = dR;
  // StiffnessMatrix2D (isotropic_material_property_card.cpp:395-426)
  if (!m.plane_strain) {
    p.c11 = m.E0 / (1. - m.nu * m.nu); p.c12 = m.E0 * m.nu / (1. - m.nu * m.nu);
  } else {
    p.c11 = m.E0 * (1. - m.nu) / ((1. + m.nu) * (1. - 2. * m.nu)); p.c12 = m.E0 * m.nu / ((1. + m.nu) * (1. - 2. * m.nu));
  }
  p.c33 = m.E0 / 2. / (1. + m.nu);
  p.shear_G = m.E0 / 2. / (1. + m.nu);
  p.thickness = m.thickness; p.offset = m.offset; p.kappa = m.kappa;
  p.local_frames = local_frames;
  p.nonlinear = ctx->mat.nonlinear_strain ? 1 : 0;
  CU(launch_quad4_plate(p, ctx->stream));
  ctx->launches++;
  CU(cudaEventRecord(ctx->ev[2], ctx->stream));
  if (dR && !ctx->pl_node.empty()) {
    CU(launch_point_loads((int)ctx->pl_node.size(), ctx->d_pl_node.p, ctx->d_pl_load.p, p.cmask, ctx->nv, dR, ctx->stream));
    ctx->launches++;
  }
  CU(cudaEventRecord(ctx->ev[3], ctx->stream));
  ctx->timed = true;
  ctx->last_R = dR;
  if ((rc = finish_out(ctx, R, n_dof, dR))) return rc;
  if ((rc = finish_out(ctx, J, nnz, dJ))) return rc;
  CU(cudaStreamSynchronize(ctx->stream));
  return MASTB200_OK;
}

// QUAD4 heat conduction (row gather) and HEX27 elasticity (coloured DMMA scatter)
int run_hex8_gather(mastb200_context* ctx, const double* X, double* R, double* J);
int run_generic_sweep(mastb200_context* ctx, const double* X, double* R, double* J) {
  int rc;
  const bool heat = ctx->elem_type == MASTB200_QUAD4;
  if (heat && ctx->nv == 6) return run_plate_sweep(ctx, X, R, J);
  if (heat && ctx->nv != 1) return ctx->fail(MASTB200_ERR_ARG, "QUAD4: nv = 1 (heat conduction) or 6 (Mindlin plate)");
  if (!heat && ctx->nv == 1) return ctx->fail(MASTB200_ERR_ARG, "HEX27 elasticity needs nv = 3 or 6");
  if (!heat && ctx->mat.nonlinear_strain) return run_hex8_gather(ctx, X, R, J);  // HEX27, Green-Lagrange: row-owner gather
  if (heat) detect_quad4_struct(ctx);
  const bool q4s = heat && ctx->q4s_ok;
  ctx->last_q4s = q4s;
  if (!q4s && (rc = ensure_generic_plan(ctx))) return rc;
  mastb200_context::GenericPlan& G = ctx->gplan;
  const size_t n_dof = size_t(ctx->n_nodes) * ctx->nv;
  const size_t nnz = (size_t)plan_nnz(ctx);
  const double* dX = nullptr;
  double *dR = nullptr, *dJ = nullptr;
  if ((rc = resolve_in(ctx, X, n_dof, ctx->d_X, &dX))) return rc;
  if ((rc = localize_if_needed(ctx, dX, ctx->nv))) return rc;
  if ((rc = resolve_out(ctx, R, n_dof, ctx->d_R, &dR))) return rc;
  if ((rc = resolve_out(ctx, J, nnz, ctx->d_J, &dJ))) return rc;
  const mastb200_material& m = ctx->mat;
  CU(cudaEventRecord(ctx->ev[0], ctx->stream));
  if (heat) {
    if (dR && ctx->has_ghosts) CU(cudaMemsetAsync(dR, 0, n_dof * sizeof(double), ctx->stream));
    CU(cudaEventRecord(ctx->ev[1], ctx->stream));
    GatherParams p;
    p.n_owned = ctx->pat.n_owned; p.owned_nodes = ctx->d_owned_nodes.p;
    if (!q4s) {
      p.nbr_ptr = G.nbr_ptr.p; p.nbr = G.nbr.p; p.n2e_ptr = G.n2e_ptr.p; p.n2e = G.n2e.p; p.conn = G.conn.p;
      p.diag_elem = ctx->has_diag ? G.diag.p : nullptr;
    }
    p.cmask = ctx->has_constraints ? ctx->d_cmask.p : nullptr;
    p.xyz = ctx->d_xyz.p; p.X = dX; p.J = dJ; p.R = dR;
    p.k0 = m.k0; p.k_beta = m.k_beta; p.thickness = m.thickness; p.heat_source = m.heat_source; p.k_of_T = m.k_of_T;
    if (ctx->have_heat_sides) { p.sl_ptr = ctx->d_hs_ptr.p; p.sl_side = ctx->d_hs_side.p; p.sl_val = ctx->d_hs_val.p; p.rad_sbc = ctx->rad_sbc; p.rad_T0 = ctx->rad_T0; }
    if (q4s) CU(launch_quad4_heat_struct(p, ctx->q4s_nx, ctx->q4s_ny, ctx->stream));
    else CU(launch_quad4_heat(p, ctx->stream));
    ctx->launches++;
  } else {
    // Two-pass assembly of the 3-dof Jacobian (default when J is requested): chunks of consecutive element ids; pass 1 stores
    // K_e and f_e element-major (one launch per chunk, no colours, no read-modify-write), pass 2 sums them per owned row and
    // writes the CSR rows as whole runs.  The coloured scatter touches every CSR value 1.42 times in 24-byte pieces
    // (partial-sector fills: 160 KB of DRAM traffic per element); here everything streams (profiles/r02g_hex27_two_pass.md).
    // MASTB200_H27_TWO_PASS=0 keeps the coloured scatter; MASTB200_H27_CHUNK sets the elements per chunk.
    bool two_pass = ctx->nv == 3 && !ctx->has_diag && !G.ordered && dJ != nullptr;
    if (const char* e2 = getenv("MASTB200_H27_TWO_PASS")) if (e2[0] == '0') two_pass = false;
    if (two_pass && !G.tp_built) {
      int64_t chunk = 131072;
      if (const char* ce = getenv("MASTB200_H27_CHUNK")) chunk = std::max<int64_t>(1, atoll(ce));
      chunk = std::min<int64_t>(chunk, ctx->n_elem);
      const Pattern& P = ctx->pat;
      std::vector<int32_t> rows, ea;
      std::vector<int64_t> ea_ptr;
      std::vector<uint8_t> row_first;
      std::vector<int32_t> stamp((size_t)ctx->n_nodes, -1);
      G.tp_row_ptr.assign(1, 0);
      for (int64_t c0 = 0, c = 0; c0 < ctx->n_elem; c0 += chunk, ++c) {
        const int64_t c1 = std::min<int64_t>(c0 + chunk, ctx->n_elem);
        const size_t first = rows.size();
        for (int64_t e = c0; e < c1; ++e)
          for (int a = 0; a < 27; ++a) {
            const int32_t nd = ctx->conn[e * 27 + a];
            if (stamp[nd] == (int32_t)c) continue;
            stamp[nd] = (int32_t)c;
            if (P.node_row[nd] >= 0) rows.push_back(P.node_row[nd]);
          }
        std::sort(rows.begin() + first, rows.end());
        G.tp_row_ptr.push_back((int64_t)rows.size());
        // per row of the chunk: its elements inside the chunk (ascending id) with the node's local index, so that the gather
        // starts its loads without walking n2e / conn
        ea_ptr.push_back((int64_t)ea.size());
        for (size_t k = first; k < rows.size(); ++k) {
          const int32_t nd = P.owned_nodes[rows[k]];
          const int64_t q0 = P.n2e_ptr[nd], q1 = P.n2e_ptr[nd + 1];
          row_first.push_back((q0 == q1 || P.n2e[q0] >= c0) ? 1 : 0);
          for (int64_t q = q0; q < q1; ++q) {
            const int64_t e = P.n2e[q];
            if (e < c0) continue;
            if (e >= c1) break;
            int a = 0;
            for (int b = 0; b < 27; ++b)
              if (ctx->conn[e * 27 + b] == nd) a = b;
            ea.push_back((int32_t)((e - c0) * 27 + a));
          }
          ea_ptr.push_back((int64_t)ea.size());
        }
      }
      if (rows.empty()) rows.push_back(0);
      if (ea.empty()) ea.push_back(0);
      if (row_first.empty()) row_first.push_back(0);
      CU(G.tp_rows.upload(rows, ctx->stream));
      CU(G.tp_ea.upload(ea, ctx->stream));
      CU(G.tp_ea_ptr.upload(ea_ptr, ctx->stream));
      CU(G.tp_first.upload(row_first, ctx->stream));
      CU(G.tp_K.ensure((size_t)chunk * 27 * 27 * 9));
      CU(G.tp_f.ensure((size_t)chunk * 27 * 3));
      if (!ctx->d_owned_nodes.p) return ctx->fail(MASTB200_ERR_PLAN, "two-pass HEX27 assembly: owned-node list missing");
      G.tp_chunk = (int32_t)chunk;
      G.tp_built = true;
    }
    // the scatter accumulates: R and J start from zero (nonlinear_implicit_assembly.cpp:81-82) -- except that the ordered
    // sweep of the 3-dof system stores the first contribution to every CSR block, so J needs no memset there
    const bool fw = G.first_writer && ctx->nv == 3;
    if (dR) CU(cudaMemsetAsync(dR, 0, n_dof * sizeof(double), ctx->stream));
    if (dJ && !fw && !two_pass) CU(cudaMemsetAsync(dJ, 0, nnz * sizeof(double), ctx->stream));
    CU(cudaEventRecord(ctx->ev[1], ctx->stream));
    Hex27Params p;
    p.conn = G.conn.p; p.posmap = G.posmap.p; p.node_row = G.node_row.p; p.nbr_ptr = G.nbr_ptr.p; p.tab = G.tab.p;
    p.diag_elem = ctx->has_diag ? G.diag.p : nullptr;
    p.cmask = ctx->has_constraints ? ctx->d_cmask.p : nullptr;
    p.xyz = ctx->d_xyz.p; p.X = dX; p.J = dJ; p.R = dR;
    p.rho = ctx->rho_set ? (ctx->rho_ext ? ctx->rho_ext : ctx->d_rho.p) : nullptr;
    p.temperature = (ctx->thermal && ctx->T_nodal) ? (ctx->T_ext ? ctx->T_ext : ctx->d_T.p) : nullptr;
    p.E0 = m.E0;
    p.lam1 = m.nu / (1. - m.nu - 2. * m.nu * m.nu);
    p.mu1 = 1. / 2. / (1. + m.nu);
    p.rho_min = m.rho_min; p.penalty = m.penalty;
    p.alpha_k = m.alpha * (3. * p.lam1 + 2. * p.mu1);
    p.T_const = ctx->T_const; p.T_ref = ctx->T_ref;
    p.nv = ctx->nv; p.thermal = ctx->thermal ? 1 : 0;
    if (two_pass) {
      Hex27GatherParams g;
      g.owned_nodes = ctx->d_owned_nodes.p; g.nbr_ptr = G.nbr_ptr.p; g.n2e_ptr = G.n2e_ptr.p; g.n2e = G.n2e.p;
      g.conn = G.conn.p; g.posmap = G.posmap.p; g.pos_flagged = G.first_writer ? 1 : 0;
      g.tmp_K = G.tp_K.p; g.tmp_f = dR ? G.tp_f.p : nullptr; g.J = dJ; g.R = dR;
      p.J = nullptr; p.R = nullptr;
      p.tmp_K = G.tp_K.p; p.tmp_f = dR ? G.tp_f.p : nullptr;
      p.pos_flagged = G.first_writer ? 1 : 0;
      bool h27_sym = true;  // MASTB200_H27_SYM=0: pass 1 contracts all sixteen tile pairs (the colour kernel's tiling)
      if (const char* es = getenv("MASTB200_H27_SYM")) if (es[0] == '0') h27_sym = false;
      const int64_t n_chunks = (int64_t)G.tp_row_ptr.size() - 1;
      for (int64_t c = 0; c < n_chunks; ++c) {
        p.elems = nullptr; p.elem_first = (int32_t)(c * G.tp_chunk);
        p.n_elems = (int32_t)std::min<int64_t>(G.tp_chunk, ctx->n_elem - (int64_t)c * G.tp_chunk);
        if (h27_sym) CU(launch_hex27_sym_pass1(p, ctx->sm_count, ctx->stream));
        else CU(launch_hex27_sweep(p, ctx->sm_count, ctx->stream));
        g.elem_first = p.elem_first; g.n_elems = p.n_elems;
        g.rows = G.tp_rows.p + G.tp_row_ptr[c]; g.n_rows = G.tp_row_ptr[c + 1] - G.tp_row_ptr[c];
        g.ea_ptr = G.tp_ea_ptr.p + G.tp_row_ptr[c] + c; g.ea = G.tp_ea.p; g.row_first = G.tp_first.p + G.tp_row_ptr[c];
        CU(launch_hex27_row_gather(g, ctx->stream));
        ctx->launches += 2;
      }
    } else if (G.ordered) {  // one launch: the plan's order with done flags
      if (++G.epoch == INT32_MAX) {
        G.epoch = 1;
        CU(cudaMemsetAsync(G.done.p, 0, size_t(ctx->n_elem) * sizeof(int32_t), ctx->stream));
      }
      p.elems = G.order.p; p.n_elems = G.n_order;
      p.pred_ptr = G.pred_ptr.p; p.pred = G.pred.p; p.done = G.done.p; p.epoch = G.epoch;
      p.first_writer = fw ? 1 : 0;
      p.pos_flagged = G.first_writer ? 1 : 0;
      CU(launch_hex27_sweep(p, ctx->sm_count, ctx->stream));
      ctx->launches++;
    } else {
      const int nc = (int)G.color_ptr.size() - 1;
      p.first_writer = fw ? 1 : 0;
      p.pos_flagged = G.first_writer ? 1 : 0;
      for (int c = 0; c < nc; ++c) {
        p.elems = G.color_elems.p + G.color_ptr[c];
        p.n_elems = G.color_ptr[c + 1] - G.color_ptr[c];
        CU(launch_hex27_sweep(p, ctx->sm_count, ctx->stream));
        ctx->launches++;
      }
    }
  }
  CU(cudaEventRecord(ctx->ev[2], ctx->stream));
  if (!heat && (rc = apply_external_loads(ctx, dR, ctx->has_constraints ? ctx->d_cmask.p : nullptr))) return rc;  // SURFACE_PRESSURE on HEX27 faces, point loads
  CU(cudaEventRecord(ctx->ev[3], ctx->stream));
  ctx->timed = true;
  ctx->last_R = dR;
  if ((rc = finish_out(ctx, R, n_dof, dR))) return rc;
  if ((rc = finish_out(ctx, J, nnz, dJ))) return rc;
  CU(cudaStreamSynchronize(ctx->stream));
  return MASTB200_OK;
}

// Recognises libMesh build_cube numbering (node id i + j(nx+1) + k(nx+1)(ny+1), element id i + j nx + k nx ny, HEX8
// node order) with every node owned or the owned nodes forming a range of z-planes, and builds the work units of the
// implicit-connectivity sweep: columns of <= 7x7 owned nodes cut into z-segments.  MASTB200_NO_STRUCT=1 disables it.
int ensure_struct_plan(mastb200_context* ctx) {
  mastb200_context::StructPlan& S = ctx->splan;
  if (S.tried) return MASTB200_OK;
  S.tried = true;
  if (ctx->elem_type != MASTB200_HEX8 || ctx->has_diag || ctx->planning_only) return MASTB200_OK;
  if (const char* e = getenv("MASTB200_NO_STRUCT")) if (e[0] == '1') return MASTB200_OK;
  if ((size_t)ctx->max_optin < hex8_struct_smem_bytes()) return MASTB200_OK;
  const int32_t* c = ctx->conn.data();
  const int64_t ne = ctx->n_elem, nn = ctx->n_nodes;
  if (c[0] != 0 || c[1] != 1) return MASTB200_OK;
  const int64_t npx = c[3], npxy = c[4];
  if (npx < 2 || npxy < 2 * npx || npxy % npx) return MASTB200_OK;
  const int64_t nx = npx - 1, ny = npxy / npx - 1;
  if (ny < 1 || ne % (nx * ny)) return MASTB200_OK;
  const int64_t nz = ne / (nx * ny);
  if (npxy * (nz + 1) != nn || nx > 32000 || ny > 32000) return MASTB200_OK;
  bool match = true;
  for (int64_t e = 0; e < ne && match; ++e) {
    const int64_t i = e % nx, j = (e / nx) % ny, k = e / (nx * ny);
    const int64_t n0 = i + j * npx + k * npxy;
    const int32_t* q = c + e * 8;
    match = q[0] == n0 && q[1] == n0 + 1 && q[2] == n0 + npx + 1 && q[3] == n0 + npx && q[4] == n0 + npxy &&
            q[5] == n0 + npxy + 1 && q[6] == n0 + npxy + npx + 1 && q[7] == n0 + npxy + npx;
  }
  if (!match) return MASTB200_OK;
  int64_t oz0 = 0, oz1 = nz + 1;
  if (!ctx->owned.empty()) {  // owned nodes must be whole z-planes [oz0, oz1)
    oz0 = -1; oz1 = -1;
    for (int64_t k = 0; k <= nz; ++k) {
      const bool o = ctx->owned[k * npxy] != 0;
      for (int64_t n = k * npxy; n < (k + 1) * npxy; ++n)
        if ((ctx->owned[n] != 0) != o) return MASTB200_OK;
      if (o) {
        if (oz0 < 0) oz0 = k;
        else if (oz1 >= 0) return MASTB200_OK;  // second range
      } else if (oz0 >= 0 && oz1 < 0) {
        oz1 = k;
      }
    }
    if (oz0 < 0) return MASTB200_OK;
    if (oz1 < 0) oz1 = nz + 1;
  }
  // owned nodes of a direction in chunks of at most (tile elements - 1), as even as possible
  auto chunks = [](int64_t n_nodes_dir, int64_t mx, std::vector<std::pair<int, int>>& out) {
    const int64_t nt = (n_nodes_dir + mx - 1) / mx, base = n_nodes_dir / nt, rem = n_nodes_dir % nt;
    int64_t o = 0;
    for (int64_t t = 0; t < nt; ++t) {
      const int64_t len = base + (t < rem ? 1 : 0);
      out.emplace_back((int)o, (int)len);
      o += len;
    }
  };
  std::vector<std::pair<int, int>> cx, cy;
  chunks(nx + 1, 7, cx);                  // 8 elements (9 node columns) per tile row
  chunks(ny + 1, kStructTileY - 1, cy);   // kStructTileY elements per tile column
  const int64_t n_tiles = (int64_t)cx.size() * cy.size(), planes = oz1 - oz0;
  // z-segments per tile: large meshes get ~20 units per SM with at least 24 planes each (a segment costs one extra element
  // plane); meshes too small for that take the split that minimises waves x (planes per segment + 1)
  auto pick_segments = [&](int64_t tiles, int64_t ctas) {
    int64_t ns = (20 * ctas + tiles - 1) / tiles;
    const int64_t cap = std::max<int64_t>(1, planes / 24);
    ns = std::max<int64_t>(1, std::min<int64_t>(ns, cap));
    if (tiles * cap >= 2 * ctas) return ns;
    int64_t best = 1, best_cost = INT64_MAX;
    for (int64_t c = 1; c <= std::max<int64_t>(1, planes / 2); ++c) {
      const int64_t waves = (tiles * c + ctas - 1) / ctas, cost = waves * ((planes + c - 1) / c + 1);
      if (cost < best_cost) { best_cost = cost; best = c; }
    }
    return best;
  };
  const int64_t nseg = pick_segments(n_tiles, (int64_t)ctx->sm_count * kStructCtasPerSm);
  std::vector<StructUnit> units;
  std::vector<int64_t> cost;
  S.n_visits = 0;
  for (size_t ty = 0; ty < cy.size(); ++ty)
    for (size_t tx = 0; tx < cx.size(); ++tx)
      for (int64_t sg = 0; sg < nseg; ++sg) {
        StructUnit u;
        u.o0x = (int16_t)cx[tx].first; u.nox = (int16_t)cx[tx].second;
        u.o0y = (int16_t)cy[ty].first; u.noy = (int16_t)cy[ty].second;
        const int64_t e0x = std::max<int64_t>(u.o0x - 1, 0), e1x = std::min<int64_t>(u.o0x + u.nox, nx);
        const int64_t e0y = std::max<int64_t>(u.o0y - 1, 0), e1y = std::min<int64_t>(u.o0y + u.noy, ny);
        u.e0x = (int16_t)e0x; u.nex = (int16_t)(e1x - e0x);
        u.e0y = (int16_t)e0y; u.ney = (int16_t)(e1y - e0y);
        u.z0 = (int32_t)(oz0 + planes * sg / nseg); u.z1 = (int32_t)(oz0 + planes * (sg + 1) / nseg);
        if (u.z1 <= u.z0) continue;
        const int64_t s_begin = u.z0 > 0 ? u.z0 - 1 : 0, e_planes = std::min<int64_t>(u.z1, nz) - s_begin;
        units.push_back(u);
        cost.push_back((int64_t)u.nex * u.ney * (u.z1 - s_begin));
        S.n_visits += (int64_t)u.nex * u.ney * std::max<int64_t>(e_planes, 0);
      }
  std::vector<int32_t> order(units.size());
  for (size_t i = 0; i < order.size(); ++i) order[i] = (int32_t)i;
  std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return cost[a] > cost[b]; });
  std::vector<StructUnit> sorted(units.size());
  for (size_t i = 0; i < order.size(); ++i) sorted[i] = units[order[i]];
  {  // units of the segmented sweep of host-buffer calls, by z-segment (ascending), cost-sorted inside a segment.  Small
     // meshes: the same units as above.  Tall meshes get their own, uneven split -- short segments at both ends (the planes of
     // X the first kernel waits for and the rows of R that leave after the last one are the exposed copies), long ones in
     // the middle (every segment is a launch with a partly idle tail), each at most 2.5 x its predecessor (the planes of a
     // segment arrive over PCIe while the one before is swept: ~19 us against ~54 us per plane at 200^2 nodes).
    std::vector<int32_t> zb;  // segment boundaries (node planes)
    if (planes >= 160 && !getenv("MASTB200_SEG_UNIFORM")) {  // (at 100 planes per rank on two GPUs the even split measured faster)
      const double f[7] = {0., 0.05, 0.175, 0.475, 0.775, 0.925, 1.};
      for (int k = 0; k < 7; ++k) zb.push_back((int32_t)(oz0 + (int64_t)(planes * f[k] + 0.5)));
    } else {
      for (int64_t sg = 0; sg <= nseg; ++sg) zb.push_back((int32_t)(oz0 + planes * sg / nseg));
    }
    std::vector<StructUnit> grouped;
    S.g_ptr.assign(1, 0); S.g_z0.clear(); S.g_z1.clear();
    for (size_t sg = 0; sg + 1 < zb.size(); ++sg) {
      const int32_t z0 = zb[sg], z1 = zb[sg + 1];
      if (z1 <= z0) continue;
      std::vector<StructUnit> seg_units;
      for (size_t ty = 0; ty < cy.size(); ++ty)
        for (size_t tx = 0; tx < cx.size(); ++tx) {
          StructUnit u;
          u.o0x = (int16_t)cx[tx].first; u.nox = (int16_t)cx[tx].second;
          u.o0y = (int16_t)cy[ty].first; u.noy = (int16_t)cy[ty].second;
          const int64_t e0x = std::max<int64_t>(u.o0x - 1, 0), e1x = std::min<int64_t>(u.o0x + u.nox, nx);
          const int64_t e0y = std::max<int64_t>(u.o0y - 1, 0), e1y = std::min<int64_t>(u.o0y + u.noy, ny);
          u.e0x = (int16_t)e0x; u.nex = (int16_t)(e1x - e0x);
          u.e0y = (int16_t)e0y; u.ney = (int16_t)(e1y - e0y);
          u.z0 = z0; u.z1 = z1;
          seg_units.push_back(u);
        }
      std::stable_sort(seg_units.begin(), seg_units.end(),
                       [](const StructUnit& a, const StructUnit& b) { return (int)a.nex * a.ney > (int)b.nex * b.ney; });
      grouped.insert(grouped.end(), seg_units.begin(), seg_units.end());
      S.g_ptr.push_back((int32_t)grouped.size()); S.g_z0.push_back(z0); S.g_z1.push_back(z1);
    }
    CU(S.g_units.upload(grouped, ctx->stream));
  }
  {  // units of the structured sensitivity sweep: columns of <= 15 x 11 owned nodes (16 x 12 elements), two CTAs per SM
    std::vector<std::pair<int, int>> sx, sy;
    chunks(nx + 1, kSensTileX, sx);
    chunks(ny + 1, kSensTileY, sy);
    const int64_t s_tiles = (int64_t)sx.size() * sy.size();
    const int64_t sseg = pick_segments(s_tiles, (int64_t)ctx->sm_count * 2);
    std::vector<StructUnit> su;
    std::vector<int64_t> sc;
    for (size_t ty = 0; ty < sy.size(); ++ty)
      for (size_t tx = 0; tx < sx.size(); ++tx)
        for (int64_t sg = 0; sg < sseg; ++sg) {
          StructUnit u;
          u.o0x = (int16_t)sx[tx].first; u.nox = (int16_t)sx[tx].second;
          u.o0y = (int16_t)sy[ty].first; u.noy = (int16_t)sy[ty].second;
          const int64_t e0x = std::max<int64_t>(u.o0x - 1, 0), e1x = std::min<int64_t>(u.o0x + u.nox, nx);
          const int64_t e0y = std::max<int64_t>(u.o0y - 1, 0), e1y = std::min<int64_t>(u.o0y + u.noy, ny);
          u.e0x = (int16_t)e0x; u.nex = (int16_t)(e1x - e0x);
          u.e0y = (int16_t)e0y; u.ney = (int16_t)(e1y - e0y);
          u.z0 = (int32_t)(oz0 + planes * sg / sseg); u.z1 = (int32_t)(oz0 + planes * (sg + 1) / sseg);
          if (u.z1 <= u.z0) continue;
          const int64_t s_begin = u.z0 > 0 ? u.z0 - 1 : 0;
          su.push_back(u);
          sc.push_back((int64_t)u.nex * u.ney * (u.z1 - s_begin));
        }
    std::vector<int32_t> so(su.size());
    for (size_t i = 0; i < so.size(); ++i) so[i] = (int32_t)i;
    std::stable_sort(so.begin(), so.end(), [&](int32_t a, int32_t b) { return sc[a] > sc[b]; });
    std::vector<StructUnit> ss(su.size());
    for (size_t i = 0; i < so.size(); ++i) ss[i] = su[so[i]];
    CU(S.s_units.upload(ss, ctx->stream));
    S.s_n_units = (int32_t)ss.size();
  }
  // units of the register-accumulating march (3-dof system): columns of <= 16 x 4 owned nodes cut into z-segments.
  // Opt-in with MASTB200_STRUCT_KERNEL=march: parity-green but slower than the lock-step kernel so far
  // (profiles/r01h_hex8_march_ncu.md).
  {
    bool want = false;
    if (const char* e = getenv("MASTB200_STRUCT_KERNEL"))
      want = e[0] == 'm' && ctx->nv == 3 && (size_t)ctx->max_optin >= hex8_march_smem_bytes();
    S.march = want;
    // MASTB200_STRUCT_KERNEL=slots: element results in shared-memory slots, gather straight to the CSR (sweep_hex8_slots.cu);
    // same units as the lock-step kernel
    S.slots = false;
    if (const char* e = getenv("MASTB200_STRUCT_KERNEL"))
      S.slots = e[0] == 's' && ctx->nv == 3 && (size_t)ctx->max_optin >= hex8_slots_smem_bytes();
    if (want) {
      std::vector<std::pair<int, int>> mx, my;
      chunks(nx + 1, kMarchTileX, mx);
      chunks(ny + 1, kMarchTileY, my);
      const int64_t m_tiles = (int64_t)mx.size() * my.size();
      const int64_t mseg = pick_segments(m_tiles, ctx->sm_count);
      std::vector<StructUnit> mu;
      std::vector<int64_t> mc;
      S.m_n_visits = 0;
      for (size_t ty = 0; ty < my.size(); ++ty)
        for (size_t tx = 0; tx < mx.size(); ++tx)
          for (int64_t sg = 0; sg < mseg; ++sg) {
            StructUnit u;
            u.o0x = (int16_t)mx[tx].first; u.nox = (int16_t)mx[tx].second;
            u.o0y = (int16_t)my[ty].first; u.noy = (int16_t)my[ty].second;
            const int64_t e0x = std::max<int64_t>(u.o0x - 1, 0), e1x = std::min<int64_t>(u.o0x + u.nox, nx);
            const int64_t e0y = std::max<int64_t>(u.o0y - 1, 0), e1y = std::min<int64_t>(u.o0y + u.noy, ny);
            u.e0x = (int16_t)e0x; u.nex = (int16_t)(e1x - e0x);
            u.e0y = (int16_t)e0y; u.ney = (int16_t)(e1y - e0y);
            u.z0 = (int32_t)(oz0 + planes * sg / mseg); u.z1 = (int32_t)(oz0 + planes * (sg + 1) / mseg);
            if (u.z1 <= u.z0) continue;
            const int64_t e_begin = u.z0 > 0 ? u.z0 - 1 : 0, e_end = std::min<int64_t>(u.z1 - 1, nz - 1);
            mu.push_back(u);
            mc.push_back((int64_t)u.nex * u.ney * (e_end - e_begin + 1));
            S.m_n_visits += (int64_t)u.nex * u.ney * std::max<int64_t>(e_end - e_begin + 1, 0);
          }
      std::vector<int32_t> mo(mu.size());
      for (size_t i = 0; i < mo.size(); ++i) mo[i] = (int32_t)i;
      std::stable_sort(mo.begin(), mo.end(), [&](int32_t a, int32_t b) { return mc[a] > mc[b]; });
      std::vector<StructUnit> ms(mu.size());
      for (size_t i = 0; i < mo.size(); ++i) ms[i] = mu[mo[i]];
      CU(S.m_units.upload(ms, ctx->stream));
      S.m_n_units = (int32_t)ms.size();
      S.m_n_ctas = std::max(1, std::min(S.m_n_units, ctx->sm_count));
    }
  }
  std::vector<int32_t> arr_off(124, 0);
  for (int a = 0; a < 42; ++a) arr_off[a] = kStructPosIn[a] * kStructSlots;
  for (int a = 0; a < 81; ++a) arr_off[42 + a] = kStructPosCr[a] * kStructSlots;
  CU(S.units.upload(sorted, ctx->stream));
  CU(S.arr_off.upload(arr_off, ctx->stream));
  CU(S.zeros.ensure(kStructZeroDoubles));
  CU(cudaMemsetAsync(S.zeros.p, 0, kStructZeroDoubles * sizeof(double), ctx->stream));
  CU(S.nbr_ptr.upload(ctx->pat.nbr_ptr, ctx->stream));
  CU(ctx->d_tile_counter.ensure(40));
  CU(cudaStreamSynchronize(ctx->stream));
  S.nx = (int32_t)nx; S.ny = (int32_t)ny; S.nz = (int32_t)nz; S.oz0 = (int32_t)oz0; S.oz1 = (int32_t)oz1;
  S.n_units = (int32_t)sorted.size();
  S.n_ctas = std::max(1, std::min(S.n_units, ctx->sm_count * kStructCtasPerSm));
  S.ok = true;
  return MASTB200_OK;
}

// HEX8 with Green-Lagrange strain: row-owner gather (sweep_hex8_gather.cu)
int run_hex8_gather(mastb200_context* ctx, const double* X, double* R, double* J) {
  int rc;
  if ((rc = ensure_generic_plan(ctx))) return rc;
  mastb200_context::GenericPlan& G = ctx->gplan;
  const Pattern& P = ctx->pat;
  const bool h27 = ctx->elem_type == MASTB200_HEX27;
  const int max_nbr = h27 ? 128 : 32;  // one 3x3 block per lane (HEX8), four per lane (HEX27: a vertex node has 125)
  for (int64_t r = 0; r < P.n_owned; ++r)
    if (P.nbr_ptr[r + 1] - P.nbr_ptr[r] > max_nbr)
      return ctx->fail(MASTB200_ERR_PLAN, "plan: a node has more than %d neighbours (gather sweep)", max_nbr);
  if (h27 && !G.nbr.p) {  // the HEX27 plan keeps the column list on the host until the gather sweep asks for it
    CU(G.nbr.upload(P.nbr, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
  }
  const size_t n_dof = size_t(ctx->n_nodes) * ctx->nv;
  const size_t nnz = (size_t)plan_nnz(ctx);
  const double* dX = nullptr;
  double *dR = nullptr, *dJ = nullptr;
  if ((rc = resolve_in(ctx, X, n_dof, ctx->d_X, &dX))) return rc;
  if ((rc = localize_if_needed(ctx, dX, ctx->nv))) return rc;
  if ((rc = resolve_out(ctx, R, n_dof, ctx->d_R, &dR))) return rc;
  if ((rc = resolve_out(ctx, J, nnz, ctx->d_J, &dJ))) return rc;
  const mastb200_material& m = ctx->mat;
  CU(cudaEventRecord(ctx->ev[0], ctx->stream));
  if (dR && ctx->has_ghosts) CU(cudaMemsetAsync(dR, 0, n_dof * sizeof(double), ctx->stream));
  CU(cudaEventRecord(ctx->ev[1], ctx->stream));
  SolidGatherParams p;
  p.n_owned = P.n_owned; p.owned_nodes = ctx->d_owned_nodes.p;
  p.nbr_ptr = G.nbr_ptr.p; p.nbr = G.nbr.p; p.n2e_ptr = G.n2e_ptr.p; p.n2e = G.n2e.p; p.conn = G.conn.p;
  p.diag_elem = ctx->has_diag ? G.diag.p : nullptr;
  p.cmask = ctx->has_constraints ? ctx->d_cmask.p : nullptr;
  p.xyz = ctx->d_xyz.p; p.X = dX; p.J = dJ; p.R = dR;
  p.rho = ctx->rho_set ? (ctx->rho_ext ? ctx->rho_ext : ctx->d_rho.p) : nullptr;
  p.temperature = (ctx->thermal && ctx->T_nodal) ? (ctx->T_ext ? ctx->T_ext : ctx->d_T.p) : nullptr;
  p.E0 = m.E0;
  p.lam1 = m.nu / (1. - m.nu - 2. * m.nu * m.nu);
  p.mu1 = 1. / 2. / (1. + m.nu);
  p.rho_min = m.rho_min; p.penalty = m.penalty;
  p.alpha_k = m.alpha * (3. * p.lam1 + 2. * p.mu1);
  p.T_const = ctx->T_const; p.T_ref = ctx->T_ref;
  p.nv = ctx->nv; p.thermal = ctx->thermal ? 1 : 0; p.nonlinear = m.nonlinear_strain ? 1 : 0;
  p.tab = G.tab.p;
  if (h27) CU(launch_hex27_gather(p, ctx->stream));
  else CU(launch_hex8_gather(p, ctx->stream));
  ctx->launches++;
  CU(cudaEventRecord(ctx->ev[2], ctx->stream));
  if ((rc = apply_external_loads(ctx, dR, p.cmask))) return rc;
  CU(cudaEventRecord(ctx->ev[3], ctx->stream));
  ctx->timed = true;
  ctx->last_R = dR;
  if ((rc = finish_out(ctx, R, n_dof, dR))) return rc;
  if ((rc = finish_out(ctx, J, nnz, dJ))) return rc;
  CU(cudaStreamSynchronize(ctx->stream));
  return MASTB200_OK;
}

int fill_stress_params(mastb200_context* ctx, const double* X, StressParams& p) {
  int rc = check_ready(ctx);
  if (rc) return rc;
  if (!X) return ctx->fail(MASTB200_ERR_ARG, "X is NULL");
  if (ctx->elem_type != MASTB200_HEX8 && ctx->elem_type != MASTB200_HEX27)
    return ctx->fail(MASTB200_ERR_UNSUPPORTED, "stress output: 3D solids (HEX8, HEX27) only");
  if (ctx->nv == 1) return ctx->fail(MASTB200_ERR_ARG, "stress output needs nv = 3 or 6");
  CU(cudaSetDevice(ctx->device));
  if ((rc = ensure_generic_plan(ctx))) return rc;
  mastb200_context::GenericPlan& G = ctx->gplan;
  const double* dX = nullptr;
  if ((rc = resolve_in(ctx, X, size_t(ctx->n_nodes) * ctx->nv, ctx->d_X, &dX))) return rc;
  if ((rc = localize_if_needed(ctx, dX, ctx->nv))) return rc;
  p = StressParams();
  p.n_elem = ctx->n_elem; p.n_owned = ctx->pat.n_owned;
  p.nn = ctx->nn; p.nq = ctx->nn; p.nv = ctx->nv; p.nonlinear = ctx->mat.nonlinear_strain ? 1 : 0;
  p.conn = G.conn.p; p.tab = G.tab.p; p.xyz = ctx->d_xyz.p; p.X = dX;
  p.rho = ctx->rho_set ? (ctx->rho_ext ? ctx->rho_ext : ctx->d_rho.p) : nullptr;
  p.diag_elem = ctx->has_diag ? G.diag.p : nullptr;
  p.cmask = ctx->has_constraints ? ctx->d_cmask.p : nullptr;
  p.owned_nodes = ctx->d_owned_nodes.p; p.n2e_ptr = G.n2e_ptr.p; p.n2e = G.n2e.p;
  const mastb200_material& m = ctx->mat;
  p.E0 = m.E0;
  p.lam1 = m.nu / (1. - m.nu - 2. * m.nu * m.nu);
  p.mu1 = 1. / 2. / (1. + m.nu);
  p.rho_min = m.rho_min; p.penalty = m.penalty;
  return MASTB200_OK;
}
int upload_elem_mask(mastb200_context* ctx, const uint8_t* elem_mask, const uint8_t** out) {
  *out = nullptr;
  if (!elem_mask) return MASTB200_OK;
  if (is_device_ptr(elem_mask)) { *out = elem_mask; return MASTB200_OK; }
  CU(ctx->d_elem_mask.ensure(ctx->n_elem));
  CU(cudaMemcpyAsync(ctx->d_elem_mask.p, elem_mask, ctx->n_elem, cudaMemcpyHostToDevice, ctx->stream));
  *out = ctx->d_elem_mask.p;
  return MASTB200_OK;
}

// density sensitivities alone: two element-parallel launches (stress_output.cu), HEX8 and HEX27, linear strain
int run_density_sens(mastb200_context* ctx, const double* X, const double* adj, double* G) {
  int rc;
  if (!ctx->rho_set) return ctx->fail(MASTB200_ERR_ARG, "density sensitivities need a nodal density field");
  if (ctx->thermal)
    return ctx->fail(MASTB200_ERR_UNSUPPORTED,
                     "thermal_residual_sensitivity is libmesh_error() for 3D elements in the reference");
  StressParams p;
  if ((rc = fill_stress_params(ctx, X, p))) return rc;
  const size_t n_dof = size_t(ctx->n_nodes) * ctx->nv;
  const double* dadj = nullptr;
  double* dG = nullptr;
  if (adj == X) dadj = p.X;  // compliance: the adjoint is the state
  else {
    if ((rc = resolve_in(ctx, adj, n_dof, ctx->d_adj, &dadj))) return rc;
    if ((rc = localize_if_needed(ctx, dadj, ctx->nv))) return rc;
  }
  if ((rc = resolve_out(ctx, G, (size_t)ctx->n_nodes, ctx->d_G, &dG))) return rc;
  CU(ctx->d_scratch.ensure(size_t(ctx->n_elem) * p.nq));
  p.scratch = ctx->d_scratch.p;
  CU(cudaEventRecord(ctx->ev[0], ctx->stream));
  if (ctx->has_ghosts) CU(cudaMemsetAsync(dG, 0, size_t(ctx->n_nodes) * sizeof(double), ctx->stream));
  CU(cudaEventRecord(ctx->ev[1], ctx->stream));
  // build_cube-numbered HEX8 meshes, 3-dof system, linear strain: the structured sensitivity sweep (node planes staged in
  // shared memory, nothing but G leaves the SM) instead of the two element-parallel kernels (MASTB200_SENS_STRUCT=0)
  bool sens_struct = ctx->elem_type == MASTB200_HEX8 && ctx->nv == 3 && !ctx->mat.nonlinear_strain;
  if (const char* e = getenv("MASTB200_SENS_STRUCT")) if (e[0] == '0') sens_struct = false;
  if (sens_struct) sens_struct = ensure_struct_plan(ctx) == MASTB200_OK && ctx->splan.ok;
  if (sens_struct) {
    const mastb200_context::StructPlan& S = ctx->splan;
    const mastb200_material& m = ctx->mat;
    StructParams sp;
    sp.units = S.s_units.p; sp.n_units = S.s_n_units; sp.counter = ctx->d_tile_counter.p;  // its own, larger tiles
    sp.nx = S.nx; sp.ny = S.ny; sp.nz = S.nz; sp.oz0 = S.oz0; sp.nv = 3;
    sp.xyz = ctx->d_xyz.p; sp.X = p.X; sp.adj = dadj; sp.G = dG;
    sp.cmask = ctx->has_constraints ? ctx->d_cmask.p : nullptr;
    sp.rho = ctx->rho_ext ? ctx->rho_ext : ctx->d_rho.p;
    sp.E0 = m.E0; sp.lam1 = m.nu / (1. - m.nu - 2. * m.nu * m.nu); sp.mu1 = 1. / 2. / (1. + m.nu);
    sp.rho_min = m.rho_min; sp.penalty = m.penalty;
    CU(launch_hex8_sens_struct(sp, ctx->sm_count, ctx->stream));
    ctx->launches += 1;
  } else {
    CU(launch_density_sensitivities(p, dadj, dG, ctx->stream));
    ctx->launches += 2;
  }
  CU(cudaEventRecord(ctx->ev[2], ctx->stream));
  CU(cudaEventRecord(ctx->ev[3], ctx->stream));
  ctx->timed = true;
  if ((rc = finish_out(ctx, G, (size_t)ctx->n_nodes, dG))) return rc;
  CU(cudaStreamSynchronize(ctx->stream));
  return MASTB200_OK;
}

// matrix-free element force of a 3D solid mesh (linear strain) into `out` [n_nodes*nv]: see launch_internal_force
int run_internal_force(mastb200_context* ctx, const double* V, int mode, const double* dfield, double lam1, double mu1,
                       double* out, bool constrain_rows = true, bool v_is_state = false) {
  int rc;
  StressParams p;
  if ((rc = fill_stress_params(ctx, V, p))) return rc;
  if (!constrain_rows) p.cmask = nullptr;  // ComplianceOutput::evaluate dots the unconstrained element vectors with the state
  // Green-Lagrange strain: the parameter sensitivities of the internal force (modes 1-3) carry F = I + grad u of the state;
  // the tangent product J dX (mode 0 with a perturbation) and the thermoelastic force vector need more than that
  if (ctx->mat.nonlinear_strain && ((mode == 0 && !v_is_state) || mode == 4))
    return ctx->fail(MASTB200_ERR_UNSUPPORTED, "tangent products with Green-Lagrange strain are not on the device path");
  p.lam1 = lam1; p.mu1 = mu1;
  p.alpha_k = ctx->mat.alpha * (3. * p.lam1 + 2. * p.mu1);
  p.T_const = ctx->T_const; p.T_ref = ctx->T_ref;
  p.temperature = (ctx->thermal && ctx->T_nodal) ? (ctx->T_ext ? ctx->T_ext : ctx->d_T.p) : nullptr;
  const size_t n_dof = size_t(ctx->n_nodes) * ctx->nv;
  double* d = nullptr;
  if ((rc = resolve_out(ctx, out, n_dof, ctx->d_dq, &d))) return rc;
  CU(ctx->d_scratch.ensure(size_t(ctx->n_elem) * p.nq * 9));
  p.scratch = ctx->d_scratch.p;
  p.dq_dX = d;
  if (ctx->has_ghosts) CU(cudaMemsetAsync(d, 0, n_dof * sizeof(double), ctx->stream));
  CU(launch_internal_force(p, mode, dfield, ctx->stream));
  ctx->launches += 2;
  if ((rc = finish_out(ctx, out, n_dof, d))) return rc;
  CU(cudaStreamSynchronize(ctx->stream));
  return MASTB200_OK;
}

// HEX8 with Green-Lagrange strain on an ijk-numbered (build_cube) mesh, 3-dof system, no thermal load: the tangent comes from
// the slot kernel's Green-Lagrange variant (sweep_hex8_slots.cu: 54 DMMA per element visit, the slots hold K itself), the
// residual f_a = sum_q JxW (F S) grad N_a from the matrix-free force kernels of the state (stress_output.cu, mode 0) plus the
// external loads.  MASTB200_GL_STRUCT=0 keeps the row-owner gather of sweep_hex8_gather.cu (any mesh, thermal, nv = 6).
bool gl_struct_ok(mastb200_context* ctx) {
  if (ctx->nv != 3 || ctx->thermal || kStructWarps != 16) return false;
  if (const char* e = getenv("MASTB200_GL_STRUCT")) if (e[0] == '0') return false;
  if (ensure_struct_plan(ctx) != MASTB200_OK || !ctx->splan.ok) return false;
  return (size_t)ctx->max_optin >= hex8_slots_smem_bytes();
}

int run_gl_struct(mastb200_context* ctx, const double* X, double* R, double* J) {
  int rc;
  const size_t n_dof = size_t(ctx->n_nodes) * ctx->nv;
  const size_t nnz = (size_t)plan_nnz(ctx);
  const double* dX = nullptr;
  double *dR = nullptr, *dJ = nullptr;
  if ((rc = resolve_in(ctx, X, n_dof, ctx->d_X, &dX))) return rc;
  if ((rc = localize_if_needed(ctx, dX, ctx->nv))) return rc;
  if ((rc = resolve_out(ctx, R, n_dof, ctx->d_R, &dR))) return rc;
  if ((rc = resolve_out(ctx, J, nnz, ctx->d_J, &dJ))) return rc;
  const mastb200_material& m = ctx->mat;
  const double lam1 = m.nu / (1. - m.nu - 2. * m.nu * m.nu), mu1 = 1. / 2. / (1. + m.nu);
  const uint8_t* cmask = ctx->has_constraints ? ctx->d_cmask.p : nullptr;
  CU(cudaEventRecord(ctx->ev[0], ctx->stream));
  CU(cudaEventRecord(ctx->ev[1], ctx->stream));
  if (dJ) {
    const mastb200_context::StructPlan& S = ctx->splan;
    StructParams sp;
    sp.units = S.units.p; sp.n_units = S.n_units; sp.counter = ctx->d_tile_counter.p;
    sp.nx = S.nx; sp.ny = S.ny; sp.nz = S.nz; sp.oz0 = S.oz0; sp.nv = 3;
    sp.nbr_ptr = S.nbr_ptr.p;
    sp.xyz = ctx->d_xyz.p; sp.X = dX; sp.cmask = cmask; sp.J = dJ; sp.R = nullptr;
    sp.rho = ctx->rho_set ? (ctx->rho_ext ? ctx->rho_ext : ctx->d_rho.p) : nullptr;
    sp.E0 = m.E0; sp.lam1 = lam1; sp.mu1 = mu1; sp.rho_min = m.rho_min; sp.penalty = m.penalty;
    sp.gl = 1;
    CU(launch_hex8_slots(sp, std::max(1, std::min(S.n_units, ctx->sm_count)), ctx->stream));
    ctx->launches++;
  }
  CU(cudaEventRecord(ctx->ev[2], ctx->stream));
  ctx->last_struct = true; ctx->splan.last_march = false; ctx->splan.last_slots = true;
  if (dR) {
    if ((rc = run_internal_force(ctx, dX, 0, nullptr, lam1, mu1, dR, true, true))) return rc;
    if ((rc = apply_external_loads(ctx, dR, cmask))) return rc;
  }
  CU(cudaEventRecord(ctx->ev[3], ctx->stream));
  ctx->timed = true;
  ctx->last_R = dR;
  if ((rc = finish_out(ctx, R, n_dof, dR))) return rc;
  if ((rc = finish_out(ctx, J, nnz, dJ))) return rc;
  CU(cudaStreamSynchronize(ctx->stream));
  return MASTB200_OK;
}

// the common sweep: any of R, J, G may be null
int run_sweep(mastb200_context* ctx, const double* X, const double* adj, double* G, double* R, double* J) {
  int rc = check_ready(ctx);
  if (rc) return rc;
  if (!X) return ctx->fail(MASTB200_ERR_ARG, "X is NULL");
  if (G && !adj) return ctx->fail(MASTB200_ERR_ARG, "adjoint vector is NULL");
  if (ctx->elem_type != MASTB200_HEX8) {
    if (G && ctx->elem_type != MASTB200_HEX27) return ctx->fail(MASTB200_ERR_UNSUPPORTED, "density sensitivities: 3D solids only");
    if ((R || J) && (rc = run_generic_sweep(ctx, X, R, J))) return rc;
    if (G) return run_density_sens(ctx, X, adj, G);
    return MASTB200_OK;
  }
  if (ctx->nv == 1) return ctx->fail(MASTB200_ERR_ARG, "HEX8 elasticity needs nv = 3 or 6");
  if (ctx->mat.nonlinear_strain) {  // Green-Lagrange strain: R / J by the row-owner gather, G by the element-parallel kernels
    if (G && !ctx->rho_set) return ctx->fail(MASTB200_ERR_ARG, "density sensitivities need a nodal density field");
    if ((R || J) && (rc = gl_struct_ok(ctx) ? run_gl_struct(ctx, X, R, J) : run_hex8_gather(ctx, X, R, J))) return rc;
    if (!G) return MASTB200_OK;
    const double* last_R = ctx->last_R;
    rc = run_density_sens(ctx, X, adj, G);
    if (R || J) ctx->last_R = last_R;
    return rc;
  }
  if (G && !ctx->rho_set) return ctx->fail(MASTB200_ERR_ARG, "density sensitivities need a nodal density field");
  {  // sensitivities alone: element-parallel kernels instead of the streaming sweep (MASTB200_SENS_SWEEP=1: old path)
    const char* e = getenv("MASTB200_SENS_SWEEP");
    if (G && !R && !J && !(e && e[0] == '1')) return run_density_sens(ctx, X, adj, G);
  }
  if (G && ctx->thermal)
    return ctx->fail(MASTB200_ERR_UNSUPPORTED,
                     "thermal_residual_sensitivity is libmesh_error() for 3D elements in the reference");
  const size_t n_dof = size_t(ctx->n_nodes) * ctx->nv;
  const size_t nnz = (size_t)plan_nnz(ctx);
  // the plain shared-memory layout serves residual/Jacobian sweeps of the 3-dof system with a homogeneous modulus and
  // no thermal load; everything else stages the auxiliary nodal fields and per-node extras
  int aux = (G || ctx->rho_set || ctx->thermal || ctx->nv == 6) ? 1 : 0;
  // build_cube-numbered meshes: the implicit-connectivity kernel serves the 3-dof R/J sweep (constant or SIMP modulus);
  // sensitivities requested together with R/J then come from the element-parallel kernels right after it
  // (the TEMPERATURE load of linear strain changes the residual only: the structured kernel still serves R / J and the load's
  // force vector, B^T D alpha dT from the matrix-free force kernels, is subtracted afterwards)
  const bool struct_ok_cfg = (ctx->nv == 3 || ctx->nv == 6) && !(ctx->thermal && G) && (R || J);
  if (struct_ok_cfg && (rc = ensure_struct_plan(ctx))) return rc;
  const bool use_struct = struct_ok_cfg && ctx->splan.ok;
  if (use_struct) aux = 0;
  // ... or, for the 3-dof system, from the same launch: the structured kernel's sensitivity variant
  const bool fused_G = use_struct && G && ctx->nv == 3 && hex8_struct_sens_smem_bytes() != 0 &&
                       (size_t)ctx->max_optin >= hex8_struct_sens_smem_bytes() && !(ctx->splan.march) &&
                       !(getenv("MASTB200_NO_FUSED_SENS") && getenv("MASTB200_NO_FUSED_SENS")[0] == '1');
  if (use_struct && G && !fused_G) {  // R/J by the implicit-connectivity sweep, then the sensitivities of the same state
    if ((rc = run_sweep(ctx, X, nullptr, nullptr, R, J))) return rc;
    float a = 0.f;
    if (cudaEventElapsedTime(&a, ctx->ev[0], ctx->ev[3]) != cudaSuccess) a = 0.f;
    const double* last_R = ctx->last_R;
    rc = run_density_sens(ctx, X, adj, G);
    ctx->carry_ms = a;
    ctx->last_R = last_R;
    return rc;
  }
  if (!use_struct && (rc = ensure_plan(ctx, aux))) return rc;
  ctx->last_plan = aux;
  ctx->last_struct = use_struct;
  SweepParams p;
  fill_params(ctx, aux, p);
  const double *dX = nullptr, *dadj = nullptr;
  double *dR = nullptr, *dJ = nullptr, *dG = nullptr;
  // Host X in and / or host R out on the implicit-connectivity path of one rank: the sweep runs z-segment by z-segment, the
  // planes of X a segment needs arrive on a second stream while the segment before is swept, and a segment's rows of R leave
  // while the next one is swept (MASTB200_SEG_COPY=0: one copy in, one launch, one copy out).
  const bool x_host = !is_device_ptr(X), r_host = R && !is_device_ptr(R);
  bool seg = use_struct && !G && (x_host || r_host) && ctx->neighbours.empty() && !ctx->has_ghosts &&
             ctx->splan.g_ptr.size() > 2 && ctx->splan.g_ptr.size() <= 33 && !(ctx->splan.march && ctx->nv == 3) &&
             !(ctx->thermal && R);
  if (const char* e = getenv("MASTB200_SEG_COPY")) if (e[0] == '0') seg = false;
  if (seg && x_host) {
    CU(ctx->d_X.ensure(n_dof));
    dX = ctx->d_X.p;
  } else if ((rc = resolve_in(ctx, X, n_dof, ctx->d_X, &dX))) return rc;
  if ((rc = localize_if_needed(ctx, dX, ctx->nv))) return rc;
  if (G && (rc = resolve_in(ctx, adj, n_dof, ctx->d_adj, &dadj))) return rc;
  if (G && adj != X && (rc = localize_if_needed(ctx, dadj, ctx->nv))) return rc;
  if ((rc = resolve_out(ctx, R, n_dof, ctx->d_R, &dR))) return rc;
  if ((rc = resolve_out(ctx, J, nnz, ctx->d_J, &dJ))) return rc;
  if ((rc = resolve_out(ctx, G, (size_t)ctx->n_nodes, ctx->d_G, &dG))) return rc;
  p.X = dX; p.adj = dadj; p.R = dR; p.J = dJ; p.G = dG;
  CU(cudaEventRecord(ctx->ev[0], ctx->stream));
  if (ctx->has_ghosts) {
    if (dR) CU(cudaMemsetAsync(dR, 0, n_dof * sizeof(double), ctx->stream));
    if (dG) CU(cudaMemsetAsync(dG, 0, size_t(ctx->n_nodes) * sizeof(double), ctx->stream));
  }
  CU(cudaEventRecord(ctx->ev[1], ctx->stream));
  if (use_struct) {
    const mastb200_context::StructPlan& S = ctx->splan;
    StructParams sp;
    sp.units = S.units.p; sp.n_units = S.n_units; sp.counter = ctx->d_tile_counter.p;
    sp.nx = S.nx; sp.ny = S.ny; sp.nz = S.nz; sp.oz0 = S.oz0; sp.nv = ctx->nv;
    sp.nbr_ptr = S.nbr_ptr.p; sp.arr_off = S.arr_off.p;
    sp.in1_off = kStructIn1Off; sp.cr0_off = kStructCr0Off; sp.cr1_off = kStructCr1Off;
    sp.zeros = getenv("MASTB200_STRUCT_NO_TMA_CLEAR") ? nullptr : S.zeros.p;
    sp.xyz = p.xyz; sp.X = dX; sp.cmask = p.cmask; sp.J = dJ; sp.R = dR;
    sp.j_aligned32 = (reinterpret_cast<uintptr_t>(dJ) & 31) == 0 ? 1 : 0;
    sp.rho = ctx->rho_set ? (ctx->rho_ext ? ctx->rho_ext : ctx->d_rho.p) : nullptr;
    if (fused_G) { sp.adj = adj == X ? dX : dadj; sp.G = dG; }
    sp.E0 = p.E0; sp.lam1 = p.lam1; sp.mu1 = p.mu1; sp.rho_min = p.rho_min; sp.penalty = p.penalty;
    ctx->splan.last_march = S.march && ctx->nv == 3;
    ctx->splan.last_slots = S.slots && ctx->nv == 3 && !fused_G && kStructWarps == 16;
    if (seg) {
      if (!ctx->seg_ready) {
        CU(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        for (auto& ev : ctx->seg_ev) CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        ctx->seg_ready = true;
      }
      const int ng = (int)S.g_ptr.size() - 1;
      const int64_t pn = (int64_t)(S.nx + 1) * (S.ny + 1) * ctx->nv;  // doubles per node plane
      CU(cudaEventRecord(ctx->seg_ev[64], ctx->stream));                // the copy stream starts behind what is queued so far
      CU(cudaStreamWaitEvent(ctx->copy_stream, ctx->seg_ev[64], 0));
      if (x_host) {
        int64_t up = 0;  // node planes of X on the device so far
        for (int g = 0; g < ng; ++g) {
          const int64_t hi = std::min<int64_t>((int64_t)S.g_z1[g] + 2, (int64_t)S.nz + 1);  // the segment reads planes < z1 + 2
          if (hi > up)
            CU(cudaMemcpyAsync(ctx->d_X.p + up * pn, X + up * pn, size_t(hi - up) * pn * sizeof(double), cudaMemcpyHostToDevice,
                               ctx->copy_stream));
          up = std::max(up, hi);
          CU(cudaEventRecord(ctx->seg_ev[g], ctx->copy_stream));
        }
      }
      for (int g = 0; g < ng; ++g) {
        if (x_host) CU(cudaStreamWaitEvent(ctx->stream, ctx->seg_ev[g], 0));
        sp.units = S.g_units.p + S.g_ptr[g]; sp.n_units = S.g_ptr[g + 1] - S.g_ptr[g];
        sp.counter = ctx->d_tile_counter.p + 1 + g;
        const int ctas = std::max(1, std::min(sp.n_units, ctx->sm_count * (ctx->splan.last_slots ? 1 : kStructCtasPerSm)));
        if (ctx->splan.last_slots) CU(launch_hex8_slots(sp, ctas, ctx->stream));
        else CU(launch_hex8_struct(sp, ctas, ctx->stream));
        ctx->launches++;
        if (dR) {  // external loads of the segment's nodes, then its rows of R may leave
          const int64_t pnn = (int64_t)(S.nx + 1) * (S.ny + 1);
          const int lo = (int)(S.g_z0[g] * pnn), hi = (int)(S.g_z1[g] * pnn);
          if (!ctx->fl_node.empty()) {
            FaceLoadPlan fp = face_plan(ctx);
            fp.node_lo = lo; fp.node_hi = hi;
            CU(launch_face_loads(fp, ctx->d_xyz.p, nullptr, p.cmask, ctx->nv, dR, nullptr, nullptr, nullptr, ctx->stream));
            ctx->launches++;
          }
          if (!ctx->pl_node.empty()) {
            CU(launch_point_loads((int)ctx->pl_node.size(), ctx->d_pl_node.p, ctx->d_pl_load.p, p.cmask, ctx->nv, dR, ctx->stream, lo, hi));
            ctx->launches++;
          }
          if (r_host) {
            CU(cudaEventRecord(ctx->seg_ev[32 + g], ctx->stream));
            CU(cudaStreamWaitEvent(ctx->copy_stream, ctx->seg_ev[32 + g], 0));
            CU(cudaMemcpyAsync(R + S.g_z0[g] * pn, dR + S.g_z0[g] * pn, size_t(S.g_z1[g] - S.g_z0[g]) * pn * sizeof(double),
                               cudaMemcpyDeviceToHost, ctx->copy_stream));
          }
        }
      }
      CU(cudaEventRecord(ctx->seg_ev[65], ctx->copy_stream));  // the context stream is complete when the copies are
      CU(cudaStreamWaitEvent(ctx->stream, ctx->seg_ev[65], 0));
      ctx->launches--;  // the common code below counts one launch
    } else if (ctx->splan.last_march) {
      sp.units = S.m_units.p; sp.n_units = S.m_n_units;
      CU(launch_hex8_march(sp, S.m_n_ctas, ctx->stream));
    } else if (ctx->splan.last_slots) {
      CU(launch_hex8_slots(sp, std::max(1, std::min(S.n_units, ctx->sm_count)), ctx->stream));
    } else {
      CU(launch_hex8_struct(sp, S.n_ctas, ctx->stream));
    }
  } else {
    CU(launch_hex8_sweep(p, ctx->dplan[aux].n_ctas, ctx->stream));
  }
  ctx->launches++;
  CU(cudaEventRecord(ctx->ev[2], ctx->stream));
  if (!seg && (rc = apply_external_loads(ctx, dR, p.cmask))) return rc;
  if (use_struct && ctx->thermal && dR) {  // f -= JxW B^T (D alpha dT) (solid_element_3d.cpp:950-1089, linear strain)
    CU(ctx->d_vec2.ensure(n_dof));
    if ((rc = run_internal_force(ctx, dX, 4, nullptr, p.lam1, p.mu1, ctx->d_vec2.p))) return rc;
    CU(launch_sub_vector((int64_t)n_dof, dR, ctx->d_vec2.p, ctx->stream));
    ctx->launches++;
  }
  CU(cudaEventRecord(ctx->ev[3], ctx->stream));
  ctx->timed = true;
  ctx->last_R = dR;
  if (!(seg && r_host) && (rc = finish_out(ctx, R, n_dof, dR))) return rc;
  if ((rc = finish_out(ctx, J, nnz, dJ))) return rc;
  if ((rc = finish_out(ctx, G, (size_t)ctx->n_nodes, dG))) return rc;
  // stream-ordered mode: nothing left the device, so the caller's next call (or its own stream work) may queue behind this
  // sweep; errors of the launch itself were collected above, asynchronous ones surface at the next synchronising call
  const bool all_device = dX == X && (!R || dR == R) && (!J || dJ == J) && (!G || (dG == G && (adj == X || dadj == adj)));
  if (!(ctx->async_device && all_device)) CU(cudaStreamSynchronize(ctx->stream));
  return MASTB200_OK;
}


extern "C" {

static int rebuild_face_loads(mastb200_context* ctx);

const char* mastb200_version(void) { return "mast_b200 0.1 (sm_100a)"; }

int mastb200_create(int device, mastb200_context** out) {
  if (!out) return MASTB200_ERR_ARG;
  *out = nullptr;
  if (device == -1) {  // host planning context: sparsity pattern and assembly plan only
    mastb200_context* ctx = new mastb200_context();
    ctx->device = -1;
    ctx->planning_only = true;
    ctx->sm_count = 148;
    *out = ctx;
    return MASTB200_OK;
  }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    g_create_error = std::string("no CUDA device: ") + cudaGetErrorString(e) + " (there is no CPU fallback)";
    cudaGetLastError();
    return MASTB200_ERR_CUDA;
  }
  if (device < 0 || device >= ndev) { g_create_error = "device index out of range"; return MASTB200_ERR_ARG; }
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) { g_create_error = cudaGetErrorString(e); return MASTB200_ERR_CUDA; }
  if (prop.major != 10) {
    g_create_error = "device is not sm_100 (Blackwell B200); this library carries sm_100a code only";
    return MASTB200_ERR_CUDA;
  }
  if ((e = cudaSetDevice(device)) != cudaSuccess) { g_create_error = cudaGetErrorString(e); return MASTB200_ERR_CUDA; }
  mastb200_context* ctx = new mastb200_context();
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) {
    g_create_error = cudaGetErrorString(e);
    delete ctx;
    return MASTB200_ERR_CUDA;
  }
  for (auto& ev : ctx->ev) cudaEventCreate(&ev);
  *out = ctx;
  return MASTB200_OK;
}

int mastb200_destroy(mastb200_context* ctx) {
  if (!ctx) return MASTB200_OK;
  if (ctx->planning_only) { delete ctx; return MASTB200_OK; }
  cudaSetDevice(ctx->device);
  if (ctx->stream) { cudaStreamSynchronize(ctx->stream); cudaStreamDestroy(ctx->stream); }
  for (auto& ev : ctx->ev) if (ev) cudaEventDestroy(ev);
  for (auto& ev : ctx->seg_ev) if (ev) cudaEventDestroy(ev);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  comm_release(ctx);
  delete ctx;
  return MASTB200_OK;
}

const char* mastb200_last_error(const mastb200_context* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

void* mastb200_stream(const mastb200_context* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

int mastb200_set_async(mastb200_context* ctx, int32_t on) {
  if (!ctx) return MASTB200_ERR_ARG;
  ctx->async_device = on != 0;
  return MASTB200_OK;
}

int mastb200_set_mesh(mastb200_context* ctx, int32_t elem_type, int64_t n_elem, int64_t n_nodes, const double* xyz,
                      const int32_t* conn, const uint8_t* diag_elem) {
  if (!ctx) return MASTB200_ERR_ARG;
  int nn = 0;
  switch (elem_type) {
    case MASTB200_HEX8: nn = 8; break;
    case MASTB200_HEX27: nn = 27; break;
    case MASTB200_QUAD4: nn = 4; break;
    default: return ctx->fail(MASTB200_ERR_ARG, "unknown element type %d", elem_type);
  }
  if (n_elem <= 0 || n_nodes <= 0 || !xyz || !conn) return ctx->fail(MASTB200_ERR_ARG, "empty mesh or NULL arrays");
  if (!ctx->planning_only) CU(cudaSetDevice(ctx->device));
  ctx->elem_type = elem_type; ctx->nn = nn; ctx->n_elem = n_elem; ctx->n_nodes = n_nodes;
  ctx->xyz.assign(xyz, xyz + n_nodes * 3);
  ctx->conn.assign(conn, conn + n_elem * nn);
  ctx->has_diag = false;
  if (diag_elem) {
    ctx->diag_elem.assign(diag_elem, diag_elem + n_elem);
    for (uint8_t d : ctx->diag_elem) ctx->has_diag = ctx->has_diag || d;
  } else {
    ctx->diag_elem.clear();
  }
  if (!ctx->planning_only) CU(ctx->d_xyz.upload(ctx->xyz, ctx->stream));
  ctx->have_mesh = true; ctx->have_plan = false;
  // everything sized for the previous mesh goes: the caller sets the dof map, density, temperature and loads again
  ctx->have_dofs = false; ctx->owned.clear(); ctx->cmask.clear(); ctx->has_constraints = false; ctx->has_ghosts = false;
  ctx->rho_set = false; ctx->rho_ext = nullptr; ctx->thermal = false; ctx->T_nodal = false; ctx->T_ext = nullptr;
  ctx->last_R = nullptr;
  ctx->have_filter = false; ctx->have_nodal_vol = false;
  ctx->sp_elem.clear(); ctx->sp_side.clear(); ctx->sp_p.clear();
  ctx->fl_node.clear(); ctx->fl_ptr.clear(); ctx->fl_face_nodes.clear(); ctx->fl_k.clear(); ctx->fl_p.clear();
  ctx->have_heat_sides = false;
  return MASTB200_OK;
}

int mastb200_set_dof_map(mastb200_context* ctx, int32_t nv, const uint8_t* constrained, const uint8_t* owned) {
  if (!ctx) return MASTB200_ERR_ARG;
  if (!ctx->have_mesh) return ctx->fail(MASTB200_ERR_ARG, "set_mesh first");
  if (nv != 1 && nv != 3 && nv != 6) return ctx->fail(MASTB200_ERR_ARG, "nv must be 1, 3 or 6");
  if (!ctx->planning_only) CU(cudaSetDevice(ctx->device));
  ctx->nv = nv;
  ctx->cmask.assign(ctx->n_nodes, 0);
  ctx->has_constraints = false;
  if (constrained)
    for (int64_t n = 0; n < ctx->n_nodes; ++n) {
      unsigned m = 0;
      for (int v = 0; v < nv; ++v)
        if (constrained[n * nv + v]) m |= 1u << v;
      ctx->cmask[n] = (uint8_t)m;
      ctx->has_constraints = ctx->has_constraints || m;
    }
  ctx->has_ghosts = false;
  if (owned) {
    ctx->owned.assign(owned, owned + ctx->n_nodes);
    for (uint8_t o : ctx->owned) ctx->has_ghosts = ctx->has_ghosts || !o;
  } else {
    ctx->owned.clear();
  }
  if (!ctx->planning_only) CU(ctx->d_cmask.upload(ctx->cmask, ctx->stream));
  ctx->have_dofs = true; ctx->have_plan = false;
  if (!ctx->planning_only && !ctx->sp_elem.empty()) return rebuild_face_loads(ctx);  // the owned flags select the loaded nodes
  return MASTB200_OK;
}

int mastb200_set_material(mastb200_context* ctx, const mastb200_material* mat) {
  if (!ctx || !mat) return MASTB200_ERR_ARG;
  ctx->mat = *mat;
  ctx->have_mat = true;
  return MASTB200_OK;
}

// Builds the per-node gather lists of the SURFACE_PRESSURE loads from the raw (element, side, pressure) entries.  Only
// OWNED nodes get a list: ghost rows of R stay zero and the compliance sum of a rank covers its own nodes, so the
// caller's sum over ranks counts every loaded node once (ghost elements carry the load of nodes on a partition cut).
static int rebuild_face_loads(mastb200_context* ctx) {
  ctx->fl_node.clear(); ctx->fl_ptr.clear(); ctx->fl_face_nodes.clear(); ctx->fl_k.clear(); ctx->fl_p.clear();
  const int64_t n = (int64_t)ctx->sp_elem.size();
  if (n == 0) return MASTB200_OK;
  const int nfn = ctx->elem_type == MASTB200_HEX8 ? 4 : 9, nne = ctx->nn;
  const int32_t* elem = ctx->sp_elem.data();
  const int32_t* side = ctx->sp_side.data();
  // entries ordered by (element, side): the order in which the reference's element loop meets the loads
  std::vector<int64_t> order(n);
  for (int64_t i = 0; i < n; ++i) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) {
    return elem[a] != elem[b] ? elem[a] < elem[b] : side[a] < side[b];
  });
  struct Ent { int32_t node; int64_t seq; int32_t fn[9]; uint8_t k; double p; };
  std::vector<Ent> ents;
  ents.reserve(n * nfn);
  for (int64_t s = 0; s < n; ++s) {
    const int64_t i = order[s];
    Ent e;
    for (int k = 0; k < 9; ++k) e.fn[k] = 0;
    for (int k = 0; k < nfn; ++k)
      e.fn[k] = ctx->conn[(int64_t)elem[i] * nne + (nfn == 4 ? H8_SIDE[side[i]][k] : H27_SIDE[side[i]][k])];
    e.p = ctx->sp_p[i];
    for (int k = 0; k < nfn; ++k) {
      e.node = e.fn[k]; e.k = (uint8_t)k; e.seq = s;
      if (ctx->owned.empty() || ctx->owned[e.node]) ents.push_back(e);
    }
  }
  std::stable_sort(ents.begin(), ents.end(), [](const Ent& a, const Ent& b) { return a.node < b.node; });
  for (size_t i = 0; i < ents.size(); ++i) {
    if (i == 0 || ents[i].node != ents[i - 1].node) {
      ctx->fl_node.push_back(ents[i].node);
      ctx->fl_ptr.push_back((int32_t)i);
    }
    for (int k = 0; k < nfn; ++k) ctx->fl_face_nodes.push_back(ents[i].fn[k]);
    ctx->fl_k.push_back(ents[i].k);
    ctx->fl_p.push_back(ents[i].p);
  }
  ctx->fl_ptr.push_back((int32_t)ents.size());
  if (ents.empty()) { ctx->fl_ptr.clear(); return MASTB200_OK; }
  CU(ctx->d_fl_node.upload(ctx->fl_node, ctx->stream));
  CU(ctx->d_fl_ptr.upload(ctx->fl_ptr, ctx->stream));
  CU(ctx->d_fl_face_nodes.upload(ctx->fl_face_nodes, ctx->stream));
  CU(ctx->d_fl_k.upload(ctx->fl_k, ctx->stream));
  CU(ctx->d_fl_p.upload(ctx->fl_p, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return MASTB200_OK;
}

int mastb200_set_side_pressure(mastb200_context* ctx, int64_t n, const int32_t* elem, const int32_t* side,
                               const double* pressure) {
  if (!ctx) return MASTB200_ERR_ARG;
  if (!ctx->have_mesh) return ctx->fail(MASTB200_ERR_ARG, "set_mesh first");
  ctx->sp_elem.clear(); ctx->sp_side.clear(); ctx->sp_p.clear();
  ctx->fl_node.clear(); ctx->fl_ptr.clear(); ctx->fl_face_nodes.clear(); ctx->fl_k.clear(); ctx->fl_p.clear();
  if (n == 0) return MASTB200_OK;
  if (ctx->elem_type != MASTB200_HEX8 && ctx->elem_type != MASTB200_HEX27)
    return ctx->fail(MASTB200_ERR_UNSUPPORTED, "side pressure: 3D solids (HEX8, HEX27) only");
  if (ctx->planning_only) return ctx->fail(MASTB200_ERR_CUDA, "planning-only context");
  CU(cudaSetDevice(ctx->device));
  for (int64_t i = 0; i < n; ++i)
    if (elem[i] < 0 || elem[i] >= ctx->n_elem || side[i] < 0 || side[i] > 5) return ctx->fail(MASTB200_ERR_ARG, "bad side load %lld", (long long)i);
  ctx->sp_elem.assign(elem, elem + n); ctx->sp_side.assign(side, side + n); ctx->sp_p.assign(pressure, pressure + n);
  return rebuild_face_loads(ctx);
}

int mastb200_set_heat_side_loads(mastb200_context* ctx, int64_t n, const int32_t* elem, const int32_t* side,
                                 const double* heat_flux, const double* conv_coeff, const double* T_amb) {
  return mastb200_set_heat_side_loads_ex(ctx, n, elem, side, heat_flux, conv_coeff, T_amb, nullptr, 0., 0.);
}

int mastb200_set_heat_side_loads_ex(mastb200_context* ctx, int64_t n, const int32_t* elem, const int32_t* side,
                                    const double* heat_flux, const double* conv_coeff, const double* T_amb,
                                    const double* emissivity, double sb_const, double T_ref_zero) {
  if (!ctx) return MASTB200_ERR_ARG;
  if (!ctx->have_mesh) return ctx->fail(MASTB200_ERR_ARG, "set_mesh first");
  ctx->have_heat_sides = false;
  if (n == 0) return MASTB200_OK;
  if (ctx->elem_type != MASTB200_QUAD4) return ctx->fail(MASTB200_ERR_UNSUPPORTED, "heat side loads: QUAD4 heat conduction only");
  if (ctx->planning_only) return ctx->fail(MASTB200_ERR_CUDA, "planning-only context");
  if (!elem || !side) return ctx->fail(MASTB200_ERR_ARG, "NULL side load arrays");
  CU(cudaSetDevice(ctx->device));
  std::vector<int64_t> order(n);
  for (int64_t i = 0; i < n; ++i) {
    if (elem[i] < 0 || elem[i] >= ctx->n_elem || side[i] < 0 || side[i] > 3) return ctx->fail(MASTB200_ERR_ARG, "bad side load %lld", (long long)i);
    order[i] = i;
  }
  // per element in ascending side id: the order in which the reference's element loop meets them
  std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) {
    return elem[a] != elem[b] ? elem[a] < elem[b] : side[a] < side[b];
  });
  std::vector<int32_t> ptr(ctx->n_elem + 1, 0), sd(n);
  std::vector<double> val(n * 4);
  for (int64_t i = 0; i < n; ++i) ptr[elem[i] + 1]++;
  for (int64_t e = 0; e < ctx->n_elem; ++e) ptr[e + 1] += ptr[e];
  for (int64_t k = 0; k < n; ++k) {
    const int64_t i = order[k];
    sd[k] = side[i];
    val[k * 4] = heat_flux ? heat_flux[i] : 0.;
    val[k * 4 + 1] = conv_coeff ? conv_coeff[i] : 0.;
    val[k * 4 + 2] = T_amb ? T_amb[i] : 0.;
    val[k * 4 + 3] = emissivity ? emissivity[i] : 0.;
  }
  ctx->rad_sbc = sb_const; ctx->rad_T0 = T_ref_zero;
  CU(ctx->d_hs_ptr.upload(ptr, ctx->stream));
  CU(ctx->d_hs_side.upload(sd, ctx->stream));
  CU(ctx->d_hs_val.upload(val, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  ctx->have_heat_sides = true;
  return MASTB200_OK;
}

int mastb200_set_temperature(mastb200_context* ctx, int32_t enable, const double* T_nodal, double T_const, double T_ref) {
  if (!ctx) return MASTB200_ERR_ARG;
  if (!ctx->have_mesh) return ctx->fail(MASTB200_ERR_ARG, "set_mesh first");
  if (ctx->planning_only) return ctx->fail(MASTB200_ERR_CUDA, "planning-only context");
  CU(cudaSetDevice(ctx->device));
  ctx->thermal = enable != 0;
  ctx->T_const = T_const; ctx->T_ref = T_ref;
  ctx->T_nodal = false; ctx->T_ext = nullptr;
  if (enable && T_nodal) {
    ctx->T_nodal = true;
    if (is_device_ptr(T_nodal)) ctx->T_ext = T_nodal;
    else {
      CU(ctx->d_T.ensure(ctx->n_nodes));
      CU(cudaMemcpyAsync(ctx->d_T.p, T_nodal, ctx->n_nodes * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
      CU(cudaStreamSynchronize(ctx->stream));
    }
  }
  return MASTB200_OK;
}

int mastb200_set_density(mastb200_context* ctx, const double* rho) {
  if (!ctx) return MASTB200_ERR_ARG;
  if (!ctx->have_mesh) return ctx->fail(MASTB200_ERR_ARG, "set_mesh first");
  if (ctx->planning_only) return ctx->fail(MASTB200_ERR_CUDA, "planning-only context");
  CU(cudaSetDevice(ctx->device));
  ctx->rho_set = rho != nullptr; ctx->rho_ext = nullptr;
  if (!rho) return MASTB200_OK;
  if (is_device_ptr(rho)) ctx->rho_ext = rho;
  else {
    CU(ctx->d_rho.ensure(ctx->n_nodes));
    CU(cudaMemcpyAsync(ctx->d_rho.p, rho, ctx->n_nodes * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
  }
  return MASTB200_OK;
}

int mastb200_build_pattern(mastb200_context* ctx, int64_t* n_rows, int64_t* nnz) {
  if (!ctx) return MASTB200_ERR_ARG;
  if (!ctx->have_mesh || !ctx->have_dofs) return ctx->fail(MASTB200_ERR_ARG, "set_mesh and set_dof_map first");
  if (!ctx->planning_only) {
    CU(cudaSetDevice(ctx->device));
    CU(cudaDeviceGetAttribute(&ctx->max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device));
  }
  for (auto& D : ctx->dplan) D.reset();
  ctx->gplan.reset();
  ctx->splan.reset();
  ctx->q4s_tried = false; ctx->q4s_ok = false;
  const std::string e = build_pattern(ctx->nn, ctx->n_nodes, ctx->n_elem, ctx->conn.data(),
                                      ctx->owned.empty() ? nullptr : ctx->owned.data(),
                                      ctx->diag_elem.empty() ? nullptr : ctx->diag_elem.data(), ctx->pat);
  if (!e.empty()) return ctx->fail(MASTB200_ERR_PLAN, "pattern: %s", e.c_str());
  ctx->have_plan = true;
  // the plan of the layout this problem will most likely sweep with is built now; the other one on first use
  const int aux = (ctx->rho_set || ctx->thermal || ctx->nv == 6) ? 1 : 0;
  const bool struct_cfg = ctx->elem_type == MASTB200_HEX8 && (ctx->nv == 3 || ctx->nv == 6) && !ctx->thermal;
  if (struct_cfg) {
    if (int rc = ensure_struct_plan(ctx)) { ctx->have_plan = false; return rc; }
  }
  if (!(struct_cfg && ctx->splan.ok)) {  // otherwise the streaming plan is built on first use, if ever
    if (int rc = ensure_plan(ctx, aux)) { ctx->have_plan = false; return rc; }
  }
  ctx->last_plan = aux;
  if (!ctx->planning_only) {
    CU(ctx->d_owned_nodes.upload(ctx->pat.owned_nodes, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
  }
  if (n_rows) *n_rows = ctx->pat.n_owned * ctx->nv;
  if (nnz) *nnz = plan_nnz(ctx);
  return MASTB200_OK;
}

int mastb200_get_pattern(const mastb200_context* ctx, int64_t* row_ptr, int32_t* col_idx) {
  if (!ctx || !ctx->have_plan || !row_ptr) return MASTB200_ERR_ARG;
  const Pattern& P = ctx->pat;
  const int nv = ctx->nv;
  int64_t k = 0;
  row_ptr[0] = 0;
  for (int64_t r = 0; r < P.n_owned; ++r) {
    const int64_t b = P.nbr_ptr[r], e = P.nbr_ptr[r + 1];
    for (int v = 0; v < nv; ++v) {
      if (col_idx)
        for (int64_t j = b; j < e; ++j)
          for (int w = 0; w < nv; ++w) col_idx[k++] = P.nbr[j] * nv + w;
      else
        k += (e - b) * nv;
      row_ptr[r * nv + v + 1] = k;
    }
  }
  return MASTB200_OK;
}

int mastb200_get_pattern_rows(const mastb200_context* ctx, int64_t n, const int64_t* rows, int64_t* row_begin,
                              int32_t* row_len, int32_t* col_idx) {
  if (!ctx || !ctx->have_plan || n < 0 || (n > 0 && (!rows || !row_begin || !row_len))) return MASTB200_ERR_ARG;
  const Pattern& P = ctx->pat;
  const int nv = ctx->nv;
  int64_t k = 0;
  for (int64_t i = 0; i < n; ++i) {
    const int64_t r = rows[i];
    if (r < 0 || r >= P.n_owned * nv) return MASTB200_ERR_ARG;
    const int64_t node = r / nv, v = r - node * nv;
    const int64_t b = P.nbr_ptr[node], e = P.nbr_ptr[node + 1];
    row_begin[i] = b * nv * nv + v * (e - b) * nv;  // the nv rows of a node are contiguous, (e - b) * nv values each
    row_len[i] = (int32_t)((e - b) * nv);
    if (col_idx)
      for (int64_t j = b; j < e; ++j)
        for (int w = 0; w < nv; ++w) col_idx[k++] = P.nbr[j] * nv + w;
  }
  return MASTB200_OK;
}

int mastb200_get_row_dofs(const mastb200_context* ctx, int64_t* row_dof) {
  if (!ctx || !ctx->have_plan || !row_dof) return MASTB200_ERR_ARG;
  for (int64_t r = 0; r < ctx->pat.n_owned; ++r)
    for (int v = 0; v < ctx->nv; ++v) row_dof[r * ctx->nv + v] = (int64_t)ctx->pat.owned_nodes[r] * ctx->nv + v;
  return MASTB200_OK;
}

int mastb200_residual_and_jacobian(mastb200_context* ctx, const double* X, double* R, double* J_vals) {
  if (!ctx) return MASTB200_ERR_ARG;
  if (int rc0 = check_ready(ctx)) return rc0;
  CU(cudaSetDevice(ctx->device));
  ctx->carry_ms = 0.;
  return run_sweep(ctx, X, nullptr, nullptr, R, J_vals);
}

int mastb200_sensitivity_sweep(mastb200_context* ctx, const double* X, const double* adj, double* G, double* R,
                               double* J_vals) {
  if (!ctx) return MASTB200_ERR_ARG;
  if (!G) return ctx->fail(MASTB200_ERR_ARG, "G is NULL");
  if (int rc0 = check_ready(ctx)) return rc0;
  CU(cudaSetDevice(ctx->device));
  ctx->carry_ms = 0.;
  return run_sweep(ctx, X, adj, G, R, J_vals);
}

int mastb200_res_l2_norm(mastb200_context* ctx, double* norm) {
  if (!ctx || !norm) return MASTB200_ERR_ARG;
  if (int rc0 = check_ready(ctx)) return rc0;
  if (!ctx->last_R) return ctx->fail(MASTB200_ERR_ARG, "no residual has been assembled");
  CU(cudaSetDevice(ctx->device));
  const int nb = 1024;
  CU(ctx->d_partial.ensure(4096));
  CU(ctx->d_scalar.ensure(8));
  CU(launch_sumsq_owned(ctx->last_R, ctx->d_owned_nodes.p, ctx->pat.n_owned, ctx->nv, ctx->d_partial.p, nb, ctx->stream));
  CU(launch_reduce_partials(ctx->d_partial.p, nb, ctx->d_scalar.p, ctx->stream));
  ctx->launches += 2;
  double ss = 0.;
  CU(cudaMemcpyAsync(&ss, ctx->d_scalar.p, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  *norm = std::sqrt(ss);
  return MASTB200_OK;
}

int mastb200_apply_dv_map_transpose(mastb200_context* ctx, const double* G, int64_t n_dv, const int64_t* dv_ptr,
                                    const int32_t* dv_node, const double* dv_weight, double* sens) {
  if (!ctx || !G || !dv_ptr || !sens || n_dv < 0) return MASTB200_ERR_ARG;
  if (int rc0 = check_ready(ctx)) return rc0;
  CU(cudaSetDevice(ctx->device));
  if (n_dv == 0) return MASTB200_OK;
  const int64_t nnzw = dv_ptr[n_dv];
  const double* dG = nullptr;
  int rc = resolve_in(ctx, G, (size_t)ctx->n_nodes, ctx->d_G, &dG);
  if (rc) return rc;
  CU(ctx->d_dv_ptr.ensure(n_dv + 1));
  CU(ctx->d_dv_node.ensure(nnzw));
  CU(ctx->d_dv_w.ensure(nnzw));
  CU(ctx->d_sens.ensure(n_dv));
  CU(cudaMemcpyAsync(ctx->d_dv_ptr.p, dv_ptr, (n_dv + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(ctx->d_dv_node.p, dv_node, nnzw * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(ctx->d_dv_w.p, dv_weight, nnzw * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  CU(launch_dv_map_transpose(dG, n_dv, ctx->d_dv_ptr.p, ctx->d_dv_node.p, ctx->d_dv_w.p, ctx->d_sens.p, ctx->stream));
  ctx->launches++;
  CU(cudaMemcpyAsync(sens, ctx->d_sens.p, n_dv * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return MASTB200_OK;
}

int mastb200_set_point_loads(mastb200_context* ctx, int64_t n, const int32_t* node, const double* load) {
  if (!ctx) return MASTB200_ERR_ARG;
  if (!ctx->have_mesh || !ctx->have_dofs) return ctx->fail(MASTB200_ERR_ARG, "set_mesh and set_dof_map first");
  if (ctx->planning_only) return ctx->fail(MASTB200_ERR_CUDA, "planning-only context");
  ctx->pl_node.clear(); ctx->pl_load.clear();
  if (n == 0) return MASTB200_OK;
  if (!node || !load) return MASTB200_ERR_ARG;
  CU(cudaSetDevice(ctx->device));
  const int nv = ctx->nv;
  std::vector<int64_t> order(n);
  for (int64_t i = 0; i < n; ++i) {
    if (node[i] < 0 || node[i] >= ctx->n_nodes) return ctx->fail(MASTB200_ERR_ARG, "bad point load %lld", (long long)i);
    order[i] = i;
  }
  std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return node[a] < node[b]; });
  for (int64_t s = 0; s < n; ++s) {
    const int64_t i = order[s];
    if (!ctx->owned.empty() && !ctx->owned[node[i]]) continue;  // the reference zeroes components of other ranks' dofs
    if (ctx->pl_node.empty() || ctx->pl_node.back() != node[i]) {
      ctx->pl_node.push_back(node[i]);
      ctx->pl_load.insert(ctx->pl_load.end(), nv, 0.);
    }
    for (int v = 0; v < nv; ++v) ctx->pl_load[ctx->pl_load.size() - nv + v] += load[i * nv + v];
  }
  if (ctx->pl_node.empty()) return MASTB200_OK;
  CU(ctx->d_pl_node.upload(ctx->pl_node, ctx->stream));
  CU(ctx->d_pl_load.upload(ctx->pl_load, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return MASTB200_OK;
}

int mastb200_sensitivity_assemble(mastb200_context* ctx, const double* X, int32_t param, int64_t n_k, const int32_t* dv_node,
                                  const double* dv_weight, double* dRdp) {
  if (!ctx || !X || !dRdp) return MASTB200_ERR_ARG;
  int rc = check_ready(ctx);
  if (rc) return rc;
  if (ctx->elem_type == MASTB200_QUAD4 && ctx->nv == 1) {
    // Heat conduction (HeatConductionNonlinearAssemblyElemOperations::elem_sensitivity_calculations,
    // heat_conduction_nonlinear_assembly.cpp:100-116).  Every term of the element residual is linear in the conductivity k0
    // or in the section thickness h (d(k h) = dk h + k dh, solid_2d_section_element_property_card.cpp:955-975; sources and
    // side loads are value*dth + dvalue*th with dvalue = 0), so dR/dp is the residual sweep itself with the card swapped:
    //   dR/dk0 = internal residual at k0 = 1 (no source, no side loads),   dR/dh = full residual at h = 1.
    if (param != MASTB200_PARAM_K0 && param != MASTB200_PARAM_THICKNESS)
      return ctx->fail(MASTB200_ERR_ARG, "heat conduction: parameter kinds MASTB200_PARAM_K0 / MASTB200_PARAM_THICKNESS");
    const mastb200_material saved = ctx->mat;
    const bool saved_sides = ctx->have_heat_sides;
    if (param == MASTB200_PARAM_K0) { ctx->mat.k0 = 1.; ctx->mat.heat_source = 0.; ctx->have_heat_sides = false; }
    else ctx->mat.thickness = 1.;
    rc = run_generic_sweep(ctx, X, dRdp, nullptr);
    ctx->mat = saved;
    ctx->have_heat_sides = saved_sides;
    return rc;
  }
  if (ctx->elem_type != MASTB200_HEX8 && ctx->elem_type != MASTB200_HEX27)
    return ctx->fail(MASTB200_ERR_UNSUPPORTED, "sensitivity_assemble: 3D solids (HEX8, HEX27) and QUAD4 heat conduction only");
  if (ctx->thermal)
    return ctx->fail(MASTB200_ERR_UNSUPPORTED, "thermal_residual_sensitivity is libmesh_error() for 3D elements in the reference");
  CU(cudaSetDevice(ctx->device));
  const double nu = ctx->mat.nu;
  const double lam1 = nu / (1. - nu - 2. * nu * nu), mu1 = 1. / 2. / (1. + nu);
  if (param == MASTB200_PARAM_DENSITY_DV) {
    if (!ctx->rho_set) return ctx->fail(MASTB200_ERR_ARG, "a density design variable needs a nodal density field");
    if (n_k < 0 || (n_k > 0 && (!dv_node || !dv_weight))) return MASTB200_ERR_ARG;
    for (int64_t k = 0; k < n_k; ++k)
      if (dv_node[k] < 0 || dv_node[k] >= ctx->n_nodes) return ctx->fail(MASTB200_ERR_ARG, "bad design-variable node");
    // nodal derivative field d(rho_n)/dp: the filter column of the design variable, zero outside its support
    CU(ctx->d_dfield.ensure(ctx->n_nodes));
    CU(cudaMemsetAsync(ctx->d_dfield.p, 0, size_t(ctx->n_nodes) * sizeof(double), ctx->stream));
    if (n_k > 0) {
      CU(ctx->d_dvn.ensure(n_k)); CU(ctx->d_dvw.ensure(n_k));
      CU(cudaMemcpyAsync(ctx->d_dvn.p, dv_node, n_k * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
      CU(cudaMemcpyAsync(ctx->d_dvw.p, dv_weight, n_k * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
      CU(launch_scatter_field(n_k, ctx->d_dvn.p, ctx->d_dvw.p, ctx->d_dfield.p, ctx->stream));
      ctx->launches++;
    }
    return run_internal_force(ctx, X, 1, ctx->d_dfield.p, lam1, mu1, dRdp);
  }
  if (param == MASTB200_PARAM_E) return run_internal_force(ctx, X, 2, nullptr, lam1, mu1, dRdp);
  if (param == MASTB200_PARAM_NU) {
    // d/dnu of the unit-modulus Lame constants of isotropic_material_property_card.cpp:512-528 (derivative :530-567)
    const double den = 1. - nu - 2. * nu * nu;
    const double dlam = (den + nu * (1. + 4. * nu)) / (den * den), dmu = -1. / 2. / ((1. + nu) * (1. + nu));
    return run_internal_force(ctx, X, 3, nullptr, dlam, dmu, dRdp);
  }
  return ctx->fail(MASTB200_ERR_ARG, "unknown parameter kind %d", (int)param);
}

int mastb200_linearized_jacobian_solution_product(mastb200_context* ctx, const double* X, const double* dX, double* out) {
  if (!ctx || !dX || !out) return MASTB200_ERR_ARG;
  int rc = check_ready(ctx);
  if (rc) return rc;
  if (ctx->elem_type != MASTB200_HEX8 && ctx->elem_type != MASTB200_HEX27)
    return ctx->fail(MASTB200_ERR_UNSUPPORTED, "linearized_jacobian_solution_product: 3D solids (HEX8, HEX27) only");
  CU(cudaSetDevice(ctx->device));
  const double nu = ctx->mat.nu;
  const double lam1 = nu / (1. - nu - 2. * nu * nu), mu1 = 1. / 2. / (1. + nu);
  if (!ctx->mat.nonlinear_strain)  // linear strain: the tangent does not depend on the state
    return run_internal_force(ctx, dX, 0, nullptr, lam1, mu1, out);
  // Green-Lagrange strain: K_T(X) dX matrix-free (material tangent + initial-stress stiffness, mode 5 of the force kernels)
  if (!X) return MASTB200_ERR_ARG;
  if (ctx->thermal)
    return ctx->fail(MASTB200_ERR_UNSUPPORTED, "tangent product with Green-Lagrange strain and a thermal load is not on the device path");
  const size_t n_dof = size_t(ctx->n_nodes) * ctx->nv;
  const double* ddX = nullptr;
  if ((rc = resolve_in(ctx, dX, n_dof, ctx->d_adj, &ddX))) return rc;
  if ((rc = localize_if_needed(ctx, ddX, ctx->nv))) return rc;
  return run_internal_force(ctx, X, 5, ddX, lam1, mu1, out, true, true);
}

// -f_th . V over the owned nodes, f_th = the TEMPERATURE volume load as an element force vector (mode 4)
static int thermal_load_dot(mastb200_context* ctx, const double* V, double* val, double* vec_out) {
  // value: the unconstrained load vector (compliance_output.cpp:71-102); derivative: constrained rows zeroed (:265-296)
  int rc;
  const double nu = ctx->mat.nu;
  const size_t n_dof = size_t(ctx->n_nodes) * ctx->nv;
  CU(ctx->d_vec2.ensure(n_dof));
  const double* dV = nullptr;
  if (V && (rc = resolve_in(ctx, V, n_dof, ctx->d_X, &dV))) return rc;
  if ((rc = run_internal_force(ctx, dV ? dV : ctx->d_vec2.p, 4, nullptr, nu / (1. - nu - 2. * nu * nu), 1. / 2. / (1. + nu),
                               ctx->d_vec2.p, val == nullptr)))
    return rc;
  if (vec_out) CU(cudaMemcpyAsync(vec_out, ctx->d_vec2.p, n_dof * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  if (val) {
    const int nb = 1024;
    CU(ctx->d_partial.ensure(4096));
    CU(ctx->d_scalar.ensure(8));
    CU(launch_dot_owned(ctx->d_vec2.p, dV, ctx->d_owned_nodes.p, ctx->pat.n_owned, ctx->nv, ctx->d_partial.p, nb, ctx->stream));
    CU(launch_reduce_partials(ctx->d_partial.p, nb, ctx->d_scalar.p + 1, ctx->stream));
    ctx->launches += 2;
    CU(cudaMemcpyAsync(val, ctx->d_scalar.p + 1, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
  }
  return MASTB200_OK;
}

int mastb200_compliance(mastb200_context* ctx, const double* X, double* q) {
  if (!ctx || !q || !X) return MASTB200_ERR_ARG;
  int rc = check_ready(ctx);
  if (rc) return rc;
  CU(cudaSetDevice(ctx->device));
  *q = 0.;
  double q_th = 0.;
  if (ctx->thermal) {  // volume_external_residual: vec = -f_th (thermal_residual), compliance -= vec . u
    if (ctx->elem_type != MASTB200_HEX8 && ctx->elem_type != MASTB200_HEX27)
      return ctx->fail(MASTB200_ERR_UNSUPPORTED, "compliance with a thermoelastic load: 3D solids only");
    if ((rc = thermal_load_dot(ctx, X, &q_th, nullptr))) return rc;
    *q = q_th;
  }
  if (ctx->fl_node.empty()) return MASTB200_OK;
  const double* dX = nullptr;
  if ((rc = resolve_in(ctx, X, size_t(ctx->n_nodes) * ctx->nv, ctx->d_X, &dX))) return rc;
  const int nb = face_loads_blocks((int)ctx->fl_node.size());
  CU(ctx->d_partial.ensure(std::max(nb, 4096)));
  CU(ctx->d_scalar.ensure(8));
  CU(launch_face_loads(face_plan(ctx), ctx->d_xyz.p, dX, nullptr, ctx->nv, nullptr, nullptr, ctx->d_partial.p, nullptr, ctx->stream));
  CU(launch_reduce_partials(ctx->d_partial.p, nb, ctx->d_scalar.p, ctx->stream));
  ctx->launches += 2;
  CU(cudaMemcpyAsync(q, ctx->d_scalar.p, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  *q += q_th;
  return MASTB200_OK;
}

int mastb200_compliance_direct_sensitivity(mastb200_context* ctx, const double* X, const double* dXdp, double* dq) {
  // ComplianceOutput::evaluate_sensitivity (compliance_output.cpp:106-150): -d(f_ext)/dp . u - f_ext . du/dp; the loads on
  // the device path (non-follower pressure, thermal load at fixed modulus) do not depend on the parameter, and the
  // output is linear in the state
  (void)X;
  if (!ctx || !dq) return MASTB200_ERR_ARG;
  *dq = 0.;
  if (!dXdp) return MASTB200_OK;
  return mastb200_compliance(ctx, dXdp, dq);
}

int mastb200_compliance_derivative(mastb200_context* ctx, const double* X, double* dq_dX) {
  if (!ctx || !dq_dX) return MASTB200_ERR_ARG;
  (void)X;  // linear in X: -f_ext does not depend on the state for non-follower pressure loads
  int rc = check_ready(ctx);
  if (rc) return rc;
  CU(cudaSetDevice(ctx->device));
  const size_t n_dof = size_t(ctx->n_nodes) * ctx->nv;
  double* d = nullptr;
  if (ctx->thermal) {  // dq/dX = -vec = +f_th on the unconstrained dofs; the pressure part is set on its own (loaded) dofs
    if (ctx->elem_type != MASTB200_HEX8 && ctx->elem_type != MASTB200_HEX27)
      return ctx->fail(MASTB200_ERR_UNSUPPORTED, "compliance with a thermoelastic load: 3D solids only");
    if (!ctx->fl_node.empty())
      return ctx->fail(MASTB200_ERR_UNSUPPORTED, "compliance derivative with both a pressure and a thermoelastic load is not on the device path");
    if ((rc = thermal_load_dot(ctx, nullptr, nullptr, nullptr))) return rc;
    if (is_device_ptr(dq_dX)) CU(cudaMemcpyAsync(dq_dX, ctx->d_vec2.p, n_dof * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    else CU(cudaMemcpyAsync(dq_dX, ctx->d_vec2.p, n_dof * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return MASTB200_OK;
  }
  if ((rc = resolve_out(ctx, dq_dX, n_dof, ctx->d_dq, &d))) return rc;
  CU(cudaMemsetAsync(d, 0, n_dof * sizeof(double), ctx->stream));
  if (!ctx->fl_node.empty()) {
    CU(launch_face_loads(face_plan(ctx), ctx->d_xyz.p, nullptr, ctx->has_constraints ? ctx->d_cmask.p : nullptr, ctx->nv,
                         nullptr, d, nullptr, nullptr, ctx->stream));
    ctx->launches++;
  }
  if ((rc = finish_out(ctx, dq_dX, n_dof, d))) return rc;
  CU(cudaStreamSynchronize(ctx->stream));
  return MASTB200_OK;
}


int mastb200_stress(mastb200_context* ctx, const double* X, double* stress, double* strain, double* von_mises,
                    double* jxw) {
  if (!ctx) return MASTB200_ERR_ARG;
  StressParams p;
  int rc = fill_stress_params(ctx, X, p);
  if (rc) return rc;
  const size_t nqp = size_t(ctx->n_elem) * p.nq;
  if ((rc = resolve_out(ctx, stress, nqp * 6, ctx->d_st[0], &p.out_stress))) return rc;
  if ((rc = resolve_out(ctx, strain, nqp * 6, ctx->d_st[1], &p.out_strain))) return rc;
  if ((rc = resolve_out(ctx, von_mises, nqp, ctx->d_st[2], &p.out_vm))) return rc;
  if ((rc = resolve_out(ctx, jxw, nqp, ctx->d_st[3], &p.out_jxw))) return rc;
  CU(launch_stress_qp(p, ctx->stream));
  ctx->launches++;
  if ((rc = finish_out(ctx, stress, nqp * 6, p.out_stress))) return rc;
  if ((rc = finish_out(ctx, strain, nqp * 6, p.out_strain))) return rc;
  if ((rc = finish_out(ctx, von_mises, nqp, p.out_vm))) return rc;
  if ((rc = finish_out(ctx, jxw, nqp, p.out_jxw))) return rc;
  CU(cudaStreamSynchronize(ctx->stream));
  return MASTB200_OK;
}

int mastb200_stress_functional(mastb200_context* ctx, const double* X, double p_norm, double rho_exp, double sigma0,
                               const uint8_t* elem_mask, double* sums, double* q) {
  if (!ctx || !sums) return MASTB200_ERR_ARG;
  if (!(sigma0 > 0.) || !(p_norm > 0.)) return ctx->fail(MASTB200_ERR_ARG, "sigma0 and p_norm must be positive");
  StressParams p;
  int rc = fill_stress_params(ctx, X, p);
  if (rc) return rc;
  if ((rc = upload_elem_mask(ctx, elem_mask, &p.elem_mask))) return rc;
  p.p_norm = p_norm; p.rho_exp = rho_exp; p.sigma0 = sigma0;
  const int nb = stress_blocks(p.n_elem, p.nq);
  CU(ctx->d_partial.ensure(std::max<size_t>(2 * (size_t)nb, 4096)));
  CU(ctx->d_scalar.ensure(8));
  p.partial = ctx->d_partial.p;
  CU(cudaEventRecord(ctx->ev[0], ctx->stream));
  CU(cudaEventRecord(ctx->ev[1], ctx->stream));
  CU(launch_stress_qp(p, ctx->stream));
  CU(cudaEventRecord(ctx->ev[2], ctx->stream));
  CU(launch_reduce_pairs(ctx->d_partial.p, nb, ctx->d_scalar.p, ctx->stream));
  CU(cudaEventRecord(ctx->ev[3], ctx->stream));
  ctx->timed = true;
  ctx->launches += 2;
  CU(cudaMemcpyAsync(sums, ctx->d_scalar.p, 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  if (q) *q = sigma0 * std::pow(sums[0] / sums[1], 1. / p_norm);
  return MASTB200_OK;
}

int mastb200_stress_functional_derivative(mastb200_context* ctx, const double* X, double p_norm, double rho_exp,
                                          double sigma0, const uint8_t* elem_mask, const double* sums, double* dq_dX) {
  if (!ctx || !sums || !dq_dX) return MASTB200_ERR_ARG;
  if (!(sigma0 > 0.) || !(p_norm > 0.)) return ctx->fail(MASTB200_ERR_ARG, "sigma0 and p_norm must be positive");
  StressParams p;
  int rc = fill_stress_params(ctx, X, p);
  if (rc) return rc;
  if ((rc = upload_elem_mask(ctx, elem_mask, &p.elem_mask))) return rc;
  p.p_norm = p_norm; p.rho_exp = rho_exp; p.sigma0 = sigma0;
  p.sum_num = sums[0]; p.sum_den = sums[1];
  const size_t n_dof = size_t(ctx->n_nodes) * ctx->nv;
  CU(ctx->d_scratch.ensure(size_t(ctx->n_elem) * p.nq * 9));
  p.scratch = ctx->d_scratch.p;
  if ((rc = resolve_out(ctx, dq_dX, n_dof, ctx->d_dq, &p.dq_dX))) return rc;
  CU(cudaEventRecord(ctx->ev[0], ctx->stream));
  if (ctx->has_ghosts) CU(cudaMemsetAsync(p.dq_dX, 0, n_dof * sizeof(double), ctx->stream));
  CU(cudaEventRecord(ctx->ev[1], ctx->stream));
  CU(launch_stress_derivative(p, ctx->stream));
  CU(cudaEventRecord(ctx->ev[2], ctx->stream));
  CU(cudaEventRecord(ctx->ev[3], ctx->stream));
  ctx->timed = true;
  ctx->launches += 2;
  if ((rc = finish_out(ctx, dq_dX, n_dof, p.dq_dX))) return rc;
  CU(cudaStreamSynchronize(ctx->stream));
  return MASTB200_OK;
}

int mastb200_stress_functional_density_derivative(mastb200_context* ctx, const double* X, double p_norm, double rho_exp,
                                                  double sigma0, const uint8_t* elem_mask, const double* sums, double* dq_drho) {
  if (!ctx || !sums || !dq_drho) return MASTB200_ERR_ARG;
  if (!(sigma0 > 0.) || !(p_norm > 0.)) return ctx->fail(MASTB200_ERR_ARG, "sigma0 and p_norm must be positive");
  StressParams p;
  int rc = fill_stress_params(ctx, X, p);
  if (rc) return rc;
  if (!ctx->rho_set) return ctx->fail(MASTB200_ERR_ARG, "density derivatives need a nodal density field");
  if ((rc = upload_elem_mask(ctx, elem_mask, &p.elem_mask))) return rc;
  p.p_norm = p_norm; p.rho_exp = rho_exp; p.sigma0 = sigma0;
  p.sum_num = sums[0]; p.sum_den = sums[1];
  CU(ctx->d_scratch.ensure(size_t(ctx->n_elem) * p.nq));
  p.scratch = ctx->d_scratch.p;
  double* dG = nullptr;
  if ((rc = resolve_out(ctx, dq_drho, (size_t)ctx->n_nodes, ctx->d_G, &dG))) return rc;
  CU(cudaEventRecord(ctx->ev[0], ctx->stream));
  if (ctx->has_ghosts) CU(cudaMemsetAsync(dG, 0, size_t(ctx->n_nodes) * sizeof(double), ctx->stream));
  CU(cudaEventRecord(ctx->ev[1], ctx->stream));
  CU(launch_stress_density_derivative(p, dG, ctx->stream));
  CU(cudaEventRecord(ctx->ev[2], ctx->stream));
  CU(cudaEventRecord(ctx->ev[3], ctx->stream));
  ctx->timed = true; ctx->carry_ms = 0.;
  ctx->launches += 2;
  if ((rc = finish_out(ctx, dq_drho, (size_t)ctx->n_nodes, dG))) return rc;
  CU(cudaStreamSynchronize(ctx->stream));
  return MASTB200_OK;
}

int mastb200_get_stats(const mastb200_context* ctx, mastb200_stats* s) {
  if (!ctx || !s) return MASTB200_ERR_ARG;
  std::memset(s, 0, sizeof *s);
  const mastb200_context::DevicePlan& D = ctx->dplan[ctx->last_plan];
  const bool lm = ctx->last_struct && ctx->splan.last_march;  // the march kernel ran the last structured sweep
  s->n_tiles = ctx->last_struct ? (lm ? ctx->splan.m_n_units : ctx->splan.n_units) : D.n_tiles;
  s->n_visits = ctx->last_struct ? (lm ? ctx->splan.m_n_visits : ctx->splan.n_visits) : D.n_visits;
  s->n_owned_nodes = ctx->pat.n_owned;
  s->n_elem = ctx->n_elem;
  s->kernel_launches = ctx->launches;
  s->plan_bytes = ctx->last_struct ? (int64_t)(lm ? ctx->splan.m_n_units : ctx->splan.n_units) * (int64_t)sizeof(StructUnit) : D.bytes;
  s->smem_bytes = ctx->last_struct ? (int32_t)(lm ? hex8_march_smem_bytes() : hex8_struct_smem_bytes()) : (int32_t)D.smem;
  s->threads_per_cta = ctx->last_struct ? (lm ? hex8_march_threads() : hex8_struct_threads()) : hex8_sweep_threads();
  s->ctas = ctx->last_struct ? (lm ? ctx->splan.m_n_ctas : ctx->splan.n_ctas) : D.n_ctas;
  s->structured = ctx->last_struct ? (lm ? 2 : (ctx->splan.last_slots ? 3 : 1)) : (ctx->elem_type == MASTB200_QUAD4 && ctx->last_q4s ? 1 : 0);
  if (ctx->timed) {
    if (ctx->async_device) cudaStreamSynchronize(ctx->stream);  // the events of the last sweep may still be in flight
    float a = 0.f, b = 0.f;
    if (cudaEventElapsedTime(&a, ctx->ev[0], ctx->ev[3]) == cudaSuccess) s->last_sweep_ms = a + ctx->carry_ms;
    if (cudaEventElapsedTime(&b, ctx->ev[1], ctx->ev[2]) == cudaSuccess) s->last_main_kernel_ms = b + ctx->carry_ms;
  }
  return MASTB200_OK;
}

int mastb200_debug_plan_sizes(mastb200_context* ctx, int32_t aux, int64_t* sizes) {
  if (!ctx || !sizes || aux < 0 || aux > 1) return MASTB200_ERR_ARG;
  if (!ctx->planning_only) return ctx->fail(MASTB200_ERR_ARG, "plan introspection needs a planning-only context (device -1)");
  if (!ctx->have_plan) return ctx->fail(MASTB200_ERR_ARG, "build_pattern first");
  if (int rc = ensure_plan(ctx, aux)) return rc;
  const StreamPlan& P = ctx->dplan[aux].host;
  sizes[0] = (int64_t)P.tiles.size(); sizes[1] = (int64_t)P.steps.size(); sizes[2] = (int64_t)P.visits.size();
  sizes[3] = (int64_t)P.flush_rows.size(); sizes[4] = (int64_t)P.flush_ent.size(); sizes[5] = (int64_t)P.loads.size();
  sizes[6] = P.lim.arena_blocks; sizes[7] = P.lim.max_slots; sizes[8] = P.lim.n_workers;
  return MASTB200_OK;
}

int mastb200_debug_plan_copy(mastb200_context* ctx, int32_t aux, void* tiles, void* steps, void* visits, void* flush_rows,
                             void* flush_ent, void* loads) {
  int64_t sz[9];
  if (int rc = mastb200_debug_plan_sizes(ctx, aux, sz)) return rc;
  const StreamPlan& P = ctx->dplan[aux].host;
  if (tiles) std::memcpy(tiles, P.tiles.data(), P.tiles.size() * sizeof(TileDesc));
  if (steps) std::memcpy(steps, P.steps.data(), P.steps.size() * sizeof(StepDesc));
  if (visits) std::memcpy(visits, P.visits.data(), P.visits.size() * sizeof(VisitRec8));
  if (flush_rows) std::memcpy(flush_rows, P.flush_rows.data(), P.flush_rows.size() * sizeof(FlushRow));
  if (flush_ent) std::memcpy(flush_ent, P.flush_ent.data(), P.flush_ent.size() * 4);
  if (loads) std::memcpy(loads, P.loads.data(), P.loads.size() * sizeof(LoadRec));
  return MASTB200_OK;
}

}  // extern "C"
