// device_types.h -- parameter block shared by the host context and the sweep kernels.
#pragma once
#include <cstdint>

#include "plan.h"

namespace mastb200 {

struct SweepParams {
  // ---- static plan (device pointers) ----
  const TileDesc* tiles;
  int32_t n_tiles;
  const int32_t* tile_nodes;
  const int32_t* tile_acc_off;
  const int32_t* tile_nbr_cnt;
  const uint16_t* nbr_local;
  const VisitRec8* visits;
  const int64_t* nbr_ptr;     // per owned row: offset (in node blocks) of its first neighbour
  const int32_t* node_row;    // global node -> owned row or -1
  const int32_t* n_adj_elem;  // per node
  const int32_t* n_adj_diag;  // per node (diagonal_elem_subdomain_id elements)
  // ---- mesh + state ----
  const double* xyz;          // [n_nodes][3]
  const double* X;            // [n_nodes*nv]
  const double* adj;          // [n_nodes*nv] or null
  const double* rho;          // [n_nodes] or null
  const double* temperature;  // [n_nodes] or null
  const uint8_t* cmask;       // [n_nodes] bit v set => dof (node, v) constrained; or null
  // ---- outputs ----
  double* J;                  // CSR values of owned rows or null
  double* R;                  // [n_nodes*nv] or null
  double* G;                  // [n_nodes] or null
  // ---- material ----
  double E0, lam1, mu1;       // lam1/mu1: Lame constants for unit modulus
  double rho_min, penalty;
  double alpha_k;             // alpha*(3*lam1+2*mu1): thermal stress per unit modulus per kelvin
  double T_const, T_ref;
  double k0, k_beta, thickness, heat_source;  // heat conduction
  int32_t nv;
  int32_t thermal;            // TEMPERATURE volume load active
  int32_t nonlinear;          // Green-Lagrange strain
  int32_t k_of_T;
  // ---- shared-memory layout (set by the launcher) ----
  int32_t arena_doubles;
  int32_t max_tile_nodes;
};

}  // namespace mastb200
