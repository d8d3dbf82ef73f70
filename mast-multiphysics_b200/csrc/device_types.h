// device_types.h -- parameter block shared by the host context and the sweep kernels.
#pragma once
#include <cstdint>

#include "plan.h"

namespace mastb200 {

struct SweepParams {
  // ---- static plan (device pointers) ----
  const TileDesc* tiles;
  int32_t n_tiles;
  const StepDesc* steps;
  const VisitRec8* visits;
  const FlushRow* flush_rows;
  const uint32_t* flush_ent;
  const LoadRec* loads;
  int32_t* tile_counter;      // dynamic tile scheduler (zeroed before every launch)
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
  int32_t nv;
  int32_t thermal;            // TEMPERATURE volume load active
  int32_t aux;                // 1: adjoint / density / temperature state and per-node extras are staged (kernel variant)
};

}  // namespace mastb200
