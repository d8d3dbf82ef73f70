// plan.h -- static assembly plan: sparsity pattern + node tiles + element visit lists.
//
// The device sweep is a row-owner gather: the owned nodes are cut into compact
// tiles; a CTA keeps the CSR rows of one tile in shared memory, visits every element adjacent to the
// tile ("visit"), adds the element's contribution at precomputed slots (the element-to-nnz map of
// BASELINE.json's north_star, at node-block granularity) and writes each CSR value exactly once.
// Within a tile, visits are grouped into steps; the visits of one step share no owned node, so warps
// can accumulate concurrently and the summation order into every CSR entry is fixed by the plan
// (ascending step, i.e. deterministic and independent of scheduling).
//
// Everything here is integer bookkeeping built once per mesh, like libMesh's sparsity pattern /
// PETSc preallocation that the reference relies on (nonlinear_implicit_assembly.cpp:182-183).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace mastb200 {

struct TileDesc {
  int32_t node_begin;     // offset into tile_nodes (owned nodes first, sorted by id; then halo nodes)
  int32_t n_owned;
  int32_t n_nodes;        // owned + halo
  int32_t visit_begin;    // visits of the tile, in the order in which they are accumulated
  int32_t n_visits;
  int32_t arena_doubles;  // shared-memory accumulator doubles used by this tile
  int32_t nbrl_begin;     // offset into nbr_local of the first owned node's neighbour list
  int32_t blocks_total;   // sum over owned nodes of their neighbour counts
};

// One element visit of a HEX8 tile, 96 bytes (six 16-byte words; a chunk of 4 visits is loaded with 3 coalesced
// 4-byte loads per lane).  rank[a] is the number of earlier visits of the same tile that touch owned node a: the
// device accumulates into node a only once its per-node counter equals rank[a], which fixes the summation order
// into every CSR entry without CTA-wide barriers.
struct alignas(16) VisitRec8 {
  uint16_t ln[8];    // tile-local node ids
  uint8_t rank[8];
  int32_t elem;
  int32_t pad;
  uint8_t slot[64];  // [a][b]: position of node b in the neighbour list (CSR row) of node a, for owned a
};
static_assert(sizeof(VisitRec8) == 96, "VisitRec8 must be 96 bytes");

struct PlanLimits {
  int32_t arena_doubles;    // accumulator capacity in doubles
  int32_t max_tile_nodes;   // owned + halo nodes whose state is staged in shared memory
  int32_t acc_per_block;    // doubles accumulated per (row node, column node) pair: 9 solid, 1 heat
  int32_t acc_per_node;     // extra doubles per owned node (f, G, rot-diag)
};

struct Plan {
  int32_t nn = 0;           // nodes per element
  int64_t n_nodes = 0, n_elem = 0, n_owned = 0;
  // node-level sparsity of the owned rows: neighbours of owned node k (ascending node id)
  std::vector<int32_t> owned_nodes;      // [n_owned] ascending
  std::vector<int64_t> nbr_ptr;          // [n_owned+1]
  std::vector<int32_t> nbr;              // [nbr_ptr[n_owned]]
  std::vector<int32_t> node_row;         // [n_nodes] index into owned_nodes or -1
  std::vector<int32_t> n_adj_elem;       // [n_nodes] number of (non-diagonal) elements touching the node
  std::vector<int32_t> n_adj_diag_elem;  // [n_nodes] number of diagonal_elem_subdomain elements touching it
  // tiles
  std::vector<TileDesc> tiles;
  std::vector<int32_t> tile_nodes;       // global node id per tile-local node
  std::vector<int32_t> tile_acc_off;     // per tile-local node: accumulator offset (owned) or -1
  std::vector<int32_t> tile_nbr_cnt;     // per tile-local node: neighbour count (owned) or 0
  std::vector<uint16_t> nbr_local;       // per owned tile node, per neighbour: tile-local id of the neighbour
  std::vector<VisitRec8> visits8;        // HEX8 visit records
  int64_t n_visits() const { return (int64_t)visits8.size(); }
  int64_t bytes() const;
};

// Builds pattern and tiles. conn: [n_elem][nn]; owned: [n_nodes] flags or nullptr; diag_elem: [n_elem] or nullptr
// (elements of AssemblyBase::diagonal_elem_subdomain_id take no part in the sweep but are in the pattern).
// Returns empty string on success, else an error message.
std::string build_plan(int32_t nn, int64_t n_nodes, int64_t n_elem, const double* xyz, const int32_t* conn,
                       const uint8_t* owned, const uint8_t* diag_elem, const PlanLimits& lim, Plan& plan);

}  // namespace mastb200
