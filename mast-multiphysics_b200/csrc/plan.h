// plan.h -- static assembly plan: sparsity pattern + streaming node tiles ("strips") + element visit lists.
//
// The device sweep is a row-owner gather.  The owned nodes are cut into strips: a small
// cross-section of node columns swept along the mesh's longest axis, one "front" (node plane) per step.
// A CTA walks a strip step by step; per step it visits every element whose last owned front is that step,
// accumulates the element's contributions in shared memory and, one step behind the visits, writes out every CSR
// row that has become complete -- each CSR value goes to HBM exactly once and K_e is never stored.
//
// Device timeline the lifetimes below are planned for (sweep_hex8.cu): the flush of step s runs concurrently with
// the visits of step s+1; the visits of step s+2 start only after it has finished.  Hence a pair block whose last
// reader is the flush of step t may be re-initialised by visits of step >= t+2.  Nodal state is staged two steps
// ahead (state of step s+2 while step s is visited), so a state slot last read in step t (visit or flush) may be
// reloaded for a node first visited in step >= t+3.
//
// Shared-memory accumulators are kept per unordered node pair {A,B} ("pair block", 9 doubles of
// S_AB = sum_q w_q grad N_A (x) grad N_B; S_BA = S_AB^T, so one block serves rows A and B).  A pair block lives
// from the first visit that touches it until the rows of its owned endpoints are written; slots are handed out
// by the planner, so a strip only ever holds about two fronts of pairs no matter how long it is.  The same
// holds for the staged nodal state (coordinates, solution, ...), which lives in planner-assigned "state slots".
//
// Ordering: the visits of a step are laid out on a grid of n_workers (compute warps) x positions; visits at the same
// position share no owned node.  The accumulation order is (step, position, worker): rank[a] of a visit is the
// number of earlier visits of the strip touching owned node a; the device adds into the pairs of node a only once
// the node's counter equals rank[a], which fixes the summation order into every CSR entry (deterministic,
// independent of scheduling) without CTA-wide barriers inside a step.  Each warp walks its own row of the grid,
// so a visit only waits for visits at earlier positions, which the other warps have normally already finished.
//
// Everything here is integer bookkeeping built once per mesh, like libMesh's sparsity pattern /
// PETSc preallocation that the reference relies on (nonlinear_implicit_assembly.cpp:182-183).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace mastb200 {

// ---- device records ------------------------------------------------------------------------------------------
struct TileDesc {       // one strip segment
  int32_t step_begin;   // into steps
  int32_t n_steps;
  int32_t n_visits;     // total, for scheduling (tiles are sorted by descending cost)
  int32_t n_owned;
};

struct StepDesc {
  int32_t visit_begin, n_visits;   // grid of visit records [worker][position] (n_workers x n_pos); n_visits real ones
  int32_t load_begin, n_loads;     // node states that must be resident before the step's visits start
  int32_t flush_begin, n_flush;    // rows written after the step
  int32_t fent_begin, n_fent;      // flush entries of those rows
  int32_t n_pos;                   // positions per worker in this step
  int32_t pad[3];
};
static_assert(sizeof(StepDesc) == 48, "StepDesc must be 48 bytes");

// One element visit of a HEX8 strip, 96 bytes (a chunk of 4 visits = 3 coalesced 4-byte loads per lane).
struct alignas(16) VisitRec8 {
  uint16_t ln[8];     // state slots of the element's nodes (libMesh node order); ln[0] == 0xFFFF: empty grid position
  uint8_t rank[8];    // accumulation rank per node; 0xFF: node not owned by the strip
  uint16_t pair[36];  // local node pairs (a<=b), index a*(15-a)/2+b: kPairNone, or
                      // bits 0-11 arena block, bit 12: stored block is S_(b,a) (transposed), bit 13: first touch (store, don't add)
};
static_assert(sizeof(VisitRec8) == 96, "VisitRec8 must be 96 bytes");
constexpr uint16_t kPairNone = 0xFFFF;
constexpr uint16_t kPairSlotMask = 0x0FFF, kPairTransposed = 0x1000, kPairInit = 0x2000;

struct alignas(16) FlushRow {   // one owned node whose 3 (nv=3) or 6 CSR rows are complete
  int64_t dst_block;    // offset of the node's first neighbour block in the node-level CSR (nbr_ptr[row])
  int32_t gid;          // global node id
  int32_t ent_rel;      // first flush entry of the row, relative to the step's fent_begin
  uint16_t slot;        // state slot of the node
  uint16_t nnb;         // neighbour count (== number of flush entries)
  uint16_t nadj;        // elements adjacent to the node (not in diagonal_elem_subdomain_id)
  uint16_t nadjd;       // adjacent elements of diagonal_elem_subdomain_id
  int32_t pad[2];
};
static_assert(sizeof(FlushRow) == 32, "FlushRow must be 32 bytes");
// flush entry (uint32), one per neighbour in CSR column order:
//   bits 0-11 arena block (kFlushNoBlock: neighbour only through diagonal elements), bit 12: block holds S_(nbr,row),
//   bit 13: self pair (upper triangle valid), bits 16-31: state slot of the neighbour (0xFFFF: not staged)
constexpr uint32_t kFlushNoBlock = 0x0FFF, kFlushTransposed = 0x1000, kFlushSelf = 0x2000;

struct LoadRec {        // node state to stage: global id -> slot
  int32_t gid;
  uint16_t slot;
  uint16_t owned;
};

// ---- host plan -----------------------------------------------------------------------------------------------
struct PlanLimits {
  int32_t arena_blocks;     // pair blocks (9 doubles each) that fit the shared-memory arena (<= 4095)
  int32_t max_slots;        // node state slots
  int32_t flush_buf_bytes;  // per-step flush records staged in shared memory
  int32_t sm_count;         // for the tile count heuristic
  int32_t n_workers;        // compute warps of the sweep kernel: rows of the per-step visit grid
};

struct Pattern {
  int32_t nn = 0;           // nodes per element
  int64_t n_nodes = 0, n_elem = 0, n_owned = 0;
  // node-level sparsity of the owned rows: neighbours of owned node k (ascending node id)
  std::vector<int32_t> owned_nodes;      // [n_owned] ascending
  std::vector<int64_t> nbr_ptr;          // [n_owned+1]
  std::vector<int32_t> nbr;              // [nbr_ptr[n_owned]]
  std::vector<int32_t> node_row;         // [n_nodes] index into owned_nodes or -1
  std::vector<int32_t> n_adj_elem;       // [n_nodes] number of (non-diagonal) elements touching the node
  std::vector<int32_t> n_adj_diag_elem;  // [n_nodes] number of diagonal_elem_subdomain elements touching it
  std::vector<int64_t> n2e_ptr;          // node -> element adjacency (ascending element id)
  std::vector<int32_t> n2e;
};

struct StreamPlan {
  PlanLimits lim{};
  std::vector<TileDesc> tiles;
  std::vector<StepDesc> steps;
  std::vector<VisitRec8> visits;
  std::vector<FlushRow> flush_rows;
  std::vector<uint32_t> flush_ent;
  std::vector<LoadRec> loads;
  int32_t max_live_blocks = 0, max_live_slots = 0;
  int64_t n_visits() const { int64_t n = 0; for (const TileDesc& t : tiles) n += t.n_visits; return n; }
  int64_t bytes() const;
};

// Sparsity pattern of the owned rows. conn: [n_elem][nn]; owned: [n_nodes] flags or nullptr;
// diag_elem: [n_elem] or nullptr (elements of AssemblyBase::diagonal_elem_subdomain_id take no part in the sweep
// but are in the pattern).  Returns empty string on success, else an error message.
std::string build_pattern(int32_t nn, int64_t n_nodes, int64_t n_elem, const int32_t* conn, const uint8_t* owned,
                          const uint8_t* diag_elem, Pattern& pat);

// Streaming plan for HEX8 (nn == 8) within the given shared-memory limits.
std::string build_stream_plan(const Pattern& pat, const double* xyz, const int32_t* conn, const uint8_t* diag_elem,
                              const PlanLimits& lim, StreamPlan& plan);

// ---- coloured scatter plan (HEX27) -----------------------------------------------------------------------------
// Greedy colouring in element id order (smallest colour no node-sharing element has taken): no two elements of a colour
// share a node, so a colour can be scattered into the CSR without atomics.  Elements without an owned node get no colour.
// color_ptr[n_colors+1] indexes color_elems (element ids, ascending within a colour).
std::string build_coloring(const Pattern& pat, const int32_t* conn, std::vector<int32_t>& color_ptr,
                           std::vector<int32_t>& color_elems);
// element -> nnz map: posmap[e][a][b] = position of node conn[e][b] in the neighbour list of the owned node conn[e][a]
// (0 for unowned a).  One byte per node pair; fails if a row has more than 255 neighbours.
std::string build_posmap(const Pattern& pat, const int32_t* conn, std::vector<uint8_t>& posmap);
// processing order, predecessor lists and position map (bit 7: first writer of the block) of the single-launch scatter
// first-writer flags (bit 7 of the position map) for a scatter that runs group by group in the given order
bool flag_first_writers(const Pattern& pat, const int32_t* conn, const uint8_t* diag_elem, const std::vector<int32_t>& group_ptr,
                        const std::vector<int32_t>& elems, bool parallel_groups, std::vector<uint8_t>& posmap);
std::string build_ordered_scatter(const Pattern& pat, const double* xyz, const int32_t* conn, const uint8_t* diag_elem,
                                  int brick_elems, std::vector<int32_t>& order, std::vector<int32_t>& pred_ptr,
                                  std::vector<int32_t>& pred, std::vector<uint8_t>& posmap, bool& first_writer);

}  // namespace mastb200
