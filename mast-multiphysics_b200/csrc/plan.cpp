// plan.cpp -- host-side construction of the assembly plan (see plan.h).
#include "plan.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <thread>

namespace mastb200 {

int64_t Plan::bytes() const {
  return (int64_t)(tiles.size() * sizeof(TileDesc) + tile_nodes.size() * 4 + tile_acc_off.size() * 4 +
                   tile_nbr_cnt.size() * 4 + nbr_local.size() * 2 + visits8.size() * sizeof(VisitRec8));
}

namespace {

template <typename F>
void parallel_for(int64_t n, F f) {
  unsigned hw = std::thread::hardware_concurrency();
  int nt = (int)std::max(1u, std::min(hw ? hw : 1u, 32u));
  if (n < 4096) nt = 1;
  if (nt == 1) { f(0, n, 0); return; }
  std::vector<std::thread> pool;
  for (int t = 0; t < nt; ++t) pool.emplace_back([=] { f(n * t / nt, n * (t + 1) / nt, t); });
  for (auto& th : pool) th.join();
}

struct TileBuild {
  std::vector<int32_t> nodes, acc_off, nbr_cnt;
  std::vector<uint16_t> nbrl;
  std::vector<VisitRec8> visits;
  int32_t n_owned = 0, arena = 0, blocks = 0;
};

}  // namespace

std::string build_plan(int32_t nn, int64_t n_nodes, int64_t n_elem, const double* xyz, const int32_t* conn,
                       const uint8_t* owned, const uint8_t* diag_elem, const PlanLimits& lim, Plan& P) {
  P = Plan();
  P.nn = nn; P.n_nodes = n_nodes; P.n_elem = n_elem;
  if (n_nodes <= 0 || n_elem <= 0) return "empty mesh";
  if (n_nodes >= (int64_t(1) << 31) || n_elem >= (int64_t(1) << 31)) return "mesh too large for 32-bit ids";
  for (int64_t k = 0; k < n_elem * nn; ++k)
    if (conn[k] < 0 || conn[k] >= n_nodes) return "connectivity entry out of range";

  // ---- node -> element adjacency (ascending element id) -------------------------------------
  std::vector<int64_t> n2e_ptr(n_nodes + 1, 0);
  for (int64_t k = 0; k < n_elem * nn; ++k) n2e_ptr[conn[k] + 1]++;
  for (int64_t n = 0; n < n_nodes; ++n) n2e_ptr[n + 1] += n2e_ptr[n];
  std::vector<int32_t> n2e(n2e_ptr[n_nodes]);
  {
    std::vector<int64_t> pos(n2e_ptr.begin(), n2e_ptr.end() - 1);
    for (int64_t e = 0; e < n_elem; ++e)
      for (int a = 0; a < nn; ++a) n2e[pos[conn[e * nn + a]]++] = (int32_t)e;
  }
  P.n_adj_elem.assign(n_nodes, 0);
  P.n_adj_diag_elem.assign(n_nodes, 0);
  for (int64_t n = 0; n < n_nodes; ++n)
    for (int64_t k = n2e_ptr[n]; k < n2e_ptr[n + 1]; ++k) {
      if (diag_elem && diag_elem[n2e[k]]) P.n_adj_diag_elem[n]++;
      else P.n_adj_elem[n]++;
    }

  // ---- owned rows and node-level sparsity ----------------------------------------------------
  P.node_row.assign(n_nodes, -1);
  for (int64_t n = 0; n < n_nodes; ++n)
    if (!owned || owned[n]) { P.node_row[n] = (int32_t)P.owned_nodes.size(); P.owned_nodes.push_back((int32_t)n); }
  P.n_owned = (int64_t)P.owned_nodes.size();
  if (P.n_owned == 0) return "no owned nodes";
  P.nbr_ptr.assign(P.n_owned + 1, 0);
  {
    // pass 1: counts; pass 2: fill (both parallel over owned rows)
    std::vector<int32_t> cnt(P.n_owned);
    auto gather = [&](int64_t n, std::vector<int32_t>& l) {
      l.clear();
      for (int64_t k = n2e_ptr[n]; k < n2e_ptr[n + 1]; ++k) {
        const int32_t* c = conn + (int64_t)n2e[k] * nn;
        l.insert(l.end(), c, c + nn);
      }
      std::sort(l.begin(), l.end());
      l.erase(std::unique(l.begin(), l.end()), l.end());
    };
    parallel_for(P.n_owned, [&](int64_t b, int64_t e, int) {
      std::vector<int32_t> l;
      for (int64_t r = b; r < e; ++r) { gather(P.owned_nodes[r], l); cnt[r] = (int32_t)l.size(); }
    });
    for (int64_t r = 0; r < P.n_owned; ++r) {
      if (cnt[r] > 255) return "node with more than 255 neighbours";
      P.nbr_ptr[r + 1] = P.nbr_ptr[r] + cnt[r];
    }
    P.nbr.resize(P.nbr_ptr[P.n_owned]);
    parallel_for(P.n_owned, [&](int64_t b, int64_t e, int) {
      std::vector<int32_t> l;
      for (int64_t r = b; r < e; ++r) {
        gather(P.owned_nodes[r], l);
        std::copy(l.begin(), l.end(), P.nbr.begin() + P.nbr_ptr[r]);
      }
    });
  }

  if (nn != 8) return "";  // pattern only: the tile plan below is laid out for HEX8 visit records

  // ---- recursive coordinate bisection of the owned nodes into tiles ---------------------------
  // A leaf is accepted when its accumulators fit the arena; the halo limit is verified when the tile is built
  // and the leaf is split again if it fails.
  std::vector<int32_t> perm(P.owned_nodes);
  struct Range { int64_t b, e; };
  std::vector<Range> leaves, work;
  work.push_back({0, P.n_owned});
  auto arena_need = [&](int64_t b, int64_t e) {
    int64_t s = 0;
    for (int64_t k = b; k < e; ++k) {
      const int32_t r = P.node_row[perm[k]];
      s += (P.nbr_ptr[r + 1] - P.nbr_ptr[r]) * lim.acc_per_block + lim.acc_per_node;
    }
    return s;
  };
  auto split = [&](Range r, std::vector<Range>& out) {
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int64_t k = r.b; k < r.e; ++k)
      for (int c = 0; c < 3; ++c) {
        const double v = xyz[(int64_t)perm[k] * 3 + c];
        lo[c] = std::min(lo[c], v); hi[c] = std::max(hi[c], v);
      }
    int ax = 0;
    for (int c = 1; c < 3; ++c)
      if (hi[c] - lo[c] > hi[ax] - lo[ax]) ax = c;
    const int64_t cnt = r.e - r.b;
    const int64_t mid = r.b + cnt / 2;
    auto key = [&](int32_t n) { return xyz[(int64_t)n * 3 + ax]; };
    auto less = [&](int32_t a, int32_t b) { return key(a) < key(b) || (key(a) == key(b) && a < b); };
    // Cut near the median but at the widest coordinate gap inside a window of about one mesh "plane" of nodes, so
    // that tiles of (possibly jittered) structured meshes end on whole node planes and keep a small halo.
    int64_t W = (int64_t)std::ceil(std::pow((double)cnt, 2.0 / 3.0)) + 1;
    W = std::min(W, std::min(mid - r.b, r.e - mid) - 1);
    int64_t cut = mid;
    if (W >= 1) {
      auto b0 = perm.begin() + (mid - W), b1 = perm.begin() + (mid + W);
      std::nth_element(perm.begin() + r.b, b0, perm.begin() + r.e, less);
      std::nth_element(b0, b1, perm.begin() + r.e, less);
      std::sort(b0, b1, less);
      double best = -1.;
      for (int64_t k = mid - W + 1; k < mid + W; ++k) {
        const double gap = key(perm[k]) - key(perm[k - 1]);
        const bool closer = std::llabs(k - mid) < std::llabs(cut - mid);
        if (gap > best * (1. + 1e-9) + 1e-300 || (gap >= best * (1. - 1e-9) && closer)) {
          if (gap > best) best = gap;
          cut = k;
        }
      }
    } else {
      std::nth_element(perm.begin() + r.b, perm.begin() + mid, perm.begin() + r.e, less);
    }
    out.push_back({r.b, cut});
    out.push_back({cut, r.e});
  };
  while (!work.empty()) {
    Range r = work.back();
    work.pop_back();
    if (r.e - r.b <= 1 || (arena_need(r.b, r.e) <= lim.arena_doubles && r.e - r.b <= lim.max_tile_nodes)) leaves.push_back(r);
    else split(r, work);
  }

  // ---- build tiles (parallel), splitting leaves whose halo does not fit -------------------------
  std::vector<TileBuild> built;
  std::string err;
  while (!leaves.empty()) {
    std::sort(leaves.begin(), leaves.end(), [](const Range& a, const Range& b) { return a.b < b.b; });
    std::vector<TileBuild> tb(leaves.size());
    std::vector<char> too_big(leaves.size(), 0);
    std::atomic<int> fail(0);
    parallel_for((int64_t)leaves.size(), [&](int64_t tb0, int64_t tb1, int) {
      std::vector<int32_t> local(n_nodes, -1);  // global -> tile-local
      std::vector<int32_t> elems;
      std::vector<uint32_t> used;  // per owned node: bitmask of steps (colours) already taken
      for (int64_t t = tb0; t < tb1; ++t) {
        TileBuild& T = tb[t];
        const Range r = leaves[t];
        T.nodes.assign(perm.begin() + r.b, perm.begin() + r.e);
        std::sort(T.nodes.begin(), T.nodes.end());
        T.n_owned = (int32_t)T.nodes.size();
        for (int32_t l = 0; l < T.n_owned; ++l) local[T.nodes[l]] = l;
        // visits: every non-diagonal element touching an owned node, ascending id
        elems.clear();
        for (int32_t l = 0; l < T.n_owned; ++l) {
          const int32_t n = T.nodes[l];
          for (int64_t k = n2e_ptr[n]; k < n2e_ptr[n + 1]; ++k)
            if (!(diag_elem && diag_elem[n2e[k]])) elems.push_back(n2e[k]);
        }
        std::sort(elems.begin(), elems.end());
        elems.erase(std::unique(elems.begin(), elems.end()), elems.end());
        // halo nodes in first-touch order
        for (int32_t e : elems)
          for (int a = 0; a < nn; ++a) {
            const int32_t n = conn[(int64_t)e * nn + a];
            if (local[n] < 0) { local[n] = (int32_t)T.nodes.size(); T.nodes.push_back(n); }
          }
        // neighbours of owned nodes always belong to the tile, except through diagonal elements
        bool ok = (int64_t)T.nodes.size() <= lim.max_tile_nodes;
        if (ok) {
          // accumulator layout: node blocks in owned order, then per-node extras
          T.acc_off.assign(T.nodes.size(), -1);
          T.nbr_cnt.assign(T.nodes.size(), 0);
          int32_t off = 0;
          for (int32_t l = 0; l < T.n_owned; ++l) {
            const int32_t row = P.node_row[T.nodes[l]];
            const int32_t cnt = (int32_t)(P.nbr_ptr[row + 1] - P.nbr_ptr[row]);
            T.acc_off[l] = off; T.nbr_cnt[l] = cnt;
            off += cnt * lim.acc_per_block;
            T.blocks += cnt;
            for (int64_t k = P.nbr_ptr[row]; k < P.nbr_ptr[row + 1]; ++k) {
              const int32_t ll = local[P.nbr[k]];
              T.nbrl.push_back(ll < 0 ? (uint16_t)0xFFFF : (uint16_t)ll);  // 0xFFFF: neighbour only through a diagonal element
            }
          }
          T.arena = off + T.n_owned * lim.acc_per_node;
          if (T.arena > lim.arena_doubles) ok = false;
        }
        if (!ok) {
          for (int32_t n : T.nodes) local[n] = -1;
          if (T.n_owned <= 1) fail = 1;
          too_big[t] = 1;
          continue;
        }
        // greedy colouring: visits of one step share no owned node
        used.assign(T.n_owned, 0u);
        std::vector<int32_t> colour(elems.size());
        int32_t ncol = 0;
        for (size_t v = 0; v < elems.size(); ++v) {
          uint32_t m = 0;
          for (int a = 0; a < nn; ++a) {
            const int32_t l = local[conn[(int64_t)elems[v] * nn + a]];
            if (l < T.n_owned) m |= used[l];
          }
          int32_t c = 0;
          while (c < 32 && (m >> c & 1u)) ++c;
          if (c >= 32) { fail = 2; c = 31; }
          colour[v] = c;
          ncol = std::max(ncol, c + 1);
          for (int a = 0; a < nn; ++a) {
            const int32_t l = local[conn[(int64_t)elems[v] * nn + a]];
            if (l < T.n_owned) used[l] |= 1u << c;
          }
        }
        // accumulation order: colour-major (conflicting visits end up far apart in the order), ascending element id
        // within a colour; rank = position of the visit among the visits of each of its owned nodes
        std::vector<int32_t> csize(ncol, 0), start(ncol + 1, 0);
        for (size_t v = 0; v < elems.size(); ++v) csize[colour[v]]++;
        for (int32_t c = 0; c < ncol; ++c) start[c + 1] = start[c] + csize[c];
        const size_t nv_ = elems.size();
        std::vector<int32_t> order(nv_);
        for (size_t v = 0; v < nv_; ++v) order[start[colour[v]]++] = (int32_t)v;
        T.visits.resize(nv_);
        used.assign(T.n_owned, 0u);  // reused as per-node visit counter
        for (size_t o = 0; o < nv_; ++o) {
          const int32_t e = elems[order[o]];
          const int32_t* c = conn + (int64_t)e * nn;
          VisitRec8& V = T.visits[o];
          std::memset(&V, 0, sizeof V);
          V.elem = e;
          for (int a = 0; a < nn; ++a) V.ln[a] = (uint16_t)local[c[a]];
          for (int a = 0; a < nn; ++a) {
            const int32_t l = local[c[a]];
            if (l >= T.n_owned) continue;
            if (used[l] > 250) fail = 3;
            V.rank[a] = (uint8_t)used[l]++;
            const int32_t row = P.node_row[c[a]];
            const int32_t* nb = P.nbr.data() + P.nbr_ptr[row];
            const int32_t* ne = P.nbr.data() + P.nbr_ptr[row + 1];
            for (int b = 0; b < nn; ++b) V.slot[a * nn + b] = (uint8_t)(std::lower_bound(nb, ne, c[b]) - nb);
          }
        }
        for (int32_t n : T.nodes) local[n] = -1;
      }
    });
    if (fail == 1) return "a single node does not fit the shared-memory arena";
    if (fail == 2) return "tile needs more than 32 colours";
    if (fail == 3) return "node touched by more than 250 elements";
    std::vector<Range> next;
    for (size_t t = 0; t < leaves.size(); ++t) {
      if (too_big[t]) split(leaves[t], next);
      else built.push_back(std::move(tb[t]));
    }
    leaves.swap(next);
  }
  // deterministic tile order: by smallest owned node id
  std::sort(built.begin(), built.end(), [](const TileBuild& a, const TileBuild& b) { return a.nodes[0] < b.nodes[0]; });

  // ---- concatenate ---------------------------------------------------------------------------------
  size_t tot_nodes = 0, tot_vis = 0, tot_nbrl = 0;
  for (const TileBuild& T : built) {
    tot_nodes += T.nodes.size(); tot_vis += T.visits.size(); tot_nbrl += T.nbrl.size();
  }
  if (tot_nodes >= (size_t(1) << 31) || tot_vis >= (size_t(1) << 31) || tot_nbrl >= (size_t(1) << 31)) return "plan too large";
  P.tiles.reserve(built.size());
  P.tile_nodes.reserve(tot_nodes); P.tile_acc_off.reserve(tot_nodes); P.tile_nbr_cnt.reserve(tot_nodes);
  P.nbr_local.reserve(tot_nbrl);
  P.visits8.reserve(tot_vis);
  for (TileBuild& T : built) {
    TileDesc d;
    d.node_begin = (int32_t)P.tile_nodes.size();
    d.n_owned = T.n_owned;
    d.n_nodes = (int32_t)T.nodes.size();
    d.visit_begin = (int32_t)P.visits8.size();
    d.n_visits = (int32_t)T.visits.size();
    d.arena_doubles = T.arena;
    d.nbrl_begin = (int32_t)P.nbr_local.size();
    d.blocks_total = T.blocks;
    P.tile_nodes.insert(P.tile_nodes.end(), T.nodes.begin(), T.nodes.end());
    P.tile_acc_off.insert(P.tile_acc_off.end(), T.acc_off.begin(), T.acc_off.end());
    P.tile_nbr_cnt.insert(P.tile_nbr_cnt.end(), T.nbr_cnt.begin(), T.nbr_cnt.end());
    P.nbr_local.insert(P.nbr_local.end(), T.nbrl.begin(), T.nbrl.end());
    P.visits8.insert(P.visits8.end(), T.visits.begin(), T.visits.end());
    P.tiles.push_back(d);
    TileBuild().nodes.swap(T.nodes);  // release memory early
    std::vector<VisitRec8>().swap(T.visits);
  }
  return err;
}

}  // namespace mastb200
