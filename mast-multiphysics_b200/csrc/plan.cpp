// plan.cpp -- host-side construction of the sparsity pattern and the streaming assembly plan (see plan.h).
#include "plan.h"

#include <algorithm>
#include <atomic>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>

namespace mastb200 {

int64_t StreamPlan::bytes() const {
  return (int64_t)(tiles.size() * sizeof(TileDesc) + steps.size() * sizeof(StepDesc) + visits.size() * sizeof(VisitRec8) +
                   flush_rows.size() * sizeof(FlushRow) + flush_ent.size() * 4 + loads.size() * sizeof(LoadRec));
}

namespace {

int hw_threads() {
  unsigned hw = std::thread::hardware_concurrency();
  return (int)std::max(1u, std::min(hw ? hw : 1u, 32u));
}

template <typename F>
void parallel_for(int64_t n, F f) {
  int nt = hw_threads();
  if (n < 4096) nt = 1;
  if (nt == 1) { f(0, n, 0); return; }
  std::vector<std::thread> pool;
  for (int t = 0; t < nt; ++t) pool.emplace_back([=] { f(n * t / nt, n * (t + 1) / nt, t); });
  for (auto& th : pool) th.join();
}

// dynamic work queue over items [0,n): f(item, thread)
template <typename F>
void parallel_items(int64_t n, F f) {
  int nt = (int)std::min<int64_t>(hw_threads(), std::max<int64_t>(n, 1));
  if (nt <= 1) { for (int64_t i = 0; i < n; ++i) f(i, 0); return; }
  std::atomic<int64_t> next(0);
  std::vector<std::thread> pool;
  for (int t = 0; t < nt; ++t)
    pool.emplace_back([&, t] {
      for (;;) {
        const int64_t i = next.fetch_add(1);
        if (i >= n) break;
        f(i, t);
      }
    });
  for (auto& th : pool) th.join();
}

}  // namespace

std::string build_pattern(int32_t nn, int64_t n_nodes, int64_t n_elem, const int32_t* conn, const uint8_t* owned,
                          const uint8_t* diag_elem, Pattern& P) {
  P = Pattern();
  P.nn = nn; P.n_nodes = n_nodes; P.n_elem = n_elem;
  if (n_nodes <= 0 || n_elem <= 0) return "empty mesh";
  if (n_nodes >= (int64_t(1) << 31) || n_elem >= (int64_t(1) << 31)) return "mesh too large for 32-bit ids";
  for (int64_t k = 0; k < n_elem * nn; ++k)
    if (conn[k] < 0 || conn[k] >= n_nodes) return "connectivity entry out of range";

  // ---- node -> element adjacency (ascending element id) -------------------------------------
  P.n2e_ptr.assign(n_nodes + 1, 0);
  for (int64_t k = 0; k < n_elem * nn; ++k) P.n2e_ptr[conn[k] + 1]++;
  for (int64_t n = 0; n < n_nodes; ++n) P.n2e_ptr[n + 1] += P.n2e_ptr[n];
  P.n2e.resize(P.n2e_ptr[n_nodes]);
  {
    std::vector<int64_t> pos(P.n2e_ptr.begin(), P.n2e_ptr.end() - 1);
    for (int64_t e = 0; e < n_elem; ++e)
      for (int a = 0; a < nn; ++a) P.n2e[pos[conn[e * nn + a]]++] = (int32_t)e;
  }
  P.n_adj_elem.assign(n_nodes, 0);
  P.n_adj_diag_elem.assign(n_nodes, 0);
  for (int64_t n = 0; n < n_nodes; ++n)
    for (int64_t k = P.n2e_ptr[n]; k < P.n2e_ptr[n + 1]; ++k) {
      if (diag_elem && diag_elem[P.n2e[k]]) P.n_adj_diag_elem[n]++;
      else P.n_adj_elem[n]++;
    }

  // ---- owned rows and node-level sparsity ----------------------------------------------------
  P.node_row.assign(n_nodes, -1);
  for (int64_t n = 0; n < n_nodes; ++n)
    if (!owned || owned[n]) { P.node_row[n] = (int32_t)P.owned_nodes.size(); P.owned_nodes.push_back((int32_t)n); }
  P.n_owned = (int64_t)P.owned_nodes.size();
  if (P.n_owned == 0) return "no owned nodes";
  P.nbr_ptr.assign(P.n_owned + 1, 0);
  // pass 1: counts; pass 2: fill (both parallel over owned rows)
  std::vector<int32_t> cnt(P.n_owned);
  auto gather = [&](int64_t n, std::vector<int32_t>& l) {
    l.clear();
    for (int64_t k = P.n2e_ptr[n]; k < P.n2e_ptr[n + 1]; ++k) {
      const int32_t* c = conn + (int64_t)P.n2e[k] * nn;
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
    if (cnt[r] > 4000) return "node with more than 4000 neighbours";
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
  return "";
}

// ================================================================================================================
// streaming plan
// ================================================================================================================
namespace {

struct Range { int64_t b, e; };

struct TileOut {
  std::vector<StepDesc> steps;
  std::vector<VisitRec8> visits;
  std::vector<FlushRow> frows;
  std::vector<uint32_t> fent;
  std::vector<LoadRec> loads;
  int32_t n_owned = 0, max_blocks = 0, max_slots = 0, n_real_visits = 0;
};

struct Scratch {
  std::vector<int32_t> local;  // global node -> tile-local id or -1 (dense, reset after every tile)
};

struct Builder {
  const Pattern& pat;
  const double* xyz;
  const int32_t* conn;
  const uint8_t* diag;
  PlanLimits lim;
  int sweep = 2, cu = 0, cv = 1;  // sweep axis and the two cross axes
  double h[3] = {1., 1., 1.};     // typical node spacing per axis
  std::string error;              // hard (non-capacity) failure
  std::atomic<int> why{0};        // which limit the last oversized tile hit (1 slots, 2 arena, 3 flush buffer)
  std::mutex mu;

  Builder(const Pattern& p, const double* x, const int32_t* c, const uint8_t* d, const PlanLimits& l)
      : pat(p), xyz(x), conn(c), diag(d), lim(l) {}

  void fail(const std::string& s) {
    std::lock_guard<std::mutex> g(mu);
    if (error.empty()) error = s;
  }

  // ---- one tile: owned nodes `own` (sorted by (front, gid)) with front index `fr` in [0, n_fronts) --------------
  // returns false if the tile does not fit the shared-memory limits (caller splits the strip)
  bool build_tile(const int32_t* own, const int32_t* fr, int32_t n_owned, int32_t n_fronts, Scratch& sc, TileOut& T) {
    constexpr int nn = 8;
    T = TileOut();
    T.n_owned = n_owned;
    std::vector<int32_t>& local = sc.local;
    std::vector<int32_t> nodes(own, own + n_owned);  // tile-local -> gid
    for (int32_t l = 0; l < n_owned; ++l) local[own[l]] = l;
    auto cleanup = [&] { for (int32_t n : nodes) local[n] = -1; };

    // elements of the tile: every non-diagonal element touching an owned node
    std::vector<int32_t> elems;
    for (int32_t l = 0; l < n_owned; ++l) {
      const int32_t n = own[l];
      for (int64_t k = pat.n2e_ptr[n]; k < pat.n2e_ptr[n + 1]; ++k)
        if (!(diag && diag[pat.n2e[k]])) elems.push_back(pat.n2e[k]);
    }
    std::sort(elems.begin(), elems.end());
    elems.erase(std::unique(elems.begin(), elems.end()), elems.end());
    const size_t nvis = elems.size();
    // step of a visit: the last front among its owned nodes; an element whose owned nodes all lie in one front and
    // which sits beyond that front along the sweep goes one step later, so that every step completes one front only
    std::vector<int32_t> vstep(nvis);
    int32_t S = n_fronts;
    for (size_t v = 0; v < nvis; ++v) {
      const int32_t* c = conn + (int64_t)elems[v] * nn;
      int32_t smin = INT_MAX, smax = 0, no = 0;
      double cen = 0., own_c = 0.;
      for (int a = 0; a < nn; ++a) {
        const double x = xyz[(int64_t)c[a] * 3 + sweep];
        cen += x;
        const int32_t l = local[c[a]];
        if (l >= 0) { smin = std::min(smin, fr[l]); smax = std::max(smax, fr[l]); own_c += x; ++no; }
      }
      if (smin == smax && cen / nn > own_c / no) ++smax;
      vstep[v] = smax;
      S = std::max(S, smax + 1);
    }
    // Static lock-step schedule: the visits of a step are laid out on a grid of `n_workers`
    // compute warps x K positions.  Visits at the same position share no owned node, so they never touch the same
    // pair block; a visit only has to wait for visits at earlier positions, which the other warps -- walking their own
    // rows of the grid in step -- have normally finished.  Accumulation order: (step, position, worker).
    const int32_t NW = lim.n_workers;
    std::vector<int32_t> order(nvis), colour(nvis, 0), worker(nvis, 0), n_pos(S, 0);
    {
      std::vector<int32_t> idx(nvis);
      for (size_t v = 0; v < nvis; ++v) idx[v] = (int32_t)v;
      std::stable_sort(idx.begin(), idx.end(), [&](int32_t a, int32_t b) { return vstep[a] < vstep[b]; });
      std::vector<uint64_t> used(n_owned, 0ull);
      std::vector<int32_t> touched, cls(nvis, 0);
      auto owned_mask = [&](int32_t v) {
        const int32_t* c = conn + (int64_t)elems[v] * nn;
        uint64_t m = 0;
        for (int a = 0; a < nn; ++a) { const int32_t l = local[c[a]]; if (l >= 0) m |= used[l]; }
        return m;
      };
      auto mark = [&](int32_t v, int bit) {
        const int32_t* c = conn + (int64_t)elems[v] * nn;
        for (int a = 0; a < nn; ++a) { const int32_t l = local[c[a]]; if (l >= 0) { used[l] |= 1ull << bit; touched.push_back(l); } }
      };
      size_t i0 = 0;
      while (i0 < nvis) {
        size_t i1 = i0;
        const int32_t st = vstep[idx[i0]];
        while (i1 < nvis && vstep[idx[i1]] == st) ++i1;
        // (1) proper colouring of the step's conflict graph (visits sharing an owned node), greedy in element order
        touched.clear();
        std::vector<std::vector<int32_t>> classes;
        for (size_t i = i0; i < i1; ++i) {
          const uint64_t m = owned_mask(idx[i]);
          int32_t col = 0;
          while (col < 64 && (m >> col & 1ull)) ++col;
          if (col >= 64) { cleanup(); why = 4; return false; }  // a node with more than 64 elements in one step
          if ((size_t)col >= classes.size()) classes.resize(col + 1);
          classes[col].push_back(idx[i]);
          mark(idx[i], col);
        }
        for (int32_t l : touched) used[l] = 0ull;
        touched.clear();
        // (2) positions: full groups of NW visits of one class; the remainders of all classes are packed together
        // into extra positions (taken alternately from the two ends of the classes, i.e. from far-apart regions)
        int32_t K = 0;
        std::vector<int32_t> overflow;
        int flip = 0;
        for (auto& cl : classes) {
          const int32_t n = (int32_t)cl.size();
          int32_t rem = n % NW;
          if (rem * 4 >= NW * 3) rem = 0;  // nearly a full position: keep it whole
          const int32_t b = (flip & 1) ? rem : 0, e = (flip & 1) ? n : n - rem;
          if (rem) ++flip;
          for (int32_t k = 0; k < n; ++k)
            if (k < b || k >= e) overflow.push_back(cl[k]);
          for (int32_t k = b; k < e; ++k) { colour[cl[k]] = K + (k - b) / NW; worker[cl[k]] = (k - b) % NW; }
          K += (e - b + NW - 1) / NW;
        }
        {
          std::vector<int32_t> count(64, 0);
          const int32_t K0 = K;
          for (int32_t v : overflow) {
            const uint64_t m = owned_mask(v);
            int32_t col = 0;
            while (col < 64 && ((m >> col & 1ull) || count[col] >= NW)) ++col;
            if (K0 + col >= 64 || col >= 64) { cleanup(); why = 4; return false; }
            colour[v] = K0 + col; worker[v] = count[col]++;
            K = std::max(K, K0 + col + 1);
            mark(v, col);
          }
          for (int32_t l : touched) used[l] = 0ull;
        }
        n_pos[st] = K;
        std::stable_sort(idx.begin() + i0, idx.begin() + i1, [&](int32_t a, int32_t b) {
          return colour[a] != colour[b] ? colour[a] < colour[b] : worker[a] < worker[b];
        });
        i0 = i1;
      }
      order.swap(idx);
    }
    // halo nodes: tile-local ids in first-touch order
    for (size_t o = 0; o < nvis; ++o) {
      const int32_t* c = conn + (int64_t)elems[order[o]] * nn;
      for (int a = 0; a < nn; ++a)
        if (local[c[a]] < 0) { local[c[a]] = (int32_t)nodes.size(); nodes.push_back(c[a]); }
    }
    const int32_t n_tile_nodes = (int32_t)nodes.size();

    // lifetimes
    std::vector<int32_t> first(n_tile_nodes, INT_MAX), last(n_tile_nodes, -1), flush_step(n_owned), state_end(n_tile_nodes);
    for (size_t o = 0; o < nvis; ++o) {
      const int32_t s = vstep[order[o]];
      const int32_t* c = conn + (int64_t)elems[order[o]] * nn;
      for (int a = 0; a < nn; ++a) {
        const int32_t l = local[c[a]];
        first[l] = std::min(first[l], s); last[l] = std::max(last[l], s);
      }
    }
    for (int32_t l = 0; l < n_owned; ++l) {
      if (last[l] < 0) { first[l] = fr[l]; last[l] = fr[l]; }  // owned node without (non-diagonal) elements
      // rows of front f are written after step f+1 even where the mesh ends, so that no step writes two fronts
      flush_step[l] = std::max(last[l], fr[l] + 1);
      S = std::max(S, flush_step[l] + 1);
    }
    for (int32_t l = 0; l < n_tile_nodes; ++l) state_end[l] = l < n_owned ? std::max(last[l], flush_step[l]) : last[l];
    for (size_t o = 0; o < nvis; ++o) {
      const int32_t* c = conn + (int64_t)elems[order[o]] * nn;
      int32_t m = -1;
      for (int a = 0; a < nn; ++a) { const int32_t l = local[c[a]]; if (l < n_owned) m = std::max(m, flush_step[l]); }
      for (int a = 0; a < nn; ++a) { const int32_t l = local[c[a]]; state_end[l] = std::max(state_end[l], m); }
    }

    // ---- state slots: a slot last read in step t (visit or flush) may be reloaded for a node first used at step >= t+3
    // (the state of step s+2 is staged while step s is being visited and step s-1 flushed)
    std::vector<int32_t> slot(n_tile_nodes, -1);
    {
      std::vector<std::vector<int32_t>> by_first(S), release(S + 3);
      for (int32_t l = 0; l < n_tile_nodes; ++l) by_first[first[l]].push_back(l);
      std::vector<int32_t> free_slots;
      int32_t next_new = 0, live = 0;
      for (int32_t s = 0; s < S; ++s) {
        for (int32_t sl : release[s]) { free_slots.push_back(sl); --live; }
        for (int32_t l : by_first[s]) {
          int32_t sl;
          if (!free_slots.empty()) { sl = free_slots.back(); free_slots.pop_back(); }
          else if (next_new < lim.max_slots) sl = next_new++;
          else { cleanup(); why = 1; return false; }
          slot[l] = sl;
          ++live;
          T.max_slots = std::max(T.max_slots, live);
          release[std::min(state_end[l] + 3, S + 2)].push_back(sl);  // staged two steps ahead of its first visit
        }
      }
    }

    // ---- pair blocks ------------------------------------------------------------------------------------------
    std::vector<int64_t> base(n_owned + 1, 0);  // block index space: (owned node, neighbour position)
    for (int32_t l = 0; l < n_owned; ++l) {
      const int32_t row = pat.node_row[nodes[l]];
      base[l + 1] = base[l] + (pat.nbr_ptr[row + 1] - pat.nbr_ptr[row]);
    }
    auto nbr_pos = [&](int32_t l, int32_t gid) {  // position of gid in the neighbour list of owned tile node l
      const int32_t row = pat.node_row[nodes[l]];
      const int32_t* nb = pat.nbr.data() + pat.nbr_ptr[row];
      const int32_t* ne = pat.nbr.data() + pat.nbr_ptr[row + 1];
      return (int32_t)(std::lower_bound(nb, ne, gid) - nb);
    };
    std::vector<int32_t> blk_slot(base[n_owned], -1);
    n_pos.resize(S, 0);  // S may have grown by the trailing flush-only step
    T.steps.assign(S, StepDesc{0, 0, 0, 0, 0, 0, 0, 0, 0, {0, 0, 0}});
    {
      int32_t g = 0;
      for (int32_t st = 0; st < S; ++st) { T.steps[st].visit_begin = g; T.steps[st].n_pos = n_pos[st]; g += NW * n_pos[st]; }
      VisitRec8 empty;
      std::memset(&empty, 0xFF, sizeof empty);  // ln = 0xFFFF: no visit at this grid position
      T.visits.assign((size_t)g, empty);
    }
    T.n_real_visits = (int32_t)nvis;
    {
      std::vector<std::vector<int32_t>> release(S + 3);
      std::vector<int32_t> free_blocks;
      int32_t next_new = 0, live = 0;
      std::vector<uint32_t> rank(n_owned, 0u);
      int32_t cur_step = -1;
      for (size_t o = 0; o < nvis; ++o) {
        const int32_t s = vstep[order[o]];
        while (cur_step < s) {
          ++cur_step;
          for (int32_t b : release[cur_step]) { free_blocks.push_back(b); --live; }
        }
        T.steps[s].n_visits++;
        const int32_t* c = conn + (int64_t)elems[order[o]] * nn;
        VisitRec8& V = T.visits[(size_t)T.steps[s].visit_begin + (size_t)worker[order[o]] * n_pos[s] + colour[order[o]]];
        int32_t ll[8];
        for (int a = 0; a < nn; ++a) {
          ll[a] = local[c[a]];
          V.ln[a] = (uint16_t)slot[ll[a]];
          if (ll[a] < n_owned) {
            if (rank[ll[a]] > 250) { cleanup(); fail("node touched by more than 250 elements"); return true; }
            V.rank[a] = (uint8_t)rank[ll[a]]++;
          } else {
            V.rank[a] = 0xFF;
          }
        }
        for (int a = 0; a < nn; ++a)
          for (int b = a; b < nn; ++b) {
            uint16_t& out = V.pair[a * (15 - a) / 2 + b];
            const bool oa = ll[a] < n_owned, ob = ll[b] < n_owned;
            if (!oa && !ob) { out = kPairNone; continue; }
            // the block is stored as S_(P,Q): P the owned endpoint, the lower global id if both are owned
            const bool p_is_a = (oa && ob) ? (c[a] <= c[b]) : oa;
            const int p = p_is_a ? a : b, q = p_is_a ? b : a;
            const int64_t blk = base[ll[p]] + nbr_pos(ll[p], c[q]);
            uint16_t e = p_is_a ? 0 : kPairTransposed;
            if (blk_slot[blk] < 0) {
              int32_t bs;
              if (!free_blocks.empty()) { bs = free_blocks.back(); free_blocks.pop_back(); }
              else if (next_new < lim.arena_blocks) bs = next_new++;
              else { cleanup(); why = 2; return false; }
              blk_slot[blk] = bs;
              ++live;
              T.max_blocks = std::max(T.max_blocks, live);
              int32_t end = flush_step[ll[p]];
              if (ll[q] < n_owned) end = std::max(end, flush_step[ll[q]]);
              release[end + 2].push_back(bs);  // the flush of step `end` runs while step end+1 is being visited
              e |= kPairInit;
            }
            out = (uint16_t)(e | (uint16_t)blk_slot[blk]);
          }
      }
    }

    // ---- loads and flush rows per step ------------------------------------------------------------------------------
    {
      std::vector<std::vector<int32_t>> by_first(S), by_flush(S);
      for (int32_t l = 0; l < n_tile_nodes; ++l) by_first[first[l]].push_back(l);
      for (int32_t l = 0; l < n_owned; ++l) by_flush[flush_step[l]].push_back(l);
      for (int32_t s = 0; s < S; ++s) {
        StepDesc& sd = T.steps[s];
        sd.load_begin = (int32_t)T.loads.size();
        std::sort(by_first[s].begin(), by_first[s].end(), [&](int32_t a, int32_t b) { return nodes[a] < nodes[b]; });
        for (int32_t l : by_first[s]) T.loads.push_back(LoadRec{nodes[l], (uint16_t)slot[l], (uint16_t)(l < n_owned ? 1 : 0)});
        sd.n_loads = (int32_t)by_first[s].size();
        sd.flush_begin = (int32_t)T.frows.size();
        while (T.fent.size() % 4) T.fent.push_back(0u);  // 16-byte aligned per step: the device copies uint4
        sd.fent_begin = (int32_t)T.fent.size();
        std::sort(by_flush[s].begin(), by_flush[s].end(), [&](int32_t a, int32_t b) { return nodes[a] < nodes[b]; });
        for (int32_t l : by_flush[s]) {
          const int32_t ga = nodes[l];
          const int32_t row = pat.node_row[ga];
          const int64_t nb0 = pat.nbr_ptr[row], nb1 = pat.nbr_ptr[row + 1];
          FlushRow fr_;
          std::memset(&fr_, 0, sizeof fr_);
          fr_.dst_block = nb0; fr_.gid = ga; fr_.ent_rel = (int32_t)T.fent.size() - sd.fent_begin;
          fr_.slot = (uint16_t)slot[l]; fr_.nnb = (uint16_t)(nb1 - nb0);
          fr_.nadj = (uint16_t)std::min(pat.n_adj_elem[ga], 65535); fr_.nadjd = (uint16_t)std::min(pat.n_adj_diag_elem[ga], 65535);
          T.frows.push_back(fr_);
          for (int64_t k = nb0; k < nb1; ++k) {
            const int32_t gq = pat.nbr[k];
            const int32_t lq = local[gq];
            uint32_t fl = (gq == ga) ? kFlushSelf : 0u, bsl = kFlushNoBlock, st = 0xFFFFu;
            if (lq >= 0) {
              int64_t blk;
              if (lq < n_owned && gq < ga) { blk = base[lq] + nbr_pos(lq, ga); fl |= kFlushTransposed; }
              else blk = base[l] + (k - nb0);
              if (blk_slot[blk] >= 0) { bsl = (uint32_t)blk_slot[blk]; st = (uint32_t)slot[lq]; }
            }
            if (gq == ga) st = (uint32_t)slot[l];  // the row's own state is always staged
            const uint32_t e = bsl | fl | (st << 16);
            T.fent.push_back(e);
          }
        }
        sd.n_flush = (int32_t)by_flush[s].size();
        sd.n_fent = (int32_t)T.fent.size() - sd.fent_begin;
        if ((int64_t)sd.n_flush * (int64_t)sizeof(FlushRow) + (int64_t)(sd.n_fent + 3) / 4 * 16 > lim.flush_buf_bytes) { cleanup(); why = 3; return false; }
      }
      while (T.fent.size() % 4) T.fent.push_back(0u);
    }
    cleanup();
    return true;
  }

  // ---- strip handling ---------------------------------------------------------------------------------------------
  double coord(int32_t n, int ax) const { return xyz[(int64_t)n * 3 + ax]; }

  // splits perm[r.b, r.e) along axis ax at about fraction num/den of the nodes, snapping to the widest coordinate gap
  // within a window of about one node plane so that strips of (possibly jittered) structured meshes end on whole planes
  int64_t split(std::vector<int32_t>& perm, Range r, int ax, int64_t num, int64_t den, double planes) const {
    const int64_t cnt = r.e - r.b;
    const int64_t pos = r.b + std::max<int64_t>(1, std::min<int64_t>(cnt - 1, cnt * num / den));
    auto key = [&](int32_t n) { return coord(n, ax); };
    auto less = [&](int32_t a, int32_t b) { return key(a) < key(b) || (key(a) == key(b) && a < b); };
    int64_t W = (int64_t)std::ceil((double)cnt / std::max(1.0, planes)) + 1;
    W = std::min(W, std::min(pos - r.b, r.e - pos) - 1);
    int64_t cut = pos;
    if (W >= 1) {
      auto b0 = perm.begin() + (pos - W), b1 = perm.begin() + (pos + W);
      std::nth_element(perm.begin() + r.b, b0, perm.begin() + r.e, less);
      std::nth_element(b0, b1, perm.begin() + r.e, less);
      std::sort(b0, b1, less);
      // prefer the plane boundary (gap of more than half a node spacing) closest to pos; else the widest gap
      double best = -1.;
      int64_t best_k = pos, near_k = -1;
      for (int64_t k = pos - W + 1; k < pos + W; ++k) {
        const double gap = key(perm[k]) - key(perm[k - 1]);
        if (gap > 0.5 * h[ax] && (near_k < 0 || std::llabs(k - pos) < std::llabs(near_k - pos))) near_k = k;
        if (gap > best) { best = gap; best_k = k; }
      }
      cut = near_k >= 0 ? near_k : best_k;
    } else {
      std::nth_element(perm.begin() + r.b, perm.begin() + pos, perm.begin() + r.e, less);
    }
    return cut;
  }

  void extent(const std::vector<int32_t>& perm, Range r, double lo[3], double hi[3]) const {
    for (int c = 0; c < 3; ++c) { lo[c] = 1e300; hi[c] = -1e300; }
    for (int64_t k = r.b; k < r.e; ++k)
      for (int c = 0; c < 3; ++c) {
        const double v = coord(perm[k], c);
        lo[c] = std::min(lo[c], v); hi[c] = std::max(hi[c], v);
      }
  }
  double planes_of(double lo, double hi, int ax) const { return std::floor((hi - lo) / h[ax] + 0.5) + 1.; }
};

}  // namespace

std::string build_stream_plan(const Pattern& pat, const double* xyz, const int32_t* conn, const uint8_t* diag_elem,
                              const PlanLimits& lim, StreamPlan& plan) {
  plan = StreamPlan();
  plan.lim = lim;
  if (pat.nn != 8) return "streaming plan: HEX8 only";
  if (lim.arena_blocks < 64 || lim.arena_blocks > (int)kFlushNoBlock - 1) return "arena size out of range";
  if (lim.n_workers < 1 || lim.n_workers > 64) return "worker count out of range";
  Builder B(pat, xyz, conn, diag_elem, lim);
  constexpr int nn = 8;

  // ---- typical node spacing per axis from element edges (every edge is attributed to its dominant axis) ----------
  {
    static const int EDGE[12][2] = {{0, 1}, {3, 2}, {4, 5}, {7, 6}, {0, 3}, {1, 2}, {4, 7}, {5, 6}, {0, 4}, {1, 5}, {2, 6}, {3, 7}};
    double sum[3] = {0., 0., 0.};
    int64_t num[3] = {0, 0, 0};
    const int64_t stride = std::max<int64_t>(1, pat.n_elem / 50000);
    for (int64_t e = 0; e < pat.n_elem; e += stride) {
      const int32_t* c = conn + e * nn;
      for (auto& ed : EDGE) {
        double d[3];
        for (int k = 0; k < 3; ++k) d[k] = std::fabs(xyz[(int64_t)c[ed[1]] * 3 + k] - xyz[(int64_t)c[ed[0]] * 3 + k]);
        const int ax = d[0] >= d[1] ? (d[0] >= d[2] ? 0 : 2) : (d[1] >= d[2] ? 1 : 2);
        sum[ax] += d[ax]; num[ax]++;
      }
    }
    double all = 0.; int64_t nall = 0;
    for (int k = 0; k < 3; ++k) { all += sum[k]; nall += num[k]; }
    for (int k = 0; k < 3; ++k) B.h[k] = num[k] ? sum[k] / num[k] : (nall ? all / nall : 1.);
    for (int k = 0; k < 3; ++k) if (!(B.h[k] > 0.)) B.h[k] = 1.;
  }

  std::vector<int32_t> perm(pat.owned_nodes);
  double lo[3], hi[3];
  B.extent(perm, {0, pat.n_owned}, lo, hi);
  double pl[3];
  for (int k = 0; k < 3; ++k) pl[k] = B.planes_of(lo[k], hi[k], k);
  // sweep along the axis with most node planes (ties: the highest axis, whose planes are contiguous in libMesh numbering)
  B.sweep = 2;
  for (int k = 1; k >= 0; --k) if (pl[k] > pl[B.sweep]) B.sweep = k;
  B.cu = (B.sweep + 1) % 3; B.cv = (B.sweep + 2) % 3;
  if (B.cu > B.cv) std::swap(B.cu, B.cv);

  // cross-section target: the largest mu x mv node columns whose live pair blocks fit the arena.  On a hexahedral grid
  // a strip holds three fronts of in-plane pairs and three layers of cross-plane pairs at a time: two being
  // accumulated / waiting for their rows, one being flushed behind the visits.
  int mu = 1, mv = 1;
  {
    auto blocks = [](int a, int b) {
      const int both = (3 * a - 2) * (3 * b - 2), halo = 9 * a * b - both;
      return 3 * ((both + a * b) / 2 + halo) + 3 * (9 * a * b + halo);
    };
    const double cap = 1.03 * lim.arena_blocks;
    for (int a = 1; a <= 24; ++a)
      for (int b = a; b <= 24; ++b) {
        if (blocks(a, b) > cap || (a + 2) * (b + 2) * 5 > lim.max_slots) continue;
        if (a * b > mu * mv || (a * b == mu * mv && b - a < mv - mu)) { mu = a; mv = b; }
      }
    if (const char* ev = std::getenv("MASTB200_PLAN_COLS")) {
      const int cols = std::max(1, std::atoi(ev));
      mu = std::max(1, (int)std::floor(std::sqrt((double)cols)));
      mv = std::max(1, cols / mu);
    }
  }
  const int cols = mu * mv;
  // ---- cross partition into strips ----------------------------------------------------------------------------------
  std::vector<Range> strips, work;
  work.push_back({0, pat.n_owned});
  auto cross_split = [&](Range r, bool force, std::vector<Range>& out_leaf, std::vector<Range>& out_work) {
    double l[3], hh[3];
    B.extent(perm, r, l, hh);
    const double pu = B.planes_of(l[B.cu], hh[B.cu], B.cu), pv = B.planes_of(l[B.cv], hh[B.cv], B.cv);
    if (r.e - r.b <= 1 || (!force && pu <= mu && pv <= mv)) { out_leaf.push_back(r); return; }
    int ax; double p; int m;
    if (pu / mu >= pv / mv) { ax = B.cu; p = pu; m = mu; } else { ax = B.cv; p = pv; m = mv; }
    int64_t k = std::max<int64_t>(2, (int64_t)std::ceil(p / m));
    if (force && p <= 1.5) {  // a single node plane left in both cross directions: split by count
      ax = (hh[B.cu] - l[B.cu]) / B.h[B.cu] >= (hh[B.cv] - l[B.cv]) / B.h[B.cv] ? B.cu : B.cv;
      k = 2;
    }
    const int64_t cut = B.split(perm, r, ax, k / 2, k, p);
    out_work.push_back({r.b, cut});
    out_work.push_back({cut, r.e});
  };
  while (!work.empty()) {
    Range r = work.back();
    work.pop_back();
    cross_split(r, false, strips, work);
  }

  // ---- segment length along the sweep -----------------------------------------------------------------------------------
  const double total_fronts = (double)strips.size() * pl[B.sweep];
  const int seg_len = (int)std::max(3.0, std::min(128.0, std::floor(total_fronts / (6.0 * std::max(1, lim.sm_count)))));

  // ---- build the tiles of every strip (parallel over strips; a strip that does not fit is split and retried) -----------
  std::vector<std::vector<TileOut>> built;   // per finished strip
  const double gap_thresh = 0.5 * B.h[B.sweep];
  const int max_front = std::max(8, 2 * cols);
  int guard = 0;
  while (!strips.empty() && B.error.empty()) {
    if (++guard > 64) return "plan: strips keep failing to fit shared memory";
    std::sort(strips.begin(), strips.end(), [](const Range& a, const Range& b) { return a.b < b.b; });
    const size_t ns = strips.size();
    std::vector<std::vector<TileOut>> out(ns);
    std::vector<char> too_big(ns, 0);
    std::vector<Scratch> scratch(hw_threads());
    parallel_items((int64_t)ns, [&](int64_t si, int th) {
      Scratch& sc = scratch[th];
      if (sc.local.empty()) sc.local.assign(pat.n_nodes, -1);
      const Range r = strips[si];
      const int ax = B.sweep;
      std::sort(perm.begin() + r.b, perm.begin() + r.e, [&](int32_t a, int32_t b) {
        const double ka = B.coord(a, ax), kb = B.coord(b, ax);
        return ka < kb || (ka == kb && a < b);
      });
      // fronts: cut at coordinate gaps of more than half a node spacing, or when a front gets too large
      const int64_t cnt = r.e - r.b;
      std::vector<int32_t> front(cnt);
      int32_t f = 0, fsz = 0;
      for (int64_t k = 0; k < cnt; ++k) {
        if (k > 0 && (B.coord(perm[r.b + k], ax) - B.coord(perm[r.b + k - 1], ax) > gap_thresh || fsz >= max_front)) { ++f; fsz = 0; }
        front[k] = f; ++fsz;
      }
      const int32_t n_fronts = f + 1;
      const int32_t n_seg = std::max(1, (n_fronts + seg_len / 2) / seg_len);
      std::vector<int32_t> own, fr;
      int64_t k0 = 0;
      for (int32_t sg = 0; sg < n_seg && !too_big[si]; ++sg) {
        const int32_t f0 = (int32_t)((int64_t)n_fronts * sg / n_seg), f1 = (int32_t)((int64_t)n_fronts * (sg + 1) / n_seg);
        int64_t k1 = k0;
        while (k1 < cnt && front[k1] < f1) ++k1;
        // nodes sorted by (front, gid)
        std::vector<std::pair<int32_t, int32_t>> tmp;
        tmp.reserve(k1 - k0);
        for (int64_t k = k0; k < k1; ++k) tmp.emplace_back(front[k] - f0, perm[r.b + k]);
        std::sort(tmp.begin(), tmp.end());
        own.resize(tmp.size()); fr.resize(tmp.size());
        for (size_t i = 0; i < tmp.size(); ++i) { fr[i] = tmp[i].first; own[i] = tmp[i].second; }
        out[si].emplace_back();
        if (!B.build_tile(own.data(), fr.data(), (int32_t)own.size(), f1 - f0, sc, out[si].back())) too_big[si] = 1;
        k0 = k1;
      }
    });
    if (!B.error.empty()) break;
    if (std::getenv("MASTB200_PLAN_DEBUG")) {
      size_t nf = 0;
      for (char c : too_big) nf += c;
      std::fprintf(stderr, "[plan] round %d: %zu strips, %zu too big; sweep axis %d, h = %g %g %g, target %dx%d, seg_len %d, limit hit %d\n",
                   guard, ns, nf, B.sweep, B.h[0], B.h[1], B.h[2], mu, mv, seg_len, (int)B.why);
    }
    std::vector<Range> next;
    for (size_t si = 0; si < ns; ++si) {
      if (!too_big[si]) { built.push_back(std::move(out[si])); continue; }
      if (strips[si].e - strips[si].b <= 1) return "plan: a single node does not fit the shared-memory arena";
      std::vector<Range> leaf;
      cross_split(strips[si], true, leaf, next);
      for (const Range& l : leaf) next.push_back(l);  // cannot happen for ranges of more than one node
    }
    strips.swap(next);
  }
  if (!B.error.empty()) return B.error;

  // ---- concatenate; tiles ordered by descending cost for the dynamic scheduler -----------------------------------------
  std::vector<TileOut*> all;
  for (auto& s : built) for (auto& t : s) all.push_back(&t);
  std::stable_sort(all.begin(), all.end(), [](const TileOut* a, const TileOut* b) { return a->n_real_visits > b->n_real_visits; });
  size_t n_steps = 0, n_vis = 0, n_fr = 0, n_fe = 0, n_ld = 0;
  for (const TileOut* t : all) {
    n_steps += t->steps.size(); n_vis += t->visits.size(); n_fr += t->frows.size(); n_fe += t->fent.size(); n_ld += t->loads.size();
  }
  const size_t lim31 = size_t(1) << 31;
  if (n_steps >= lim31 || n_vis >= lim31 || n_fr >= lim31 || n_fe >= lim31 || n_ld >= lim31) return "plan too large";
  plan.tiles.reserve(all.size()); plan.steps.reserve(n_steps); plan.visits.reserve(n_vis);
  plan.flush_rows.reserve(n_fr); plan.flush_ent.reserve(n_fe); plan.loads.reserve(n_ld);
  for (TileOut* t : all) {
    TileDesc td;
    td.step_begin = (int32_t)plan.steps.size(); td.n_steps = (int32_t)t->steps.size();
    td.n_visits = t->n_real_visits; td.n_owned = t->n_owned;
    const int32_t v0 = (int32_t)plan.visits.size(), l0 = (int32_t)plan.loads.size(), f0 = (int32_t)plan.flush_rows.size(),
                  e0 = (int32_t)plan.flush_ent.size();
    for (StepDesc sd : t->steps) {
      sd.visit_begin += v0; sd.load_begin += l0; sd.flush_begin += f0; sd.fent_begin += e0;
      plan.steps.push_back(sd);
    }
    plan.visits.insert(plan.visits.end(), t->visits.begin(), t->visits.end());
    plan.flush_rows.insert(plan.flush_rows.end(), t->frows.begin(), t->frows.end());
    plan.flush_ent.insert(plan.flush_ent.end(), t->fent.begin(), t->fent.end());
    plan.loads.insert(plan.loads.end(), t->loads.begin(), t->loads.end());
    plan.max_live_blocks = std::max(plan.max_live_blocks, t->max_blocks);
    plan.max_live_slots = std::max(plan.max_live_slots, t->max_slots);
    plan.tiles.push_back(td);
    *t = TileOut();  // release memory early
  }
  return "";
}

// ================================================================================================================
// coloured scatter plan (HEX27)
// ================================================================================================================
std::string build_coloring(const Pattern& pat, const int32_t* conn, std::vector<int32_t>& color_ptr,
                           std::vector<int32_t>& color_elems) {
  const int nn = pat.nn;
  std::vector<uint64_t> used(pat.n_nodes, 0);  // colours taken by the elements around each node
  std::vector<int8_t> color(pat.n_elem, -1);
  int n_colors = 0;
  for (int64_t e = 0; e < pat.n_elem; ++e) {
    const int32_t* c = conn + e * nn;
    bool any_owned = false;
    uint64_t m = 0;
    for (int a = 0; a < nn; ++a) { m |= used[c[a]]; any_owned = any_owned || pat.node_row[c[a]] >= 0; }
    if (!any_owned) continue;
    if (m == ~uint64_t(0)) return "more than 64 colours needed";
    int k = 0;
    while ((m >> k) & 1u) ++k;
    color[e] = (int8_t)k;
    n_colors = std::max(n_colors, k + 1);
    for (int a = 0; a < nn; ++a) used[c[a]] |= uint64_t(1) << k;
  }
  color_ptr.assign(n_colors + 1, 0);
  for (int64_t e = 0; e < pat.n_elem; ++e)
    if (color[e] >= 0) color_ptr[color[e] + 1]++;
  for (int k = 0; k < n_colors; ++k) color_ptr[k + 1] += color_ptr[k];
  color_elems.resize(color_ptr[n_colors]);
  std::vector<int32_t> pos(color_ptr.begin(), color_ptr.end() - 1);
  for (int64_t e = 0; e < pat.n_elem; ++e)
    if (color[e] >= 0) color_elems[pos[color[e]]++] = (int32_t)e;
  return "";
}

std::string build_posmap(const Pattern& pat, const int32_t* conn, std::vector<uint8_t>& posmap) {
  const int nn = pat.nn;
  for (int64_t r = 0; r < pat.n_owned; ++r)
    if (pat.nbr_ptr[r + 1] - pat.nbr_ptr[r] > 255) return "row with more than 255 neighbouring nodes";
  posmap.assign((size_t)pat.n_elem * nn * nn, 0);
  parallel_for(pat.n_elem, [&](int64_t b0, int64_t e0, int) {
    for (int64_t e = b0; e < e0; ++e) {
      const int32_t* c = conn + e * nn;
      for (int a = 0; a < nn; ++a) {
        const int32_t r = pat.node_row[c[a]];
        if (r < 0) continue;
        const int32_t* lb = pat.nbr.data() + pat.nbr_ptr[r];
        const int32_t* le = pat.nbr.data() + pat.nbr_ptr[r + 1];
        for (int b = 0; b < nn; ++b)
          posmap[((size_t)e * nn + a) * nn + b] = (uint8_t)(std::lower_bound(lb, le, c[b]) - lb);
      }
    }
  });
  return "";
}

// Processing order of the single-launch HEX27 scatter (sweep_hex27.cu): elements grouped into spatial bricks (so that the
// CSR rows a brick adds into stay in L2 while its elements run), colour-major inside a brick (elements of one colour share
// no node).  An element may start when every element that shares a node with it and comes EARLIER in the order has
// finished: pred lists those.  The order fixes the summation order of every CSR value and residual entry (independent of
// scheduling) and is deadlock-free for CTAs that take the list round-robin, each in increasing position.
// first_flag[(e, a, b)]: this element is the first in the order to touch block (a, b) of its row -> it stores, the later
// ones add, and J needs no memset.
// Sets bit 7 of the position-map entry of the FIRST element (in the order given: group by group, elements of a group in
// any order) that touches each CSR block, so that a scatter in that order can store the first contribution instead of adding
// to a cleared value.  Needs < 128 neighbours per row and no diagonal elements; returns false (map untouched) otherwise.
// With parallel_groups the elements of a group are known not to share nodes (colours): they are flagged concurrently.
bool flag_first_writers(const Pattern& pat, const int32_t* conn, const uint8_t* diag_elem, const std::vector<int32_t>& group_ptr,
                        const std::vector<int32_t>& elems, bool parallel_groups, std::vector<uint8_t>& posmap) {
  const int nn = pat.nn;
  for (int64_t r = 0; r < pat.n_owned; ++r)
    if (pat.nbr_ptr[r + 1] - pat.nbr_ptr[r] > 127) return false;
  if (diag_elem)
    for (int64_t e = 0; e < pat.n_elem; ++e)
      if (diag_elem[e]) return false;
  std::vector<uint64_t> seen(((size_t)pat.nbr_ptr[pat.n_owned] + 63) / 64, 0);
  auto flag = [&](int64_t e) {
    const int32_t* c = conn + e * nn;
    for (int a = 0; a < nn; ++a) {
      const int32_t r = pat.node_row[c[a]];
      if (r < 0) continue;
      for (int b = 0; b < nn; ++b) {
        uint8_t& pm = posmap[((size_t)e * nn + a) * nn + b];
        const size_t blk = (size_t)pat.nbr_ptr[r] + pm;
        const uint64_t bit = uint64_t(1) << (blk & 63);
        // elements of one group touch disjoint blocks, but their blocks may share a word of the bitmap
        const uint64_t old = __atomic_fetch_or(&seen[blk >> 6], bit, __ATOMIC_RELAXED);
        if (!(old & bit)) pm |= 0x80;
      }
    }
  };
  for (size_t g = 0; g + 1 < group_ptr.size(); ++g) {
    const int64_t g0 = group_ptr[g], g1 = group_ptr[g + 1];
    if (parallel_groups) {
      parallel_for(g1 - g0, [&](int64_t b0, int64_t e0, int) {
        for (int64_t k = b0; k < e0; ++k) flag(elems[g0 + k]);
      });
    } else {
      for (int64_t k = g0; k < g1; ++k) flag(elems[k]);
    }
  }
  // every block of an owned row is touched by some element (the pattern comes from the elements), so no value stays unset
  return true;
}

std::string build_ordered_scatter(const Pattern& pat, const double* xyz, const int32_t* conn, const uint8_t* diag_elem,
                                  int brick_elems, std::vector<int32_t>& order, std::vector<int32_t>& pred_ptr,
                                  std::vector<int32_t>& pred, std::vector<uint8_t>& posmap, bool& first_writer) {
  const int nn = pat.nn;
  const int64_t ne = pat.n_elem;
  // colours (greedy, as build_coloring) of the elements that touch an owned node
  std::vector<int32_t> cptr, celems;
  std::string err = build_coloring(pat, conn, cptr, celems);
  if (!err.empty()) return err;
  std::vector<int8_t> color(ne, -1);
  for (int c = 0; c + 1 < (int)cptr.size(); ++c)
    for (int32_t k = cptr[c]; k < cptr[c + 1]; ++k) color[celems[k]] = (int8_t)c;
  // bricks: cubes of side h in the bounding box of the element centroids, about brick_elems elements each
  double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
  std::vector<double> cen((size_t)ne * 3);
  for (int64_t e = 0; e < ne; ++e)
    for (int d = 0; d < 3; ++d) {
      double c = 0.;
      for (int a = 0; a < nn; ++a) c += xyz[(int64_t)conn[e * nn + a] * 3 + d];
      c /= nn;
      cen[e * 3 + d] = c; lo[d] = std::min(lo[d], c); hi[d] = std::max(hi[d], c);
    }
  const int64_t active = (int64_t)celems.size();
  double vol = 1.;
  int ext_dims = 0;
  for (int d = 0; d < 3; ++d)
    if (hi[d] > lo[d]) { vol *= hi[d] - lo[d]; ++ext_dims; }
  const double nbricks = std::max(1.0, double(active) / brick_elems);
  const double h = ext_dims ? std::pow(vol / nbricks, 1.0 / ext_dims) : 1.;
  int64_t nb[3];
  for (int d = 0; d < 3; ++d) nb[d] = std::max<int64_t>(1, (int64_t)std::ceil((hi[d] - lo[d]) / h + 1e-9));
  std::vector<std::pair<int64_t, int32_t>> keyed;  // (brick * 64 + colour, element)
  keyed.reserve(active);
  for (int64_t e = 0; e < ne; ++e) {
    if (color[e] < 0) continue;
    int64_t b[3];
    for (int d = 0; d < 3; ++d) b[d] = std::min<int64_t>(nb[d] - 1, (int64_t)((cen[e * 3 + d] - lo[d]) / h));
    const int64_t brick = (b[2] * nb[1] + b[1]) * nb[0] + b[0];
    keyed.emplace_back(brick * 64 + color[e], (int32_t)e);
  }
  std::sort(keyed.begin(), keyed.end());
  order.resize(keyed.size());
  std::vector<int32_t> posn(ne, -1);
  for (size_t k = 0; k < keyed.size(); ++k) { order[k] = keyed[k].second; posn[keyed[k].second] = (int32_t)k; }
  // predecessors: node-sharing elements earlier in the order (two passes: count, then fill)
  pred_ptr.assign(order.size() + 1, 0);
  auto collect = [&](size_t k, std::vector<int32_t>& scratch) {
    const int64_t e = order[k];
    scratch.clear();
    for (int a = 0; a < nn; ++a) {
      const int32_t n = conn[e * nn + a];
      for (int64_t q = pat.n2e_ptr[n]; q < pat.n2e_ptr[n + 1]; ++q) {
        const int32_t o = pat.n2e[q];
        if (o != e && posn[o] >= 0 && posn[o] < (int32_t)k) scratch.push_back(o);
      }
    }
    std::sort(scratch.begin(), scratch.end());
    scratch.erase(std::unique(scratch.begin(), scratch.end()), scratch.end());
  };
  std::atomic<int> too_many{0};
  parallel_for((int64_t)order.size(), [&](int64_t b0, int64_t e0, int) {
    std::vector<int32_t> scratch;
    for (int64_t k = b0; k < e0; ++k) {
      collect((size_t)k, scratch);
      if (scratch.size() > 64) too_many = 1;
      pred_ptr[k + 1] = (int32_t)scratch.size();
    }
  });
  if (too_many) return "an element has more than 64 node-sharing predecessors";
  for (size_t k = 0; k < order.size(); ++k) pred_ptr[k + 1] += pred_ptr[k];
  pred.resize((size_t)pred_ptr[order.size()]);
  parallel_for((int64_t)order.size(), [&](int64_t b0, int64_t e0, int) {
    std::vector<int32_t> scratch;
    for (int64_t k = b0; k < e0; ++k) {
      collect((size_t)k, scratch);
      std::copy(scratch.begin(), scratch.end(), pred.begin() + pred_ptr[k]);
    }
  });
  // position map, with the first-writer flag in bit 7 when every row has < 128 neighbours and there are no diagonal elements
  err = build_posmap(pat, conn, posmap);
  if (!err.empty()) return err;
  std::vector<int32_t> one_group = {0, (int32_t)order.size()};
  first_writer = flag_first_writers(pat, conn, diag_elem, one_group, order, /*parallel_groups=*/false, posmap);
  return "";
}

}  // namespace mastb200
