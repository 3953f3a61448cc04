// Tile planner (host): see eqd_tiles.h.  No reference code path corresponds to
// this file: the reference sweeps elements in storage order and scatter-adds
// (assembleGlobalKU.f90:11-66); the tiles only re-group that same sweep so that
// one CTA can assemble a brick's nodal forces on chip.
#include "eqd_tiles.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <stdexcept>
#include <thread>

#include "eqd_box.h"
#include "eqd_dev.cuh"
#include "eqd_par.h"

namespace eqd {

namespace {

template <class F>
void parallel_for(int n, F&& fn) {
  const int nt = host_threads();
  if (n < 64 || nt == 1) { fn(0, n); return; }
  std::vector<std::thread> th;
  const int chunk = (n + nt - 1) / nt;
  for (int t = 0; t < nt; ++t) {
    const int b = t * chunk, e = std::min(n, b + chunk);
    if (b >= e) break;
    th.emplace_back([=, &fn] { fn(b, e); });
  }
  for (auto& x : th) x.join();
}

inline int balanced(int extent, int target, int& nblocks) {
  nblocks = std::max(1, (extent + target - 1) / target);
  return std::max(1, (extent + nblocks - 1) / nblocks);
}

// EQD_VERBOSE=1: wall-clock laps of the planner's phases on stderr
struct PlanLap {
  bool on = std::getenv("EQD_VERBOSE") != nullptr;
  std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
  void lap(const char* label, int n) {
    if (!on || n < 100000) return;
    auto now = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[eqd]   plan_tiles(%d): %s %.3f s\n", n, label, std::chrono::duration<double>(now - t).count());
    t = now;
  }
};

struct Tile {
  int b, e;                 // range in the sorted element order
  std::vector<int> nodes;   // ascending unique node ids
};

// Small open-addressing table node id -> value, reused tile after tile by one thread: the
// 8 x elements node ids of a tile are de-duplicated in O(ids) before the (much shorter) list of
// unique nodes is sorted, and the tile-local index of a node is then one probe instead of a
// binary search.
class NodeTable {
 public:
  NodeTable() : key_(kSize, -1), val_(kSize, 0) {}
  void clear() {
    for (int s : used_) key_[s] = -1;
    used_.clear();
  }
  // returns true when id was not in the table yet
  bool insert(int id, int value) {
    unsigned s = hash(id);
    while (key_[s] >= 0) {
      if (key_[s] == id) return false;
      s = (s + 1) & (kSize - 1);
    }
    key_[s] = id; val_[s] = value; used_.push_back((int)s);
    return true;
  }
  int find(int id) const {
    unsigned s = hash(id);
    while (key_[s] != id) {
      if (key_[s] < 0) return -1;
      s = (s + 1) & (kSize - 1);
    }
    return val_[s];
  }
  size_t size() const { return used_.size(); }

 private:
  static constexpr unsigned kSize = 1u << 14;   // > 4 x the ids of the largest tile (8 x capE = 5120 at most)
  static unsigned hash(int id) { return ((unsigned)id * 2654435761u) >> 18; }
  std::vector<int> key_, val_, used_;
};

void unique_nodes(const int* conn, const std::vector<int>& elems, const raw_vector<int>& order, int b, int e,
                  NodeTable& tab, std::vector<int>& out) {
  out.clear();
  tab.clear();
  for (int k = b; k < e; ++k) {
    const int* c = conn + 8 * (size_t)elems[order[k]];
    for (int j = 0; j < 8; ++j)
      if (tab.insert(c[j], 0)) out.push_back(c[j]);
  }
  std::sort(out.begin(), out.end());
}

void split_group(const int* conn, const std::vector<int>& elems, const raw_vector<int>& order, int b, int e,
                 const TileShape& sh, NodeTable& tab, std::vector<Tile>& out) {
  if (b >= e) return;
  Tile t;
  t.b = b; t.e = e;
  if (e - b <= sh.capE && e - b <= 640) {   // 640: the table's capacity contract (8 x 640 ids)
    unique_nodes(conn, elems, order, b, e, tab, t.nodes);
    if ((int)t.nodes.size() <= sh.capN || e - b == 1) { out.push_back(std::move(t)); return; }
  }
  const int m = b + (e - b) / 2;
  split_group(conn, elems, order, b, m, sh, tab, out);
  split_group(conn, elems, order, m, e, sh, tab, out);
}

}  // namespace

bool infer_grid(const int* conn, const int* etype, int Ne, int Nn, int& ny, int& nz) {
  for (int e = 0; e < Ne; ++e) {
    if (etype[e] != 1 && etype[e] != 2) continue;
    int a[8];
    for (int k = 0; k < 8; ++k) a[k] = conn[8 * (size_t)e + k];
    std::sort(a, a + 8);
    const int p = a[2] - a[0], q = a[4] - a[0];
    if (a[1] - a[0] != 1 || p <= 1 || q <= p + 1 || q % p != 0) continue;
    if (a[3] != a[0] + p + 1 || a[5] != a[0] + q + 1 || a[6] != a[0] + q + p || a[7] != a[0] + q + p + 1) continue;
    ny = p; nz = q / p;
    (void)Nn;
    return true;
  }
  return false;
}

void plan_tiles(const int* conn, const std::vector<int>& elems, int Nn, int ny, int nz, bool gridOk,
                const TileShape& sh, int NT, TilePlan& P) {
  const int n = (int)elems.size();
  P = TilePlan();
  P.n = n;
  if (n == 0) { P.S = 32; P.PFS = 4; P.LS = 1; P.refId.assign(P.S, -1); P.tileNode.assign(1, 0); P.tnode.assign(P.PFS, -1); P.lconn.assign(8 * (size_t)P.S, 0); return; }
  if (sh.capN > (int)EQD_LN_MASK) throw std::runtime_error("tile node cap exceeds the 12-bit local index");
  PlanLap lap;
  // ---- brick key per element
  raw_vector<int> key(n);
  raw_vector<int> cx, cy, cz;   // grid cell of every element (structured part of the mesh)
  int nKeys = 1;
  if (gridOk) {
    const long nynz = (long)ny * nz;
    cx.resize(n); cy.resize(n); cz.resize(n);
    parallel_for(n, [&](int b, int e) {
      for (int j = b; j < e; ++j) {
        const int* c = conn + 8 * (size_t)elems[j];
        int n0 = c[0];
        for (int k = 1; k < 8; ++k) n0 = std::min(n0, c[k]);
        cx[j] = (int)(n0 / nynz); cz[j] = (int)((n0 % nynz) / ny); cy[j] = (int)(n0 % ny);
      }
    });
    int mn[3] = {cx[0], cz[0], cy[0]}, mx[3] = {cx[0], cz[0], cy[0]};
    {
      std::mutex mu;
      parallel_for(n, [&](int b, int e) {
        int lmn[3] = {cx[b], cz[b], cy[b]}, lmx[3] = {cx[b], cz[b], cy[b]};
        for (int j = b; j < e; ++j) {
          lmn[0] = std::min(lmn[0], cx[j]); lmx[0] = std::max(lmx[0], cx[j]);
          lmn[1] = std::min(lmn[1], cz[j]); lmx[1] = std::max(lmx[1], cz[j]);
          lmn[2] = std::min(lmn[2], cy[j]); lmx[2] = std::max(lmx[2], cy[j]);
        }
        std::lock_guard<std::mutex> g(mu);
        for (int k = 0; k < 3; ++k) { mn[k] = std::min(mn[k], lmn[k]); mx[k] = std::max(mx[k], lmx[k]); }
      });
    }
    int nb[3];
    const int bxe = balanced(mx[0] - mn[0] + 1, sh.bx, nb[0]);
    const int bze = balanced(mx[1] - mn[1] + 1, sh.bz, nb[1]);
    const int bye = balanced(mx[2] - mn[2] + 1, sh.by, nb[2]);
    const double tot = (double)nb[0] * nb[1] * nb[2];
    if (tot > 1.0e9) throw std::runtime_error("tile planner: brick grid too large");
    nKeys = nb[0] * nb[1] * nb[2];
    parallel_for(n, [&](int b, int e) {
      for (int j = b; j < e; ++j)
        key[j] = (((cx[j] - mn[0]) / bxe) * nb[1] + (cz[j] - mn[1]) / bze) * nb[2] + (cy[j] - mn[2]) / bye;
    });
  } else {
    const int per = std::max(32, std::min(sh.capE, sh.bx * sh.bz * sh.by));
    nKeys = (n + per - 1) / per;
    for (int j = 0; j < n; ++j) key[j] = j / per;
  }
  lap.lap("brick keys", n);
  // ---- stable counting sort by key (elements keep ascending reference order inside a brick)
  std::vector<int> start(nKeys + 1, 0);
  for (int j = 0; j < n; ++j) start[key[j] + 1]++;
  for (int k = 0; k < nKeys; ++k) start[k + 1] += start[k];
  raw_vector<int> order(n);
  {
    std::vector<int> fill(start.begin(), start.end() - 1);
    for (int j = 0; j < n; ++j) order[fill[key[j]]++] = j;
  }
  lap.lap("counting sort", n);
  // ---- non-empty groups -> tiles
  std::vector<int> groups;
  for (int k = 0; k < nKeys; ++k) if (start[k + 1] > start[k]) groups.push_back(k);
  const int nG = (int)groups.size();
  std::vector<std::vector<Tile>> gt(nG);
  parallel_for(nG, [&](int b, int e) {
    NodeTable tab;
    for (int g = b; g < e; ++g) split_group(conn, elems, order, start[groups[g]], start[groups[g] + 1], sh, tab, gt[g]);
  });
  lap.lap("split into tiles (node lists)", n);
  std::vector<Tile*> tiles;
  for (auto& v : gt) for (auto& t : v) tiles.push_back(&t);
  // ---- option bankOrder == 2: residue numbering of the tile-local nodes.  For a tile that is a complete brick of
  // ex x ez x ey grid cells (elements in ascending id = x, z, y order; every node a regular grid node) the local
  // index of node (ix, iz, iy) is chosen with  index mod 16 == (iy + ey*iz + ey*ez*ix) mod 16 : the corner-c nodes
  // of 16 consecutive elements then have 16 consecutive residues, i.e. every half-warp access is conflict free.
  // Nodes of one residue class are numbered 16 apart; classes are not equally large, the gaps stay empty (-1).
  if (sh.bankOrder == 2 && gridOk) {
    const long nynz = (long)ny * nz;
    parallel_for((int)tiles.size(), [&](int tb, int te) {
      std::vector<int> layout;
      for (int t = tb; t < te; ++t) {
        Tile& T = *tiles[t];
        const int ne = T.e - T.b;
        int lo[3] = {1 << 30, 1 << 30, 1 << 30}, hi[3] = {-1, -1, -1};
        for (int k = T.b; k < T.e; ++k) {
          const int j = order[k];
          const int c3[3] = {cx[j], cz[j], cy[j]};
          for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], c3[a]); hi[a] = std::max(hi[a], c3[a]); }
        }
        const int ex = hi[0] - lo[0] + 1, ez = hi[1] - lo[1] + 1, ey = hi[2] - lo[2] + 1;
        if ((long)ex * ez * ey != ne || (long)(ex + 1) * (ez + 1) * (ey + 1) != (long)T.nodes.size()) continue;
        // every corner must be the regular grid node of its cell corner (no split-node masters, no wedges),
        // and the elements must come in x, z, y order
        bool ok = true;
        for (int k = 0; k < ne && ok; ++k) {
          const int j = order[T.b + k];
          const int px = cx[j] - lo[0], pz = cz[j] - lo[1], py = cy[j] - lo[2];
          if (k != (px * ez + pz) * ey + py) { ok = false; break; }
          const int* c = conn + 8 * (size_t)elems[j];
          for (int i = 0; i < 8; ++i) {
            const long want = (long)(cx[j] + (box_px(i) ? 1 : 0)) * nynz + (long)(cz[j] + (box_pz(i) ? 1 : 0)) * ny + cy[j] + (box_py(i) ? 1 : 0);
            if (c[i] != want) { ok = false; break; }
          }
        }
        if (!ok) continue;
        int cnt[16] = {0};
        auto cls = [&](int ix, int iz, int iy) { return (iy + ey * iz + ey * ez * ix) & 15; };
        for (int ix = 0; ix <= ex; ++ix) for (int iz = 0; iz <= ez; ++iz) for (int iy = 0; iy <= ey; ++iy) cnt[cls(ix, iz, iy)]++;
        int mx = 0;
        for (int q = 0; q < 16; ++q) mx = std::max(mx, cnt[q]);
        if (16 * mx > sh.capN) continue;                       // does not fit the shared-memory rows: keep ascending ids
        layout.assign(16 * (size_t)mx, -1);
        int fill[16] = {0};
        for (int ix = 0; ix <= ex; ++ix) for (int iz = 0; iz <= ez; ++iz) for (int iy = 0; iy <= ey; ++iy) {
          const int q = cls(ix, iz, iy);
          layout[16 * (size_t)fill[q]++ + q] = (int)((long)(lo[0] + ix) * nynz + (long)(lo[1] + iz) * ny + lo[2] + iy);
        }
        while (!layout.empty() && layout.back() < 0) layout.pop_back();
        T.nodes = layout;                                      // position = tile-local index; -1 = unused slot
      }
    });
    lap.lap("residue numbering", n);
  }
  // ---- shared-memory bank model of a tile's element order, and (option) the order that minimises it.
  // A warp's 8-byte shared-memory access is served half-warp by half-warp; lanes whose words lie in the same
  // bank pair (local node index mod 16) but at different addresses need one wavefront each.  The kernels read
  // and update the tile's node rows at index lconn(corner, element), lanes = consecutive elements of a stage.
  if (sh.bankOrder || std::getenv("EQD_VERBOSE")) {   // the model costs a pass over the elements: not on the default path
    static_assert(EQD_STAGE % 16 == 0 && EQD_STAGE_PML % 16 == 0, "half-warps must not straddle stages");
    std::vector<long> ideal(tiles.size(), 0), asc(tiles.size(), 0), chosen(tiles.size(), 0);
    parallel_for((int)tiles.size(), [&](int tb, int te) {
      NodeTable tab;
      std::vector<int> li, perm, best, idx;
      auto cost = [&](const std::vector<int>& ord, const Tile& T) {
        long w = 0;
        const int ne = T.e - T.b;
        for (int base = 0; base < ne; base += 16) {
          const int m = std::min(16, ne - base);
          for (int i = 0; i < 8; ++i) {
            int mult[16] = {0}, seen[16][16];
            int wf = 1;
            for (int l = 0; l < m; ++l) {
              const int v = li[8 * (size_t)ord[base + l] + i], bk = v & 15;
              bool dup = false;
              for (int q = 0; q < mult[bk]; ++q) dup = dup || seen[bk][q] == v;
              if (!dup) { seen[bk][mult[bk]] = v; wf = std::max(wf, ++mult[bk]); }
            }
            w += wf;
          }
        }
        return w;
      };
      for (int t = tb; t < te; ++t) {
        Tile& T = *tiles[t];
        const int ne = T.e - T.b;
        tab.clear();
        for (int i = 0; i < (int)T.nodes.size(); ++i) if (T.nodes[i] >= 0) tab.insert(T.nodes[i], i);
        li.resize(8 * (size_t)ne);
        for (int k = 0; k < ne; ++k)
          for (int i = 0; i < 8; ++i) li[8 * (size_t)k + i] = tab.find(conn[8 * (size_t)elems[order[T.b + k]] + i]);
        idx.resize(ne);
        for (int k = 0; k < ne; ++k) idx[k] = k;
        ideal[t] = 8L * ((ne + 15) / 16);
        asc[t] = cost(idx, T);
        chosen[t] = asc[t];
        if (sh.bankOrder != 1 || !gridOk || asc[t] == ideal[t]) continue;
        best = idx;
        // the six orders of the grid axes (slowest .. fastest); ties keep the earlier candidate, ascending id first
        static const int axes[6][3] = {{0, 1, 2}, {0, 2, 1}, {1, 0, 2}, {1, 2, 0}, {2, 0, 1}, {2, 1, 0}};   // 0 = x, 1 = z, 2 = y
        for (int c = 1; c < 6; ++c) {
          perm = idx;
          auto coord = [&](int k, int a) { const int j = order[T.b + k]; return a == 0 ? cx[j] : a == 1 ? cz[j] : cy[j]; };
          std::stable_sort(perm.begin(), perm.end(), [&](int p, int q) {
            for (int a = 0; a < 3; ++a) { const int u = coord(p, axes[c][a]), v = coord(q, axes[c][a]); if (u != v) return u < v; }
            return false;
          });
          const long w = cost(perm, T);
          if (w < chosen[t]) { chosen[t] = w; best = perm; }
        }
        if (chosen[t] < asc[t]) {
          std::vector<int> moved(ne);
          for (int k = 0; k < ne; ++k) moved[k] = order[T.b + best[k]];
          for (int k = 0; k < ne; ++k) order[T.b + k] = moved[k];
        }
      }
    });
    for (size_t t = 0; t < tiles.size(); ++t) { P.bankIdeal += ideal[t]; P.bankAscending += asc[t]; P.bankChosen += chosen[t]; }
    if (lap.on && n >= 100000)
      std::fprintf(stderr, "[eqd]   plan_tiles(%d): modelled corner wavefronts ideal %ld, ascending order %ld, chosen order %ld\n", n,
                   P.bankIdeal, P.bankAscending, P.bankChosen);
    lap.lap("bank model / element order", n);
  }
  const int nT = (int)tiles.size();
  P.nTiles = nT;
  P.tileElem.resize(nT); P.tileCnt.resize(nT); P.tileNode.resize(nT + 1); P.tileColours.assign(nT, 1);
  long slot = 0, nslot = 0;
  int LS = 1;
  for (int t = 0; t < nT; ++t) {
    P.tileElem[t] = (int)slot; P.tileCnt[t] = tiles[t]->e - tiles[t]->b;
    slot += (P.tileCnt[t] + 31) / 32 * 32;
    P.tileNode[t] = (int)nslot;
    const int ln = (int)tiles[t]->nodes.size();
    if (ln > (int)EQD_LN_MASK) throw std::runtime_error("tile planner: one element brick touches more than 4095 nodes");
    LS = std::max(LS, (ln + 3) / 4 * 4);
    nslot += (ln + 3) / 4 * 4;
    if (slot > (1L << 29) || nslot > (1L << 30)) throw std::runtime_error("tile planner: class too large for 32-bit slots");
  }
  P.tileNode[nT] = (int)nslot;
  P.S = (int)std::max(slot, 32L);
  P.PFS = (int)std::max(nslot, 4L);
  P.LS = LS;
  P.refId.resize(P.S); P.tnode.resize(P.PFS); P.lconn.resize(8 * (size_t)P.S);
  parallel_range((size_t)P.S, [&](size_t b, size_t e) { std::fill(P.refId.begin() + b, P.refId.begin() + e, -1); });
  parallel_range((size_t)P.PFS, [&](size_t b, size_t e) { std::fill(P.tnode.begin() + b, P.tnode.begin() + e, -1); });
  parallel_range(8 * (size_t)P.S, [&](size_t b, size_t e) { std::fill(P.lconn.begin() + b, P.lconn.begin() + e, (uint16_t)0); });
  lap.lap("slots + fills", n);
  const size_t S = P.S;
  std::vector<int> bad(1, 0);
  parallel_for(nT, [&](int tb, int te) {
    std::vector<int> stamp, cnt;
    NodeTable tab;
    for (int t = tb; t < te; ++t) {
      const Tile& T = *tiles[t];
      const int ln = (int)T.nodes.size();
      std::copy(T.nodes.begin(), T.nodes.end(), P.tnode.begin() + P.tileNode[t]);
      tab.clear();
      for (int i = 0; i < ln; ++i) if (T.nodes[i] >= 0) tab.insert(T.nodes[i], i);   // node id -> tile-local index
      stamp.assign(8 * (size_t)ln, -1); cnt.assign(8 * (size_t)ln, 0);
      int maxc = 1;
      for (int k = T.b; k < T.e; ++k) {
        const int le = k - T.b;
        const size_t s = (size_t)P.tileElem[t] + le;
        const int e = elems[order[k]];
        P.refId[s] = e;
        const int pass = le / NT;
        for (int i = 0; i < 8; ++i) {
          const int nd = conn[8 * (size_t)e + i];
          const int li = tab.find(nd);
          // colour = how many earlier elements of this (pass, phase) hit the same node
          const size_t q = 8 * (size_t)li + i;
          if (stamp[q] != pass) { stamp[q] = pass; cnt[q] = 0; }
          const int col = cnt[q]++;
          if (col > 15) { bad[0] = 1; continue; }
          maxc = std::max(maxc, col + 1);
          P.lconn[(size_t)i * S + s] = (uint16_t)(li | (col << EQD_LN_BITS));
        }
      }
      P.tileColours[t] = (uint8_t)maxc;
    }
  });
  lap.lap("local connectivity + colours", n);
  if (bad[0]) throw std::runtime_error("tile planner: more than 16 elements of one phase share a node");
  (void)Nn;
}

}  // namespace eqd

// ----------------------------------------------------------------------------
// Host-only self-check of the planner on a sub-domain's connectivity (needs no
// GPU; used by the CPU test-suite and by `eqd_set_option(h, "check_tiles", 1)`).
// stats[0..7] = tiles, elements, padded slots, tile-node slots, max tile nodes,
// max colours, tiles with more than one colour, grid inferred (0/1), per class
// c at stats[8*c ..].  Returns 0 when every invariant holds, else a line number.
extern "C" int eqd_plan_check(int32_t Nn, int32_t Ne, const int32_t* nodeElemIdRelation, const int32_t* elemTypeArr,
                              const int32_t* numOfDofPerNodeArr, int64_t* stats) {
  using namespace eqd;
  if (Nn <= 0 || Ne <= 0 || !nodeElemIdRelation || !elemTypeArr || !numOfDofPerNodeArr || !stats) return __LINE__;
  try {
    std::vector<int> conn(8 * (size_t)Ne);
    for (size_t k = 0; k < conn.size(); ++k) {
      conn[k] = nodeElemIdRelation[k] - 1;
      if (conn[k] < 0 || conn[k] >= Nn) return __LINE__;
    }
    std::vector<int> members[3];
    for (int e = 0; e < Ne; ++e) {
      int c = CLS_PML;
      if (elemTypeArr[e] != 2) {
        c = CLS_REG;
        for (int k = 0; k < 8; ++k) if (numOfDofPerNodeArr[conn[8 * (size_t)e + k]] == 12) c = CLS_REGX;
      }
      members[c].push_back(e);
    }
    int ny = 0, nz = 0;
    const bool ok = infer_grid(conn.data(), elemTypeArr, Ne, Nn, ny, nz);
    std::vector<char> seen(Ne, 0);
    for (int c = 0; c < 3; ++c) {
      TileShape sh;
      if (c == CLS_PML) { sh.bx = kPmlBrick[0]; sh.bz = kPmlBrick[1]; sh.by = kPmlBrick[2]; sh.capE = 320; sh.capN = EQD_PML_LS; }
      else { sh.bx = kRegBrick[0]; sh.bz = kRegBrick[1]; sh.by = kRegBrick[2]; sh.capE = 384; sh.capN = EQD_REG_LS; }
      const int NTP = c == CLS_PML ? EQD_STAGE_PML : EQD_STAGE;
      TilePlan P;
      plan_tiles(conn.data(), members[c], Nn, ny, nz, ok, sh, NTP, P);
      int64_t* st = stats + 8 * c;
      st[0] = P.nTiles; st[1] = P.n; st[2] = P.S; st[3] = P.PFS; st[4] = P.LS; st[5] = 0; st[6] = 0; st[7] = ok;
      if (P.n == 0) continue;
      for (int t = 0; t < P.nTiles; ++t) {
        const int nb = P.tileNode[t], ln = P.tileNode[t + 1] - nb;
        if (P.tileElem[t] % 32 || nb % 4 || ln > P.LS || P.tileCnt[t] > sh.capE) return __LINE__;
        for (int i = 1; i < ln; ++i)
          if (P.tnode[nb + i] >= 0 && P.tnode[nb + i] <= P.tnode[nb + i - 1]) return __LINE__;
        st[5] = std::max<int64_t>(st[5], P.tileColours[t]);
        st[6] += P.tileColours[t] > 1;
        // replay the kernel's assembly schedule and look for write conflicts
        std::vector<int> owner(ln);
        for (int base = 0; base < P.tileCnt[t]; base += NTP)
          for (int i = 0; i < 8; ++i)
            for (int col = 0; col < P.tileColours[t]; ++col) {
              std::fill(owner.begin(), owner.end(), -1);
              for (int le = base; le < std::min(P.tileCnt[t], base + NTP); ++le) {
                const size_t s = (size_t)P.tileElem[t] + le;
                const unsigned u = P.lconn[(size_t)i * P.S + s];
                if ((int)(u >> EQD_LN_BITS) != col) continue;
                const int li = u & EQD_LN_MASK;
                if (li >= ln || owner[li] >= 0) return __LINE__;
                owner[li] = le;
              }
            }
        for (int le = 0; le < P.tileCnt[t]; ++le) {
          const size_t s = (size_t)P.tileElem[t] + le;
          const int e = P.refId[s];
          if (e < 0 || e >= Ne || seen[e]) return __LINE__;
          seen[e] = 1;
          for (int i = 0; i < 8; ++i) {
            const unsigned u = P.lconn[(size_t)i * P.S + s];
            if (P.tnode[nb + (u & EQD_LN_MASK)] != conn[8 * (size_t)e + i]) return __LINE__;
            if ((int)(u >> EQD_LN_BITS) >= P.tileColours[t]) return __LINE__;
          }
        }
      }
    }
    for (int e = 0; e < Ne; ++e) if (!seen[e]) return __LINE__;
  } catch (const std::exception&) {
    return __LINE__;
  }
  return 0;
}

// ----------------------------------------------------------------------------
// Host-only model of the tile kernels' shared-memory bank behaviour (no GPU): plans the tiles of the three
// classes as eqd_set_mesh would, with the element order inside a tile ascending (bank_order = 0) or chosen to
// minimise conflicts (bank_order = 1, eqd_set_option "bank_order"), and returns per class the modelled
// wavefronts of the corner accesses: out[3*c + 0..2] = conflict-free, ascending order, chosen order.
// Also re-checks the planner's invariants for the chosen order.  Returns 0 or a line number.
extern "C" int eqd_plan_bank_model(int32_t Nn, int32_t Ne, const int32_t* nodeElemIdRelation, const int32_t* elemTypeArr,
                                   const int32_t* numOfDofPerNodeArr, int32_t bank_order, int64_t* out) {
  using namespace eqd;
  if (Nn <= 0 || Ne <= 0 || !nodeElemIdRelation || !elemTypeArr || !numOfDofPerNodeArr || !out) return __LINE__;
  try {
    std::vector<int> conn(8 * (size_t)Ne);
    for (size_t k = 0; k < conn.size(); ++k) {
      conn[k] = nodeElemIdRelation[k] - 1;
      if (conn[k] < 0 || conn[k] >= Nn) return __LINE__;
    }
    std::vector<int> members[3];
    for (int e = 0; e < Ne; ++e) {
      int c = CLS_PML;
      if (elemTypeArr[e] != 2) {
        c = CLS_REG;
        for (int k = 0; k < 8; ++k) if (numOfDofPerNodeArr[conn[8 * (size_t)e + k]] == 12) c = CLS_REGX;
      }
      members[c].push_back(e);
    }
    int ny = 0, nz = 0;
    const bool ok = infer_grid(conn.data(), elemTypeArr, Ne, Nn, ny, nz);
    for (int c = 0; c < 3; ++c) {
      TileShape sh;
      if (c == CLS_PML) { sh.bx = kPmlBrick[0]; sh.bz = kPmlBrick[1]; sh.by = kPmlBrick[2]; sh.capE = 320; sh.capN = EQD_PML_LS; }
      else { sh.bx = kRegBrick[0]; sh.bz = kRegBrick[1]; sh.by = kRegBrick[2]; sh.capE = 384; sh.capN = EQD_REG_LS; }
      sh.bankOrder = bank_order == 2 ? 2 : 1;   // always run the model ...
      TilePlan P;
      plan_tiles(conn.data(), members[c], Nn, ny, nz, ok, sh, c == CLS_PML ? EQD_STAGE_PML : EQD_STAGE, P);
      // bank_order 2: "ascending" already is the count under the residue numbering (the element order is kept)
      out[3 * c] = P.bankIdeal; out[3 * c + 1] = P.bankAscending; out[3 * c + 2] = bank_order == 1 ? P.bankChosen : P.bankAscending;
      if (!bank_order || P.n == 0) continue; // ... the chosen order is only checked when asked for
      // every element exactly once, local connectivity consistent, assembly schedule conflict free
      std::vector<char> seen(Ne, 0);
      const int NTP = c == CLS_PML ? EQD_STAGE_PML : EQD_STAGE;
      for (int t = 0; t < P.nTiles; ++t) {
        const int nb = P.tileNode[t], ln = P.tileNode[t + 1] - nb;
        std::vector<int> owner(ln);
        for (int base = 0; base < P.tileCnt[t]; base += NTP)
          for (int i = 0; i < 8; ++i)
            for (int col = 0; col < P.tileColours[t]; ++col) {
              std::fill(owner.begin(), owner.end(), -1);
              for (int le = base; le < std::min(P.tileCnt[t], base + NTP); ++le) {
                const unsigned u = P.lconn[(size_t)i * P.S + (size_t)P.tileElem[t] + le];
                if ((int)(u >> EQD_LN_BITS) != col) continue;
                const int li = u & EQD_LN_MASK;
                if (li >= ln || owner[li] >= 0) return __LINE__;
                owner[li] = le;
              }
            }
        for (int le = 0; le < P.tileCnt[t]; ++le) {
          const size_t s = (size_t)P.tileElem[t] + le;
          const int e = P.refId[s];
          if (e < 0 || e >= Ne || seen[e]) return __LINE__;
          seen[e] = 1;
          for (int i = 0; i < 8; ++i)
            if (P.tnode[nb + (P.lconn[(size_t)i * P.S + s] & EQD_LN_MASK)] != conn[8 * (size_t)e + i]) return __LINE__;
        }
      }
      for (int e : members[c]) if (!seen[e]) return __LINE__;
    }
  } catch (const std::exception&) {
    return __LINE__;
  }
  return 0;
}

// ----------------------------------------------------------------------------
// Host-only check of the closed-form box operators (eqd_box.h) against the
// reference's precomputed ones (needs no GPU; CPU test-suite).  For every element
// that passes the geometric box test, the strain, the B^T t nodal forces
// (calcElemKU.f90:44-60,175-189) and the hourglass forces (hrglss.f90:20-54) of a
// pseudo-random nodal field are evaluated both ways.  dev[0..2] = the largest
// deviation of the three, relative to the element's largest component;
// *nBox = elements that passed the test.  Returns 0, or a line number on bad input.
extern "C" int eqd_box_check(int32_t Nn, int32_t Ne, const double* meshCoor, const int32_t* nodeElemIdRelation,
                             const int32_t* elemTypeArr, const double* eleshp, const double* phi, const double* ss,
                             int64_t* nBox, double* dev) {
  using namespace eqd;
  if (Nn <= 0 || Ne <= 0 || !meshCoor || !nodeElemIdRelation || !elemTypeArr || !eleshp || !phi || !ss || !nBox || !dev) return __LINE__;
  auto field = [](int node, int c) {   // deterministic values in (-1, 1), no pattern along the grid
    uint32_t x = (uint32_t)node * 2654435761u + (uint32_t)c * 40503u + 12345u;
    x ^= x >> 15; x *= 2246822519u; x ^= x >> 13; x *= 3266489917u; x ^= x >> 16;
    return (double)x / 2147483648.0 - 1.0;
  };
  int64_t count = 0;
  double d[3] = {0, 0, 0};
  for (int e = 0; e < Ne; ++e) {
    int c[8];
    for (int k = 0; k < 8; ++k) { c[k] = nodeElemIdRelation[8 * (size_t)e + k] - 1; if (c[k] < 0 || c[k] >= Nn) return __LINE__; }
    if (elemTypeArr[e] == 11 || elemTypeArr[e] == 12 || !box_element(c, meshCoor)) continue;
    ++count;
    const double* shp = eleshp + 24 * (size_t)e;   // eleshp(3,8,e)
    const double* ph = phi + 32 * (size_t)e;       // phi(8,4,e)
    const double* s6 = ss + 6 * (size_t)e;
    double u[8][3];
    for (int i = 0; i < 8; ++i) for (int k = 0; k < 3; ++k) u[i][k] = field(c[i], k);
    // general forms
    double sr[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 8; ++i) {
      const double s1 = shp[3 * i], s2 = shp[3 * i + 1], s3 = shp[3 * i + 2];
      sr[0] += s1 * u[i][0]; sr[1] += s2 * u[i][1]; sr[2] += s3 * u[i][2];
      sr[3] += s3 * u[i][1] + s2 * u[i][2]; sr[4] += s3 * u[i][0] + s1 * u[i][2]; sr[5] += s2 * u[i][0] + s1 * u[i][1];
    }
    double t[6];
    for (int k = 0; k < 6; ++k) t[k] = field(c[0] + 7 * k, 3 + k);
    double fg[8][3], hg[8][3];
    for (int i = 0; i < 8; ++i) {
      const double s1 = shp[3 * i], s2 = shp[3 * i + 1], s3 = shp[3 * i + 2];
      fg[i][0] = s1 * t[0] + s3 * t[4] + s2 * t[5];
      fg[i][1] = s2 * t[1] + s3 * t[3] + s1 * t[5];
      fg[i][2] = s3 * t[2] + s2 * t[3] + s1 * t[4];
    }
    double phid[4][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, hv[4][3];
    for (int i = 0; i < 8; ++i) for (int m = 0; m < 4; ++m) for (int k = 0; k < 3; ++k) phid[m][k] += ph[8 * m + i] * u[i][k];
    for (int m = 0; m < 4; ++m) {
      hv[m][0] = s6[0] * phid[m][0] + s6[1] * phid[m][1] + s6[2] * phid[m][2];
      hv[m][1] = s6[1] * phid[m][0] + s6[3] * phid[m][1] + s6[4] * phid[m][2];
      hv[m][2] = s6[2] * phid[m][0] + s6[4] * phid[m][1] + s6[5] * phid[m][2];
    }
    for (int i = 0; i < 8; ++i) for (int k = 0; k < 3; ++k) {
      double h = 0.0;
      for (int m = 0; m < 4; ++m) h -= ph[8 * m + i] * hv[m][k];
      hg[i][k] = h;
    }
    // closed forms
    const double ax = shp[BOX_AX], ay = shp[BOX_AY], az = shp[BOX_AZ];
    double g[3][3], sb[6], fb[8][3], hb[8][3];
    box_grad(u, g);
    box_strain(g, ax, ay, az, sb);
    box_force(t, ax, ay, az, fb);
    box_hourglass(u, s6[0], s6[3], s6[5], hb);
    double sc = 0, df = 0;
    for (int k = 0; k < 6; ++k) { sc = std::max(sc, std::fabs(sr[k])); df = std::max(df, std::fabs(sr[k] - sb[k])); }
    if (sc > 0) d[0] = std::max(d[0], df / sc);
    sc = 0; df = 0;
    for (int i = 0; i < 8; ++i) for (int k = 0; k < 3; ++k) { sc = std::max(sc, std::fabs(fg[i][k])); df = std::max(df, std::fabs(fg[i][k] - fb[i][k])); }
    if (sc > 0) d[1] = std::max(d[1], df / sc);
    sc = 0; df = 0;
    for (int i = 0; i < 8; ++i) for (int k = 0; k < 3; ++k) { sc = std::max(sc, std::fabs(hg[i][k])); df = std::max(df, std::fabs(hg[i][k] - hb[i][k])); }
    if (sc > 0) d[2] = std::max(d[2], df / sc);
  }
  *nBox = count;
  for (int k = 0; k < 3; ++k) dev[k] = d[k];
  return 0;
}
