// Marching kernel (eqd_march.h) and its host-side planner.
//
// The planner is pure index work on the reference's connectivity (nodeElemIdRelation,
// meshgen.f90:702-741) and coordinates (meshCoor): it finds the axis-aligned hexahedra of the
// regular class that sit on the structured node lattice (node id = ix*ny*nz + iz*ny + iy,
// meshgen.f90:904-919), cuts the lattice into column tiles that never straddle a plane where
// neighbouring elements do not share their nodes (the fault: elements on its + side reference the
// appended split-node masters, replaceSlaveWithMasterNode :743-766), and splits the tiles' x extents
// into bundles so that every CTA of the persistent grid marches the same number of element planes.
// The node lattice of every bundle is read off the connectivity and checked against every element
// that touches a lattice position; nothing is assumed about the ids (masters are welcome).
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <mutex>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "eqd_march.h"
#include "eqd_march_pml.h"
#include "eqd_march_plan.h"
#include "eqd_par.h"
#include "eqd_tiles.h"

namespace eqd {

__global__ void __launch_bounds__(MK_NT, 3) k_march(MarchArgs A) {
  extern __shared__ __align__(128) unsigned char mk_smraw[];
  MarchShared& sm = *reinterpret_cast<MarchShared*>(mk_smraw);
  const int tid = threadIdx.x;
  MarchRegs R;
  unsigned par = 0;   // bit s = phase parity of operator stage s
  if (tid == 0) { mk_bar_init(&sm.bar[0]); mk_bar_init(&sm.bar[1]); }
  __syncthreads();
  // this CTA's share of the boundary list, then of the interior list (either may be switched off by the launch)
  for (int part = 0; part < 2; ++part) {
  const int* first = part == 0 ? A.ctaFirstA : A.ctaFirstB;
  if (!first) continue;
  const int b0 = first[blockIdx.x], b1 = first[blockIdx.x + 1];
  for (int b = b0; b < b1; ++b) {
    const MarchBundle B = A.rec[b];
#define MK_RUN(body) do { body; __syncthreads(); } while (0)
#define MK_RUNNS(body) do { body; } while (0)
#define MK_WAITN mk_wait_all()
#define MK_WAITO(p) do { mk_bar_wait(&sm.bar[(p) & 1], (par >> ((p) & 1)) & 1u); par ^= 1u << ((p) & 1); } while (0)
    MARCH_BUNDLE(MK_RUN, MK_RUNNS, MK_WAITN, MK_WAITO, A, B, sm, R);
#undef MK_RUN
#undef MK_RUNNS
#undef MK_WAITN
#undef MK_WAITO
  }
  }
}

// PML bundles (eqd_march_pml.h): same driver
__global__ void __launch_bounds__(MK_NT, 2) k_march_pml(MarchPmlArgs A) {
  extern __shared__ __align__(128) unsigned char mk_smraw[];
  MarchPmlShared& sm = *reinterpret_cast<MarchPmlShared*>(mk_smraw);
  const int tid = threadIdx.x;
  MarchPmlRegs R;
  unsigned par = 0;
  if (tid == 0) { mk_bar_init(&sm.bar[0]); mk_bar_init(&sm.bar[1]); }
  __syncthreads();
  for (int part = 0; part < 2; ++part) {
  const int* first = part == 0 ? A.ctaFirstA : A.ctaFirstB;
  if (!first) continue;
  const int b0 = first[blockIdx.x], b1 = first[blockIdx.x + 1];
  for (int b = b0; b < b1; ++b) {
    const MarchBundle B = A.rec[b];
#define MK_RUN(body) do { body; __syncthreads(); } while (0)
#define MK_RUNNS(body) do { body; } while (0)
#define MK_WAITN mk_wait_all()
#define MK_WAITO(p) do { mk_bar_wait(&sm.bar[(p) & 1], (par >> ((p) & 1)) & 1u); par ^= 1u << ((p) & 1); } while (0)
    MARCH_PML_BUNDLE(MK_RUN, MK_RUNNS, MK_WAITN, MK_WAITO, A, B, sm, R);
#undef MK_RUN
#undef MK_RUNNS
#undef MK_WAITN
#undef MK_WAITO
  }
  }
}

// Lumped mass of the bundle nodes from the element masses em[8][S] (contm, assembleGlobalMass.f90:376-406),
// summed in a fixed order (x-, then x+ element plane; z-, z+; y-, y+).  One thread per node slot.  A fused
// node's mass is complete and goes to mass[]; the others leave a partial for k_node_mass.
__global__ void __launch_bounds__(128) k_march_mass(const MarchBundle* __restrict__ rec, int nBundles, const int* __restrict__ slotBundle,
                                                    const int* __restrict__ code, const double* __restrict__ em, size_t S, int PFS,
                                                    double* __restrict__ pm, double* __restrict__ mass) {
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= PFS) return;
  const int b = slotBundle[slot / MK_PN];   // node planes never straddle bundles: one entry per plane
  if (b < 0) return;
  const MarchBundle B = rec[b];
  const int k = slot - B.n0, pl = k / MK_PN, i = k - pl * MK_PN, rowlen = mk_rowlen(B), iz = i / rowlen, iy = i - iz * rowlen;
  const int c = code[slot];
  if (c < 0 || (c & MK_GHOST)) { pm[slot] = 0.0; return; }
  const int bz = B.shape & 0xff, by = (B.shape >> 8) & 0xff;
  // corner of the element at (pl - dx, iz - dz, iy - dy) that is this node: signs (dx, dy, dz)
  const int corner[2][2][2] = {{{0, 4}, {3, 7}}, {{1, 5}, {2, 6}}};   // [sx][sy][sz]
  double sum = 0.0;
  for (int dx = 1; dx >= 0; --dx)
    for (int dz = 1; dz >= 0; --dz)
      for (int dy = 1; dy >= 0; --dy) {
        const int p = pl - dx, cz = iz - dz, cy = iy - dy;
        if (p < 0 || p >= B.Lx || cz < 0 || cz >= bz || cy < 0 || cy >= by) continue;
        const size_t e = (size_t)B.e0 + (size_t)p * mk_es(B) + cz * by + cy;
        sum = sum + em[(size_t)corner[dx][dy][dz] * S + e];
      }
  pm[slot] = sum;
  if (mass && (c & MK_FUSED)) mass[c & MK_IDMASK] = sum;
}

size_t march_smem_bytes() { return sizeof(MarchShared) + 128; }
size_t march_pml_smem_bytes() { return sizeof(MarchPmlShared) + 128; }
int march_pml_ctas_per_sm() {
  int n = 0;
  if (cudaFuncSetAttribute(k_march_pml, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)march_pml_smem_bytes()) != cudaSuccess) return 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_march_pml, MK_NT, march_pml_smem_bytes()) != cudaSuccess) return 0;
  return n;
}
void launch_march_pml(const MarchPmlArgs& A, int grid, cudaStream_t s) {
  if (grid <= 0) return;
  k_march_pml<<<grid, MK_NT, march_pml_smem_bytes(), s>>>(A);
}

int march_ctas_per_sm() {
  int n = 0;
  if (cudaFuncSetAttribute(k_march, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)march_smem_bytes()) != cudaSuccess) return 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_march, MK_NT, march_smem_bytes()) != cudaSuccess) return 0;
  return n;
}

void launch_march(const MarchArgs& A, int grid, cudaStream_t s) {
  if (grid <= 0) return;
  k_march<<<grid, MK_NT, march_smem_bytes(), s>>>(A);
}

void launch_march_mass(const MarchBundle* rec, int nBundles, const int* slotBundle, const int* code, const double* em, size_t S, int PFS,
                       double* pm, double* mass, cudaStream_t s) {
  if (PFS > 0) k_march_mass<<<(PFS + 127) / 128, 128, 0, s>>>(rec, nBundles, slotBundle, code, em, S, PFS, pm, mass);
}

// ------------------------------------------------------------------------------------------------
// planner
namespace {

// cells [x0, x0+len) x [z0, z0+bz) x [y0, y0+by) of the lattice box.  gz / gy: the first row / column of the
// rectangle is a GHOST copy of the last row / column of the tile before (which owns those elements); hz / hy: the
// tile after took such a copy of this strip's last row / column, so this strip leaves the nodes of its last node
// row / column to it.  tz, ty: tile of the base (ghost-free) tiling.
struct Strip { int x0, len, z0, bz, y0, by, tz, ty; unsigned char gz, gy, hz, hy, orient; };

// nearly equal parts of at most `cap` cells between consecutive seams of one lattice axis
void cut_axis(int n, const std::vector<char>& seam, int cap, std::vector<std::pair<int, int>>& out) {
  out.clear();
  int a = 0;
  for (int k = 1; k <= n; ++k) {
    if (k < n && !seam[k]) continue;
    const int len = k - a, parts = (len + cap - 1) / cap;
    for (int q = 0; q < parts; ++q) {
      const int b0 = a + (int)((long)len * q / parts), b1 = a + (int)((long)len * (q + 1) / parts);
      out.emplace_back(b0, b1 - b0);
    }
    a = k;
  }
}

}  // namespace

void plan_march(const int* conn, const int* etype, const double* coor, const int* info, const std::vector<int>& elems, int Nn, int ny,
                int nz, int nxg, int grid, int share, MarchPlan& P, bool pml) {
  P = MarchPlan();
  const int n = (int)elems.size();
  const long nynz = (long)ny * nz;
  auto reject_all = [&] { P.leftover = elems; };
  if (n == 0 || ny <= 1 || nz <= 1 || grid <= 0) { reject_all(); return; }
  struct Lap {
    bool on = std::getenv("EQD_VERBOSE") != nullptr;
    std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
    void lap(const char* label) {
      if (!on) return;
      auto now = std::chrono::steady_clock::now();
      std::fprintf(stderr, "[eqd]   plan_march: %s %.3f s\n", label, std::chrono::duration<double>(now - t).count());
      t = now;
    }
  } lap;
  // ---- lattice cell of every candidate: the grid position of corner 7 (+,+,+) minus one; the y+ face
  // (corners 3, 4, 7, 8) must be the lattice nodes of that cell (the y- face may hold split-node masters)
  raw_vector<int> cx(n), cy(n), cz(n);
  parallel_range((size_t)n, [&](size_t b, size_t e) {
    for (size_t j = b; j < e; ++j) {
      const int el = elems[j];
      const int* c = conn + 8 * (size_t)el;
      cx[j] = -1;
      if (etype[el] == 11 || etype[el] == 12) continue;
      const long id6 = c[6];
      const int ix = (int)(id6 / nynz), iz = (int)((id6 % nynz) / ny), iy = (int)(id6 % ny);
      if (ix < 1 || iz < 1 || iy < 1) continue;
      auto id = [&](int dx, int dz, int dy) { return (long)(ix - 1 + dx) * nynz + (long)(iz - 1 + dz) * ny + (iy - 1 + dy); };
      if (c[2] != id(1, 0, 1) || c[3] != id(0, 0, 1) || c[7] != id(0, 1, 1)) continue;
      bool free3 = true;   // regular bundles: 3-dof nodes only; PML bundles: type-2 elements on any nodes
      for (int k = 0; k < 8 && !pml; ++k) free3 = free3 && EQD_INFO_KIND(info[c[k]]) != KIND_PML12;
      if (pml && etype[el] != 2) continue;
      if (!free3 || !box_element(c, coor)) continue;
      cx[j] = ix - 1; cz[j] = iz - 1; cy[j] = iy - 1;
    }
  });
  lap.lap("lattice cells + box test");
  int lo[3] = {1 << 30, 1 << 30, 1 << 30}, hi[3] = {-1, -1, -1};
  {
    std::mutex mu;
    parallel_range((size_t)n, [&](size_t b, size_t e) {
      int l[3] = {1 << 30, 1 << 30, 1 << 30}, h3[3] = {-1, -1, -1};
      for (size_t j = b; j < e; ++j) {
        if (cx[j] < 0) continue;
        l[0] = std::min(l[0], cx[j]); h3[0] = std::max(h3[0], cx[j]);
        l[1] = std::min(l[1], cz[j]); h3[1] = std::max(h3[1], cz[j]);
        l[2] = std::min(l[2], cy[j]); h3[2] = std::max(h3[2], cy[j]);
      }
      std::lock_guard<std::mutex> g(mu);
      for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], l[k]); hi[k] = std::max(hi[k], h3[k]); }
    });
  }
  if (hi[0] < 0) { reject_all(); return; }
  const int ex = hi[0] - lo[0] + 1, ez = hi[1] - lo[1] + 1, ey = hi[2] - lo[2] + 1;
  if ((double)ex * ez * ey > 1.5e9) { reject_all(); return; }
  raw_vector<int> cell((size_t)ex * ez * ey);   // lattice cell -> index into elems, -1 = no candidate
  parallel_range(cell.size(), [&](size_t b, size_t e) { std::fill(cell.begin() + b, cell.begin() + e, -1); });
  auto at = [&](int x, int z, int y) -> int& { return cell[((size_t)x * ez + z) * ey + y]; };
  parallel_range((size_t)n, [&](size_t b, size_t e) {   // distinct cells: no two candidates share one
    for (size_t j = b; j < e; ++j)
      if (cx[j] >= 0) at(cx[j] - lo[0], cz[j] - lo[1], cy[j] - lo[2]) = (int)j;
  });
  lap.lap("cell map");
  // ---- where neighbouring candidates do not share their four nodes (the fault: its + side references the
  // split-node masters): per cell, bit 0 / 1 / 2 = mismatch with the y+ / z+ / x+ neighbour
  raw_vector<unsigned char> mis(cell.size());
  std::vector<char> seamZ(ez + 1, 0), seamY(ey + 1, 0), edgeZ(ez + 1, 0), edgeY(ey + 1, 0);
  {
    const int nth = host_threads();
    std::vector<std::vector<char>> sz(nth, std::vector<char>(ez + 1, 0)), sy(nth, std::vector<char>(ey + 1, 0));
    std::vector<std::vector<char>> bz2(nth, std::vector<char>(ez + 1, 0)), by2(nth, std::vector<char>(ey + 1, 0));
    std::vector<std::thread> th;
    const int per = (ex + nth - 1) / nth;
    for (int t = 0; t < nth; ++t)
      th.emplace_back([&, t] {
        for (int x = t * per; x < std::min(ex, (t + 1) * per); ++x)
          for (int z = 0; z < ez; ++z)
            for (int y = 0; y < ey; ++y) {
              const int j = at(x, z, y);
              unsigned char m = 0;
              if (j >= 0) {
                const int* c = conn + 8 * (size_t)elems[j];
                // lattice planes where the candidate set ends (a PML slab against the interior, say)
                if (y + 1 < ey && at(x, z, y + 1) < 0) by2[t][y + 1] = 1;
                if (y > 0 && at(x, z, y - 1) < 0) by2[t][y] = 1;
                if (z + 1 < ez && at(x, z + 1, y) < 0) bz2[t][z + 1] = 1;
                if (z > 0 && at(x, z - 1, y) < 0) bz2[t][z] = 1;
                if (y + 1 < ey && at(x, z, y + 1) >= 0) {
                  const int* d = conn + 8 * (size_t)elems[at(x, z, y + 1)];
                  if (c[3] != d[0] || c[2] != d[1] || c[7] != d[4] || c[6] != d[5]) { m |= 1; sy[t][y + 1] = 1; }
                }
                if (z + 1 < ez && at(x, z + 1, y) >= 0) {
                  const int* d = conn + 8 * (size_t)elems[at(x, z + 1, y)];
                  if (c[4] != d[0] || c[5] != d[1] || c[6] != d[2] || c[7] != d[3]) { m |= 2; sz[t][z + 1] = 1; }
                }
                if (x + 1 < ex && at(x + 1, z, y) >= 0) {
                  const int* d = conn + 8 * (size_t)elems[at(x + 1, z, y)];
                  if (c[1] != d[0] || c[2] != d[3] || c[5] != d[4] || c[6] != d[7]) m |= 4;
                }
              }
              mis[((size_t)x * ez + z) * ey + y] = m;
            }
      });
    for (auto& x : th) x.join();
    for (int t = 0; t < nth; ++t) {
      for (int k = 0; k <= ez; ++k) { seamZ[k] |= sz[t][k]; edgeZ[k] |= bz2[t][k]; }
      for (int k = 0; k <= ey; ++k) { seamY[k] |= sy[t][k]; edgeY[k] |= by2[t][k]; }
    }
  }
  lap.lap("neighbour joins");
  // A planar fault along a lattice plane (vertical strike-slip faults: y = 0) shows as ONE such plane: the column
  // tiles are then cut there.  A dipping fault crosses many planes; tiles then stay on the plain lattice and the
  // local checks below reject the stretches of a tile that the fault runs through (the tile kernels sweep those).
  auto few = [](std::vector<char>& seam) {
    int cnt = 0;
    for (char c : seam) cnt += c != 0;
    if (cnt > 4) std::fill(seam.begin(), seam.end(), 0);
  };
  few(seamZ); few(seamY);
  // the same for the planes where the candidate set ends: a PML slab is six cells thick, tiles that straddle its inner
  // face would never be complete; an irregular outline (a dipping fault zone) gives many such planes and is left alone
  few(edgeZ); few(edgeY);
  for (int k = 0; k <= ez; ++k) seamZ[k] |= edgeZ[k];
  for (int k = 0; k <= ey; ++k) seamY[k] |= edgeY[k];
  // ---- column tiles and their strips: runs in x of complete cross-sections whose cells share all their nodes
  std::vector<std::pair<int, int>> zt, yt, ztB;
  // base tiles are one cell short of the kernel's cross-section: room for the ghost row / column
  cut_axis(ez, seamZ, (share & 2) ? MK_BZ - 1 : MK_BZ, zt);   // share: bit 0 = ghost columns (y), bit 1 = ghost rows (z)
  cut_axis(ey, seamY, (share & 1) ? MK_BY - 1 : MK_BY, yt);
  // PML bundles: a y segment no wider than 7 cells (a slab) is swept with the node plane turned -- 16 rows of 8 nodes, up
  // to 15 x 7 element columns -- so that the threads are filled along z instead
  std::vector<char> turned(yt.size(), 0);
  if (pml && !share) {
    cut_axis(ez, seamZ, MK_BY, ztB);
    for (size_t k = 0; k < yt.size(); ++k) turned[k] = yt[k].second <= MK_BZ && ztB.size() < zt.size();
  }
  struct TileRef { int zi, yi; };
  std::vector<TileRef> tiles;
  if (share || !pml) {
    for (size_t zi = 0; zi < zt.size(); ++zi) for (size_t yi = 0; yi < yt.size(); ++yi) tiles.push_back({(int)zi, (int)yi});   // index = zi*nty + yi
  } else {
    for (size_t yi = 0; yi < yt.size(); ++yi)
      for (size_t zi = 0; zi < (turned[yi] ? ztB.size() : zt.size()); ++zi) tiles.push_back({(int)zi, (int)yi});
  }
  std::vector<Strip> strips;
  {
    const int nT = (int)tiles.size();
    std::vector<std::vector<Strip>> per(nT);
    parallel_range((size_t)nT, [&](size_t tb, size_t te) {
      for (size_t t = tb; t < te; ++t) {
        const unsigned char orient = turned[tiles[t].yi];
        const auto& Z = (orient ? ztB : zt)[tiles[t].zi];
        const auto& Y = yt[tiles[t].yi];
        const int z1 = Z.first + Z.second, y1 = Y.first + Y.second;
        int run0 = -1;
        for (int x = 0; x <= ex; ++x) {
          bool full = x < ex, joined = true;   // joined: every cell shares its x- face with the plane before
          for (int z = Z.first; full && z < z1; ++z)
            for (int y = Y.first; full && y < y1; ++y) {
              full = at(x, z, y) >= 0;
              const unsigned char m = mis[((size_t)x * ez + z) * ey + y];
              if ((m & 1) && y + 1 < y1) full = false;
              if ((m & 2) && z + 1 < z1) full = false;
              if (x > 0 && (mis[((size_t)(x - 1) * ez + z) * ey + y] & 4)) joined = false;
            }
          if (run0 >= 0 && (!full || !joined)) {
            if (x - run0 >= MK_MINLX)
              per[t].push_back({run0, x - run0, Z.first, Z.second, Y.first, Y.second, tiles[t].zi, tiles[t].yi, 0, 0, 0, 0, orient});
            run0 = -1;
          }
          if (full && run0 < 0) run0 = x;
        }
      }
    }, 1);
    // ---- ghost sharing (option): a strip whose y- (then z-) neighbour tile has a strip over exactly the same x range
    // takes a private copy of that strip's last element column (row): the node column (row) between the two tiles is
    // then interior to this strip -- complete inside one CTA, updated by it -- instead of two partial sums that the node
    // update has to gather.  Both copies of a shared element see the same nodal values and run the same instruction
    // sequence, so their stresses stay bit-identical; only the owner's copy is ever fetched.
    if (share) {
      const int nty = (int)yt.size(), ntz = (int)zt.size();
      // An interface between two tile columns (rows) is shared as a whole or not at all: every pair of tiles across it
      // must hold strips over the same x ranges whose facing cells are joined.  That makes the choices around every tile
      // corner agree (the corner element is then reported to the corner node by exactly one of the four tiles).
      auto same_lists = [](const std::vector<Strip>& a, const std::vector<Strip>& b) {
        if (a.size() != b.size() || a.empty()) return false;
        for (size_t k = 0; k < a.size(); ++k) if (a[k].x0 != b[k].x0 || a[k].len != b[k].len) return false;
        return true;
      };
      std::vector<char> yShared(nty + 1, 0), zShared(ntz + 1, 0);
      for (int ty = 1; ty < nty; ++ty) {
        bool ok = (share & 1) != 0;
        for (int tz = 0; ok && tz < ntz; ++tz) {
          const auto& a = per[tz * nty + ty - 1];
          const auto& b = per[tz * nty + ty];
          if (a.empty() && b.empty()) continue;
          ok = same_lists(a, b);
          for (size_t k = 0; ok && k < b.size(); ++k)
            for (int x = b[k].x0; ok && x < b[k].x0 + b[k].len; ++x)
              for (int z = b[k].z0; ok && z < b[k].z0 + b[k].bz; ++z) ok = !(mis[((size_t)x * ez + z) * ey + b[k].y0 - 1] & 1);
        }
        yShared[ty] = ok;
      }
      for (int tz = 1; tz < ntz; ++tz) {
        bool ok = (share & 2) != 0;
        for (int ty = 0; ok && ty < nty; ++ty) {
          const auto& a = per[(tz - 1) * nty + ty];
          const auto& b = per[tz * nty + ty];
          if (a.empty() && b.empty()) continue;
          ok = same_lists(a, b);
          for (size_t k = 0; ok && k < b.size(); ++k)
            for (int x = b[k].x0; ok && x < b[k].x0 + b[k].len; ++x)
              for (int y = b[k].y0; ok && y < b[k].y0 + b[k].by; ++y) ok = !(mis[((size_t)x * ez + b[k].z0 - 1) * ey + y] & 2);
        }
        zShared[tz] = ok;
      }
      for (int t = 0; t < nT; ++t)
        for (Strip& q : per[t]) { q.gy = yShared[q.ty]; q.gz = zShared[q.tz]; q.hy = yShared[q.ty + 1]; q.hz = zShared[q.tz + 1]; }
      for (int t = 0; t < nT; ++t)
        for (Strip& q : per[t]) { q.z0 -= q.gz; q.bz += q.gz; q.y0 -= q.gy; q.by += q.gy; }
    }
    for (auto& v : per) strips.insert(strips.end(), v.begin(), v.end());
  }
  lap.lap("tiles + strips");
  long total = 0;
  for (const Strip& s : strips) total += s.len;
  if (total == 0) { reject_all(); return; }
  // ---- two work lists.  A: what may touch a rank face -- strips along the y / z boundary of the node grid, and
  // short caps (MK_MINLX planes) where a strip reaches the grid's first or last node plane in x (the domain's outer
  // boundary is PML, so in practice only rank faces and the free surface qualify).  B: the rest.  With rank
  // neighbours the step sweeps A first and exchanges the face forces while B is swept (eqd_api.cu).
  struct Piece { int strip, a, len; };
  std::vector<Piece> listA, listB;
  {
    const int gx1 = nxg > 1 ? nxg - 2 : -1;   // last cell of the node grid in x
    for (size_t k = 0; k < strips.size(); ++k) {
      const Strip& s = strips[k];
      const int z0 = lo[1] + s.z0, y0 = lo[2] + s.y0, x0 = lo[0] + s.x0;
      const bool edge = z0 == 0 || y0 == 0 || z0 + s.bz == nz - 1 || y0 + s.by == ny - 1;
      if (edge) { listA.push_back({(int)k, 0, s.len}); continue; }
      int a = 0, b = s.len;
      const bool capLo = x0 == 0 && s.len >= 3 * MK_MINLX, capHi = x0 + s.len - 1 == gx1 && s.len >= 3 * MK_MINLX;
      if (x0 == 0 && !capLo) { listA.push_back({(int)k, 0, s.len}); continue; }          // too short to cap: all of it
      if (x0 + s.len - 1 == gx1 && !capHi) { listA.push_back({(int)k, 0, s.len}); continue; }
      if (capLo) { listA.push_back({(int)k, 0, MK_MINLX}); a = MK_MINLX; }
      if (capHi) { listA.push_back({(int)k, s.len - MK_MINLX, MK_MINLX}); b = s.len - MK_MINLX; }
      listB.push_back({(int)k, a, b - a});
    }
  }
  // ---- balanced static schedule of one list: CTA c marches the element planes [c*total/grid, (c+1)*total/grid) of the
  // piece sequence; a cut closer than MK_MINLX to a piece end moves to that end
  P.grid = grid;
  auto schedule = [&](const std::vector<Piece>& pieces, std::vector<int>& ctaFirst) {
    const int first = (int)P.rec.size();
    ctaFirst.assign(grid + 1, first);
    long tot = 0;
    for (const Piece& q : pieces) tot += q.len;
    if (tot == 0) return;
    std::vector<long> cutpos;   // snapped global plane positions where CTA c = 1 .. grid-1 starts (non-decreasing)
    long base = 0, prev = 0;
    size_t si = 0;
    for (int c = 1; c < grid; ++c) {
      const long g = tot * c / grid;
      while (si + 1 < pieces.size() && g >= base + pieces[si].len) { base += pieces[si].len; ++si; }
      long off = std::min<long>(g - base, pieces[si].len);
      if (off < MK_MINLX) off = 0;
      else if (pieces[si].len - off < MK_MINLX) off = pieces[si].len;
      long sgl = std::max(base + off, prev);
      if (sgl > prev && sgl - prev < MK_MINLX && prev > base) sgl = prev;   // two cuts of one piece too close: the CTA stays empty
      cutpos.push_back(sgl);
      prev = sgl;
    }
    std::vector<int> ctaOf;
    size_t ci = 0;
    base = 0;
    for (const Piece& q : pieces) {
      const Strip& st = strips[q.strip];
      int a = 0;
      while (a < q.len) {
        while (ci < cutpos.size() && cutpos[ci] <= base + a) ++ci;
        int end = q.len;
        if (ci < cutpos.size() && cutpos[ci] < base + q.len) end = (int)(cutpos[ci] - base);
        MarchBundle B{};
        B.Lx = end - a;
        B.shape = st.bz | (st.by << 8) | ((int)st.orient << 16);
        B.e0 = st.x0 + q.a + a;   // provisional: lattice x of the first plane (replaced by the slot below)
        B.n0 = q.strip;           // provisional: strip
        P.rec.push_back(B);
        ctaOf.push_back((int)ci);
        a = end;
      }
      base += q.len;
    }
    size_t bi = 0;
    for (int c = 0; c <= grid; ++c) {
      while (bi < ctaOf.size() && ctaOf[bi] < c) ++bi;
      ctaFirst[c] = first + (int)bi;
    }
    ctaFirst[grid] = (int)P.rec.size();
  };
  schedule(listA, P.ctaFirstA);
  P.nBundlesA = (int)P.rec.size();
  schedule(listB, P.ctaFirstB);
  lap.lap("schedule");
  // ---- slots
  const int nB = (int)P.rec.size();
  std::vector<int> stripOf(nB), xOf(nB);
  long eslot = 0, nslot = 0;
  for (int b = 0; b < nB; ++b) {
    stripOf[b] = P.rec[b].n0; xOf[b] = P.rec[b].e0;
    P.rec[b].e0 = (int)eslot; P.rec[b].n0 = (int)nslot;
    eslot += (long)P.rec[b].Lx * mk_es(P.rec[b]);
    nslot += (long)(P.rec[b].Lx + 1) * MK_PN;
    if (eslot > (1L << 29) || nslot > (1L << 30)) throw std::runtime_error("march planner: class too large for 32-bit slots");
  }
  P.S = (int)std::max(eslot, 32L);
  P.PFS = (int)std::max(nslot, 4L);
  P.refId.resize(P.S); P.owner.resize(P.S); P.code.resize(P.PFS); P.slotBundle.assign((size_t)(P.PFS + MK_PN - 1) / MK_PN, -1);
  parallel_range((size_t)P.S, [&](size_t b, size_t e) { std::fill(P.refId.begin() + b, P.refId.begin() + e, -1); std::fill(P.owner.begin() + b, P.owner.begin() + e, (unsigned char)0); });
  parallel_range((size_t)P.PFS, [&](size_t b, size_t e) { std::fill(P.code.begin() + b, P.code.begin() + e, -1); });
  std::vector<char> taken(n, 0);
  std::vector<long> fusedT(host_threads() + 1, 0), elemsT(host_threads() + 1, 0);
  std::vector<int> bad(1, 0);
  const bool fullCheck = std::getenv("EQD_MARCH_CHECK") != nullptr;
  {
    const int nth = host_threads();
    std::vector<std::thread> th;
    std::atomic<int> next(0);   // bundles differ in length: hand them out one by one
    for (int t = 0; t < nth; ++t)
      th.emplace_back([&, t] {
        const int corner[2][2][2] = {{{0, 4}, {3, 7}}, {{1, 5}, {2, 6}}};   // [sx][sy][sz]
        for (int b = next.fetch_add(1); b < nB; b = next.fetch_add(1)) {
          const MarchBundle& B = P.rec[b];
          const Strip& s = strips[stripOf[b]];
          const int bz = s.bz, by = s.by, x0 = xOf[b], rowlen = mk_rowlen(B);
          for (int p = 0; p < B.Lx; ++p)
            for (int z = 0; z < bz; ++z)
              for (int y = 0; y < by; ++y) {
                const int j = at(x0 + p, s.z0 + z, s.y0 + y);
                const size_t slot = (size_t)B.e0 + (size_t)p * mk_es(B) + z * by + y;
                P.refId[slot] = elems[j];
                const bool ghost = (z == 0 && s.gz) || (y == 0 && s.gy);     // the tile before owns this element
                P.owner[slot] = ghost ? 0 : 1;
                if (!ghost) { taken[j] = 1; elemsT[t]++; }
              }
          for (int pl = 0; pl <= B.Lx; ++pl) {
            P.slotBundle[(size_t)B.n0 / MK_PN + pl] = b;
            for (int iz = 0; iz <= bz; ++iz)
              for (int iy = 0; iy <= by; ++iy) {
                // the node id as any adjacent element of the strip has it (the joins between neighbouring cells were
                // checked when the strips were made; EQD_MARCH_CHECK=1 compares all of them again)
                int id = -1;
                for (int dx = 0; dx < 2 && (id < 0 || fullCheck); ++dx)
                  for (int dz = 0; dz < 2 && (id < 0 || fullCheck); ++dz)
                    for (int dy = 0; dy < 2 && (id < 0 || fullCheck); ++dy) {
                      const int p = pl - dx, z = iz - dz, y = iy - dy;
                      if (p < 0 || p >= B.Lx || z < 0 || z >= bz || y < 0 || y >= by) continue;
                      const int nd = conn[8 * (size_t)elems[at(x0 + p, s.z0 + z, s.y0 + y)] + corner[dx][dy][dz]];
                      if (id >= 0 && nd != id) bad[0] = 1;
                      id = nd;
                    }
                if (id < 0 || id >= Nn || id > MK_IDMASK) { bad[0] = 1; continue; }
                // a node of the first / last row or column is this strip's to report unless a neighbour tile shares it
                const bool outZ = iz == 0 ? !s.gz : (iz == bz ? !s.hz : true), outY = iy == 0 ? !s.gy : (iy == by ? !s.hy : true);
                const bool interior = pl > 0 && pl < B.Lx && iz > 0 && iz < bz && iy > 0 && iy < by;
                const bool fused = !pml && interior && EQD_INFO_KIND(info[id]) == KIND_FREE3;
                if (fused) fusedT[t]++;
                P.code[(size_t)B.n0 + (size_t)pl * MK_PN + iz * rowlen + iy] = id | (fused ? MK_FUSED : 0) | (outZ && outY ? 0 : MK_GHOST);
              }
          }
        }
      });
    for (auto& x : th) x.join();
  }
  lap.lap("slots, codes");
  if (bad[0]) throw std::runtime_error("march planner: the elements of a bundle disagree on a lattice node (internal)");
  if (share) {
    // every (node, adjacent element) pair of the elements in bundles must be reported exactly once: count, per node,
    // the corners the owned elements put on it, and the computed elements behind every slot that reports it
    std::vector<int> want(Nn, 0), got(Nn, 0);
    for (int b = 0; b < nB; ++b) {
      const MarchBundle& B = P.rec[b];
      const int bz = mk_bz(B), by = mk_by(B);
      for (int p = 0; p < B.Lx; ++p)
        for (int z = 0; z < bz; ++z)
          for (int y = 0; y < by; ++y) {
            const size_t slot = (size_t)B.e0 + (size_t)p * mk_es(B) + z * by + y;
            if (P.owner[slot]) for (int k = 0; k < 8; ++k) want[conn[8 * (size_t)P.refId[slot] + k]]++;
          }
      for (int pl = 0; pl <= B.Lx; ++pl)
        for (int iz = 0; iz <= bz; ++iz)
          for (int iy = 0; iy <= by; ++iy) {
            const int c = P.code[(size_t)B.n0 + (size_t)pl * MK_PN + iz * mk_rowlen(B) + iy];
            if (c < 0 || (c & MK_GHOST)) continue;
            const int nx_ = (pl > 0) + (pl < B.Lx), nz_ = (iz > 0) + (iz < bz), ny_ = (iy > 0) + (iy < by);
            got[c & MK_IDMASK] += nx_ * nz_ * ny_;
          }
    }
    bool okc = true;
    for (int nd = 0; nd < Nn && okc; ++nd) okc = want[nd] == got[nd];
    if (!okc && std::getenv("EQD_MARCH_DEBUG")) {
      int shown = 0;
      for (int nd = 0; nd < Nn && shown < 12; ++nd)
        if (want[nd] != got[nd]) {
          std::fprintf(stderr, "  node %d (ix %ld iz %ld iy %ld; lattice x %ld z %ld y %ld): want %d got %d\n", nd, nd / nynz, (nd % nynz) / ny, (long)nd % ny,
                       nd / nynz - lo[0], (nd % nynz) / ny - lo[1], (long)nd % ny - lo[2], want[nd], got[nd]);
          ++shown;
        }
      for (size_t k = 0; k < strips.size() && k < 12; ++k)
        std::fprintf(stderr, "  strip %zu: x %d+%d z %d+%d y %d+%d tile (%d,%d) g %d%d h %d%d\n", k, strips[k].x0, strips[k].len, strips[k].z0, strips[k].bz,
                     strips[k].y0, strips[k].by, strips[k].tz, strips[k].ty, strips[k].gz, strips[k].gy, strips[k].hz, strips[k].hy);
    }
    if (!okc) {   // cannot happen on lattices the sharing rules were made for; stay correct on the others
      if (std::getenv("EQD_VERBOSE")) std::fprintf(stderr, "[eqd]   plan_march: ghost sharing inconsistent on this mesh, planning without it\n");
      plan_march(conn, etype, coor, info, elems, Nn, ny, nz, nxg, grid, 0, P, pml);
      return;
    }
  }
  for (long v : fusedT) P.nFused += v;
  for (long v : elemsT) P.n += (int)v;
  P.leftover.reserve(n - P.n);
  for (int j = 0; j < n; ++j) if (!taken[j]) P.leftover.push_back(elems[j]);
  lap.lap("check + leftover");
  if (std::getenv("EQD_VERBOSE"))
    std::fprintf(stderr, "[eqd]   plan_march: %d of %d regular elements in %d bundles (%d of them boundary work; %zu strips, lattice %dx%dx%d, "
                         "%zu x %zu column tiles), %ld element planes over %d CTAs, %ld fused nodes of %ld node slots\n",
                 P.n, n, nB, P.nBundlesA, strips.size(), ex, ez, ey, zt.size(), yt.size(), total, grid, P.nFused, nslot);
}

}  // namespace eqd

// ------------------------------------------------------------------------------------------------
// Host-only self-check of the marching planner AND of the kernel source (no GPU): plans the bundles of a
// sub-domain, then runs the phases of eqd_march.h over all thread ids, CTA by CTA, on the given nodal
// fields for ONE step, and returns what the device would have produced: updated stresses (reference element
// order), updated v / d of the fused nodes, partial forces summed per node into fsum (3,Nn), fused flags.
// stats[8]: elements in bundles, bundles, node slots, fused nodes, leftover regular elements, grid, S, 0.
extern "C" int eqd_march_emulate(int32_t Nn, int32_t Ne, const double* meshCoor, const int32_t* nodeElemIdRelation,
                                 const int32_t* elemTypeArr, const int32_t* numOfDofPerNodeArr, int32_t grid, const double* eleshp,
                                 const double* ss, const double* eledet, const double* mat, double* stress6 /* (6,Ne) in/out */,
                                 double* vel /* (3,Nn) in/out */, double* disp /* (3,Nn) in/out */, const double* mass /* (Nn) */,
                                 double dt, double rdampk, double w, int32_t update, double* fsum /* (3,Nn) out */,
                                 int32_t* fusedFlag /* (Nn) out */, int32_t* inBundle /* (Ne) out */, int64_t* stats) {
  using namespace eqd;
  try {
    std::vector<int> conn(8 * (size_t)Ne), info(Nn), elems;
    for (size_t k = 0; k < conn.size(); ++k) conn[k] = nodeElemIdRelation[k] - 1;
    for (int nd = 0; nd < Nn; ++nd) info[nd] = numOfDofPerNodeArr[nd] == 12 ? KIND_PML12 : KIND_FREE3;
    for (int e = 0; e < Ne; ++e) {
      if (elemTypeArr[e] == 2) continue;
      bool reg = true;
      for (int k = 0; k < 8; ++k) reg = reg && numOfDofPerNodeArr[conn[8 * (size_t)e + k]] == 3;
      if (reg) elems.push_back(e);
    }
    int ny = 0, nz = 0;
    if (!infer_grid(conn.data(), elemTypeArr, Ne, Nn, ny, nz)) return __LINE__;
    MarchPlan P;
    int nxg = 0;
    for (int e = 0; e < Ne; ++e) nxg = std::max(nxg, (int)(conn[8 * (size_t)e + 6] / ((long)ny * nz)) + 1);
    plan_march(conn.data(), elemTypeArr, meshCoor, info.data(), elems, Nn, ny, nz, nxg, grid & 0xffff, (grid >> 16) & 3, P, false);
    stats[0] = P.n; stats[1] = (int64_t)P.rec.size(); stats[2] = P.PFS; stats[3] = P.nFused; stats[4] = (int64_t)P.leftover.size();
    stats[5] = P.grid; stats[6] = P.S; stats[7] = P.nBundlesA;
    for (int e = 0; e < Ne; ++e) inBundle[e] = 0;
    for (int nd = 0; nd < Nn; ++nd) fusedFlag[nd] = 0;
    if (P.n == 0) return 0;
    const size_t S = P.S, NnS = ((size_t)Nn + 31) / 32 * 32, PFS = P.PFS;
    // class SoA as the library uploads it
    std::vector<double> a(3 * S, 0.0), s3(3 * S, 0.0), lam(S, 0.0), mu(S, 0.0), det(S, 1.0), sg(6 * S, 0.0);
    for (size_t s = 0; s < S; ++s) {
      const int e = P.refId[s];
      if (e < 0) continue;
      if (P.owner[s]) inBundle[e] += 1;
      a[s] = eleshp[BOX_AX + 24 * (size_t)e]; a[S + s] = eleshp[BOX_AY + 24 * (size_t)e]; a[2 * S + s] = eleshp[BOX_AZ + 24 * (size_t)e];
      s3[s] = ss[0 + 6 * (size_t)e]; s3[S + s] = ss[3 + 6 * (size_t)e]; s3[2 * S + s] = ss[5 + 6 * (size_t)e];
      lam[s] = mat[(size_t)e + 3 * (size_t)Ne]; mu[s] = mat[(size_t)e + 4 * (size_t)Ne]; det[s] = eledet[e];
      for (int k = 0; k < 6; ++k) sg[(size_t)k * S + s] = stress6[k + 6 * (size_t)e];
    }
    std::vector<double> v(3 * NnS, 0.0), d(3 * NnS, 0.0), pf(3 * PFS, 0.0), force(3 * NnS, 0.0);
    for (int nd = 0; nd < Nn; ++nd)
      for (int c = 0; c < 3; ++c) { v[c * NnS + nd] = vel[c + 3 * (size_t)nd]; d[c * NnS + nd] = disp[c + 3 * (size_t)nd]; }
    std::vector<double> vOut(v), dOut(d);   // the second buffer of the library's pair: updated nodes go here
    StepState st{};
    MarchArgs A{};
    A.rec = P.rec.data(); A.ctaFirstA = P.ctaFirstA.data(); A.ctaFirstB = P.ctaFirstB.data(); A.code = P.code.data();
    A.S = S; A.NnS = NnS; A.PFS = PFS;
    A.a = a.data(); A.ss = s3.data(); A.lam = lam.data(); A.mu = mu.data(); A.det = det.data(); A.stress = sg.data();
    A.vel = v.data(); A.disp = d.data(); A.velOut = vOut.data(); A.dispOut = dOut.data(); A.mass = mass; A.pf = pf.data(); A.force = force.data();
    A.dt = dt; A.rdampk = rdampk; A.w = w; A.update = update; A.st = &st;
    // the kernel reads v and d of nodes that other CTAs update in the same launch (ghost sharing): the updates go to
    // the second buffer.  The host reading runs the CTAs one after the other.
    std::vector<MarchShared> smv(1);
    MarchShared& sm = smv[0];
    std::vector<MarchRegs> regs(MK_NT);
    for (int cta = 0; cta < P.grid; ++cta)
      for (int part = 0; part < 2; ++part)
      for (int b = (part ? P.ctaFirstB : P.ctaFirstA)[cta]; b < (part ? P.ctaFirstB : P.ctaFirstA)[cta + 1]; ++b) {
        const MarchBundle B = P.rec[b];
#define MK_RUN(body) do { for (int tid = 0; tid < MK_NT; ++tid) { MarchRegs& R = regs[tid]; (void)R; body; } } while (0)
#define MK_WAITN ((void)0)
#define MK_WAITO(p) ((void)0)
        MARCH_BUNDLE(MK_RUN, MK_RUN, MK_WAITN, MK_WAITO, A, B, sm, R);
#undef MK_RUN
#undef MK_WAITN
#undef MK_WAITO
      }
    if (st.nanFlag) return __LINE__;
    for (size_t s = 0; s < S; ++s) {      // the owner's copy is the element's stress ...
      const int e = P.refId[s];
      if (e < 0 || !P.owner[s]) continue;
      for (int k = 0; k < 6; ++k) stress6[k + 6 * (size_t)e] = sg[(size_t)k * S + s];
    }
    for (size_t s = 0; s < S; ++s) {      // ... and every ghost copy carries the same bits
      const int e = P.refId[s];
      if (e < 0 || P.owner[s]) continue;
      for (int k = 0; k < 6; ++k) if (stress6[k + 6 * (size_t)e] != sg[(size_t)k * S + s]) return __LINE__;
    }
    for (int e = 0; e < Ne; ++e) if (inBundle[e] > 1) return __LINE__;     // one owner per element
    for (int nd = 0; nd < Nn; ++nd)
      for (int c = 0; c < 3; ++c) { vel[c + 3 * (size_t)nd] = vOut[c * NnS + nd]; disp[c + 3 * (size_t)nd] = dOut[c * NnS + nd]; fsum[c + 3 * (size_t)nd] = 0.0; }
    for (size_t slot = 0; slot < PFS; ++slot) {
      const int code = P.code[slot];
      if (code < 0 || (code & MK_GHOST)) continue;
      const int nd = code & MK_IDMASK;
      if (code & MK_FUSED) {
        if (fusedFlag[nd]) return __LINE__;          // a node is fused by one bundle only
        fusedFlag[nd] = 1;
        for (int c = 0; c < 3; ++c) fsum[c + 3 * (size_t)nd] += force[c * NnS + nd];
      } else {
        for (int c = 0; c < 3; ++c) fsum[c + 3 * (size_t)nd] += pf[c * PFS + slot];
      }
    }
    return 0;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "eqd_march_emulate: %s\n", e.what());
    return __LINE__;
  }
}

// ------------------------------------------------------------------------------------------------
// The same self-check for the PML bundles (eqd_march_pml.h): plans the bundles of the type-2 elements and runs the
// kernel's phases on the host for one step.  damps(3,Ne) is the damping profile at the element centroids,
// stress21(21,Ne) the split stresses (slots 1..21 of every PML element, in/out); f12(12,Nn) receives the twelve
// partial force rows summed per node.  stats[8] as eqd_march_emulate.
extern "C" int eqd_march_pml_emulate(int32_t Nn, int32_t Ne, const double* meshCoor, const int32_t* nodeElemIdRelation,
                                     const int32_t* elemTypeArr, const int32_t* numOfDofPerNodeArr, int32_t grid, const double* eleshp,
                                     const double* ss, const double* eledet, const double* mat, const double* damps, double* stress21,
                                     const double* vel, const double* disp, double dt, double rdampk, double w, double* f12,
                                     int32_t* inBundle, int64_t* stats) {
  using namespace eqd;
  try {
    std::vector<int> conn(8 * (size_t)Ne), info(Nn), elems;
    for (size_t k = 0; k < conn.size(); ++k) conn[k] = nodeElemIdRelation[k] - 1;
    for (int nd = 0; nd < Nn; ++nd) info[nd] = numOfDofPerNodeArr[nd] == 12 ? KIND_PML12 : KIND_FREE3;
    for (int e = 0; e < Ne; ++e) if (elemTypeArr[e] == 2) elems.push_back(e);
    int ny = 0, nz = 0;
    if (!infer_grid(conn.data(), elemTypeArr, Ne, Nn, ny, nz)) return __LINE__;
    int nxg = 0;
    for (int e = 0; e < Ne; ++e) nxg = std::max(nxg, (int)(conn[8 * (size_t)e + 6] / ((long)ny * nz)) + 1);
    MarchPlan P;
    plan_march(conn.data(), elemTypeArr, meshCoor, info.data(), elems, Nn, ny, nz, nxg, grid, 0, P, true);
    stats[0] = P.n; stats[1] = (int64_t)P.rec.size(); stats[2] = P.PFS; stats[3] = P.nFused; stats[4] = (int64_t)P.leftover.size();
    stats[5] = P.grid; stats[6] = P.S; stats[7] = P.nBundlesA;
    for (int e = 0; e < Ne; ++e) inBundle[e] = 0;
    for (size_t k = 0; k < 12 * (size_t)Nn; ++k) f12[k] = 0.0;
    if (P.n == 0) return 0;
    if (P.nFused != 0) return __LINE__;
    const size_t S = P.S, NnS = ((size_t)Nn + 31) / 32 * 32, PFS = P.PFS;
    std::vector<double> a(3 * S, 0.0), s3(3 * S, 0.0), lam(S, 0.0), mu(S, 0.0), det(S, 1.0), dm(3 * S, 0.0), sg(21 * S, 0.0);
    for (size_t s = 0; s < S; ++s) {
      const int e = P.refId[s];
      if (e < 0) continue;
      inBundle[e] += 1;
      a[s] = eleshp[BOX_AX + 24 * (size_t)e]; a[S + s] = eleshp[BOX_AY + 24 * (size_t)e]; a[2 * S + s] = eleshp[BOX_AZ + 24 * (size_t)e];
      s3[s] = ss[0 + 6 * (size_t)e]; s3[S + s] = ss[3 + 6 * (size_t)e]; s3[2 * S + s] = ss[5 + 6 * (size_t)e];
      lam[s] = mat[(size_t)e + 3 * (size_t)Ne]; mu[s] = mat[(size_t)e + 4 * (size_t)Ne]; det[s] = eledet[e];
      for (int k = 0; k < 3; ++k) dm[(size_t)k * S + s] = damps[k + 3 * (size_t)e];
      for (int k = 0; k < 21; ++k) sg[(size_t)k * S + s] = stress21[k + 21 * (size_t)e];
    }
    for (int e = 0; e < Ne; ++e) if (inBundle[e] > 1) return __LINE__;
    std::vector<double> v(3 * NnS, 0.0), d(3 * NnS, 0.0), pf(12 * PFS, 0.0);
    for (int nd = 0; nd < Nn; ++nd)
      for (int c = 0; c < 3; ++c) { v[c * NnS + nd] = vel[c + 3 * (size_t)nd]; d[c * NnS + nd] = disp[c + 3 * (size_t)nd]; }
    MarchPmlArgs A{};
    A.rec = P.rec.data(); A.ctaFirstA = P.ctaFirstA.data(); A.ctaFirstB = P.ctaFirstB.data(); A.code = P.code.data();
    A.S = S; A.NnS = NnS; A.PFS = PFS; A.slotBase = 0;
    A.a = a.data(); A.ss = s3.data(); A.lam = lam.data(); A.mu = mu.data(); A.det = det.data(); A.damps = dm.data(); A.stress = sg.data();
    A.vel = v.data(); A.disp = d.data(); A.pf = pf.data();
    A.dt = dt; A.rdampk = rdampk; A.w = w;
    std::vector<MarchPmlShared> smv(1);
    MarchPmlShared& sm = smv[0];
    std::vector<MarchPmlRegs> regs(MK_NT);
    for (int cta = 0; cta < P.grid; ++cta)
      for (int part = 0; part < 2; ++part)
      for (int b = (part ? P.ctaFirstB : P.ctaFirstA)[cta]; b < (part ? P.ctaFirstB : P.ctaFirstA)[cta + 1]; ++b) {
        const MarchBundle B = P.rec[b];
#define MK_RUN(body) do { for (int tid = 0; tid < MK_NT; ++tid) { MarchPmlRegs& R = regs[tid]; (void)R; body; } } while (0)
#define MK_WAITN ((void)0)
#define MK_WAITO(p) ((void)0)
        MARCH_PML_BUNDLE(MK_RUN, MK_RUN, MK_WAITN, MK_WAITO, A, B, sm, R);
#undef MK_RUN
#undef MK_WAITN
#undef MK_WAITO
      }
    for (size_t s = 0; s < S; ++s) {
      const int e = P.refId[s];
      if (e < 0) continue;
      for (int k = 0; k < 21; ++k) stress21[k + 21 * (size_t)e] = sg[(size_t)k * S + s];
    }
    for (size_t slot = 0; slot < PFS; ++slot) {
      const int code = P.code[slot];
      if (code < 0) continue;
      const int nd = code & MK_IDMASK;
      for (int r = 0; r < 12; ++r) f12[r + 12 * (size_t)nd] += pf[(size_t)r * PFS + slot];
    }
    return 0;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "eqd_march_pml_emulate: %s\n", e.what());
    return __LINE__;
  }
}
