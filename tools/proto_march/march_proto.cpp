// PROTOTYPE driver (CPU): builds bundles on a structured mesh, runs the SINGLE-SOURCE marching kernel of
// march_kernel.cuh phase by phase over all thread ids (the host reading of its MARCH_BUNDLE schedule) and
// compares partial forces and updated stresses with an element-by-element evaluation from the reference's
// stored eleshp / phi / ss (calcElemKU.f90:44-189, hrglss.f90:20-54).  Not part of the product.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "march_kernel.cuh"

using namespace eqd;
using namespace march;

// fm / fr: (3,Nn) nodal forces of the bundle elements by the marching kernel / element by element.
// dev[0] = largest |stress_march - stress_ref| / largest |stress_ref| after one step with dt.
// stats: bundles, elements in bundles, rejected lattice positions, stress deviation (1e-18 units), regular box
// elements of the mesh (the most bundles could cover).  Returns 0 or a line number.
extern "C" int march_proto(int32_t Nn, int32_t Ne, int32_t nx, int32_t ny, int32_t nz, const double* coor, const int32_t* conn,
                           const int32_t* etype, const int32_t* ndof, const double* shp, const double* phi, const double* ss,
                           const double* det, const double* mat, const double* v, const double* d, double rdampk, double w,
                           double* fm, double* fr, int64_t* stats) {
  const int LX = 32;
  const double dt = 0.004;
  const int ncx = nx - 1, ncy = ny - 1, ncz = nz - 1;
  if ((long)ncx * ncy * ncz != Ne) return __LINE__;          // one element per cell: meshes without wedges
  const int grid = nx * ny * nz;
  auto node_id = [&](int ix, int iy, int iz) { return ix * ny * nz + iz * ny + iy; };
  auto elem_of = [&](int cx, int cy, int cz) { return (cx * ncz + cz) * ncy + cy; };
  for (int k = 0; k < 64; ++k) {
    int cx = (k * 7) % ncx, cy = (k * 13) % ncy, cz = (k * 5) % ncz;
    const int32_t* c = conn + 8 * (size_t)elem_of(cx, cy, cz);
    // corner 1 is the cell's (x-,y-,z-) node, unless the fault replaced it by a split-node master (id beyond the grid)
    if (c[0] - 1 != node_id(cx, cy, cz) && c[0] - 1 < grid) return __LINE__;
  }
  // ---- bundles: the lattice starts at the first regular cell of every axis (behind the PML), restarts in y at
  // the fault plane (cells on its + side reference split-node masters), and its last bundle in x is shortened
  // to what is left (>= 8 elements).  A bundle qualifies when all its elements are boxes of type 1 on 3-dof
  // nodes and agree on their shared corners.  Node ids are taken from the connectivity, so masters are welcome.
  auto regular_cell = [&](int cx, int cy, int cz) {
    const int e = elem_of(cx, cy, cz);
    if (etype[e] != 1) return false;
    for (int j = 0; j < 8; ++j) if (ndof[conn[8 * (size_t)e + j] - 1] != 3) return false;
    return true;
  };
  auto span = [&](int n, auto&& ok, int& lo, int& hi) {
    lo = 0; hi = n - 1;
    while (lo < n && !ok(lo)) ++lo;
    while (hi > lo && !ok(hi)) --hi;
  };
  int xlo, xhi, ylo, yhi, zlo, zhi;
  span(ncx, [&](int c) { return regular_cell(c, ncy / 4, ncz / 2); }, xlo, xhi);
  span(ncy, [&](int c) { return regular_cell(ncx / 2, c, ncz / 2); }, ylo, yhi);
  span(ncz, [&](int c) { return regular_cell(ncx / 2, ncy / 4, c); }, zlo, zhi);
  int yfault = -1;   // first cell row on the + side of the fault: one of its nodes is a master (id beyond the grid)
  for (int cy = ylo; cy <= yhi && yfault < 0; ++cy) {
    const int e = elem_of(ncx / 2, cy, ncz / 2);
    for (int j = 0; j < 8; ++j) if (conn[8 * (size_t)e + j] - 1 >= grid) yfault = cy;
  }
  std::vector<int> ystarts;
  for (int seg = 0; seg < 2; ++seg) {
    const int a = seg == 0 ? ylo : (yfault > ylo ? yfault : yhi + 1), b = seg == 0 ? (yfault > ylo ? yfault - 1 : yhi) : yhi;
    for (int by0 = a; by0 + BY <= b + 1; by0 += BY) ystarts.push_back(by0);
  }
  std::vector<Bundle> rec;
  std::vector<int> tnode, elemOf;      // plane-ordered node ids; reference element of every bundle element slot
  int64_t nrej = 0;
  for (int bx0 = xlo; bx0 <= xhi;) {
    const int LXb = std::min(LX, xhi - bx0 + 1);
    if (LXb < 8) break;
    for (int bz0 = zlo; bz0 + BZ <= zhi + 1; bz0 += BZ)
      for (int by0 : ystarts) {
        bool ok = true;
        std::vector<int> nodes((size_t)(LXb + 1) * PN, -1), elems;
        for (int p = 0; p < LXb && ok; ++p)
          for (int cz = 0; cz < BZ && ok; ++cz)
            for (int cy = 0; cy < BY && ok; ++cy) {
              const int e = elem_of(bx0 + p, by0 + cy, bz0 + cz);
              int c0[8];
              for (int j = 0; j < 8; ++j) c0[j] = conn[8 * (size_t)e + j] - 1;
              if (etype[e] != 1 || !box_element(c0, coor)) { ok = false; break; }
              for (int j = 0; j < 8; ++j) {
                if (ndof[c0[j]] != 3) { ok = false; break; }
                const int pp = p + (box_px(j) ? 1 : 0), li = (cz + (box_pz(j) ? 1 : 0)) * (BY + 1) + cy + (box_py(j) ? 1 : 0);
                int& slot = nodes[(size_t)pp * PN + li];
                if (slot >= 0 && slot != c0[j]) { ok = false; break; }   // two elements disagree on a shared corner (fault inside the bundle)
                slot = c0[j];
              }
              elems.push_back(e);
            }
        if (!ok) { ++nrej; continue; }
        rec.push_back(Bundle{(int)elemOf.size(), (int)tnode.size(), LXb, 0});
        elemOf.insert(elemOf.end(), elems.begin(), elems.end());
        tnode.insert(tnode.end(), nodes.begin(), nodes.end());
      }
    bx0 += LXb;
  }
  const size_t S = std::max<size_t>(elemOf.size(), 1), NnS = Nn, PFS = std::max<size_t>(tnode.size(), 1);
  // ---- class SoA of the bundle elements (what eqd_compute_elem_ops / eqd_set_elem_ops would fill)
  std::vector<double> ax(S), ay(S), az(S), s0(S), s3(S), s5(S), lam(S), mu(S), dt_(S), stress(6 * S), stressRef;
  for (size_t s = 0; s < elemOf.size(); ++s) {
    const size_t e = elemOf[s];
    ax[s] = shp[BOX_AX + 24 * e]; ay[s] = shp[BOX_AY + 24 * e]; az[s] = shp[BOX_AZ + 24 * e];
    s0[s] = ss[0 + 6 * e]; s3[s] = ss[3 + 6 * e]; s5[s] = ss[5 + 6 * e];
    lam[s] = mat[e + 3 * (size_t)Ne]; mu[s] = mat[e + 4 * (size_t)Ne]; dt_[s] = det[e];
    for (int k = 0; k < 6; ++k) stress[k * S + s] = 1.0e3 * std::sin(0.001 * (double)e + k);   // some pre-stress
  }
  stressRef = stress;
  std::vector<double> vel(3 * NnS), disp(3 * NnS), pf(3 * PFS, 0.0);
  for (int n = 0; n < Nn; ++n)
    for (int c = 0; c < 3; ++c) { vel[c * NnS + n] = v[c + 3 * (size_t)n]; disp[c * NnS + n] = d[c + 3 * (size_t)n]; }
  Args A{};
  A.nBundles = (int)rec.size(); A.rec = rec.data(); A.tnode = tnode.data(); A.S = S; A.NnS = NnS; A.PFS = PFS;
  A.ax = ax.data(); A.ay = ay.data(); A.az = az.data(); A.ss0 = s0.data(); A.ss3 = s3.data(); A.ss5 = s5.data();
  A.lam = lam.data(); A.mu = mu.data(); A.det = dt_.data(); A.stress = stress.data();
  A.vel = vel.data(); A.disp = disp.data(); A.pf = pf.data(); A.dt = dt; A.rdampk = rdampk; A.w = w;
  // ---- the kernel's schedule, read on the host: every phase over all thread ids, then the next phase
  {
    std::vector<Shared> smv(1);
    Shared& sm = smv[0];
    std::vector<Regs> R(NT);
    for (int b = 0; b < A.nBundles; ++b) {
      const Bundle B = A.rec[b];
#define MK_RUN(body) do { for (int tid = 0; tid < NT; ++tid) { body; } } while (0)
      const char* variant = std::getenv("MARCH_PREFETCH");        // which schedule of march_kernel.cuh
      if (variant && variant[0] == '2') MARCH_BUNDLE_PF2(MK_RUN, A, B, sm, R[tid]);
      else if (variant && variant[0] == '1') MARCH_BUNDLE_PF(MK_RUN, A, B, sm, R[tid]);
      else MARCH_BUNDLE(MK_RUN, A, B, sm, R[tid]);
#undef MK_RUN
    }
  }
  std::fill(fm, fm + 3 * (size_t)Nn, 0.0);
  std::fill(fr, fr + 3 * (size_t)Nn, 0.0);
  for (size_t k = 0; k < tnode.size(); ++k)
    if (tnode[k] >= 0)
      for (int c = 0; c < 3; ++c) fm[c + 3 * (size_t)tnode[k]] += pf[c * PFS + k];   // what the node update sums
  // ---- reference: the same elements one by one with the stored operators
  for (size_t s = 0; s < elemOf.size(); ++s) {
    const size_t e = elemOf[s];
    const int32_t* c8 = conn + 8 * e;
    const double* SH = shp + 24 * e;
    const double* PH = phi + 32 * e;
    const double* S6 = ss + 6 * e;
    double sr[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 8; ++i) {
      const size_t n = c8[i] - 1;
      const double s1 = SH[3 * i], s2 = SH[3 * i + 1], s3_ = SH[3 * i + 2];
      const double vx = v[3 * n], vy = v[1 + 3 * n], vz = v[2 + 3 * n];
      sr[0] += s1 * vx; sr[1] += s2 * vy; sr[2] += s3_ * vz;
      sr[3] += s3_ * vy + s2 * vz; sr[4] += s3_ * vx + s1 * vz; sr[5] += s2 * vx + s1 * vy;
    }
    const double la = mat[e + 3 * (size_t)Ne], m_ = mat[e + 4 * (size_t)Ne], l2m = la + 2 * m_;
    const double rate[6] = {l2m * sr[0] + la * sr[1] + la * sr[2], la * sr[0] + l2m * sr[1] + la * sr[2],
                            la * sr[0] + la * sr[1] + l2m * sr[2], m_ * sr[3], m_ * sr[4], m_ * sr[5]};
    const double temp = (-det[e]) * w;
    double t[6];
    for (int k = 0; k < 6; ++k) {
      const double sg = stressRef[k * S + s] + rate[k] * dt;
      stressRef[k * S + s] = sg;
      t[k] = temp * (sg + rdampk * rate[k]);
    }
    double phid[4][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, hv[4][3];
    for (int i = 0; i < 8; ++i) {
      const size_t n = c8[i] - 1;
      for (int mm = 0; mm < 4; ++mm)
        for (int c = 0; c < 3; ++c) phid[mm][c] += PH[8 * mm + i] * (d[c + 3 * n] + rdampk * v[c + 3 * n]);
    }
    for (int mm = 0; mm < 4; ++mm) {
      hv[mm][0] = S6[0] * phid[mm][0] + S6[1] * phid[mm][1] + S6[2] * phid[mm][2];
      hv[mm][1] = S6[1] * phid[mm][0] + S6[3] * phid[mm][1] + S6[4] * phid[mm][2];
      hv[mm][2] = S6[2] * phid[mm][0] + S6[4] * phid[mm][1] + S6[5] * phid[mm][2];
    }
    for (int i = 0; i < 8; ++i) {
      const size_t n = c8[i] - 1;
      const double s1 = SH[3 * i], s2 = SH[3 * i + 1], s3_ = SH[3 * i + 2];
      double f0 = s1 * t[0] + s3_ * t[4] + s2 * t[5], f1 = s2 * t[1] + s3_ * t[3] + s1 * t[5], f2 = s3_ * t[2] + s2 * t[3] + s1 * t[4];
      for (int mm = 0; mm < 4; ++mm) { f0 -= PH[8 * mm + i] * hv[mm][0]; f1 -= PH[8 * mm + i] * hv[mm][1]; f2 -= PH[8 * mm + i] * hv[mm][2]; }
      fr[3 * n] += f0; fr[1 + 3 * n] += f1; fr[2 + 3 * n] += f2;
    }
  }
  double smax = 0, sdev = 0;
  for (size_t k = 0; k < 6 * S && !elemOf.empty(); ++k) { smax = std::max(smax, std::fabs(stressRef[k])); sdev = std::max(sdev, std::fabs(stress[k] - stressRef[k])); }
  int64_t nRegBox = 0;   // what the bundles could cover at best: type-1 boxes on 3-dof nodes
  for (int e = 0; e < Ne; ++e) {
    if (etype[e] != 1) continue;
    int c0[8];
    bool three = true;
    for (int j = 0; j < 8; ++j) { c0[j] = conn[8 * (size_t)e + j] - 1; three = three && ndof[c0[j]] == 3; }
    if (three && box_element(c0, coor)) ++nRegBox;
  }
  stats[0] = (int64_t)rec.size(); stats[1] = (int64_t)elemOf.size(); stats[2] = nrej; stats[4] = nRegBox;
  stats[3] = smax > 0 ? (int64_t)std::llround(1e18 * std::min(sdev / smax, 1.0)) : 0;   // stress deviation in units of 1e-18
  return 0;
}
