// PROTOTYPE (round-2 preparation, CPU only, not part of the product): the marching scheme for box tiles
// described in DESIGN.md section 3d, emulated thread by thread, checked against the element-by-element
// forces of the reference's stored operators.
//
// A "bundle" is a column bundle of the structured grid: BZ x BY element columns, LX elements long in x
// (x = the slowest node index, node id = ix*ny*nz + iz*ny + iy, meshgen.f90:64-107).  One emulated thread
// per column and role (0 = stress part, calcElemKU.f90; 1 = hourglass part, hrglss.f90).  At step p a
// thread holds its element's four x- corners in "registers" (they were the x+ corners of step p-1), reads
// the four x+ corners from node plane p+1, evaluates the closed-form forces (eqd_box.h), completes the
// forces on the x- corners with the partial it carried, adds them to force plane p in four phases
// (one per (y,z) corner: no two columns touch the same node in a phase) and carries the x+ forces on.
// Output: one partial force per (bundle, node), summed here per node for the comparison.
//
// What this validates before any GPU time is spent: corner pairing (1->0, 2->3, 5->4, 6->7), plane-local
// node indexing, the 4-phase assembly, the carry across steps and the element numbering assumed for a
// bundle (x-major, z, y).  It does not model shared memory, barriers or timing.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <vector>

#include "../../eqdyna_b200/csrc/cuda/eqd_box.h"

using namespace eqd;

namespace {
constexpr int BZ = 4, BY = 16, LX = 32;
// (y,z) position of the x- corners 0,3,4,7 and their x+ partners 1,2,5,6
constexpr int XM[4] = {0, 3, 4, 7}, XP[4] = {1, 2, 5, 6};
constexpr int DY[4] = {0, 1, 0, 1}, DZ[4] = {0, 0, 1, 1};

struct Mesh {
  int Nn, Ne, nx, ny, nz;
  const double* coor; const int32_t* conn; const int32_t* etype; const int32_t* ndof;
  const double* shp; const double* phi; const double* ss; const double* det; const double* mat;  // mat(Ne,5)
};

inline int node_id(const Mesh& m, int ix, int iy, int iz) { return ix * m.ny * m.nz + iz * m.ny + iy; }   // 0-based
}  // namespace

// forces of all qualifying bundles: marching emulation -> fm, element-by-element with the stored operators -> fr
// (both (3,Nn), zero elsewhere).  v, d: nodal fields (3,Nn).  stats: bundles, elements in bundles, rejected
// bundles.  Returns 0, or a line number when an assumption about the numbering does not hold.
extern "C" int march_proto(int32_t Nn, int32_t Ne, int32_t nx, int32_t ny, int32_t nz, const double* coor, const int32_t* conn,
                           const int32_t* etype, const int32_t* ndof, const double* shp, const double* phi, const double* ss,
                           const double* det, const double* mat, const double* v, const double* d, double rdampk, double w,
                           double* fm, double* fr, int64_t* stats) {
  Mesh m{Nn, Ne, nx, ny, nz, coor, conn, etype, ndof, shp, phi, ss, det, mat};
  const int ncx = nx - 1, ncy = ny - 1, ncz = nz - 1;     // cells
  // element of cell (cx,cy,cz): meshes without wedges create one element per cell in node sweep order
  if ((long)ncx * ncy * ncz != Ne) return __LINE__;
  auto elem_of = [&](int cx, int cy, int cz) { return (cx * ncz + cz) * ncy + cy; };
  for (int k = 0; k < 64; ++k) {   // spot-check the numbering assumption on the connectivity
    int cx = (k * 7) % ncx, cy = (k * 13) % ncy, cz = (k * 5) % ncz;
    const int32_t* c = conn + 8 * (size_t)elem_of(cx, cy, cz);
    // corner 1 is the cell's (x-,y-,z-) node, unless the fault replaced it by a split-node master (id beyond the grid)
    if (c[0] - 1 != node_id(m, cx, cy, cz) && c[0] - 1 < nx * ny * nz) return __LINE__;
  }
  std::fill(fm, fm + 3 * (size_t)Nn, 0.0);
  std::fill(fr, fr + 3 * (size_t)Nn, 0.0);
  int64_t nb = 0, nel = 0, nrej = 0;
  const int grid = nx * ny * nz;
  for (int bx0 = 0; bx0 + LX <= ncx; bx0 += LX)
    for (int bz0 = 0; bz0 + BZ <= ncz; bz0 += BZ)
      for (int by0 = 0; by0 + BY <= ncy; by0 += BY) {
        // ---- qualify: all elements regular boxes on 3-dof regular-grid nodes with the structured connectivity
        bool ok = true;
        for (int p = 0; p < LX && ok; ++p)
          for (int cz = 0; cz < BZ && ok; ++cz)
            for (int cy = 0; cy < BY && ok; ++cy) {
              const int e = elem_of(bx0 + p, by0 + cy, bz0 + cz);
              const int32_t* c = conn + 8 * (size_t)e;
              if (etype[e] != 1) { ok = false; break; }
              int c0[8];
              for (int j = 0; j < 8; ++j) c0[j] = c[j] - 1;
              for (int j = 0; j < 8; ++j) {
                const int ix = bx0 + p + (box_px(j) ? 1 : 0), iy = by0 + cy + (box_py(j) ? 1 : 0), iz = bz0 + cz + (box_pz(j) ? 1 : 0);
                if (c0[j] >= grid || c0[j] != node_id(m, ix, iy, iz) || ndof[c0[j]] != 3) { ok = false; break; }
              }
              if (ok && !box_element(c0, coor)) ok = false;
            }
        if (!ok) { ++nrej; continue; }
        ++nb; nel += (int64_t)LX * BZ * BY;
        // ---- marching emulation
        const int PN = (BZ + 1) * (BY + 1);                       // nodes of a plane
        std::vector<double> fplane((size_t)2 * 2 * 3 * PN, 0.0);  // [plane parity][role][3][PN]
        struct Regs { double u[4][3]; double fc[4][3]; };         // x- corner values and carried forces per thread
        std::vector<Regs> regs((size_t)2 * BZ * BY);
        auto plane_node = [&](int p, int iy, int iz) { return node_id(m, bx0 + p, by0 + iy, bz0 + iz); };
        auto value = [&](int role, int n, int c) { return role == 0 ? v[c + 3 * (size_t)n] : d[c + 3 * (size_t)n] + rdampk * v[c + 3 * (size_t)n]; };
        for (int role = 0; role < 2; ++role)
          for (int cz = 0; cz < BZ; ++cz)
            for (int cy = 0; cy < BY; ++cy) {
              Regs& R = regs[(size_t)(role * BZ + cz) * BY + cy];
              for (int q = 0; q < 4; ++q)
                for (int c = 0; c < 3; ++c) { R.u[q][c] = value(role, plane_node(0, cy + DY[q], cz + DZ[q]), c); R.fc[q][c] = 0.0; }
            }
        auto flush = [&](int p) {
          double* F = &fplane[(size_t)(p & 1) * 2 * 3 * PN];
          for (int iz = 0; iz <= BZ; ++iz)
            for (int iy = 0; iy <= BY; ++iy) {
              const int n = plane_node(p, iy, iz), li = iz * (BY + 1) + iy;
              for (int c = 0; c < 3; ++c) {
                fm[c + 3 * (size_t)n] += F[(0 * 3 + c) * PN + li] + F[(1 * 3 + c) * PN + li];   // one partial per (bundle, node)
                F[(0 * 3 + c) * PN + li] = 0.0; F[(1 * 3 + c) * PN + li] = 0.0;
              }
            }
        };
        for (int p = 0; p < LX; ++p) {
          std::vector<double> fxm((size_t)2 * BZ * BY * 12);      // this step's completed x- corner forces per thread
          for (int role = 0; role < 2; ++role)
            for (int cz = 0; cz < BZ; ++cz)
              for (int cy = 0; cy < BY; ++cy) {
                Regs& R = regs[(size_t)(role * BZ + cz) * BY + cy];
                const int e = elem_of(bx0 + p, by0 + cy, bz0 + cz);
                double u[8][3], f[8][3];
                double up[4][3];
                for (int q = 0; q < 4; ++q)
                  for (int c = 0; c < 3; ++c) {
                    up[q][c] = value(role, plane_node(p + 1, cy + DY[q], cz + DZ[q]), c);   // "shared plane p+1"
                    u[XM[q]][c] = R.u[q][c];
                    u[XP[q]][c] = up[q][c];
                  }
                const double ax = shp[BOX_AX + 24 * (size_t)e], ay = shp[BOX_AY + 24 * (size_t)e], az = shp[BOX_AZ + 24 * (size_t)e];
                if (role == 0) {
                  // elastic stress part with zero stored stress: t = -det*w*(rdampk-weighted rate), enough to exercise B and B^T
                  double g[3][3], sr[6], t[6];
                  box_grad(u, g);
                  box_strain(g, ax, ay, az, sr);
                  const double lam = mat[(size_t)e + 3 * (size_t)Ne], mu = mat[(size_t)e + 4 * (size_t)Ne], l2m = lam + 2 * mu;
                  const double rate[6] = {l2m * sr[0] + lam * sr[1] + lam * sr[2], lam * sr[0] + l2m * sr[1] + lam * sr[2],
                                          lam * sr[0] + lam * sr[1] + l2m * sr[2], mu * sr[3], mu * sr[4], mu * sr[5]};
                  const double temp = (-det[e]) * w;
                  for (int k = 0; k < 6; ++k) t[k] = temp * rate[k];
                  box_force(t, ax, ay, az, f);
                } else {
                  box_hourglass(u, ss[0 + 6 * (size_t)e], ss[3 + 6 * (size_t)e], ss[5 + 6 * (size_t)e], f);
                }
                double* out = &fxm[((size_t)(role * BZ + cz) * BY + cy) * 12];
                for (int q = 0; q < 4; ++q)
                  for (int c = 0; c < 3; ++c) {
                    out[q * 3 + c] = R.fc[q][c] + f[XM[q]][c];   // carried partial + this element's x- corner
                    R.fc[q][c] = f[XP[q]][c];                    // carry the x+ corner forces
                    R.u[q][c] = up[q][c];                        // x+ values become next step's x- values
                  }
              }
          // four assembly phases into force plane p: in phase q every column adds to its (DY[q], DZ[q]) node
          double* F = &fplane[(size_t)(p & 1) * 2 * 3 * PN];
          for (int q = 0; q < 4; ++q)
            for (int role = 0; role < 2; ++role)
              for (int cz = 0; cz < BZ; ++cz)
                for (int cy = 0; cy < BY; ++cy) {
                  const double* in = &fxm[((size_t)(role * BZ + cz) * BY + cy) * 12];
                  const int li = (cz + DZ[q]) * (BY + 1) + (cy + DY[q]);
                  for (int c = 0; c < 3; ++c) F[(role * 3 + c) * PN + li] += in[q * 3 + c];
                }
          flush(p);
        }
        // last plane: the carried x+ forces of the final step
        {
          double* F = &fplane[(size_t)(LX & 1) * 2 * 3 * PN];
          for (int q = 0; q < 4; ++q)
            for (int role = 0; role < 2; ++role)
              for (int cz = 0; cz < BZ; ++cz)
                for (int cy = 0; cy < BY; ++cy) {
                  const Regs& R = regs[(size_t)(role * BZ + cz) * BY + cy];
                  const int li = (cz + DZ[q]) * (BY + 1) + (cy + DY[q]);
                  for (int c = 0; c < 3; ++c) F[(role * 3 + c) * PN + li] += R.fc[q][c];
                }
          flush(LX);
        }
        // ---- reference: the same elements, one by one, with the stored eleshp / phi / ss (general forms)
        for (int p = 0; p < LX; ++p)
          for (int cz = 0; cz < BZ; ++cz)
            for (int cy = 0; cy < BY; ++cy) {
              const int e = elem_of(bx0 + p, by0 + cy, bz0 + cz);
              const int32_t* c8 = conn + 8 * (size_t)e;
              const double* S = shp + 24 * (size_t)e;
              const double* PH = phi + 32 * (size_t)e;
              const double* S6 = ss + 6 * (size_t)e;
              double sr[6] = {0, 0, 0, 0, 0, 0};
              for (int i = 0; i < 8; ++i) {
                const int n = c8[i] - 1;
                const double s1 = S[3 * i], s2 = S[3 * i + 1], s3 = S[3 * i + 2];
                const double vx = v[3 * (size_t)n], vy = v[1 + 3 * (size_t)n], vz = v[2 + 3 * (size_t)n];
                sr[0] += s1 * vx; sr[1] += s2 * vy; sr[2] += s3 * vz;
                sr[3] += s3 * vy + s2 * vz; sr[4] += s3 * vx + s1 * vz; sr[5] += s2 * vx + s1 * vy;
              }
              const double lam = mat[(size_t)e + 3 * (size_t)Ne], mu = mat[(size_t)e + 4 * (size_t)Ne], l2m = lam + 2 * mu;
              const double rate[6] = {l2m * sr[0] + lam * sr[1] + lam * sr[2], lam * sr[0] + l2m * sr[1] + lam * sr[2],
                                      lam * sr[0] + lam * sr[1] + l2m * sr[2], mu * sr[3], mu * sr[4], mu * sr[5]};
              const double temp = (-det[e]) * w;
              double t[6];
              for (int k = 0; k < 6; ++k) t[k] = temp * rate[k];
              double phid[4][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, hv[4][3];
              for (int i = 0; i < 8; ++i) {
                const int n = c8[i] - 1;
                for (int mm = 0; mm < 4; ++mm)
                  for (int c = 0; c < 3; ++c) phid[mm][c] += PH[8 * mm + i] * (d[c + 3 * (size_t)n] + rdampk * v[c + 3 * (size_t)n]);
              }
              for (int mm = 0; mm < 4; ++mm) {
                hv[mm][0] = S6[0] * phid[mm][0] + S6[1] * phid[mm][1] + S6[2] * phid[mm][2];
                hv[mm][1] = S6[1] * phid[mm][0] + S6[3] * phid[mm][1] + S6[4] * phid[mm][2];
                hv[mm][2] = S6[2] * phid[mm][0] + S6[4] * phid[mm][1] + S6[5] * phid[mm][2];
              }
              for (int i = 0; i < 8; ++i) {
                const int n = c8[i] - 1;
                const double s1 = S[3 * i], s2 = S[3 * i + 1], s3 = S[3 * i + 2];
                double f0 = s1 * t[0] + s3 * t[4] + s2 * t[5], f1 = s2 * t[1] + s3 * t[3] + s1 * t[5], f2 = s3 * t[2] + s2 * t[3] + s1 * t[4];
                for (int mm = 0; mm < 4; ++mm) { f0 -= PH[8 * mm + i] * hv[mm][0]; f1 -= PH[8 * mm + i] * hv[mm][1]; f2 -= PH[8 * mm + i] * hv[mm][2]; }
                fr[3 * (size_t)n] += f0; fr[1 + 3 * (size_t)n] += f1; fr[2 + 3 * (size_t)n] += f2;
              }
            }
      }
  stats[0] = nb; stats[1] = nel; stats[2] = nrej;
  return 0;
}
