// One-time operator precompute of the stand-in host: lumped mass, Jacobian
// determinant, one-point shape-function derivatives and the Kosloff-Frazier
// hourglass operators.  Restates src/assembleGlobalMass.f90:3-56,283-406,
// src/calcGlobalShapeFunc.f90, src/calcLocalShapeFunc.f90 and
// src/library.f90:60-93 (vlm); plus the init-time shared-node sums
// (MPI4NodalQuant for nodalMassArr and fnms, assembleGlobalMass.f90:40-41)
// for in-process multi-sub-domain worlds, and init_vel (eqdyna3d.f90:191-212).
#include <cmath>
#include <stdexcept>

#include "eqh_state.h"

namespace eqh {

namespace {

// calcLocalShapeFunc.f90:19-25: N_i,xi = acoor/8, N_i = 1/8
static const double ACOOR[8][3] = {{-1, -1, -1}, {1, -1, -1}, {1, 1, -1}, {-1, 1, -1},
                                   {-1, -1, 1},  {1, -1, 1},  {1, 1, 1},  {-1, 1, 1}};

// vlm, library.f90:60-93 (Belytschko et al. 1984)
double vlm(const double xl[8][3]) {
  static const int it[8][8] = {{1, 2, 3, 4, 5, 6, 7, 8}, {2, 3, 4, 1, 6, 7, 8, 5}, {3, 4, 1, 2, 7, 8, 5, 6},
                               {4, 1, 2, 3, 8, 5, 6, 7}, {5, 8, 7, 6, 1, 4, 3, 2}, {6, 5, 8, 7, 2, 1, 4, 3},
                               {7, 6, 5, 8, 3, 2, 1, 4}, {8, 7, 6, 5, 4, 3, 2, 1}};
  // Fortran it(a,i) with reshape column-major: it(a,i) = it_c[i-1][a-1]
  auto IT = [&](int a, int i) { return it[i - 1][a - 1] - 1; };
  auto Y = [&](int n) { return xl[n][1]; };
  auto Z = [&](int n) { return xl[n][2]; };
  double bb[8];
  for (int i = 1; i <= 8; ++i) {
    bb[i - 1] = Y(IT(2, i)) * (Z(IT(6, i)) - Z(IT(3, i)) + Z(IT(5, i)) - Z(IT(4, i))) +
                Y(IT(3, i)) * (Z(IT(2, i)) - Z(IT(4, i))) +
                Y(IT(4, i)) * (Z(IT(3, i)) - Z(IT(8, i)) + Z(IT(2, i)) - Z(IT(5, i))) +
                Y(IT(5, i)) * (Z(IT(8, i)) - Z(IT(6, i)) + Z(IT(4, i)) - Z(IT(2, i))) +
                Y(IT(6, i)) * (Z(IT(5, i)) - Z(IT(2, i))) + Y(IT(8, i)) * (Z(IT(4, i)) - Z(IT(5, i)));
  }
  double volume = 0.0;
  for (int i = 0; i < 8; ++i) volume = volume + xl[i][0] * bb[i];
  return volume / 12.0;
}

}  // namespace

void assemble_global_mass(const CaseInput& in, RankState& s) {
  const int Ne = s.totalNumOfElements;
  s.eledet.assign(Ne, 0.0);
  s.elemass.assign((size_t)24 * Ne, 0.0);
  s.eleshp.assign((size_t)24 * Ne, 0.0);
  s.ss.assign((size_t)6 * Ne, 0.0);
  s.phi.assign((size_t)32 * Ne, 0.0);
  const double cst = 1.0 / 8.0;
  double lshg[8][4];  // localShapeFunc(j,i): [i][j-1]
  for (int i = 0; i < 8; ++i) {
    lshg[i][3] = cst;
    for (int j = 0; j < 3; ++j) lshg[i][j] = cst * ACOOR[i][j];
  }
  static const int ha[4][8] = {{1, 1, -1, -1, -1, -1, 1, 1},
                               {1, -1, -1, 1, -1, 1, 1, -1},
                               {1, -1, 1, -1, 1, -1, 1, -1},
                               {-1, 1, -1, 1, 1, -1, 1, -1}};
  for (int nel = 1; nel <= Ne; ++nel) {
    const int* conn = &s.nodeElemIdRelation[8 * (size_t)(nel - 1)];
    double xl[8][3];
    for (int i = 0; i < 8; ++i)
      for (int j = 0; j < 3; ++j) xl[i][j] = s.meshCoor[j + 3 * (size_t)(conn[i] - 1)];
    // calcGlobalShapeFunc.f90:19-75
    double shg[8][4];
    for (int i = 0; i < 8; ++i)
      for (int j = 0; j < 4; ++j) shg[i][j] = lshg[i][j];
    int et = s.elemTypeArr[nel - 1];
    if (et == 11 || et == 12) {
      for (int j = 0; j < 4; ++j) {
        shg[2][j] = lshg[2][j] + lshg[3][j];
        shg[3][j] = 0.0;
        shg[6][j] = lshg[6][j] + lshg[7][j];
        shg[7][j] = 0.0;
      }
    }
    double xs[3][3];  // xs(j,i) -> xs[j-1][i-1]
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        double temp = 0.0;
        for (int k = 0; k < 8; ++k) temp = temp + shg[k][i] * xl[k][j];
        xs[j][i] = temp;
      }
    double cof11 = xs[1][1] * xs[2][2] - xs[1][2] * xs[2][1];
    double cof12 = xs[1][2] * xs[2][0] - xs[1][0] * xs[2][2];
    double cof13 = xs[1][0] * xs[2][1] - xs[1][1] * xs[2][0];
    double cof21 = xs[2][1] * xs[0][2] - xs[2][2] * xs[0][1];
    double cof22 = xs[2][2] * xs[0][0] - xs[2][0] * xs[0][2];
    double cof23 = xs[2][0] * xs[0][1] - xs[2][1] * xs[0][0];
    double cof31 = xs[0][1] * xs[1][2] - xs[0][2] * xs[1][1];
    double cof32 = xs[0][2] * xs[1][0] - xs[0][0] * xs[1][2];
    double cof33 = xs[0][0] * xs[1][1] - xs[0][1] * xs[1][0];
    double det = xs[0][0] * cof11 + xs[0][1] * cof12 + xs[0][2] * cof13;
    if (det <= 0.0) throw std::runtime_error("Non-positive determinant; element " + std::to_string(nel));
    double tmpS[8][4];
    for (int i = 0; i < 8; ++i)
      for (int j = 0; j < 4; ++j) tmpS[i][j] = shg[i][j];
    for (int i = 0; i < 8; ++i) {
      shg[i][0] = (tmpS[i][0] * cof11 + tmpS[i][1] * cof12 + tmpS[i][2] * cof13) / det;
      shg[i][1] = (tmpS[i][0] * cof21 + tmpS[i][1] * cof22 + tmpS[i][2] * cof23) / det;
      shg[i][2] = (tmpS[i][0] * cof31 + tmpS[i][1] * cof32 + tmpS[i][2] * cof33) / det;
    }
    // xs = reshape((/cof11,cof12,cof13, cof21,.../),(/3,3/))/det  => xs(1,1)=cof11, xs(2,1)=cof12, ...
    double xsn[3][3];  // xsn[a-1][b-1] = xs(a,b)
    xsn[0][0] = cof11 / det; xsn[1][0] = cof12 / det; xsn[2][0] = cof13 / det;
    xsn[0][1] = cof21 / det; xsn[1][1] = cof22 / det; xsn[2][1] = cof23 / det;
    xsn[0][2] = cof31 / det; xsn[1][2] = cof32 / det; xsn[2][2] = cof33 / det;
    // contm, assembleGlobalMass.f90:376-406
    double eleffm[24];
    {
      const double constm = s.mat[(size_t)(nel - 1) + (size_t)Ne * 2];
      double dsum = 0.0, totmas = constm * in.w * det, work[8];
      for (int j = 0; j < 8; ++j) {
        double temp2 = totmas * shg[j][3] * shg[j][3];
        dsum = dsum + temp2;
        work[j] = 0.0 + temp2;
      }
      double temp1 = totmas / dsum;
      for (int j = 0; j < 8; ++j) {
        double temp2 = temp1 * work[j];
        for (int k = 0; k < 3; ++k) eleffm[3 * j + k] = temp2;
      }
    }
    // assembleElementMassDetShg, assembleGlobalMass.f90:283-326
    for (int i = 0; i < 8; ++i) {
      int nodeID = conn[i];
      int st = s.eqNumStartIndexLoc[nodeID - 1];
      if (s.numOfDofPerNodeArr[nodeID - 1] == 12) {
        for (int ixyz = 1; ixyz <= 3; ++ixyz) {
          for (int j = 3 * (ixyz - 1) + 1; j <= 3 * (ixyz - 1) + 3; ++j) {
            int eq = s.eqNumIndexArr[st + j - 1];
            if (eq > 0) s.nodalMassArr[eq - 1] += eleffm[3 * i + ixyz - 1];
          }
          int eq = s.eqNumIndexArr[st + ixyz + 9 - 1];
          if (eq > 0) s.nodalMassArr[eq - 1] += eleffm[3 * i + ixyz - 1];
        }
      } else if (s.numOfDofPerNodeArr[nodeID - 1] == 3) {
        for (int j = 1; j <= 3; ++j) {
          int eq = s.eqNumIndexArr[st + j - 1];
          s.nodalMassArr[eq - 1] += eleffm[3 * i + j - 1];
        }
      }
      s.fnms[nodeID - 1] += eleffm[3 * i];
    }
    for (int i = 0; i < 24; ++i) s.elemass[i + 24 * (size_t)(nel - 1)] = eleffm[i];
    s.eledet[nel - 1] = det;
    for (int i = 0; i < 8; ++i)
      for (int j = 0; j < 3; ++j) s.eleshp[j + 3 * (i + 8 * (size_t)(nel - 1))] = shg[i][j];
    // calcSSPhi4Hrgls, assembleGlobalMass.f90:328-374
    {
      double vol = vlm(xl);
      double lam = s.mat[(size_t)(nel - 1) + (size_t)Ne * 3], miu = s.mat[(size_t)(nel - 1) + (size_t)Ne * 4];
      double ce = miu * (3 * lam + 2 * miu) / (lam + miu);
      ce = 16.0 * ce / 15.0;
      double co = ce * vol / 48.0;
      double* SS = &s.ss[6 * (size_t)(nel - 1)];
      auto XS = [&](int a, int b) { return xsn[a - 1][b - 1]; };
      SS[0] = co * (XS(1, 1) * XS(1, 1) + XS(2, 1) * XS(2, 1) + XS(3, 1) * XS(3, 1));
      SS[1] = co * (XS(1, 1) * XS(1, 2) + XS(2, 1) * XS(2, 2) + XS(3, 1) * XS(3, 2));
      SS[2] = co * (XS(1, 1) * XS(1, 3) + XS(2, 1) * XS(2, 3) + XS(3, 1) * XS(3, 3));
      SS[3] = co * (XS(1, 2) * XS(1, 2) + XS(2, 2) * XS(2, 2) + XS(3, 2) * XS(3, 2));
      SS[4] = co * (XS(1, 2) * XS(1, 3) + XS(2, 2) * XS(2, 3) + XS(3, 2) * XS(3, 3));
      SS[5] = co * (XS(1, 3) * XS(1, 3) + XS(2, 3) * XS(2, 3) + XS(3, 3) * XS(3, 3));
      double* PHI = &s.phi[32 * (size_t)(nel - 1)];  // phi(j,i,nel) -> PHI[(j-1)+8*(i-1)]
      for (int i = 0; i < 4; ++i) {
        for (int j = 0; j < 8; ++j) {
          double v = 0.0;
          for (int k = 0; k < 8; ++k)
            v = v + ha[i][k] * (xl[k][0] * shg[j][0] + xl[k][1] * shg[j][1] + xl[k][2] * shg[j][2]);
          PHI[j + 8 * i] = ha[i][j] - v;
        }
        double v = 0.0;
        for (int j = 0; j < 8; ++j) v = v + PHI[j + 8 * i] * PHI[j + 8 * i];
        v = std::sqrt(v / 8.0);
        for (int j = 0; j < 8; ++j) PHI[j + 8 * i] = PHI[j + 8 * i] / v;
      }
    }
  }
  // assembleGlobalMass.f90:43-55 sanity check of wedges
  for (int nel = 1; nel <= Ne; ++nel) {
    int et = s.elemTypeArr[nel - 1];
    if (et >= 11 && et <= 12) {
      const int* conn = &s.nodeElemIdRelation[8 * (size_t)(nel - 1)];
      if (conn[2] != conn[3]) throw std::runtime_error("Wrongly created wedge; nel " + std::to_string(nel));
    }
  }
}

// MPI4NodalQuant (assembleGlobalMass.f90:58-281) emulated across an in-process
// world, for the two init-time calls.  which: 0 = nodalMassArr (numDof 3),
// 1 = fnms (numDof 1).  Per axis, the "-" then "+" faces; both faces of one rank
// are disjoint node sets, so snapshot-then-add equals the blocking sequence.
void exchange_nodal(const CaseInput& in, std::vector<RankState*>& world, int which, int numDof) {
  const int stride[3] = {in.npy * in.npz, in.npz, 1};
  const int npxyz[3] = {in.npx, in.npy, in.npz};
  auto face_slots = [&](RankState& s, int a, int side, std::vector<double*>& out) {
    out.clear();
    const int nx = s.nx, ny = s.ny, nz = s.nz;
    const int n[3] = {nx, ny, nz};
    const int b = side == 0 ? 1 : n[a];
    double* q = which == 0 ? s.nodalMassArr.data() : s.fnms.data();
    auto push_node = [&](int node) {
      if (numDof == 1) { out.push_back(&q[node - 1]); return; }
      int st = s.eqNumStartIndexLoc[node - 1];
      for (int d = 1; d <= s.numOfDofPerNodeArr[node - 1]; ++d) {
        int eq = s.eqNumIndexArr[st + d - 1];
        if (eq > 0) out.push_back(&q[eq - 1]);
      }
    };
    if (a == 0) {
      for (int iz = 1; iz <= nz; ++iz)
        for (int iy = 1; iy <= ny; ++iy) push_node((b - 1) * ny * nz + (iz - 1) * ny + iy);
    } else if (a == 1) {
      for (int ix = 1; ix <= nx; ++ix)
        for (int iz = 1; iz <= nz; ++iz) push_node((ix - 1) * ny * nz + (iz - 1) * ny + b);
    } else {
      for (int ix = 1; ix <= nx; ++ix)
        for (int iy = 1; iy <= ny; ++iy) push_node((ix - 1) * ny * nz + (b - 1) * ny + iy);
    }
    if (s.fltMPI[2 * a + side])
      for (int i : s.fltface[2 * a + side]) push_node(nx * ny * nz + i);
  };
  for (int a = 0; a < 3; ++a) {
    if (npxyz[a] <= 1) continue;
    std::vector<std::vector<double>> snap[2];
    snap[0].resize(world.size());
    snap[1].resize(world.size());
    std::vector<double*> slots;
    for (size_t r = 0; r < world.size(); ++r)
      for (int side = 0; side < 2; ++side) {
        face_slots(*world[r], a, side, slots);
        for (double* p : slots) snap[side][r].push_back(*p);
      }
    for (size_t r = 0; r < world.size(); ++r) {
      RankState& s = *world[r];
      int mexyz[3] = {s.mex, s.mey, s.mez};
      for (int side = 0; side < 2; ++side) {
        bool active = side == 0 ? (mexyz[a] != 0) : (mexyz[a] != npxyz[a] - 1);
        if (!active) continue;
        int nb = s.me + (side == 0 ? -stride[a] : stride[a]);
        face_slots(s, a, side, slots);
        const std::vector<double>& rv = snap[1 - side][nb];
        if (rv.size() != slots.size()) throw std::runtime_error("exchange_nodal: face size mismatch");
        for (size_t k = 0; k < slots.size(); ++k) *slots[k] = *slots[k] + rv[k];
      }
    }
  }
}

// init_vel, eqdyna3d.f90:191-212
void init_vel(const CaseInput& in, RankState& s) {
  for (int ift = 0; ift < in.ntotft; ++ift)
    for (int i = 1; i <= s.nftnd[ift]; ++i) {
      size_t pb = (size_t)(i - 1) + (size_t)s.nftmx * ift;
      const double* f = &s.fric[100 * pb];
      int st = s.eqNumStartIndexLoc[s.nsmp[0 + 2 * pb] - 1];
      for (int k = 0; k < 3; ++k) s.v1[s.eqNumIndexArr[st + k] - 1] = f[33 + k];
      st = s.eqNumStartIndexLoc[s.nsmp[1 + 2 * pb] - 1];
      for (int k = 0; k < 3; ++k) s.v1[s.eqNumIndexArr[st + k] - 1] = f[30 + k];
    }
}

}  // namespace eqh
