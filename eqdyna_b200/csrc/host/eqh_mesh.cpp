// Mesh sizing and generation for one sub-domain: the stand-in host's
// restatement of src/mesh4num.f90, src/meshgen.f90 and
// src/library_degeneration.f90 (there is no Fortran compiler in the build
// container).  Node, element and equation numbering must be
// bit-identical to the reference, so the sweep order (ix outer, iz, iy inner),
// the order of counters and every comparison are kept as in the Fortran; the
// data structures are ordinary C++ (direct index arithmetic instead of the
// plane1/plane2 sliding planes, a hash map instead of the O(pairs) search of
// replaceSlaveWithMasterNode).
#include <cmath>
#include <cstdio>
#include <stdexcept>
#include <unordered_map>

#include "eqh_state.h"

namespace eqh {

namespace {

inline int nint(double x) { return (int)std::lround(x); }

struct Line {
  int nGlobal = 0, nUni = 0, frontEdge = 0, nLocal = 0;
  std::vector<double> local;  // 1-based use: local[i-1]
};

// getLocalOneDimCoorArrAndSize, meshgen.f90:447-560
Line local_line(const CaseInput& in, int dimId, int mpiId, double bound[2], double* PMLb) {
  const double* flt = in.fltxyz.data();  // fault 1
  const int np = 1000000;                // globalvar.f90:47
  Line L;
  double gridSize, frontEdgeCoor, backEdgeCoor, minCoor, maxCoor;
  int numMPI;
  if (dimId == 1) {
    L.nUni = nint((flt[1] - flt[0]) / in.dx) + 1;
    gridSize = in.dx; frontEdgeCoor = flt[0]; backEdgeCoor = flt[1];
    minCoor = in.xmin; maxCoor = in.xmax; numMPI = in.npx;
  } else if (dimId == 2) {
    L.nUni = in.dis4uniF + in.dis4uniB + 1;
    gridSize = in.dy; frontEdgeCoor = -in.dis4uniF * in.dy; backEdgeCoor = in.dis4uniB * in.dy;
    minCoor = in.ymin; maxCoor = in.ymax; numMPI = in.npy;
  } else {
    L.nUni = nint((flt[5] - flt[4]) / in.dz) + 1;
    gridSize = in.dz; frontEdgeCoor = flt[4]; backEdgeCoor = flt[5];
    minCoor = in.zmin; maxCoor = in.zmax; numMPI = in.npz;
  }
  double coorTmp = frontEdgeCoor, gridSizeTmp = gridSize;
  int i;
  for (i = 1; i <= np; ++i) {
    gridSizeTmp = gridSizeTmp * in.rat;
    coorTmp = coorTmp - gridSizeTmp;
    if (coorTmp <= minCoor) break;
  }
  L.frontEdge = i + in.nPML;
  coorTmp = backEdgeCoor; gridSizeTmp = gridSize;
  for (i = 1; i <= np; ++i) {
    gridSizeTmp = gridSizeTmp * in.rat;
    coorTmp = coorTmp + gridSizeTmp;
    if (coorTmp >= maxCoor) break;
  }
  if (dimId == 3) i = -in.nPML;
  L.nGlobal = L.nUni + L.frontEdge + i + in.nPML;
  std::vector<double> g(L.nGlobal + 1, 0.0);  // 1-based
  int perMPI = (L.nGlobal + numMPI - 1) / numMPI;
  int resid = (L.nGlobal + numMPI - 1) - perMPI * numMPI;
  L.nLocal = (mpiId < (numMPI - resid)) ? perMPI : perMPI + 1;
  g[L.frontEdge + 1] = frontEdgeCoor;
  gridSizeTmp = gridSize;
  for (i = L.frontEdge; i >= 1; --i) {
    gridSizeTmp = gridSizeTmp * in.rat;
    g[i] = g[i + 1] - gridSizeTmp;
  }
  for (i = L.frontEdge + 2; i <= L.frontEdge + L.nUni; ++i) g[i] = g[i - 1] + gridSize;
  if (dimId < 3) {
    gridSizeTmp = gridSize;
    for (i = L.frontEdge + L.nUni + 1; i <= L.nGlobal; ++i) {
      gridSizeTmp = gridSizeTmp * in.rat;
      g[i] = g[i - 1] + gridSizeTmp;
    }
  }
  int N = L.nGlobal;
  bound[0] = g[1]; bound[1] = g[N];
  if (dimId == 1) {
    PMLb[0] = g[N - in.nPML]; PMLb[1] = g[in.nPML + 1]; PMLb[5] = g[N] - g[N - 1];
  } else if (dimId == 2) {
    PMLb[2] = g[N - in.nPML]; PMLb[3] = g[in.nPML + 1]; PMLb[6] = g[N] - g[N - 1];
  } else {
    PMLb[4] = g[in.nPML + 1]; PMLb[7] = g[2] - g[1];
  }
  L.local.resize(L.nLocal);
  if (mpiId <= (numMPI - resid)) {
    for (i = 1; i <= L.nLocal; ++i) L.local[i - 1] = g[(perMPI - 1) * mpiId + i];
  } else {
    for (i = 1; i <= L.nLocal; ++i)
      L.local[i - 1] = g[(perMPI - 1) * mpiId + i + (mpiId - numMPI + resid)];
  }
  return L;
}

struct Lines {
  Line x, y, z;
};

Lines all_lines(const CaseInput& in, RankState& s) {
  // calcXyzMPIId, meshgen.f90:437-445
  s.mex = s.me / (in.npy * in.npz);
  s.mey = (s.me - s.mex * in.npy * in.npz) / in.npz;
  s.mez = s.me - s.mex * in.npy * in.npz - s.mey * in.npz;
  Lines L;
  double b[2];
  L.x = local_line(in, 1, s.mex, b, s.PMLb); s.xminB = b[0]; s.xmaxB = b[1];
  L.y = local_line(in, 2, s.mey, b, s.PMLb); s.yminB = b[0]; s.ymaxB = b[1];
  L.z = local_line(in, 3, s.mez, b, s.PMLb); s.zminB = b[0]; s.zmaxB = b[1];
  s.nx = L.x.nLocal; s.ny = L.y.nLocal; s.nz = L.z.nLocal;
  return L;
}

// checkIsOnFault, meshgen.f90:768-787 (also the in-line copy in mesh4num.f90:48-58)
inline bool is_on_fault(const CaseInput& in, const double* c, int ift) {
  const double* f = &in.fltxyz[8 * (size_t)ift];
  const double tol = in.tol;
  if (c[0] >= (f[0] - tol) && c[0] <= (f[1] + tol) && c[1] >= (f[2] - tol) && c[1] <= (f[3] + tol) &&
      c[2] >= (f[4] - tol) && c[2] <= (f[5] + tol)) {
    if (in.C_degen == 0.0 && c[1] == 0.0) return true;
    if (in.C_degen > 3.0) {
      double t = std::tan(in.C_degen / 180.0 * in.pi);
      double d = std::fabs(c[2] + c[1] * t);
      d = d / std::sqrt(1.0 + t * t);
      if (d < in.dx / 100.0) return true;
    }
  }
  return false;
}

inline bool in_pml(const double* PMLb, double x, double y, double z) {
  return x > PMLb[0] || x < PMLb[1] || y > PMLb[2] || y < PMLb[3] || z < PMLb[4];
}

}  // namespace

eqd_params RankState::params() const {
  const CaseInput& c = *in;
  eqd_params p{};
  p.dt = c.dt; p.nstep = c.nstep; p.me = me; p.npx = c.npx; p.npy = c.npy; p.npz = c.npz;
  p.rdampk = c.rdampk; p.rdampm = c.rdampm; p.w = c.w; p.grav = c.grav; p.roumax = c.roumax;
  p.rhow = c.rhow; p.gamar = c.gamar; p.ccosphi = c.ccosphi; p.sinphi = c.sinphi; p.tv = c.tv;
  p.kapa_hg = c.kapa_hg; p.dx = c.dx; p.C_elastic = c.C_elastic; p.C_Q = c.C_Q; p.C_hg = c.C_hg;
  for (int i = 0; i < 8; ++i) p.PMLb[i] = PMLb[i];
  p.nPML = c.nPML; p.R = c.R; p.vmaxPML = c.vmaxPML;
  p.friclaw = c.friclaw; p.C_nuclea = c.C_nuclea; p.nucfault = c.nucfault; p.TPV = c.TPV;
  p.insertFaultType = c.insertFaultType; p.ntotft = c.ntotft;
  p.nucR = c.nucR; p.nucT = c.nucT; p.nucRuptVel = c.nucRuptVel; p.nucdtau0 = c.nucdtau0;
  p.xsource = c.xsource; p.ysource = c.ysource; p.zsource = c.zsource;
  p.slipRateThres = c.slipRateThres; p.tol = c.tol; p.fric_tp_h = c.fric_tp_h;
  p.outputGroundMotion = c.outputGroundMotion;
  return p;
}

// mesh4num.f90:3-89 -- counts only
void mesh4num(const CaseInput& in, RankState& s) {
  Lines L = all_lines(in, s);
  const int nx = s.nx, ny = s.ny, nz = s.nz;
  const double tol = in.tol;
  long nodeCount = 0, elementCount = 0, equationNumCount = 0, eqSize = 0;
  s.nftnd.assign(in.ntotft, 0);
  const double tanDip = std::tan(in.C_degen / 180.0 * in.pi);
  const double* f1 = in.fltxyz.data();
  for (int ix = 1; ix <= nx; ++ix)
    for (int iz = 1; iz <= nz; ++iz)
      for (int iy = 1; iy <= ny; ++iy) {
        double c[3] = {L.x.local[ix - 1], L.y.local[iy - 1], L.z.local[iz - 1]};
        nodeCount++;
        int numOfDof = in_pml(s.PMLb, c[0], c[1], c[2]) ? 12 : 3;
        bool fixed = std::fabs(c[0] - s.xminB) < tol || std::fabs(c[0] - s.xmaxB) < tol ||
                     std::fabs(c[1] - s.yminB) < tol || std::fabs(c[1] - s.ymaxB) < tol ||
                     std::fabs(c[2] - s.zminB) < tol;
        eqSize += numOfDof;
        if (!fixed) equationNumCount += numOfDof;
        for (int ift = 0; ift < in.ntotft; ++ift) {
          if (is_on_fault(in, c, ift)) {
            s.nftnd[ift]++;
            nodeCount++;
            equationNumCount += 3;
            eqSize += 3;
            break;
          }
        }
        if (ix >= 2 && iy >= 2 && iz >= 2) {
          elementCount++;
          if (in.C_degen > 3.0) {
            // wedge4num, library_degeneration.f90:57-75
            double cenx = c[0] - in.dx / 2.0, ceny = c[1] - in.dy / 2.0, cenz = c[2] - in.dz / 2.0;
            double d = std::fabs(ceny * tanDip + cenz) / std::sqrt(1.0 + tanDip * tanDip);
            if (cenx > f1[0] && cenx < f1[1] && ceny > f1[2] && ceny < f1[3] && cenz > f1[4] &&
                d < in.dx / 100.0)
              elementCount++;
          }
        }
      }
  if (nodeCount > 2000000000L || eqSize > 2000000000L / 5)
    throw std::runtime_error("sub-domain too large for 4-byte indices (reference limit)");
  s.sizeOfEqNumIndexArr = (int)eqSize;
  s.totalNumOfNodes = (int)nodeCount;
  s.totalNumOfElements = (int)elementCount;
  s.totalNumOfEquations = (int)equationNumCount;
}

// meshgen.f90:3-156 with its helper subroutines
void meshgen(const CaseInput& in, RankState& s) {
  Lines L = all_lines(in, s);
  const int nx = s.nx, ny = s.ny, nz = s.nz;
  const int Nn = s.totalNumOfNodes, Ne = s.totalNumOfElements;
  const double tol = in.tol;
  const int ntotft = in.ntotft;
  // allocInit, eqdyna3d.f90:98-150
  s.eqNumIndexArr.assign(s.sizeOfEqNumIndexArr, 0);
  s.eqNumStartIndexLoc.assign(Nn, 0);
  s.numOfDofPerNodeArr.assign(Nn, 0);
  s.meshCoor.assign((size_t)3 * Nn, 0.0);
  s.fnms.assign(Nn, 0.0);
  s.surfaceNodeIdArr.clear();
  s.nodeElemIdRelation.assign((size_t)8 * Ne, 0);
  s.mat.assign((size_t)Ne * 5, 0.0);
  s.elemTypeArr.assign(Ne, 0);
  s.eleporep.assign(Ne, 0.0);
  s.pstrain.assign(Ne, 0.0);
  s.stressCompIndexArr.assign(Ne, 0);
  s.stressArr.assign((size_t)5 * s.sizeOfEqNumIndexArr, 0.0);
  int nftmx = 0;
  for (int v : s.nftnd) nftmx = std::max(nftmx, v);
  if (nftmx <= 0) nftmx = 1;
  s.nftmx = nftmx;
  int nonmx = 0;
  for (int v : in.nonfs) nonmx += v;
  s.nonmx = nonmx;
  int mxOn = 0;
  for (int v : in.nonfs) mxOn = std::max(mxOn, v);
  s.nsmp.assign((size_t)2 * nftmx * ntotft, 0);
  s.fnft.assign((size_t)nftmx * ntotft, 99999.0);
  s.un.assign((size_t)3 * nftmx * ntotft, 0.0);
  s.us.assign((size_t)3 * nftmx * ntotft, 1000.0);
  s.ud.assign((size_t)3 * nftmx * ntotft, 0.0);
  s.fric.assign((size_t)100 * nftmx * ntotft, 0.0);
  s.arn.assign((size_t)nftmx * ntotft, 0.0);
  s.anonfs.assign((size_t)3 * std::max(nonmx, 1), 0);
  s.fltgm.assign(nftmx, 0);
  s.OffFaultStNodeIdIndex.assign((size_t)2 * std::max(in.totalNumOfOffSt, 1), 0);
  for (int k = 0; k < 9; ++k) s.numcount[k] = 0;
  s.numcount[0] = nx; s.numcount[1] = ny; s.numcount[2] = nz;
  for (int k = 0; k < 6; ++k) s.fltnum[k] = 0;

  int nodeCount = 0, elemCount = 0, equationNumCount = 0, eqTag = 0, stressDofCount = 0;
  int msnode = nx * ny * nz;
  s.numOfOnFaultStCount = 0;
  s.numOfOffFaultStCount = 0;
  std::vector<int> nftnd0(ntotft, 0), ixfi(ntotft, 0), izfi(ntotft, 0), ifs(ntotft, 0), ifd(ntotft, 0);
  std::vector<int> n4yn(std::max(in.totalNumOfOffSt, 1), 0);
  const int nxuni = L.x.nUni, nzuni = L.z.nUni;
  std::vector<int> fltrc((size_t)2 * nxuni * nzuni * ntotft, 0);
  auto FLTRC = [&](int a, int j, int i, int ift) -> int& {
    return fltrc[(a - 1) + 2 * ((size_t)(j - 1) + (size_t)nxuni * ((i - 1) + (size_t)nzuni * ift))];
  };
  std::unordered_map<int, int> slave2master;
  double pfx = 0.0, pfz = 0.0, ycoort = 0.0;
  const double* xline = L.x.local.data();
  const double* yline = L.y.local.data();
  const double* zline = L.z.local.data();
  const double* f1 = in.fltxyz.data();
  const double tanDip = std::tan(in.C_degen / 180.0 * in.pi);
  auto nodeId = [&](int ix, int iy, int iz) { return (ix - 1) * ny * nz + (iz - 1) * ny + iy; };
  double* X = s.meshCoor.data();

  for (int ix = 1; ix <= nx; ++ix) {
    for (int iz = 1; iz <= nz; ++iz) {
      for (int iy = 1; iy <= ny; ++iy) {
        // createNode, meshgen.f90:904-919
        double nodeCoor[3] = {xline[ix - 1], yline[iy - 1], zline[iz - 1]};
        nodeCount++;
        X[0 + 3 * (size_t)(nodeCount - 1)] = nodeCoor[0];
        X[1 + 3 * (size_t)(nodeCount - 1)] = nodeCoor[1];
        X[2 + 3 * (size_t)(nodeCount - 1)] = nodeCoor[2];
        if (in.insertFaultType > 0) {
          // insertFaultInterface, meshgen.f90:921-962
          double fx1 = in.rough_fx_min, fx2 = in.rough_fx_max, fz1 = in.rough_fz_min;
          int ixx = 1, izz = 1;
          double x = nodeCoor[0], z = nodeCoor[2];
          if ((x < fx2 + tol) && (x > fx1 - tol) && (z > fz1 - tol)) {
            ixx = nint((x - fx1) / in.dx) + 1; izz = nint((z - fz1) / in.dz) + 1;
          } else if ((x < fx1 - tol) && (z > fz1 - tol)) {
            ixx = 1; izz = nint((z - fz1) / in.dz) + 1;
          } else if ((x > fx2 + tol) && (z > fz1 - tol)) {
            ixx = in.nnx; izz = nint((z - fz1) / in.dz) + 1;
          } else if ((x < fx2 + tol) && (x > fx1 - tol) && (z < fz1 - tol)) {
            ixx = nint((x - fx1) / in.dx) + 1; izz = 1;
          } else if ((x < fx1 - tol) && (z < fz1 - tol)) {
            ixx = 1; izz = 1;
          } else if ((x > fx2 + tol) && (z < fz1 - tol)) {
            ixx = in.nnx; izz = 1;
          }
          size_t idx = (size_t)in.nnz * (ixx - 1) + izz - 1;
          double peak = in.rough_geo[0 + 3 * idx];
          pfx = in.rough_geo[1 + 3 * idx];
          pfz = in.rough_geo[2 + 3 * idx];
          // ymax / ymin are the globals overwritten with modelBoundCoor at meshgen.f90:33-34
          if (nodeCoor[1] > -tol) ycoort = nodeCoor[1] * (s.ymaxB - peak) / s.ymaxB + peak;
          else if (nodeCoor[1] < -tol) ycoort = nodeCoor[1] * (peak - s.yminB) / (-s.yminB) + peak;
          X[1 + 3 * (size_t)(nodeCount - 1)] = ycoort;
        }
        // setNumDof, meshgen.f90:562-572
        int numDof = in_pml(s.PMLb, nodeCoor[0], nodeCoor[1], nodeCoor[2]) ? 12 : 3;
        s.eqNumStartIndexLoc[nodeCount - 1] = eqTag;
        s.numOfDofPerNodeArr[nodeCount - 1] = numDof;
        // setEquationNumber, meshgen.f90:660-699 (xmin..zmin are modelBoundCoor here, :31-36)
        {
          bool fixed = std::fabs(nodeCoor[0] - s.xminB) < tol || std::fabs(nodeCoor[0] - s.xmaxB) < tol ||
                       std::fabs(nodeCoor[1] - s.yminB) < tol || std::fabs(nodeCoor[1] - s.ymaxB) < tol ||
                       std::fabs(nodeCoor[2] - s.zminB) < tol;
          for (int iDof = 1; iDof <= numDof; ++iDof) {
            if (fixed) {
              s.eqNumIndexArr[eqTag++] = -1;
            } else {
              equationNumCount++;
              s.eqNumIndexArr[eqTag++] = equationNumCount;
              if (ix == 1) s.numcount[3]++;
              if (ix == nx) s.numcount[4]++;
              if (iy == 1) s.numcount[5]++;
              if (iy == ny) s.numcount[6]++;
              if (iz == 1) s.numcount[7]++;
              if (iz == nz) s.numcount[8]++;
            }
          }
        }
        // setSurfaceStation, meshgen.f90:574-658
        if (in.totalNumOfOffSt > 0) {
          auto yMatch = [&](int i) {
            double ys = in.x4nds[1 + 3 * (size_t)i];
            return std::fabs(nodeCoor[1] - ys) < tol ||
                   (ys > yline[iy - 2] && ys < nodeCoor[1] && (nodeCoor[1] - ys) < (ys - yline[iy - 2])) ||
                   (ys > nodeCoor[1] && ys < yline[iy] && (ys - nodeCoor[1]) < (yline[iy] - ys));
          };
          auto found = [&](int i) {
            n4yn[i] = 1;
            s.numOfOffFaultStCount++;
            s.OffFaultStNodeIdIndex[0 + 2 * (size_t)(s.numOfOffFaultStCount - 1)] = i + 1;
            s.OffFaultStNodeIdIndex[1 + 2 * (size_t)(s.numOfOffFaultStCount - 1)] = nodeCount;
          };
          if (ix > 1 && ix < nx && iy > 1 && iy < ny) {
            for (int i = 0; i < in.totalNumOfOffSt; ++i) {
              if (n4yn[i] != 0) continue;
              double xs = in.x4nds[0 + 3 * (size_t)i];
              if (std::fabs(nodeCoor[2] - in.x4nds[2 + 3 * (size_t)i]) < tol) {
                if (std::fabs(nodeCoor[0] - xs) < tol ||
                    (xs > xline[ix - 2] && xs < nodeCoor[0] && (nodeCoor[0] - xs) < (xs - xline[ix - 2])) ||
                    (xs > nodeCoor[0] && xs < xline[ix] && (xs - nodeCoor[0]) < (xline[ix] - xs))) {
                  if (yMatch(i)) { found(i); break; }
                }
              }
            }
          }
          if (ix == 1 && iy > 1 && iy < ny) {
            for (int i = 0; i < in.totalNumOfOffSt; ++i) {
              if (n4yn[i] != 0) continue;
              double xs = in.x4nds[0 + 3 * (size_t)i];
              if (std::fabs(nodeCoor[2] - in.x4nds[2 + 3 * (size_t)i]) < tol) {
                if (std::fabs(nodeCoor[0] - xs) < tol ||
                    (xs > nodeCoor[0] && xs < xline[ix] && (xs - nodeCoor[0]) < (xline[ix] - xs))) {
                  if (yMatch(i)) { found(i); break; }
                }
              }
            }
          }
          if (ix == nx && iy > 1 && iy < ny) {
            for (int i = 0; i < in.totalNumOfOffSt; ++i) {
              if (n4yn[i] != 0) continue;
              double xs = in.x4nds[0 + 3 * (size_t)i];
              if (std::fabs(nodeCoor[2] - in.x4nds[2 + 3 * (size_t)i]) < tol) {
                if (xs > xline[ix - 2] && xs < nodeCoor[0] && (nodeCoor[0] - xs) < (xs - xline[ix - 2])) {
                  if (yMatch(i)) { found(i); break; }
                }
              }
            }
          }
        }
        // createMasterNode, meshgen.f90:789-902
        for (int ift = 0; ift < ntotft; ++ift) {
          if (!is_on_fault(in, nodeCoor, ift)) continue;
          nftnd0[ift]++;
          const int ip = nftnd0[ift];  // 1-based pair index
          if (ip > s.nftnd[ift]) throw std::runtime_error("meshgen: more fault nodes than mesh4num counted");
          size_t pb = (size_t)(ip - 1) + (size_t)nftmx * ift;
          s.nsmp[0 + 2 * pb] = nodeCount;
          msnode = nx * ny * nz + ip;
          if (ift > 0) throw std::runtime_error("msnode cannot handle iFault>1");
          s.eqNumStartIndexLoc[msnode - 1] = eqTag;
          s.numOfDofPerNodeArr[msnode - 1] = 3;
          s.nsmp[1 + 2 * pb] = msnode;
          slave2master[nodeCount] = msnode;
          X[0 + 3 * (size_t)(msnode - 1)] = nodeCoor[0];
          X[1 + 3 * (size_t)(msnode - 1)] = nodeCoor[1];
          X[2 + 3 * (size_t)(msnode - 1)] = nodeCoor[2];
          if (in.insertFaultType > 0) X[1 + 3 * (size_t)(msnode - 1)] = ycoort;
          for (int i = 0; i < 3; ++i) {
            equationNumCount++;
            s.eqNumIndexArr[eqTag++] = equationNumCount;
          }
          if (ix == 1) { s.fltgm[ip - 1] += 1; s.fltnum[0]++; }
          if (ix == nx) { s.fltgm[ip - 1] += 2; s.fltnum[1]++; }
          if (iy == 1) { s.fltgm[ip - 1] += 10; s.fltnum[2]++; }
          if (iy == ny) { s.fltgm[ip - 1] += 20; s.fltnum[3]++; }
          if (iz == 1) { s.fltgm[ip - 1] += 100; s.fltnum[4]++; }
          if (iz == nz) { s.fltgm[ip - 1] += 200; s.fltnum[5]++; }
          for (int i = 0; i < in.nonfs[ift]; ++i) {
            if (std::fabs(nodeCoor[0] - in.xonfs[0 + 2 * (i + (size_t)mxOn * ift)]) < tol &&
                std::fabs(nodeCoor[2] - in.xonfs[1 + 2 * (i + (size_t)mxOn * ift)]) < tol) {
              s.numOfOnFaultStCount++;
              int k = s.numOfOnFaultStCount - 1;
              s.anonfs[0 + 3 * (size_t)k] = ip;
              s.anonfs[1 + 3 * (size_t)k] = i + 1;
              s.anonfs[2 + 3 * (size_t)k] = ift + 1;
              break;
            }
          }
          const double strike = in.fltxyz[6 + 8 * (size_t)ift], dip = in.fltxyz[7 + 8 * (size_t)ift];
          double* un = &s.un[3 * pb]; double* us = &s.us[3 * pb]; double* ud = &s.ud[3 * pb];
          un[0] = std::cos(strike) * std::sin(dip);
          un[1] = -std::sin(strike) * std::sin(dip);
          un[2] = std::cos(dip);
          us[0] = -std::sin(strike);
          us[1] = -std::cos(strike);
          us[2] = 0.0;
          ud[0] = std::cos(strike) * std::cos(dip);
          ud[1] = std::sin(strike) * std::cos(dip);
          ud[2] = std::sin(dip);
          if (in.insertFaultType > 0) {
            double a = std::sqrt(pfx * pfx + 1.0 + pfz * pfz);
            un[0] = -pfx / a; un[1] = 1.0 / a; un[2] = -pfz / a;
            double b = std::sqrt(1.0 + pfx * pfx);
            us[0] = 1.0 / b; us[1] = pfx / b; us[2] = 0.0;
            ud[0] = us[1] * un[2] - us[2] * un[1];
            ud[1] = us[2] * un[0] - us[0] * un[2];
            ud[2] = us[0] * un[1] - us[1] * un[0];
          }
          if (ixfi[ift] == 0) ixfi[ift] = ix;
          if (izfi[ift] == 0) izfi[ift] = iz;
          ifs[ift] = ix - ixfi[ift] + 1;
          ifd[ift] = iz - izfi[ift] + 1;
          if (ifs[ift] < 1 || ifs[ift] > nxuni || ifd[ift] < 1 || ifd[ift] > nzuni)
            throw std::runtime_error("meshgen: fltrc index out of range");
          FLTRC(1, ifs[ift], ifd[ift], ift) = msnode;
          FLTRC(2, ifs[ift], ifd[ift], ift) = ip;
        }
        // elements
        if (ix >= 2 && iy >= 2 && iz >= 2) {
          // createElement, meshgen.f90:702-741
          elemCount++;
          if (elemCount > Ne) throw std::runtime_error("meshgen: more elements than mesh4num counted");
          int brick[8] = {nodeId(ix - 1, iy - 1, iz - 1), nodeId(ix, iy - 1, iz - 1), nodeId(ix, iy, iz - 1),
                          nodeId(ix - 1, iy, iz - 1),     nodeId(ix - 1, iy - 1, iz), nodeId(ix, iy - 1, iz),
                          nodeId(ix, iy, iz),             nodeId(ix - 1, iy, iz)};
          int* conn = &s.nodeElemIdRelation[8 * (size_t)(elemCount - 1)];
          s.elemTypeArr[elemCount - 1] = 1;
          for (int k = 0; k < 8; ++k) conn[k] = brick[k];
          s.stressCompIndexArr[elemCount - 1] = stressDofCount;
          double cen[3] = {0, 0, 0};
          for (int i = 0; i < 8; ++i)
            for (int j = 0; j < 3; ++j) cen[j] = cen[j] + X[j + 3 * (size_t)(conn[i] - 1)];
          for (int j = 0; j < 3; ++j) cen[j] = cen[j] / 8.0;
          if (in_pml(s.PMLb, cen[0], cen[1], cen[2])) {
            s.elemTypeArr[elemCount - 1] = 2;
            stressDofCount += 15 + 6;
          } else {
            stressDofCount += 12;
          }
          // setElementMaterial, meshgen.f90:161-207
          auto MAT = [&](int e, int j) -> double& { return s.mat[(size_t)(e - 1) + (size_t)Ne * (j - 1)]; };
          auto MATERIAL = [&](int i, int j) { return in.material[(size_t)(i - 1) + (size_t)in.nmat * (j - 1)]; };
          if (in.nmat == 1 && in.n2mat == 3) {
            MAT(elemCount, 1) = MATERIAL(1, 1); MAT(elemCount, 2) = MATERIAL(1, 2); MAT(elemCount, 3) = MATERIAL(1, 3);
          } else if (in.nmat > 1 && in.n2mat == 4) {
            double az = std::fabs(cen[2]);
            if (az < MATERIAL(1, 1)) {
              MAT(elemCount, 1) = MATERIAL(1, 2); MAT(elemCount, 2) = MATERIAL(1, 3); MAT(elemCount, 3) = MATERIAL(1, 4);
            } else {
              for (int i = 2; i <= in.nmat; ++i)
                if (az < MATERIAL(i, 1) && az >= MATERIAL(i - 1, 1)) {
                  MAT(elemCount, 1) = MATERIAL(i, 2); MAT(elemCount, 2) = MATERIAL(i, 3); MAT(elemCount, 3) = MATERIAL(i, 4);
                }
            }
          }
          MAT(elemCount, 5) = MAT(elemCount, 2) * MAT(elemCount, 2) * MAT(elemCount, 3);
          MAT(elemCount, 4) = MAT(elemCount, 1) * MAT(elemCount, 1) * MAT(elemCount, 3) - 2.0 * MAT(elemCount, 5);
          if (in.C_degen > 3.0) {
            // wedge, library_degeneration.f90:3-55
            double d = std::fabs(cen[1] * tanDip + cen[2]) / std::sqrt(1.0 + tanDip * tanDip);
            if (cen[0] > f1[0] && cen[0] < f1[1] && cen[1] > f1[2] && cen[1] < f1[3] && cen[2] > f1[4] && d < tol) {
              static const int o11[8] = {5, 1, 4, 4, 6, 2, 3, 3};
              static const int o12[8] = {4, 8, 5, 5, 3, 7, 6, 6};
              s.elemTypeArr[elemCount - 1] = 11;
              for (int k = 0; k < 8; ++k) conn[k] = brick[o11[k] - 1];
              elemCount++;
              if (elemCount > Ne) throw std::runtime_error("meshgen: more elements than mesh4num counted");
              s.elemTypeArr[elemCount - 1] = 12;
              conn = &s.nodeElemIdRelation[8 * (size_t)(elemCount - 1)];
              for (int k = 0; k < 8; ++k) conn[k] = brick[o12[k] - 1];
              s.stressCompIndexArr[elemCount - 1] = stressDofCount;
              stressDofCount += 12;
              MAT(elemCount, 1) = MATERIAL(1, 1); MAT(elemCount, 2) = MATERIAL(1, 2); MAT(elemCount, 3) = MATERIAL(1, 3);
              MAT(elemCount, 5) = MAT(elemCount, 2) * MAT(elemCount, 2) * MAT(elemCount, 3);
              MAT(elemCount, 4) = MAT(elemCount, 1) * MAT(elemCount, 1) * MAT(elemCount, 3) - 2 * MAT(elemCount, 5);
            }
            // meshgen.f90:91-97
            for (int k = 0; k < 2; ++k) {
              if (is_on_fault(in, &X[3 * (size_t)(conn[k] - 1)], 0) && s.elemTypeArr[elemCount - 1] == 1)
                s.elemTypeArr[elemCount - 1] = 13;
            }
          }
          // replaceSlaveWithMasterNode, meshgen.f90:743-766
          {
            int et = s.elemTypeArr[elemCount - 1];
            if ((et == 1 && (nodeCoor[0] > (f1[0] - tol) && nodeCoor[0] < (f1[1] + in.dx + tol) &&
                             nodeCoor[2] > (f1[4] - tol) && nodeCoor[1] > 0.0 &&
                             std::fabs(nodeCoor[1] - in.dy) < tol)) ||
                et == 12 || et == 13) {
              for (int k = 0; k < 8; ++k) {
                auto it = slave2master.find(conn[k]);
                if (it != slave2master.end()) conn[k] = it->second;
              }
            }
          }
          // setPlasticStress, meshgen.f90:976-996
          if (in.C_elastic == 0) {
            double depth = -0.5 * (zline[iz - 1] + zline[iz - 2]) + 7.3215;
            int etTag = (s.elemTypeArr[elemCount - 1] == 2) ? 1 : 0;
            s.eleporep[elemCount - 1] = 0.0;
            double strVert = -(in.roumax - in.rhow * (in.gamar + 1.0)) * depth * in.grav;
            double devStr = std::fabs(strVert) * in.devStrToStrVertRatio;
            double* sa = &s.stressArr[(size_t)s.stressCompIndexArr[elemCount - 1] + 15 * etTag];
            sa[2] = strVert;
            sa[0] = strVert - devStr * std::cos(2.0 * in.str1ToFaultAngle);
            sa[1] = strVert + devStr * std::cos(2.0 * in.str1ToFaultAngle);
            sa[5] = devStr * std::sin(2.0 * in.str1ToFaultAngle);
          }
        }
      }
    }
  }
  s.sizeOfStressDofIndexArr = stressDofCount;
  // meshGenError, meshgen.f90:405-435
  if (stressDofCount >= 5 * (long)s.sizeOfEqNumIndexArr) throw std::runtime_error("meshgen stop 2002");
  if (nodeCount != nx * ny * nz || elemCount != s.totalNumOfElements || equationNumCount != s.totalNumOfEquations)
    throw std::runtime_error("meshgen stop 2003: inconsistency between meshgen and mesh4num");
  {
    int ms = nx * ny * nz;
    for (int v : nftnd0) ms += v;
    if (ms != s.totalNumOfNodes) throw std::runtime_error("meshgen stop 2003: msnode /= totalNumOfNodes");
  }
  if (eqTag != s.sizeOfEqNumIndexArr) throw std::runtime_error("meshgen stop 2004");
  for (int i = 0; i < ntotft; ++i)
    if (nftnd0[i] != s.nftnd[i]) throw std::runtime_error("meshgen stop 2005");

  // fault node areas, meshgen.f90:114-152
  for (int ift = 0; ift < ntotft; ++ift) {
    if (nftnd0[ift] <= 0) continue;
    auto dist = [&](int a, int b) {
      const double* p = &X[3 * (size_t)(a - 1)];
      const double* q = &X[3 * (size_t)(b - 1)];
      return std::sqrt((q[0] - p[0]) * (q[0] - p[0]) + (q[1] - p[1]) * (q[1] - p[1]) + (q[2] - p[2]) * (q[2] - p[2]));
    };
    for (int i = 2; i <= ifd[ift]; ++i)
      for (int j = 2; j <= ifs[ift]; ++j) {
        int n1 = FLTRC(1, j, i, ift), n2 = FLTRC(1, j - 1, i, ift), n3 = FLTRC(1, j - 1, i - 1, ift), n4 = FLTRC(1, j, i - 1, ift);
        int m1 = FLTRC(2, j, i, ift), m2 = FLTRC(2, j - 1, i, ift), m3 = FLTRC(2, j - 1, i - 1, ift), m4 = FLTRC(2, j, i - 1, ift);
        if (!n1 || !n2 || !n3 || !n4) throw std::runtime_error("meshgen: incomplete fault grid (fltrc)");
        double aa1 = dist(n1, n2), bb1 = dist(n2, n3), cc1 = dist(n3, n4), dd1 = dist(n4, n1);
        double p1 = dist(n2, n4), q1 = dist(n1, n3);
        double t = (bb1 * bb1 + dd1 * dd1 - aa1 * aa1 - cc1 * cc1);
        double area = 0.25 * std::sqrt(4 * p1 * p1 * q1 * q1 - t * t);
        area = 0.25 * area;
        double* arn = &s.arn[(size_t)nftmx * ift];
        arn[m1 - 1] += area; arn[m2 - 1] += area; arn[m3 - 1] += area; arn[m4 - 1] += area;
      }
  }
  // first half of MPI4arn (meshgen.f90:221-254): face lists of split-node pairs
  for (int k = 0; k < 6; ++k) s.fltface[k].clear();
  for (int k = 0; k < 6; ++k) s.fltnum[k] = 0;
  {
    int n = nftnd0[ntotft - 1];
    for (int i = 1; i <= n; ++i) {
      int g = s.fltgm[i - 1];
      if (g % 10 == 1) { s.fltnum[0]++; s.fltface[0].push_back(i); }
      if (g % 10 == 2) { s.fltnum[1]++; s.fltface[1].push_back(i); }
      if (g % 100 - g % 10 == 10) { s.fltnum[2]++; s.fltface[2].push_back(i); }
      if (g % 100 - g % 10 == 20) { s.fltnum[3]++; s.fltface[3].push_back(i); }
      if (g - g % 100 == 100) { s.fltnum[4]++; s.fltface[4].push_back(i); }
      if (g - g % 100 == 200) { s.fltnum[5]++; s.fltface[5].push_back(i); }
    }
  }
  // fltMPI flags (meshgen.f90:255-402): a face exchanges split nodes iff it is
  // an interior rank face and holds at least one pair
  {
    int npxyz[3] = {in.npx, in.npy, in.npz};
    int mexyz[3] = {s.mex, s.mey, s.mez};
    for (int a = 0; a < 3; ++a) {
      bool lo = npxyz[a] > 1 && mexyz[a] != 0;
      bool hi = npxyz[a] > 1 && mexyz[a] != npxyz[a] - 1;
      // reference quirk kept: `if (mex == 0) bndl=0 elseif (mex == np-1) bndr=0`
      s.fltMPI[2 * a] = (lo && s.fltnum[2 * a] > 0) ? 1 : 0;
      s.fltMPI[2 * a + 1] = (hi && s.fltnum[2 * a + 1] > 0) ? 1 : 0;
    }
  }
}

// second half of MPI4arn, meshgen.f90:255-402: add the neighbours' arn on shared
// rank faces, axis x then y then z.  Within one axis a rank first exchanges its
// "-" face, then its "+" face; the two faces hold disjoint pairs, so every value
// sent is the pre-phase value and a snapshot-then-add reproduces the blocking
// mpi_sendrecv sequence exactly.
void exchange_arn(const CaseInput& in, std::vector<RankState*>& world) {
  const int stride[3] = {in.npy * in.npz, in.npz, 1};
  for (int ift = 0; ift < in.ntotft; ++ift)
    for (int a = 0; a < 3; ++a) {
      std::vector<std::vector<double>> lo(world.size()), hi(world.size());
      for (size_t r = 0; r < world.size(); ++r) {
        RankState& s = *world[r];
        for (int i : s.fltface[2 * a]) lo[r].push_back(s.arn[(size_t)(i - 1) + (size_t)s.nftmx * ift]);
        for (int i : s.fltface[2 * a + 1]) hi[r].push_back(s.arn[(size_t)(i - 1) + (size_t)s.nftmx * ift]);
      }
      for (size_t r = 0; r < world.size(); ++r) {
        RankState& s = *world[r];
        for (int side = 0; side < 2; ++side) {
          if (!s.fltMPI[2 * a + side]) continue;
          int nb = s.me + (side == 0 ? -stride[a] : stride[a]);
          const std::vector<double>& rv = side == 0 ? hi[nb] : lo[nb];
          const std::vector<int>& mine = s.fltface[2 * a + side];
          if (rv.size() != mine.size()) throw std::runtime_error("exchange_arn: face list mismatch");
          for (size_t k = 0; k < rv.size(); ++k) s.arn[(size_t)(mine[k] - 1) + (size_t)s.nftmx * ift] += rv[k];
        }
      }
    }
}

// netcdf_read_on_fault_eqdyna, netcdf_io.f90:70-107
void load_on_fault(const CaseInput& in, RankState& s) {
  const int fnx = in.fnx, fnz = in.fnz;
  auto OFV = [&](int ii, int jj, int v) {
    return in.on_fault_vars[(size_t)(ii - 1) + (size_t)fnx * ((jj - 1) + (size_t)fnz * (v - 1))];
  };
  for (int ift = 0; ift < in.ntotft; ++ift)
    for (int i = 1; i <= s.nftnd[ift]; ++i) {
      int slave = s.nsmp[0 + 2 * ((size_t)(i - 1) + (size_t)s.nftmx * ift)];
      double xc = s.meshCoor[0 + 3 * (size_t)(slave - 1)], zc = s.meshCoor[2 + 3 * (size_t)(slave - 1)];
      int ii = nint((xc - in.fxmin[ift]) / in.dx) + 1;
      int jj = nint((zc - in.fzmin[ift]) / in.dz) + 1;
      if (ii < 1 || ii > fnx || jj < 1 || jj > fnz) throw std::runtime_error("load_on_fault: index out of range");
      // reference writes fault 1's slab for every ift (fric(k,i,1)), netcdf_io.f90:76
      double* f = &s.fric[100 * (size_t)(i - 1)];
      auto F = [&](int k) -> double& { return f[k - 1]; };
      F(1) = OFV(ii, jj, 1); F(2) = OFV(ii, jj, 2); F(3) = OFV(ii, jj, 3);
      F(9) = OFV(ii, jj, 4); F(10) = OFV(ii, jj, 5); F(11) = OFV(ii, jj, 6);
      F(12) = OFV(ii, jj, 7); F(13) = OFV(ii, jj, 8); F(14) = OFV(ii, jj, 9);
      F(15) = OFV(ii, jj, 10); F(16) = OFV(ii, jj, 11); F(17) = OFV(ii, jj, 12);
      F(18) = OFV(ii, jj, 13); F(19) = OFV(ii, jj, 14); F(40) = OFV(ii, jj, 15);
      F(41) = OFV(ii, jj, 16); F(42) = OFV(ii, jj, 17); F(46) = OFV(ii, jj, 18);
      F(8) = OFV(ii, jj, 19); F(7) = OFV(ii, jj, 20); F(20) = OFV(ii, jj, 21);
      F(47) = F(46); F(25) = 0.0; F(26) = F(46); F(27) = 0.0;
      F(5) = OFV(ii, jj, 22); F(4) = OFV(ii, jj, 23); F(49) = OFV(ii, jj, 24);
      F(23) = std::fabs(F(7));
    }
}

// netcdf_read_on_fault_eqdyna_restart, netcdf_io.f90:152-178 (mode == 2, eqdyna3d.f90:60): initial
// tractions, slip rate, state variables and the split-node velocities of the previous cycle.
// Unlike load_on_fault this one indexes fric by the fault it loops over (fric(k,i,ift)).
void load_on_fault_restart(const CaseInput& in, RankState& s) {
  const int fnx = in.fnx, fnz = in.fnz;
  if (in.restart_vars.size() != (size_t)fnx * fnz * 12) throw std::runtime_error("load_on_fault_restart: no restart fields");
  auto RV = [&](int ii, int jj, int v) {
    return in.restart_vars[(size_t)(ii - 1) + (size_t)fnx * ((jj - 1) + (size_t)fnz * (v - 1))];
  };
  static const int slot[12] = {8, 49, 7, 47, 20, 23, 31, 32, 33, 34, 35, 36};
  for (int ift = 0; ift < in.ntotft; ++ift)
    for (int i = 1; i <= s.nftnd[ift]; ++i) {
      const size_t pb = (size_t)(i - 1) + (size_t)s.nftmx * ift;
      int slave = s.nsmp[0 + 2 * pb];
      double xc = s.meshCoor[0 + 3 * (size_t)(slave - 1)], zc = s.meshCoor[2 + 3 * (size_t)(slave - 1)];
      int ii = nint((xc - in.fxmin[ift]) / in.dx) + 1;
      int jj = nint((zc - in.fzmin[ift]) / in.dz) + 1;
      if (ii < 1 || ii > fnx || jj < 1 || jj > fnz) throw std::runtime_error("load_on_fault_restart: index out of range");
      double* f = &s.fric[100 * pb];
      for (int v = 1; v <= 12; ++v) f[slot[v - 1] - 1] = RV(ii, jj, v);
    }
}

// find_surfaceNodeIdArr, library_output.f90:247-264
void find_surface_nodes(const CaseInput& in, RankState& s) {
  s.surface_nnode = 0;
  s.surfaceNodeIdArr.clear();
  if (!(in.outputGroundMotion == 1 || in.outputFinalSurfDisp == 1)) return;
  const double* f = in.fltxyz.data();
  for (int i = 1; i <= s.totalNumOfNodes; ++i) {
    const double* c = &s.meshCoor[3 * (size_t)(i - 1)];
    if (c[0] < f[1] + 20.0e3 && c[0] > f[0] - 20.0e3 && c[1] < f[3] + 20.0e3 && c[1] > f[2] - 20.0e3 &&
        std::fabs(c[2]) < in.dx / 1000) {
      s.surface_nnode++;
      s.surfaceNodeIdArr.push_back(i);
    }
  }
}

// allocInitAfterMeshGen, eqdyna3d.f90:153-189
void alloc_after_meshgen(const CaseInput& in, RankState& s) {
  s.nOnAlloc = s.numOfOnFaultStCount <= 0 ? 1 : s.numOfOnFaultStCount;
  s.onFaultQuantHistSCECForm.assign((size_t)12 * in.nstep * s.nOnAlloc, 0.0);
  s.nodalForceArr.assign(s.totalNumOfEquations, 0.0);
  s.v1.assign(s.totalNumOfEquations, 0.0);
  s.nodalMassArr.assign(s.totalNumOfEquations, 0.0);
  s.velArr.assign((size_t)3 * s.totalNumOfNodes, 0.0);
  s.dispArr.assign((size_t)3 * s.totalNumOfNodes, 0.0);
  s.hypoLog.assign((size_t)13 * in.nstep, 0.0);
  if (in.friclaw == 5) s.onFaultTPHist.assign((size_t)2 * s.nftmx * in.nstep * in.ntotft, 0.0);
  s.nGmAlloc = 0; s.nGmSamples = 0;
  s.gmHist.clear(); s.srcEvolHist.clear();
  if (in.outputGroundMotion == 1) {   // one sample at every nt with mod(nt,10) == 1 (driver.f90:30)
    s.nGmAlloc = in.nstep / 10 + 1;
    s.gmHist.assign((size_t)3 * s.surface_nnode * s.nGmAlloc, 0.0);
    s.srcEvolHist.assign((size_t)(s.nftnd.empty() ? 0 : s.nftnd[0]) * s.nGmAlloc, 0.0);
  }
  s.idhist.clear();
  s.OffFaultStGramSCEC.clear();
  if (s.numOfOffFaultStCount > 0) {
    int n = s.numOfOffFaultStCount * 6;
    s.idhist.assign((size_t)3 * n, 0);
    s.OffFaultStGramSCEC.assign((size_t)(n + 1) * in.nstep, 0.0);
    int row = 0;
    for (int iSt = 1; iSt <= s.numOfOffFaultStCount; ++iSt)
      for (int iDof = 1; iDof <= 3; ++iDof)
        for (int dv = 1; dv <= 2; ++dv) {
          s.idhist[0 + 3 * (size_t)row] = s.OffFaultStNodeIdIndex[1 + 2 * (size_t)(iSt - 1)];
          s.idhist[1 + 3 * (size_t)row] = iDof;
          s.idhist[2 + 3 * (size_t)row] = dv;
          row++;
        }
  }
}

}  // namespace eqh
