// ORACLE -- TEST INFRASTRUCTURE ONLY.  CPU restatement of EQdyna's explicit
// time-stepping loop, used as the checker for the CUDA step library.  Only
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs may load this; the product never does.
//
// Parity pin: the reference itself cannot be built here (no gfortran / MPI /
// netCDF), so this restatement is pinned against the reference's
// golden results test.reference.results/*/frt.txt* (+ the drv.a6 station
// series) at the reference's own tolerance (check.test.py:40-52); see
// tests/test_oracle_golden.py and DESIGN.md for what stays unpinned (wedges, Q).
//
// Every function cites the reference lines it follows.  Arrays arrive through
// eqh_view (include/eqdyna_host.h) in the Fortran layout with 1-based ids, and
// the loop order, operation order and quirks of the Fortran are kept
// (compile with -O2 -ffp-contract=off: the reference's ubuntu build has no FMA).
// Multi-sub-domain runs are in-process: one eqh_view per MPI rank of the
// reference, an in-memory exchange reproducing the x -> y -> z blocking
// mpi_sendrecv + add of MPI4NodalQuant, and one OpenMP thread per sub-domain.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "eqdyna_host.h"

namespace {

struct Ctx {
  eqh_view* v;
  int nt;
  double timeElapsed;
};

#define P (c.v->params)

inline double sq(double x) { return x * x; }

// comdampv.f90:1-120
int comdampv(const Ctx& c, double x2, double y, double z, double dv[9]) {
  const double xmax0 = P.PMLb[0], xmin0 = P.PMLb[1], ymax0 = P.PMLb[2], ymin0 = P.PMLb[3], zmin0 = P.PMLb[4];
  const double maxd[3] = {P.PMLb[5], P.PMLb[6], P.PMLb[7]};
  double damp[3] = {0, 0, 0};
  bool any = false;
  if (z <= zmin0) { damp[2] = std::fabs(z - zmin0); any = true; }
  else if (z > zmin0) { damp[2] = 0.0; any = true; }
  if (any) {
    if (x2 >= xmax0 && y >= ymax0) { damp[0] = std::fabs(x2 - xmax0); damp[1] = std::fabs(y - ymax0); }
    else if (x2 >= xmax0 && y <= ymin0) { damp[0] = std::fabs(x2 - xmax0); damp[1] = std::fabs(y - ymin0); }
    else if (x2 <= xmin0 && y <= ymin0) { damp[0] = std::fabs(x2 - xmin0); damp[1] = std::fabs(y - ymin0); }
    else if (x2 <= xmin0 && y >= xmax0) { damp[0] = std::fabs(x2 - xmin0); damp[1] = std::fabs(y - ymax0); }  // sic: y>=xmax0 (:29,:61)
    else if (x2 >= xmax0 && y > ymin0 && y < ymax0) { damp[0] = std::fabs(x2 - xmax0); damp[1] = 0.0; }
    else if (y <= ymin0 && x2 > xmin0 && x2 < xmax0) { damp[0] = 0.0; damp[1] = std::fabs(y - ymin0); }
    else if (x2 <= xmin0 && y > ymin0 && y < ymax0) { damp[0] = std::fabs(x2 - xmin0); damp[1] = 0.0; }
    else if (y >= ymax0 && x2 > xmin0 && x2 < xmax0) { damp[0] = 0.0; damp[1] = std::fabs(y - ymax0); }
    else { damp[0] = 0.0; damp[1] = 0.0; }
  }
  for (int i = 0; i < 3; ++i) {
    double delta = P.nPML * maxd[i];
    damp[i] = 3.0 * P.vmaxPML / 2.0 / delta * std::log(1.0 / P.R) * ((damp[i] / delta) * (damp[i] / delta));
  }
  for (int i = 0; i < 9; ++i) dv[i] = damp[i % 3];
  for (int i = 0; i < 9; ++i)
    if (dv[i] < 0.0) return EQD_ERR_DAMP;
  return 0;
}

// velDispUpdate, driver.f90:89-155
int vel_disp_update(Ctx& c) {
  eqh_view& v = *c.v;
  const double dt = P.dt;
  for (int i = 1; i <= v.Nn; ++i) {
    const int st = v.eqNumStartIndexLoc[i - 1];
    double* vel = &v.velArr[3 * (size_t)(i - 1)];
    double* dis = &v.dispArr[3 * (size_t)(i - 1)];
    if (v.numOfDofPerNodeArr[i - 1] == 3) {
      for (int j = 1; j <= 3; ++j) {
        int eq = v.eqNumIndexArr[st + j - 1];
        v.v1[eq - 1] = v.v1[eq - 1] + v.nodalForceArr[eq - 1] * dt;
        vel[j - 1] = v.v1[eq - 1];
        dis[j - 1] = dis[j - 1] + v.v1[eq - 1] * dt;
      }
    } else if (v.numOfDofPerNodeArr[i - 1] == 12) {
      double dampv[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
      int eq1 = v.eqNumIndexArr[st];
      if (eq1 > 0) {
        const double* x = &v.meshCoor[3 * (size_t)(i - 1)];
        int rc = comdampv(c, x[0], x[1], x[2], dampv);
        if (rc) return rc;
      }
      for (int j = 1; j <= 9; ++j) {
        int eq = v.eqNumIndexArr[st + j - 1];
        if (eq > 0)
          v.v1[eq - 1] = (v.nodalForceArr[eq - 1] + v.v1[eq - 1] * (1.0 / dt - dampv[j - 1] / 2.0)) /
                         (1.0 / dt + dampv[j - 1] / 2.0);
      }
      for (int j = 10; j <= 12; ++j) {
        int eq = v.eqNumIndexArr[st + j - 1];
        if (eq > 0) v.v1[eq - 1] = v.v1[eq - 1] + v.nodalForceArr[eq - 1] * dt;
      }
      if (eq1 > 0) {
        auto V = [&](int j) { return v.v1[v.eqNumIndexArr[st + j - 1] - 1]; };
        vel[0] = V(1) + V(2) + V(3) + V(10);
        vel[1] = V(4) + V(5) + V(6) + V(11);
        vel[2] = V(7) + V(8) + V(9) + V(12);
        dis[0] = dis[0] + vel[0] * dt;
        dis[1] = dis[1] + vel[1] * dt;
        dis[2] = dis[2] + vel[2] * dt;
      } else if (eq1 == -1) {
        vel[0] = vel[1] = vel[2] = 0.0;
        dis[0] = dis[1] = dis[2] = 0.0;
      }
    }
    if (vel[0] != vel[0] || vel[1] != vel[1] || vel[2] != vel[2]) return EQD_ERR_NAN;
  }
  return 0;
}

// storeOffFaultStData, driver.f90:157-180
void store_off_fault(Ctx& c) {
  eqh_view& v = *c.v;
  int n = v.nOff * 6;
  if (n <= 0) return;
  double* row = &v.OffFaultStGramSCEC[(size_t)(n + 1) * (c.nt - 1)];
  row[0] = c.timeElapsed;
  for (int i = 1; i <= n; ++i) {
    int node = v.idhist[0 + 3 * (size_t)(i - 1)], k = v.idhist[1 + 3 * (size_t)(i - 1)], q = v.idhist[2 + 3 * (size_t)(i - 1)];
    if (q == 1) row[i] = v.dispArr[(k - 1) + 3 * (size_t)(node - 1)];
    else if (q == 2) row[i] = v.velArr[(k - 1) + 3 * (size_t)(node - 1)];
    else if (q == 3) row[i] = v.nodalForceArr[v.eqNumIndexArr[v.eqNumStartIndexLoc[node - 1] + k - 1] - 1];
  }
}

// qconstant.f90:3-35 -- the tables are SINGLE-precision literals in the reference
void qconstant(double Q, double& rtaok, double& rwk, int k, double& c1) {
  static const float taok[8] = {1.72333e-3f, 1.80701e-3f, 5.38887e-3f, 1.99322e-2f, 8.49833e-2f, 4.09335e-1f, 2.05951f, 13.2629f};
  static const float alfk[8] = {1.66958e-2f, 3.81644e-2f, 9.84666e-3f, -1.36803e-2f, -2.85125e-2f, -5.37309e-2f, -6.65035e-2f, -1.33696e-1f};
  static const float betk[8] = {8.98758e-2f, 6.84635e-2f, 9.67052e-2f, 1.20172e-1f, 1.30728e-1f, 1.38746e-1f, 1.40705e-1f, 2.14647e-1f};
  const double pi = 4 * std::atan(1.0);
  double kapa = (double)3.071f + (double)1.433f * std::pow(Q, (double)(-1.158f)) * std::log(Q / 5);
  kapa = kapa / (1 + (double)0.415f * Q);
  rwk = kapa * (kapa * (double)alfk[k - 1] + (double)betk[k - 1]);
  rtaok = (double)taok[k - 1];
  double ref = 2.0 * pi;
  double ak0 = 1.0 - rwk * 8.0 / (1.0 + sq(rtaok * ref));
  double bk0 = rwk * 8.0 * ref * rtaok / (1.0 + sq(rtaok * ref));
  c1 = 0.5 * std::pow(ak0 * ak0 + bk0 * bk0, -0.5);
  c1 = c1 * (1.0 + ak0 * std::pow(ak0 * ak0 + bk0 * bk0, -0.5));
}

// calcElemKU.f90:3-191 (calcB.f90 folded in: bb(1,j)=shp(1,i) ... standard sparsity)
void calc_elem_ku(const Ctx& c, const double* shp /*(3,8)*/, const double mate[5], const double vl[24],
                  const double dl[24], double* stress /*(12)*/, double elresf[24], double constk, double porep,
                  double& pstrmag, const double ex[24]) {
  const double dt = P.dt;
  pstrmag = 0.0;
  double stressrate[6] = {0, 0, 0, 0, 0, 0}, strain[6] = {0, 0, 0, 0, 0, 0}, strainrate[6] = {0, 0, 0, 0, 0, 0};
  const double lam = mate[3], miu = mate[4];
  double cc[6][6];
  std::memset(cc, 0, sizeof cc);
  for (int i = 0; i < 3; ++i) { cc[i][i] = lam + 2 * miu; cc[i + 3][i + 3] = miu; }
  cc[0][1] = cc[1][0] = cc[0][2] = cc[2][0] = cc[1][2] = cc[2][1] = lam;
  for (int i = 0; i < 8; ++i) {
    const double s1 = shp[0 + 3 * i], s2 = shp[1 + 3 * i], s3 = shp[2 + 3 * i];
    const int j1 = 3 * i, j2 = 3 * i + 1, j3 = 3 * i + 2;
    strainrate[0] = strainrate[0] + s1 * vl[j1];
    strainrate[1] = strainrate[1] + s2 * vl[j2];
    strainrate[2] = strainrate[2] + s3 * vl[j3];
    strainrate[3] = strainrate[3] + s3 * vl[j2] + s2 * vl[j3];
    strainrate[4] = strainrate[4] + s3 * vl[j1] + s1 * vl[j3];
    strainrate[5] = strainrate[5] + s2 * vl[j1] + s1 * vl[j2];
    strain[0] = strain[0] + s1 * dl[j1];
    strain[1] = strain[1] + s2 * dl[j2];
    strain[2] = strain[2] + s3 * dl[j3];
    strain[3] = strain[3] + s3 * dl[j2] + s2 * dl[j3];
    strain[4] = strain[4] + s3 * dl[j1] + s1 * dl[j3];
    strain[5] = strain[5] + s2 * dl[j1] + s1 * dl[j2];
  }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) stressrate[i] = stressrate[i] + cc[i][j] * strainrate[j];
  for (int i = 3; i < 6; ++i) stressrate[i] = cc[i][i] * strainrate[i];
  double strdev[6];
  if (P.C_Q == 0) {
    for (int i = 0; i < 6; ++i) {
      stress[i] = stress[i] + stressrate[i] * dt;
      strdev[i] = stress[i];
    }
  } else if (P.C_Q == 1) {
    double xc[3] = {0, 0, 0};
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 8; ++j) xc[i] = xc[i] + ex[i + 3 * j];
    for (int i = 0; i < 3; ++i) xc[i] = xc[i] / 8.0;
    double Qs, Qp;
    if (xc[2] > -1000.0) { Qs = 10.0; Qp = 20.0; } else { Qs = 50.0; Qp = 100.0; }
    // integer assignment of a real expression truncates toward zero
    int ip = (int)((xc[0] - (P.PMLb[1] + P.dx / 2)) / P.dx + 1);
    int iq = (int)((xc[1] - (P.PMLb[3] + P.dx / 2)) / P.dx + 1);
    int ir = (int)((xc[2] - (P.PMLb[4] + P.dx / 2)) / P.dx + 1);
    int k = 1 + ip % 2 + 2 * (iq % 2) + 4 * (ir % 2);
    double taok, wkp, wks, cv, cs;
    qconstant(Qp, taok, wkp, k, cv);
    qconstant(Qs, taok, wks, k, cs);
    wkp = wkp * 8.0;
    wks = wks * 8.0;
    double miuu = miu * cs, Mu = (lam + 2 * miu) * cv;
    double vols = strain[0] + strain[1] + strain[2];
    double anestr1[6], anestr[6];
    for (int i = 0; i < 6; ++i) anestr1[i] = stress[i + 6];
    for (int i = 0; i < 3; ++i)
      anestr[i] = std::exp(-dt / taok) * anestr1[i] +
                  (1 - std::exp(-dt / taok)) * (2 * miuu * strain[i] * wks + (Mu * wkp - 2 * miuu * wks) * vols);
    for (int i = 3; i < 6; ++i)
      anestr[i] = std::exp(-dt / taok) * anestr1[i] + (1 - std::exp(-dt / taok)) * (miuu * strain[i] * wks);
    for (int i = 0; i < 6; ++i) stress[i + 6] = anestr[i];
    for (int i = 0; i < 3; ++i)
      stress[i] = 2.0 * miuu * strain[i] + (Mu - 2.0 * miuu) * vols - 0.5 * (anestr[i] + anestr1[i]);
    for (int i = 3; i < 6; ++i) stress[i] = 2.0 * miuu * strain[i] / 2.0 - 0.5 * (anestr[i] + anestr1[i]);
    // note: strdev is NOT assigned in this branch in the reference (only used when C_elastic==0,
    // which warning.f90 forbids together with C_Q==1)
    for (int i = 0; i < 6; ++i) strdev[i] = stress[i];
  }
  if (P.C_elastic == 0) {
    double strmea = (stress[0] + stress[1] + stress[2]) / 3.0;
    for (int i = 0; i < 3; ++i) strdev[i] = stress[i] - strmea;
    double taomax = 0.5 * (sq(strdev[0]) + sq(strdev[1]) + sq(strdev[2])) + sq(strdev[3]) + sq(strdev[4]) + sq(strdev[5]);
    taomax = std::sqrt(taomax);
    double yield = P.ccosphi - P.sinphi * (strmea + porep);
    if (yield < 0.0) yield = 0.0;
    if (taomax > yield) {
      double rjust = yield / taomax + (1 - yield / taomax) * std::exp(-dt / P.tv);
      double pstrinc[6];
      for (int i = 0; i < 6; ++i) {
        stress[i] = strdev[i] * rjust;
        pstrinc[i] = (strdev[i] - stress[i]) / miu;
        if (i < 3) stress[i] = stress[i] + strmea;
      }
      double pstrmea = (pstrinc[0] + pstrinc[1] + pstrinc[2]) / 3.0;
      for (int i = 0; i < 6; ++i) pstrinc[i] = pstrinc[i] - pstrmea;
      pstrmag = 0.5 * (sq(pstrinc[0]) + sq(pstrinc[1]) + sq(pstrinc[2])) + sq(pstrinc[3]) + sq(pstrinc[4]) + sq(pstrinc[5]);
      pstrmag = std::sqrt(pstrmag);
    }
  }
  const double temp = constk * P.w;
  double strtemp[6];
  for (int i = 0; i < 6; ++i) strtemp[i] = temp * (stress[i] + P.rdampk * stressrate[i]);
  for (int i = 0; i < 8; ++i) {
    const double s1 = shp[0 + 3 * i], s2 = shp[1 + 3 * i], s3 = shp[2 + 3 * i];
    double w1 = s1 * strtemp[0] + s3 * strtemp[4] + s2 * strtemp[5];
    double w2 = s2 * strtemp[1] + s3 * strtemp[3] + s1 * strtemp[5];
    double w3 = s3 * strtemp[2] + s2 * strtemp[3] + s1 * strtemp[4];
    elresf[3 * i] = elresf[3 * i] + w1;
    elresf[3 * i + 1] = elresf[3 * i + 1] + w2;
    elresf[3 * i + 2] = elresf[3 * i + 2] + w3;
  }
}

// calcPMLElemKU, assembleGlobalKU.f90:70-346
void calc_pml_elem_ku(const Ctx& c, const double vl[24], double f[96], double* s /*(21)*/, const double ex[24],
                      const double mat1[5], const double* shp /*(3,8)*/, double det) {
  const double dt = P.dt, w = P.w;
  const double lam = mat1[3], miu = mat1[4];
  const double xmax2 = P.PMLb[0], xmin2 = P.PMLb[1], ymax2 = P.PMLb[2], ymin2 = P.PMLb[3], zmin2 = P.PMLb[4];
  const double maxd[3] = {P.PMLb[5], P.PMLb[6], P.PMLb[7]};
  double xc[3] = {0, 0, 0};
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 8; ++j) xc[i] = xc[i] + ex[i + 3 * j];
  for (int i = 0; i < 3; ++i) xc[i] = xc[i] / 8;
  double damps[3] = {0, 0, 0};
  bool any = false;
  if (xc[2] < zmin2) { damps[2] = std::fabs(xc[2] - zmin2); any = true; }
  else if (xc[2] > zmin2) { damps[2] = 0.0; any = true; }
  if (any) {
    if (xc[0] > xmax2 && xc[1] > ymax2) { damps[0] = std::fabs(xc[0] - xmax2); damps[1] = std::fabs(xc[1] - ymax2); }
    else if (xc[0] > xmax2 && xc[1] < ymin2) { damps[0] = std::fabs(xc[0] - xmax2); damps[1] = std::fabs(xc[1] - ymin2); }
    else if (xc[0] < xmin2 && xc[1] < ymin2) { damps[0] = std::fabs(xc[0] - xmin2); damps[1] = std::fabs(xc[1] - ymin2); }
    else if (xc[0] < xmin2 && xc[1] > xmax2) { damps[0] = std::fabs(xc[0] - xmin2); damps[1] = std::fabs(xc[1] - ymax2); }  // sic (:150,:182)
    else if (xc[0] > xmax2 && xc[1] > ymin2 && xc[1] < ymax2) { damps[0] = std::fabs(xc[0] - xmax2); damps[1] = 0.0; }
    else if (xc[1] < ymin2 && xc[0] > xmin2 && xc[0] < xmax2) { damps[0] = 0.0; damps[1] = std::fabs(xc[1] - ymin2); }
    else if (xc[0] < xmin2 && xc[1] > ymin2 && xc[1] < ymax2) { damps[0] = std::fabs(xc[0] - xmin2); damps[1] = 0.0; }
    else if (xc[1] > ymax2 && xc[0] > xmin2 && xc[0] < xmax2) { damps[0] = 0.0; damps[1] = std::fabs(xc[1] - ymax2); }
    else { damps[0] = 0.0; damps[1] = 0.0; }
  }
  for (int i = 0; i < 3; ++i) {
    double delta = P.nPML * maxd[i];
    damps[i] = 3 * P.vmaxPML / 2 / delta * std::log(1 / P.R) * ((damps[i] / delta) * (damps[i] / delta));
  }
  double strainrate[6] = {0, 0, 0, 0, 0, 0}, stressrate[6] = {0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 8; ++i) {
    const double s1 = shp[0 + 3 * i], s2 = shp[1 + 3 * i], s3 = shp[2 + 3 * i];
    const int j1 = 3 * i, j2 = 3 * i + 1, j3 = 3 * i + 2;
    strainrate[0] = strainrate[0] + s1 * vl[j1];
    strainrate[1] = strainrate[1] + s2 * vl[j2];
    strainrate[2] = strainrate[2] + s3 * vl[j3];
    strainrate[3] = strainrate[3] + s3 * vl[j2] + s2 * vl[j3];
    strainrate[4] = strainrate[4] + s3 * vl[j1] + s1 * vl[j3];
    strainrate[5] = strainrate[5] + s2 * vl[j1] + s1 * vl[j2];
  }
  double cm[6][6];
  std::memset(cm, 0, sizeof cm);
  for (int i = 0; i < 3; ++i) { cm[i][i] = lam + 2.0 * miu; cm[i + 3][i + 3] = miu; }
  cm[0][1] = cm[1][0] = cm[1][2] = cm[2][1] = cm[0][2] = cm[2][0] = lam;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) stressrate[i] = stressrate[i] + cm[i][j] * strainrate[j];
  for (int i = 3; i < 6; ++i) stressrate[i] = cm[i][i] * strainrate[i];
  double Dx_vx = 0, Dy_vy = 0, Dz_vz = 0, Dx_vy = 0, Dy_vx = 0, Dx_vz = 0, Dz_vx = 0, Dy_vz = 0, Dz_vy = 0;
  for (int i = 0; i < 8; ++i) {
    const double s1 = shp[0 + 3 * i], s2 = shp[1 + 3 * i], s3 = shp[2 + 3 * i];
    Dx_vx = Dx_vx + s1 * vl[3 * i];
    Dy_vy = Dy_vy + s2 * vl[3 * i + 1];
    Dz_vz = Dz_vz + s3 * vl[3 * i + 2];
    Dx_vy = Dx_vy + s1 * vl[3 * i + 1];
    Dy_vx = Dy_vx + s2 * vl[3 * i];
    Dx_vz = Dx_vz + s1 * vl[3 * i + 2];
    Dz_vx = Dz_vx + s3 * vl[3 * i];
    Dy_vz = Dy_vz + s2 * vl[3 * i + 2];
    Dz_vy = Dz_vy + s3 * vl[3 * i + 1];
  }
  auto upd = [&](int k, double coef, double D, int a) {
    s[k - 1] = coef * D + (1 / dt - damps[a] / 2) * s[k - 1];
    s[k - 1] = s[k - 1] / (1 / dt + damps[a] / 2);
  };
  upd(1, lam + 2 * miu, Dx_vx, 0); upd(2, lam, Dy_vy, 1); upd(3, lam, Dz_vz, 2);
  upd(4, lam, Dx_vx, 0); upd(5, lam + 2 * miu, Dy_vy, 1); upd(6, lam, Dz_vz, 2);
  upd(7, lam, Dx_vx, 0); upd(8, lam, Dy_vy, 1); upd(9, lam + 2 * miu, Dz_vz, 2);
  upd(10, miu, Dx_vy, 0); upd(11, miu, Dy_vx, 1);
  upd(12, miu, Dx_vz, 0); upd(13, miu, Dz_vx, 2);
  upd(14, miu, Dy_vz, 1); upd(15, miu, Dz_vy, 2);
  const double sxx = s[0] + s[1] + s[2], syy = s[3] + s[4] + s[5], szz = s[6] + s[7] + s[8];
  const double sxy = s[9] + s[10], sxz = s[11] + s[12], syz = s[13] + s[14];
  double s0[6];
  for (int i = 0; i < 6; ++i) s0[i] = s[15 + i] + P.rdampk * stressrate[i];
  for (int i = 0; i < 8; ++i) {
    const double s1 = shp[0 + 3 * i], s2 = shp[1 + 3 * i], s3 = shp[2 + 3 * i];
    double* fi = &f[12 * i];
    fi[0] = fi[0] - det * w * s1 * sxx;
    fi[1] = fi[1] - det * w * s2 * sxy;
    fi[2] = fi[2] - det * w * s3 * sxz;
    fi[3] = fi[3] - det * w * s1 * sxy;
    fi[4] = fi[4] - det * w * s2 * syy;
    fi[5] = fi[5] - det * w * s3 * syz;
    fi[6] = fi[6] - det * w * s1 * sxz;
    fi[7] = fi[7] - det * w * s2 * syz;
    fi[8] = fi[8] - det * w * s3 * szz;
    fi[9] = fi[9] - det * w * (s1 * s0[0] + s3 * s0[4] + s2 * s0[5]);
    fi[10] = fi[10] - det * w * (s2 * s0[1] + s3 * s0[3] + s1 * s0[5]);
    fi[11] = fi[11] - det * w * (s3 * s0[2] + s2 * s0[3] + s1 * s0[4]);
  }
}

// assembleGlobalKU, assembleGlobalKU.f90:3-67 (+ calcElemMass.f90)
void assemble_global_ku(Ctx& c) {
  eqh_view& v = *c.v;
  const int Ne = v.Ne;
  for (int nel = 1; nel <= Ne; ++nel) {
    const int* conn = &v.nodeElemIdRelation[8 * (size_t)(nel - 1)];
    double al[24], vl[24], dl[24], ex[24], elresf[24];
    for (int i = 0; i < 8; ++i)
      for (int j = 0; j < 3; ++j) {
        vl[3 * i + j] = v.velArr[j + 3 * (size_t)(conn[i] - 1)];
        dl[3 * i + j] = v.dispArr[j + 3 * (size_t)(conn[i] - 1)];
        ex[3 * i + j] = v.meshCoor[j + 3 * (size_t)(conn[i] - 1)];
        al[3 * i + j] = P.rdampm * vl[3 * i + j];
      }
    for (int i = 0; i < 8; ++i)
      al[3 * i + 2] = al[3 * i + 2] + (1.0 - P.C_elastic) * P.grav * (P.roumax - (P.gamar + 1.0) * P.rhow) / P.roumax;
    const double* em = &v.elemass[24 * (size_t)(nel - 1)];
    for (int k = 0; k < 24; ++k) elresf[k] = 0.0 - al[k] * em[k];
    double mate[5];
    for (int k = 0; k < 5; ++k) mate[k] = v.mat[(size_t)(nel - 1) + (size_t)Ne * k];
    const int et = v.elemTypeArr[nel - 1];
    const double* shp = &v.eleshp[24 * (size_t)(nel - 1)];
    double* stress = &v.stressArr[v.stressCompIndexArr[nel - 1]];
    if (et == 1 || et > 10) {
      double pstrinc;
      calc_elem_ku(c, shp, mate, vl, dl, stress, elresf, -v.eledet[nel - 1], v.eleporep[nel - 1], pstrinc, ex);
      v.pstrain[nel - 1] = v.pstrain[nel - 1] + pstrinc;
      for (int i = 0; i < 8; ++i)
        for (int j = 1; j <= 3; ++j) {
          int eq = v.eqNumIndexArr[v.eqNumStartIndexLoc[conn[i] - 1] + j - 1];
          if (eq > 0) v.nodalForceArr[eq - 1] = v.nodalForceArr[eq - 1] + elresf[3 * i + j - 1];
        }
    } else if (et == 2) {
      double efPML[96];
      for (int k = 0; k < 96; ++k) efPML[k] = 0.0;
      for (int i = 0; i < 8; ++i)
        for (int j = 0; j < 3; ++j) efPML[12 * i + 9 + j] = elresf[3 * i + j];
      calc_pml_elem_ku(c, vl, efPML, stress, ex, mate, shp, v.eledet[nel - 1]);
      for (int i = 0; i < 8; ++i) {
        const int st = v.eqNumStartIndexLoc[conn[i] - 1];
        const double* e = &efPML[12 * i];
        if (v.numOfDofPerNodeArr[conn[i] - 1] == 12) {
          for (int j = 1; j <= 12; ++j) {
            int eq = v.eqNumIndexArr[st + j - 1];
            if (eq > 0) v.nodalForceArr[eq - 1] = v.nodalForceArr[eq - 1] + e[j - 1];
          }
        } else if (v.numOfDofPerNodeArr[conn[i] - 1] == 3) {
          int eq = v.eqNumIndexArr[st];
          v.nodalForceArr[eq - 1] = v.nodalForceArr[eq - 1] + e[0] + e[1] + e[2] + e[9];
          eq = v.eqNumIndexArr[st + 1];
          v.nodalForceArr[eq - 1] = v.nodalForceArr[eq - 1] + e[3] + e[4] + e[5] + e[10];
          eq = v.eqNumIndexArr[st + 2];
          v.nodalForceArr[eq - 1] = v.nodalForceArr[eq - 1] + e[6] + e[7] + e[8] + e[11];
        }
      }
    }
  }
}

// hrglss.f90:3-100
void hrglss(Ctx& c) {
  eqh_view& v = *c.v;
  const int Ne = v.Ne;
  for (int nel = 1; nel <= Ne; ++nel) {
    const int* conn = &v.nodeElemIdRelation[8 * (size_t)(nel - 1)];
    double vl[8][3], dl[8][3];
    for (int i = 0; i < 8; ++i)
      for (int j = 0; j < 3; ++j) {
        vl[i][j] = v.velArr[j + 3 * (size_t)(conn[i] - 1)];
        dl[i][j] = v.dispArr[j + 3 * (size_t)(conn[i] - 1)] + P.rdampk * vl[i][j];
      }
    auto target = [&](int i, int j) {  // j = 1..3
      int node = conn[i];
      int itag = 0;
      if (v.numOfDofPerNodeArr[node - 1] == 3) itag = v.eqNumStartIndexLoc[node - 1] + j;
      else if (v.numOfDofPerNodeArr[node - 1] == 12) itag = v.eqNumStartIndexLoc[node - 1] + j + 9;
      return v.eqNumIndexArr[itag - 1];
    };
    if (P.C_hg == 1) {
      const double* SS = &v.ss[6 * (size_t)(nel - 1)];
      const double* PHI = &v.phi[32 * (size_t)(nel - 1)];
      for (int m = 0; m < 4; ++m) {
        double phid[3];
        for (int i = 0; i < 3; ++i) {
          phid[i] = 0.0;
          for (int j = 0; j < 8; ++j) phid[i] = phid[i] + PHI[j + 8 * m] * dl[j][i];
        }
        for (int i = 0; i < 8; ++i) {
          double fhr[3];
          fhr[0] = PHI[i + 8 * m] * (SS[0] * phid[0] + SS[1] * phid[1] + SS[2] * phid[2]);
          fhr[1] = PHI[i + 8 * m] * (SS[1] * phid[0] + SS[3] * phid[1] + SS[4] * phid[2]);
          fhr[2] = PHI[i + 8 * m] * (SS[2] * phid[0] + SS[4] * phid[1] + SS[5] * phid[2]);
          for (int j = 1; j <= 3; ++j) {
            int k = target(i, j);
            if (k > 0) v.nodalForceArr[k - 1] = v.nodalForceArr[k - 1] - fhr[j - 1];
          }
        }
      }
    } else if (P.C_hg == 2) {
      static const int fi[4][8] = {{1, 1, -1, -1, -1, -1, 1, 1}, {1, -1, -1, 1, -1, 1, 1, -1},
                                   {1, -1, 1, -1, 1, -1, 1, -1}, {1, -1, 1, -1, -1, 1, -1, 1}};
      double rho = v.mat[(size_t)(nel - 1) + (size_t)Ne * 2], vp = v.mat[(size_t)(nel - 1)];
      double coef = 0.25 * P.kapa_hg * rho * vp * std::pow(v.eledet[nel - 1] * P.w, 2.0 / 3.0);
      double q[3][4], f[24];
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 4; ++j) {
          q[i][j] = 0.0;
          for (int k = 0; k < 8; ++k) q[i][j] = q[i][j] + vl[k][i] * fi[j][k];
        }
      for (int k = 0; k < 24; ++k) f[k] = 0.0;
      for (int k = 0; k < 8; ++k)
        for (int i = 0; i < 3; ++i)
          for (int j = 0; j < 4; ++j) f[3 * k + i] = f[3 * k + i] - coef * q[i][j] * fi[j][k];
      for (int i = 0; i < 8; ++i)
        for (int j = 1; j <= 3; ++j) {
          int k = target(i, j);
          if (k > 0) v.nodalForceArr[k - 1] = v.nodalForceArr[k - 1] + f[3 * i + j - 1];
        }
    }
  }
}

// MPI4NodalQuant(nodalForceArr, 3), assembleGlobalMass.f90:58-281, one axis.
// Face dof order: processNodalQuantArr over face nodes ((iz,iy) | (ix,iz) | (ix,iy)),
// then the face's split-node masters.
void face_slots(eqh_view& v, int a, int side, std::vector<double*>& out) {
  out.clear();
  const int nx = v.nx, ny = v.ny, nz = v.nz;
  const int n[3] = {nx, ny, nz};
  const int b = side == 0 ? 1 : n[a];
  auto push_node = [&](int node) {
    int st = v.eqNumStartIndexLoc[node - 1];
    for (int d = 1; d <= v.numOfDofPerNodeArr[node - 1]; ++d) {
      int eq = v.eqNumIndexArr[st + d - 1];
      if (eq > 0) out.push_back(&v.nodalForceArr[eq - 1]);
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
  if (v.fltMPI[2 * a + side])
    for (int k = 0; k < v.fltnum[2 * a + side]; ++k) push_node(nx * ny * nz + v.fltface[2 * a + side][k]);
}

int exchange_forces(eqh_view* views, int nranks) {
  const eqd_params& p0 = views[0].params;
  const int npxyz[3] = {p0.npx, p0.npy, p0.npz};
  const int stride[3] = {p0.npy * p0.npz, p0.npz, 1};
  for (int a = 0; a < 3; ++a) {
    if (npxyz[a] <= 1) continue;
    // The "-" and "+" faces of one rank are disjoint dof sets, so the blocking
    // sendrecv sequence (ib=1 then ib=2) sends pre-phase values: snapshot, then add.
    std::vector<std::vector<double>> snap[2];
    snap[0].resize(nranks);
    snap[1].resize(nranks);
#pragma omp parallel for schedule(static)
    for (int r = 0; r < nranks; ++r) {
      std::vector<double*> slots;
      for (int side = 0; side < 2; ++side) {
        face_slots(views[r], a, side, slots);
        snap[side][r].reserve(slots.size());
        for (double* q : slots) snap[side][r].push_back(*q);
      }
    }
    int bad = 0;
#pragma omp parallel for schedule(static)
    for (int r = 0; r < nranks; ++r) {
      eqh_view& v = views[r];
      const int me = v.params.me;
      const int mex = me / (p0.npy * p0.npz), mey = (me - mex * p0.npy * p0.npz) / p0.npz,
                mez = me - mex * p0.npy * p0.npz - mey * p0.npz;
      const int mexyz[3] = {mex, mey, mez};
      std::vector<double*> slots;
      for (int side = 0; side < 2; ++side) {
        bool active = side == 0 ? (mexyz[a] != 0) : (mexyz[a] != npxyz[a] - 1);
        if (!active) continue;
        int nb = me + (side == 0 ? -stride[a] : stride[a]);
        face_slots(v, a, side, slots);
        const std::vector<double>& rv = snap[1 - side][nb];
        if (rv.size() != slots.size()) { bad = 1; continue; }
        for (size_t k = 0; k < slots.size(); ++k) *slots[k] = *slots[k] + rv[k];
      }
    }
    if (bad) return EQD_ERR_ARG;
  }
  return 0;
}

// thermop.f90:1-40
void thermop(Ctx& c) {
  eqh_view& v = *c.v;
  const double dt = P.dt, pi = 4 * std::atan(1.0), h = P.fric_tp_h;
  const int nt = c.nt;
  for (int ift = 0; ift < v.ntotft; ++ift)
    for (int i = 1; i <= v.nftnd[ift]; ++i) {
      size_t pb = (size_t)(i - 1) + (size_t)v.nftmx * ift;
      double* f = &v.fric[100 * pb];
      double gama = f[18] / f[17], omega = f[15], kapa = f[16];
      auto H = [&](int a, int j) {  // onFaultTPHist(a,i,j,ift)
        return v.onFaultTPHist[(a - 1) + 2 * ((size_t)(i - 1) + (size_t)v.nftmx * ((size_t)(j - 1) + (size_t)v.nstep * ift))];
      };
      double tmp = 0.0;
      for (int j = 1; j <= nt - 1; ++j) {
        double ker = -kapa / (omega - kapa) / std::sqrt(4.0 * kapa * (nt - j) * dt + 2.0 * (h * h));
        ker = ker + omega / (omega - kapa) / std::sqrt(4.0 * omega * (nt - j) * dt + 2.0 * (h * h));
        tmp = tmp + std::fabs(H(2, j)) * H(1, j) * ker * dt;
      }
      double patnode = tmp * gama / std::sqrt(pi);
      tmp = 0.0;
      for (int j = 1; j <= nt - 1; ++j) {
        double ker = 1.0 / std::sqrt(4.0 * kapa * (nt - j) * dt + 2.0 * (h * h));
        tmp = tmp + std::fabs(H(2, j)) * H(1, j) * ker * dt;
      }
      double Tatnode = tmp / f[17] / std::sqrt(pi);
      f[50] = patnode;
      f[51] = Tatnode + f[40];
    }
}

// ---- fric.f90 --------------------------------------------------------------
void slip_weak(double slip, const double* fr, double& xmu) {  // fric.f90:3-19
  if (std::fabs(slip) < (double)1.0e-10f) xmu = fr[0];
  else if (slip < fr[2]) xmu = fr[0] - (fr[0] - fr[1]) * slip / fr[2];
  if (slip >= fr[2]) xmu = fr[1];
}
void time_weak(double trupt, const double* fr, double& xmu) {  // fric.f90:21-37
  if (trupt <= 0.0) xmu = fr[0];
  else if (trupt < fr[4]) xmu = fr[0] - (fr[0] - fr[1]) * trupt / fr[4];
  else xmu = fr[1];
}
void rate_state_ageing_law(double V2, double& theta, const double* fr, double& xmu, double& dxmudv, double dt) {  // fric.f90:39-61
  double A = fr[8], B = fr[9], L = fr[10], f0 = fr[12], V0 = fr[11];
  double tmpc = 1.0 / (2.0 * V0) * std::exp((f0 + B * std::log(V0 * theta / L)) / A);
  double tmp = (V2 + 1.e-30) * tmpc;
  xmu = A * std::log(tmp + std::sqrt(tmp * tmp + 1.0));
  dxmudv = A * tmpc / std::sqrt(1.0 + tmp * tmp);
  theta = L / V2 + (theta - L / V2) * std::exp(-V2 * dt / L);
}
void rate_state_slip_law(double V2, double& psi, const double* fr, double& xmu, double& dxmudv, double dt) {  // fric.f90:63-95
  double A = fr[8], B = fr[9], L = fr[10], f0 = fr[12], V0 = fr[11], fw = fr[13], Vw = fr[14];
  double tmpc = 1.0 / (2.0 * V0) * std::exp(psi / A);
  double tmp = (V2 + 1.e-30) * tmpc;
  xmu = A * std::log(tmp + std::sqrt(tmp * tmp + 1.0));
  dxmudv = A * tmpc / std::sqrt(1.0 + tmp * tmp);
  double fLV = f0 - (B - A) * std::log(V2 / V0);
  double r = V2 / Vw, r2 = r * r, r4 = r2 * r2;
  double fss = fw + (fLV - fw) / std::pow(1.0 + r4 * r4, 0.125);
  double fssa = fss / A;
  double psiss = A * std::log(2.0 * V0 / V2 * (std::exp(fssa) - std::exp(-fssa)) / 2.0);
  psi = psiss + (psi - psiss) * std::exp(-V2 * dt / L);
}

// rate_state_normal_stress, faulting.f90:35-52
void rate_state_normal_stress(double V2, double& theta_pc, double& theta_pc_dot, double tnrm, const double* fr, double dt) {
  double L = fr[10];
  theta_pc_dot = -V2 / L * (theta_pc - std::fabs(tnrm));
  theta_pc = theta_pc + theta_pc_dot * dt;
}

struct Pair {
  int ift, i;      // 0-based fault, 1-based pair
  double* fric;    // fric(1:100,i,ift)
  const double *un, *us, *ud;
  double arn;
  int slave, master;
  double *fS, *fM;  // nodal force of slave / master, x,y,z via eq map
};

// faulting.f90:3-541 for one pair
void fault_pair(Ctx& c, int ift, int i, double* thetaPcGarbage) {
  eqh_view& v = *c.v;
  const double dt = P.dt;
  size_t pb = (size_t)(i - 1) + (size_t)v.nftmx * ift;
  double* fr = &v.fric[100 * pb];
  auto F = [&](int k) -> double& { return fr[k - 1]; };
  const double* un = &v.un[3 * pb];
  const double* us = &v.us[3 * pb];
  const double* ud = &v.ud[3 * pb];
  const double arn = v.arn[pb];
  const int nodeS = v.nsmp[0 + 2 * pb], nodeM = v.nsmp[1 + 2 * pb];
  int eqS[3], eqM[3];
  for (int k = 0; k < 3; ++k) {
    eqS[k] = v.eqNumIndexArr[v.eqNumStartIndexLoc[nodeS - 1] + k];
    eqM[k] = v.eqNumIndexArr[v.eqNumStartIndexLoc[nodeM - 1] + k];
  }
  const double dtau = 0.0;  // faulting.f90:9, never changed
  // ---- getNsdSlipSliprateTraction, faulting.f90:54-134
  double initT[3] = {F(7), F(8) + dtau, F(49)};
  const double massSlave = v.fnms[nodeS - 1], massMaster = v.fnms[nodeM - 1];
  const double totalMass = (massSlave + massMaster) * arn;
  double xyz[3][2][3];  // [quant][slave/master][k]
  for (int k = 0; k < 3; ++k) {
    xyz[0][0][k] = v.nodalForceArr[eqS[k] - 1];
    xyz[0][1][k] = v.nodalForceArr[eqM[k] - 1];
    xyz[1][0][k] = v.velArr[k + 3 * (size_t)(nodeS - 1)];
    xyz[1][1][k] = v.velArr[k + 3 * (size_t)(nodeM - 1)];
    xyz[2][0][k] = v.dispArr[k + 3 * (size_t)(nodeS - 1)];
    xyz[2][1][k] = v.dispArr[k + 3 * (size_t)(nodeM - 1)];
  }
  double nsd[3][2][3];  // [quant][slave/master][n,s,d]
  for (int q = 0; q < 3; ++q)
    for (int k = 0; k < 2; ++k) {
      nsd[q][k][0] = xyz[q][k][0] * un[0] + xyz[q][k][1] * un[1] + xyz[q][k][2] * un[2];
      nsd[q][k][1] = xyz[q][k][0] * us[0] + xyz[q][k][1] * us[1] + xyz[q][k][2] * us[2];
      nsd[q][k][2] = xyz[q][k][0] * ud[0] + xyz[q][k][1] * ud[1] + xyz[q][k][2] * ud[2];
    }
  double slip[4], rate[4], T[4];
  for (int j = 0; j < 3; ++j) slip[j] = nsd[2][1][j] - nsd[2][0][j];
  slip[3] = std::sqrt(sq(slip[0]) + sq(slip[1]) + sq(slip[2]));
  for (int j = 0; j < 3; ++j) rate[j] = nsd[1][1][j] - nsd[1][0][j];
  rate[3] = std::sqrt(sq(rate[0]) + sq(rate[1]) + sq(rate[2]));
  F(71) = slip[1]; F(72) = slip[2]; F(73) = slip[0];
  F(74) = rate[1]; F(75) = rate[2];
  if (rate[3] > F(76)) F(76) = rate[3];
  F(77) = F(77) + rate[3] * dt;
  T[0] = (massSlave * massMaster * ((nsd[1][1][0] - nsd[1][0][0]) + (nsd[2][1][0] - nsd[2][0][0]) / dt) / dt +
          massSlave * nsd[0][1][0] - massMaster * nsd[0][0][0]) / totalMass + initT[0] * P.C_elastic;
  T[1] = (massSlave * massMaster * (nsd[1][1][1] - nsd[1][0][1]) / dt + massSlave * nsd[0][1][1] -
          massMaster * nsd[0][0][1]) / totalMass + initT[1] * P.C_elastic;
  T[2] = (massSlave * massMaster * (nsd[1][1][2] - nsd[1][0][2]) / dt + massSlave * nsd[0][1][2] -
          massMaster * nsd[0][0][2]) / totalMass + initT[2] * P.C_elastic;
  const double* xs = &v.meshCoor[3 * (size_t)(nodeS - 1)];
  if (P.friclaw >= 3 && P.C_nuclea == 1 && (ift + 1) == P.nucfault) {
    // rsfNucleation, faulting.f90:367-414
    double dtau2 = 0.0;
    double radius = std::sqrt(sq(xs[0] - P.xsource) + sq(xs[1] - P.ysource) + sq(xs[2] - P.zsource));
    double Fq = 0.0, G = 1.0;
    if (radius < P.nucR) Fq = std::exp(radius * radius / (radius * radius - P.nucR * P.nucR));
    if (c.timeElapsed <= P.nucT) G = std::exp(sq(c.timeElapsed - P.nucT) / (c.timeElapsed * (c.timeElapsed - 2.0 * P.nucT)));
    if (P.TPV == 105 || P.TPV == 104) dtau2 = P.nucdtau0 * Fq * G;
    else if (P.TPV == 2802) {
      if (c.nt == 1) {
        F(81) = P.nucdtau0;
        double ttao = std::sqrt(sq(T[1]) + sq(T[2]));
        double back = std::sqrt(sq(rate[1] + F(26)) + sq(rate[2] + F(27)));
        F(20) = F(9) * std::log(2.0 * F(12) / back * std::sinh(ttao / std::fabs(T[0]) / F(9)));
        F(23) = std::fabs(T[0]);
      }
      dtau2 = F(81) * Fq * G;
    }
    T[1] = T[1] + dtau2;
  }
  T[3] = std::sqrt(sq(T[1]) + sq(T[2]));

  if (P.friclaw <= 2) {
    // ---- solveSWTW, faulting.f90:136-188
    double mu = 0.0;
    if (P.friclaw == 1) slip_weak(F(77), fr, mu);
    else if (P.friclaw == 2) time_weak(c.timeElapsed - v.fnft[pb], fr, mu);
    if (P.C_nuclea == 1 && (ift + 1) == P.nucfault) {
      // swtwNucleation, faulting.f90:416-443
      double radius = std::sqrt(sq(xs[0] - P.xsource) + sq(xs[1] - P.ysource) + sq(xs[2] - P.zsource));
      double tr = 1.0e9;
      if (radius <= P.nucR) {
        if (P.TPV == 201 || P.TPV == 36 || P.TPV == 37)
          tr = (radius + 0.081 * P.nucR * (1.0 / (1.0 - sq(radius / P.nucR)) - 1.0)) / (0.7 * 3464.0);
        if (P.TPV == 202) tr = radius / P.nucRuptVel;
      }
      double tc = 1.0;
      if (c.timeElapsed < tr) tc = 0.0;
      else if ((c.timeElapsed < (tr + F(5))) && (c.timeElapsed >= tr)) tc = (c.timeElapsed - tr) / F(5);
      mu = std::fmin(F(1) + (F(2) - F(1)) * tc, mu);
    }
    double effN;
    if ((T[0] + F(6)) > 0) effN = 0.0;
    else effN = T[0] + F(6);
    double trial = F(4) - mu * effN;
    if (T[3] > trial) {
      T[1] = T[1] * trial / T[3];
      T[2] = T[2] * trial / T[3];
    }
    for (int j = 0; j < 3; ++j) {
      double xt = (T[0] * un[j] + T[1] * us[j] + T[2] * ud[j]) * arn;
      double x0 = (initT[0] * un[j] + initT[1] * us[j] + initT[2] * ud[j]) * arn;
      v.nodalForceArr[eqS[j] - 1] = v.nodalForceArr[eqS[j] - 1] + xt - x0 * P.C_elastic;
      v.nodalForceArr[eqM[j] - 1] = v.nodalForceArr[eqM[j] - 1] - xt + x0 * P.C_elastic;
    }
    for (int j = 0; j < 3; ++j) F(78 + j) = T[j];
  } else {
    // ---- solveRSF, faulting.f90:190-327
    if (P.friclaw == 5) T[0] = T[0] + F(51);
    else T[0] = T[0] + F(6);
    if (P.insertFaultType > 0 && P.C_elastic == 1) {
      const double max_norm = -40.0e6, min_norm = -10.0e6;
      if (T[0] >= min_norm) T[0] = min_norm;
      else if (T[0] <= max_norm) T[0] = max_norm;
    }
    if (T[0] > 0.0) T[0] = 0.0;
    for (int j = 0; j < 3; ++j) {
      slip[j] = slip[j] + F(25 + j) * c.timeElapsed;
      rate[j] = rate[j] + F(25 + j);
    }
    slip[3] = std::sqrt(sq(slip[1]) + sq(slip[2]));
    rate[3] = std::sqrt(sq(rate[1]) + sq(rate[2]));
    double v_trial = rate[3];
    double theta_pc_tmp = F(23), theta_pc_dot;
    rate_state_normal_stress(v_trial, F(23), theta_pc_dot, T[0], fr, dt);
    F(24) = theta_pc_dot;
    double statetmp = F(20);
    double xmu = 0, dxmudv = 0;
    if (P.friclaw == 3) rate_state_ageing_law(v_trial, F(20), fr, xmu, dxmudv, dt);
    else rate_state_slip_law(v_trial, F(20), fr, xmu, dxmudv, dt);
    double taoc_old;
    if (P.friclaw == 5) taoc_old = F(4) - xmu * T[0];
    else taoc_old = xmu * theta_pc_tmp;
    double mr = massMaster * massSlave / (massMaster + massSlave);
    double T_coeff = arn * dt / mr;
    double trialT[4] = {0, 0, 0, 0};
    for (int j = 1; j <= 2; ++j) trialT[j] = T[j] - taoc_old * 0.5 * (rate[j] / rate[3]) + F(25 + j) / T_coeff;
    trialT[3] = std::sqrt(sq(trialT[1]) + sq(trialT[2]));
    // NewtonRaphson, faulting.f90:459-516
    double taoc_new = 0.0;
    {
      const double state0 = statetmp, thetaPc0 = theta_pc_tmp;
      double stateTmp = state0;
      // thetaPcTmp is an UNINITIALISED local in the reference when friclaw==5; what the
      // golden tpv1053d run shows in frt column 22 is modelled by *thetaPcGarbage.
      double thetaPcTmp = *thetaPcGarbage;
      for (int iv = 1; iv <= 20; ++iv) {
        stateTmp = state0;
        if (P.friclaw == 3) rate_state_ageing_law(v_trial, stateTmp, fr, xmu, dxmudv, dt);
        else rate_state_slip_law(v_trial, stateTmp, fr, xmu, dxmudv, dt);
        double rsfeq, drsfeqdv;
        if (P.friclaw < 5) {
          thetaPcTmp = thetaPc0;
          rate_state_normal_stress(v_trial, thetaPcTmp, theta_pc_dot, T[0], fr, dt);
          taoc_new = xmu * thetaPcTmp;
          rsfeq = v_trial + T_coeff * (taoc_new * 0.5 - trialT[3]);
          drsfeqdv = 1.0 + T_coeff * (dxmudv * thetaPcTmp) * 0.5;
        } else {
          taoc_new = F(4) - xmu * std::fmin(T[0], 0.0);
          rsfeq = v_trial + T_coeff * (taoc_new * 0.5 - trialT[3]);
          drsfeqdv = 1.0 + T_coeff * (-dxmudv * std::fmin(T[0], 0.0)) * 0.5;
        }
        if (std::fabs(rsfeq / drsfeqdv) < 1.e-14 * std::fabs(v_trial) && std::fabs(rsfeq) < 1.e-6 * std::fabs(v_trial)) break;
        double newSliprate = v_trial - rsfeq / drsfeqdv;
        if (newSliprate <= 0.0) v_trial = v_trial / 2.0;
        else v_trial = newSliprate;
      }
      if (P.TPV == 105 && v_trial < F(46)) v_trial = F(46);
      statetmp = stateTmp;
      theta_pc_tmp = thetaPcTmp;
      *thetaPcGarbage = thetaPcTmp;
    }
    F(20) = statetmp;
    F(23) = theta_pc_tmp;
    for (int j = 1; j <= 2; ++j) T[j] = taoc_old * 0.5 * (rate[j] / rate[3]) + taoc_new * 0.5 * (trialT[j] / trialT[3]);
    for (int j = 0; j < 3; ++j) F(78 + j) = T[j];
    F(47) = v_trial;
    F(48) = std::sqrt(sq(T[1]) + sq(T[2]));
    if (v.onFaultTPHist) {
      size_t hb = 2 * ((size_t)(i - 1) + (size_t)v.nftmx * ((size_t)(c.nt - 1) + (size_t)v.nstep * ift));
      v.onFaultTPHist[hb] = F(47);
      v.onFaultTPHist[hb + 1] = F(48);
    }
    double acc[3];
    acc[0] = -rate[0] / dt - slip[0] / dt / dt;
    acc[1] = (v_trial * (trialT[1] / trialT[3]) - rate[1]) / dt;
    acc[2] = (v_trial * (trialT[2] / trialT[3]) - rate[2]) / dt;
    double xyzAcc[3], xyzR[3];
    for (int j = 0; j < 3; ++j) {
      xyzAcc[j] = acc[0] * un[j] + acc[1] * us[j] + acc[2] * ud[j];
      xyzR[j] = v.nodalForceArr[eqS[j] - 1] + v.nodalForceArr[eqM[j] - 1];
    }
    for (int j = 0; j < 3; ++j) {
      v.nodalForceArr[eqS[j] - 1] = (-xyzAcc[j] + xyzR[j] / massMaster) * mr;
      v.nodalForceArr[eqM[j] - 1] = (xyzAcc[j] + xyzR[j] / massSlave) * mr;
      F(31 + j) = v.velArr[j + 3 * (size_t)(nodeM - 1)] + (xyzAcc[j] + xyzR[j] / massSlave) * dt;
      F(34 + j) = v.velArr[j + 3 * (size_t)(nodeS - 1)] + (-xyzAcc[j] + xyzR[j] / massMaster) * dt;
    }
  }
  // showSourceDynamics, faulting.f90:343-365 (recorded instead of printed)
  if (std::fabs(xs[0] - P.xsource) < P.tol && std::fabs(xs[2] - P.zsource) < P.tol && v.hypoLog) {
    double* h = &v.hypoLog[13 * (size_t)(c.nt - 1)];
    h[0] = c.timeElapsed; h[1] = F(78); h[2] = F(79); h[3] = F(80); h[4] = F(23); h[5] = F(73); h[6] = F(71);
    h[7] = F(72); h[8] = F(74); h[9] = F(75); h[10] = F(76); h[11] = F(77); h[12] = F(20);
  }
  // storeOnFaultStationQuantSCEC, faulting.f90:518-541
  for (int j = 0; j < v.nOn; ++j) {
    if (v.anonfs[0 + 3 * (size_t)j] == i && v.anonfs[2 + 3 * (size_t)j] == ift + 1) {
      double* q = &v.onFaultQuantHistSCECForm[12 * ((size_t)(c.nt - 1) + (size_t)v.nstep * j)];
      q[0] = c.timeElapsed; q[1] = rate[1]; q[2] = rate[2]; q[3] = F(20);
      q[4] = slip[1]; q[5] = slip[2]; q[6] = slip[0];
      q[7] = T[1]; q[8] = T[2]; q[9] = T[0];
      q[10] = F(51) + F(42); q[11] = F(52);
    }
  }
  // storeRuptureTime, faulting.f90:329-341
  if (v.fnft[pb] > 5000.0)
    if (rate[3] >= P.slipRateThres) v.fnft[pb] = c.timeElapsed;
}

void faulting(Ctx& c) {
  eqh_view& v = *c.v;
  double garbage = 0.0;
  for (int ift = 0; ift < v.ntotft; ++ift)
    for (int i = 1; i <= v.nftnd[ift]; ++i) fault_pair(c, ift, i, &garbage);
}

// driver.f90:30-33 -> output_gm, output_src_evol (library_output.f90:267-279,297-312): what the
// reference appends to gm<me> / src_evol<me> at every step with mod(nt,10) == 1, kept in the
// host's sample arrays (the host writes the files, eqh_write_outputs)
void sample_gm(Ctx& c) {
  eqh_view& v = *c.v;
  if (v.params.outputGroundMotion != 1 || c.nt % 10 != 1 || !v.nGmSamples) return;
  const int k = *v.nGmSamples;
  if (k >= v.nGmAlloc) return;
  if (v.gmHist)
    for (int i = 0; i < v.nSurf; ++i) {
      const int node = v.surfaceNodeIdArr[i];
      for (int j = 0; j < 3; ++j) v.gmHist[j + 3 * ((size_t)i + (size_t)v.nSurf * k)] = v.velArr[j + 3 * (size_t)(node - 1)];
    }
  if (v.srcEvolHist && v.nftnd[0] > 0)
    for (int i = 0; i < v.nftnd[0]; ++i) v.srcEvolHist[(size_t)i + (size_t)v.nftnd[0] * k] = v.fric[46 + 100 * (size_t)i];
  *v.nGmSamples = k + 1;
}

}  // namespace

extern "C" {

// Run steps nt_begin..nt_end (1-based, inclusive) of driver.f90:9-34 for all
// sub-domains in `views`.  time_elapsed is the reference's accumulated
// timeElapsed (driver.f90:11), in/out.  Returns 0 or an EQD_ERR_* code.
int orc_run(eqh_view* views, int nranks, int nt_begin, int nt_end, double* time_elapsed) {
  std::vector<Ctx> ctx(nranks);
  for (int r = 0; r < nranks; ++r) ctx[r] = Ctx{&views[r], 0, *time_elapsed};
  int rc_all = 0;
  for (int nt = nt_begin; nt <= nt_end; ++nt) {
#pragma omp parallel for schedule(static)
    for (int r = 0; r < nranks; ++r) {
      Ctx& c = ctx[r];
      c.nt = nt;
      c.timeElapsed = c.timeElapsed + c.v->params.dt;
      int rc = vel_disp_update(c);
      if (rc) {
#pragma omp critical
        rc_all = rc;
      }
      store_off_fault(c);
      std::memset(c.v->nodalForceArr, 0, sizeof(double) * (size_t)c.v->Neq);
      assemble_global_ku(c);
      hrglss(c);
    }
    if (rc_all) return rc_all;
    int rc = exchange_forces(views, nranks);
    if (rc) return rc;
#pragma omp parallel for schedule(static)
    for (int r = 0; r < nranks; ++r) {
      Ctx& c = ctx[r];
      if (c.v->params.friclaw == 5) thermop(c);
      faulting(c);
      eqh_view& v = *c.v;
      for (int k = 0; k < v.Neq; ++k) v.nodalForceArr[k] = v.nodalForceArr[k] / v.nodalMassArr[k];
      sample_gm(c);
    }
  }
  *time_elapsed = ctx[0].timeElapsed;
  return 0;
}

// ---- phase-split entry points for a one-sub-domain-per-process host (tests of
// the N>1 host logic over torch.distributed/gloo): the same step, with the
// MPI4NodalQuant exchange left to the caller.
int orc_step_pre(eqh_view* v, int nt, double* time_elapsed) {
  Ctx c{v, nt, *time_elapsed + v->params.dt};
  int rc = vel_disp_update(c);
  if (rc) return rc;
  store_off_fault(c);
  std::memset(v->nodalForceArr, 0, sizeof(double) * (size_t)v->Neq);
  assemble_global_ku(c);
  hrglss(c);
  *time_elapsed = c.timeElapsed;
  return 0;
}

// number of doubles on face (axis a, side 0/1); buf != NULL: pack them in
// processNodalQuantArr order (assembleGlobalMass.f90:141-188)
int orc_face_pack(eqh_view* v, int a, int side, double* buf) {
  std::vector<double*> slots;
  face_slots(*v, a, side, slots);
  if (buf)
    for (size_t k = 0; k < slots.size(); ++k) buf[k] = *slots[k];
  return (int)slots.size();
}

int orc_face_add(eqh_view* v, int a, int side, const double* buf, int n) {
  std::vector<double*> slots;
  face_slots(*v, a, side, slots);
  if ((int)slots.size() != n) return EQD_ERR_ARG;
  for (int k = 0; k < n; ++k) *slots[k] = *slots[k] + buf[k];
  return 0;
}

int orc_step_post(eqh_view* v, int nt, double time_elapsed) {
  Ctx c{v, nt, time_elapsed};
  if (v->params.friclaw == 5) thermop(c);
  faulting(c);
  for (int k = 0; k < v->Neq; ++k) v->nodalForceArr[k] = v->nodalForceArr[k] / v->nodalMassArr[k];
  sample_gm(c);
  return 0;
}

}  // extern "C"
