// Device-side data model of the B200 step library (libeqdyna_b200.so).
//
// The Fortran host hands over module `globalvar` arrays (AoS, 1-based,
// equation-indirected: src/globalvar.f90:81-101).  On the device the same state
// lives as 0-based structure-of-arrays laid out for coalesced FP64 streaming:
//
//   nodes     vel[3][NnS] disp[3][NnS] mass[Nn]            (3-dof nodes: v1 == vel)
//   PML nodes v1p[12][NpS] dampp[3][NpS]                   (12-dof split field, comdampv precomputed)
//   elements  three classes, each SoA with the element index fastest:
//             REG  (types 1,11,12,13 touching only 3-dof nodes)  ef[24]
//             REGX (regular element touching a 12-dof node)      ef[48] = KU | hourglass
//             PML  (type 2)                                      ef[96]
//   assembly  node -> (element, local node) CSR in ascending reference element
//             id; a node's force is the ordered sum of its elements' `ef`
//             entries (no atomics => fixed summation order)
//   special   nodes whose force is needed between the element sweep and the
//             next node update (split-node pairs, rank-face halo nodes): their
//             sums are materialised in force[3][NnS] / forcep[12][NpS]
#pragma once
#include <cstdint>

namespace eqd {

enum { KIND_FREE3 = 0, KIND_PML12 = 1, KIND_FIXED = 2 };
enum { CLS_REG = 0, CLS_REGX = 1, CLS_PML = 2 };
#define EQD_INFO_KIND(i) ((i) & 3)
#define EQD_INFO_SPECIAL(i) (((i) >> 2) & 1)
#define EQD_INFO_SLOT(i) ((i) >> 3)
// adjacency entry: class (2 bits) | local node (3 bits) | element index in class (27 bits)
#define EQD_ADJ(cls, ln, idx) ((uint32_t)(cls) | ((uint32_t)(ln) << 2) | ((uint32_t)(idx) << 5))

struct StepState {
  double timeElapsed;  // driver.f90:11
  int nt;              // current step, 1-based
  int nanFlag;         // driver.f90:147-152
  int nanNode;         // 1-based node id of the first NaN seen
  int pad;
};

struct NodeArgs {
  int Nn, NnS, Np, NpS;
  const int* info;       // [Nn] kind | special<<2 | pmlSlot<<3
  double* vel;           // [3][NnS]
  double* disp;          // [3][NnS]
  const double* mass;    // [Nn]
  double* v1p;           // [12][NpS]
  const double* dampp;   // [3][NpS]
  double* force;         // [3][NnS] + [12][NpS] (forcep = force + 3*NnS)
  const int* adjStart;   // [Nn+1]
  const uint32_t* adj;
  const double* efR; int SR;
  const double* efX; int SX;
  const double* efP; int SP;
  const double* accel0;  // optional uploaded acceleration, same layout as force; first step only
  double dt;
  StepState* st;
};

struct ElemArgs {
  int n, S;
  const int* conn;       // [8][S] 0-based node ids
  const double* shp;     // [24][S]  eleshp(3,8,e): row 3*i+j
  const double* phi;     // [32][S]  phi(8,4,e):    row 8*m+i
  const double* ss;      // [6][S]
  const double* lam; const double* mu; const double* det;
  const double* rho; const double* vp;  // viscous hourglass only (C_hg==2)
  double* stress;        // REG/REGX [6][S]; PML [21][S]
  double* qmem;          // [6][S]  memory variables (C_Q==1)
  const uint8_t* qcls;   // [S]     Q class 0..15 -> c_qtab
  const double* porep; double* pstrain;  // plastic
  const double* emass;   // [8][S] nodal lumped element mass (body force / Rayleigh mass damping)
  const double* damps;   // [3][S] PML damping profile at the centroid
  double* ef;            // [24|48|96][S]
  const double* vel; const double* disp; int NnS;
  double dt, rdampk, rdampm, w, bodyz, ccosphi, sinphi, expdttv, kapa_hg;
};

struct FaultArgs {
  int nPairs, PS, NnS;
  const int* nodeS; const int* nodeM;  // 0-based node ids
  const int* ift;                      // 1-based fault id per pair
  const double* un; const double* us; const double* ud;  // [3][PS]
  const double* arn; const double* massS; const double* massM;
  const double* xs;                    // [3][PS] slave node coordinates
  double* fric;                        // [100][PS]
  double* fnft;                        // [PS]
  const double* vel; const double* disp; double* force;
  const int* pairStation;              // [PS] -1 or on-fault station index
  double* onHist; int nstep;           // (12,nstep,nOnAlloc)
  double* hypoLog;                     // (13,nstep)
  double* tphist;                      // [nstep][2][PS]
  double* srcEvol; int nSrc;           // not used by the kernel
  StepState* st;
  double dt, nucR, nucT, nucRuptVel, nucdtau0, xsource, ysource, zsource, slipRateThres, tol, fric_tp_h;
  int friclaw, C_nuclea, nucfault, TPV, insertFaultType, C_elastic;
};

// Q constants per class (qconstant.f90), filled by the host at create time
struct QTab { double taok, wkp, wks, cv, cs, expdt; };

}  // namespace eqd
