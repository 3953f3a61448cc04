// Device-side data model of the B200 step library (libeqdyna_b200.so).
//
// The Fortran host hands over module `globalvar` arrays (AoS, 1-based,
// equation-indirected: src/globalvar.f90:81-101).  On the device the same state
// lives as 0-based structure-of-arrays laid out for coalesced FP64 streaming:
//
//   nodes     vel[3][NnS] disp[3][NnS] mass[Nn]            (3-dof nodes: v1 == vel)
//   PML nodes v1p[12][NpS] dampp[3][NpS]                   (12-dof split field, comdampv precomputed)
//   elements  three classes, each SoA with the element index fastest:
//             REG  (types 1,11,12,13 touching only 3-dof nodes)  3 force rows / node
//             REGX (regular element touching a 12-dof node)      6 rows = KU | hourglass
//             PML  (type 2)                                      12 rows
//   tiles     every class is cut into bricks of ~4x4x32 (PML 3x3x32) elements
//             stored contiguously; one CTA owns one tile: it stages the tile's
//             nodes (v, d + rdampk v) in shared memory, sweeps the tile's
//             elements and accumulates their nodal forces in shared memory in a
//             FIXED order (8 local-node phases, host-coloured where two elements
//             of one phase share a node: no atomics), then writes ONE partial
//             force row per tile node: pf[rows][slot]
//   assembly  node -> tile-node slots, stored by rank (class, tile ascending); a node's
//             force is the ordered sum of its 1..8 tile partials, fused into
//             the next node update
//   special   nodes whose force is needed between the element sweep and the
//             next node update (split-node pairs, rank-face halo nodes): their
//             sums are materialised in force[3][NnS] / forcep[12][NpS]
#pragma once
#include <cstdint>

#include <vector_types.h>

namespace eqd {

enum { KIND_FREE3 = 0, KIND_PML12 = 1, KIND_FIXED = 2 };
// CLS_MARCH: bundles of box elements (eqd_march.h); CLS_MARCHP: bundles of PML box elements (eqd_march_pml.h), whose
// partial force rows live behind the PML tile class's in the same buffer (slot entries carry CLS_PML)
enum { CLS_REG = 0, CLS_REGX = 1, CLS_PML = 2, CLS_MARCH = 3, CLS_MARCHP = 4, NCLS = 5 };
#define EQD_INFO_KIND(i) ((i) & 3)
#define EQD_INFO_SPECIAL(i) (((i) >> 2) & 1)
#define EQD_INFO_SLOT(i) ((i) >> 3)
// free 3-dof nodes only (they have no PML slot): the node is updated by its marching bundle (eqd_march.h)
#define EQD_INFO_FUSED_BIT 8
#define EQD_INFO_FUSED(i) (((i) >> 3) & 1)
// node -> tile-node slot entry: class (2 bits) | slot in the class's partial buffer (30 bits)
#define EQD_SLOT(cls, slot) ((uint32_t)(cls) | ((uint32_t)(slot) << 2))
// local connectivity entry: tile-local node index (12 bits) | colour of the phase (4 bits)
#define EQD_LN_BITS 12
#define EQD_LN_MASK 0x0fffu
#define EQD_STAGE 128          // regular tile kernel: elements per streamed stage (two threads each)
#define EQD_STAGE_PML 64       // PML tile kernel: elements per streamed stage (four threads each)
#define EQD_PML_LS 400         // PML tile kernel: node cap of a tile
#define EQD_REG_LS 400         // regular tile kernel: node cap of a tile = shared-memory row stride (two CTAs per SM)

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
  const int* pmlNode;    // [Np] PML slot -> node
  const double* vel;     // [3][NnS] v(nt-1), d(nt-1) as read ...
  const double* disp;
  double* velOut;        // ... and where v(nt), d(nt) go: the same arrays, or the other buffer of the pair (eqd_api.cu)
  double* dispOut;
  const double* mass;    // [Nn]
  double* v1p;           // [12][NpS]
  const double* dampp;   // [3][NpS]
  double* force;         // [3][NnS] + [12][NpS] (forcep = force + 3*NnS)
  const uint8_t* slotCnt;   // [Nn] tile-node slots of the node (0 for fixed nodes)
  const uint32_t* slotTab;  // [maxCnt][NnS] EQD_SLOT entries by rank (class, then ascending tile id)
  const double* pfR; int SR;   // [3][SR]  tile partials, REG
  const double* pfX; int SX;   // [6][SX]  REGX: KU | hourglass
  const double* pfP; int SP;   // [12][SP] PML
  const double* pfM; int SM;   // [3][SM]  marching bundles (non-fused node slots)
  const int* list; int nList;  // list variant of k_node_update3: the free 3-dof nodes no bundle updates, ascending
  int fusedMode;               // nodes flagged EQD_INFO_FUSED: 1 = take their force from force[] (the last sweep did not
                               // update them), 2 = skip them (it did)
  const double* accel0;  // optional uploaded acceleration, same layout as force; first step only
  int variant;           // launch-bounds variant of k_node_update3 (tuning)
  int skipSpecial;       // 1: k_node_update3/12 leave the special nodes to k_node_update_special
  double dt;
  StepState* st;
};

struct ElemArgs {
  int n, S;
  // tiles: elements of tile t are slots tileElem[t] .. tileElem[t]+tileCnt[t]-1,
  // its nodes are tnode[tileNode[t] .. tileNode[t+1]-1] (-1 = padding)
  const int4* tileRec;   // per tile in LAUNCH order (rank-face tiles first):
                         // {first slot, elements | colours << 16, first tile-node slot, tile nodes}
  int tile0, ntiles;     // tiles [tile0, tile0+ntiles) of tileRec belong to this launch
  int maxGrid;           // cap on the persistent grid (0 = two CTAs per SM); the overlapped interior sweep leaves room for the halo kernels
  const int* tnode;
  const uint16_t* lconn; // [8][S] local node | colour << 12
  int LS;                // shared-memory row stride (max tile nodes of the class)
  double* pf; int PFS;   // [3|6|12][PFS] partial nodal forces, one row per tile node
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
  const double* vel; const double* disp; int NnS;
  double dt, rdampk, rdampm, w, bodyz, ccosphi, sinphi, expdttv, kapa_hg;
  const uint8_t* tileBox;  // per tile in launch order: 1 = every element is an axis-aligned hexahedron (eqd_box.h); null = off
  int allBox;              // 1: every tile of the class is a box tile and the compact stage buffer was asked for ("box_compact")
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

// one axis phase of the nodal-force exchange over peer memory (k_halo_send / k_halo_recv), [side]
struct HaloAxisArgs {
  int n[2];                    // doubles exchanged with the -side / +side neighbour (0 = none)
  const uint32_t* idx[2];      // dof positions in force[]
  double* remoteRecv[2];       // the neighbour's receive buffer for this exchange (its memory, mapped through CUDA IPC)
  unsigned* remoteFlag[2];     // ... and its flag word
  const double* localRecv[2];  // this sub-domain's receive buffer for this exchange
  const unsigned* localFlag[2];
  unsigned* counter[2];        // blocks of k_halo_send that have finished
  double* force;
  unsigned seq;                // exchange number (1, 2, ...)
};

// operator precompute on the device (eqd_ops.cu)
struct OpsArgs {
  int S, Ne;
  const int* refId;      // [S] class slot -> reference element (0-based), -1 = padding
  const int* conn;       // (8,Ne) 0-based
  const int* etype;      // [Ne]
  const double* coor;    // (3,Nn)
  const double* mat;     // (Ne,5), element index fastest
  double w;
  double* shp; double* phi; double* ss; double* det; double* lam; double* mu;
  double* rho; double* vp;   // optional (C_hg == 2)
  double* em;            // [8][S] lumped element mass per local node
  int* badElem;          // smallest element id with a non-positive determinant
  int compact;           // 1 (marching class): shp = [3][S] a_x a_y a_z (eleshp rows 3, 7, 14), ss = [3][S] ss1 ss4 ss6, no phi
};
struct TileMassArgs {
  const int4* tileRec; const double* em; const uint16_t* lconn; int S, capE;
  double* pm;            // [PFS] lumped-mass partial per tile-node slot
};
struct NodeMassArgs {
  int Nn, NnS;
  const uint8_t* slotCnt; const uint32_t* slotTab;
  const double* pm[NCLS];
  double* mass;
};

// Q constants per class (qconstant.f90), filled by the host at create time
struct QTab { double taok, wkp, wks, cv, cs, expdt; };

}  // namespace eqd
