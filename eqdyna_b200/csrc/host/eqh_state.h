// Host-side state of one sub-domain (= one MPI rank of the reference).
//
// This is the C++ stand-in for EQdyna's Fortran host (module globalvar,
// src/globalvar.f90:81-101).  Arrays keep the FORTRAN layout (column-major)
// and integer arrays keep 1-based node / element / equation ids, so that the
// pointers can be handed to the C ABI (include/eqdyna_b200.h) exactly as a
// Fortran host would hand `c_loc(array)`.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "eqdyna_b200.h"

namespace eqh {

// Everything read from the b*.txt files (src/readInputFiles.f90) plus the
// compile-time initialisers of src/globalvar.f90 that the code reads.
struct CaseInput {
  // bGlobal.txt (readInputFiles.f90:28-58)
  int mode = 1, C_elastic = 1, C_nuclea = 1, insertFaultType = 0, friclaw = 1,
      ntotft = 1, nucfault = 1, TPV = -1, output_plastic = 0,
      outputGroundMotion = 0, outputFinalSurfDisp = 0;
  double C_degen = 0.0;
  int npx = 1, npy = 1, npz = 1;
  double totalSimuTime = 0, dt = 0;
  int nmat = 1, n2mat = 3;
  double roumax = 0, rhow = 0, gamar = 0, rdampk = 0, vmaxPML = 0;
  double xsource = 0, ysource = 0, zsource = 0;
  double nucR = 0, nucRuptVel = 0, nucdtau0 = 0, nucT = 0;
  double str1ToFaultAngle = 0, devStrToStrVertRatio = 0, bulk = 0, coheplas = 0;
  double fstrike = 0, fdip = 0, slipRateThres = 0;
  // bModelGeometry.txt (readInputFiles.f90:87-95)
  double xmin = 0, xmax = 0, ymin = 0, ymax = 0, zmin = 0, zmax = 0;
  int dis4uniF = 0, dis4uniB = 0;
  double rat = 1, dx = 0, dy = 0, dz = 0;
  // bFaultGeometry.txt (readInputFiles.f90:124-146)
  std::vector<double> fxmin, fxmax, fymin, fymax, fzmin, fzmax;
  std::vector<double> fltxyz;  // (2,4,ntotft)
  // bMaterial.txt
  std::vector<double> material;  // (nmat,n2mat) column-major
  // bStations.txt
  int totalNumOfOffSt = 0;
  std::vector<int> nonfs;        // (ntotft)
  std::vector<double> xonfs;     // (2,max(nonfs),ntotft), metres
  std::vector<double> x4nds;     // (3,totalNumOfOffSt), metres
  // bFault_Rough_Geometry.txt
  int nnx = 0, nnz = 0;
  double dxtmp = 0, rough_fx_min = 0, rough_fx_max = 0, rough_fz_min = 0;
  std::vector<double> rough_geo;  // (3,nnx*nnz)
  // on_fault_vars_input (24 fields, each (fnx,fnz) column-major = ix fastest)
  int fnx = 0, fnz = 0;
  std::vector<double> on_fault_vars;  // (fnx,fnz,24)
  // mode == 2: the 12 fields of the restart file fault.r.nc spun off by EQquasi
  // (netcdf_io.f90:116-185), from the raw dump fault.r.bin
  std::vector<double> restart_vars;   // (fnx,fnz,12)
  // derived (readInputFiles.f90:182-186)
  double ccosphi = 0, sinphi = 0, tv = 0;
  int nstep = 0;
  // globalvar.f90 initialisers
  double pi = 0, tol = 1.0e-5, R = 0.01, kapa_hg = 0.1, grav = 9.8,
         rdampm = 0.0, w = 8.0, fric_tp_h = 0.0;
  int nPML = 6, C_Q = 0, C_hg = 1;
  std::string dir;
};

struct RankState {
  int me = 0, mex = 0, mey = 0, mez = 0;
  // sizes (mesh4num.f90:84-87)
  int totalNumOfNodes = 0, totalNumOfElements = 0, totalNumOfEquations = 0,
      sizeOfEqNumIndexArr = 0, sizeOfStressDofIndexArr = 0;
  int nx = 0, ny = 0, nz = 0;  // local node grid
  int nftmx = 1, nonmx = 0;
  double PMLb[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  double xminB = 0, xmaxB = 0, yminB = 0, ymaxB = 0, zminB = 0, zmaxB = 0;  // modelBoundCoor
  int numcount[9] = {0};
  int fltnum[6] = {0};
  int fltMPI[6] = {0};
  std::vector<int> fltface[6];  // fltl, fltr, fltf, fltb, fltd, fltu
  std::vector<int> nftnd;       // (ntotft)
  // mesh
  std::vector<double> meshCoor;             // (3,Nn)
  std::vector<int> nodeElemIdRelation;      // (8,Ne)
  std::vector<int> elemTypeArr;             // (Ne)
  std::vector<int> numOfDofPerNodeArr;      // (Nn)
  std::vector<int> eqNumStartIndexLoc;      // (Nn)
  std::vector<int> eqNumIndexArr;           // (sizeOfEqNumIndexArr)
  std::vector<int> stressCompIndexArr;      // (Ne)
  std::vector<int> surfaceNodeIdArr;        // (Nn), first surface_nnode used
  int surface_nnode = 0;
  // element data
  std::vector<double> mat;       // (Ne,5) element index fastest
  std::vector<double> eleporep, pstrain, eledet;  // (Ne)
  std::vector<double> elemass;   // (24,Ne)
  std::vector<double> eleshp;    // (3,8,Ne)
  std::vector<double> ss;        // (6,Ne)
  std::vector<double> phi;       // (8,4,Ne)
  std::vector<double> stressArr; // (5*sizeOfEqNumIndexArr)
  // nodal
  std::vector<double> fnms;      // (Nn)
  std::vector<double> nodalForceArr, nodalMassArr, v1;  // (Neq)
  std::vector<double> velArr, dispArr;                  // (3,Nn)
  // fault
  std::vector<int> nsmp;         // (2,nftmx,ntotft)
  std::vector<int> fltgm;        // (nftmx)
  std::vector<double> fnft, arn; // (nftmx,ntotft)
  std::vector<double> un, us, ud;  // (3,nftmx,ntotft)
  std::vector<double> fric;      // (100,nftmx,ntotft)
  std::vector<int> anonfs;       // (3,nonmx)
  int numOfOnFaultStCount = 0, numOfOffFaultStCount = 0;
  int nOnAlloc = 1;              // allocated station slots (>=1, eqdyna3d.f90:157)
  std::vector<int> OffFaultStNodeIdIndex;  // (2,totalNumOfOffSt)
  std::vector<int> idhist;       // (3,6*nOff)
  std::vector<double> onFaultQuantHistSCECForm;  // (12,nstep,nOnAlloc)
  std::vector<double> OffFaultStGramSCEC;        // (6*nOff+1,nstep)
  std::vector<double> hypoLog;   // (13,nstep)
  std::vector<double> onFaultTPHist;  // (2,nftmx,nstep,ntotft) only friclaw==5
  // what output_gm / output_src_evol append to gm<me> / src_evol<me> every 10th step
  // (driver.f90:30-33), kept in memory until eqh_write_outputs
  std::vector<double> gmHist;       // (3,surface_nnode,nGmAlloc)
  std::vector<double> srcEvolHist;  // (nftnd(1),nGmAlloc)
  int nGmAlloc = 0, nGmSamples = 0;
  double compTime[10] = {0};        // compTimeInSeconds(1:9), MPICommTimeInSeconds (library_output.f90:208-218)
  bool compTimeSet = false;
  eqd_params params() const;
  const CaseInput* in = nullptr;
};

// eqh_io.cpp
void read_case(const std::string& dir, CaseInput& in);
void write_frt(const RankState& s, const std::string& dir);
void write_onfault_stations(const RankState& s, const std::string& dir);
void write_offfault_stations(const RankState& s, const std::string& dir);
void write_surface_coor(const RankState& s, const std::string& dir);
void write_gm(const RankState& s, const std::string& dir);
void write_src_evol(const RankState& s, const std::string& dir);
void write_final_surf_disp(const RankState& s, const std::string& dir);
void write_plastic_strain(const RankState& s, const std::string& dir);
void write_comp_time(const RankState& s, const std::string& dir);
// eqh_mesh.cpp
void mesh4num(const CaseInput& in, RankState& s);
void meshgen(const CaseInput& in, RankState& s);
void exchange_arn(const CaseInput& in, std::vector<RankState*>& world);
void load_on_fault(const CaseInput& in, RankState& s);
void load_on_fault_restart(const CaseInput& in, RankState& s);
void find_surface_nodes(const CaseInput& in, RankState& s);
void alloc_after_meshgen(const CaseInput& in, RankState& s);
// eqh_mass.cpp
void assemble_global_mass(const CaseInput& in, RankState& s);
void exchange_nodal(const CaseInput& in, std::vector<RankState*>& world, int which, int numDof);
void init_vel(const CaseInput& in, RankState& s);

}  // namespace eqh
