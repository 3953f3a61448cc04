/*
 * eqdyna_host.h -- C API of libeqdyna_host.so, the stand-in for EQdyna's
 * Fortran host (the build container has no Fortran compiler).  It does what
 * src/eqdyna3d.f90:33-70 does before `call driver` -- read the case files,
 * size and generate the mesh, precompute mass and element operators, load the
 * on-fault fields -- and what eqdyna3d.f90:75-79 does after it (text outputs).
 * It owns the `globalvar` arrays of each sub-domain and exposes them as raw
 * Fortran-layout pointers (`eqh_view`), which the caller hands to the step
 * library (include/eqdyna_b200.h) exactly as the Fortran host would.
 * No CUDA dependency.
 */
#ifndef EQDYNA_HOST_H
#define EQDYNA_HOST_H

#include <stdint.h>

#include "eqdyna_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct eqh_world eqh_world;

/* Pointers into one sub-domain's state; names are the globalvar names. */
typedef struct eqh_view {
  eqd_params params;
  int32_t Nn, Ne, Neq, sizeEq, sizeStress; /* sizeStress = 5*sizeEq as allocated */
  int32_t nx, ny, nz;                      /* local node grid (numcount(1:3))    */
  int32_t nftmx, ntotft, nOn, nOnAlloc, nOff, nSurf, nstep;
  int32_t stressUsed;                      /* sizeOfStressDofIndexArr            */
  int32_t pad_;
  double* meshCoor;
  int32_t* nodeElemIdRelation;
  int32_t* elemTypeArr;
  int32_t* numOfDofPerNodeArr;
  int32_t* eqNumStartIndexLoc;
  int32_t* eqNumIndexArr;
  int32_t* stressCompIndexArr;
  double* eleshp;
  double* eledet;
  double* elemass;
  double* mat;
  double* ss;
  double* phi;
  double* eleporep;
  double* stressArr;
  double* pstrain;
  double* nodalMassArr;
  double* fnms;
  double* v1;
  double* velArr;
  double* dispArr;
  double* nodalForceArr;
  int32_t* nftnd;
  int32_t* nsmp;
  double* un;
  double* us;
  double* ud;
  double* arn;
  double* fric;
  double* fnft;
  int32_t* numcount; /* (9) */
  int32_t* fltnum;   /* (6) */
  int32_t* fltMPI;   /* (6) */
  int32_t* fltface[6]; /* fltl, fltr, fltf, fltb, fltd, fltu */
  int32_t* idhist;
  int32_t* anonfs;
  int32_t* surfaceNodeIdArr;
  double* onFaultQuantHistSCECForm;
  double* OffFaultStGramSCEC;
  double* hypoLog;
  double* onFaultTPHist;
  /* samples of output_gm / output_src_evol (driver.f90:30-33): velArr of the surface
   * nodes and fric(47,:,1) at every step with mod(nt,10) == 1, in file order      */
  double* gmHist;        /* (3,nSurf,nGmAlloc)     */
  double* srcEvolHist;   /* (nftnd(1),nGmAlloc)    */
  int32_t* nGmSamples;   /* samples taken so far   */
  int32_t nGmAlloc;      /* nstep/10 + 1, 0 = outputGroundMotion off */
  int32_t pad2_;
} eqh_view;

/* Read the case directory (b*.txt + on_fault_vars_input.bin [+ rough geometry]).
 * npx/npy/npz > 0 override the decomposition of bGlobal.txt; nstep > 0 overrides
 * the number of time steps (term/dt).                                          */
int eqh_world_create(const char* case_dir, int npx, int npy, int npz, int nstep, eqh_world** out);
int eqh_world_destroy(eqh_world* w);
/* emulate a host built with other compile-time switches of globalvar.f90:
 * "C_Q", "C_hg", "kapa_hg", "rdampm", "outputGroundMotion"                    */
int eqh_world_set_switch(eqh_world* w, const char* name, double value);
const char* eqh_last_error(void);
int eqh_world_size(const eqh_world* w);
/* Build sub-domain `rank` (mesh4num, meshgen, on-fault load, mass/operators,
 * init_vel), rank = -1 builds all of them.  With all ranks in-process call
 * eqh_world_sum_shared afterwards (the MPI sums of arn, nodalMassArr, fnms);
 * a one-rank-per-process host calls eqd_sum_shared on the step library instead. */
int eqh_world_build(eqh_world* w, int rank);
int eqh_world_sum_shared(eqh_world* w);
int eqh_get_view(eqh_world* w, int rank, eqh_view* out);
/* What eqdyna3d.f90:75-84 and driver.f90:30-33 write for one rank, in the reference's
 * formats (library_output.f90): faultst*.txt, body*.txt, frt.txt<me>; with
 * outputGroundMotion gm<me>, src_evol<me> (raw float64 streams) and surface_coor.txt<me>;
 * with outputFinalSurfDisp finalSurfDisp.txt<me>; with output_plastic pstr.txt<me>;
 * compTime<me> once eqh_set_comp_time supplied the ten timings.                      */
int eqh_write_outputs(eqh_world* w, int rank, const char* out_dir);
/* compTimeInSeconds(1:9) and MPICommTimeInSeconds of output_timeanalysis (library_output.f90:208-218) */
int eqh_set_comp_time(eqh_world* w, int rank, const double* t10);
/* free the big element-operator arrays of a rank once they live on the device */
int eqh_release_operators(eqh_world* w, int rank);

#ifdef __cplusplus
}
#endif
#endif
