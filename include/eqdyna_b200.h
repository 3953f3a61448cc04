/*
 * eqdyna_b200.h -- C ABI of libeqdyna_b200.so, the B200 (sm_100a) step library
 * that replaces the body of EQdyna's `subroutine driver`.
 *
 * Reference boundary: src/driver.f90:3-36 (the step loop), called from
 * src/eqdyna3d.f90:72.  Everything the loop reads is module `globalvar`
 * (src/globalvar.f90:81-101, shapes from src/eqdyna3d.f90:98-172); every
 * eqd_set_* argument below is one of those arrays, passed AS ALLOCATED BY THE
 * FORTRAN HOST: column-major, 1-based indices inside integer arrays.  The
 * library copies at eqd_set_*; the host keeps ownership.  Results come back
 * only through eqd_fetch into caller-owned buffers of the original shape.
 *
 * All entry points return 0 on success or one of the EQD_ERR_* codes; no C++
 * exception crosses this boundary.  The matching ISO_C_BINDING interface is
 * eqdyna_b200/csrc/fortran/eqdyna_cuda_iface.f90; INTEGRATION.md shows the
 * replacement driver.
 */
#ifndef EQDYNA_B200_H
#define EQDYNA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- error codes (mirror the reference's `stop` sites) ------------------- */
#define EQD_OK 0
#define EQD_ERR_NAN 1      /* velocity NaN            (driver.f90:147-152)   */
#define EQD_ERR_DAMP 2     /* negative PML damping    (comdampv.f90:114-118) */
#define EQD_ERR_CUDA 3     /* CUDA / NCCL failure                            */
#define EQD_ERR_ARG 4      /* bad argument / wrong call order                */

/* ---- scalars of module globalvar the step loop reads --------------------- */
typedef struct eqd_params {
  /* time stepping (readInputFiles.f90:182-186) */
  double dt;            /* globalvar dt                                      */
  int32_t nstep;        /* idnint(totalSimuTime/dt)                          */
  int32_t me;           /* MPI rank id = mex*npy*npz + mey*npz + mez         */
  int32_t npx, npy, npz;
  /* damping / element constants */
  double rdampk;        /* ALREADY multiplied by dt (readInputFiles.f90:185) */
  double rdampm;        /* 0 in the reference (globalvar.f90:16)             */
  double w;             /* 8.0, one-point Gauss weight                       */
  double grav, roumax, rhow, gamar;
  double ccosphi, sinphi, tv; /* Drucker-Prager (readInputFiles.f90:182-186) */
  double kapa_hg;       /* viscous hourglass coefficient (C_hg==2)           */
  double dx;            /* used by the Q class index only                    */
  int32_t C_elastic, C_Q, C_hg;
  /* PML */
  double PMLb[8];       /* meshgen.f90:532-549                               */
  int32_t nPML;
  double R;             /* theoretical reflection coefficient                */
  double vmaxPML;
  /* fault */
  int32_t friclaw, C_nuclea, nucfault, TPV, insertFaultType, ntotft;
  double nucR, nucT, nucRuptVel, nucdtau0;
  double xsource, ysource, zsource;
  double slipRateThres;
  double tol;           /* 1e-5, hypocentre-pair match (faulting.f90:349)    */
  double fric_tp_h;     /* global never assigned in the reference => 0       */
  int32_t outputGroundMotion;
  int32_t reserved_i[7];
  double reserved_d[8];
} eqd_params;

typedef struct eqd_handle eqd_handle;

/* which-codes of eqd_fetch: destination must have the Fortran shape given.  */
enum {
  EQD_F_DISP = 1,        /* dispArr(3,Nn)                                    */
  EQD_F_VEL = 2,         /* velArr(3,Nn)                                     */
  EQD_F_V1 = 3,          /* v1(Neq)                                          */
  EQD_F_FORCE = 4,       /* nodalForceArr(Neq)  (= f/m after the last step)  */
  EQD_F_FRIC = 5,        /* fric(100,nftmx,ntotft)                           */
  EQD_F_FNFT = 6,        /* fnft(nftmx,ntotft)                               */
  EQD_F_PSTRAIN = 7,     /* pstrain(Ne)                                      */
  EQD_F_STRESS = 8,      /* stressArr(sizeStress)                            */
  EQD_F_ONFAULT_HIST = 9,  /* onFaultQuantHistSCECForm(12,nstep,nOn)         */
  EQD_F_OFFFAULT_HIST = 10,/* OffFaultStGramSCEC(6*nOff+1,nstep)             */
  EQD_F_HYPO_LOG = 11,   /* (13,nstep): what showSourceDynamics prints       */
  EQD_F_GM = 12,         /* (3,nSurf,nGmSamples) output_gm samples           */
  EQD_F_SRC_EVOL = 13,   /* (nftnd(1),nGmSamples) output_src_evol samples    */
  EQD_F_TPHIST = 14,     /* onFaultTPHist(2,nftmx,nstep,ntotft), friclaw 5   */
  EQD_F_MASS = 15,       /* nodalMassArr(Neq) (after eqd_sum_shared)         */
  EQD_F_FNMS = 16,       /* fnms(Nn); after eqd_compute_elem_ops: 0 on nodes whose
                          * dofs are all fixed (the loop never reads it there; the
                          * reference accumulates it, assembleGlobalMass.f90:322)      */
  EQD_F_ARN = 17,        /* arn(nftmx,ntotft)                                */
  EQD_F_ELEDET = 18,     /* eledet(Ne)       | element operators as the device  */
  EQD_F_ELESHP = 19,     /* eleshp(3,8,Ne)   | holds them (uploaded by          */
  EQD_F_SS = 20,         /* ss(6,Ne)         | eqd_set_elem_ops or computed by  */
  EQD_F_PHI = 21         /* phi(8,4,Ne)      | eqd_compute_elem_ops)            */
};

/* timing slots of eqd_get_timing (ms, CUDA events) -- the reference's
 * compTimeInSeconds(3..6)+MPICommTimeInSeconds (library_output.f90:208-218) */
enum {
  EQD_T_TOTAL = 0, EQD_T_NODE = 1, EQD_T_ELEM = 2 /* regular hex/wedge kernel */, EQD_T_ASSEMBLE = 3,
  EQD_T_HALO = 4, EQD_T_FAULT = 5, EQD_T_ELEM_PML = 6, EQD_T_ELEM_REGX = 7,
  EQD_T_MARCH = 8 /* marching kernel: box bundles, element sweep + assembly + update of their inner nodes */,
  EQD_T_MARCH_PML = 9 /* marching kernel of the PML bundles */, EQD_T_NSLOTS = 10
};

/* -- lifecycle ------------------------------------------------------------- */
int eqd_create(const eqd_params* p, int device, eqd_handle** out);
int eqd_destroy(eqd_handle* h);
int eqd_last_error(const eqd_handle* h, char* buf, int n);

/* -- state upload (each replaces reading the named globalvar arrays) ------- */
/* mesh: meshgen.f90 outputs.  nodeElemIdRelation(8,Ne), eqNumIndexArr(sizeEq) */
int eqd_set_mesh(eqd_handle* h, int32_t Nn, int32_t Ne, int32_t Neq, int32_t sizeEq,
                 const double* meshCoor, const int32_t* nodeElemIdRelation,
                 const int32_t* elemTypeArr, const int32_t* numOfDofPerNodeArr,
                 const int32_t* eqNumStartIndexLoc, const int32_t* eqNumIndexArr,
                 const int32_t* stressCompIndexArr, int32_t sizeStress);
/* element operators: assembleGlobalMass.f90:283-374 outputs.
 * eleshp(3,8,Ne) eledet(Ne) elemass(24,Ne) mat(Ne,5) ss(6,Ne) phi(8,4,Ne)    */
int eqd_set_elem_ops(eqd_handle* h, const double* eleshp, const double* eledet,
                     const double* elemass, const double* mat, const double* ss,
                     const double* phi, const double* eleporep,
                     const double* stressArr, const double* pstrain);
/* Alternative to eqd_set_elem_ops: compute eleshp, eledet, elemass, ss, phi and the
 * lumped nodal mass ON THE DEVICE from the mesh given to eqd_set_mesh, i.e. the
 * work of assembleGlobalMass (assembleGlobalMass.f90:3-56,283-406),
 * calcGlobalShapeFunc.f90:19-75 and vlm (library.f90:60-93); the host then neither
 * computes nor uploads the ~600 B of operators per element.  Per-element values are
 * bit-identical to the Fortran ones (no FMA contraction, same operation order); the
 * nodal mass is summed in tile order (last-bit differences).  Follow with
 * eqd_set_nodal(h, NULL, NULL, v1, velArr, dispArr, ...) to keep that mass, and
 * with eqd_sum_shared when sub-domains share nodes.  Error 4 with the element
 * number on a non-positive determinant (calcGlobalShapeFunc.f90:57-61).          */
int eqd_compute_elem_ops(eqd_handle* h, const double* mat, const double* eleporep,
                         const double* stressArr, const double* pstrain);
/* nodal state: nodalMassArr(Neq) fnms(Nn) v1(Neq) velArr(3,Nn) dispArr(3,Nn)
 * nodalForceArr(Neq) (acceleration left by a previous run; may be NULL = 0) */
int eqd_set_nodal(eqd_handle* h, const double* nodalMassArr, const double* fnms,
                  const double* v1, const double* velArr, const double* dispArr,
                  const double* nodalForceArr);
/* fault: nsmp(2,nftmx,ntotft) un/us/ud(3,nftmx,ntotft) arn(nftmx,ntotft)
 * fric(100,nftmx,ntotft) fnft(nftmx,ntotft)                                  */
int eqd_set_fault(eqd_handle* h, int32_t nftmx, const int32_t* nftnd,
                  const int32_t* nsmp, const double* un, const double* us,
                  const double* ud, const double* arn, const double* fric,
                  const double* fnft);
/* halo description: numcount(9) fltnum(6) fltMPI(6 as int) and the six face
 * split-node lists (meshgen.f90:209-254); NULL lists allowed when count==0   */
int eqd_set_halo(eqd_handle* h, const int32_t* numcount, const int32_t* fltnum,
                 const int32_t* fltMPI, const int32_t* fltl, const int32_t* fltr,
                 const int32_t* fltf, const int32_t* fltb, const int32_t* fltd,
                 const int32_t* fltu);
/* stations: idhist(3,6*nOff) anonfs(3,nOn) surfaceNodeIdArr(nSurf)           */
int eqd_set_stations(eqd_handle* h, const int32_t* idhist, int32_t nOff,
                     const int32_t* anonfs, int32_t nOn,
                     const int32_t* surfaceNodeIdArr, int32_t nSurf);

/* -- multi-GPU: one process (or thread) per sub-domain ---------------------- */
/* 128-byte ncclUniqueId made by rank 0 and broadcast by the host (MPI_Bcast) */
int eqd_get_unique_id(void* id128);
int eqd_set_comm(eqd_handle* h, const void* id128, int32_t nranks, int32_t rank);
/* Alternative to eqd_get_unique_id / eqd_set_comm for a host that already has a communicator over the same
 * npx*npy*npz ranks (the reference's MPI_COMM_WORLD, eqdyna3d.f90:19-21): the library does its SET-UP exchanges
 * (the CUDA-IPC handles of the peer-memory areas, the sums of eqd_sum_shared) through the host's all-gather
 *     fn(ctx, send, bytes, recv)  ==  MPI_Allgather(send, bytes, MPI_BYTE, recv, bytes, MPI_BYTE, comm),  0 = ok,
 * called collectively (every rank, same order, from the thread that calls eqd_sum_shared / the first eqd_run),
 * and the step-loop exchange of MPI4NodalQuant runs over peer memory.  No NCCL communicator is created (its start-up
 * costs seconds on an 8-GPU box); where peer memory cannot be mapped, also call eqd_set_comm.                       */
typedef int32_t (*eqd_allgather_fn)(void* ctx, const void* send, int64_t bytes, void* recv);
int eqd_set_host_comm(eqd_handle* h, int32_t nranks, int32_t rank, eqd_allgather_fn fn, void* ctx);
/* init-time shared-node sums the reference does with MPI before the loop:
 * nodalMassArr, fnms (assembleGlobalMass.f90:40-41) and arn (meshgen.f90:154).
 * Call once after eqd_set_* if the host has NOT already summed them.         */
int eqd_sum_shared(eqd_handle* h);

/* -- run steps nt_begin..nt_end inclusive (1-based like `do nt = 1, nstep`) - */
int eqd_run(eqd_handle* h, int32_t nt_begin, int32_t nt_end);
/* Several sub-domains driven by ONE process (handles ordered by rank id, all
 * npx*npy*npz of them; one device or one device each): lock-step steps with
 * device-to-device halo copies instead of NCCL.  The host must have summed the
 * shared mass / fnms / arn itself (as the in-process stand-in host does).     */
int eqd_run_group(eqd_handle** hs, int32_t n, int32_t nt_begin, int32_t nt_end);
int eqd_fetch(eqd_handle* h, int32_t which, void* dst, int64_t dst_bytes);
/* algorithmic counters for the benchmark: elements by type, kernel launches */
int eqd_get_counts(const eqd_handle* h, int64_t* n_regular, int64_t* n_pml,
                   int64_t* n_pairs, int64_t* launches);
/* elements swept with closed-form box operators (option "box"); valid after the first run */
int eqd_get_box_counts(const eqd_handle* h, int64_t* n_regular_box, int64_t* n_pml_box);
/* marching classes (option "march"): out8 = elements in bundles, bundles, node slots, nodes the bundles update
 * themselves, CTAs of the persistent launch; PML elements in bundles, PML bundles, PML node slots */
int eqd_get_march_counts(const eqd_handle* h, int64_t* out8);
/* transport of the step-loop exchange once the first run / eqd_sum_shared has set it up: 0 = no rank neighbours,
 * 1 = ncclSend / ncclRecv, 2 = peer memory (CUDA IPC, k_halo_send / k_halo_recv)                                */
int eqd_get_halo_mode(const eqd_handle* h);
int eqd_get_timing(const eqd_handle* h, double* ms_slots /*[EQD_T_NSLOTS]*/);
/* options: "timing" 1 = CUDA-event timing of every phase (2 = also reset the
 * accumulated slots and the launch counter); "overlap" -1 auto (default) / 0 serial
 * step / 1 halo + fault solver on a second stream under the next bulk node update /
 * 2 also rank-face tiles first; "reserve" CTAs left free by the overlapped interior
 * sweep (mode 2); "node_variant" launch-bounds variant of the node update; and,
 * before eqd_set_mesh only, the tile bricks "reg_bx/bz/by", "pml_bx/bz/by";
 * "box" (before the first eqd_run / eqd_sum_shared) 0 off (default) / 1 regular
 * classes / 2 also PML: tiles whose elements are all axis-aligned hexahedra
 * (exact test on meshCoor) use eleshp = sign*a_d, phi = ha, ss = diag in closed
 * form and stream 15 instead of 71 (PML: 33 instead of 89) operator rows;
 * "box_compact" 1 = a regular class whose tiles are ALL box tiles uses the kernel
 * variant whose stage buffer holds only those 15 rows (three CTAs per SM).
 * "bank_order" (before eqd_set_mesh) 1 = bank-aware element order inside the tiles,
 * 2 = residue numbering of the tile-local nodes of complete bricks (conflict free by
 * construction; see eqd_plan_bank_model; default 0 = ascending ids).
 * "march" (before eqd_set_mesh) 1 = elastic runs (C_elastic = 1, C_Q = 0, C_hg = 1, rdampm = 0) sweep the
 * axis-aligned hexahedra of the structured grid with the marching kernel (eqd_march.h): bundles of element
 * columns whose inner nodes are updated by the element sweep itself, connectivity implicit (default 0).
 * Unknown keys return 4.  See DESIGN.md sections 3-4.                          */
int eqd_set_option(eqd_handle* h, const char* key, int32_t value);

/* Host-only self-check of the tile planner (no GPU needed): cuts the given
 * connectivity (nodeElemIdRelation(8,Ne), 1-based, meshgen.f90:702-741) into the
 * tiles the element kernels sweep and replays their shared-memory assembly
 * schedule looking for write conflicts.  stats[24]: per class (regular,
 * regular-on-PML-node, PML) tiles, elements, padded slots, tile-node slots,
 * max tile nodes, max colours, multi-colour tiles, grid inferred.  0 = ok.    */
int eqd_plan_check(int32_t Nn, int32_t Ne, const int32_t* nodeElemIdRelation,
                   const int32_t* elemTypeArr, const int32_t* numOfDofPerNodeArr,
                   int64_t* stats);

/* Host-only model of the tile kernels' shared-memory bank conflicts (no GPU needed): plans
 * the tiles as eqd_set_mesh would and counts, per class, the wavefronts of the half-warps'
 * 8-byte corner accesses (gathers and force updates index the tile's node rows by local
 * node): out[9] = {conflict-free, ascending element order, order chosen under bank_order}
 * x {regular, regular-on-PML-node, PML}.  With bank_order = 1 (eqd_set_option
 * "bank_order", before eqd_set_mesh) the elements inside a tile are ordered along the grid
 * axis permutation that minimises the count; with bank_order = 2 the element order is kept and
 * the tile-local nodes of complete bricks are numbered by residue instead (out[.+2] is then
 * the count under that numbering).  The planner's invariants are re-checked.            */
int eqd_plan_bank_model(int32_t Nn, int32_t Ne, const int32_t* nodeElemIdRelation,
                        const int32_t* elemTypeArr, const int32_t* numOfDofPerNodeArr,
                        int32_t bank_order, int64_t* out);

/* Host-only check of the closed-form operators the regular tile kernel uses on
 * axis-aligned hexahedra under eqd_set_option("box", 1) (no GPU needed): for every
 * element whose eight nodes form an exact box in meshCoor, the strain, B^T t
 * forces (calcElemKU.f90:44-60,175-189) and hourglass forces (hrglss.f90:20-54)
 * of a pseudo-random field are computed from eleshp / phi / ss and in closed
 * form.  dev[3] = largest relative deviations, *nBox = box elements.  0 = ok. */
int eqd_box_check(int32_t Nn, int32_t Ne, const double* meshCoor, const int32_t* nodeElemIdRelation,
                  const int32_t* elemTypeArr, const double* eleshp, const double* phi, const double* ss,
                  int64_t* nBox, double* dev);

/* Host-only self-check of the marching kernel (option "march", eqd_march.h) -- no GPU needed.  Plans the bundles of
 * a sub-domain from its connectivity as eqd_set_mesh would for a launch of `grid` CTAs, then runs the kernel's phases
 * over all thread ids, CTA after CTA, for ONE step on the given nodal fields (asynchronous copies complete at issue).
 * In/out: stress6(6,Ne) (slots 1..6 of every element's stresses), vel / disp (3,Nn): updated where the device would
 * update them.  Out: fsum(3,Nn) = the partial forces summed per node (+ the complete force of the fused nodes when
 * update = 0), fusedFlag(Nn), inBundle(Ne), stats[8] = elements in bundles, bundles, node slots, fused nodes, regular
 * elements left to the tile kernels, grid, element slots, 0.  Returns 0 or a line number of eqd_march.cu.          */
int eqd_march_emulate(int32_t Nn, int32_t Ne, const double* meshCoor, const int32_t* nodeElemIdRelation,
                      const int32_t* elemTypeArr, const int32_t* numOfDofPerNodeArr, int32_t grid, const double* eleshp,
                      const double* ss, const double* eledet, const double* mat, double* stress6, double* vel, double* disp,
                      const double* mass, double dt, double rdampk, double w, int32_t update, double* fsum,
                      int32_t* fusedFlag, int32_t* inBundle, int64_t* stats);

/* The same for the PML bundles (eqd_march_pml.h): damps(3,Ne) = damping profile at the element centroids
 * (assembleGlobalKU.f90:130-213), stress21(21,Ne) in/out = the 21 stresses of every PML element, f12(12,Nn) out = the
 * twelve split-field force rows (assembleGlobalKU.f90:328-344) summed per node over the bundle elements.           */
int eqd_march_pml_emulate(int32_t Nn, int32_t Ne, const double* meshCoor, const int32_t* nodeElemIdRelation,
                          const int32_t* elemTypeArr, const int32_t* numOfDofPerNodeArr, int32_t grid, const double* eleshp,
                          const double* ss, const double* eledet, const double* mat, const double* damps, double* stress21,
                          const double* vel, const double* disp, double dt, double rdampk, double w, double* f12,
                          int32_t* inBundle, int64_t* stats);

#ifdef __cplusplus
}
#endif
#endif
