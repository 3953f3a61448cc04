// eqdyna_host -- standalone driver: the sequence of the reference's main program
// (src/eqdyna3d.f90:33-84) for a case directory, with `call driver` replaced by the
// CUDA step library.  It is the C++ counterpart of eqdyna_b200/csrc/fortran/
// driver_cuda.f90 for an image without a Fortran compiler:
//
//   eqdyna_host <case_dir> [-o out_dir] [-np npx npy npz] [-nstep n] [-device-ops]
//               [-box 0|1|2] [-box-compact 0|1] [-march 0|1|2] [-chunk n] [-gm]
//
//   1. libeqdyna_host.so  reads b*.txt + on_fault_vars_input.bin, builds every
//      sub-domain (mesh4num, meshgen, on-fault load, assembleGlobalMass, init_vel)
//   2. one eqd_handle per sub-domain (device = rank mod #devices), fed through the
//      same eqd_set_* calls the Fortran host would make
//   3. eqd_run_group steps all sub-domains in lock step (device-to-device halos);
//      one-rank-per-process runs use eqd_set_comm + eqd_run instead (bench.py)
//   4. eqd_fetch into the host arrays, then the writers of library_output.f90
//
// Exit codes mirror the reference's stop sites: 1 NaN velocity (driver.f90:147-152),
// 2 negative PML damping (comdampv.f90:114-118), 3 CUDA failure, 4 bad input.
// There is no CPU path: without a CUDA device eqd_create fails and so does this.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <string>
#include <vector>

#include "eqdyna_b200.h"
#include "eqdyna_host.h"

namespace {

double now() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

struct Opts {
  std::string dir, out = ".";
  int np[3] = {0, 0, 0};
  int nstep = 0, deviceOps = 0, box = 0, boxCompact = 0, chunk = 100, gm = 0, march = 0;
};

int usage() {
  std::fprintf(stderr,
               "usage: eqdyna_host <case_dir> [-o out_dir] [-np npx npy npz] [-nstep n] [-device-ops]\n"
               "                   [-box 0|1|2] [-box-compact 0|1] [-march 0|1|2] [-chunk n] [-gm]\n");
  return EQD_ERR_ARG;
}

int fail(eqd_handle* h, const char* what, int rc) {
  char msg[1024] = "";
  if (h) eqd_last_error(h, msg, sizeof msg);
  std::fprintf(stderr, "eqdyna_host: %s failed with code %d: %s\n", what, rc, msg);
  return rc;
}

#define HOSTCK(call)                                                                  \
  do {                                                                                \
    if ((call) != 0) {                                                                \
      std::fprintf(stderr, "eqdyna_host: %s: %s\n", #call, eqh_last_error());         \
      return EQD_ERR_ARG;                                                             \
    }                                                                                 \
  } while (0)
#define DEVCK(h, call)                                       \
  do {                                                       \
    int rc_ = (call);                                        \
    if (rc_ != EQD_OK) return fail(h, #call, rc_);           \
  } while (0)

// the eqd_set_* sequence of driver_cuda.f90 for one sub-domain
int upload(eqd_handle* h, const eqh_view& v, const Opts& o) {
  DEVCK(h, eqd_set_option(h, "march", o.march));   // shapes the element classes: before eqd_set_mesh
  DEVCK(h, eqd_set_mesh(h, v.Nn, v.Ne, v.Neq, v.sizeEq, v.meshCoor, v.nodeElemIdRelation, v.elemTypeArr, v.numOfDofPerNodeArr,
                        v.eqNumStartIndexLoc, v.eqNumIndexArr, v.stressCompIndexArr, v.sizeStress));
  if (o.deviceOps) {
    DEVCK(h, eqd_compute_elem_ops(h, v.mat, v.eleporep, v.stressArr, v.pstrain));
    // in-process worlds have host-summed masses (eqh_world_sum_shared): keep them, so that shared
    // nodes carry the same mass on both sides without a device-side exchange
    DEVCK(h, eqd_set_nodal(h, v.nodalMassArr, v.fnms, v.v1, v.velArr, v.dispArr, v.nodalForceArr));
  } else {
    DEVCK(h, eqd_set_elem_ops(h, v.eleshp, v.eledet, v.elemass, v.mat, v.ss, v.phi, v.eleporep, v.stressArr, v.pstrain));
    DEVCK(h, eqd_set_nodal(h, v.nodalMassArr, v.fnms, v.v1, v.velArr, v.dispArr, v.nodalForceArr));
  }
  long npairs = 0;
  for (int i = 0; i < v.ntotft; ++i) npairs += v.nftnd[i];
  if (npairs > 0) DEVCK(h, eqd_set_fault(h, v.nftmx, v.nftnd, v.nsmp, v.un, v.us, v.ud, v.arn, v.fric, v.fnft));
  DEVCK(h, eqd_set_halo(h, v.numcount, v.fltnum, v.fltMPI, v.fltface[0], v.fltface[1], v.fltface[2], v.fltface[3], v.fltface[4],
                        v.fltface[5]));
  DEVCK(h, eqd_set_stations(h, v.nOff ? v.idhist : nullptr, v.nOff, v.nOn ? v.anonfs : nullptr, v.nOn,
                            v.nSurf ? v.surfaceNodeIdArr : nullptr, v.nSurf));
  DEVCK(h, eqd_set_option(h, "box", o.box));
  DEVCK(h, eqd_set_option(h, "box_compact", o.boxCompact));
  return EQD_OK;
}

// what library_output.f90 reads after `call driver` (eqdyna3d.f90:75-79)
int download(eqd_handle* h, const eqh_view& v, int ntDone) {
  const int64_t d = sizeof(double);
  DEVCK(h, eqd_fetch(h, EQD_F_DISP, v.dispArr, d * 3 * v.Nn));
  DEVCK(h, eqd_fetch(h, EQD_F_VEL, v.velArr, d * 3 * v.Nn));
  DEVCK(h, eqd_fetch(h, EQD_F_V1, v.v1, d * v.Neq));
  DEVCK(h, eqd_fetch(h, EQD_F_STRESS, v.stressArr, d * v.sizeStress));
  if (v.params.C_elastic == 0) DEVCK(h, eqd_fetch(h, EQD_F_PSTRAIN, v.pstrain, d * v.Ne));
  long npairs = 0;
  for (int i = 0; i < v.ntotft; ++i) npairs += v.nftnd[i];
  if (npairs > 0) {
    DEVCK(h, eqd_fetch(h, EQD_F_FRIC, v.fric, d * 100 * (int64_t)v.nftmx * v.ntotft));
    DEVCK(h, eqd_fetch(h, EQD_F_FNFT, v.fnft, d * (int64_t)v.nftmx * v.ntotft));
    DEVCK(h, eqd_fetch(h, EQD_F_ONFAULT_HIST, v.onFaultQuantHistSCECForm, d * 12 * (int64_t)v.nstep * v.nOnAlloc));
  }
  if (v.nOff) DEVCK(h, eqd_fetch(h, EQD_F_OFFFAULT_HIST, v.OffFaultStGramSCEC, d * (6 * (int64_t)v.nOff + 1) * v.nstep));
  if (v.params.outputGroundMotion == 1 && v.nGmAlloc > 0 && ntDone > 0) {
    const int k = (ntDone - 1) / 10 + 1;   // steps with mod(nt,10) == 1 (driver.f90:30)
    if (k <= v.nGmAlloc) {
      if (v.gmHist && v.nSurf) DEVCK(h, eqd_fetch(h, EQD_F_GM, v.gmHist, d * 3 * (int64_t)v.nSurf * k));
      if (v.srcEvolHist && v.nftnd[0] > 0) DEVCK(h, eqd_fetch(h, EQD_F_SRC_EVOL, v.srcEvolHist, d * (int64_t)v.nftnd[0] * k));
      *v.nGmSamples = k;
    }
  }
  return EQD_OK;
}

}  // namespace

int main(int argc, char** argv) {
  Opts o;
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i];
    auto need = [&](int n) { return i + n < argc; };
    if (a == "-o" && need(1)) o.out = argv[++i];
    else if (a == "-np" && need(3)) { for (int k = 0; k < 3; ++k) o.np[k] = std::atoi(argv[++i]); }
    else if (a == "-nstep" && need(1)) o.nstep = std::atoi(argv[++i]);
    else if (a == "-device-ops") o.deviceOps = 1;
    else if (a == "-box" && need(1)) o.box = std::atoi(argv[++i]);
    else if (a == "-box-compact" && need(1)) o.boxCompact = std::atoi(argv[++i]);
    else if (a == "-march" && need(1)) o.march = std::atoi(argv[++i]);
    else if (a == "-chunk" && need(1)) o.chunk = std::max(1, std::atoi(argv[++i]));
    else if (a == "-gm") o.gm = 1;
    else if (!a.empty() && a[0] != '-' && o.dir.empty()) o.dir = a;
    else return usage();
  }
  if (o.dir.empty()) return usage();
  {
    std::error_code ec;
    std::filesystem::create_directories(o.out, ec);   // the reference writes into its working directory
    if (ec) {
      std::fprintf(stderr, "eqdyna_host: cannot create %s: %s\n", o.out.c_str(), ec.message().c_str());
      return EQD_ERR_ARG;
    }
  }

  const double tStart = now();
  double comp[10] = {0};   // compTimeInSeconds(1:9), MPICommTimeInSeconds (eqdyna3d.f90:59-84)
  eqh_world* w = nullptr;
  HOSTCK(eqh_world_create(o.dir.c_str(), o.np[0], o.np[1], o.np[2], o.nstep, &w));
  if (o.gm) HOSTCK(eqh_world_set_switch(w, "outputGroundMotion", 1.0));
  double t0 = now();
  HOSTCK(eqh_world_build(w, -1));
  HOSTCK(eqh_world_sum_shared(w));
  comp[0] = now() - t0;    // input + mesh (the stand-in does not time mass assembly apart: slot 2 stays 0)
  const int nr = eqh_world_size(w);
  std::vector<eqh_view> views(nr);
  for (int r = 0; r < nr; ++r) HOSTCK(eqh_get_view(w, r, &views[r]));
  const int nstep = views[0].nstep;
  long ne = 0;
  for (const eqh_view& v : views) ne += v.Ne;
  std::printf(" eqdyna_host: %d sub-domain(s), %ld elements, %d steps, dt = %g s\n", nr, ne, nstep, views[0].params.dt);

  std::vector<eqd_handle*> hs(nr, nullptr);
  auto cleanup = [&] {
    for (eqd_handle* h : hs) if (h) eqd_destroy(h);
    eqh_world_destroy(w);
  };
  t0 = now();
  for (int r = 0; r < nr; ++r) {
    int rc = eqd_create(&views[r].params, -1, &hs[r]);   // -1: me modulo the visible devices
    if (rc != EQD_OK) {
      std::fprintf(stderr, "eqdyna_host: eqd_create failed with code %d (no CUDA device? the step library has no CPU path)\n", rc);
      cleanup();
      return rc;
    }
    rc = upload(hs[r], views[r], o);
    if (rc != EQD_OK) { cleanup(); return rc; }
  }
  comp[1] = now() - t0;

  // the loop of driver.f90:9-34, in chunks so that the banner keeps appearing
  t0 = now();
  for (int nt0 = 1; nt0 <= nstep; nt0 += o.chunk) {
    const int nt1 = std::min(nt0 + o.chunk - 1, nstep);
    std::printf(" =     Current time in dynamic rupture                               =\n");
    std::printf(" = %40s%7.3f    s\n", "", (nt0 - 1) * views[0].params.dt + views[0].params.dt);
    std::fflush(stdout);
    const int rc = nr == 1 ? eqd_run(hs[0], nt0, nt1) : eqd_run_group(hs.data(), nr, nt0, nt1);
    if (rc != EQD_OK) {
      fail(hs[0], "eqd_run", rc);
      cleanup();
      return rc;
    }
  }
  const double loop = now() - t0;
  comp[2] = loop;
  std::printf(" eqdyna_host: step loop %.3f s = %.3e element-steps/s\n", loop, loop > 0 ? (double)ne * nstep / loop : 0.0);

  t0 = now();
  for (int r = 0; r < nr; ++r) {
    int rc = download(hs[r], views[r], nstep);
    if (rc != EQD_OK) { cleanup(); return rc; }
    double tms[EQD_T_NSLOTS] = {0};
    eqd_get_timing(hs[r], tms);
    double c[10];
    std::memcpy(c, comp, sizeof c);
    c[9] = tms[EQD_T_HALO] * 1e-3;          // MPICommTimeInSeconds (only with option "timing")
    c[7] = now() - t0;
    c[8] = now() - tStart;
    eqh_set_comp_time(w, r, c);
    HOSTCK(eqh_write_outputs(w, r, o.out.c_str()));
  }
  cleanup();
  return 0;
}
