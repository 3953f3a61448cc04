// C API of the stand-in host (include/eqdyna_host.h): sequence of
// src/eqdyna3d.f90:33-70 for one or all sub-domains of a case.
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>

#include "eqdyna_host.h"
#include "eqh_state.h"

struct eqh_world {
  eqh::CaseInput in;
  std::vector<std::unique_ptr<eqh::RankState>> ranks;
};

namespace {
thread_local std::string g_err;
template <class F>
int guarded(F&& f) {
  try {
    f();
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return EQD_ERR_ARG;
  } catch (...) {
    g_err = "unknown error";
    return EQD_ERR_ARG;
  }
}
}  // namespace

extern "C" {

const char* eqh_last_error(void) { return g_err.c_str(); }

int eqh_world_create(const char* case_dir, int npx, int npy, int npz, int nstep, eqh_world** out) {
  return guarded([&] {
    auto w = std::make_unique<eqh_world>();
    eqh::read_case(case_dir, w->in);
    if (npx > 0 && npy > 0 && npz > 0) { w->in.npx = npx; w->in.npy = npy; w->in.npz = npz; }
    if (nstep > 0) w->in.nstep = nstep;
    // warning.f90:6-19
    if (w->in.C_elastic == 0 && w->in.C_Q == 1) throw std::runtime_error("Q model can only work with elastic code (stop 1001)");
    if (w->in.C_Q == 1 && w->in.rat > 1) throw std::runtime_error("Q model can only work with uniform element size (stop 1002)");
    if (w->in.output_plastic == 1 && w->in.C_elastic != 0) throw std::runtime_error("Only output plastic strains for C_elastic == 0 (stop 1003)");
    w->ranks.resize((size_t)w->in.npx * w->in.npy * w->in.npz);
    *out = w.release();
  });
}

// The reference's compile-time switches (globalvar.f90:16,59-77) are plain
// initialisers there; a host build with other values is emulated by setting
// them before eqh_get_view.
int eqh_world_set_switch(eqh_world* w, const char* name, double value) {
  return guarded([&] {
    std::string n = name ? name : "";
    if (n == "C_Q") w->in.C_Q = (int)value;
    else if (n == "C_hg") w->in.C_hg = (int)value;
    else if (n == "kapa_hg") w->in.kapa_hg = value;
    else if (n == "rdampm") w->in.rdampm = value;
    else if (n == "outputGroundMotion") w->in.outputGroundMotion = (int)value;   // before eqh_world_build: it sizes the sample arrays
    else if (n == "outputFinalSurfDisp") w->in.outputFinalSurfDisp = (int)value;
    else if (n == "output_plastic") w->in.output_plastic = (int)value;
    // run-time parameters of bGlobal.txt (readInputFiles.f90:28-58,182-186): tests use them to reach the
    // branches no shipped case takes (friclaw 3, TPV 2802 / 201 / 202, Drucker-Prager yielding)
    else if (n == "friclaw") w->in.friclaw = (int)value;
    else if (n == "TPV") w->in.TPV = (int)value;
    else if (n == "C_nuclea") w->in.C_nuclea = (int)value;
    else if (n == "ccosphi") w->in.ccosphi = value;
    else if (n == "sinphi") w->in.sinphi = value;
    else if (n == "tv") w->in.tv = value;
    else if (n == "nucR") w->in.nucR = value;
    else if (n == "nucT") w->in.nucT = value;
    else if (n == "nucRuptVel") w->in.nucRuptVel = value;
    else if (n == "nucdtau0") w->in.nucdtau0 = value;
    else throw std::runtime_error("eqh_world_set_switch: unknown switch " + n);
    if (w->in.C_elastic == 0 && w->in.C_Q == 1) throw std::runtime_error("Q model can only work with elastic code (stop 1001)");
  });
}

int eqh_world_destroy(eqh_world* w) {
  delete w;
  return 0;
}

int eqh_world_size(const eqh_world* w) { return (int)w->ranks.size(); }

static void build_one(eqh_world* w, int r) {
  auto s = std::make_unique<eqh::RankState>();
  s->me = r;
  s->in = &w->in;
  eqh::mesh4num(w->in, *s);
  eqh::meshgen(w->in, *s);
  eqh::load_on_fault(w->in, *s);
  if (w->in.mode == 2) eqh::load_on_fault_restart(w->in, *s);   // eqdyna3d.f90:60
  eqh::find_surface_nodes(w->in, *s);
  eqh::alloc_after_meshgen(w->in, *s);
  eqh::assemble_global_mass(w->in, *s);
  eqh::init_vel(w->in, *s);
  w->ranks[r] = std::move(s);
}

int eqh_world_build(eqh_world* w, int rank) {
  return guarded([&] {
    if (rank >= (int)w->ranks.size()) throw std::runtime_error("rank out of range");
    if (rank >= 0) build_one(w, rank);
    else
      for (int r = 0; r < (int)w->ranks.size(); ++r) build_one(w, r);
  });
}

int eqh_world_sum_shared(eqh_world* w) {
  return guarded([&] {
    std::vector<eqh::RankState*> world;
    for (auto& p : w->ranks) {
      if (!p) throw std::runtime_error("eqh_world_sum_shared needs every rank built in-process");
      world.push_back(p.get());
    }
    eqh::exchange_arn(w->in, world);
    eqh::exchange_nodal(w->in, world, 0, 3);
    eqh::exchange_nodal(w->in, world, 1, 1);
  });
}

int eqh_get_view(eqh_world* w, int rank, eqh_view* v) {
  return guarded([&] {
    if (rank < 0 || rank >= (int)w->ranks.size() || !w->ranks[rank]) throw std::runtime_error("rank not built");
    eqh::RankState& s = *w->ranks[rank];
    std::memset(v, 0, sizeof *v);
    v->params = s.params();
    v->Nn = s.totalNumOfNodes; v->Ne = s.totalNumOfElements; v->Neq = s.totalNumOfEquations;
    v->sizeEq = s.sizeOfEqNumIndexArr; v->sizeStress = 5 * s.sizeOfEqNumIndexArr;
    v->stressUsed = s.sizeOfStressDofIndexArr;
    v->nx = s.nx; v->ny = s.ny; v->nz = s.nz;
    v->nftmx = s.nftmx; v->ntotft = w->in.ntotft; v->nOn = s.numOfOnFaultStCount; v->nOnAlloc = s.nOnAlloc;
    v->nOff = s.numOfOffFaultStCount; v->nSurf = s.surface_nnode; v->nstep = w->in.nstep;
    v->meshCoor = s.meshCoor.data(); v->nodeElemIdRelation = s.nodeElemIdRelation.data();
    v->elemTypeArr = s.elemTypeArr.data(); v->numOfDofPerNodeArr = s.numOfDofPerNodeArr.data();
    v->eqNumStartIndexLoc = s.eqNumStartIndexLoc.data(); v->eqNumIndexArr = s.eqNumIndexArr.data();
    v->stressCompIndexArr = s.stressCompIndexArr.data();
    v->eleshp = s.eleshp.data(); v->eledet = s.eledet.data(); v->elemass = s.elemass.data();
    v->mat = s.mat.data(); v->ss = s.ss.data(); v->phi = s.phi.data(); v->eleporep = s.eleporep.data();
    v->stressArr = s.stressArr.data(); v->pstrain = s.pstrain.data();
    v->nodalMassArr = s.nodalMassArr.data(); v->fnms = s.fnms.data(); v->v1 = s.v1.data();
    v->velArr = s.velArr.data(); v->dispArr = s.dispArr.data(); v->nodalForceArr = s.nodalForceArr.data();
    v->nftnd = s.nftnd.data(); v->nsmp = s.nsmp.data(); v->un = s.un.data(); v->us = s.us.data();
    v->ud = s.ud.data(); v->arn = s.arn.data(); v->fric = s.fric.data(); v->fnft = s.fnft.data();
    v->numcount = s.numcount; v->fltnum = s.fltnum; v->fltMPI = s.fltMPI;
    for (int k = 0; k < 6; ++k) v->fltface[k] = s.fltface[k].empty() ? nullptr : s.fltface[k].data();
    v->idhist = s.idhist.empty() ? nullptr : s.idhist.data();
    v->anonfs = s.anonfs.data();
    v->surfaceNodeIdArr = s.surfaceNodeIdArr.empty() ? nullptr : s.surfaceNodeIdArr.data();
    v->onFaultQuantHistSCECForm = s.onFaultQuantHistSCECForm.data();
    v->OffFaultStGramSCEC = s.OffFaultStGramSCEC.empty() ? nullptr : s.OffFaultStGramSCEC.data();
    v->hypoLog = s.hypoLog.data();
    v->onFaultTPHist = s.onFaultTPHist.empty() ? nullptr : s.onFaultTPHist.data();
    v->gmHist = s.gmHist.empty() ? nullptr : s.gmHist.data();
    v->srcEvolHist = s.srcEvolHist.empty() ? nullptr : s.srcEvolHist.data();
    v->nGmSamples = &s.nGmSamples;
    v->nGmAlloc = s.nGmAlloc;
  });
}

int eqh_write_outputs(eqh_world* w, int rank, const char* out_dir) {
  return guarded([&] {
    if (rank < 0 || rank >= (int)w->ranks.size() || !w->ranks[rank]) throw std::runtime_error("rank not built");
    eqh::RankState& s = *w->ranks[rank];
    // eqdyna3d.f90:75-77
    eqh::write_onfault_stations(s, out_dir);
    eqh::write_offfault_stations(s, out_dir);
    eqh::write_frt(s, out_dir);
    if (w->in.output_plastic == 1) eqh::write_plastic_strain(s, out_dir);
    if (w->in.outputFinalSurfDisp == 1) eqh::write_final_surf_disp(s, out_dir);
    // written while stepping / before it in the reference (driver.f90:30-33, eqdyna3d.f90:61)
    eqh::write_surface_coor(s, out_dir);
    eqh::write_gm(s, out_dir);
    eqh::write_src_evol(s, out_dir);
    if (s.compTimeSet) eqh::write_comp_time(s, out_dir);
  });
}

int eqh_set_comp_time(eqh_world* w, int rank, const double* t10) {
  return guarded([&] {
    if (rank < 0 || rank >= (int)w->ranks.size() || !w->ranks[rank] || !t10) throw std::runtime_error("rank not built");
    for (int k = 0; k < 10; ++k) w->ranks[rank]->compTime[k] = t10[k];
    w->ranks[rank]->compTimeSet = true;
  });
}

int eqh_release_operators(eqh_world* w, int rank) {
  return guarded([&] {
    if (rank < 0 || rank >= (int)w->ranks.size() || !w->ranks[rank]) throw std::runtime_error("rank not built");
    eqh::RankState& s = *w->ranks[rank];
    std::vector<double>().swap(s.eleshp);
    std::vector<double>().swap(s.phi);
    std::vector<double>().swap(s.elemass);
    std::vector<double>().swap(s.ss);
  });
}

}  // extern "C"
