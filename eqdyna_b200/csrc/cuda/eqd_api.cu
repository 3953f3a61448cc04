// C ABI of libeqdyna_b200.so (include/eqdyna_b200.h): state upload, the step
// loop of src/driver.f90:9-34 as a sequence of sm_100a kernels, halo exchange
// (MPI4NodalQuant, src/assembleGlobalMass.f90:58-281) over NCCL send/recv or,
// for several sub-domains driven by one process, over device-to-device copies.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "eqd_box.h"
#include "eqd_dev.cuh"
#include "eqd_kernels.h"
#include "eqd_march_plan.h"
#include "eqd_par.h"
#include "eqd_tiles.h"
#include "eqdyna_b200.h"

#ifndef EQD_NCCL_PATH
#define EQD_NCCL_PATH ""
#endif

using namespace eqd;

namespace {

struct CudaError : std::runtime_error { using std::runtime_error::runtime_error; };
struct ArgError : std::runtime_error { using std::runtime_error::runtime_error; };

#define CK(x)                                                                                      \
  do {                                                                                             \
    cudaError_t e_ = (x);                                                                          \
    if (e_ != cudaSuccess)                                                                         \
      throw CudaError(std::string(#x) + ": " + cudaGetErrorString(e_) + " (" __FILE__ ":" + std::to_string(__LINE__) + ")"); \
  } while (0)

// the calling handle's pinned staging buffers while a set-up call is running (else pageable copies)
thread_local StagedCopy* t_stage = nullptr;
thread_local cudaStream_t t_stageStream = nullptr;

// the pinned double buffer of a device: allocated at its first use in the process, kept for every later handle (pinning
// 256 MB costs 0.1-0.2 s, more than the uploads of a mid-sized sub-domain take)
std::shared_ptr<StagedCopy> shared_stage(int device) {
  static std::mutex mu;
  static std::map<int, std::shared_ptr<StagedCopy>> pool;
  std::lock_guard<std::mutex> g(mu);
  auto& p = pool[device];
  if (!p) p.reset(new StagedCopy());
  return p;
}

inline void h2d(void* dst, const void* src, size_t bytes) {
  if (t_stage && bytes >= (8u << 20)) {
    cudaError_t e_ = t_stage->h2d(dst, src, bytes, t_stageStream);
    if (e_ != cudaSuccess) throw CudaError(std::string("staged host->device copy: ") + cudaGetErrorString(e_));
  } else {
    CK(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice));
  }
}

template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
  void alloc(size_t count, bool zero = true) {
    release();
    n = count;
    if (count == 0) return;
    CK(cudaMalloc(&p, count * sizeof(T)));
    if (zero) CK(cudaMemset(p, 0, count * sizeof(T)));
  }
  template <class A>
  void upload(const std::vector<T, A>& h) {
    alloc(h.size(), false);
    if (!h.empty()) h2d(p, h.data(), h.size() * sizeof(T));
  }
  std::vector<T> download() const {
    std::vector<T> h(n);
    if (n) CK(cudaMemcpy(h.data(), p, n * sizeof(T), cudaMemcpyDeviceToHost));
    return h;
  }
};

// EQD_VERBOSE=1: wall-clock laps of the set-up phases on stderr
struct Lap {
  bool on; const char* what; std::chrono::steady_clock::time_point t, ts;
  explicit Lap(const char* w) : on(getenv("EQD_VERBOSE") != nullptr), what(w), t(std::chrono::steady_clock::now()), ts(t) {}
  void lap(const char* label) {
    if (!on) return;
    cudaDeviceSynchronize();
    auto n = std::chrono::steady_clock::now();
    fprintf(stderr, "[eqd] %s: %s %.3f s\n", what, label, std::chrono::duration<double>(n - t).count());
    t = ts = n;
  }
  void sub(const char* label) {   // a part of the lap that follows: printed, the lap's own clock runs on
    if (!on) return;
    cudaDeviceSynchronize();
    auto n = std::chrono::steady_clock::now();
    fprintf(stderr, "[eqd] %s:   . %s %.3f s\n", what, label, std::chrono::duration<double>(n - ts).count());
    ts = n;
  }
};

inline int pad32(long n) { return (int)((n + 31) / 32 * 32); }

// ---- NCCL through dlopen: the library must load (and export its symbols) on a
// box without NCCL; the communicator is only needed for multi-process runs.
struct Nccl {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string why;
  bool load() {
    if (lib) return true;
    const char* names[] = {getenv("EQD_NCCL_LIB"), "libnccl.so.2", EQD_NCCL_PATH, "libnccl.so"};
    for (const char* nm : names) {
      if (!nm || !*nm) continue;
      lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (lib) break;
      why = dlerror();
    }
    if (!lib) return false;
#define SYM(f) f = (decltype(f))dlsym(lib, "nccl" #f); if (!f) { why = "missing nccl" #f; lib = nullptr; return false; }
    SYM(GetUniqueId) SYM(CommInitRank) SYM(CommDestroy) SYM(Send) SYM(Recv) SYM(AllGather) SYM(GroupStart) SYM(GroupEnd) SYM(GetErrorString)
#undef SYM
    return true;
  }
};
Nccl g_nccl;

#define NK(x)                                                                               \
  do {                                                                                      \
    ncclResult_t r_ = (x);                                                                  \
    if (r_ != ncclSuccess) throw CudaError(std::string(#x) + ": " + g_nccl.GetErrorString(r_)); \
  } while (0)

struct ElemClass {
  int n = 0, S = 0, nf = 3, nstress = 6;
  // tiles (eqd_tiles.h): launch order = rank-face tiles first, then interior
  int nTiles = 0, nFaceTiles = 0, LS = 1, PFS = 4;
  raw_vector<int> refId;  // [S] slot -> reference element (0-based), -1 = padding
  std::vector<int> tileNodeH;
  raw_vector<int> tnodeH;
  std::vector<int4> tileRecH;  // planner order
  DevBuf<int4> tileRec;        // launch order
  std::vector<uint8_t> tileBoxH;  // planner order: 1 = all elements of the tile are axis-aligned hexahedra (option "box")
  DevBuf<uint8_t> tileBox;        // launch order; empty = closed-form box operators off
  long nBoxElems = 0;
  DevBuf<int> tnode;
  DevBuf<uint16_t> lconn;
  DevBuf<double> shp, phi, ss, lam, mu, det, rho, vp, stress, qmem, porep, pstrain, emass, damps, pf;
  DevBuf<uint8_t> qcls;
};

struct Face {
  int n = 0;       // doubles exchanged
  int nb = -1;     // neighbour rank, -1 = inactive
  DevBuf<uint32_t> idx;
  DevBuf<double> send, recv;
  // init-time node-wise exchange (mass, fnms) and pair-wise (arn)
  std::vector<int> nodes;   // face node ids (0-based), reference order, incl. face masters
  std::vector<int> pairs;   // fault pair slots on this face (device pair index)
};

}  // namespace

struct eqd_handle {
  eqd_params p{};
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  // ---- host copies (0-based where noted)
  int Nn = 0, Ne = 0, Neq = 0, sizeEq = 0, sizeStress = 0;
  raw_vector<int> conn;        // (8,Ne) 0-based
  raw_vector<int> etype, ndof, eqStart, eqIdx, stressIdx;
  raw_vector<double> coor;     // (3,Nn)
  std::vector<int> info;       // per node
  std::vector<int> elemCode;   // per element: class | slot<<2
  std::vector<double> fnmsH, massH;
  bool massFromDevice = false;  // eqd_compute_elem_ops filled massH / fnmsH / dMass
  int NnS = 0, Np = 0, NpS = 0;
  bool body = false, plastic = false, qmode = false;
  // ---- device
  ElemClass cls[NCLS];
  // marching class (CLS_MARCH, eqd_march.h): bundles of box elements; cls[CLS_MARCH] holds its SoA rows
  // (shp = [3][S] a_x a_y a_z, ss = [3][S] the diagonal, no phi), refId, partial buffer and node ids
  DevBuf<MarchBundle> mRec;
  DevBuf<int> mCtaFirstA, mCtaFirstB, mCode, mSlotBundle;   // work lists: A = bundles that may touch a rank face, B = the rest
  int mGrid = 0, mBundles = 0;
  int mNxg = 0;             // node planes of the grid in x (for the planner's end caps)
  int mSlotB0 = 0;          // first node slot of the interior work list
  bool mSplitOk = true;     // no strip of the interior work list touches a rank face (checked when the faces are known)
  long mFused = 0;
  int optMarch = 0;         // option "march" (before eqd_set_mesh)
  bool sweepFused = false;  // the last element sweep updated the fused nodes itself
  // PML bundles (CLS_MARCHP, eqd_march_pml.h): cls[CLS_MARCHP] holds shp = [3][S], ss = [3][S], damps, 21 stress rows; their
  // node slots follow the PML tile class's in cls[CLS_PML].pf (first one: pSlotBase)
  DevBuf<MarchBundle> pRec;
  DevBuf<int> pCtaFirstA, pCtaFirstB, pCode, pSlotBundle;
  int pGrid = 0, pBundles = 0, pSlotBase = 0, pSlotB0 = 0;
  raw_vector<unsigned char> mOwner;   // [S] 1: the slot is its element's own copy (0: padding / ghost copy of a neighbouring strip's element)
  DevBuf<int> dNodeList;    // free 3-dof nodes no bundle updates, ascending (list variant of the node update)
  int nNodeList = 0;
  DevBuf<int> dInfo, dPmlNode, dSpecial;
  DevBuf<uint8_t> dSlotCnt;
  DevBuf<uint32_t> dSlotTab;
  // velArr / dispArr.  With ghost sharing between marching strips (option march = 2) a node is read, in the sweep
  // that updates it, by strips other than the one updating it: v and d are then double-buffered -- every kernel that
  // advances nodes reads buffer `cur` and writes the other one, and the two swap after the node-update phase.
  DevBuf<double> dVelB[2], dDispB[2];
  int cur = 0;
  bool pingpong = false;
  DevBuf<double>& velCur() { return dVelB[cur]; }
  DevBuf<double>& dispCur() { return dDispB[cur]; }
  const DevBuf<double>& velCur() const { return dVelB[cur]; }
  const DevBuf<double>& dispCur() const { return dDispB[cur]; }
  const DevBuf<double>& velNext() const { return dVelB[pingpong ? cur ^ 1 : cur]; }
  const DevBuf<double>& dispNext() const { return dDispB[pingpong ? cur ^ 1 : cur]; }
  DevBuf<double> dMass, dV1p, dDampp, dForce, dAccel0;
  int nSpecial = 0;
  DevBuf<StepState> dState;
  // fault
  int nftmx = 0, ntotft = 0, nPairs = 0, PS = 0;
  std::vector<int> nftnd, pairRef;  // pairRef: device pair -> (i-1) + nftmx*ift
  std::vector<int> pairNodeS, pairNodeM;
  DevBuf<int> dNodeS, dNodeM, dIft, dPairStation;
  DevBuf<double> dUn, dUs, dUd, dArn, dMassS, dMassM, dXs, dFric, dFnft;
  // stations / histories
  int nOff = 0, nOn = 0, nOnAlloc = 1, nSurf = 0, nGm = 0;
  std::vector<int> idhistH, anonfsH, surfH;
  std::vector<int> stationAlias;  // on-fault station -> the station whose device row holds its pair's record
  DevBuf<int> dIdhist, dSurf;
  DevBuf<double> dOnHist, dOffHist, dHypo, dTpHist, dGm, dSrc;
  // halo
  bool haloSet = false;
  int numcount[9] = {0}, fltnum[6] = {0}, fltMPI[6] = {0};
  std::vector<int> fltface[6];
  Face face[3][2];
  ncclComm_t comm = nullptr;
  int nranks = 1, rank = 0;
  // ncclCommInitRank runs on a helper thread from eqd_set_comm on, so that a host which calls eqd_set_comm right after
  // eqd_create overlaps the communicator's start-up with its uploads; comm_ready() joins before the first use
  std::thread commThread;
  ncclResult_t commRc = ncclSuccess;
  // eqd_set_host_comm: the host's own all-gather (MPI_Allgather) carries the set-up exchanges; no NCCL communicator
  eqd_allgather_fn hostAg = nullptr;
  void* hostCtx = nullptr;
  // Step-loop exchange over peer memory (option "halo", default on between processes of one box): every face has two
  // receive buffers and two flag words in ONE allocation of this sub-domain, which the neighbours map through CUDA IPC;
  // k_halo_send writes into the neighbour's copy, k_halo_recv polls the own one (eqd_kernels.cu).  NCCL is then used at
  // set-up only (unique id, the all-gather of the IPC handles, eqd_sum_shared).
  struct P2P {
    bool on = false;
    DevBuf<unsigned char> area;            // [face][parity] receive buffers | flag words | send counters
    size_t recvOff[3][2][2] = {{{0}}};     // byte offsets inside `area`
    size_t flagOff = 0, counterOff = 0;
    std::vector<void*> peerBase;           // [rank] the neighbour's area as mapped here (nullptr: not a neighbour)
    size_t peerRecvOff[3][2][2] = {{{0}}}; // [axis][my side][parity]: offset of the FACING face's buffer inside the neighbour's area
    size_t peerFlagOff[3][2] = {{0}};      // flag words of the facing face (two parities, consecutive)
    unsigned seq = 0;
  } p2p;
  int optHalo = -1;                        // -1 auto (peer memory when it can be set up), 0 ncclSend/ncclRecv, 1 peer memory or fail
  cudaEvent_t evPacked[3] = {nullptr, nullptr, nullptr};
  // overlap of the halo with the interior element sweep: tiles that
  // touch a rank-face node are swept first, their face sums are exchanged on
  // commStream while the main stream sweeps the interior tiles
  cudaStream_t commStream = nullptr;
  cudaEvent_t evFace = nullptr, evElem = nullptr, evComm = nullptr;
  bool anyFace = false;
  DevBuf<int> dSpecialA, dSpecialB;  // special nodes on active rank faces | the others (split-node pairs)
  int nSpecialA = 0, nSpecialB = 0;
  int optOverlap = -1, optReserve = 8, smCount = 148;  // overlap -1 = auto: 1 with rank neighbours, else 0
  std::shared_ptr<StagedCopy> stage;  // (one per device and process, shared by the handles on it) pinned staging of the set-up uploads (and of large eqd_fetch reads); small sub-domains drop it once the run starts
  // run state
  bool meshSet = false, opsSet = false, nodalSet = false, faultSet = false, finalized = false;
  int hostNt = 0;
  double hostTime = 0.0;
  long launches = 0;
  bool timing = false;
  int optNodeVariant = 4;
  int optBankOrder = 0;    // 1: bank-aware element order inside the tiles (eqd_tiles.h: TileShape::bankOrder); before eqd_set_mesh
  int optBoxCompact = 0;   // 1: classes whose tiles are ALL box tiles use the compact stage buffer (three CTAs per SM)
  int optBox = 0;   // closed-form operators on all-box tiles (eqd_box.h): 1 = regular classes, 2 = also PML; set before the first eqd_run
  int optTile[2][3] = {{kRegBrick[0], kRegBrick[1], kRegBrick[2]}, {kPmlBrick[0], kPmlBrick[1], kPmlBrick[2]}};  // brick of a regular / PML tile in elements along x, z, y (before eqd_set_mesh)
  double tms[EQD_T_NSLOTS] = {0};
  std::vector<cudaEvent_t> evs;  // timing events in flight: (start, stop, slot) triples
  std::vector<cudaEvent_t> evPool;  // timing events, created once and reused run after run (no create / destroy inside the step loop)
  size_t evPoolUsed = 0;

  NodeArgs nodeArgs() const;
  ElemArgs elemArgs(int c) const;
  MarchArgs marchArgs(bool update, int part) const;
  MarchPmlArgs marchPmlArgs(int part) const;
  FaultArgs faultArgs() const;
};

namespace {

template <class F>
int guarded(eqd_handle* h, F&& f) {
  struct StageScope {
    StageScope(eqd_handle* h) { if (h && h->stage) { t_stage = h->stage.get(); t_stageStream = h->stream; } }
    ~StageScope() { t_stage = nullptr; t_stageStream = nullptr; }
  } scope(h);
  try {
    if (h) CK(cudaSetDevice(h->device));
    f();
    return EQD_OK;
  } catch (const ArgError& e) {
    if (h) h->err = e.what();
    return EQD_ERR_ARG;
  } catch (const CudaError& e) {
    if (h) h->err = e.what();
    return EQD_ERR_CUDA;
  } catch (const std::exception& e) {
    if (h) h->err = e.what();
    return EQD_ERR_ARG;
  }
}

// comdampv.f90:1-120 for one PML node (returns the three distinct profile values)
bool comdampv(const eqd_params& P, double x2, double y, double z, double out[3]) {
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
    else if (x2 <= xmin0 && y >= xmax0) { damp[0] = std::fabs(x2 - xmin0); damp[1] = std::fabs(y - ymax0); }  // sic, comdampv.f90:29,61
    else if (x2 >= xmax0 && y > ymin0 && y < ymax0) { damp[0] = std::fabs(x2 - xmax0); damp[1] = 0.0; }
    else if (y <= ymin0 && x2 > xmin0 && x2 < xmax0) { damp[0] = 0.0; damp[1] = std::fabs(y - ymin0); }
    else if (x2 <= xmin0 && y > ymin0 && y < ymax0) { damp[0] = std::fabs(x2 - xmin0); damp[1] = 0.0; }
    else if (y >= ymax0 && x2 > xmin0 && x2 < xmax0) { damp[0] = 0.0; damp[1] = std::fabs(y - ymax0); }
    else { damp[0] = 0.0; damp[1] = 0.0; }
  }
  for (int i = 0; i < 3; ++i) {
    const double delta = P.nPML * maxd[i];
    out[i] = 3.0 * P.vmaxPML / 2.0 / delta * std::log(1.0 / P.R) * ((damp[i] / delta) * (damp[i] / delta));
  }
  return !(out[0] < 0.0 || out[1] < 0.0 || out[2] < 0.0);
}

// damping profile at a PML element's centroid, assembleGlobalKU.f90:121-213
void pml_elem_damps(const eqd_params& P, const double xc[3], double out[3]) {
  const double xmax2 = P.PMLb[0], xmin2 = P.PMLb[1], ymax2 = P.PMLb[2], ymin2 = P.PMLb[3], zmin2 = P.PMLb[4];
  const double maxd[3] = {P.PMLb[5], P.PMLb[6], P.PMLb[7]};
  double d[3] = {0, 0, 0};
  bool any = false;
  if (xc[2] < zmin2) { d[2] = std::fabs(xc[2] - zmin2); any = true; }
  else if (xc[2] > zmin2) { d[2] = 0.0; any = true; }
  if (any) {
    if (xc[0] > xmax2 && xc[1] > ymax2) { d[0] = std::fabs(xc[0] - xmax2); d[1] = std::fabs(xc[1] - ymax2); }
    else if (xc[0] > xmax2 && xc[1] < ymin2) { d[0] = std::fabs(xc[0] - xmax2); d[1] = std::fabs(xc[1] - ymin2); }
    else if (xc[0] < xmin2 && xc[1] < ymin2) { d[0] = std::fabs(xc[0] - xmin2); d[1] = std::fabs(xc[1] - ymin2); }
    else if (xc[0] < xmin2 && xc[1] > xmax2) { d[0] = std::fabs(xc[0] - xmin2); d[1] = std::fabs(xc[1] - ymax2); }  // sic, :150,:182
    else if (xc[0] > xmax2 && xc[1] > ymin2 && xc[1] < ymax2) { d[0] = std::fabs(xc[0] - xmax2); d[1] = 0.0; }
    else if (xc[1] < ymin2 && xc[0] > xmin2 && xc[0] < xmax2) { d[0] = 0.0; d[1] = std::fabs(xc[1] - ymin2); }
    else if (xc[0] < xmin2 && xc[1] > ymin2 && xc[1] < ymax2) { d[0] = std::fabs(xc[0] - xmin2); d[1] = 0.0; }
    else if (xc[1] > ymax2 && xc[0] > xmin2 && xc[0] < xmax2) { d[0] = 0.0; d[1] = std::fabs(xc[1] - ymax2); }
    else { d[0] = 0.0; d[1] = 0.0; }
  }
  for (int i = 0; i < 3; ++i) {
    const double delta = P.nPML * maxd[i];
    out[i] = 3 * P.vmaxPML / 2 / delta * std::log(1 / P.R) * ((d[i] / delta) * (d[i] / delta));
  }
}

// qconstant.f90:3-35 (single-precision literal tables, as in the reference)
void qconstant(double Q, int k, double& rtaok, double& rwk, double& c1) {
  static const float taok[8] = {1.72333e-3f, 1.80701e-3f, 5.38887e-3f, 1.99322e-2f, 8.49833e-2f, 4.09335e-1f, 2.05951f, 13.2629f};
  static const float alfk[8] = {1.66958e-2f, 3.81644e-2f, 9.84666e-3f, -1.36803e-2f, -2.85125e-2f, -5.37309e-2f, -6.65035e-2f, -1.33696e-1f};
  static const float betk[8] = {8.98758e-2f, 6.84635e-2f, 9.67052e-2f, 1.20172e-1f, 1.30728e-1f, 1.38746e-1f, 1.40705e-1f, 2.14647e-1f};
  const double pi = 4 * std::atan(1.0);
  double kapa = (double)3.071f + (double)1.433f * std::pow(Q, (double)(-1.158f)) * std::log(Q / 5);
  kapa = kapa / (1 + (double)0.415f * Q);
  rwk = kapa * (kapa * (double)alfk[k - 1] + (double)betk[k - 1]);
  rtaok = (double)taok[k - 1];
  const double ref = 2.0 * pi;
  const double ak0 = 1.0 - rwk * 8.0 / (1.0 + (rtaok * ref) * (rtaok * ref));
  const double bk0 = rwk * 8.0 * ref * rtaok / (1.0 + (rtaok * ref) * (rtaok * ref));
  c1 = 0.5 * std::pow(ak0 * ak0 + bk0 * bk0, -0.5);
  c1 = c1 * (1.0 + ak0 * std::pow(ak0 * ak0 + bk0 * bk0, -0.5));
}

void need(bool ok, const char* msg) { if (!ok) throw ArgError(msg); }

}  // namespace

NodeArgs eqd_handle::nodeArgs() const {
  NodeArgs A{};
  A.Nn = Nn; A.NnS = NnS; A.Np = Np; A.NpS = NpS;
  A.info = dInfo.p; A.pmlNode = dPmlNode.p; A.vel = velCur().p; A.disp = dispCur().p; A.mass = dMass.p;
  A.velOut = velNext().p; A.dispOut = dispNext().p;
  A.v1p = dV1p.p; A.dampp = dDampp.p; A.force = dForce.p;
  A.slotCnt = dSlotCnt.p; A.slotTab = dSlotTab.p;
  A.pfR = cls[CLS_REG].pf.p; A.SR = cls[CLS_REG].PFS;
  A.pfX = cls[CLS_REGX].pf.p; A.SX = cls[CLS_REGX].PFS;
  A.pfP = cls[CLS_PML].pf.p; A.SP = cls[CLS_PML].PFS;
  A.pfM = cls[CLS_MARCH].pf.p; A.SM = cls[CLS_MARCH].PFS;
  A.list = nullptr; A.nList = 0; A.fusedMode = 1;
  A.accel0 = nullptr; A.skipSpecial = 0; A.variant = optNodeVariant;
  A.dt = p.dt;
  A.st = dState.p;
  return A;
}

ElemArgs eqd_handle::elemArgs(int c) const {
  const ElemClass& C = cls[c];
  ElemArgs A{};
  A.n = C.n; A.S = C.S;
  A.tileRec = C.tileRec.p; A.tile0 = 0; A.ntiles = C.nTiles; A.maxGrid = 0;
  A.tnode = C.tnode.p; A.lconn = C.lconn.p; A.LS = C.LS; A.pf = C.pf.p; A.PFS = C.PFS;
  A.shp = C.shp.p; A.phi = C.phi.p; A.ss = C.ss.p;
  A.lam = C.lam.p; A.mu = C.mu.p; A.det = C.det.p; A.rho = C.rho.p; A.vp = C.vp.p;
  A.stress = C.stress.p; A.qmem = C.qmem.p; A.qcls = C.qcls.p;
  A.porep = C.porep.p; A.pstrain = C.pstrain.p; A.emass = C.emass.p; A.damps = C.damps.p;
  A.vel = velCur().p; A.disp = dispCur().p; A.NnS = NnS;
  A.dt = p.dt; A.rdampk = p.rdampk; A.rdampm = p.rdampm; A.w = p.w;
  // assembleGlobalKU.f90:16
  A.bodyz = (1.0 - p.C_elastic) * p.grav * (p.roumax - (p.gamar + 1.0) * p.rhow) / p.roumax;
  A.ccosphi = p.ccosphi; A.sinphi = p.sinphi;
  A.expdttv = p.tv != 0.0 ? std::exp(-p.dt / p.tv) : 0.0;
  A.kapa_hg = p.kapa_hg;
  A.tileBox = C.tileBox.n ? C.tileBox.p : nullptr;
  A.allBox = (optBoxCompact && C.tileBox.n && C.nBoxElems == C.n) ? 1 : 0;
  return A;
}

MarchArgs eqd_handle::marchArgs(bool update, int part) const {   // part 0: both work lists, 1: boundary list, 2: interior list
  const ElemClass& C = cls[CLS_MARCH];
  MarchArgs A{};
  A.rec = mRec.p; A.code = mCode.p;
  A.ctaFirstA = part != 2 ? mCtaFirstA.p : nullptr;
  A.ctaFirstB = part != 1 ? mCtaFirstB.p : nullptr;
  A.S = (size_t)C.S; A.NnS = (size_t)NnS; A.PFS = (size_t)C.PFS;
  A.a = C.shp.p; A.ss = C.ss.p; A.lam = C.lam.p; A.mu = C.mu.p; A.det = C.det.p; A.stress = C.stress.p;
  A.vel = velCur().p; A.disp = dispCur().p; A.velOut = velNext().p; A.dispOut = dispNext().p;
  A.mass = dMass.p; A.pf = C.pf.p; A.force = dForce.p;
  A.dt = p.dt; A.rdampk = p.rdampk; A.w = p.w; A.update = update ? 1 : 0; A.st = dState.p;
  return A;
}

MarchPmlArgs eqd_handle::marchPmlArgs(int part) const {
  const ElemClass& C = cls[CLS_MARCHP];
  MarchPmlArgs A{};
  A.rec = pRec.p; A.code = pCode.p;
  A.ctaFirstA = part != 2 ? pCtaFirstA.p : nullptr;
  A.ctaFirstB = part != 1 ? pCtaFirstB.p : nullptr;
  A.S = (size_t)C.S; A.NnS = (size_t)NnS; A.PFS = (size_t)cls[CLS_PML].PFS; A.slotBase = (size_t)pSlotBase;
  A.a = C.shp.p; A.ss = C.ss.p; A.lam = C.lam.p; A.mu = C.mu.p; A.det = C.det.p; A.damps = C.damps.p; A.stress = C.stress.p;
  A.vel = velCur().p; A.disp = dispCur().p; A.pf = cls[CLS_PML].pf.p;
  A.dt = p.dt; A.rdampk = p.rdampk; A.w = p.w;
  return A;
}

FaultArgs eqd_handle::faultArgs() const {
  FaultArgs A{};
  A.nPairs = nPairs; A.PS = PS; A.NnS = NnS;
  A.nodeS = dNodeS.p; A.nodeM = dNodeM.p; A.ift = dIft.p;
  A.un = dUn.p; A.us = dUs.p; A.ud = dUd.p; A.arn = dArn.p; A.massS = dMassS.p; A.massM = dMassM.p;
  A.xs = dXs.p; A.fric = dFric.p; A.fnft = dFnft.p;
  A.vel = velCur().p; A.disp = dispCur().p; A.force = dForce.p;
  A.pairStation = dPairStation.p; A.onHist = dOnHist.p; A.nstep = p.nstep;
  A.hypoLog = dHypo.p; A.tphist = dTpHist.p;
  A.st = dState.p;
  A.dt = p.dt; A.nucR = p.nucR; A.nucT = p.nucT; A.nucRuptVel = p.nucRuptVel; A.nucdtau0 = p.nucdtau0;
  A.xsource = p.xsource; A.ysource = p.ysource; A.zsource = p.zsource;
  A.slipRateThres = p.slipRateThres; A.tol = p.tol; A.fric_tp_h = p.fric_tp_h;
  A.friclaw = p.friclaw; A.C_nuclea = p.C_nuclea; A.nucfault = p.nucfault; A.TPV = p.TPV;
  A.insertFaultType = p.insertFaultType; A.C_elastic = p.C_elastic;
  return A;
}

namespace {

// ---------------------------------------------------------------------------
// face dof / node lists in the order of processNodalQuantArr
// (assembleGlobalMass.f90:141-188): x-face (iz,iy), y-face (ix,iz), z-face
// (ix,iy), then the face's split-node masters.
void face_nodes(const eqd_handle& h, int a, int side, std::vector<int>& out) {
  out.clear();
  const int nx = h.numcount[0], ny = h.numcount[1], nz = h.numcount[2];
  const int n[3] = {nx, ny, nz};
  const int b = side == 0 ? 1 : n[a];
  auto id = [&](int ix, int iy, int iz) { return (ix - 1) * ny * nz + (iz - 1) * ny + iy - 1; };
  if (a == 0) {
    for (int iz = 1; iz <= nz; ++iz) for (int iy = 1; iy <= ny; ++iy) out.push_back(id(b, iy, iz));
  } else if (a == 1) {
    for (int ix = 1; ix <= nx; ++ix) for (int iz = 1; iz <= nz; ++iz) out.push_back(id(ix, b, iz));
  } else {
    for (int ix = 1; ix <= nx; ++ix) for (int iy = 1; iy <= ny; ++iy) out.push_back(id(ix, iy, b));
  }
  if (h.fltMPI[2 * a + side])
    for (int k : h.fltface[2 * a + side]) out.push_back(nx * ny * nz + k - 1);
}

void p2p_setup(eqd_handle* h);

void comm_ready(eqd_handle* h) {
  if (h->commThread.joinable()) {
    h->commThread.join();
    if (h->commRc != ncclSuccess) { h->comm = nullptr; throw CudaError(std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(h->commRc)); }
  }
}

void finalize(eqd_handle* h) {
  if (h->finalized) return;
  need(h->meshSet && h->opsSet && h->nodalSet, "eqd_run: eqd_set_mesh, eqd_set_elem_ops and eqd_set_nodal must be called first");
  const eqd_params& P = h->p;
  const int Nn = h->Nn;
  // ---- special nodes: split-node pairs + active rank faces
  std::vector<char> special(Nn, 0);
  for (int q = 0; q < h->nPairs; ++q) { special[h->pairNodeS[q]] = 1; special[h->pairNodeM[q]] = 1; }
  const int np[3] = {P.npx, P.npy, P.npz};
  const int me = P.me;
  const int mex = me / (P.npy * P.npz), mey = (me - mex * P.npy * P.npz) / P.npz, mez = me - mex * P.npy * P.npz - mey * P.npz;
  const int mexyz[3] = {mex, mey, mez};
  const int stride[3] = {P.npy * P.npz, P.npz, 1};
  const size_t NS = h->NnS, PSn = h->NpS;
  if (h->haloSet) {
    for (int a = 0; a < 3; ++a)
      for (int side = 0; side < 2; ++side) {
        Face& F = h->face[a][side];
        const bool active = np[a] > 1 && (side == 0 ? mexyz[a] != 0 : mexyz[a] != np[a] - 1);
        if (!active) { F.nb = -1; F.n = 0; continue; }
        F.nb = me + (side == 0 ? -stride[a] : stride[a]);
        face_nodes(*h, a, side, F.nodes);
        std::vector<uint32_t> idx;
        for (int n : F.nodes) {
          need(n >= 0 && n < Nn, "eqd_set_halo: face node out of range");
          const int kind = EQD_INFO_KIND(h->info[n]);
          if (kind == KIND_FIXED) continue;
          special[n] = 1;
          if (kind == KIND_FREE3) for (int j = 0; j < 3; ++j) idx.push_back((uint32_t)(j * NS + n));
          else { const size_t slot = EQD_INFO_SLOT(h->info[n]); for (int j = 0; j < 12; ++j) idx.push_back((uint32_t)(3 * NS + j * PSn + slot)); }
        }
        // cross-check with the reference's own count (numcount(4:9) + 3*fltnum)
        const int expect = h->numcount[3 + 2 * a + side] + (h->fltMPI[2 * a + side] ? 3 * h->fltnum[2 * a + side] : 0);
        if ((int)idx.size() != expect)
          throw ArgError("halo face dof count differs from numcount/fltnum: " + std::to_string(idx.size()) + " vs " + std::to_string(expect));
        F.n = (int)idx.size();
        F.idx.upload(idx);
        F.send.alloc(F.n); F.recv.alloc(F.n);
        // pairs on this face (for the init-time arn exchange, MPI4arn)
        F.pairs.clear();
        if (h->fltMPI[2 * a + side])
          for (int k : h->fltface[2 * a + side]) {
            // pair (k, fault ntotft) -> device pair slot
            const int ref = (k - 1) + h->nftmx * (h->ntotft - 1);
            int slot = -1;
            for (int q = 0; q < h->nPairs; ++q) if (h->pairRef[q] == ref) { slot = q; break; }
            need(slot >= 0, "eqd_set_halo: face pair not found");
            F.pairs.push_back(slot);
          }
      }
    for (int a = 0; a < 3; ++a) CK(cudaEventCreateWithFlags(&h->evPacked[a], cudaEventDisableTiming));
  }
  p2p_setup(h);
  // ---- launch order of the tiles: those touching an active rank face first
  std::vector<char> onFace(Nn, 0);
  bool& anyFace = h->anyFace;
  anyFace = false;
  if (h->haloSet)
    for (int a = 0; a < 3; ++a)
      for (int side = 0; side < 2; ++side)
        if (h->face[a][side].nb >= 0)
          for (int n : h->face[a][side].nodes) { onFace[n] = 1; anyFace = true; }
  // ---- option "box": flag the tiles (1: regular classes, 2: also PML) whose elements are all
  // axis-aligned hexahedra (exact test on the reference's coordinates, eqd_box.h)
  if (h->optBox)
    for (int c = 0; c < 3; ++c) {
      ElemClass& C = h->cls[c];
      C.nBoxElems = 0;
      if (!C.n || (c == CLS_PML && h->optBox < 2)) continue;
      C.tileBoxH.assign(C.nTiles, 0);
      parallel_range((size_t)C.nTiles, [&](size_t tb, size_t te) {
        for (size_t t = tb; t < te; ++t) {
          const int e0 = C.tileRecH[t].x, ne = C.tileRecH[t].y & 0xffff;
          bool all = true;
          for (int k = 0; k < ne && all; ++k) {
            const int e = C.refId[(size_t)e0 + k];
            all = e >= 0 && h->etype[e] != 11 && h->etype[e] != 12 && box_element(&h->conn[8 * (size_t)e], h->coor.data());
          }
          C.tileBoxH[t] = all ? 1 : 0;
        }
      });
      for (int t = 0; t < C.nTiles; ++t) if (C.tileBoxH[t]) C.nBoxElems += C.tileRecH[t].y & 0xffff;
      if (!anyFace) C.tileBox.upload(C.tileBoxH);
    }
  if (anyFace) {
    for (int c = 0; c < 3; ++c) {
      ElemClass& C = h->cls[c];
      if (!C.n) continue;
      std::vector<int4> first, rest;
      std::vector<uint8_t> firstB, restB;
      const bool boxed = !C.tileBoxH.empty();
      for (int t = 0; t < C.nTiles; ++t) {
        bool f = false;
        for (int k = C.tileNodeH[t]; k < C.tileNodeH[t + 1] && !f; ++k) f = C.tnodeH[k] >= 0 && onFace[C.tnodeH[k]];
        (f ? first : rest).push_back(C.tileRecH[t]);
        if (boxed) (f ? firstB : restB).push_back(C.tileBoxH[t]);
      }
      C.nFaceTiles = (int)first.size();
      first.insert(first.end(), rest.begin(), rest.end());
      C.tileRec.upload(first);
      if (boxed) { firstB.insert(firstB.end(), restB.begin(), restB.end()); C.tileBox.upload(firstB); }
    }
  }
  {
    // communication stream (halo + fault solver); a blocking stream like the main one
    int lo = 0, hi = 0;
    CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CK(cudaStreamCreateWithPriority(&h->commStream, cudaStreamDefault, hi));
    for (cudaEvent_t* e : {&h->evFace, &h->evElem, &h->evComm}) CK(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    CK(cudaDeviceGetAttribute(&h->smCount, cudaDevAttrMultiProcessorCount, h->device));
  }
  std::vector<int> list, listA, listB;
  for (int n = 0; n < Nn; ++n)
    if (special[n] && EQD_INFO_KIND(h->info[n]) != KIND_FIXED) {
      h->info[n] |= 4;
      list.push_back(n);
      (onFace[n] ? listA : listB).push_back(n);
    }
  if (h->cls[CLS_MARCH].n || h->cls[CLS_MARCHP].n) {
    // a node a bundle updates itself is interior to the bundle: no split node (the planner cuts where elements stop
    // sharing nodes), no rank-face node (a face bounds the sub-domain's elements)
    for (int n : list)
      if (EQD_INFO_KIND(h->info[n]) == KIND_FREE3 && EQD_INFO_FUSED(h->info[n])) throw ArgError("internal: node " + std::to_string(n + 1) + " is both fused and special");
    // the interior work list is swept while the face forces travel: none of its strips may touch a rank-face node
    {
      const raw_vector<int>& tn = h->cls[CLS_MARCH].tnodeH;
      std::vector<char> hit(1, 0);
      parallel_range(tn.size() - std::min(tn.size(), (size_t)h->mSlotB0), [&](size_t b, size_t e) {
        for (size_t k = b + h->mSlotB0; k < e + h->mSlotB0; ++k) if (tn[k] >= 0 && onFace[tn[k]]) hit[0] = 1;
      });
      const raw_vector<int>& tp = h->cls[CLS_MARCHP].tnodeH;
      if (h->cls[CLS_MARCHP].n)
        parallel_range(tp.size() - std::min(tp.size(), (size_t)h->pSlotB0), [&](size_t b, size_t e) {
          for (size_t k = b + h->pSlotB0; k < e + h->pSlotB0; ++k) if (tp[k] >= 0 && onFace[tp[k]]) hit[0] = 1;
        });
      h->mSplitOk = !hit[0];   // (a lattice the boundary criterion of the planner does not fit: sweep everything before the exchange)
    }
    std::vector<int> rest;
    rest.reserve((size_t)Nn - (size_t)h->mFused);
    for (int n = 0; n < Nn; ++n)
      if (EQD_INFO_KIND(h->info[n]) == KIND_FREE3 && !EQD_INFO_FUSED(h->info[n])) rest.push_back(n);
    h->nNodeList = (int)rest.size();
    h->dNodeList.upload(rest);
  }
  h->nSpecial = (int)list.size();
  h->dSpecial.upload(list);
  h->nSpecialA = (int)listA.size(); h->nSpecialB = (int)listB.size();
  h->dSpecialA.upload(listA); h->dSpecialB.upload(listB);
  h->dInfo.upload(h->info);
  // ---- stations
  std::vector<int> pairStation(std::max(h->PS, 1), -1);
  for (int j = 0; j < h->nOn; ++j) {
    const int i = h->anonfsH[3 * j], ift = h->anonfsH[3 * j + 2];
    const int ref = (i - 1) + h->nftmx * (ift - 1);
    for (int q = 0; q < h->nPairs; ++q)
      if (h->pairRef[q] == ref) pairStation[q] = j;  // the device records one row per pair: the last matching station's
  }
  // storeOnFaultStationQuantSCEC (faulting.f90:524-539) writes the record of a pair into the slot of EVERY station
  // that snapped to it; stations sharing a pair get a copy of the recorded row when the history is fetched
  h->stationAlias.assign(std::max(h->nOn, 1), 0);
  for (int j = 0; j < h->nOn; ++j) {
    const int ref = (h->anonfsH[3 * j] - 1) + h->nftmx * (h->anonfsH[3 * j + 2] - 1);
    h->stationAlias[j] = j;
    for (int q = 0; q < h->nPairs; ++q)
      if (h->pairRef[q] == ref && pairStation[q] >= 0) h->stationAlias[j] = pairStation[q];
  }
  h->dPairStation.upload(pairStation);
  h->nOnAlloc = std::max(h->nOn, 1);
  const size_t nstep = (size_t)std::max(P.nstep, 1);
  h->dOnHist.alloc(12 * nstep * h->nOnAlloc);
  if (h->nOff > 0) h->dOffHist.alloc((size_t)(6 * h->nOff + 1) * nstep);
  h->dHypo.alloc(13 * nstep);
  if (P.friclaw == 5 && h->nPairs > 0) h->dTpHist.alloc(2 * nstep * h->PS);
  if (P.outputGroundMotion) {
    h->nGm = 0;
    const size_t ns = nstep / 10 + 1;
    if (h->nSurf > 0) h->dGm.alloc(3 * (size_t)h->nSurf * ns);
    if (h->nPairs > 0) h->dSrc.alloc((size_t)h->nftnd[0] * ns);
  }
  std::vector<StepState> st(1);
  st[0].timeElapsed = 0.0; st[0].nt = 0; st[0].nanFlag = 0; st[0].nanNode = 0; st[0].pad = 0;
  h->dState.upload(st);
  h->hostNt = 0; h->hostTime = 0.0;
  // cudaMemset / cudaMemcpy above ran on the legacy stream, which does not order
  // against the non-blocking step stream
  CK(cudaDeviceSynchronize());
  // small sub-domains give their pinned staging buffers back; large ones keep them for the staged
  // reads of eqd_fetch (re-pinning 256 MB costs more than the reads)
  if ((size_t)h->Nn * 3 * sizeof(double) < (64u << 20)) {
    if (t_stage == h->stage.get()) { t_stage = nullptr; t_stageStream = nullptr; }  // finalize may run inside a guarded call
    h->stage.reset();
  }
  h->finalized = true;
}

// transport of one axis phase between processes (NCCL) -- pack/unpack by caller
void halo_axis_nccl(eqd_handle* h, int a, cudaStream_t st) {
  comm_ready(h);
  if (!h->comm)
    throw ArgError(h->hostAg ? "eqd_run: peer memory (CUDA IPC) cannot be mapped between the ranks and no NCCL communicator was given (eqd_set_comm)"
                             : "eqd_run: sub-domain has neighbours but eqd_set_comm was not called");
  NK(g_nccl.GroupStart());
  for (int side = 0; side < 2; ++side) {
    Face& F = h->face[a][side];
    if (F.nb < 0 || F.n == 0) continue;
    NK(g_nccl.Send(F.send.p, F.n, ncclDouble, F.nb, h->comm, st));
    NK(g_nccl.Recv(F.recv.p, F.n, ncclDouble, F.nb, h->comm, st));
  }
  NK(g_nccl.GroupEnd());
}

bool has_neighbours(const eqd_handle* h) {
  if (!h->haloSet) return false;
  for (int a = 0; a < 3; ++a) for (int s = 0; s < 2; ++s) if (h->face[a][s].nb >= 0) return true;
  return false;
}

struct Timer {
  eqd_handle* h; int slot; cudaEvent_t a = nullptr, b = nullptr; cudaStream_t s;
  Timer(eqd_handle* h_, int slot_, cudaStream_t s_ = nullptr) : h(h_), slot(slot_), s(s_ ? s_ : h_->stream) {
    if (!h->timing) return;
    if (h->evPoolUsed + 2 > h->evPool.size()) {
      const size_t grow = std::max<size_t>(64, h->evPool.size());
      for (size_t k = 0; k < grow; ++k) { cudaEvent_t e; CK(cudaEventCreate(&e)); h->evPool.push_back(e); }
    }
    a = h->evPool[h->evPoolUsed++]; b = h->evPool[h->evPoolUsed++];
    CK(cudaEventRecord(a, s));
  }
  void stop() {
    if (!h->timing) return;
    CK(cudaEventRecord(b, s));
    h->evs.push_back(a); h->evs.push_back(b); h->evs.push_back((cudaEvent_t)(intptr_t)slot);
  }
};

void collect_timing(eqd_handle* h) {
  for (size_t i = 0; i + 3 <= h->evs.size(); i += 3) {
    float ms = 0;
    cudaEventElapsedTime(&ms, h->evs[i], h->evs[i + 1]);
    h->tms[(int)(intptr_t)h->evs[i + 2]] += ms;
  }
  h->evs.clear();
  h->evPoolUsed = 0;
}

void halo_pack(eqd_handle* h, int a, cudaStream_t st = nullptr) {
  if (!st) st = h->stream;
  for (int side = 0; side < 2; ++side) {
    Face& F = h->face[a][side];
    if (F.nb < 0 || F.n == 0) continue;
    launch_pack(h->dForce.p, F.idx.p, F.n, F.send.p, st); h->launches++;
  }
}
void halo_unpack(eqd_handle* h, int a, cudaStream_t st = nullptr) {
  if (!st) st = h->stream;
  for (int side = 0; side < 2; ++side) {
    Face& F = h->face[a][side];
    if (F.nb < 0 || F.n == 0) continue;
    launch_unpack_add(h->dForce.p, F.idx.p, F.n, F.recv.p, st); h->launches++;
  }
}
// Peer-memory set-up of the step-loop exchange (collective: every rank of the communicator calls it once, from
// finalize).  Each rank lays its receive buffers, flag words and send counters out in one allocation, publishes the
// allocation's IPC handle and the layout through one ncclAllGather, and maps its neighbours' allocations.
// set-up all-gather of `bytes` (a multiple of 8) per rank: the host's communicator when eqd_set_host_comm gave one,
// else NCCL through device memory
void setup_allgather(eqd_handle* h, const void* send, size_t bytes, void* recv) {
  if (h->hostAg) {
    if (h->hostAg(h->hostCtx, send, (int64_t)bytes, recv) != 0) throw CudaError("eqd_set_host_comm: the host's all-gather reported a failure");
    return;
  }
  DevBuf<unsigned long long> dMine, dAll;
  dMine.alloc(bytes / 8); dAll.alloc(bytes / 8 * h->nranks);
  CK(cudaMemcpy(dMine.p, send, bytes, cudaMemcpyHostToDevice));
  NK(g_nccl.AllGather(dMine.p, dAll.p, bytes / 8, ncclUint64, h->comm, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaMemcpy(recv, dAll.p, bytes * h->nranks, cudaMemcpyDeviceToHost));
}

struct P2PRecord { cudaIpcMemHandle_t handle; unsigned long long recvOff[3][2][2]; unsigned long long flagOff; unsigned long long pad[2]; };
void p2p_setup(eqd_handle* h) {
  auto& Q = h->p2p;
  Q.on = false;
  if (!h->hostAg) comm_ready(h);
  if ((!h->comm && !h->hostAg) || h->optHalo == 0 || h->nranks < 2) return;
  static_assert(sizeof(P2PRecord) % 8 == 0, "record is exchanged as 8-byte words");
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
  for (int a = 0; a < 3; ++a)
    for (int s = 0; s < 2; ++s)
      for (int par = 0; par < 2; ++par) Q.recvOff[a][s][par] = take(sizeof(double) * (size_t)std::max(h->face[a][s].n, 1));
  Q.flagOff = take(sizeof(unsigned) * 12);      // [axis][side][parity]
  Q.counterOff = take(sizeof(unsigned) * 6);    // [axis][side]
  Q.area.alloc(off);                            // zeroed: flags 0, counters 0
  P2PRecord mine{};
  bool ok = cudaIpcGetMemHandle(&mine.handle, Q.area.p) == cudaSuccess;
  for (int a = 0; a < 3; ++a) for (int s = 0; s < 2; ++s) for (int par = 0; par < 2; ++par) mine.recvOff[a][s][par] = Q.recvOff[a][s][par];
  mine.flagOff = Q.flagOff;
  mine.pad[0] = ok ? 1 : 0;
  std::vector<P2PRecord> all(h->nranks);
  setup_allgather(h, &mine, sizeof mine, all.data());
  for (const P2PRecord& r : all) ok = ok && r.pad[0] == 1;
  Q.peerBase.assign(h->nranks, nullptr);
  for (int a = 0; a < 3 && ok; ++a)
    for (int s = 0; s < 2 && ok; ++s) {
      const Face& F = h->face[a][s];
      if (F.nb < 0) continue;
      if (!Q.peerBase[F.nb]) {
        void* base = nullptr;
        if (cudaIpcOpenMemHandle(&base, all[F.nb].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = false; break; }
        Q.peerBase[F.nb] = base;
      }
      // my -side face talks to the neighbour's +side face of the same axis, and vice versa
      for (int par = 0; par < 2; ++par) Q.peerRecvOff[a][s][par] = all[F.nb].recvOff[a][1 - s][par];
      Q.peerFlagOff[a][s] = all[F.nb].flagOff + sizeof(unsigned) * (size_t)((a * 2 + (1 - s)) * 2);
    }
  // the choice must be the same on every rank: one more tiny all-gather of the outcome
  unsigned long long okw = ok ? 1 : 0;
  std::vector<unsigned long long> oks(h->nranks);
  setup_allgather(h, &okw, 8, oks.data());
  for (unsigned long long v : oks) ok = ok && v == 1;
  if (!ok) {
    for (void*& b : Q.peerBase) if (b) { cudaIpcCloseMemHandle(b); b = nullptr; }
    if (h->optHalo == 1) throw CudaError("option halo = 1: peer memory (CUDA IPC) between the ranks is not available");
    if (getenv("EQD_VERBOSE")) fprintf(stderr, "[eqd] rank %d: peer memory not available, step-loop exchange over ncclSend/ncclRecv\n", h->rank);
    return;
  }
  Q.on = true;
  Q.seq = 0;
  if (getenv("EQD_VERBOSE")) fprintf(stderr, "[eqd] rank %d: step-loop exchange over peer memory (CUDA IPC), %zu bytes of receive area\n", h->rank, off);
}

// the three axis phases over peer memory: pack into the neighbour's buffer + flag, wait for the own flag + add
void halo_all_p2p(eqd_handle* h, cudaStream_t st) {
  auto& Q = h->p2p;
  Timer t(h, EQD_T_HALO, st);
  const unsigned seq = ++Q.seq;
  const int par = (int)(seq & 1u);
  unsigned char* base = Q.area.p;
  for (int a = 0; a < 3; ++a) {
    if (h->face[a][0].nb < 0 && h->face[a][1].nb < 0) continue;
    HaloAxisArgs A{};
    for (int s = 0; s < 2; ++s) {
      const Face& F = h->face[a][s];
      A.n[s] = F.nb >= 0 ? F.n : 0;
      if (F.nb < 0) continue;
      A.idx[s] = F.idx.p;
      unsigned char* peer = (unsigned char*)Q.peerBase[F.nb];
      A.remoteRecv[s] = (double*)(peer + Q.peerRecvOff[a][s][par]);
      A.remoteFlag[s] = (unsigned*)(peer + Q.peerFlagOff[a][s]) + par;
      A.localRecv[s] = (const double*)(base + Q.recvOff[a][s][par]);
      A.localFlag[s] = (const unsigned*)(base + Q.flagOff) + (a * 2 + s) * 2 + par;
      A.counter[s] = (unsigned*)(base + Q.counterOff) + a * 2 + s;
    }
    A.force = h->dForce.p; A.seq = seq;
    launch_halo_send(A, st); launch_halo_recv(A, st); h->launches += 2;
  }
  t.stop();
}

// the three axis phases of MPI4NodalQuant between processes, in order, on stream st
void halo_all_nccl(eqd_handle* h, cudaStream_t st) {
  if (h->p2p.on) { halo_all_p2p(h, st); return; }
  Timer t(h, EQD_T_HALO, st);
  for (int a = 0; a < 3; ++a) {
    if (h->face[a][0].nb < 0 && h->face[a][1].nb < 0) continue;
    halo_pack(h, a, st);
    halo_axis_nccl(h, a, st);
    halo_unpack(h, a, st);
  }
  t.stop();
}

// Everything of a step that follows the element sweep and needs the summed
// face forces: thermop + faulting (driver.f90:23-28) and the source-evolution
// sample (driver.f90:30-33), on stream st.
void launch_fault_phase(eqd_handle* h, cudaStream_t st) {
  const eqd_params& P = h->p;
  Timer t(h, EQD_T_FAULT, st);
  if (h->nPairs > 0) {
    FaultArgs A = h->faultArgs();
    if (P.friclaw == 5) { launch_thermop(A, st); h->launches++; }
    launch_fault(A, st); h->launches++;
  }
  if (P.outputGroundMotion && (h->hostNt + 1) % 10 == 1 && h->dSrc.p) {  // driver.f90:30-33: mod(nt,10) == 1
    launch_sample_src(h->dFric.p, h->PS, h->nftnd[0], h->dSrc.p + (size_t)h->nftnd[0] * h->nGm, st); h->launches++;
  }
  t.stop();
}

// One time step (driver.f90:9-34) as a launch sequence.
//   ov == 0  serial: everything on the main stream.  `external` = the caller does
//            the halo between step_pre and step_post (in-process groups).
//   ov >= 1  the halo and the fault solver run on commStream and are joined only
//            by the special-node update of the NEXT step: they hide under that
//            step's bulk node update, which touches no split-node or rank-face
//            node (those are updated by k_node_update_special after the join).
//   ov == 2  additionally the tiles touching a rank face are swept first and the
//            halo starts while the interior tiles are still being swept.
//   last     the last step of this eqd_run call: the marching kernel then leaves the complete force of the
//            nodes it would update itself in force[] instead, so that the caller sees v(nt), d(nt) and
//            f/m as the reference leaves them; the next call's first node update picks them up there.
void step_pre(eqd_handle* h, int ov, bool multi, bool last) {
  const eqd_params& P = h->p;
  cudaStream_t s = h->stream, c = h->commStream;
  const bool marching = h->cls[CLS_MARCH].n > 0 || h->cls[CLS_MARCHP].n > 0;
  {
    Timer t(h, EQD_T_NODE);
    NodeArgs A = h->nodeArgs();
    if (h->dAccel0.p) A.accel0 = h->dAccel0.p;
    A.skipSpecial = ov ? 1 : 0;
    if (h->cls[CLS_MARCH].n && h->sweepFused) { A.list = h->dNodeList.p; A.nList = h->nNodeList; A.fusedMode = 2; }
    launch_node_update(A, s); h->launches += h->Np > 0 ? 2 : 1;
    t.stop();
    if (ov) {
      // join: halo + fault solver of the previous step
      CK(cudaStreamWaitEvent(s, h->evComm, 0));
    }
    launch_advance(h->dState.p, P.dt, s); h->launches++;
    Timer t2(h, EQD_T_NODE);
    if (ov && h->nSpecial) { launch_node_update_special(A, h->dSpecial.p, h->nSpecial, s); h->launches++; }
    if (h->pingpong) h->cur ^= 1;   // every node's v(nt), d(nt) now sits in the other buffer: the sweep reads that one
    if (h->nOff > 0) {
      launch_store_offfault(h->dIdhist.p, 6 * h->nOff, h->dOffHist.p, h->velCur().p, h->dispCur().p, h->NnS, h->dState.p, s);
      h->launches++;
    }
    if (P.outputGroundMotion && (h->hostNt + 1) % 10 == 1 && h->dGm.p) {
      // output_gm (driver.f90:30-33): velArr of this step, before the element sweep moves the fused nodes on
      launch_sample_gm(h->dSurf.p, h->nSurf, h->velCur().p, h->NnS, h->dGm.p + 3 * (size_t)h->nSurf * h->nGm, s); h->launches++;
    }
    t2.stop();
  }
  if (h->dAccel0.p) { CK(cudaStreamSynchronize(s)); h->dAccel0.release(); }
  auto sweep = [&](int part) {   // 0: all tiles, 1: rank-face tiles, 2: interior tiles
    if (h->cls[CLS_MARCH].n) {
      Timer t(h, EQD_T_MARCH);
      launch_march(h->marchArgs(!last, part), h->mGrid, s); h->launches++;
      t.stop();
    }
    if (h->cls[CLS_MARCHP].n) {
      Timer t(h, EQD_T_MARCH_PML);
      launch_march_pml(h->marchPmlArgs(part), h->pGrid, s); h->launches++;
      t.stop();
    }
    for (int k = 0; k < 3; ++k) {
      ElemClass& C = h->cls[k];
      if (!C.n) continue;
      ElemArgs A = h->elemArgs(k);
      if (part == 1) { A.tile0 = 0; A.ntiles = C.nFaceTiles; }
      if (part == 2) { A.tile0 = C.nFaceTiles; A.ntiles = C.nTiles - C.nFaceTiles; A.maxGrid = std::max(2 * h->smCount - h->optReserve, 2); }
      if (A.ntiles <= 0) continue;
      Timer t(h, k == CLS_REG ? EQD_T_ELEM : k == CLS_REGX ? EQD_T_ELEM_REGX : EQD_T_ELEM_PML);
      if (k == CLS_PML) launch_elem_pml(A, h->body, P.C_hg, s);
      else launch_elem_reg(A, k == CLS_REGX, h->plastic, h->qmode, h->body, P.C_hg, s);
      h->launches++;
      t.stop();
    }
  };
  auto assemble = [&](const DevBuf<int>& list, int n) {
    Timer t(h, EQD_T_ASSEMBLE);
    if (n) { launch_assemble_special(h->nodeArgs(), list.p, n, s); h->launches++; }
    t.stop();
  };
  const bool split = ov == 2 && multi && h->anyFace && (!marching || h->mSplitOk);
  if (!split) {
    sweep(0);
    assemble(h->dSpecial, h->nSpecial);
  } else {
    sweep(1);
    assemble(h->dSpecialA, h->nSpecialA);
    CK(cudaEventRecord(h->evFace, s));
    CK(cudaStreamWaitEvent(c, h->evFace, 0));
    halo_all_nccl(h, c);
    sweep(2);
    assemble(h->dSpecialB, h->nSpecialB);
  }
  h->sweepFused = h->cls[CLS_MARCH].n > 0 && !last;
  if (ov) {
    CK(cudaEventRecord(h->evElem, s));
    CK(cudaStreamWaitEvent(c, h->evElem, 0));
    if (multi && !split) halo_all_nccl(h, c);
  }
}

void step_post(eqd_handle* h, int ov) {
  const eqd_params& P = h->p;
  cudaStream_t s = h->stream;
  launch_fault_phase(h, ov ? h->commStream : s);
  if (ov) CK(cudaEventRecord(h->evComm, h->commStream));
  h->hostNt++;
  h->hostTime = h->hostTime + P.dt;
  if (P.outputGroundMotion && h->hostNt % 10 == 1) h->nGm++;   // the samples of this step: step_pre (gm), launch_fault_phase (src_evol)
}

void prepare_run(eqd_handle* h, int nt_begin) {
  finalize(h);
  if (nt_begin != h->hostNt + 1) {
    // jump: rebuild the accumulated time exactly as repeated `timeElapsed + dt`
    double t = 0.0;
    for (int k = 1; k < nt_begin; ++k) t = t + h->p.dt;
    StepState st{};
    CK(cudaMemcpy(&st, h->dState.p, sizeof st, cudaMemcpyDeviceToHost));
    st.nt = nt_begin - 1; st.timeElapsed = t;
    CK(cudaMemcpy(h->dState.p, &st, sizeof st, cudaMemcpyHostToDevice));
    h->hostNt = nt_begin - 1; h->hostTime = t;
  }
}

int finish_run(eqd_handle* h) {
  if (h->commStream) CK(cudaStreamSynchronize(h->commStream));
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaGetLastError());
  if (h->timing) collect_timing(h);
  StepState st{};
  CK(cudaMemcpy(&st, h->dState.p, sizeof st, cudaMemcpyDeviceToHost));
  if (st.nanFlag) {
    h->err = "NaN velocity at node " + std::to_string(st.nanNode) + " (driver.f90:147-152), step <= " + std::to_string(st.nt);
    return EQD_ERR_NAN;
  }
  return EQD_OK;
}

// init-time shared sums between in-process handles or over NCCL use this generic
// "exchange doubles per face" helper: vals[a][side] in/out
}  // namespace

extern "C" {

int eqd_create(const eqd_params* p, int device, eqd_handle** out) {
  if (!p || !out) return EQD_ERR_ARG;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    fprintf(stderr, "eqdyna_b200: no CUDA device (%s); the step library has no CPU path\n", cudaGetErrorString(e));
    return EQD_ERR_CUDA;
  }
  eqd_handle* h = new eqd_handle();
  h->p = *p;
  h->device = device < 0 ? (p->me % ndev) : device;
  int rc = guarded(h, [&] {
    need(h->device < ndev, "eqd_create: device index out of range");
    need(p->dt > 0 && p->npx > 0 && p->npy > 0 && p->npz > 0, "eqd_create: bad dt / decomposition");
    need(!(p->C_elastic == 0 && p->C_Q == 1), "Q model can only work with elastic code (warning.f90:6-9)");
    need(p->C_hg == 1 || p->C_hg == 2, "eqd_create: C_hg must be 1 or 2");
    // A *blocking* stream on purpose: the setup path uses plain cudaMemcpy /
    // cudaMemset on the legacy stream (a pageable H2D copy may return before its
    // DMA has landed), and a blocking stream is ordered after those.
    CK(cudaStreamCreate(&h->stream));
    // L2 -> DRAM fetch size.  The node kernels that follow a node list (perimeter nodes of the marching bundles, rank
    // faces, split nodes) read single 8-byte values a row of the lattice apart; with the default of 64 B every such
    // read pulls two sectors from DRAM (ncu, r02_j: 2.0 L2 sectors per L1 sector in k_node_update3's list variant,
    // 1.2 in the streaming kernels).  EQD_L2_FETCH = 32 / 64 / 128 sets the device limit (a hint; per context).
    if (const char* e = getenv("EQD_L2_FETCH")) {
      const int v = atoi(e);
      if (v == 32 || v == 64 || v == 128) {
        if (cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)v) != cudaSuccess) cudaGetLastError();
      }
      size_t now = 0;
      if (cudaDeviceGetLimit(&now, cudaLimitMaxL2FetchGranularity) != cudaSuccess) cudaGetLastError();
      if (getenv("EQD_VERBOSE")) fprintf(stderr, "[eqd] L2 fetch granularity: asked %d, device reports %zu\n", v, now);
    }
    h->stage = shared_stage(h->device);
    h->body = (p->C_elastic == 0) || (p->rdampm != 0.0);
    h->plastic = p->C_elastic == 0;
    h->qmode = p->C_Q == 1;
    if (h->qmode) {
      QTab tab[16];
      for (int depth = 0; depth < 2; ++depth)
        for (int k = 1; k <= 8; ++k) {
          // calcElemKU.f90:86-93: Qs/Qp by depth
          const double Qs = depth == 0 ? 10.0 : 50.0, Qp = depth == 0 ? 20.0 : 100.0;
          double taok, wkp, wks, cv, cs;
          qconstant(Qp, k, taok, wkp, cv);
          qconstant(Qs, k, taok, wks, cs);
          QTab& t = tab[depth * 8 + k - 1];
          t.taok = taok; t.wkp = wkp * 8.0; t.wks = wks * 8.0; t.cv = cv; t.cs = cs; t.expdt = std::exp(-p->dt / taok);
        }
      upload_qtab(tab);
    }
  });
  if (rc != EQD_OK) { fprintf(stderr, "eqd_create: %s\n", h->err.c_str()); delete h; return rc; }
  *out = h;
  return EQD_OK;
}

int eqd_destroy(eqd_handle* h) {
  if (!h) return EQD_OK;
  cudaSetDevice(h->device);
  if (h->commThread.joinable()) h->commThread.join();
  for (void* b : h->p2p.peerBase) if (b) cudaIpcCloseMemHandle(b);
  if (h->p2p.area.p) {
    // a neighbour may still have this allocation mapped (freeing an exported allocation before every importer has closed
    // it is undefined): the few MB stay allocated until the process ends
    static std::mutex mu;
    static std::vector<void*> keep;
    std::lock_guard<std::mutex> g(mu);
    keep.push_back(h->p2p.area.p);
    h->p2p.area.p = nullptr; h->p2p.area.n = 0;
  }
  if (h->comm && g_nccl.lib) g_nccl.CommDestroy(h->comm);
  for (int a = 0; a < 3; ++a) if (h->evPacked[a]) cudaEventDestroy(h->evPacked[a]);
  for (cudaEvent_t e : {h->evFace, h->evElem, h->evComm}) if (e) cudaEventDestroy(e);
  for (cudaEvent_t e : h->evPool) cudaEventDestroy(e);
  if (h->commStream) cudaStreamDestroy(h->commStream);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return EQD_OK;
}

int eqd_last_error(const eqd_handle* h, char* buf, int n) {
  if (!h || !buf || n <= 0) return EQD_ERR_ARG;
  snprintf(buf, n, "%s", h->err.c_str());
  return EQD_OK;
}

int eqd_set_mesh(eqd_handle* h, int32_t Nn, int32_t Ne, int32_t Neq, int32_t sizeEq, const double* meshCoor,
                 const int32_t* nodeElemIdRelation, const int32_t* elemTypeArr, const int32_t* numOfDofPerNodeArr,
                 const int32_t* eqNumStartIndexLoc, const int32_t* eqNumIndexArr, const int32_t* stressCompIndexArr,
                 int32_t sizeStress) {
  if (!h) return EQD_ERR_ARG;
  return guarded(h, [&] {
    need(Nn > 0 && Ne > 0 && meshCoor && nodeElemIdRelation && elemTypeArr && numOfDofPerNodeArr && eqNumStartIndexLoc &&
             eqNumIndexArr && stressCompIndexArr, "eqd_set_mesh: null / empty argument");
    const eqd_params& P = h->p;
    Lap lap("eqd_set_mesh");
    h->Nn = Nn; h->Ne = Ne; h->Neq = Neq; h->sizeEq = sizeEq; h->sizeStress = sizeStress;
    h->coor.resize(3 * (size_t)Nn); h->etype.resize(Ne); h->ndof.resize(Nn); h->eqStart.resize(Nn);
    h->eqIdx.resize(sizeEq); h->stressIdx.resize(Ne); h->conn.resize(8 * (size_t)Ne);
    parallel_memcpy(h->coor.data(), meshCoor, sizeof(double) * 3 * (size_t)Nn);
    parallel_memcpy(h->etype.data(), elemTypeArr, sizeof(int) * (size_t)Ne);
    parallel_memcpy(h->ndof.data(), numOfDofPerNodeArr, sizeof(int) * (size_t)Nn);
    parallel_memcpy(h->eqStart.data(), eqNumStartIndexLoc, sizeof(int) * (size_t)Nn);
    parallel_memcpy(h->eqIdx.data(), eqNumIndexArr, sizeof(int) * (size_t)sizeEq);
    parallel_memcpy(h->stressIdx.data(), stressCompIndexArr, sizeof(int) * (size_t)Ne);
    {
      std::vector<char> badv(host_threads() + 1, 0);
      parallel_range(8 * (size_t)Ne, [&](size_t b, size_t e) {
        bool bad = false;
        for (size_t k = b; k < e; ++k) {
          const int n = nodeElemIdRelation[k] - 1;
          bad |= (n < 0 || n >= Nn);
          h->conn[k] = n;
        }
        if (bad) badv[0] = 1;
      });
      need(!badv[0], "eqd_set_mesh: connectivity out of range");
    }
    lap.lap("host copies");
    // ---- node kinds
    h->info.resize(Nn);   // (every entry is written by the loop below)
    h->Np = 0;
    {
      std::vector<char> badv(2, 0);
      parallel_range((size_t)Nn, [&](size_t b, size_t e) {
        for (size_t n = b; n < e; ++n) {
          const int nd = h->ndof[n], st = h->eqStart[n];
          h->info[n] = 0;
          if (!((nd == 3 || nd == 12) && st >= 0 && st + nd <= sizeEq)) { badv[0] = 1; continue; }
          int nfix = 0;
          for (int j = 0; j < nd; ++j) if (h->eqIdx[st + j] <= 0) nfix++;
          if (!(nfix == 0 || nfix == nd)) { badv[1] = 1; continue; }
          h->info[n] = nfix ? KIND_FIXED : (nd == 3 ? KIND_FREE3 : KIND_PML12);
        }
      });
      need(!badv[0], "eqd_set_mesh: bad dof table");
      need(!badv[1], "eqd_set_mesh: partially fixed node is not supported");
      // PML slots in ascending node order: per-chunk counts, then an exclusive scan
      const int nth = host_threads();
      std::vector<int> cntT(nth + 1, 0);
      const size_t per = ((size_t)Nn + nth - 1) / nth;
      parallel_range((size_t)nth, [&](size_t tb, size_t te) {
        for (size_t t = tb; t < te; ++t) {
          int c = 0;
          for (size_t n = std::min((size_t)Nn, t * per); n < std::min((size_t)Nn, (t + 1) * per); ++n) c += h->info[n] == KIND_PML12;
          cntT[t + 1] = c;
        }
      }, 1);
      for (int t = 0; t < nth; ++t) cntT[t + 1] += cntT[t];
      h->Np = cntT[nth];
      parallel_range((size_t)nth, [&](size_t tb, size_t te) {
        for (size_t t = tb; t < te; ++t) {
          int slot = cntT[t];
          for (size_t n = std::min((size_t)Nn, t * per); n < std::min((size_t)Nn, (t + 1) * per); ++n)
            if (h->info[n] == KIND_PML12) h->info[n] = KIND_PML12 | (slot++ << 3);
        }
      }, 1);
    }
    h->NnS = pad32(Nn);
    h->NpS = pad32(std::max(h->Np, 1));
    {
      std::vector<int> pn(std::max(h->Np, 1), 0);
      parallel_range((size_t)Nn, [&](size_t b, size_t e) {
        for (size_t n = b; n < e; ++n) if (EQD_INFO_KIND(h->info[n]) == KIND_PML12) pn[EQD_INFO_SLOT(h->info[n])] = (int)n;
      });
      h->dPmlNode.upload(pn);
    }
    need(3 * (double)h->NnS + 12 * (double)h->NpS < 4.0e9, "eqd_set_mesh: sub-domain too large for 32-bit halo offsets");
    // ---- element classes, cut into tiles (eqd_tiles.h)
    std::vector<int> members[3];
    {
      raw_vector<signed char> ccode(Ne);
      raw_vector<unsigned char> is12(Nn);   // one byte per node: the eight gathers per element stay in cache
      parallel_range((size_t)Nn, [&](size_t b, size_t e) { for (size_t n = b; n < e; ++n) is12[n] = h->ndof[n] == 12; });
      parallel_range((size_t)Ne, [&](size_t b, size_t e) {
        for (size_t el = b; el < e; ++el) {
          const int t = h->etype[el];
          int c;
          if (t == 2) c = CLS_PML;
          else if (t == 1 || (t >= 11 && t <= 13)) {
            const int* cn = &h->conn[8 * el];
            const int any = is12[cn[0]] | is12[cn[1]] | is12[cn[2]] | is12[cn[3]] | is12[cn[4]] | is12[cn[5]] | is12[cn[6]] | is12[cn[7]];
            c = any ? CLS_REGX : CLS_REG;
          } else c = -1;
          ccode[el] = (signed char)c;
        }
      });
      // member lists in ascending element order: per-chunk counts per class, exclusive scan, parallel fill
      const int nth = host_threads();
      const size_t per = ((size_t)Ne + nth - 1) / nth;
      std::vector<size_t> cnt3((size_t)(nth + 1) * 3, 0);
      std::vector<char> unknown(1, 0);
      parallel_range((size_t)nth, [&](size_t tb, size_t te) {
        for (size_t t = tb; t < te; ++t)
          for (size_t e = std::min((size_t)Ne, t * per); e < std::min((size_t)Ne, (t + 1) * per); ++e) {
            if (ccode[e] < 0) { unknown[0] = 1; continue; }
            cnt3[(t + 1) * 3 + ccode[e]]++;
          }
      }, 1);
      need(!unknown[0], "eqd_set_mesh: unknown element type");
      for (int t = 0; t < nth; ++t) for (int c = 0; c < 3; ++c) cnt3[(size_t)(t + 1) * 3 + c] += cnt3[(size_t)t * 3 + c];
      for (int c = 0; c < 3; ++c) members[c].resize(cnt3[(size_t)nth * 3 + c]);
      parallel_range((size_t)nth, [&](size_t tb, size_t te) {
        for (size_t t = tb; t < te; ++t) {
          size_t pos[3] = {cnt3[t * 3], cnt3[t * 3 + 1], cnt3[t * 3 + 2]};
          for (size_t e = std::min((size_t)Ne, t * per); e < std::min((size_t)Ne, (t + 1) * per); ++e) members[ccode[e]][pos[ccode[e]]++] = (int)e;
        }
      }, 1);
    }
    lap.lap("node kinds + classes");
    int gny = 0, gnz = 0;
    const bool gridOk = infer_grid(h->conn.data(), h->etype.data(), Ne, Nn, gny, gnz);
    h->elemCode.assign(Ne, 0);
    lap.sub("lattice strides");
    // ---- marching class: the box elements of the structured lattice leave the regular tile class (option "march";
    // the kernel covers the elastic path with Kosloff-Frazier hourglass control and no body force)
    {
      ElemClass& M = h->cls[CLS_MARCH];
      M.n = 0; M.S = 32; M.PFS = 4; M.nTiles = 0; M.nf = 3; M.nstress = 6;
      h->mGrid = 0; h->mBundles = 0; h->mFused = 0;
      const bool eligible = h->optMarch && gridOk && !h->plastic && !h->qmode && !h->body && P.C_hg == 1;
      if (eligible) {
        int nxg = 0;   // node planes of the grid in x: the (+,+,+) corner of every element is a lattice node (meshgen.f90:702-741)
        {
          std::vector<int> mx(host_threads() + 1, 0);
          const long nynz = (long)gny * gnz;
          parallel_range((size_t)Ne, [&](size_t b, size_t e) {
            int m = 0;
            for (size_t el = b; el < e; ++el) {
              // only corners that sit on the lattice count (a split-node master has an id beyond the grid): the y+ face
              // of the cell must be the four lattice nodes around corner 7
              const int* c = &h->conn[8 * el];
              const long id6 = c[6];
              const long ix = id6 / nynz, iz = (id6 % nynz) / gny, iy = id6 % gny;
              if (ix < 1 || iz < 1 || iy < 1) continue;
              if (c[2] != id6 - gny || c[7] != id6 - nynz || c[3] != id6 - nynz - gny) continue;
              m = std::max(m, (int)ix + 1);
            }
            static std::mutex mu;
            std::lock_guard<std::mutex> g(mu);
            mx[0] = std::max(mx[0], m);
          });
          nxg = mx[0];
        }
        h->mNxg = nxg;
      }
      lap.sub("node planes in x");
      if (eligible && !members[CLS_REG].empty()) {
        int dev = 0, sms = 148;
        CK(cudaGetDevice(&dev));
        CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        const int perSm = march_ctas_per_sm();
        need(perSm > 0, "option march: the marching kernel does not fit this device");
        MarchPlan MP;
        plan_march(h->conn.data(), h->etype.data(), h->coor.data(), h->info.data(), members[CLS_REG], Nn, gny, gnz, h->mNxg, perSm * sms,
                   h->optMarch == 2 ? 3 : h->optMarch == 3 ? 1 : 0, MP);   // 2: ghost rows and columns, 3: ghost columns (y) only
        lap.sub("plan_march");
        if (MP.n > 0) {
          M.n = MP.n; M.S = MP.S; M.PFS = MP.PFS;
          h->mGrid = MP.grid; h->mBundles = (int)MP.rec.size(); h->mFused = MP.nFused;
          h->mSlotB0 = MP.nBundlesA < (int)MP.rec.size() ? MP.rec[MP.nBundlesA].n0 : MP.PFS;
          h->mRec.upload(MP.rec); h->mCtaFirstA.upload(MP.ctaFirstA); h->mCtaFirstB.upload(MP.ctaFirstB);
          h->mCode.upload(MP.code); h->mSlotBundle.upload(MP.slotBundle);
          lap.sub("plan uploads");
          // host copy of the node slots for the slot table: only the nodes that emit a partial
          M.tnodeH.resize(MP.code.size());
          parallel_range(MP.code.size(), [&](size_t b, size_t e) {
            for (size_t k = b; k < e; ++k) {
              const int c = MP.code[k];
              M.tnodeH[k] = (c < 0 || (c & (MK_FUSED | MK_GHOST))) ? -1 : (c & MK_IDMASK);
              if (c >= 0 && (c & MK_FUSED)) h->info[c & MK_IDMASK] |= EQD_INFO_FUSED_BIT;   // one bundle per fused node: no race
            }
          });
          M.refId = std::move(MP.refId);
          h->mOwner = std::move(MP.owner);
          parallel_range((size_t)M.S, [&](size_t sb, size_t se) {
            for (size_t sl = sb; sl < se; ++sl) if (M.refId[sl] >= 0 && h->mOwner[sl]) h->elemCode[M.refId[sl]] = CLS_MARCH | ((int)sl << 2);
          });
          lap.sub("slot nodes + element codes");
          M.pf.alloc((size_t)3 * M.PFS);
          M.stress.alloc((size_t)6 * M.S);
          M.nBoxElems = M.n;
          members[CLS_REG] = std::move(MP.leftover);
          lap.sub("partial + stress buffers");
        }
      }
      h->pingpong = h->optMarch >= 2 && M.n > 0;
      lap.lap("march plan");
      // ---- the same for the PML class (eqd_march_pml.h)
      ElemClass& Q = h->cls[CLS_MARCHP];
      Q.n = 0; Q.S = 32; Q.PFS = 0; Q.nTiles = 0; Q.nf = 12; Q.nstress = 21;
      h->pGrid = 0; h->pBundles = 0;
      if (eligible && !members[CLS_PML].empty()) {
        int dev = 0, sms = 148;
        CK(cudaGetDevice(&dev));
        CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        const int perSm = march_pml_ctas_per_sm();
        need(perSm > 0, "option march: the PML marching kernel does not fit this device");
        MarchPlan MP;
        plan_march(h->conn.data(), h->etype.data(), h->coor.data(), h->info.data(), members[CLS_PML], Nn, gny, gnz, h->mNxg, perSm * sms, 0, MP, true);
        if (MP.n > 0) {
          Q.n = MP.n; Q.S = MP.S; Q.PFS = MP.PFS;
          h->pGrid = MP.grid; h->pBundles = (int)MP.rec.size();
          h->pSlotB0 = MP.nBundlesA < (int)MP.rec.size() ? MP.rec[MP.nBundlesA].n0 : MP.PFS;
          h->pRec.upload(MP.rec); h->pCtaFirstA.upload(MP.ctaFirstA); h->pCtaFirstB.upload(MP.ctaFirstB);
          h->pCode.upload(MP.code); h->pSlotBundle.upload(MP.slotBundle);
          Q.tnodeH.resize(MP.code.size());
          parallel_range(MP.code.size(), [&](size_t b, size_t e) {
            for (size_t k = b; k < e; ++k) Q.tnodeH[k] = MP.code[k] < 0 ? -1 : (MP.code[k] & MK_IDMASK);
          });
          Q.refId = std::move(MP.refId);
          parallel_range((size_t)Q.S, [&](size_t sb, size_t se) {
            for (size_t sl = sb; sl < se; ++sl) if (Q.refId[sl] >= 0) h->elemCode[Q.refId[sl]] = CLS_MARCH | ((int)sl << 2);   // (never matched by the scatter uploads)
          });
          Q.stress.alloc((size_t)21 * Q.S);
          Q.nBoxElems = Q.n;
          members[CLS_PML] = std::move(MP.leftover);
          // damping profile at the centroids, as for the PML tile class
          std::vector<double> dm(3 * (size_t)Q.S, 0.0);
          parallel_range((size_t)Q.S, [&](size_t sb, size_t se) {
            for (size_t sl = sb; sl < se; ++sl) {
              const int e = Q.refId[sl];
              if (e < 0) continue;
              double xc[3] = {0, 0, 0};
              for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 8; ++j) xc[i] = xc[i] + h->coor[i + 3 * (size_t)h->conn[8 * (size_t)e + j]];
              for (int i = 0; i < 3; ++i) xc[i] = xc[i] / 8;
              double d[3];
              pml_elem_damps(P, xc, d);
              for (int i = 0; i < 3; ++i) dm[(size_t)i * Q.S + sl] = d[i];
            }
          });
          Q.damps.upload(dm);
        }
      }
      lap.lap("PML march plan");
    }
    const int nf[3] = {3, 6, 12}, nstr[3] = {6, 6, 21};
    for (int c = 0; c < 3; ++c) {
      ElemClass& C = h->cls[c];
      TileShape sh;
      const int* ts = h->optTile[c == CLS_PML ? 1 : 0];
      sh.bx = ts[0]; sh.bz = ts[1]; sh.by = ts[2];
      if (c == CLS_PML) { sh.capE = 320; sh.capN = EQD_PML_LS; }
      else { sh.capE = 384; sh.capN = EQD_REG_LS; }  // two CTAs per SM
      sh.bankOrder = h->optBankOrder;
      TilePlan T;
      plan_tiles(h->conn.data(), members[c], Nn, gny, gnz, gridOk, sh, c == CLS_PML ? EQD_STAGE_PML : EQD_STAGE, T);
      C.n = T.n; C.S = T.S; C.nf = nf[c]; C.nstress = nstr[c];
      C.nTiles = T.nTiles; C.nFaceTiles = 0; C.LS = T.LS; C.PFS = T.PFS;
      need(tile_smem_bytes(c, h->qmode, C.LS) <= 227 * 1024, "eqd_set_mesh: tile does not fit shared memory");
      C.refId = std::move(T.refId);
      parallel_range((size_t)C.S, [&](size_t sb, size_t se) {
        for (size_t s = sb; s < se; ++s) if (C.refId[s] >= 0) h->elemCode[C.refId[s]] = c | ((int)s << 2);
      });
      C.tileNodeH = T.tileNode;
      if (!C.n) { C.tnodeH = std::move(T.tnode); continue; }
      need(C.LS <= (c == CLS_PML ? EQD_PML_LS : EQD_REG_LS), "eqd_set_mesh: tile has too many nodes");
      C.tileRecH.resize(C.nTiles);
      for (int t = 0; t < C.nTiles; ++t)
        C.tileRecH[t] = make_int4(T.tileElem[t], T.tileCnt[t] | ((int)T.tileColours[t] << 16), T.tileNode[t], T.tileNode[t + 1] - T.tileNode[t]);
      C.tileRec.upload(C.tileRecH);
      C.tnode.upload(T.tnode); C.lconn.upload(T.lconn);
      C.tnodeH = std::move(T.tnode);
      C.pf.alloc((size_t)C.nf * C.PFS);
      C.stress.alloc((size_t)C.nstress * C.S);
      if (c == CLS_PML) {
        std::vector<double> dm(3 * (size_t)C.S, 0.0);
        parallel_range((size_t)C.S, [&](size_t sb, size_t se) {
          for (size_t s = sb; s < se; ++s) {
            const int e = C.refId[s];
            if (e < 0) continue;
            double xc[3] = {0, 0, 0};
            for (int i = 0; i < 3; ++i)
              for (int j = 0; j < 8; ++j) xc[i] = xc[i] + h->coor[i + 3 * (size_t)h->conn[8 * (size_t)e + j]];
            for (int i = 0; i < 3; ++i) xc[i] = xc[i] / 8;
            double d[3];
            pml_elem_damps(P, xc, d);
            for (int i = 0; i < 3; ++i) dm[(size_t)i * C.S + s] = d[i];
          }
        });
        C.damps.upload(dm);
      } else if (h->qmode) {
        std::vector<uint8_t> qc(C.S, 0);
        for (int s = 0; s < C.S; ++s) {
          const int e = C.refId[s];
          if (e < 0) continue;
          double xc[3] = {0, 0, 0};
          for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 8; ++j) xc[i] = xc[i] + h->coor[i + 3 * (size_t)h->conn[8 * (size_t)e + j]];
          for (int i = 0; i < 3; ++i) xc[i] = xc[i] / 8.0;
          // calcElemKU.f90:86-99 (integer assignment truncates toward zero)
          const int depth = xc[2] > -1000.0 ? 0 : 1;
          const int ip = (int)((xc[0] - (P.PMLb[1] + P.dx / 2)) / P.dx + 1);
          const int iq = (int)((xc[1] - (P.PMLb[3] + P.dx / 2)) / P.dx + 1);
          const int ir = (int)((xc[2] - (P.PMLb[4] + P.dx / 2)) / P.dx + 1);
          const int k = 1 + ip % 2 + 2 * (iq % 2) + 4 * (ir % 2);
          need(k >= 1 && k <= 8, "Q class index out of range (element outside the uniform grid)");
          qc[s] = (uint8_t)(depth * 8 + k - 1);
        }
        C.qcls.upload(qc);
        C.qmem.alloc(6 * (size_t)C.S);
      }
    }
    {
      // the PML bundles' node slots follow the PML tiles' in one partial buffer (one class id for the node update's gather)
      ElemClass& T = h->cls[CLS_PML];
      ElemClass& Q = h->cls[CLS_MARCHP];
      h->pSlotBase = 0;
      if (Q.n) {
        h->pSlotBase = T.n ? T.PFS : 0;
        T.PFS = h->pSlotBase + Q.PFS;
        T.nf = 12;
        T.pf.alloc((size_t)12 * T.PFS);
      }
    }
    lap.lap("tile plan + uploads");
    // ---- node -> tile-node slots by rank, class then ascending tile id
    {
      // every thread owns a contiguous range of nodes.  The tile-node lists are first split, chunk by chunk, into one
      // bucket per owner thread; an owner then walks its buckets in (class, chunk) order, so each node's slots are still
      // visited in (class, ascending slot) order and every list entry is touched once, not once per thread
      std::vector<uint8_t> cnt(Nn, 0);
      const int nth = host_threads();
      std::vector<int> maxv(nth + 1, 2), badv(nth + 1, 0);
      const size_t NS = h->NnS;
      const size_t per = ((size_t)Nn + nth - 1) / nth;
      struct Pair { int nd; uint32_t code; };
      std::vector<std::vector<Pair>> bucket((size_t)NCLS * nth * nth);   // [class][chunk][owner]
      {
        std::vector<std::thread> th;
        for (int w = 0; w < nth; ++w)
          th.emplace_back([&, w] {
            for (int c = 0; c < NCLS; ++c) {
              const raw_vector<int>& tn = h->cls[c].tnodeH;
              const size_t chunk = (tn.size() + nth - 1) / nth, b = std::min(tn.size(), w * chunk), e = std::min(tn.size(), (w + 1) * chunk);
              for (size_t sl = b; sl < e; ++sl) {
                const int nd = tn[sl];
                if (nd < 0 || EQD_INFO_KIND(h->info[nd]) == KIND_FIXED) continue;
                bucket[((size_t)c * nth + w) * nth + nd / per].push_back({nd, c == CLS_MARCHP ? EQD_SLOT(CLS_PML, h->pSlotBase + sl) : EQD_SLOT(c, sl)});
              }
            }
          });
        for (auto& x : th) x.join();
      }
      auto scan = [&](raw_vector<uint32_t>* tab) {
        std::vector<std::thread> th;
        for (int t = 0; t < nth; ++t)
          th.emplace_back([&, t] {
            for (int c = 0; c < NCLS; ++c)
              for (int w = 0; w < nth; ++w)
                for (const Pair& q : bucket[((size_t)c * nth + w) * nth + t]) {
                  const int nd = q.nd;
                  if (!tab) {
                    if (cnt[nd] == 255) { badv[t] = 1; continue; }
                    maxv[t] = std::max(maxv[t], (int)++cnt[nd]);
                  } else {
                    if ((c == CLS_REG || c == CLS_MARCH) && EQD_INFO_KIND(h->info[nd]) == KIND_PML12) { badv[t] = 2; continue; }
                    (*tab)[(size_t)cnt[nd]++ * NS + nd] = q.code;
                  }
                }
          });
        for (auto& x : th) x.join();
      };
      scan(nullptr);
      int maxCnt = 2;
      for (int t = 0; t < nth; ++t) { maxCnt = std::max(maxCnt, maxv[t]); need(!badv[t], "eqd_set_mesh: a node belongs to more than 255 tiles"); }
      raw_vector<uint32_t> tab((size_t)maxCnt * NS);   // first touched (zeroed) by all threads, not serially
      parallel_range(tab.size(), [&](size_t b, size_t e) { std::fill(tab.begin() + b, tab.begin() + e, 0u); });
      parallel_range((size_t)Nn, [&](size_t b, size_t e) { std::fill(cnt.begin() + b, cnt.begin() + e, (uint8_t)0); });
      scan(&tab);
      for (int t = 0; t < nth; ++t) need(!badv[t], "internal: REG element on a 12-dof node");
      h->dSlotCnt.upload(cnt);
      h->dSlotTab.upload(tab);
    }
    lap.lap("slot table");
    // ---- PML node damping profile (comdampv, recomputed every step in the reference)
    std::vector<double> dp(3 * (size_t)h->NpS, 0.0);
    {
      std::vector<char> neg(1, 0);
      parallel_range((size_t)Nn, [&](size_t b, size_t e) {
        for (size_t n = b; n < e; ++n)
          if (EQD_INFO_KIND(h->info[n]) == KIND_PML12) {
            double d[3];
            if (!comdampv(P, h->coor[3 * n], h->coor[3 * n + 1], h->coor[3 * n + 2], d)) { neg[0] = 1; continue; }
            const size_t slot = EQD_INFO_SLOT(h->info[n]);
            for (int i = 0; i < 3; ++i) dp[(size_t)i * h->NpS + slot] = d[i];
          }
      });
      if (neg[0]) throw ArgError("negative PML damping (comdampv.f90:114-118)");
    }
    h->dDampp.upload(dp);
    for (int b = 0; b < (h->pingpong ? 2 : 1); ++b) { h->dVelB[b].alloc(3 * (size_t)h->NnS); h->dDispB[b].alloc(3 * (size_t)h->NnS); }
    h->cur = 0;
    h->dV1p.alloc(12 * (size_t)h->NpS);
    h->dForce.alloc(3 * (size_t)h->NnS + 12 * (size_t)h->NpS);
    h->dMass.alloc(Nn);
    lap.lap("comdampv + nodal allocs");
    h->meshSet = true;
  });
}

}  // extern "C"

namespace {
// porep / pstrain / stresses of eqd_set_elem_ops and eqd_compute_elem_ops
void upload_elem_state(eqd_handle* h, DevBuf<int>& dCode, const double* eleporep, const double* stressArr, const double* pstrain) {
  const int Ne = h->Ne;
  DevBuf<double> tmp;
  auto spread = [&](const double* src, int K, auto member, int k0, int nk, int rowOff) {
    tmp.alloc((size_t)K * Ne, false);
    h2d(tmp.p, src, sizeof(double) * (size_t)K * Ne);
    for (int c = 0; c < 3; ++c) {
      ElemClass& C = h->cls[c];
      DevBuf<double>& dst = C.*member;
      if (!C.n || !dst.p) continue;
      launch_aos_to_soa(tmp.p, K, Ne, dCode.p, c, dst.p + (size_t)rowOff * C.S, C.S, k0, nk, h->stream);
    }
    CK(cudaStreamSynchronize(h->stream));
  };
  if (h->plastic) {
    spread(eleporep, 1, &ElemClass::porep, 0, 1, 0);
    spread(pstrain, 1, &ElemClass::pstrain, 0, 1, 0);
  }
  tmp.release();
  // stresses: stressArr(stressCompIndexArr(e) + k).  Elastic runs start from rest (all zeros): then the device rows
  // are simply cleared instead of being gathered on the host and uploaded
  bool allZero = true;
  {
    std::vector<char> nz(1, 0);
    parallel_range((size_t)h->sizeStress, [&](size_t b, size_t e) {
      for (size_t k = b; k < e && !nz[0]; ++k) if (stressArr[k] != 0.0) nz[0] = 1;
    });
    allZero = !nz[0];
  }
  for (int c = 0; c < NCLS; ++c) {
    ElemClass& C = h->cls[c];
    if (!C.n) continue;
    if (allZero) {
      CK(cudaMemsetAsync(C.stress.p, 0, sizeof(double) * (size_t)C.nstress * C.S, h->stream));
      if (h->qmode && c != CLS_PML && C.qmem.p) CK(cudaMemsetAsync(C.qmem.p, 0, sizeof(double) * 6 * (size_t)C.S, h->stream));
      continue;
    }
    // first touched inside the parallel loop (no serial zero fill of ~1 GB); padding slots get zeros there
    raw_vector<double> sg((size_t)C.nstress * C.S), qm;
    if (h->qmode && c != CLS_PML && c != CLS_MARCHP) qm.resize(6 * (size_t)C.S);
    std::vector<char> badv(1, 0);
    parallel_range((size_t)C.S, [&](size_t sb, size_t se) {
      for (size_t s = sb; s < se; ++s) {
        const int base = C.refId[s] < 0 ? -1 : h->stressIdx[C.refId[s]];
        const bool ok = base >= 0 && base + (c == CLS_PML || c == CLS_MARCHP ? 21 : 12) <= h->sizeStress;
        if (C.refId[s] >= 0 && !ok) badv[0] = 1;
        for (int k = 0; k < C.nstress; ++k) sg[(size_t)k * C.S + s] = ok ? stressArr[base + k] : 0.0;
        if (!qm.empty()) for (int k = 0; k < 6; ++k) qm[(size_t)k * C.S + s] = ok ? stressArr[base + 6 + k] : 0.0;
      }
    });
    need(!badv[0], "eqd_set_elem_ops: stress index out of range");
    C.stress.upload(sg);
    if (!qm.empty()) C.qmem.upload(qm);
  }
}
void alloc_elem_ops(eqd_handle* h) {
  for (int c = 0; c < 3; ++c) {
    ElemClass& C = h->cls[c];
    if (!C.n) continue;
    C.shp.alloc(24 * (size_t)C.S); C.phi.alloc(32 * (size_t)C.S); C.ss.alloc(6 * (size_t)C.S);
    C.lam.alloc(C.S); C.mu.alloc(C.S); C.det.alloc(C.S);
    if (h->p.C_hg == 2) { C.rho.alloc(C.S); C.vp.alloc(C.S); }
    if (h->body) C.emass.alloc(8 * (size_t)C.S);
    if (h->plastic && c != CLS_PML) { C.porep.alloc(C.S); C.pstrain.alloc(C.S); }
  }
  for (int c : {CLS_MARCH, CLS_MARCHP}) {
    ElemClass& M = h->cls[c];
    if (M.n) { M.shp.alloc(3 * (size_t)M.S); M.ss.alloc(3 * (size_t)M.S); M.lam.alloc(M.S); M.mu.alloc(M.S); M.det.alloc(M.S); }
  }
}
}  // namespace

extern "C" {

// Operators computed on the device from the mesh of eqd_set_mesh (eqd_ops.cu); see the header.
int eqd_compute_elem_ops(eqd_handle* h, const double* mat, const double* eleporep, const double* stressArr,
                         const double* pstrain) {
  if (!h) return EQD_ERR_ARG;
  return guarded(h, [&] {
    need(h->meshSet, "eqd_compute_elem_ops: call eqd_set_mesh first");
    need(mat && stressArr, "eqd_compute_elem_ops: null argument");
    need(!h->plastic || (eleporep && pstrain), "eqd_compute_elem_ops: eleporep/pstrain required for C_elastic==0");
    const int Ne = h->Ne, Nn = h->Nn;
    Lap lap("eqd_compute_elem_ops");
    alloc_elem_ops(h);
    DevBuf<int> dConn, dEtype, dBad;
    DevBuf<double> dCoor, dMat;
    dConn.upload(h->conn); dEtype.upload(h->etype); dCoor.upload(h->coor);
    dMat.alloc(5 * (size_t)Ne, false);
    h2d(dMat.p, mat, sizeof(double) * 5 * (size_t)Ne);
    std::vector<int> big(1, 0x7fffffff);
    dBad.upload(big);
    lap.lap("mesh upload");
    DevBuf<double> emTmp[3], pm[3];
    for (int c = 0; c < 3; ++c) {
      ElemClass& C = h->cls[c];
      if (!C.n) continue;
      DevBuf<int> dRef; dRef.upload(C.refId);
      if (!h->body) emTmp[c].alloc(8 * (size_t)C.S);
      OpsArgs A{};
      A.S = C.S; A.Ne = Ne; A.refId = dRef.p; A.conn = dConn.p; A.etype = dEtype.p; A.coor = dCoor.p; A.mat = dMat.p;
      A.w = h->p.w;
      A.shp = C.shp.p; A.phi = C.phi.p; A.ss = C.ss.p; A.det = C.det.p; A.lam = C.lam.p; A.mu = C.mu.p;
      A.rho = C.rho.p; A.vp = C.vp.p;
      A.em = h->body ? C.emass.p : emTmp[c].p;
      A.badElem = dBad.p;
      launch_elem_ops(A, h->stream);
      // lumped mass partial per tile node, fixed order
      pm[c].alloc(C.PFS);
      TileMassArgs T{};
      T.tileRec = C.tileRec.p; T.em = A.em; T.lconn = C.lconn.p; T.S = C.S; T.capE = 384; T.pm = pm[c].p;
      launch_tile_mass(T, C.nTiles, h->stream);
      CK(cudaStreamSynchronize(h->stream));
    }
    DevBuf<double> emM[2], pmM;
    for (int k = 0; k < 2; ++k) {
      ElemClass& M = h->cls[k == 0 ? CLS_MARCH : CLS_MARCHP];
      if (M.n) {
        DevBuf<int> dRef; dRef.upload(M.refId);
        emM[k].alloc(8 * (size_t)M.S);
        OpsArgs A{};
        A.S = M.S; A.Ne = Ne; A.refId = dRef.p; A.conn = dConn.p; A.etype = dEtype.p; A.coor = dCoor.p; A.mat = dMat.p;
        A.w = h->p.w;
        A.shp = M.shp.p; A.phi = nullptr; A.ss = M.ss.p; A.det = M.det.p; A.lam = M.lam.p; A.mu = M.mu.p;
        A.rho = nullptr; A.vp = nullptr; A.em = emM[k].p; A.badElem = dBad.p; A.compact = 1;
        launch_elem_ops(A, h->stream);
        CK(cudaStreamSynchronize(h->stream));
      }
    }
    if (h->cls[CLS_MARCH].n) pmM.alloc(h->cls[CLS_MARCH].PFS);
    if (h->cls[CLS_MARCHP].n) {
      // mass partials of the PML bundles: behind the PML tiles' in one array, as their force partials are
      DevBuf<double> both;
      both.alloc(h->cls[CLS_PML].PFS);
      if (pm[CLS_PML].p && h->pSlotBase) CK(cudaMemcpy(both.p, pm[CLS_PML].p, sizeof(double) * h->pSlotBase, cudaMemcpyDeviceToDevice));
      const ElemClass& Q = h->cls[CLS_MARCHP];
      launch_march_mass(h->pRec.p, h->pBundles, h->pSlotBundle.p, h->pCode.p, emM[1].p, (size_t)Q.S, Q.PFS, both.p + h->pSlotBase, nullptr, h->stream);
      CK(cudaStreamSynchronize(h->stream));
      std::swap(pm[CLS_PML].p, both.p); std::swap(pm[CLS_PML].n, both.n);
    }
    CK(cudaGetLastError());
    const int bad = dBad.download()[0];
    if (bad != 0x7fffffff) throw ArgError("Non-positive determinant; element " + std::to_string(bad + 1) + " (calcGlobalShapeFunc.f90:57-61)");
    NodeMassArgs M{};
    M.Nn = Nn; M.NnS = h->NnS; M.slotCnt = h->dSlotCnt.p; M.slotTab = h->dSlotTab.p;
    for (int c = 0; c < 3; ++c) M.pm[c] = pm[c].p;
    M.pm[CLS_MARCH] = pmM.p;
    M.mass = h->dMass.p;
    if (h->cls[CLS_MARCH].n) {   // partials of the bundle surfaces first (k_node_mass sums them) ...
      const ElemClass& MC = h->cls[CLS_MARCH];
      launch_march_mass(h->mRec.p, h->mBundles, h->mSlotBundle.p, h->mCode.p, emM[0].p, (size_t)MC.S, MC.PFS, pmM.p, nullptr, h->stream);
    }
    launch_node_mass(M, h->stream);
    if (h->cls[CLS_MARCH].n) {   // ... then the complete masses of the fused nodes, which have no slot-table entry
      const ElemClass& MC = h->cls[CLS_MARCH];
      launch_march_mass(h->mRec.p, h->mBundles, h->mSlotBundle.p, h->mCode.p, emM[0].p, (size_t)MC.S, MC.PFS, pmM.p, h->dMass.p, h->stream);
    }
    CK(cudaStreamSynchronize(h->stream));
    h->massH = h->dMass.download();
    h->fnmsH = h->massH;                       // fnms adds the same element masses (assembleGlobalMass.f90:322)
    parallel_range((size_t)Nn, [&](size_t b, size_t e) {
      for (size_t n = b; n < e; ++n) if (EQD_INFO_KIND(h->info[n]) == KIND_FIXED) h->massH[n] = 1.0;
    });
    h->dMass.upload(h->massH);
    h->massFromDevice = true;
    lap.lap("operators + mass");
    DevBuf<int> dCode; dCode.upload(h->elemCode);
    upload_elem_state(h, dCode, eleporep, stressArr, pstrain);
    lap.lap("stresses");
    h->opsSet = true;
  });
}

int eqd_set_elem_ops(eqd_handle* h, const double* eleshp, const double* eledet, const double* elemass, const double* mat,
                     const double* ss, const double* phi, const double* eleporep, const double* stressArr,
                     const double* pstrain) {
  if (!h) return EQD_ERR_ARG;
  return guarded(h, [&] {
    need(h->meshSet, "eqd_set_elem_ops: call eqd_set_mesh first");
    need(eleshp && eledet && mat && ss && phi && stressArr, "eqd_set_elem_ops: null argument");
    need(!h->body || elemass, "eqd_set_elem_ops: elemass required (gravity / mass damping)");
    need(!h->plastic || (eleporep && pstrain), "eqd_set_elem_ops: eleporep/pstrain required for C_elastic==0");
    const int Ne = h->Ne;
    Lap lap("eqd_set_elem_ops");
    DevBuf<int> dCode; dCode.upload(h->elemCode);
    DevBuf<double> tmp;
    alloc_elem_ops(h);
    auto spread = [&](const double* src, int K, auto member, int k0, int nk, int rowOff) {
      tmp.alloc((size_t)K * Ne, false);
      h2d(tmp.p, src, sizeof(double) * (size_t)K * Ne);
      for (int c = 0; c < 3; ++c) {
        ElemClass& C = h->cls[c];
        DevBuf<double>& dst = C.*member;
        if (!C.n || !dst.p) continue;
        launch_aos_to_soa(tmp.p, K, Ne, dCode.p, c, dst.p + (size_t)rowOff * C.S, C.S, k0, nk, h->stream);
      }
      CK(cudaStreamSynchronize(h->stream));
    };
    // marching class: of eleshp only rows 3, 7, 14 (a_x, a_y, a_z), of ss the diagonal, no phi (eqd_box.h)
    DevBuf<int> dRefM[2];
    for (int k = 0; k < 2; ++k) if (h->cls[CLS_MARCH + k].n) dRefM[k].upload(h->cls[CLS_MARCH + k].refId);
    auto march_rows = [&](DevBuf<double> ElemClass::*member, int K, const int* rows, int nrows) {
      for (int m = 0; m < 2; ++m) {   // by slot, not by element: the ghost copies of an element get its rows too
        ElemClass& M = h->cls[CLS_MARCH + m];
        if (!M.n) continue;
        DevBuf<double>& dst = M.*member;
        for (int k = 0; k < nrows; ++k)
          launch_gather_rows(tmp.p, K, dRefM[m].p, M.S, dst.p + (size_t)k * M.S, rows[k], h->stream);
      }
      CK(cudaStreamSynchronize(h->stream));
    };
    const int shpRows[3] = {BOX_AX, BOX_AY, BOX_AZ}, ssRows[3] = {0, 3, 5}, row0[1] = {0};
    spread(eleshp, 24, &ElemClass::shp, 0, 24, 0);
    march_rows(&ElemClass::shp, 24, shpRows, 3);
    spread(phi, 32, &ElemClass::phi, 0, 32, 0);
    spread(ss, 6, &ElemClass::ss, 0, 6, 0);
    march_rows(&ElemClass::ss, 6, ssRows, 3);
    spread(eledet, 1, &ElemClass::det, 0, 1, 0);
    march_rows(&ElemClass::det, 1, row0, 1);
    spread(mat + 3 * (size_t)Ne, 1, &ElemClass::lam, 0, 1, 0);  // mat(Ne,5): element index fastest
    march_rows(&ElemClass::lam, 1, row0, 1);
    spread(mat + 4 * (size_t)Ne, 1, &ElemClass::mu, 0, 1, 0);
    march_rows(&ElemClass::mu, 1, row0, 1);
    if (h->p.C_hg == 2) {
      spread(mat + 2 * (size_t)Ne, 1, &ElemClass::rho, 0, 1, 0);
      spread(mat, 1, &ElemClass::vp, 0, 1, 0);
    }
    if (h->body) {
      tmp.alloc((size_t)24 * Ne, false);
      h2d(tmp.p, elemass, sizeof(double) * (size_t)24 * Ne);
      for (int c = 0; c < 3; ++c) {
        ElemClass& C = h->cls[c];
        if (!C.n) continue;
        for (int i = 0; i < 8; ++i) launch_aos_to_soa(tmp.p, 24, Ne, dCode.p, c, C.emass.p + (size_t)i * C.S, C.S, 3 * i, 1, h->stream);
      }
      CK(cudaStreamSynchronize(h->stream));
    }
    tmp.release();
    lap.lap("operator rows");
    upload_elem_state(h, dCode, eleporep, stressArr, pstrain);
    lap.lap("stresses");
    h->opsSet = true;
  });
}

int eqd_set_nodal(eqd_handle* h, const double* nodalMassArr, const double* fnms, const double* v1, const double* velArr,
                  const double* dispArr, const double* nodalForceArr) {
  if (!h) return EQD_ERR_ARG;
  return guarded(h, [&] {
    need(h->meshSet, "eqd_set_nodal: call eqd_set_mesh first");
    need(v1 && velArr && dispArr, "eqd_set_nodal: null argument");
    // NULL mass arrays = keep the lumped mass eqd_compute_elem_ops assembled on the device
    const bool devMass = !nodalMassArr && !fnms && h->massFromDevice;
    need(devMass || (nodalMassArr && fnms), "eqd_set_nodal: nodalMassArr / fnms missing (and no eqd_compute_elem_ops before)");
    const int Nn = h->Nn;
    const size_t NS = h->NnS, PS = h->NpS;
    Lap lap("eqd_set_nodal");
    if (!devMass) {
      h->fnmsH.assign(fnms, fnms + Nn);
      h->massH.assign(Nn, 1.0);
      h->massFromDevice = false;
    }
    // first touched inside the parallel loops below (no serial zero fill of ~1 GB)
    raw_vector<double> vel(3 * NS), disp(3 * NS), v1p(12 * PS);
    std::vector<double> acc;
    parallel_range(12 * PS, [&](size_t b, size_t e) { std::fill(v1p.begin() + b, v1p.begin() + e, 0.0); });
    bool anyAcc = false;
    if (nodalForceArr) {
      std::vector<char> any(1, 0);
      parallel_range((size_t)h->Neq, [&](size_t b, size_t e) {
        for (size_t k = b; k < e && !any[0]; ++k) if (nodalForceArr[k] != 0.0) any[0] = 1;
      });
      anyAcc = any[0] != 0;
    }
    if (anyAcc) acc.assign(3 * NS + 12 * PS, 0.0);
    std::vector<char> badv(1, 0);
    parallel_range(NS, [&](size_t nb, size_t ne) {
      for (size_t n = nb; n < ne; ++n) {
        const int kind = n < (size_t)Nn ? EQD_INFO_KIND(h->info[n]) : KIND_FIXED;
        if (kind == KIND_FIXED) {   // fixed nodes and the padding of the rows stay at rest
          for (int j = 0; j < 3; ++j) { vel[j * NS + n] = 0.0; disp[j * NS + n] = 0.0; }
          continue;
        }
        const int st = h->eqStart[n], nd = h->ndof[n];
        if (!devMass) {
          const double m = nodalMassArr[h->eqIdx[st] - 1];
          for (int j = 1; j < nd; ++j)
            if (nodalMassArr[h->eqIdx[st + j] - 1] != m) badv[0] = 1;
          h->massH[n] = m;
        }
        for (int j = 0; j < 3; ++j) disp[j * NS + n] = dispArr[j + 3 * n];
        if (kind == KIND_FREE3) {
          // v1 and velArr are the same quantity for a 3-dof node after the first update (driver.f90:102-103)
          for (int j = 0; j < 3; ++j) vel[j * NS + n] = v1[h->eqIdx[st + j] - 1];
          if (anyAcc) for (int j = 0; j < 3; ++j) acc[j * NS + n] = nodalForceArr[h->eqIdx[st + j] - 1];
        } else {
          const size_t slot = EQD_INFO_SLOT(h->info[n]);
          for (int j = 0; j < 12; ++j) v1p[j * PS + slot] = v1[h->eqIdx[st + j] - 1];
          for (int j = 0; j < 3; ++j) vel[j * NS + n] = velArr[j + 3 * n];
          if (anyAcc) for (int j = 0; j < 12; ++j) acc[3 * NS + j * PS + slot] = nodalForceArr[h->eqIdx[st + j] - 1];
        }
      }
    });
    need(!badv[0], "eqd_set_nodal: dofs of one node carry different lumped masses");
    // (both buffers of a double-buffered pair: the nodes no kernel ever writes -- fixed ones -- must read the same in either)
    for (int b = 0; b < (h->pingpong ? 2 : 1); ++b) { h->dVelB[b].upload(vel); h->dDispB[b].upload(disp); }
    h->cur = 0;
    h->dV1p.upload(v1p); h->dMass.upload(h->massH);
    if (anyAcc) h->dAccel0.upload(acc); else h->dAccel0.release();
    lap.lap("all");
    h->nodalSet = true;
  });
}

int eqd_set_fault(eqd_handle* h, int32_t nftmx, const int32_t* nftnd, const int32_t* nsmp, const double* un,
                  const double* us, const double* ud, const double* arn, const double* fric, const double* fnft) {
  if (!h) return EQD_ERR_ARG;
  return guarded(h, [&] {
    need(h->meshSet && h->nodalSet, "eqd_set_fault: call eqd_set_mesh and eqd_set_nodal first");
    need(!h->finalized, "eqd_set_fault: already running");
    const int ntotft = h->p.ntotft;
    need(nftmx >= 1 && ntotft >= 1 && nftnd && nsmp && un && us && ud && arn && fric && fnft, "eqd_set_fault: null / empty argument");
    h->nftmx = nftmx; h->ntotft = ntotft;
    h->nftnd.assign(nftnd, nftnd + ntotft);
    h->pairRef.clear();
    for (int ift = 0; ift < ntotft; ++ift)
      for (int i = 0; i < nftnd[ift]; ++i) h->pairRef.push_back(i + nftmx * ift);
    h->nPairs = (int)h->pairRef.size();
    h->PS = pad32(std::max(h->nPairs, 1));
    const size_t PS = h->PS;
    std::vector<int> nS(PS, 0), nM(PS, 0), ift(PS, 0);
    std::vector<double> u1(3 * PS, 0.0), u2(3 * PS, 0.0), u3(3 * PS, 0.0), ar(PS, 1.0), mS(PS, 1.0), mM(PS, 1.0), xs(3 * PS, 0.0),
        fr(100 * PS, 0.0), ft(PS, 0.0);
    h->pairNodeS.resize(h->nPairs); h->pairNodeM.resize(h->nPairs);
    for (int q = 0; q < h->nPairs; ++q) {
      const size_t pb = h->pairRef[q];
      const int s = nsmp[2 * pb] - 1, m = nsmp[2 * pb + 1] - 1;
      need(s >= 0 && s < h->Nn && m >= 0 && m < h->Nn, "eqd_set_fault: nsmp out of range");
      need(EQD_INFO_KIND(h->info[s]) == KIND_FREE3 && EQD_INFO_KIND(h->info[m]) == KIND_FREE3, "eqd_set_fault: split nodes must be free 3-dof nodes");
      nS[q] = s; nM[q] = m; ift[q] = (int)(pb / nftmx) + 1;
      h->pairNodeS[q] = s; h->pairNodeM[q] = m;
      for (int k = 0; k < 3; ++k) {
        u1[k * PS + q] = un[k + 3 * pb]; u2[k * PS + q] = us[k + 3 * pb]; u3[k * PS + q] = ud[k + 3 * pb];
        xs[k * PS + q] = h->coor[k + 3 * (size_t)s];
      }
      ar[q] = arn[pb]; mS[q] = h->fnmsH[s]; mM[q] = h->fnmsH[m]; ft[q] = fnft[pb];
      for (int k = 0; k < 100; ++k) fr[k * PS + q] = fric[k + 100 * pb];
    }
    h->dNodeS.upload(nS); h->dNodeM.upload(nM); h->dIft.upload(ift);
    h->dUn.upload(u1); h->dUs.upload(u2); h->dUd.upload(u3); h->dArn.upload(ar);
    h->dMassS.upload(mS); h->dMassM.upload(mM); h->dXs.upload(xs); h->dFric.upload(fr); h->dFnft.upload(ft);
    h->faultSet = true;
  });
}

int eqd_set_halo(eqd_handle* h, const int32_t* numcount, const int32_t* fltnum, const int32_t* fltMPI, const int32_t* fltl,
                 const int32_t* fltr, const int32_t* fltf, const int32_t* fltb, const int32_t* fltd, const int32_t* fltu) {
  if (!h) return EQD_ERR_ARG;
  return guarded(h, [&] {
    need(h->meshSet && !h->finalized, "eqd_set_halo: call after eqd_set_mesh and before eqd_run");
    need(numcount && fltnum && fltMPI, "eqd_set_halo: null argument");
    for (int k = 0; k < 9; ++k) h->numcount[k] = numcount[k];
    const int32_t* lists[6] = {fltl, fltr, fltf, fltb, fltd, fltu};
    for (int k = 0; k < 6; ++k) {
      h->fltnum[k] = fltnum[k]; h->fltMPI[k] = fltMPI[k];
      h->fltface[k].clear();
      if (fltnum[k] > 0 && fltMPI[k]) {
        need(lists[k] != nullptr, "eqd_set_halo: face list missing");
        h->fltface[k].assign(lists[k], lists[k] + fltnum[k]);
      }
    }
    need((long)numcount[0] * numcount[1] * numcount[2] <= h->Nn, "eqd_set_halo: numcount(1:3) exceeds the node count");
    h->haloSet = true;
  });
}

int eqd_set_stations(eqd_handle* h, const int32_t* idhist, int32_t nOff, const int32_t* anonfs, int32_t nOn,
                     const int32_t* surfaceNodeIdArr, int32_t nSurf) {
  if (!h) return EQD_ERR_ARG;
  return guarded(h, [&] {
    need(h->meshSet && !h->finalized, "eqd_set_stations: call after eqd_set_mesh and before eqd_run");
    h->nOff = nOff > 0 && idhist ? nOff : 0;
    h->nOn = nOn > 0 && anonfs ? nOn : 0;
    h->nSurf = nSurf > 0 && surfaceNodeIdArr ? nSurf : 0;
    if (h->nOff) {
      h->idhistH.assign(idhist, idhist + 18 * (size_t)h->nOff);
      for (int i = 0; i < 6 * h->nOff; ++i) {
        need(h->idhistH[3 * i] >= 1 && h->idhistH[3 * i] <= h->Nn && h->idhistH[3 * i + 1] >= 1 && h->idhistH[3 * i + 1] <= 3,
             "eqd_set_stations: idhist out of range");
        need(h->idhistH[3 * i + 2] == 1 || h->idhistH[3 * i + 2] == 2, "eqd_set_stations: idhist(3,:) must be 1 (disp) or 2 (vel)");
      }
      h->dIdhist.upload(h->idhistH);
    }
    if (h->nOn) h->anonfsH.assign(anonfs, anonfs + 3 * (size_t)h->nOn);
    if (h->nSurf) { h->surfH.assign(surfaceNodeIdArr, surfaceNodeIdArr + h->nSurf); h->dSurf.upload(h->surfH); }
  });
}

int eqd_get_unique_id(void* id128) {
  if (!id128) return EQD_ERR_ARG;
  if (!g_nccl.load()) { fprintf(stderr, "eqdyna_b200: cannot load NCCL: %s\n", g_nccl.why.c_str()); return EQD_ERR_CUDA; }
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  if (g_nccl.GetUniqueId(&id) != ncclSuccess) return EQD_ERR_CUDA;
  memcpy(id128, &id, 128);
  return EQD_OK;
}

int eqd_set_comm(eqd_handle* h, const void* id128, int32_t nranks, int32_t rank) {
  if (!h) return EQD_ERR_ARG;
  return guarded(h, [&] {
    need(id128 && nranks >= 1 && rank >= 0 && rank < nranks, "eqd_set_comm: bad arguments");
    need(nranks == h->p.npx * h->p.npy * h->p.npz && rank == h->p.me, "eqd_set_comm: rank/nranks must match me and npx*npy*npz");
    if (!g_nccl.load()) throw CudaError("cannot load NCCL: " + g_nccl.why);
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    comm_ready(h);
    h->nranks = nranks; h->rank = rank;
    h->commRc = ncclSuccess;
    h->commThread = std::thread([h, id, nranks, rank] {
      cudaSetDevice(h->device);
      h->commRc = g_nccl.CommInitRank(&h->comm, nranks, id, rank);
    });
  });
}

int eqd_set_host_comm(eqd_handle* h, int32_t nranks, int32_t rank, eqd_allgather_fn fn, void* ctx) {
  if (!h) return EQD_ERR_ARG;
  return guarded(h, [&] {
    need(fn && nranks >= 1 && rank >= 0 && rank < nranks, "eqd_set_host_comm: bad arguments");
    need(nranks == h->p.npx * h->p.npy * h->p.npz && rank == h->p.me, "eqd_set_host_comm: rank/nranks must match me and npx*npy*npz");
    need(!h->finalized, "eqd_set_host_comm: call before the first eqd_run / eqd_sum_shared");
    h->nranks = nranks; h->rank = rank;
    h->hostAg = fn; h->hostCtx = ctx;
  });
}

// The init-time sums of the reference (nodalMassArr and fnms through
// MPI4NodalQuant, assembleGlobalMass.f90:40-41; arn through MPI4arn,
// meshgen.f90:274-395), x then y then z, for a one-process-per-GPU host.
int eqd_sum_shared(eqd_handle* h) {
  if (!h) return EQD_ERR_ARG;
  return guarded(h, [&] {
    need(h->nodalSet, "eqd_sum_shared: call after eqd_set_nodal / eqd_set_fault / eqd_set_halo");
    finalize(h);
    if (!has_neighbours(h)) return;
    if (!h->hostAg) comm_ready(h);
    need(h->comm != nullptr || h->hostAg, "eqd_sum_shared: eqd_set_comm or eqd_set_host_comm first");
    // host-staged: these run once
    for (int a = 0; a < 3; ++a) {
      std::vector<double> sendv[2], recvv[2];
      for (int side = 0; side < 2; ++side) {
        Face& F = h->face[a][side];
        if (F.nb < 0) continue;
        for (int n : F.nodes) { sendv[side].push_back(h->massH[n]); sendv[side].push_back(h->fnmsH[n]); }
        std::vector<double> arnH = h->dArn.download();
        for (int q : F.pairs) sendv[side].push_back(arnH[q]);
      }
      if (h->hostAg) {
        // one all-gather of [len(-side), len(+side)] and one of the two vectors padded to the longest: a rank reads
        // what the facing side of each neighbour sent (the two sides of a face hold the same nodes in the same order)
        unsigned long long len[2] = {sendv[0].size(), sendv[1].size()};
        std::vector<unsigned long long> lens(2 * (size_t)h->nranks);
        setup_allgather(h, len, sizeof len, lens.data());
        size_t L = 0;
        for (unsigned long long v : lens) L = std::max(L, (size_t)v);
        if (L) {
          std::vector<double> mine(2 * L, 0.0), all(2 * L * (size_t)h->nranks);
          for (int side = 0; side < 2; ++side) std::copy(sendv[side].begin(), sendv[side].end(), mine.begin() + side * L);
          setup_allgather(h, mine.data(), sizeof(double) * 2 * L, all.data());
          for (int side = 0; side < 2; ++side) {
            const Face& F = h->face[a][side];
            if (F.nb < 0 || sendv[side].empty()) continue;
            need(lens[2 * (size_t)F.nb + (1 - side)] == sendv[side].size(), "eqd_sum_shared: the two sides of a rank face differ in length");
            const double* src = all.data() + (2 * (size_t)F.nb + (1 - side)) * L;
            recvv[side].assign(src, src + sendv[side].size());
          }
        }
      } else {
        DevBuf<double> ds[2], dr[2];
        for (int side = 0; side < 2; ++side) if (h->face[a][side].nb >= 0) { ds[side].upload(sendv[side]); dr[side].alloc(sendv[side].size()); }
        NK(g_nccl.GroupStart());
        for (int side = 0; side < 2; ++side) {
          Face& F = h->face[a][side];
          if (F.nb < 0 || ds[side].n == 0) continue;
          NK(g_nccl.Send(ds[side].p, ds[side].n, ncclDouble, F.nb, h->comm, h->stream));
          NK(g_nccl.Recv(dr[side].p, dr[side].n, ncclDouble, F.nb, h->comm, h->stream));
        }
        NK(g_nccl.GroupEnd());
        CK(cudaStreamSynchronize(h->stream));
        for (int side = 0; side < 2; ++side) if (h->face[a][side].nb >= 0 && ds[side].n) recvv[side] = dr[side].download();
      }
      for (int side = 0; side < 2; ++side) {
        Face& F = h->face[a][side];
        if (F.nb < 0 || sendv[side].empty()) continue;
        size_t k = 0;
        for (int n : F.nodes) {
          if (EQD_INFO_KIND(h->info[n]) != KIND_FIXED) h->massH[n] = h->massH[n] + recvv[side][k];
          h->fnmsH[n] = h->fnmsH[n] + recvv[side][k + 1];
          k += 2;
        }
        std::vector<double> arnH = h->dArn.download();
        for (int q : F.pairs) arnH[q] = arnH[q] + recvv[side][k++];
        h->dArn.upload(arnH);
      }
    }
    h->dMass.upload(h->massH);
    std::vector<double> mS(h->PS, 1.0), mM(h->PS, 1.0);
    for (int q = 0; q < h->nPairs; ++q) { mS[q] = h->fnmsH[h->pairNodeS[q]]; mM[q] = h->fnmsH[h->pairNodeM[q]]; }
    if (h->nPairs) { h->dMassS.upload(mS); h->dMassM.upload(mM); }
  });
}

int eqd_run(eqd_handle* h, int32_t nt_begin, int32_t nt_end) {
  if (!h) return EQD_ERR_ARG;
  int rc2 = EQD_OK;
  int rc = guarded(h, [&] {
    need(nt_begin >= 1 && nt_end <= h->p.nstep, "eqd_run: step range outside 1..nstep");
    prepare_run(h, nt_begin);
    const bool multi = has_neighbours(h);
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    if (h->timing) { CK(cudaEventCreate(&t0)); CK(cudaEventCreate(&t1)); CK(cudaEventRecord(t0, h->stream)); }
    for (int nt = nt_begin; nt <= nt_end; ++nt) {
      // auto: with rank neighbours the exchange and the fault solver go to the second stream, under the next step's
      // bulk node update.  Mode 2 (boundary work list first, exchange under the interior sweep) is an option only:
      // measured slower on one box -- 1.619 against 1.577 ms at 2 GPUs, 0.534 against 0.495 ms at 8 (r02_k, r02_l) --
      // because two half-filled persistent launches per class cost more than the exchange it hides
      const int ov = !h->commStream ? 0 : h->optOverlap < 0 ? (multi ? 1 : 0) : h->optOverlap;
      step_pre(h, ov, multi, nt == nt_end);
      if (multi && !ov) halo_all_nccl(h, h->stream);
      step_post(h, ov);
    }
    // the last step's fault solver ran on the communication stream
    if (h->commStream && nt_end >= nt_begin) CK(cudaStreamWaitEvent(h->stream, h->evComm, 0));
    if (h->timing) {
      CK(cudaEventRecord(t1, h->stream));
      CK(cudaEventSynchronize(t1));
      float ms = 0; CK(cudaEventElapsedTime(&ms, t0, t1));
      h->tms[EQD_T_TOTAL] += ms;
      cudaEventDestroy(t0); cudaEventDestroy(t1);
    }
    rc2 = finish_run(h);
  });
  return rc != EQD_OK ? rc : rc2;
}

// Several sub-domains driven by ONE process (all on one device, or one device
// each with peer access): lock-step run with device-to-device halo copies.
int eqd_run_group(eqd_handle** hs, int32_t n, int32_t nt_begin, int32_t nt_end) {
  if (!hs || n < 1) return EQD_ERR_ARG;
  int rcAll = EQD_OK;
  int rc = guarded(hs[0], [&] {
    need(n == hs[0]->p.npx * hs[0]->p.npy * hs[0]->p.npz, "eqd_run_group: need every sub-domain of the decomposition");
    for (int r = 0; r < n; ++r) {
      need(hs[r] && hs[r]->p.me == r, "eqd_run_group: handles must be ordered by rank id");
      CK(cudaSetDevice(hs[r]->device));
      need(nt_begin >= 1 && nt_end <= hs[r]->p.nstep, "eqd_run_group: step range outside 1..nstep");
      prepare_run(hs[r], nt_begin);
    }
    for (int nt = nt_begin; nt <= nt_end; ++nt) {
      for (int r = 0; r < n; ++r) { CK(cudaSetDevice(hs[r]->device)); step_pre(hs[r], 0, false, nt == nt_end); }
      for (int a = 0; a < 3; ++a) {
        for (int r = 0; r < n; ++r) {
          eqd_handle* h = hs[r];
          if (!h->haloSet || (h->face[a][0].nb < 0 && h->face[a][1].nb < 0)) continue;
          CK(cudaSetDevice(h->device));
          halo_pack(h, a);
          CK(cudaEventRecord(h->evPacked[a], h->stream));
        }
        for (int r = 0; r < n; ++r) {
          eqd_handle* h = hs[r];
          if (!h->haloSet) continue;
          CK(cudaSetDevice(h->device));
          for (int side = 0; side < 2; ++side) {
            Face& F = h->face[a][side];
            if (F.nb < 0 || F.n == 0) continue;
            eqd_handle* o = hs[F.nb];
            Face& G = o->face[a][1 - side];
            need(G.n == F.n, "eqd_run_group: face sizes differ between neighbours");
            CK(cudaStreamWaitEvent(h->stream, o->evPacked[a], 0));
            CK(cudaMemcpyAsync(F.recv.p, G.send.p, sizeof(double) * F.n, cudaMemcpyDefault, h->stream));
          }
          halo_unpack(h, a);
        }
        // a rank's send buffers of this axis are reused next step only: the next
        // pack on the same stream is ordered after the neighbours' copies by the
        // step_pre/step_post kernels of every rank being enqueued in between and
        // by the device-wide join below
      }
      for (int r = 0; r < n; ++r) { CK(cudaSetDevice(hs[r]->device)); step_post(hs[r], 0); }
      // lock-step join so that no rank overwrites a send buffer a neighbour still reads
      for (int r = 0; r < n; ++r) { CK(cudaSetDevice(hs[r]->device)); CK(cudaStreamSynchronize(hs[r]->stream)); }
    }
    for (int r = 0; r < n; ++r) {
      CK(cudaSetDevice(hs[r]->device));
      int rc1 = finish_run(hs[r]);
      if (rc1 != EQD_OK) { rcAll = rc1; if (r) hs[0]->err = hs[r]->err; }
    }
  });
  return rc != EQD_OK ? rc : rcAll;
}

int eqd_fetch(eqd_handle* h, int32_t which, void* dst, int64_t dst_bytes) {
  if (!h) return EQD_ERR_ARG;
  return guarded(h, [&] {
    need(dst != nullptr, "eqd_fetch: null destination");
    need(h->meshSet && h->nodalSet, "eqd_fetch: nothing uploaded yet");
    CK(cudaStreamSynchronize(h->stream));
    double* out = (double*)dst;
    const int Nn = h->Nn;
    const size_t NS = h->NnS, PSn = h->NpS;
    auto want = [&](size_t nDoubles) { need((size_t)dst_bytes >= nDoubles * sizeof(double), "eqd_fetch: destination too small"); };
    switch (which) {
      case EQD_F_DISP: case EQD_F_VEL: {
        want(3 * (size_t)Nn);
        const double* dev = (which == EQD_F_DISP ? h->dispCur() : h->velCur()).p;
        if ((size_t)Nn * 3 * sizeof(double) < (64u << 20)) {
          std::vector<double> v = (which == EQD_F_DISP ? h->dispCur() : h->velCur()).download();
          parallel_range((size_t)Nn, [&](size_t b, size_t e) {
            for (size_t n = b; n < e; ++n) for (int j = 0; j < 3; ++j) out[j + 3 * n] = v[j * NS + n];
          });
          break;
        }
        // large sub-domains: chunks of the three SoA rows land in a pinned buffer (DMA at link
        // speed, no pageable bounce, no serial first touch of a 0.5 GB temporary) and are
        // interleaved into the caller's (3,Nn) array by all host threads
        if (!h->stage) h->stage = shared_stage(h->device);
        void* pinv = nullptr;
        CK(h->stage->pinned(0, &pinv));
        double* pin = (double*)pinv;
        const size_t cap = StagedCopy::kChunk / sizeof(double) / 3;   // nodes per chunk
        for (size_t n0 = 0; n0 < (size_t)Nn; n0 += cap) {
          const size_t m = std::min(cap, (size_t)Nn - n0);
          for (int j = 0; j < 3; ++j)
            CK(cudaMemcpyAsync(pin + j * m, dev + j * NS + n0, m * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
          CK(cudaStreamSynchronize(h->stream));
          parallel_range(m, [&](size_t b, size_t e) {
            for (size_t i = b; i < e; ++i) for (int j = 0; j < 3; ++j) out[j + 3 * (n0 + i)] = pin[j * m + i];
          });
        }
        break;
      }
      case EQD_F_V1: case EQD_F_FORCE: case EQD_F_MASS: {
        want(h->Neq);
        std::vector<double> a3, a12;
        if (which == EQD_F_V1) { a3 = h->velCur().download(); a12 = h->dV1p.download(); }
        else if (which == EQD_F_FORCE) {
          finalize(h);
          DevBuf<double> acc; acc.alloc(3 * NS + 12 * PSn);
          NodeArgs A = h->nodeArgs();
          if (h->dAccel0.p) A.accel0 = h->dAccel0.p;
          launch_materialize_accel(A, acc.p, h->stream);
          CK(cudaStreamSynchronize(h->stream));
          std::vector<double> all = acc.download();
          a3.assign(all.begin(), all.begin() + 3 * NS);
          a12.assign(all.begin() + 3 * NS, all.end());
        }
        for (int n = 0; n < Nn; ++n) {
          const int kind = EQD_INFO_KIND(h->info[n]);
          if (kind == KIND_FIXED) continue;
          const int st = h->eqStart[n];
          if (which == EQD_F_MASS) { for (int j = 0; j < h->ndof[n]; ++j) out[h->eqIdx[st + j] - 1] = h->massH[n]; continue; }
          if (kind == KIND_FREE3) for (int j = 0; j < 3; ++j) out[h->eqIdx[st + j] - 1] = a3[j * NS + n];
          else { const size_t slot = EQD_INFO_SLOT(h->info[n]); for (int j = 0; j < 12; ++j) out[h->eqIdx[st + j] - 1] = a12[j * PSn + slot]; }
        }
        break;
      }
      case EQD_F_ELEDET: case EQD_F_ELESHP: case EQD_F_SS: case EQD_F_PHI: {
        need(h->opsSet, "eqd_fetch: no element operators yet");
        const int K = which == EQD_F_ELEDET ? 1 : which == EQD_F_ELESHP ? 24 : which == EQD_F_SS ? 6 : 32;
        want((size_t)K * h->Ne);
        for (int mc : {CLS_MARCH, CLS_MARCHP}) {
          // the marching classes keep a_x, a_y, a_z and the diagonal of ss only; the rest is the closed form of eqd_box.h
          ElemClass& C = h->cls[mc];
          if (!C.n) continue;
          std::vector<double> a = (which == EQD_F_ELESHP ? C.shp : which == EQD_F_SS ? C.ss : C.det).download();
          parallel_range((size_t)C.S, [&](size_t sb, size_t se) {
            for (size_t s = sb; s < se; ++s) {
              const int e = C.refId[s];
              if (e < 0 || (mc == CLS_MARCH && !h->mOwner[s])) continue;
              double* o = out + (size_t)K * e;
              if (which == EQD_F_ELEDET) o[0] = a[s];
              else if (which == EQD_F_ELESHP)
                for (int i = 0; i < 8; ++i) {
                  o[3 * i] = box_px(i) ? a[s] : -a[s];
                  o[3 * i + 1] = box_py(i) ? a[C.S + s] : -a[C.S + s];
                  o[3 * i + 2] = box_pz(i) ? a[2 * (size_t)C.S + s] : -a[2 * (size_t)C.S + s];
                }
              else if (which == EQD_F_SS) { o[0] = a[s]; o[1] = 0; o[2] = 0; o[3] = a[C.S + s]; o[4] = 0; o[5] = a[2 * (size_t)C.S + s]; }
              else
                for (int m = 0; m < 4; ++m) for (int i = 0; i < 8; ++i) o[8 * m + i] = box_hp(m, i) ? 1.0 : -1.0;
            }
          });
        }
        for (int c = 0; c < 3; ++c) {
          ElemClass& C = h->cls[c];
          if (!C.n) continue;
          const DevBuf<double>& src = which == EQD_F_ELEDET ? C.det : which == EQD_F_ELESHP ? C.shp : which == EQD_F_SS ? C.ss : C.phi;
          std::vector<double> v = src.download();
          parallel_range((size_t)C.S, [&](size_t sb, size_t se) {
            for (size_t s = sb; s < se; ++s) {
              if (C.refId[s] < 0) continue;
              for (int k = 0; k < K; ++k) out[k + (size_t)K * C.refId[s]] = v[(size_t)k * C.S + s];
            }
          });
        }
        break;
      }
      case EQD_F_FNMS: want(Nn); memcpy(out, h->fnmsH.data(), sizeof(double) * Nn); break;
      case EQD_F_FRIC: {
        need(h->faultSet, "eqd_fetch: no fault");
        want(100 * (size_t)h->nftmx * h->ntotft);
        std::vector<double> fr = h->dFric.download();
        for (int q = 0; q < h->nPairs; ++q) for (int k = 0; k < 100; ++k) out[k + 100 * (size_t)h->pairRef[q]] = fr[(size_t)k * h->PS + q];
        break;
      }
      case EQD_F_FNFT: case EQD_F_ARN: {
        need(h->faultSet, "eqd_fetch: no fault");
        want((size_t)h->nftmx * h->ntotft);
        std::vector<double> v = (which == EQD_F_FNFT ? h->dFnft : h->dArn).download();
        for (int q = 0; q < h->nPairs; ++q) out[h->pairRef[q]] = v[q];
        break;
      }
      case EQD_F_PSTRAIN: {
        want(h->Ne);
        for (int c = 0; c < 3; ++c) {
          ElemClass& C = h->cls[c];
          if (!C.n || !C.pstrain.p) continue;
          std::vector<double> v = C.pstrain.download();
          for (int s = 0; s < C.S; ++s) if (C.refId[s] >= 0) out[C.refId[s]] = v[s];
        }
        break;
      }
      case EQD_F_STRESS: {
        want(h->sizeStress);
        for (int c = 0; c < NCLS; ++c) {
          ElemClass& C = h->cls[c];
          if (!C.n) continue;
          std::vector<double> sg = C.stress.download(), qm;
          if (C.qmem.p) qm = C.qmem.download();
          for (int s = 0; s < C.S; ++s) {
            if (C.refId[s] < 0 || (c == CLS_MARCH && !h->mOwner[s])) continue;
            const int base = h->stressIdx[C.refId[s]];
            for (int k = 0; k < C.nstress; ++k) out[base + k] = sg[(size_t)k * C.S + s];
            if (!qm.empty()) for (int k = 0; k < 6; ++k) out[base + 6 + k] = qm[(size_t)k * C.S + s];
          }
        }
        break;
      }
      case EQD_F_ONFAULT_HIST: {
        finalize(h); want(h->dOnHist.n);
        CK(cudaMemcpy(out, h->dOnHist.p, sizeof(double) * h->dOnHist.n, cudaMemcpyDeviceToHost));
        const size_t row = 12 * (size_t)std::max(h->p.nstep, 1);
        for (int j = 0; j < h->nOn; ++j)
          if (h->stationAlias[j] != j) memcpy(out + row * j, out + row * h->stationAlias[j], sizeof(double) * row);
        break;
      }
      case EQD_F_OFFFAULT_HIST: finalize(h); want(h->dOffHist.n); if (h->dOffHist.n) CK(cudaMemcpy(out, h->dOffHist.p, sizeof(double) * h->dOffHist.n, cudaMemcpyDeviceToHost)); break;
      case EQD_F_HYPO_LOG: finalize(h); want(h->dHypo.n); CK(cudaMemcpy(out, h->dHypo.p, sizeof(double) * h->dHypo.n, cudaMemcpyDeviceToHost)); break;
      case EQD_F_GM: finalize(h); want(3 * (size_t)h->nSurf * h->nGm); if (h->dGm.p && h->nGm) CK(cudaMemcpy(out, h->dGm.p, sizeof(double) * 3 * (size_t)h->nSurf * h->nGm, cudaMemcpyDeviceToHost)); break;
      case EQD_F_SRC_EVOL: finalize(h); need(h->faultSet, "eqd_fetch: no fault"); want((size_t)h->nftnd[0] * h->nGm); if (h->dSrc.p && h->nGm) CK(cudaMemcpy(out, h->dSrc.p, sizeof(double) * (size_t)h->nftnd[0] * h->nGm, cudaMemcpyDeviceToHost)); break;
      case EQD_F_TPHIST: {
        need(h->faultSet && h->dTpHist.p, "eqd_fetch: no thermal-pressurization history (friclaw != 5)");
        const size_t nstep = h->p.nstep;
        want(2 * (size_t)h->nftmx * nstep * h->ntotft);
        std::vector<double> v = h->dTpHist.download();
        for (int q = 0; q < h->nPairs; ++q) {
          const size_t i = h->pairRef[q] % h->nftmx, ift = h->pairRef[q] / h->nftmx;
          for (size_t j = 0; j < nstep; ++j)
            for (int a = 0; a < 2; ++a) out[a + 2 * (i + h->nftmx * (j + nstep * ift))] = v[(j * 2 + a) * h->PS + q];
        }
        break;
      }
      default: throw ArgError("eqd_fetch: unknown which-code");
    }
  });
}

int eqd_get_counts(const eqd_handle* h, int64_t* n_regular, int64_t* n_pml, int64_t* n_pairs, int64_t* launches) {
  if (!h) return EQD_ERR_ARG;
  if (n_regular) *n_regular = h->cls[CLS_REG].n + h->cls[CLS_REGX].n + h->cls[CLS_MARCH].n;
  if (n_pml) *n_pml = h->cls[CLS_PML].n + h->cls[CLS_MARCHP].n;
  if (n_pairs) *n_pairs = h->nPairs;
  if (launches) *launches = h->launches;
  return EQD_OK;
}

int eqd_get_box_counts(const eqd_handle* h, int64_t* n_regular_box, int64_t* n_pml_box) {
  if (!h) return EQD_ERR_ARG;
  if (n_regular_box) *n_regular_box = h->cls[CLS_REG].nBoxElems + h->cls[CLS_REGX].nBoxElems + h->cls[CLS_MARCH].nBoxElems;
  if (n_pml_box) *n_pml_box = h->cls[CLS_PML].nBoxElems + h->cls[CLS_MARCHP].nBoxElems;
  return EQD_OK;
}

int eqd_get_march_counts(const eqd_handle* h, int64_t* out5) {   // out5[0..7]
  if (!h || !out5) return EQD_ERR_ARG;
  out5[0] = h->cls[CLS_MARCH].n; out5[1] = h->mBundles; out5[2] = h->cls[CLS_MARCH].n ? h->cls[CLS_MARCH].PFS : 0;
  out5[3] = h->mFused; out5[4] = h->mGrid;
  out5[5] = h->cls[CLS_MARCHP].n; out5[6] = h->pBundles; out5[7] = h->cls[CLS_MARCHP].n ? h->cls[CLS_MARCHP].PFS : 0;
  return EQD_OK;
}

int eqd_get_halo_mode(const eqd_handle* h) {
  if (!h || !h->finalized || !has_neighbours(h)) return 0;
  return h->p2p.on ? 2 : 1;
}

int eqd_get_timing(const eqd_handle* h, double* ms_slots) {
  if (!h || !ms_slots) return EQD_ERR_ARG;
  for (int k = 0; k < EQD_T_NSLOTS; ++k) ms_slots[k] = h->tms[k];
  return EQD_OK;
}

int eqd_set_option(eqd_handle* h, const char* key, int32_t value) {
  if (!h || !key) return EQD_ERR_ARG;
  if (!strcmp(key, "timing")) { h->timing = value != 0; if (value == 2) { for (double& t : h->tms) t = 0; h->launches = 0; } return EQD_OK; }
  if (!strcmp(key, "overlap")) { h->optOverlap = value; return EQD_OK; }
  {
    const char* names[6] = {"reg_bx", "reg_bz", "reg_by", "pml_bx", "pml_bz", "pml_by"};
    for (int k = 0; k < 6; ++k)
      if (!strcmp(key, names[k])) {
        if (value < 1 || value > 64 || h->meshSet) return EQD_ERR_ARG;
        h->optTile[k / 3][k % 3] = value;
        return EQD_OK;
      }
  }
  if (!strcmp(key, "node_variant")) { h->optNodeVariant = value; return EQD_OK; }
  if (!strcmp(key, "bank_order")) {
    if (h->meshSet) { h->err = "eqd_set_option: bank_order must be set before eqd_set_mesh"; return EQD_ERR_ARG; }
    h->optBankOrder = value < 0 ? 0 : value > 2 ? 2 : value;
    return EQD_OK;
  }
  if (!strcmp(key, "box_compact")) { h->optBoxCompact = value != 0; return EQD_OK; }
  if (!strcmp(key, "march")) {
    if (h->meshSet) { h->err = "eqd_set_option: march must be set before eqd_set_mesh"; return EQD_ERR_ARG; }
    h->optMarch = value < 0 ? 0 : value > 3 ? 3 : value;
    return EQD_OK;
  }
  if (!strcmp(key, "box")) {
    if (h->finalized) { h->err = "eqd_set_option: box must be set before the first eqd_run / eqd_sum_shared"; return EQD_ERR_ARG; }
    h->optBox = value < 0 ? 0 : value > 2 ? 2 : value;
    return EQD_OK;
  }
  if (!strcmp(key, "reserve")) { h->optReserve = value; return EQD_OK; }
  if (!strcmp(key, "halo")) {
    if (h->finalized) { h->err = "eqd_set_option: halo must be set before the first eqd_run / eqd_sum_shared"; return EQD_ERR_ARG; }
    h->optHalo = value;
    return EQD_OK;
  }
  h->err = std::string("eqd_set_option: unknown key ") + key;
  return EQD_ERR_ARG;
}

}  // extern "C"
