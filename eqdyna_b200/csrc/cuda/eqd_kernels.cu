// Hand-written FP64 sm_100a kernels of the EQdyna step loop (src/driver.f90:9-34).
// Every kernel cites the reference routine it replaces.  One thread per work
// item (element / node / split-node pair), SoA operands with the work-item
// index fastest so every warp-wide access is a run of consecutive doubles;
// element kernels work tile by tile (one CTA = one brick of elements) and
// assemble nodal forces in shared memory in a fixed order.
#include <cuda_runtime.h>
#include <math.h>

#include <algorithm>
#include <map>
#include <mutex>
#include <set>
#include <utility>

#include "eqd_box.h"
#include "eqd_dev.cuh"
#include "eqd_kernels.h"

namespace eqd {

__constant__ QTab c_qtab[16];

void upload_qtab(const QTab* t) { cudaMemcpyToSymbol(c_qtab, t, sizeof(QTab) * 16); }

#define LDG(p) __ldg(p)

// loads the compiler must issue where they are written (it otherwise sinks them
// below early exits and branches, which turns independent DRAM round trips into a chain)
__device__ __forceinline__ double ld_now(const double* p) { double v; asm volatile("ld.global.f64 %0, [%1];" : "=d"(v) : "l"(p)); return v; }
__device__ __forceinline__ double ldc_now(const double* p) { double v; asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(p)); return v; }
__device__ __forceinline__ uint32_t ldc_now(const uint32_t* p) { uint32_t v; asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(p)); return v; }
__device__ __forceinline__ int ldc_now(const int* p) { int v; asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(v) : "l"(p)); return v; }
__device__ __forceinline__ int ldc_now(const uint8_t* p) { uint32_t v; asm volatile("ld.global.nc.u8 %0, [%1];" : "=r"(v) : "l"(p)); return (int)v; }

// ----------------------------------------------------------------------------
// step counter: driver.f90:11 (timeElapsed = timeElapsed + dt)
__global__ void k_advance(StepState* st, double dt) {
  st->timeElapsed = st->timeElapsed + dt;
  st->nt = st->nt + 1;
}

// ----------------------------------------------------------------------------
// Ordered gather of a node's force from the tile partials (replaces the
// scatter-add of assembleGlobalKU.f90:28-35,47-63 and hrglss.f90:42-54: same
// contributions; fixed order = 8 local-node phases inside a tile, then the
// node's tiles by class and ascending tile id).
__device__ __forceinline__ void add3(const NodeArgs& A, uint32_t u, double f[3]) {
  const int cls = u & 3;
  const size_t idx = u >> 2;
  if (cls == CLS_REG) {
    const double* p = A.pfR + idx;
    f[0] += p[0]; f[1] += p[A.SR]; f[2] += p[2 * (size_t)A.SR];
  } else if (cls == CLS_MARCH) {
    const double* p = A.pfM + idx;
    f[0] += p[0]; f[1] += p[A.SM]; f[2] += p[2 * (size_t)A.SM];
  } else if (cls == CLS_PML) {
    // 3-dof node of a PML element: assembleGlobalKU.f90:55-61
    const double* p = A.pfP + idx;
    const size_t S = A.SP;
    f[0] = f[0] + p[0] + p[S] + p[2 * S] + p[9 * S];
    f[1] = f[1] + p[3 * S] + p[4 * S] + p[5 * S] + p[10 * S];
    f[2] = f[2] + p[6 * S] + p[7 * S] + p[8 * S] + p[11 * S];
  } else {
    const double* p = A.pfX + idx;
    const size_t S = A.SX;
    f[0] = f[0] + p[0] + p[3 * S];
    f[1] = f[1] + p[S] + p[4 * S];
    f[2] = f[2] + p[2 * S] + p[5 * S];
  }
}
// The slot table is stored by rank: slotTab[k][n] = k-th tile-node slot of node n,
// so a warp of consecutive nodes reads consecutive words and the first two ranks
// can be fetched together with `info` (one level of the dependent-load chain less).
__device__ __forceinline__ void gather3(const NodeArgs& A, int n, int cnt, uint32_t u0, uint32_t u1, double f[3]) {
  f[0] = f[1] = f[2] = 0.0;
  if (cnt > 0) add3(A, u0, f);
  if (cnt > 1) add3(A, u1, f);
  for (int k = 2; k < cnt; ++k) add3(A, LDG(A.slotTab + (size_t)k * A.NnS + n), f);
}
__device__ __forceinline__ void gather3(const NodeArgs& A, int n, int cnt, double f[3]) {
  f[0] = f[1] = f[2] = 0.0;
  for (int k = 0; k < cnt; ++k) add3(A, LDG(A.slotTab + (size_t)k * A.NnS + n), f);
}

__device__ __forceinline__ void gather12(const NodeArgs& A, int n, int cnt, double f[12]) {
#pragma unroll
  for (int j = 0; j < 12; ++j) f[j] = 0.0;
  for (int k = 0; k < cnt; ++k) {
    const uint32_t u = LDG(A.slotTab + (size_t)k * A.NnS + n);
    const int cls = u & 3;
    const size_t idx = u >> 2;
    if (cls == CLS_PML) {
      const double* p = A.pfP + idx;
#pragma unroll
      for (int j = 0; j < 12; ++j) f[j] += p[(size_t)j * A.SP];
    } else if (cls == CLS_REGX) {
      // regular element on a 12-dof node: KU goes to dofs 1-3
      // (assembleGlobalKU.f90:28-35), hourglass to dofs 10-12 (hrglss.f90:44-48)
      const double* p = A.pfX + idx;
      const size_t S = A.SX;
      f[0] += p[0]; f[1] += p[S]; f[2] += p[2 * S];
      f[9] += p[3 * S]; f[10] += p[4 * S]; f[11] += p[5 * S];
    }
  }
}

// ----------------------------------------------------------------------------
// velDispUpdate (driver.f90:89-155) fused with the force assembly and with the
// previous step's `nodalForceArr/nodalMassArr` (driver.f90:29).  Two kernels so
// that the 3-dof one (almost every node) stays light enough for full occupancy:
// its gather is a chain of dependent loads that only parallelism hides.
// LIST: one thread per entry of A.list (the free 3-dof nodes that no marching bundle updates itself) instead of
// one per node; used after a sweep that updated the fused nodes (eqd_march.h).
template <bool SKIP, bool LIST, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) k_node_update3(NodeArgs A) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (LIST ? A.nList : A.Nn)) return;
  const int n = LIST ? ldc_now(A.list + i) : i;
  const size_t NS = A.NnS;
  // every independent load first: the stores below may alias them for the compiler,
  // which would otherwise serialise six DRAM round trips per node
  const int info = ldc_now(A.info + n);
  const uint32_t u0 = ldc_now(A.slotTab + n), u1 = ldc_now(A.slotTab + NS + n);
  const int cnt0 = ldc_now(A.slotCnt + n);
  const double m = ldc_now(A.mass + n);
  double v[3], d[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) { v[j] = ld_now(A.vel + j * NS + n); d[j] = ld_now(A.disp + j * NS + n); }
  // no early exit: fixed and PML nodes run through with an empty gather and skip the
  // stores (vel = disp = 0 for ever for fixed nodes, driver.f90:142-145)
  const bool fused = !LIST && EQD_INFO_KIND(info) == KIND_FREE3 && EQD_INFO_FUSED(info);
  const bool mine = EQD_INFO_KIND(info) == KIND_FREE3 && !(SKIP && EQD_INFO_SPECIAL(info)) && !(fused && A.fusedMode == 2);
  const int cnt = mine ? cnt0 : 0;
  const double dt = A.dt;
  double a[3] = {0.0, 0.0, 0.0};
  if (A.accel0) {
    if (mine) { a[0] = A.accel0[n]; a[1] = A.accel0[NS + n]; a[2] = A.accel0[2 * NS + n]; }
  } else {
    double f[3];
    if (mine && (fused || (!SKIP && EQD_INFO_SPECIAL(info)))) {
      f[0] = A.force[n]; f[1] = A.force[NS + n]; f[2] = A.force[2 * NS + n];
    } else {
      gather3(A, n, cnt, u0, u1, f);
    }
    a[0] = f[0] / m; a[1] = f[1] / m; a[2] = f[2] / m;
  }
  bool bad = false;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    v[j] = v[j] + a[j] * dt;                // driver.f90:102
    d[j] = d[j] + v[j] * dt;                // :104
    bad |= (v[j] != v[j]);
  }
  if (mine) {
#pragma unroll
    for (int j = 0; j < 3; ++j) { A.velOut[j * NS + n] = v[j]; A.dispOut[j * NS + n] = d[j]; }  // :103
    if (bad) {
      if (atomicExch(&A.st->nanFlag, 1) == 0) A.st->nanNode = n + 1;
    }
  }
}

// 12-dof PML nodes: split-field velocities with damping (driver.f90:105-141), one thread per PML slot
template <bool SKIP, int MINB>
__global__ void __launch_bounds__(128, MINB) k_node_update12(NodeArgs A) {
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= A.Np) return;
  const size_t NS = A.NnS, PS = A.NpS;
  const int n = ldc_now(A.pmlNode + slot);
  double v1[12], d[3];
#pragma unroll
  for (int j = 0; j < 12; ++j) v1[j] = ld_now(A.v1p + j * PS + slot);
  d[0] = ldc_now(A.dampp + slot); d[1] = ldc_now(A.dampp + PS + slot); d[2] = ldc_now(A.dampp + 2 * PS + slot);
  const int info = ldc_now(A.info + n);
  const int cnt = ldc_now(A.slotCnt + n);
  const double m = ldc_now(A.mass + n);
  if (SKIP && EQD_INFO_SPECIAL(info)) return;
  double dis[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) dis[j] = ld_now(A.disp + j * NS + n);
  const double dt = A.dt;
  double a[12];
  if (A.accel0) {
    const double* ap = A.accel0 + 3 * NS;
#pragma unroll
    for (int j = 0; j < 12; ++j) a[j] = ap[j * PS + slot];
  } else {
    double f[12];
    if (!SKIP && EQD_INFO_SPECIAL(info)) {
      const double* fp = A.force + 3 * NS;
#pragma unroll
      for (int j = 0; j < 12; ++j) f[j] = fp[j * PS + slot];
    } else {
      gather12(A, n, cnt, f);
    }
#pragma unroll
    for (int j = 0; j < 12; ++j) a[j] = f[j] / m;
  }
  double v[12];
  const double rdt = 1.0 / dt;
#pragma unroll
  for (int j = 0; j < 9; ++j) {             // driver.f90:112-117, dampv(j) = damp(mod(j-1,3)+1)
    const double dj = d[j % 3];
    v[j] = (a[j] + v1[j] * (rdt - dj / 2.0)) / (rdt + dj / 2.0);
  }
#pragma unroll
  for (int j = 9; j < 12; ++j) v[j] = v1[j] + a[j] * dt;   // :118-123
#pragma unroll
  for (int j = 0; j < 12; ++j) A.v1p[j * PS + slot] = v[j];
  double vel[3];
  vel[0] = v[0] + v[1] + v[2] + v[9];       // :125-141
  vel[1] = v[3] + v[4] + v[5] + v[10];
  vel[2] = v[6] + v[7] + v[8] + v[11];
  bool bad = false;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    A.velOut[j * NS + n] = vel[j];
    A.dispOut[j * NS + n] = dis[j] + vel[j] * dt;
    bad |= (vel[j] != vel[j]);
  }
  if (bad) {
    if (atomicExch(&A.st->nanFlag, 1) == 0) A.st->nanNode = n + 1;
  }
}

// The special nodes (split-node pairs, rank-face nodes) of the same update: their
// force comes from force[] / forcep[], which the halo and the fault solver of the
// previous step finish on the communication stream.  Same arithmetic as above.
__global__ void __launch_bounds__(128) k_node_update_special(NodeArgs A, const int* __restrict__ list, int nList) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nList) return;
  const int n = list[i];
  const int info = LDG(A.info + n);
  const int kind = EQD_INFO_KIND(info);
  const size_t NS = A.NnS, PS = A.NpS;
  const double dt = A.dt;
  const double m = LDG(A.mass + n);
  bool bad = false;
  if (kind == KIND_FREE3) {
    double a[3];
    if (A.accel0) { a[0] = A.accel0[n]; a[1] = A.accel0[NS + n]; a[2] = A.accel0[2 * NS + n]; }
    else { a[0] = A.force[n] / m; a[1] = A.force[NS + n] / m; a[2] = A.force[2 * NS + n] / m; }
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      double v = A.vel[j * NS + n], d = A.disp[j * NS + n];
      v = v + a[j] * dt;                      // driver.f90:102
      d = d + v * dt;                         // :104
      bad |= (v != v);
      A.velOut[j * NS + n] = v; A.dispOut[j * NS + n] = d;
    }
  } else if (kind == KIND_PML12) {
    const int slot = EQD_INFO_SLOT(info);
    double a[12], v[12];
    const double* src = (A.accel0 ? A.accel0 : A.force) + 3 * NS;
#pragma unroll
    for (int j = 0; j < 12; ++j) a[j] = A.accel0 ? src[j * PS + slot] : src[j * PS + slot] / m;
    const double dmp[3] = {LDG(A.dampp + slot), LDG(A.dampp + PS + slot), LDG(A.dampp + 2 * PS + slot)};
    const double rdt = 1.0 / dt;
#pragma unroll
    for (int j = 0; j < 9; ++j) {             // driver.f90:112-117
      const double dj = dmp[j % 3];
      v[j] = (a[j] + A.v1p[j * PS + slot] * (rdt - dj / 2.0)) / (rdt + dj / 2.0);
    }
#pragma unroll
    for (int j = 9; j < 12; ++j) v[j] = A.v1p[j * PS + slot] + a[j] * dt;   // :118-123
#pragma unroll
    for (int j = 0; j < 12; ++j) A.v1p[j * PS + slot] = v[j];
    double vel[3];
    vel[0] = v[0] + v[1] + v[2] + v[9];       // :125-141
    vel[1] = v[3] + v[4] + v[5] + v[10];
    vel[2] = v[6] + v[7] + v[8] + v[11];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      A.velOut[j * NS + n] = vel[j];
      A.dispOut[j * NS + n] = A.disp[j * NS + n] + vel[j] * dt;
      bad |= (vel[j] != vel[j]);
    }
  }
  if (bad) {
    if (atomicExch(&A.st->nanFlag, 1) == 0) A.st->nanNode = n + 1;
  }
}

// materialise the ordered force sums of the "special" nodes
__global__ void __launch_bounds__(128) k_assemble_special(NodeArgs A, const int* __restrict__ list, int nList) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nList) return;
  const int n = list[i];
  const int info = LDG(A.info + n);
  const int kind = EQD_INFO_KIND(info);
  const size_t NS = A.NnS;
  if (kind == KIND_FREE3) {
    double f[3];
    gather3(A, n, LDG(A.slotCnt + n), f);
    A.force[n] = f[0]; A.force[NS + n] = f[1]; A.force[2 * NS + n] = f[2];
  } else if (kind == KIND_PML12) {
    double f[12];
    gather12(A, n, LDG(A.slotCnt + n), f);
    const int slot = EQD_INFO_SLOT(info);
    double* fp = A.force + 3 * NS;
#pragma unroll
    for (int j = 0; j < 12; ++j) fp[(size_t)j * A.NpS + slot] = f[j];
  }
}

// nodalForceArr as the reference leaves it after driver.f90:29 (f/m), for eqd_fetch
__global__ void k_materialize_accel(NodeArgs A, double* __restrict__ out) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= A.Nn) return;
  const int info = LDG(A.info + n);
  const int kind = EQD_INFO_KIND(info);
  const size_t NS = A.NnS;
  const double m = LDG(A.mass + n);
  if (kind == KIND_FREE3) {
    double f[3];
    if (A.accel0) { out[n] = A.accel0[n]; out[NS + n] = A.accel0[NS + n]; out[2 * NS + n] = A.accel0[2 * NS + n]; return; }
    if (EQD_INFO_SPECIAL(info) || EQD_INFO_FUSED(info)) { f[0] = A.force[n]; f[1] = A.force[NS + n]; f[2] = A.force[2 * NS + n]; }
    else gather3(A, n, LDG(A.slotCnt + n), f);
    out[n] = f[0] / m; out[NS + n] = f[1] / m; out[2 * NS + n] = f[2] / m;
  } else if (kind == KIND_PML12) {
    const int slot = EQD_INFO_SLOT(info);
    const size_t PS = A.NpS;
    double f[12];
    if (A.accel0) {
      for (int j = 0; j < 12; ++j) out[3 * NS + j * PS + slot] = A.accel0[3 * NS + j * PS + slot];
      return;
    }
    if (EQD_INFO_SPECIAL(info)) { for (int j = 0; j < 12; ++j) f[j] = A.force[3 * NS + j * PS + slot]; }
    else gather12(A, n, LDG(A.slotCnt + n), f);
    for (int j = 0; j < 12; ++j) out[3 * NS + j * PS + slot] = f[j] / m;
  }
}

// ----------------------------------------------------------------------------
// ---- async-proxy helpers (sm_90+): 1-D bulk copies global -> shared that
// signal an mbarrier with their byte count (SASS: UBLKCP / SYNCS)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// named barrier over the `nthreads` threads of one role (ids 1..4; 0 is __syncthreads): the roles of a
// tile kernel accumulate into disjoint force rows, so an assembly phase only has to order the warps
// of its own role
__device__ __forceinline__ void role_sync(int role, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(role + 1), "r"(nthreads) : "memory");
}

// per-thread asynchronous 8-byte copy global -> shared (LDGSTS): gathers that need no register staging
__device__ __forceinline__ void cp_async8(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// ----------------------------------------------------------------------------
// Tile kernels.  One CTA owns one tile (a brick of a few hundred elements of one
// class, contiguous in the class's SoA) and walks it in stages of EQD_STAGE
// elements:
//   stream   the stage's operator rows (shape derivatives, hourglass vectors
//            and stiffness, moduli, stresses, local connectivity) arrive in
//            shared memory through bulk asynchronous copies signalling an
//            mbarrier; the next stage is requested as soon as the current one
//            has been consumed, so HBM latency hides behind the assembly phases
//            and the second CTA of the SM
//   stage    the tile's nodes: v and d + rdampk*v (hrglss.f90:20-27) -> shared
//   sweep    two threads per element: warps 0-3 do the constitutive part
//            (calcElemKU), warps 4-7 the hourglass part (hrglss) of the same
//            EQD_STAGE elements
//   assemble 8 local-node phases; in phase i every element adds its force on
//            local node i to the tile's shared force rows (one set per role).
//            Two elements of one phase never touch the same node in a
//            structured hexahedral brick; where they do (wedge pairs) the host
//            gave them different colours and the phase is split.  Plain
//            read-modify-write, no atomics: the summation order is fixed.
//   flush    one partial force row per tile node -> pf (coalesced)
// tile record (int4): first element slot, elements | colours << 16, first tile-node slot, tile nodes
#define TR_E0(r) ((r).x)
#define TR_NE(r) ((r).y & 0xffff)
#define TR_NC(r) ((r).y >> 16)
#define TR_NB(r) ((r).z)
#define TR_LN(r) ((r).w)

// operator rows of one regular-element stage, in shared-memory order
enum { RR_SHP = 0, RR_PHI = 24, RR_SS = 56, RR_LAM = 62, RR_MU = 63, RR_DET = 64, RR_STRESS = 65, RR_ROWS = 71 };
__device__ __forceinline__ const double* reg_row_src(const ElemArgs& A, int r) {
  const size_t S = A.S;
  if (r < RR_PHI) return A.shp + (size_t)r * S;
  if (r < RR_SS) return A.phi + (size_t)(r - RR_PHI) * S;
  if (r < RR_LAM) return A.ss + (size_t)(r - RR_SS) * S;
  if (r == RR_LAM) return A.lam;
  if (r == RR_MU) return A.mu;
  if (r == RR_DET) return A.det;
  return A.stress + (size_t)(r - RR_STRESS) * S;
}

// Regular hexahedron / degenerate wedge: calcElemKU.f90:3-191 (+ calcB.f90,
// calcElemMass.f90) and hrglss.f90:13-98.  The 6x24 B matrix is never
// materialised.
// BOX: tiles flagged in A.tileBox hold only axis-aligned hexahedra; their shape
// derivatives, hourglass vectors and hourglass stiffness are used in closed form
// (eqd_box.h) and only BOX_ROWS of the RR_ROWS operator rows are streamed.
// BOX == 1: per-tile flag (meshes with warped regions); BOX == 2: every tile of the
// launch is a box tile -- the stage buffer holds only the BOX_ROWS rows, so three
// CTAs fit an SM (A.allBox, option "box_compact").
template <bool PLASTIC, bool QMODE, bool BODY, bool SPLIT, int CHG, int BOX>
__global__ void __launch_bounds__(2 * EQD_STAGE, BOX == 2 ? 3 : 2) k_tile_reg(ElemArgs A) {
  extern __shared__ __align__(128) unsigned char smraw[];
  constexpr int SE = EQD_STAGE;
  constexpr int NT = 2 * SE;
  constexpr int LS = EQD_REG_LS;    // shared-memory row stride = node cap of a regular tile
  constexpr int NPT = (LS + NT - 1) / NT;
  const int tid = threadIdx.x;
  const int role = tid / SE;        // 0: constitutive (KU), 1: hourglass; warp-uniform
  const int lane = tid - role * SE; // element of the stage
  const int tEnd = A.tile0 + A.ntiles;
  int tnext = A.tile0 + blockIdx.x;
  if (tnext >= tEnd) return;
  const size_t S = A.S, NS = A.NnS;
  const double dt = A.dt, rdampk = A.rdampk;
  uint64_t* bar = (uint64_t*)smraw;                             // operator stage landed
  uint64_t* barN = bar + 1;                                     // node-id list of the next tile landed
  constexpr int NROWS = BOX == 2 ? (int)BOX_ROWS : (int)RR_ROWS;
#define ISBOX(flag) (BOX == 2 || (BOX == 1 && (flag)))
  double* ops = (double*)(smraw + 128);                         // [NROWS][SE]
  uint16_t* slc = (uint16_t*)(ops + NROWS * SE);                // [8][SE]
  int* tnS = (int*)(slc + 8 * SE);                              // [LS] node ids of the tile whose nodes are gathered next
  double* sv = (double*)(tnS + LS);                             // [3][LS] velocity
  double* sl = sv + 3 * LS;                                     // [3][LS] d + rdampk*v
  double* sd = QMODE ? sl + 3 * LS : sl;                        // [3][LS] displacement: gather target (kept for Q)
  double* sf = sl + (QMODE ? 6 : 3) * LS;                       // [2][3][LS] force accumulators: KU | hourglass
  auto request = [&](const int4& rec, int base, bool bx) {
    // one thread per row; sizes are whole 32-element groups (the class SoA is padded)
    const int cnt = min(SE, (TR_NE(rec) - base + 31) & ~31);
    if (tid == 0) mbar_expect_tx(bar, (uint32_t)(cnt * ((ISBOX(bx) ? (int)BOX_ROWS : (int)RR_ROWS) * 8 + 8 * 2)));
    __syncwarp();
    if (tid < RR_ROWS) { if (!ISBOX(bx) || box_row(tid)) bulk_g2s(ops + (BOX == 2 ? box_slot(tid) : tid) * SE, reg_row_src(A, tid) + TR_E0(rec) + base, cnt * 8, bar); }
    else if (tid < RR_ROWS + 8) bulk_g2s(slc + (tid - RR_ROWS) * SE, A.lconn + (size_t)(tid - RR_ROWS) * S + TR_E0(rec) + base, cnt * 2, bar);
  };
  auto request_ids = [&](const int4& rec) {
    if (tid == 0) {
      mbar_expect_tx(barN, (uint32_t)(TR_LN(rec) * 4));
      bulk_g2s(tnS, A.tnode + TR_NB(rec), TR_LN(rec) * 4, barN);
    }
  };
  // gather the nodal values of tile `rec` (ids in tnS) straight into shared memory
  auto gather_nodes = [&](const int4& rec) {
#pragma unroll
    for (int k = 0; k < NPT; ++k) {
      const int i = tid + k * NT;
      if (i < LS) {
        const int nd = i < TR_LN(rec) ? tnS[i] : -1;
        if (nd >= 0) {
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            cp_async8(sv + j * LS + i, A.vel + j * NS + nd);
            cp_async8(sd + j * LS + i, A.disp + j * NS + nd);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 3; ++j) { sv[j * LS + i] = 0.0; sd[j * LS + i] = 0.0; }
        }
      }
    }
    cp_async_commit();
  };
  // own nodes have landed: d -> d + rdampk*v (hrglss.f90:20-27)
  auto finish_nodes = [&]() {
    cp_async_wait_all();
#pragma unroll
    for (int k = 0; k < NPT; ++k) {
      const int i = tid + k * NT;
      if (i < LS) {
#pragma unroll
        for (int j = 0; j < 3; ++j) sl[j * LS + i] = sd[j * LS + i] + rdampk * sv[j * LS + i];
      }
    }
  };
  const int4 zero4 = make_int4(0, 0, 0, 0);
  // software pipeline over this CTA's tiles: rc = tile being swept, rn = the next
  // one (its operators and nodes are requested while rc's forces are assembled)
  int4 rc = __ldg(A.tileRec + tnext);
  bool bc = BOX == 1 && __ldg(A.tileBox + tnext) != 0;   // CTA-uniform
  tnext += gridDim.x;
  int4 rn = tnext < tEnd ? __ldg(A.tileRec + tnext) : zero4;
  bool bn = BOX == 1 && tnext < tEnd && __ldg(A.tileBox + tnext) != 0;
  if (tid == 0) { mbar_init(bar, 1); mbar_init(barN, 1); }
  __syncthreads();
  request(rc, 0, bc);
  request_ids(rc);
#pragma unroll
  for (int k = 0; k < NPT; ++k) {
    const int i = tid + k * NT;
    if (i < LS) {
#pragma unroll
      for (int j = 0; j < 6; ++j) sf[j * LS + i] = 0.0;
    }
  }
  uint32_t parity = 0, parityN = 0;
  mbar_wait(barN, parityN);
  parityN ^= 1;
  gather_nodes(rc);
  finish_nodes();
  __syncthreads();
  double* acc = sf + role * 3 * LS;
  while (true) {
  const bool more = tnext < tEnd;
  const int4 rnn = (tnext + (int)gridDim.x < tEnd) ? __ldg(A.tileRec + tnext + gridDim.x) : zero4;
  const bool bnn = BOX == 1 && (tnext + (int)gridDim.x < tEnd) && __ldg(A.tileBox + tnext + gridDim.x) != 0;
  if (more) request_ids(rn);   // tnS is free: rc's nodes are already in sv/sl
  const int ne = TR_NE(rc), NC = TR_NC(rc);
  for (int base = 0; base < ne; base += SE) {
    const int le = base + lane;
    const bool act = le < ne;
    const size_t e = (size_t)TR_E0(rc) + (act ? le : 0);
    unsigned lc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    double f[8][3];
    mbar_wait(bar, parity);
    parity ^= 1;
    if (act) {
#pragma unroll
      for (int i = 0; i < 8; ++i) lc[i] = slc[i * SE + lane];
#define OP(r) ops[(BOX == 2 ? box_slot(r) : (r)) * SE + lane]
      if (role == 0) {
        // ---------------- constitutive part: calcElemKU.f90
        double sr[6] = {0, 0, 0, 0, 0, 0}, sn[6] = {0, 0, 0, 0, 0, 0};
        double body[BODY ? 24 : 1];
        if (ISBOX(bc)) {
          // axis-aligned hexahedra: eleshp(d,i) = sign_d(i)*a_d (eqd_box.h)
          double u[8][3], g[3][3];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int li = lc[i] & EQD_LN_MASK;
            u[i][0] = sv[li]; u[i][1] = sv[LS + li]; u[i][2] = sv[2 * LS + li];
            if (BODY) {
              const double em = LDG(A.emass + i * S + e);
              body[3 * i] = 0.0 - (A.rdampm * u[i][0]) * em;
              body[3 * i + 1] = 0.0 - (A.rdampm * u[i][1]) * em;
              body[3 * i + 2] = 0.0 - (A.rdampm * u[i][2] + A.bodyz) * em;
            }
          }
          box_grad(u, g);
          box_strain(g, OP(BOX_AX), OP(BOX_AY), OP(BOX_AZ), sr);
          if (QMODE) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int li = lc[i] & EQD_LN_MASK;
              u[i][0] = sd[li]; u[i][1] = sd[LS + li]; u[i][2] = sd[2 * LS + li];
            }
            box_grad(u, g);
            box_strain(g, OP(BOX_AX), OP(BOX_AY), OP(BOX_AZ), sn);
          }
        } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int li = lc[i] & EQD_LN_MASK;
          const double vx = sv[li], vy = sv[LS + li], vz = sv[2 * LS + li];
          const double s1 = OP(RR_SHP + 3 * i), s2 = OP(RR_SHP + 3 * i + 1), s3 = OP(RR_SHP + 3 * i + 2);
          // calcElemKU.f90:44-60 (B with engineering shear, calcB.f90:10-25)
          sr[0] = sr[0] + s1 * vx;
          sr[1] = sr[1] + s2 * vy;
          sr[2] = sr[2] + s3 * vz;
          sr[3] = sr[3] + s3 * vy + s2 * vz;
          sr[4] = sr[4] + s3 * vx + s1 * vz;
          sr[5] = sr[5] + s2 * vx + s1 * vy;
          if (QMODE) {
            const double dx = sd[li], dy = sd[LS + li], dz = sd[2 * LS + li];
            sn[0] = sn[0] + s1 * dx;
            sn[1] = sn[1] + s2 * dy;
            sn[2] = sn[2] + s3 * dz;
            sn[3] = sn[3] + s3 * dy + s2 * dz;
            sn[4] = sn[4] + s3 * dx + s1 * dz;
            sn[5] = sn[5] + s2 * dx + s1 * dy;
          }
          if (BODY) {
            // assembleGlobalKU.f90:15-16 + calcElemMass.f90: elresf = -al*elemass
            const double em = LDG(A.emass + i * S + e);
            body[3 * i] = 0.0 - (A.rdampm * vx) * em;
            body[3 * i + 1] = 0.0 - (A.rdampm * vy) * em;
            body[3 * i + 2] = 0.0 - (A.rdampm * vz + A.bodyz) * em;
          }
        }
        }
        const double lam = OP(RR_LAM), mu = OP(RR_MU);
        const double l2m = lam + 2 * mu;
        double rate[6];
        // calcElemKU.f90:63-70
        rate[0] = 0.0 + l2m * sr[0] + lam * sr[1] + lam * sr[2];
        rate[1] = 0.0 + lam * sr[0] + l2m * sr[1] + lam * sr[2];
        rate[2] = 0.0 + lam * sr[0] + lam * sr[1] + l2m * sr[2];
        rate[3] = mu * sr[3];
        rate[4] = mu * sr[4];
        rate[5] = mu * sr[5];
        double sg[6];
        if (!QMODE) {
#pragma unroll
          for (int i = 0; i < 6; ++i) sg[i] = OP(RR_STRESS + i) + rate[i] * dt;  // :72-76
        } else {
          // calcElemKU.f90:77-132, constants tabulated per class (qconstant.f90)
          const QTab q = c_qtab[A.qcls[e]];
          const double miuu = mu * q.cs, Mu = l2m * q.cv;
          const double vols = sn[0] + sn[1] + sn[2];
          const double ex = q.expdt;
#pragma unroll
          for (int i = 0; i < 6; ++i) {
            const double a1 = A.qmem[i * S + e];
            double an;
            if (i < 3) an = ex * a1 + (1 - ex) * (2 * miuu * sn[i] * q.wks + (Mu * q.wkp - 2 * miuu * q.wks) * vols);
            else an = ex * a1 + (1 - ex) * (miuu * sn[i] * q.wks);
            A.qmem[i * S + e] = an;
            if (i < 3) sg[i] = 2.0 * miuu * sn[i] + (Mu - 2.0 * miuu) * vols - 0.5 * (an + a1);
            else sg[i] = 2.0 * miuu * sn[i] / 2.0 - 0.5 * (an + a1);
          }
        }
        if (PLASTIC) {
          // Drucker-Prager viscoplasticity, calcElemKU.f90:133-167
          const double strmea = (sg[0] + sg[1] + sg[2]) / 3.0;
          double dv[6] = {sg[0] - strmea, sg[1] - strmea, sg[2] - strmea, sg[3], sg[4], sg[5]};
          double taomax = 0.5 * (dv[0] * dv[0] + dv[1] * dv[1] + dv[2] * dv[2]) + dv[3] * dv[3] + dv[4] * dv[4] + dv[5] * dv[5];
          taomax = sqrt(taomax);
          double yield = A.ccosphi - A.sinphi * (strmea + LDG(A.porep + e));
          if (yield < 0.0) yield = 0.0;
          if (taomax > yield) {
            const double rjust = yield / taomax + (1 - yield / taomax) * A.expdttv;
            double pi[6];
#pragma unroll
            for (int i = 0; i < 6; ++i) {
              sg[i] = dv[i] * rjust;
              pi[i] = (dv[i] - sg[i]) / mu;
              if (i < 3) sg[i] = sg[i] + strmea;
            }
            const double pm = (pi[0] + pi[1] + pi[2]) / 3.0;
#pragma unroll
            for (int i = 0; i < 6; ++i) pi[i] = pi[i] - pm;
            double mag = 0.5 * (pi[0] * pi[0] + pi[1] * pi[1] + pi[2] * pi[2]) + pi[3] * pi[3] + pi[4] * pi[4] + pi[5] * pi[5];
            A.pstrain[e] = A.pstrain[e] + sqrt(mag);   // assembleGlobalKU.f90:26
          }
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) A.stress[i * S + e] = sg[i];
        // calcElemKU.f90:169-173, constk = -eledet
        const double temp = (-OP(RR_DET)) * A.w;
        double t[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) t[i] = temp * (sg[i] + rdampk * rate[i]);
        if (ISBOX(bc)) {
          box_force(t, OP(BOX_AX), OP(BOX_AY), OP(BOX_AZ), f);
          if (BODY) {
#pragma unroll
            for (int i = 0; i < 8; ++i) { f[i][0] = body[3 * i] + f[i][0]; f[i][1] = body[3 * i + 1] + f[i][1]; f[i][2] = body[3 * i + 2] + f[i][2]; }
          }
        } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const double s1 = OP(RR_SHP + 3 * i), s2 = OP(RR_SHP + 3 * i + 1), s3 = OP(RR_SHP + 3 * i + 2);
          // calcElemKU.f90:175-189
          double f0 = s1 * t[0] + s3 * t[4] + s2 * t[5];
          double f1 = s2 * t[1] + s3 * t[3] + s1 * t[5];
          double f2 = s3 * t[2] + s2 * t[3] + s1 * t[4];
          if (BODY) { f0 = body[3 * i] + f0; f1 = body[3 * i + 1] + f1; f2 = body[3 * i + 2] + f2; }
          f[i][0] = f0; f[i][1] = f1; f[i][2] = f2;
        }
        }
      } else {
        // ---------------- hourglass part: hrglss.f90
        if (CHG == 1) {
          if (ISBOX(bc)) {
            // phi = ha, ss diagonal (eqd_box.h)
            double l[8][3];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int li = lc[i] & EQD_LN_MASK;
              l[i][0] = sl[li]; l[i][1] = sl[LS + li]; l[i][2] = sl[2 * LS + li];
            }
            box_hourglass(l, OP(BOX_SS0), OP(BOX_SS3), OP(BOX_SS5), f);
          } else {
          double phid[4][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int li = lc[i] & EQD_LN_MASK;
            // hrglss.f90:20-33: dl = d + rdampk*v ; phid = sum_j phi(j,m)*dl(j)
            const double lx = sl[li], ly = sl[LS + li], lz = sl[2 * LS + li];
#pragma unroll
            for (int m = 0; m < 4; ++m) {
              const double p = OP(RR_PHI + 8 * m + i);
              phid[m][0] = phid[m][0] + p * lx;
              phid[m][1] = phid[m][1] + p * ly;
              phid[m][2] = phid[m][2] + p * lz;
            }
          }
          double ssv[6], hv[4][3];
#pragma unroll
          for (int i = 0; i < 6; ++i) ssv[i] = OP(RR_SS + i);
#pragma unroll
          for (int m = 0; m < 4; ++m) {  // hrglss.f90:35-40
            hv[m][0] = ssv[0] * phid[m][0] + ssv[1] * phid[m][1] + ssv[2] * phid[m][2];
            hv[m][1] = ssv[1] * phid[m][0] + ssv[3] * phid[m][1] + ssv[4] * phid[m][2];
            hv[m][2] = ssv[2] * phid[m][0] + ssv[4] * phid[m][1] + ssv[5] * phid[m][2];
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            double h0 = 0.0, h1 = 0.0, h2 = 0.0;
#pragma unroll
            for (int m = 0; m < 4; ++m) {  // hrglss.f90:41-54: force -= phi(i,m)*(SS.phid)
              const double p = OP(RR_PHI + 8 * m + i);
              h0 = h0 - p * hv[m][0];
              h1 = h1 - p * hv[m][1];
              h2 = h2 - p * hv[m][2];
            }
            f[i][0] = h0; f[i][1] = h1; f[i][2] = h2;
          }
          }
        } else {
          // viscous hourglass, hrglss.f90:57-80
          const int fi[4][8] = {{1, 1, -1, -1, -1, -1, 1, 1}, {1, -1, -1, 1, -1, 1, 1, -1},
                                {1, -1, 1, -1, 1, -1, 1, -1}, {1, -1, 1, -1, -1, 1, -1, 1}};
          double qv[3][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}};
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int li = lc[i] & EQD_LN_MASK;
            const double vx = sv[li], vy = sv[LS + li], vz = sv[2 * LS + li];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              qv[0][j] = qv[0][j] + vx * fi[j][i];
              qv[1][j] = qv[1][j] + vy * fi[j][i];
              qv[2][j] = qv[2][j] + vz * fi[j][i];
            }
          }
          const double coef = 0.25 * A.kapa_hg * LDG(A.rho + e) * LDG(A.vp + e) * pow(OP(RR_DET) * A.w, 2.0 / 3.0);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            double h0 = 0.0, h1 = 0.0, h2 = 0.0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              h0 = h0 - coef * qv[0][j] * fi[j][i];
              h1 = h1 - coef * qv[1][j] * fi[j][i];
              h2 = h2 - coef * qv[2][j] * fi[j][i];
            }
            f[i][0] = h0; f[i][1] = h1; f[i][2] = h2;
          }
        }
      }
#undef OP
    }
    // the stage buffer is consumed: request the next stage, or the next tile's first
    // stage and its nodal values; they land while the forces are assembled
    __syncthreads();
    const bool last = base + SE >= ne;
    if (!last) request(rc, base + SE, bc);
    else if (more) {
      request(rn, 0, bn);
      mbar_wait(barN, parityN);
      parityN ^= 1;
      gather_nodes(rn);   // sv / sd are dead once the last stage has been consumed
    }
    // ---- ordered assembly into the tile's shared force rows (one set per role)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int li = lc[i] & EQD_LN_MASK, col = lc[i] >> EQD_LN_BITS;
      for (int c = 0; c < NC; ++c) {
        if (act && col == c) { acc[li] += f[i][0]; acc[LS + li] += f[i][1]; acc[2 * LS + li] += f[i][2]; }
        role_sync(role, SE);
      }
    }
  }
  // ---- flush the tile's partial forces (both roles' rows)
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NPT; ++k) {
    const int i = tid + k * NT;
    if (i < LS) {
      if (i < TR_LN(rc)) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          if (SPLIT) {
            A.pf[(size_t)j * A.PFS + TR_NB(rc) + i] = sf[j * LS + i];
            A.pf[(size_t)(3 + j) * A.PFS + TR_NB(rc) + i] = sf[(3 + j) * LS + i];
          } else {
            A.pf[(size_t)j * A.PFS + TR_NB(rc) + i] = sf[j * LS + i] + sf[(3 + j) * LS + i];
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 6; ++j) sf[j * LS + i] = 0.0;
    }
  }
  if (!more) break;
  finish_nodes();
  __syncthreads();
  rc = rn; rn = rnn; tnext += gridDim.x;
  bc = bn; bn = bnn;
  }
#undef ISBOX
}

// ----------------------------------------------------------------------------
// PML element: calcPMLElemKU (assembleGlobalKU.f90:70-346) and hrglss.  The
// damping profile at the centroid (:130-213, constant in time) is precomputed by
// the host into damps[3][S].  Same streamed-tile scheme as the regular kernel,
// stages of EQD_STAGE_PML elements and FOUR threads per element:
//   warps 0-1  normal split stresses  s(1:9)   -> force dofs 1,5,9
//   warps 2-3  shear split stresses   s(10:15) -> force dofs 2,3,4,6,7,8
//   warps 4-5  regular stress part    s(16:21) + rdampk*rate -> dofs 10-12
//   warps 6-7  hourglass              (hrglss.f90)           -> dofs 10-12 (own rows, merged at the flush)
enum { PR_SHP = 0, PR_PHI = 24, PR_SS = 56, PR_LAM = 62, PR_MU = 63, PR_DET = 64, PR_DAMP = 65, PR_STRESS = 68, PR_ROWS = 89 };
__device__ __forceinline__ const double* pml_row_src(const ElemArgs& A, int r) {
  const size_t S = A.S;
  if (r < PR_PHI) return A.shp + (size_t)r * S;
  if (r < PR_SS) return A.phi + (size_t)(r - PR_PHI) * S;
  if (r < PR_LAM) return A.ss + (size_t)(r - PR_SS) * S;
  if (r == PR_LAM) return A.lam;
  if (r == PR_MU) return A.mu;
  if (r == PR_DET) return A.det;
  if (r < PR_STRESS) return A.damps + (size_t)(r - PR_DAMP) * S;
  return A.stress + (size_t)(r - PR_STRESS) * S;
}

// BOX: as in k_tile_reg -- tiles flagged in A.tileBox stream 33 of the 89 rows and
// rebuild eleshp = sign*a_d, phi = ha, ss = diag in registers (eqd_box.h).
constexpr int PR_BOX_ROWS = BOX_ROWS + (PR_ROWS - RR_ROWS);
template <bool BODY, int CHG, bool BOX>
__global__ void __launch_bounds__(4 * EQD_STAGE_PML, 2) k_tile_pml(ElemArgs A) {
  extern __shared__ __align__(128) unsigned char smraw[];
  constexpr int SE = EQD_STAGE_PML;
  constexpr int NT = 4 * SE;
  constexpr int LS = EQD_PML_LS;
  constexpr int NPT = (LS + NT - 1) / NT;
  constexpr int NROW = 15;          // force rows: 12 dofs + 3 hourglass rows
  const int tid = threadIdx.x;
  const int role = tid / SE;        // warp-uniform
  const int lane = tid - role * SE;
  const int tEnd = A.tile0 + A.ntiles;
  int tnext = A.tile0 + blockIdx.x;
  if (tnext >= tEnd) return;
  const size_t S = A.S, NS = A.NnS;
  const double dt = A.dt, rdampk = A.rdampk;
  uint64_t* bar = (uint64_t*)smraw;
  uint64_t* barN = bar + 1;
  double* ops = (double*)(smraw + 128);                         // [PR_ROWS][SE]
  uint16_t* slc = (uint16_t*)(ops + PR_ROWS * SE);              // [8][SE]
  int* tnS = (int*)(slc + 8 * SE);                              // [LS]
  double* sv = (double*)(tnS + LS);                             // [3][LS] velocity
  double* sl = sv + 3 * LS;                                     // [3][LS] d + rdampk*v
  double* sf = sl + 3 * LS;                                     // [NROW][LS]
  auto request = [&](const int4& rec, int base, bool bx) {
    const int cnt = min(SE, (TR_NE(rec) - base + 31) & ~31);
    if (tid == 0) mbar_expect_tx(bar, (uint32_t)(cnt * ((BOX && bx ? (int)PR_BOX_ROWS : (int)PR_ROWS) * 8 + 8 * 2)));
    __syncwarp();
    if (tid < PR_ROWS) { if (!(BOX && bx) || box_row(tid)) bulk_g2s(ops + tid * SE, pml_row_src(A, tid) + TR_E0(rec) + base, cnt * 8, bar); }
    else if (tid < PR_ROWS + 8) bulk_g2s(slc + (tid - PR_ROWS) * SE, A.lconn + (size_t)(tid - PR_ROWS) * S + TR_E0(rec) + base, cnt * 2, bar);
  };
  auto request_ids = [&](const int4& rec) {
    if (tid == 0) {
      mbar_expect_tx(barN, (uint32_t)(TR_LN(rec) * 4));
      bulk_g2s(tnS, A.tnode + TR_NB(rec), TR_LN(rec) * 4, barN);
    }
  };
  auto gather_nodes = [&](const int4& rec) {
#pragma unroll
    for (int k = 0; k < NPT; ++k) {
      const int i = tid + k * NT;
      if (i < LS) {
        const int nd = i < TR_LN(rec) ? tnS[i] : -1;
        if (nd >= 0) {
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            cp_async8(sv + j * LS + i, A.vel + j * NS + nd);
            if (CHG == 1) cp_async8(sl + j * LS + i, A.disp + j * NS + nd);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 3; ++j) { sv[j * LS + i] = 0.0; sl[j * LS + i] = 0.0; }
        }
      }
    }
    cp_async_commit();
  };
  auto finish_nodes = [&]() {
    cp_async_wait_all();
    if (CHG == 1) {
#pragma unroll
      for (int k = 0; k < NPT; ++k) {
        const int i = tid + k * NT;
        if (i < LS) {
#pragma unroll
          for (int j = 0; j < 3; ++j) sl[j * LS + i] = sl[j * LS + i] + rdampk * sv[j * LS + i];
        }
      }
    }
  };
  const int4 zero4 = make_int4(0, 0, 0, 0);
  int4 rc = __ldg(A.tileRec + tnext);
  bool bc = BOX && __ldg(A.tileBox + tnext) != 0;   // CTA-uniform
  tnext += gridDim.x;
  int4 rn = tnext < tEnd ? __ldg(A.tileRec + tnext) : zero4;
  bool bn = BOX && tnext < tEnd && __ldg(A.tileBox + tnext) != 0;
  if (tid == 0) { mbar_init(bar, 1); mbar_init(barN, 1); }
  __syncthreads();
  request(rc, 0, bc);
  request_ids(rc);
#pragma unroll
  for (int k = 0; k < NPT; ++k) {
    const int i = tid + k * NT;
    if (i < LS) {
#pragma unroll
      for (int j = 0; j < NROW; ++j) sf[j * LS + i] = 0.0;
    }
  }
  uint32_t parity = 0, parityN = 0;
  mbar_wait(barN, parityN);
  parityN ^= 1;
  gather_nodes(rc);
  finish_nodes();
  __syncthreads();
  while (true) {
    const bool more = tnext < tEnd;
    const int4 rnn = (tnext + (int)gridDim.x < tEnd) ? __ldg(A.tileRec + tnext + gridDim.x) : zero4;
    const bool bnn = BOX && (tnext + (int)gridDim.x < tEnd) && __ldg(A.tileBox + tnext + gridDim.x) != 0;
    if (more) request_ids(rn);
    const int ne = TR_NE(rc), NC = TR_NC(rc);
    for (int base = 0; base < ne; base += SE) {
      const int le = base + lane;
      const bool act = le < ne;
      const size_t e = (size_t)TR_E0(rc) + (act ? le : 0);
      unsigned lc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      // per role: up to 6 force rows per node, kept as coefficients so that few registers stay live
      // role 0: c[0..2] = detw*(sxx, syy, szz)      force(i) = 0 - c * shp
      // role 1: c[0..2] = detw*(sxy, sxz, syz)
      // role 2: s0[6], body                          role 3: hv[4][3]
      double c0 = 0, c1 = 0, c2 = 0, detw = 0;
      double s0[6] = {0, 0, 0, 0, 0, 0};
      double body[BODY ? 24 : 1];
      // roles 0-2: the element's shape derivatives; role 3: its 8x3 hourglass forces.  Held in
      // registers so that the stage buffer is free (and the next stage in flight) during assembly.
      double r24[24];
#pragma unroll
      for (int k = 0; k < 24; ++k) r24[k] = 0.0;
      mbar_wait(bar, parity);
      parity ^= 1;
#define OP(r) ops[(r) * SE + lane]
      if (act) {
#pragma unroll
        for (int i = 0; i < 8; ++i) lc[i] = slc[i * SE + lane];
        const double lam = OP(PR_LAM), mu = OP(PR_MU);
        const double l2m = lam + 2.0 * mu;
        const double rdt = 1 / dt;
        detw = OP(PR_DET) * A.w;
        // :277-311  s <- (coef*D + (1/dt - d/2) s) / (1/dt + d/2)
        // the three denominators are formed once per element and applied as reciprocals
        // (15 divisions -> 3; differs from the reference's `/` by at most one rounding)
        const double dh[3] = {OP(PR_DAMP) / 2, OP(PR_DAMP + 1) / 2, OP(PR_DAMP + 2) / 2};
        const double rden[3] = {1.0 / (rdt + dh[0]), 1.0 / (rdt + dh[1]), 1.0 / (rdt + dh[2])};
#define PML_UPD(k, cf, D, a, out)                                         \
  {                                                                       \
    const double x = ((cf) * (D) + (rdt - dh[a]) * OP(PR_STRESS + (k))) * rden[a]; \
    out = x;                                                              \
    A.stress[(size_t)(k) * S + e] = x;                                    \
  }
        if (role < 3) {
          if (BOX && bc) {
            const double ax = OP(BOX_AX), ay = OP(BOX_AY), az = OP(BOX_AZ);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              r24[3 * i] = box_px(i) ? ax : -ax;
              r24[3 * i + 1] = box_py(i) ? ay : -ay;
              r24[3 * i + 2] = box_pz(i) ? az : -az;
            }
          } else {
#pragma unroll
          for (int k = 0; k < 24; ++k) r24[k] = OP(PR_SHP + k);
          }
        }
        if (role == 0) {
          double g00 = 0, g11 = 0, g22 = 0;  // assembleGlobalKU.f90:248-275
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int li = lc[i] & EQD_LN_MASK;
            g00 = g00 + r24[3 * i] * sv[li];
            g11 = g11 + r24[3 * i + 1] * sv[LS + li];
            g22 = g22 + r24[3 * i + 2] * sv[2 * LS + li];
          }
          double s[9];
          PML_UPD(0, l2m, g00, 0, s[0]) PML_UPD(1, lam, g11, 1, s[1]) PML_UPD(2, lam, g22, 2, s[2])
          PML_UPD(3, lam, g00, 0, s[3]) PML_UPD(4, l2m, g11, 1, s[4]) PML_UPD(5, lam, g22, 2, s[5])
          PML_UPD(6, lam, g00, 0, s[6]) PML_UPD(7, lam, g11, 1, s[7]) PML_UPD(8, l2m, g22, 2, s[8])
          c0 = s[0] + s[1] + s[2]; c1 = s[3] + s[4] + s[5]; c2 = s[6] + s[7] + s[8];  // sxx, syy, szz
        } else if (role == 1) {
          double g01 = 0, g10 = 0, g02 = 0, g20 = 0, g12 = 0, g21 = 0;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int li = lc[i] & EQD_LN_MASK;
            const double vx = sv[li], vy = sv[LS + li], vz = sv[2 * LS + li];
            const double s1 = r24[3 * i], s2 = r24[3 * i + 1], s3 = r24[3 * i + 2];
            g01 = g01 + s1 * vy; g10 = g10 + s2 * vx;
            g02 = g02 + s1 * vz; g20 = g20 + s3 * vx;
            g12 = g12 + s2 * vz; g21 = g21 + s3 * vy;
          }
          double s[6];
          PML_UPD(9, mu, g01, 0, s[0]) PML_UPD(10, mu, g10, 1, s[1])
          PML_UPD(11, mu, g02, 0, s[2]) PML_UPD(12, mu, g20, 2, s[3])
          PML_UPD(13, mu, g12, 1, s[4]) PML_UPD(14, mu, g21, 2, s[5])
          c0 = s[0] + s[1]; c1 = s[2] + s[3]; c2 = s[4] + s[5];  // sxy, sxz, syz
        } else if (role == 2) {
          double sr[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int li = lc[i] & EQD_LN_MASK;
            const double vx = sv[li], vy = sv[LS + li], vz = sv[2 * LS + li];
            const double s1 = r24[3 * i], s2 = r24[3 * i + 1], s3 = r24[3 * i + 2];
            // assembleGlobalKU.f90:215-246
            sr[0] = sr[0] + s1 * vx;
            sr[1] = sr[1] + s2 * vy;
            sr[2] = sr[2] + s3 * vz;
            sr[3] = sr[3] + s3 * vy + s2 * vz;
            sr[4] = sr[4] + s3 * vx + s1 * vz;
            sr[5] = sr[5] + s2 * vx + s1 * vy;
            if (BODY) {
              const double em = LDG(A.emass + i * S + e);
              body[3 * i] = 0.0 - (A.rdampm * vx) * em;
              body[3 * i + 1] = 0.0 - (A.rdampm * vy) * em;
              body[3 * i + 2] = 0.0 - (A.rdampm * vz + A.bodyz) * em;
            }
          }
          double rate[6];
          rate[0] = 0.0 + l2m * sr[0] + lam * sr[1] + lam * sr[2];
          rate[1] = 0.0 + lam * sr[0] + l2m * sr[1] + lam * sr[2];
          rate[2] = 0.0 + lam * sr[0] + lam * sr[1] + l2m * sr[2];
          rate[3] = mu * sr[3];
          rate[4] = mu * sr[4];
          rate[5] = mu * sr[5];
#pragma unroll
          for (int i = 0; i < 6; ++i) s0[i] = OP(PR_STRESS + 15 + i) + rdampk * rate[i];  // :320-325 (read-only slots)
        } else {
          if (CHG == 1) {
            if (BOX && bc) {
              double l[8][3], fh[8][3];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int li = lc[i] & EQD_LN_MASK;
                l[i][0] = sl[li]; l[i][1] = sl[LS + li]; l[i][2] = sl[2 * LS + li];
              }
              box_hourglass(l, OP(BOX_SS0), OP(BOX_SS3), OP(BOX_SS5), fh);
#pragma unroll
              for (int i = 0; i < 8; ++i) { r24[3 * i] = fh[i][0]; r24[3 * i + 1] = fh[i][1]; r24[3 * i + 2] = fh[i][2]; }
            } else {
            double phid[4][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int li = lc[i] & EQD_LN_MASK;
              const double lx = sl[li], ly = sl[LS + li], lz = sl[2 * LS + li];
#pragma unroll
              for (int m = 0; m < 4; ++m) {
                const double p = OP(PR_PHI + 8 * m + i);
                phid[m][0] = phid[m][0] + p * lx;
                phid[m][1] = phid[m][1] + p * ly;
                phid[m][2] = phid[m][2] + p * lz;
              }
            }
            double ssv[6], hv[4][3];
#pragma unroll
            for (int i = 0; i < 6; ++i) ssv[i] = OP(PR_SS + i);
#pragma unroll
            for (int m = 0; m < 4; ++m) {
              hv[m][0] = ssv[0] * phid[m][0] + ssv[1] * phid[m][1] + ssv[2] * phid[m][2];
              hv[m][1] = ssv[1] * phid[m][0] + ssv[3] * phid[m][1] + ssv[4] * phid[m][2];
              hv[m][2] = ssv[2] * phid[m][0] + ssv[4] * phid[m][1] + ssv[5] * phid[m][2];
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              double h0 = 0.0, h1 = 0.0, h2 = 0.0;
#pragma unroll
              for (int m = 0; m < 4; ++m) {   // hrglss.f90:41-54
                const double p = OP(PR_PHI + 8 * m + i);
                h0 = h0 - p * hv[m][0];
                h1 = h1 - p * hv[m][1];
                h2 = h2 - p * hv[m][2];
              }
              r24[3 * i] = h0; r24[3 * i + 1] = h1; r24[3 * i + 2] = h2;
            }
            }
          } else {
            const int fi[4][8] = {{1, 1, -1, -1, -1, -1, 1, 1}, {1, -1, -1, 1, -1, 1, 1, -1},
                                  {1, -1, 1, -1, 1, -1, 1, -1}, {1, -1, 1, -1, -1, 1, -1, 1}};
            double qv[3][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}};
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int li = lc[i] & EQD_LN_MASK;
              const double vx = sv[li], vy = sv[LS + li], vz = sv[2 * LS + li];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                qv[0][j] = qv[0][j] + vx * fi[j][i];
                qv[1][j] = qv[1][j] + vy * fi[j][i];
                qv[2][j] = qv[2][j] + vz * fi[j][i];
              }
            }
            const double qcoef = 0.25 * A.kapa_hg * LDG(A.rho + e) * LDG(A.vp + e) * pow(OP(PR_DET) * A.w, 2.0 / 3.0);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              double h0 = 0.0, h1 = 0.0, h2 = 0.0;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                h0 = h0 - qcoef * qv[0][j] * fi[j][i];
                h1 = h1 - qcoef * qv[1][j] * fi[j][i];
                h2 = h2 - qcoef * qv[2][j] * fi[j][i];
              }
              r24[3 * i] = h0; r24[3 * i + 1] = h1; r24[3 * i + 2] = h2;
            }
          }
        }
#undef PML_UPD
      }
#undef OP
      // the stage buffer and the nodal values are consumed: request the next stage, or the
      // next tile's first stage and its nodes; they land while the forces are assembled
      __syncthreads();
      const bool last = base + SE >= ne;
      if (!last) request(rc, base + SE, bc);
      else if (more) {
        request(rn, 0, bn);
        mbar_wait(barN, parityN);
        parityN ^= 1;
        gather_nodes(rn);
      }
      // ---- ordered assembly (assembleGlobalKU.f90:328-344): 8 local-node phases
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int li = lc[i] & EQD_LN_MASK, col = lc[i] >> EQD_LN_BITS;
        const double s1 = r24[3 * i], s2 = r24[3 * i + 1], s3 = r24[3 * i + 2];
        for (int c = 0; c < NC; ++c) {
          if (act && col == c) {
            if (role == 0) {         // dofs 1,5,9
              sf[0 * LS + li] += 0.0 - detw * s1 * c0;
              sf[4 * LS + li] += 0.0 - detw * s2 * c1;
              sf[8 * LS + li] += 0.0 - detw * s3 * c2;
            } else if (role == 1) {  // dofs 2,3,4,6,7,8
              sf[1 * LS + li] += 0.0 - detw * s2 * c0;
              sf[2 * LS + li] += 0.0 - detw * s3 * c1;
              sf[3 * LS + li] += 0.0 - detw * s1 * c0;
              sf[5 * LS + li] += 0.0 - detw * s3 * c2;
              sf[6 * LS + li] += 0.0 - detw * s1 * c1;
              sf[7 * LS + li] += 0.0 - detw * s2 * c2;
            } else if (role == 2) {  // dofs 10-12
              double b0 = 0.0, b1 = 0.0, b2 = 0.0;
              if (BODY) { b0 = body[3 * i]; b1 = body[3 * i + 1]; b2 = body[3 * i + 2]; }
              sf[9 * LS + li] += b0 - detw * (s1 * s0[0] + s3 * s0[4] + s2 * s0[5]);
              sf[10 * LS + li] += b1 - detw * (s2 * s0[1] + s3 * s0[3] + s1 * s0[5]);
              sf[11 * LS + li] += b2 - detw * (s3 * s0[2] + s2 * s0[3] + s1 * s0[4]);
            } else {                 // hourglass, own rows 13-15
              sf[12 * LS + li] += s1; sf[13 * LS + li] += s2; sf[14 * LS + li] += s3;
            }
          }
          role_sync(role, SE);   // the four roles own disjoint force rows
        }
      }
    }
    __syncthreads();
    // ---- flush
#pragma unroll
    for (int k = 0; k < NPT; ++k) {
      const int i = tid + k * NT;
      if (i < LS) {
        if (i < TR_LN(rc)) {
#pragma unroll
          for (int j = 0; j < 9; ++j) A.pf[(size_t)j * A.PFS + TR_NB(rc) + i] = sf[j * LS + i];
#pragma unroll
          for (int j = 9; j < 12; ++j) A.pf[(size_t)j * A.PFS + TR_NB(rc) + i] = sf[j * LS + i] + sf[(j + 3) * LS + i];
        }
#pragma unroll
        for (int j = 0; j < NROW; ++j) sf[j * LS + i] = 0.0;
      }
    }
    if (!more) break;
    finish_nodes();
    __syncthreads();
    rc = rn; rn = rnn; tnext += gridDim.x;
    bc = bn; bn = bnn;
  }
}

// ----------------------------------------------------------------------------
// storeOffFaultStData, driver.f90:157-180
__global__ void k_store_offfault(const int* __restrict__ idhist, int n, double* __restrict__ out,
                                 const double* __restrict__ vel, const double* __restrict__ disp, int NnS,
                                 const StepState* st) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // 0..n   (0 = time)
  if (i > n) return;
  double* row = out + (size_t)(n + 1) * (st->nt - 1);
  if (i == 0) { row[0] = st->timeElapsed; return; }
  const int node = idhist[3 * (i - 1)] - 1, k = idhist[3 * (i - 1) + 1] - 1, q = idhist[3 * (i - 1) + 2];
  if (q == 1) row[i] = disp[(size_t)k * NnS + node];
  else if (q == 2) row[i] = vel[(size_t)k * NnS + node];
}

// output_gm / output_src_evol samples (library_output.f90:267-279,297-312)
__global__ void k_sample_gm(const int* __restrict__ surf, int nSurf, const double* __restrict__ vel, int NnS,
                            double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nSurf) return;
  const int n = surf[i] - 1;
  out[3 * (size_t)i] = vel[n]; out[3 * (size_t)i + 1] = vel[(size_t)NnS + n]; out[3 * (size_t)i + 2] = vel[2 * (size_t)NnS + n];
}
__global__ void k_sample_src(const double* __restrict__ fric, int PS, int n, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = fric[(size_t)46 * PS + i];  // fric(47,i,1)
}

// ----------------------------------------------------------------------------
// halo pack / unpack-add (processNodalQuantArr, assembleGlobalMass.f90:145-236)
__global__ void k_pack(const double* __restrict__ src, const uint32_t* __restrict__ idx, int n, double* __restrict__ buf) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) buf[i] = src[idx[i]];
}
__global__ void k_unpack_add(double* __restrict__ dst, const uint32_t* __restrict__ idx, int n, const double* __restrict__ buf) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[idx[i]] = dst[idx[i]] + buf[i];
}

// ----------------------------------------------------------------------------
// The same exchange without a communication library in the step loop: the sub-domains of one box see each
// other's memory over NVLink (CUDA IPC, set up once in eqd_api.cu).  k_halo_send packs the face dofs straight
// into the NEIGHBOUR's receive buffer (remote stores) and, when the last block has done so, publishes the step
// number in the neighbour's flag word; k_halo_recv waits for its own flag word to show this step and adds the
// received values (processNodalQuantArr's "+", assembleGlobalMass.f90:190-236).  blockIdx.y = side of the axis.
// Two receive buffers per face, used alternately: a sender can be at most one exchange ahead of its receiver
// (it needs the receiver's data of exchange n+1 before it can send exchange n+2).
__global__ void __launch_bounds__(256) k_halo_send(HaloAxisArgs A) {
  const int side = blockIdx.y;
  const int n = A.n[side];
  if (n <= 0) return;
  const int nblk = (n + 255) / 256;
  if ((int)blockIdx.x >= nblk) return;
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i < n) A.remoteRecv[side][i] = A.force[A.idx[side][i]];
  __threadfence_system();          // this thread's remote stores are visible system-wide ...
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned done = atomicAdd(A.counter[side], 1u) + 1u;
    if (done == (unsigned)nblk) {  // ... so the block that arrives last may announce the whole message
      *A.counter[side] = 0u;
      __threadfence_system();
      *(volatile unsigned*)A.remoteFlag[side] = A.seq;
    }
  }
}
__global__ void __launch_bounds__(256) k_halo_recv(HaloAxisArgs A) {
  const int side = blockIdx.y;
  const int n = A.n[side];
  if (n <= 0) return;
  if ((int)blockIdx.x >= (n + 255) / 256) return;
  if (threadIdx.x == 0) {
    const volatile unsigned* f = A.localFlag[side];
    while (*f != A.seq) __nanosleep(100);
    __threadfence_system();
  }
  __syncthreads();
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i < n) {
    const uint32_t k = A.idx[side][i];
    A.force[k] = A.force[k] + __ldcg(A.localRecv[side] + i);   // written by the peer: never through this SM's L1
  }
}

// ----------------------------------------------------------------------------
// thermop.f90:1-40: one thread per pair, O(nt) history convolution
__global__ void __launch_bounds__(128) k_thermop(FaultArgs A) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= A.nPairs) return;
  const size_t PS = A.PS;
#define FR(k) A.fric[(size_t)((k) - 1) * PS + p]
  const double dt = A.dt, pi = 4 * atan(1.0), h = A.fric_tp_h;
  const int nt = A.st->nt;
  const double gama = FR(19) / FR(18), omega = FR(16), kapa = FR(17);
  double tmp = 0.0, tmp2 = 0.0;
  for (int j = 1; j <= nt - 1; ++j) {
    const double V = A.tphist[((size_t)(j - 1) * 2) * PS + p], tau = A.tphist[((size_t)(j - 1) * 2 + 1) * PS + p];
    double ker = -kapa / (omega - kapa) / sqrt(4.0 * kapa * (nt - j) * dt + 2.0 * (h * h));
    ker = ker + omega / (omega - kapa) / sqrt(4.0 * omega * (nt - j) * dt + 2.0 * (h * h));
    tmp = tmp + fabs(tau) * V * ker * dt;
    const double ker2 = 1.0 / sqrt(4.0 * kapa * (nt - j) * dt + 2.0 * (h * h));
    tmp2 = tmp2 + fabs(tau) * V * ker2 * dt;
  }
  FR(51) = tmp * gama / sqrt(pi);
  FR(52) = tmp2 / FR(18) / sqrt(pi) + FR(41);
#undef FR
}

// ----------------------------------------------------------------------------
// fric.f90
__device__ __forceinline__ double d_slip_weak(double slip, double fs, double fd, double d0) {  // fric.f90:3-19
  double xmu = 0.0;
  if (fabs(slip) < (double)1.0e-10f) xmu = fs;
  else if (slip < d0) xmu = fs - (fs - fd) * slip / d0;
  if (slip >= d0) xmu = fd;
  return xmu;
}
__device__ __forceinline__ double d_time_weak(double trupt, double fs, double fd, double t0) {  // fric.f90:21-37
  if (trupt <= 0.0) return fs;
  else if (trupt < t0) return fs - (fs - fd) * trupt / t0;
  return fd;
}
struct RsfPar { double A, B, L, f0, V0, fw, Vw; };
// fric.f90:39-95 split into its two halves: the friction coefficient and its slip-rate derivative
// (all the Newton solve needs, with the state-dependent factor tmpc hoisted: the state is reset to
// state0 at the top of every iteration, faulting.f90:471), and the state evolution (only the LAST
// iteration's value survives, faulting.f90:511, so it is evaluated once after the loop).  Same
// expressions on the same inputs: bit-identical to evaluating both halves in every iteration.
__device__ __forceinline__ double d_rsf_tmpc_ageing(double theta, const RsfPar& r) {
  return 1.0 / (2.0 * r.V0) * exp((r.f0 + r.B * log(r.V0 * theta / r.L)) / r.A);
}
__device__ __forceinline__ double d_rsf_tmpc_slip(double psi, const RsfPar& r) { return 1.0 / (2.0 * r.V0) * exp(psi / r.A); }
__device__ __forceinline__ void d_rsf_mu(double V2, double tmpc, const RsfPar& r, double& xmu, double& dxmudv) {
  const double tmp = (V2 + 1.e-30) * tmpc;
  xmu = r.A * log(tmp + sqrt(tmp * tmp + 1.0));
  dxmudv = r.A * tmpc / sqrt(1.0 + tmp * tmp);
}
__device__ __forceinline__ double d_rsf_state_ageing(double V2, double theta, const RsfPar& r, double dt) {
  return r.L / V2 + (theta - r.L / V2) * exp(-V2 * dt / r.L);
}
__device__ __forceinline__ double d_rsf_state_slip(double V2, double psi, const RsfPar& r, double dt) {
  const double fLV = r.f0 - (r.B - r.A) * log(V2 / r.V0);
  const double q = V2 / r.Vw, q2 = q * q, q4 = q2 * q2;
  const double fss = r.fw + (fLV - r.fw) / pow(1.0 + q4 * q4, 0.125);
  const double fssa = fss / r.A;
  const double psiss = r.A * log(2.0 * r.V0 / V2 * (exp(fssa) - exp(-fssa)) / 2.0);
  return psiss + (psi - psiss) * exp(-V2 * dt / r.L);
}
__device__ __forceinline__ void d_rsf_ageing(double V2, double& theta, const RsfPar& r, double& xmu, double& dxmudv, double dt) {  // fric.f90:39-61
  d_rsf_mu(V2, d_rsf_tmpc_ageing(theta, r), r, xmu, dxmudv);
  theta = d_rsf_state_ageing(V2, theta, r, dt);
}
__device__ __forceinline__ void d_rsf_slip(double V2, double& psi, const RsfPar& r, double& xmu, double& dxmudv, double dt) {  // fric.f90:63-95
  d_rsf_mu(V2, d_rsf_tmpc_slip(psi, r), r, xmu, dxmudv);
  psi = d_rsf_state_slip(V2, psi, r, dt);
}

// faulting.f90:3-541, one thread per split-node pair
__global__ void __launch_bounds__(64, 8) k_fault(FaultArgs A) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= A.nPairs) return;
  const size_t PS = A.PS, NS = A.NnS;
#define FR(k) A.fric[(size_t)((k) - 1) * PS + p]
  const double dt = A.dt;
  const int nt = A.st->nt;
  const double timeElapsed = A.st->timeElapsed;
  const int nS = A.nodeS[p], nM = A.nodeM[p];
  const int ift = A.ift[p];
  double un[3], us[3], ud[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) { un[k] = A.un[k * PS + p]; us[k] = A.us[k * PS + p]; ud[k] = A.ud[k * PS + p]; }
  const double arn = A.arn[p];
  const double Cel = (double)A.C_elastic;
  // ---- getNsdSlipSliprateTraction, faulting.f90:54-134
  const double initT[3] = {FR(7), FR(8) + 0.0, FR(49)};
  const double massSlave = A.massS[p], massMaster = A.massM[p];
  const double totalMass = (massSlave + massMaster) * arn;
  double fS[3], fM[3], vS[3], vM[3], dS[3], dM[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    fS[k] = A.force[k * NS + nS]; fM[k] = A.force[k * NS + nM];
    vS[k] = A.vel[k * NS + nS]; vM[k] = A.vel[k * NS + nM];
    dS[k] = A.disp[k * NS + nS]; dM[k] = A.disp[k * NS + nM];
  }
#define ROT(x, u) ((x)[0] * (u)[0] + (x)[1] * (u)[1] + (x)[2] * (u)[2])
  const double fSn[3] = {ROT(fS, un), ROT(fS, us), ROT(fS, ud)}, fMn[3] = {ROT(fM, un), ROT(fM, us), ROT(fM, ud)};
  const double vSn[3] = {ROT(vS, un), ROT(vS, us), ROT(vS, ud)}, vMn[3] = {ROT(vM, un), ROT(vM, us), ROT(vM, ud)};
  const double dSn[3] = {ROT(dS, un), ROT(dS, us), ROT(dS, ud)}, dMn[3] = {ROT(dM, un), ROT(dM, us), ROT(dM, ud)};
#undef ROT
  double slip[4], rate[4], T[4];
  for (int j = 0; j < 3; ++j) slip[j] = dMn[j] - dSn[j];
  slip[3] = sqrt(slip[0] * slip[0] + slip[1] * slip[1] + slip[2] * slip[2]);
  for (int j = 0; j < 3; ++j) rate[j] = vMn[j] - vSn[j];
  rate[3] = sqrt(rate[0] * rate[0] + rate[1] * rate[1] + rate[2] * rate[2]);
  FR(71) = slip[1]; FR(72) = slip[2]; FR(73) = slip[0];
  FR(74) = rate[1]; FR(75) = rate[2];
  double f76 = FR(76);
  if (rate[3] > f76) { f76 = rate[3]; FR(76) = f76; }
  const double f77 = FR(77) + rate[3] * dt;
  FR(77) = f77;
  T[0] = (massSlave * massMaster * ((vMn[0] - vSn[0]) + (dMn[0] - dSn[0]) / dt) / dt + massSlave * fMn[0] -
          massMaster * fSn[0]) / totalMass + initT[0] * Cel;
  T[1] = (massSlave * massMaster * (vMn[1] - vSn[1]) / dt + massSlave * fMn[1] - massMaster * fSn[1]) / totalMass +
         initT[1] * Cel;
  T[2] = (massSlave * massMaster * (vMn[2] - vSn[2]) / dt + massSlave * fMn[2] - massMaster * fSn[2]) / totalMass +
         initT[2] * Cel;
  const double xs0 = A.xs[p], xs1 = A.xs[PS + p], xs2 = A.xs[2 * PS + p];
  const double radius = sqrt((xs0 - A.xsource) * (xs0 - A.xsource) + (xs1 - A.ysource) * (xs1 - A.ysource) +
                             (xs2 - A.zsource) * (xs2 - A.zsource));
  double f20 = 0.0, f23 = 0.0;
  if (A.friclaw >= 3) { f20 = FR(20); f23 = FR(23); }
  if (A.friclaw >= 3 && A.C_nuclea == 1 && ift == A.nucfault) {
    // rsfNucleation, faulting.f90:367-414
    double dtau2 = 0.0, Fq = 0.0, G = 1.0;
    if (radius < A.nucR) Fq = exp(radius * radius / (radius * radius - A.nucR * A.nucR));
    if (timeElapsed <= A.nucT)
      G = exp((timeElapsed - A.nucT) * (timeElapsed - A.nucT) / (timeElapsed * (timeElapsed - 2.0 * A.nucT)));
    if (A.TPV == 105 || A.TPV == 104) dtau2 = A.nucdtau0 * Fq * G;
    else if (A.TPV == 2802) {
      if (nt == 1) {
        FR(81) = A.nucdtau0;
        const double ttao = sqrt(T[1] * T[1] + T[2] * T[2]);
        const double back = sqrt((rate[1] + FR(26)) * (rate[1] + FR(26)) + (rate[2] + FR(27)) * (rate[2] + FR(27)));
        f20 = FR(9) * log(2.0 * FR(12) / back * sinh(ttao / fabs(T[0]) / FR(9)));
        f23 = fabs(T[0]);
      }
      dtau2 = FR(81) * Fq * G;
    }
    T[1] = T[1] + dtau2;
  }
  T[3] = sqrt(T[1] * T[1] + T[2] * T[2]);

  if (A.friclaw <= 2) {
    // ---- solveSWTW, faulting.f90:136-188
    double mu = 0.0;
    const double fs = FR(1), fd = FR(2);
    if (A.friclaw == 1) mu = d_slip_weak(f77, fs, fd, FR(3));
    else mu = d_time_weak(timeElapsed - A.fnft[p], fs, fd, FR(5));
    if (A.C_nuclea == 1 && ift == A.nucfault) {
      // swtwNucleation, faulting.f90:416-443
      double tr = 1.0e9;
      if (radius <= A.nucR) {
        if (A.TPV == 201 || A.TPV == 36 || A.TPV == 37)
          tr = (radius + 0.081 * A.nucR * (1.0 / (1.0 - (radius / A.nucR) * (radius / A.nucR)) - 1.0)) / (0.7 * 3464.0);
        if (A.TPV == 202) tr = radius / A.nucRuptVel;
      }
      const double t0 = FR(5);
      double tc = 1.0;
      if (timeElapsed < tr) tc = 0.0;
      else if ((timeElapsed < (tr + t0)) && (timeElapsed >= tr)) tc = (timeElapsed - tr) / t0;
      mu = fmin(fs + (fd - fs) * tc, mu);
    }
    const double pore = FR(6);
    double effN;
    if ((T[0] + pore) > 0) effN = 0.0;
    else effN = T[0] + pore;
    const double trial = FR(4) - mu * effN;
    if (T[3] > trial) {
      T[1] = T[1] * trial / T[3];
      T[2] = T[2] * trial / T[3];
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const double xt = (T[0] * un[j] + T[1] * us[j] + T[2] * ud[j]) * arn;
      const double x0 = (initT[0] * un[j] + initT[1] * us[j] + initT[2] * ud[j]) * arn;
      A.force[j * NS + nS] = fS[j] + xt - x0 * Cel;
      A.force[j * NS + nM] = fM[j] - xt + x0 * Cel;
    }
    FR(78) = T[0]; FR(79) = T[1]; FR(80) = T[2];
  } else {
    // ---- solveRSF, faulting.f90:190-327
    if (A.friclaw == 5) T[0] = T[0] + FR(51);
    else T[0] = T[0] + FR(6);
    if (A.insertFaultType > 0 && A.C_elastic == 1) {
      const double max_norm = -40.0e6, min_norm = -10.0e6;
      if (T[0] >= min_norm) T[0] = min_norm;
      else if (T[0] <= max_norm) T[0] = max_norm;
    }
    if (T[0] > 0.0) T[0] = 0.0;
    const double bg[3] = {FR(25), FR(26), FR(27)};
    for (int j = 0; j < 3; ++j) {
      slip[j] = slip[j] + bg[j] * timeElapsed;
      rate[j] = rate[j] + bg[j];
    }
    slip[3] = sqrt(slip[1] * slip[1] + slip[2] * slip[2]);
    rate[3] = sqrt(rate[1] * rate[1] + rate[2] * rate[2]);
    double v_trial = rate[3];
    RsfPar r;
    r.A = FR(9); r.B = FR(10); r.L = FR(11); r.V0 = FR(12); r.f0 = FR(13); r.fw = FR(14); r.Vw = FR(15);
    const double coh = FR(4);
    double theta_pc_tmp = f23, theta_pc_dot;
    // rate_state_normal_stress, faulting.f90:35-52
    theta_pc_dot = -v_trial / r.L * (f23 - fabs(T[0]));
    f23 = f23 + theta_pc_dot * dt;
    FR(24) = theta_pc_dot;
    double statetmp = f20;
    double xmu = 0, dxmudv = 0;
    if (A.friclaw == 3) d_rsf_ageing(v_trial, f20, r, xmu, dxmudv, dt);
    else d_rsf_slip(v_trial, f20, r, xmu, dxmudv, dt);
    double taoc_old;
    if (A.friclaw == 5) taoc_old = coh - xmu * T[0];
    else taoc_old = xmu * theta_pc_tmp;
    const double mr = massMaster * massSlave / (massMaster + massSlave);
    const double T_coeff = arn * dt / mr;
    double trialT[4] = {0, 0, 0, 0};
    for (int j = 1; j <= 2; ++j) trialT[j] = T[j] - taoc_old * 0.5 * (rate[j] / rate[3]) + bg[j] / T_coeff;
    trialT[3] = sqrt(trialT[1] * trialT[1] + trialT[2] * trialT[2]);
    // NewtonRaphson, faulting.f90:459-516
    double taoc_new = 0.0;
    {
      const double state0 = statetmp, thetaPc0 = theta_pc_tmp;
      double stateTmp = state0;
      double thetaPcTmp = 0.0;  // uninitialised local in the reference when friclaw==5 (never read by the solve)
      const double tmpc0 = A.friclaw == 3 ? d_rsf_tmpc_ageing(state0, r) : d_rsf_tmpc_slip(state0, r);
      double v_state = v_trial;   // slip rate the surviving state evolution is evaluated at
      for (int iv = 1; iv <= 20; ++iv) {
        v_state = v_trial;
        d_rsf_mu(v_trial, tmpc0, r, xmu, dxmudv);
        double rsfeq, drsfeqdv;
        if (A.friclaw < 5) {
          thetaPcTmp = thetaPc0;
          const double dot = -v_trial / r.L * (thetaPcTmp - fabs(T[0]));
          thetaPcTmp = thetaPcTmp + dot * dt;
          taoc_new = xmu * thetaPcTmp;
          rsfeq = v_trial + T_coeff * (taoc_new * 0.5 - trialT[3]);
          drsfeqdv = 1.0 + T_coeff * (dxmudv * thetaPcTmp) * 0.5;
        } else {
          taoc_new = coh - xmu * fmin(T[0], 0.0);
          rsfeq = v_trial + T_coeff * (taoc_new * 0.5 - trialT[3]);
          drsfeqdv = 1.0 + T_coeff * (-dxmudv * fmin(T[0], 0.0)) * 0.5;
        }
        if (fabs(rsfeq / drsfeqdv) < 1.e-14 * fabs(v_trial) && fabs(rsfeq) < 1.e-6 * fabs(v_trial)) break;
        const double newSliprate = v_trial - rsfeq / drsfeqdv;
        if (newSliprate <= 0.0) v_trial = v_trial / 2.0;
        else v_trial = newSliprate;
      }
      stateTmp = A.friclaw == 3 ? d_rsf_state_ageing(v_state, state0, r, dt) : d_rsf_state_slip(v_state, state0, r, dt);
      if (A.TPV == 105 && v_trial < FR(46)) v_trial = FR(46);
      statetmp = stateTmp;
      theta_pc_tmp = thetaPcTmp;
    }
    f20 = statetmp;
    f23 = theta_pc_tmp;
    for (int j = 1; j <= 2; ++j) T[j] = taoc_old * 0.5 * (rate[j] / rate[3]) + taoc_new * 0.5 * (trialT[j] / trialT[3]);
    FR(78) = T[0]; FR(79) = T[1]; FR(80) = T[2];
    FR(47) = v_trial;
    const double tmag = sqrt(T[1] * T[1] + T[2] * T[2]);
    FR(48) = tmag;
    if (A.tphist) {
      A.tphist[((size_t)(nt - 1) * 2) * PS + p] = v_trial;
      A.tphist[((size_t)(nt - 1) * 2 + 1) * PS + p] = tmag;
    }
    double acc[3];
    acc[0] = -rate[0] / dt - slip[0] / dt / dt;
    acc[1] = (v_trial * (trialT[1] / trialT[3]) - rate[1]) / dt;
    acc[2] = (v_trial * (trialT[2] / trialT[3]) - rate[2]) / dt;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const double xa = acc[0] * un[j] + acc[1] * us[j] + acc[2] * ud[j];
      const double xr = fS[j] + fM[j];
      const double aS = (-xa + xr / massMaster), aM = (xa + xr / massSlave);
      A.force[j * NS + nS] = aS * mr;
      A.force[j * NS + nM] = aM * mr;
      FR(31 + j) = vM[j] + aM * dt;
      FR(34 + j) = vS[j] + aS * dt;
    }
  }
  if (A.friclaw >= 3) { FR(20) = f20; FR(23) = f23; }
  // showSourceDynamics, faulting.f90:343-365 (recorded; the host prints)
  if (fabs(xs0 - A.xsource) < A.tol && fabs(xs2 - A.zsource) < A.tol && A.hypoLog) {
    double* h = A.hypoLog + 13 * (size_t)(nt - 1);
    h[0] = timeElapsed; h[1] = FR(78); h[2] = FR(79); h[3] = FR(80); h[4] = FR(23); h[5] = FR(73); h[6] = FR(71);
    h[7] = FR(72); h[8] = FR(74); h[9] = FR(75); h[10] = FR(76); h[11] = FR(77); h[12] = FR(20);
  }
  // storeOnFaultStationQuantSCEC, faulting.f90:518-541
  const int stn = A.pairStation[p];
  if (stn >= 0) {
    double* q = A.onHist + 12 * ((size_t)(nt - 1) + (size_t)A.nstep * stn);
    q[0] = timeElapsed; q[1] = rate[1]; q[2] = rate[2]; q[3] = FR(20);
    q[4] = slip[1]; q[5] = slip[2]; q[6] = slip[0];
    q[7] = T[1]; q[8] = T[2]; q[9] = T[0];
    q[10] = FR(51) + FR(42); q[11] = FR(52);
  }
  // storeRuptureTime, faulting.f90:329-341
  if (A.fnft[p] > 5000.0)
    if (rate[3] >= A.slipRateThres) A.fnft[p] = timeElapsed;
#undef FR
}

// ----------------------------------------------------------------------------
// layout conversion helpers (upload / fetch): AoS(k fastest) <-> SoA(element fastest)
__global__ void k_aos_to_soa(const double* __restrict__ src, int K, int n, const int* __restrict__ dstIdx, int cls,
                             double* __restrict__ dst, int S, int k0, int nk) {
  // src(K, n) column-major: src[k + K*e]; element e goes to class slot dstIdx[e] if its class matches
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const int code = dstIdx[e];
  if ((code & 3) != cls) return;
  const size_t d = (size_t)(code >> 2);
  for (int k = 0; k < nk; ++k) dst[(size_t)k * S + d] = src[(size_t)(k0 + k) + (size_t)K * e];
}

// one row of an AoS array by class slot: dst[s] = src(row, refId[s]) (the marching class: ghost copies included)
__global__ void k_gather_rows(const double* __restrict__ src, int K, const int* __restrict__ refId, int S, double* __restrict__ dst, int row) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  const int e = refId[s];
  if (e >= 0) dst[s] = src[(size_t)row + (size_t)K * e];
}

// ----------------------------------------------------------------------------
// launch wrappers
static inline int nblk(long n, int b) { return (int)((n + b - 1) / b); }

void launch_advance(StepState* st, double dt, cudaStream_t s) { k_advance<<<1, 1, 0, s>>>(st, dt); }

template <bool SKIP>
static void launch_node3(const NodeArgs& A, cudaStream_t s) {
  if (A.list) {   // the nodes no marching bundle updates (the last sweep updated the others)
    if (A.nList > 0) k_node_update3<SKIP, true, 128, 10><<<nblk(A.nList, 128), 128, 0, s>>>(A);
    return;
  }
  // occupancy variants of the same kernel (eqd_set_option "node_variant"); the gather is latency bound
  switch (A.variant % 10) {
    case 1: k_node_update3<SKIP, false, 256, 5><<<nblk(A.Nn, 256), 256, 0, s>>>(A); break;
    case 2: k_node_update3<SKIP, false, 256, 6><<<nblk(A.Nn, 256), 256, 0, s>>>(A); break;
    case 3: k_node_update3<SKIP, false, 128, 8><<<nblk(A.Nn, 128), 128, 0, s>>>(A); break;
    case 4: k_node_update3<SKIP, false, 128, 10><<<nblk(A.Nn, 128), 128, 0, s>>>(A); break;
    case 5: k_node_update3<SKIP, false, 128, 12><<<nblk(A.Nn, 128), 128, 0, s>>>(A); break;
    default: k_node_update3<SKIP, false, 256, 4><<<nblk(A.Nn, 256), 256, 0, s>>>(A); break;
  }
}
template <bool SKIP>
static void launch_node12(const NodeArgs& A, cudaStream_t s) {
  switch (A.variant / 10) {
    case 1: k_node_update12<SKIP, 6><<<nblk(A.Np, 128), 128, 0, s>>>(A); break;
    case 2: k_node_update12<SKIP, 8><<<nblk(A.Np, 128), 128, 0, s>>>(A); break;
    default: k_node_update12<SKIP, 5><<<nblk(A.Np, 128), 128, 0, s>>>(A); break;
  }
}
void launch_node_update(const NodeArgs& A, cudaStream_t s) {
  if (A.skipSpecial) {
    if (A.Nn > 0) launch_node3<true>(A, s);
    if (A.Np > 0) launch_node12<true>(A, s);
  } else {
    if (A.Nn > 0) launch_node3<false>(A, s);
    if (A.Np > 0) launch_node12<false>(A, s);
  }
}
void launch_node_update_special(const NodeArgs& A, const int* list, int n, cudaStream_t s) {
  if (n > 0) k_node_update_special<<<nblk(n, 128), 128, 0, s>>>(A, list, n);
}
void launch_assemble_special(const NodeArgs& A, const int* list, int n, cudaStream_t s) {
  if (n > 0) k_assemble_special<<<nblk(n, 128), 128, 0, s>>>(A, list, n);
}
void launch_materialize_accel(const NodeArgs& A, double* out, cudaStream_t s) {
  if (A.Nn > 0) k_materialize_accel<<<nblk(A.Nn, 128), 128, 0, s>>>(A, out);
}

// dynamic shared memory of a tile kernel of class `cls` whose largest tile has LS nodes
size_t tile_smem_bytes(int cls, bool q, int LS) {
  if (cls == CLS_PML)
    return 128 + (size_t)PR_ROWS * EQD_STAGE_PML * sizeof(double) + 8 * EQD_STAGE_PML * sizeof(uint16_t) + (size_t)EQD_PML_LS * sizeof(int) +
           (size_t)21 * EQD_PML_LS * sizeof(double);
  (void)LS;
  return 128 + (size_t)RR_ROWS * EQD_STAGE * sizeof(double) + 8 * EQD_STAGE * sizeof(uint16_t) + (size_t)EQD_REG_LS * sizeof(int) +
         (size_t)(q ? 15 : 12) * EQD_REG_LS * sizeof(double);
}
static int sm_count() {
  static std::mutex mu;
  static std::map<int, int> cache;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> g(mu);
  auto it = cache.find(dev);
  if (it != cache.end()) return it->second;
  int n = 148;
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  cache[dev] = n;
  return n;
}
static void tile_launch(void (*kern)(ElemArgs), const ElemArgs& A, int ntiles, int threads, size_t sm, cudaStream_t s) {
  // opt in to the large dynamic shared memory once per (device, kernel)
  static std::mutex mu;
  static std::set<std::pair<int, const void*>> done;
  int dev = 0;
  cudaGetDevice(&dev);
  {
    std::lock_guard<std::mutex> g(mu);
    if (done.insert({dev, (const void*)kern}).second)
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024));
  }
  kern<<<ntiles, threads, sm, s>>>(A);
}
template <bool PL, bool Q, bool BODY, bool SPLIT>
static void launch_reg_chg(const ElemArgs& A, int chg, cudaStream_t s) {
  const int ntiles = A.ntiles;
  const size_t sm = tile_smem_bytes(CLS_REG, Q, A.LS);
  // persistent CTAs, two per SM, each walking tiles blockIdx.x, +gridDim.x, ...
  const int grid = std::min(ntiles, A.maxGrid > 0 ? A.maxGrid : 2 * sm_count());
  if (A.tileBox && A.allBox) {
    // every tile is a box tile: compact stage buffer, three CTAs per SM (option "box_compact")
    const size_t smc = sm - (size_t)(RR_ROWS - BOX_ROWS) * EQD_STAGE * sizeof(double);
    const int gridc = std::min(ntiles, A.maxGrid > 0 ? A.maxGrid : 3 * sm_count());
    if (chg == 2) tile_launch(k_tile_reg<PL, Q, BODY, SPLIT, 2, 2>, A, gridc, 2 * EQD_STAGE, smc, s);
    else tile_launch(k_tile_reg<PL, Q, BODY, SPLIT, 1, 2>, A, gridc, 2 * EQD_STAGE, smc, s);
    return;
  }
  if (A.tileBox) {
    // closed-form operators on the tiles flagged as all-box (eqd_set_option "box")
    if (chg == 2) tile_launch(k_tile_reg<PL, Q, BODY, SPLIT, 2, 1>, A, grid, 2 * EQD_STAGE, sm, s);
    else tile_launch(k_tile_reg<PL, Q, BODY, SPLIT, 1, 1>, A, grid, 2 * EQD_STAGE, sm, s);
    return;
  }
  if (chg == 2) tile_launch(k_tile_reg<PL, Q, BODY, SPLIT, 2, 0>, A, grid, 2 * EQD_STAGE, sm, s);
  else tile_launch(k_tile_reg<PL, Q, BODY, SPLIT, 1, 0>, A, grid, 2 * EQD_STAGE, sm, s);
}
template <bool SPLIT>
static void launch_reg_split(const ElemArgs& A, bool plastic, bool q, bool body, int chg, cudaStream_t s) {
  // C_Q==1 with C_elastic==0 is rejected by the reference (warning.f90:6-9)
  if (q) { if (body) launch_reg_chg<false, true, true, SPLIT>(A, chg, s); else launch_reg_chg<false, true, false, SPLIT>(A, chg, s); }
  else if (plastic) { if (body) launch_reg_chg<true, false, true, SPLIT>(A, chg, s); else launch_reg_chg<true, false, false, SPLIT>(A, chg, s); }
  else { if (body) launch_reg_chg<false, false, true, SPLIT>(A, chg, s); else launch_reg_chg<false, false, false, SPLIT>(A, chg, s); }
}
void launch_elem_reg(const ElemArgs& A, bool split, bool plastic, bool q, bool body, int chg, cudaStream_t s) {
  if (A.ntiles <= 0) return;
  if (split) launch_reg_split<true>(A, plastic, q, body, chg, s);
  else launch_reg_split<false>(A, plastic, q, body, chg, s);
}
void launch_elem_pml(const ElemArgs& A, bool body, int chg, cudaStream_t s) {
  int ntiles = A.ntiles;
  if (ntiles <= 0) return;
  const size_t sm = tile_smem_bytes(CLS_PML, false, A.LS);
  const int nt = 4 * EQD_STAGE_PML;
  ntiles = std::min(ntiles, A.maxGrid > 0 ? A.maxGrid : 2 * sm_count());
  if (A.tileBox) {
    if (body) { if (chg == 2) tile_launch(k_tile_pml<true, 2, true>, A, ntiles, nt, sm, s); else tile_launch(k_tile_pml<true, 1, true>, A, ntiles, nt, sm, s); }
    else { if (chg == 2) tile_launch(k_tile_pml<false, 2, true>, A, ntiles, nt, sm, s); else tile_launch(k_tile_pml<false, 1, true>, A, ntiles, nt, sm, s); }
    return;
  }
  if (body) { if (chg == 2) tile_launch(k_tile_pml<true, 2, false>, A, ntiles, nt, sm, s); else tile_launch(k_tile_pml<true, 1, false>, A, ntiles, nt, sm, s); }
  else { if (chg == 2) tile_launch(k_tile_pml<false, 2, false>, A, ntiles, nt, sm, s); else tile_launch(k_tile_pml<false, 1, false>, A, ntiles, nt, sm, s); }
}
void launch_store_offfault(const int* idhist, int n, double* out, const double* vel, const double* disp, int NnS,
                           const StepState* st, cudaStream_t s) {
  if (n > 0) k_store_offfault<<<nblk(n + 1, 128), 128, 0, s>>>(idhist, n, out, vel, disp, NnS, st);
}
void launch_sample_gm(const int* surf, int nSurf, const double* vel, int NnS, double* out, cudaStream_t s) {
  if (nSurf > 0) k_sample_gm<<<nblk(nSurf, 128), 128, 0, s>>>(surf, nSurf, vel, NnS, out);
}
void launch_sample_src(const double* fric, int PS, int n, double* out, cudaStream_t s) {
  if (n > 0) k_sample_src<<<nblk(n, 128), 128, 0, s>>>(fric, PS, n, out);
}
void launch_pack(const double* src, const uint32_t* idx, int n, double* buf, cudaStream_t s) {
  if (n > 0) k_pack<<<nblk(n, 256), 256, 0, s>>>(src, idx, n, buf);
}
void launch_unpack_add(double* dst, const uint32_t* idx, int n, const double* buf, cudaStream_t s) {
  if (n > 0) k_unpack_add<<<nblk(n, 256), 256, 0, s>>>(dst, idx, n, buf);
}
void launch_halo_send(const HaloAxisArgs& A, cudaStream_t s) {
  const int n = std::max(A.n[0], A.n[1]);
  if (n > 0) k_halo_send<<<dim3(nblk(n, 256), 2), 256, 0, s>>>(A);
}
void launch_halo_recv(const HaloAxisArgs& A, cudaStream_t s) {
  const int n = std::max(A.n[0], A.n[1]);
  if (n > 0) k_halo_recv<<<dim3(nblk(n, 256), 2), 256, 0, s>>>(A);
}
void launch_thermop(const FaultArgs& A, cudaStream_t s) {
  if (A.nPairs > 0) k_thermop<<<nblk(A.nPairs, 128), 128, 0, s>>>(A);
}
void launch_fault(const FaultArgs& A, cudaStream_t s) {
  // 178 registers per thread: 64-thread CTAs pack five to an SM where 128-thread ones pack two
  if (A.nPairs > 0) k_fault<<<nblk(A.nPairs, 64), 64, 0, s>>>(A);
}
void launch_gather_rows(const double* src, int K, const int* refId, int S, double* dst, int row, cudaStream_t s) {
  if (S > 0) k_gather_rows<<<nblk(S, 256), 256, 0, s>>>(src, K, refId, S, dst, row);
}
void launch_aos_to_soa(const double* src, int K, int n, const int* dstIdx, int cls, double* dst, int S, int k0, int nk,
                       cudaStream_t s) {
  if (n > 0) k_aos_to_soa<<<nblk(n, 256), 256, 0, s>>>(src, K, n, dstIdx, cls, dst, S, k0, nk);
}

}  // namespace eqd
