// Marching kernel for bundles of axis-aligned hexahedra: element sweep (calcElemKU.f90:44-189 elastic
// branch + hrglss.f90:20-54), assembly (assembleGlobalKU.f90:28-35) AND the central-difference update of
// the nodes that belong to one bundle only (driver.f90:29,89-104) in ONE pass over the data.
//
// EQdyna's built-in mesh is a structured grid (meshgen.f90:64-107, element order createElement :702-741);
// away from the fault and the PML every element is an axis-aligned box whose connectivity follows from its
// grid cell.  A BUNDLE is a block of bz x by such element columns (bz <= MK_BZ, by <= MK_BY), Lx elements
// long in x.  One CTA marches along x, one thread per column:
//   * the thread keeps, in registers, the x- face of its element in transformed form (a 4-point
//     Walsh-Hadamard transform of the four corner values: sums and differences along y and z) -- it was
//     the x+ face of the previous element -- and reads only the four x+ corners (v and d) from shared memory;
//   * the strain rate needs three, the hourglass modes four of the eight Hadamard coefficients of the
//     element's 8 corner values, each one add away from the two face transforms (eqd_box.h gives the sign
//     tables: eleshp = sign*a_d, phi = ha = products of the signs);
//   * the nodal forces are formed in the same transformed space: the force plane between elements p-1 and
//     p is  (x+ face of p-1, carried in registers) + (x- face of p), inverse-transformed once, and stored --
//     no read-modify-write, no atomics -- into one of four shared buffers (one per (dy,dz) corner);
//   * after a barrier the node plane is complete: F = ((b0 + b1) + b2) + b3 in that fixed order.  A node
//     interior to the bundle has then received all eight of its elements: it is updated in place
//     (v += F/m dt, d += v dt, the values are still in the shared ring).  A node on the bundle's surface
//     writes one partial force, summed with the other tiles' partials by the node-update kernel as before.
// Implicit local connectivity: an element's eight nodes are positions of the bundle's
// node lattice; neither nodeElemIdRelation nor a local copy of it is read.  The lattice's node ids are taken
// from the reference's connectivity when the bundle is planned and verified against it (eqd_march.cu).
//
// Data movement: operator rows of the next element plane arrive by bulk asynchronous copies (mbarrier),
// node planes two steps ahead by cp.async; per element-step the kernel reads 15 doubles of operators and
// stresses, ~1.2 nodes x (v, d, m, id) and writes 6 stresses + the updated v, d of the fused nodes.
//
// The kernel body is written once as barrier-separated PHASES, each a function of (thread id, that thread's
// registers).  nvcc runs them on the CTA's threads with __syncthreads() between them; g++ (tests/,
// tools/march_emul) runs every phase over all thread ids in turn, with the asynchronous copies done at issue
// time, so the CPU suite checks the indexing, the carries and the arithmetic of the source the GPU executes.
#pragma once
#include <cstddef>
#include <cstdint>

#include "eqd_box.h"
#include "eqd_dev.cuh"

namespace eqd {

constexpr int MK_NZ = 8, MK_NY = 16;                 // nodes of a bundle cross-section (z, y)
constexpr int MK_BZ = MK_NZ - 1, MK_BY = MK_NY - 1;  // element columns of a full bundle cross-section: 7 x 15
constexpr int MK_PN = MK_NZ * MK_NY;                 // nodes of a plane = threads of the CTA: thread t owns node t of every plane
constexpr int MK_NT = MK_PN;                         //   and the element column whose (y-, z-) corner is node t
constexpr int MK_ES = MK_BZ * MK_NY;                 // room of one operator row in the stage buffer (>= bz*by of any bundle)
constexpr int MK_OPROWS = 15;                        // a_x a_y a_z ss1 ss4 ss6 lam mu det stress(6)
constexpr int MK_FUSED = 0x40000000;                 // node code: updated by the bundle itself
constexpr int MK_GHOST = 0x20000000;                 // node code: read only -- a neighbouring strip reports this node
constexpr int MK_IDMASK = 0x1fffffff;
constexpr int MK_MINLX = 4;                          // shortest bundle the planner makes

struct MarchBundle {
  int e0;      // first element slot of the class SoA (even): slot = e0 + p*mk_es(B) + cz*by + cy -- the planes of a bundle are
               // packed (no slots for inactive columns), so an operator row of a plane is one contiguous, 16-byte aligned run
  int n0;      // first node slot: slot = n0 + p*MK_PN + iz*MK_NY + iy
  int Lx;      // elements along x (node planes 0..Lx)
  int shape;   // bz | by << 8 | orient << 16: active columns cz < bz, cy < by.  orient 0: the node plane is MK_NZ rows of MK_NY
               // nodes (bz <= 7, by <= 15); orient 1 (PML bundles only): MK_NY rows of MK_NZ nodes (bz <= 15, by <= 7), for slabs
               // that are thin in y
};

struct MarchArgs {
  const MarchBundle* rec;
  const int* ctaFirstA;     // [grid+1]: CTA b marches bundles rec[ctaFirstA[b]] .. rec[ctaFirstA[b+1]-1] (boundary work list)
  const int* ctaFirstB;     // [grid+1]: ... then rec[ctaFirstB[b]] .. (interior work list); either may be null
  const int* code;          // [PFS] per node slot: -1 = no node, else node id | MK_FUSED
  size_t S, NnS, PFS;
  const double* a;          // [3][S] a_x, a_y, a_z  (eleshp rows 3, 7, 14)
  const double* ss;         // [3][S] ss1, ss4, ss6
  const double* lam; const double* mu; const double* det;
  double* stress;           // [6][S]
  const double* vel; const double* disp;   // [3][NnS] v(nt), d(nt): read by every bundle that touches the node
  double* velOut; double* dispOut;         // [3][NnS] where the updated nodes go (another buffer when strips share ghosts)
  const double* mass;       // [Nn]
  double* pf;               // [3][PFS] partial force of every non-fused node slot
  double* force;            // [3][NnS] complete force of the fused nodes when update == 0
  double dt, rdampk, w;
  int update;               // 1: fused nodes are updated in place; 0: their force goes to `force` (last step of a run)
  StepState* st;
};

struct MarchShared {
  double ops[2][MK_OPROWS][MK_ES];     // operator stage of element plane p in ops[p & 1]
  double ring[3][6][MK_PN];            // node plane pl in ring[pl % 3]: v(3) d(3)
  double frc[4][3][MK_PN];             // force plane, one buffer per (dy,dz) corner
  unsigned long long bar[2];           // mbarrier of each operator stage
};

// what a thread carries from one element to the next
struct MarchRegs {
  double wv[3][3];   // [S0 | Sy | Sz][component] of the x- face velocities
  double wl[3][3];   // [Sy | Sz | Syz][component] of the x- face l = d + rdampk v
  double cf[4][3];   // x+ face forces of the previous element, transformed: [1 | sy | sz | sy sz][component]
  int c0, c1, c2, c3;   // codes of this thread's node in planes p, p+1, p+2, p+3
  double m;          // lumped mass of its node in plane p
};

// ---- asynchronous copies: real on the device, immediate in the host reading
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ unsigned mk_s32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mk_async8(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(mk_s32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void mk_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void mk_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void mk_bar_init(unsigned long long* bar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mk_s32(bar)));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mk_bar_expect(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mk_s32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mk_bar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "MKW_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra MKD_%=;\n"
      "bra MKW_%=;\n"
      "MKD_%=:\n"
      "}\n" ::"r"(mk_s32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void mk_bulk(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(mk_s32(dst)), "l"(src),
               "r"(bytes), "r"(mk_s32(bar))
               : "memory");
}
__device__ __forceinline__ void mk_nan(StepState* st, int node) {
  if (atomicExch(&st->nanFlag, 1) == 0) st->nanNode = node + 1;
}
#else
EQD_HD void mk_async8(void* dst, const void* src) { *(double*)dst = *(const double*)src; }
EQD_HD void mk_commit() {}
EQD_HD void mk_wait_all() {}
EQD_HD void mk_bar_init(unsigned long long*) {}
EQD_HD void mk_bar_expect(unsigned long long*, unsigned) {}
EQD_HD void mk_bar_wait(unsigned long long*, unsigned) {}
EQD_HD void mk_bulk(void* dst, const void* src, unsigned bytes, unsigned long long*) {
  for (unsigned k = 0; k < bytes / 8; ++k) ((double*)dst)[k] = ((const double*)src)[k];
}
EQD_HD void mk_nan(StepState* st, int node) {
  if (!st->nanFlag) { st->nanFlag = 1; st->nanNode = node + 1; }
}
#endif

EQD_HD int mk_bz(const MarchBundle& B) { return B.shape & 0xff; }
EQD_HD int mk_by(const MarchBundle& B) { return (B.shape >> 8) & 0xff; }
EQD_HD int mk_es(const MarchBundle& B) { return (mk_bz(B) * mk_by(B) + 1) & ~1; }   // element slots per plane
EQD_HD int mk_rowlen(const MarchBundle& B) { return ((B.shape >> 16) & 1) ? MK_NZ : MK_NY; }   // nodes per row of the node plane
EQD_HD const double* mk_op_row(const MarchArgs& A, int r) {
  if (r < 3) return A.a + (size_t)r * A.S;
  if (r < 6) return A.ss + (size_t)(r - 3) * A.S;
  if (r == 6) return A.lam;
  if (r == 7) return A.mu;
  if (r == 8) return A.det;
  return A.stress + (size_t)(r - 9) * A.S;
}

// code of this thread's node in plane pl (-1 beyond the bundle)
EQD_HD int mk_code(const MarchArgs& A, const MarchBundle& B, int tid, int pl) {
  return pl <= B.Lx ? A.code[(size_t)B.n0 + (size_t)pl * MK_PN + tid] : -1;
}
// ---- issue: v, d of this thread's node (code) -> ring slot rs
EQD_HD void mk_issue_values(const MarchArgs& A, MarchShared& sm, int tid, int rs, int code) {
  double(*dst)[MK_PN] = sm.ring[rs];
  if (code >= 0) {
    const double* v = A.vel + (size_t)(code & MK_IDMASK);
    const double* d = A.disp + (size_t)(code & MK_IDMASK);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      mk_async8(&dst[c][tid], v + c * A.NnS);
      mk_async8(&dst[3 + c][tid], d + c * A.NnS);
    }
  } else {
#pragma unroll
    for (int c = 0; c < 6; ++c) dst[c][tid] = 0.0;
  }
}
// ---- issue: operator rows of element plane p -> ops[p & 1]
EQD_HD void mk_issue_ops(const MarchArgs& A, const MarchBundle& B, MarchShared& sm, int tid, int p) {
  if (p >= B.Lx) return;
  unsigned long long* bar = &sm.bar[p & 1];
  const int es = mk_es(B);
  if (tid == 0) mk_bar_expect(bar, (unsigned)(MK_OPROWS * es * sizeof(double)));
#if defined(__CUDA_ARCH__)
  __syncwarp();
#endif
  if (tid < MK_OPROWS)
    mk_bulk(sm.ops[p & 1][tid], mk_op_row(A, tid) + (size_t)B.e0 + (size_t)p * es, (unsigned)(es * sizeof(double)), bar);
}

// 4-point transform of a face: a[q], q = dy + 2 dz
#define MK_FACE_SUMS(a0, a1, a2, a3)          \
  const double s0_ = (a0) + (a1), s1_ = (a2) + (a3), d0_ = (a1) - (a0), d1_ = (a3) - (a2);

EQD_HD bool mk_active(const MarchBundle& B, int tid) { return (tid / MK_NY) < mk_bz(B) && (tid % MK_NY) < mk_by(B); }

// ---- phase: bundle start.  Clears the force buffers and takes the transformed x- face of the first element.
EQD_HD void mk_phase_begin(const MarchArgs& A, const MarchBundle& B, MarchShared& sm, MarchRegs& R, int tid) {
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int c = 0; c < 3; ++c) sm.frc[q][c][tid] = 0.0;
  if (!mk_active(B, tid)) return;
  double(*pl)[MK_PN] = sm.ring[0];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const double v0 = pl[c][tid], v1 = pl[c][tid + 1], v2 = pl[c][tid + MK_NY], v3 = pl[c][tid + MK_NY + 1];
    {
      MK_FACE_SUMS(v0, v1, v2, v3)
      R.wv[0][c] = s0_ + s1_; R.wv[1][c] = d0_ + d1_; R.wv[2][c] = s1_ - s0_;
    }
    const double m0 = pl[3 + c][tid] + A.rdampk * v0, m1 = pl[3 + c][tid + 1] + A.rdampk * v1;
    const double m2 = pl[3 + c][tid + MK_NY] + A.rdampk * v2, m3 = pl[3 + c][tid + MK_NY + 1] + A.rdampk * v3;
    {
      MK_FACE_SUMS(m0, m1, m2, m3)
      R.wl[0][c] = d0_ + d1_; R.wl[1][c] = s1_ - s0_; R.wl[2][c] = d1_ - d0_;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) R.cf[k][c] = 0.0;
  }
}

// inverse face transform of G[k], k = 1 | sy | sz | sy sz, stored to the four corner buffers
EQD_HD void mk_store_face(MarchShared& sm, int c, int tid, double G0, double G1, double G2, double G3) {
  const double um = G0 - G1, up = G0 + G1, wm = G2 - G3, wp = G2 + G3;
  sm.frc[0][c][tid] = um - wm;                      // (y-, z-)
  sm.frc[1][c][tid + 1] = up - wp;                  // (y+, z-)
  sm.frc[2][c][tid + MK_NY] = um + wm;              // (y-, z+)
  sm.frc[3][c][tid + MK_NY + 1] = up + wp;          // (y+, z+)
}

// ---- phase: element p of this thread's column; rs1 = ring slot of node plane p+1
EQD_HD void mk_phase_element(const MarchArgs& A, const MarchBundle& B, MarchShared& sm, MarchRegs& R, int tid, int p, int rs1) {
  if (!mk_active(B, tid)) return;
  double(*pl)[MK_PN] = sm.ring[rs1];
  const double(*op)[MK_ES] = sm.ops[p & 1];
  double gx[3], gy[3], gz[3], P[4][3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const double v0 = pl[c][tid], v1 = pl[c][tid + 1], v2 = pl[c][tid + MK_NY], v3 = pl[c][tid + MK_NY + 1];
    {
      MK_FACE_SUMS(v0, v1, v2, v3)
      const double n0 = s0_ + s1_, ny = d0_ + d1_, nz = s1_ - s0_;
      // g[c][d] = sum_i sign_d(i) v_i[c]: x+ face minus x- face, y and z differences of both faces
      gx[c] = n0 - R.wv[0][c]; gy[c] = ny + R.wv[1][c]; gz[c] = nz + R.wv[2][c];
      R.wv[0][c] = n0; R.wv[1][c] = ny; R.wv[2][c] = nz;
    }
    // hrglss.f90:20-27: l = d + rdampk v
    const double m0 = pl[3 + c][tid] + A.rdampk * v0, m1 = pl[3 + c][tid + 1] + A.rdampk * v1;
    const double m2 = pl[3 + c][tid + MK_NY] + A.rdampk * v2, m3 = pl[3 + c][tid + MK_NY + 1] + A.rdampk * v3;
    {
      MK_FACE_SUMS(m0, m1, m2, m3)
      const double ny = d0_ + d1_, nz = s1_ - s0_, nyz = d1_ - d0_;
      // phid(m) = sum_i ha(m,i) l_i, ha = sy sz | sx sz | sx sy | sx sy sz (assembleGlobalMass.f90:336-339)
      P[0][c] = nyz + R.wl[2][c]; P[1][c] = nz - R.wl[1][c]; P[2][c] = ny - R.wl[0][c]; P[3][c] = nyz - R.wl[2][c];
      R.wl[0][c] = ny; R.wl[1][c] = nz; R.wl[2][c] = nyz;
    }
  }
  const int es = (tid / MK_NY) * mk_by(B) + (tid % MK_NY);   // element slot inside the (packed) plane
  const double ax = op[0][es], ay = op[1][es], az = op[2][es];
  // calcElemKU.f90:44-60 with eleshp(d,i) = sign_d(i) a_d
  double sr[6];
  sr[0] = ax * gx[0];
  sr[1] = ay * gy[1];
  sr[2] = az * gz[2];
  sr[3] = az * gz[1] + ay * gy[2];
  sr[4] = az * gz[0] + ax * gx[2];
  sr[5] = ay * gy[0] + ax * gx[1];
  const double lam = op[6][es], mu = op[7][es], l2m = lam + 2 * mu;
  double rate[6];   // calcElemKU.f90:63-70
  rate[0] = 0.0 + l2m * sr[0] + lam * sr[1] + lam * sr[2];
  rate[1] = 0.0 + lam * sr[0] + l2m * sr[1] + lam * sr[2];
  rate[2] = 0.0 + lam * sr[0] + lam * sr[1] + l2m * sr[2];
  rate[3] = mu * sr[3];
  rate[4] = mu * sr[4];
  rate[5] = mu * sr[5];
  const double temp = (-op[8][es]) * A.w;   // calcElemKU.f90:169-173, constk = -eledet
  double* sp = A.stress + (size_t)B.e0 + (size_t)p * mk_es(B) + es;
  double t[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const double sg = op[9 + k][es] + rate[k] * A.dt;   // :72-76
    sp[(size_t)k * A.S] = sg;
    t[k] = temp * (sg + A.rdampk * rate[k]);
  }
  const double ssd[3] = {op[3][es], op[4][es], op[5][es]};
  // B^T t (calcElemKU.f90:175-189): f_c = sx X + sy Y + sz Z ; hourglass (hrglss.f90:35-54): - sum_m ha(m) ss_c P_m
  const double X[3] = {ax * t[0], ax * t[5], ax * t[4]};
  const double Y[3] = {ay * t[5], ay * t[1], ay * t[3]};
  const double Z[3] = {az * t[4], az * t[3], az * t[2]};
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const double H0 = ssd[c] * P[0][c], H1 = ssd[c] * P[1][c], H2 = ssd[c] * P[2][c], H3 = ssd[c] * P[3][c];
    // x- face of this element (sx = -1) + the carried x+ face of the previous one
    const double G0 = R.cf[0][c] - X[c], G1 = R.cf[1][c] + (Y[c] + H2), G2 = R.cf[2][c] + (Z[c] + H1), G3 = R.cf[3][c] + (H3 - H0);
    mk_store_face(sm, c, tid, G0, G1, G2, G3);
    R.cf[0][c] = X[c]; R.cf[1][c] = Y[c] - H2; R.cf[2][c] = Z[c] - H1; R.cf[3][c] = 0.0 - H0 - H3;
  }
}

// ---- phase: the last node plane receives the carried x+ faces only
EQD_HD void mk_phase_last(const MarchBundle& B, MarchShared& sm, MarchRegs& R, int tid) {
  if (!mk_active(B, tid)) return;
#pragma unroll
  for (int c = 0; c < 3; ++c) mk_store_face(sm, c, tid, R.cf[0][c], R.cf[1][c], R.cf[2][c], R.cf[3][c]);
}

// ---- phase: node plane pl is complete; rs = its ring slot, code / m = this thread's node
EQD_HD void mk_phase_flush(const MarchArgs& A, const MarchBundle& B, MarchShared& sm, int tid, int pl, int rs, int code, double m) {
  if (code < 0 || (code & MK_GHOST)) return;
  double F[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) F[c] = ((sm.frc[0][c][tid] + sm.frc[1][c][tid]) + sm.frc[2][c][tid]) + sm.frc[3][c][tid];
  const size_t n = (size_t)(code & MK_IDMASK);
  if (code & MK_FUSED) {
    if (A.update) {
      double(*rg)[MK_PN] = sm.ring[rs];
      bool bad = false;
      const double rm = 1.0 / m;              // one reciprocal instead of three divisions: f/m within one rounding
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        double v = rg[c][tid], d = rg[3 + c][tid];
        v = v + (F[c] * rm) * A.dt;           // driver.f90:29,102
        d = d + v * A.dt;                     // :104
        bad |= (v != v);
        A.velOut[c * A.NnS + n] = v; A.dispOut[c * A.NnS + n] = d;
      }
      if (bad) mk_nan(A.st, (int)n);
    } else {
#pragma unroll
      for (int c = 0; c < 3; ++c) A.force[c * A.NnS + n] = F[c];
    }
  } else {
    const size_t slot = (size_t)B.n0 + (size_t)pl * MK_PN + tid;
#pragma unroll
    for (int c = 0; c < 3; ++c) A.pf[c * A.PFS + slot] = F[c];
  }
}

// The schedule.  RUN(body) executes `body` for every thread of the CTA and then synchronises: a statement +
// __syncthreads() on the device, a loop over the thread ids on the host; RUNNS(body) does the same without
// the barrier.  WAIT_NODES / WAIT_OPS(p) are the cp.async / mbarrier waits (nothing on the host, where copies
// complete at issue).  Thread t owns node t of every plane: it fetches the node's code (three planes ahead,
// into a register), issues the copies of its v, d (two planes ahead, into the shared ring), loads its mass
// (during the element work of the step that completes it) and finally flushes it.
//   step p:  [plane p+1 and operator stage p have landed] barrier
//            issue plane p+2 values and operator stage p+1, load the mass of plane p ; element p -> force buffers ; barrier
//            flush node plane p (update or partial) ; fetch the code of plane p+4
// Ring slot (p+2)%3 was last read by the flush of plane p-1, the operator stage (p+1)&1 by element p-1, the
// force buffers by that flush: all before the barrier that opens step p.
#define MARCH_BUNDLE(RUN, RUNNS, WAIT_NODES, WAIT_OPS, A, B, sm, R)                                                     \
  do {                                                                                                                 \
    RUN(R.c0 = mk_code(A, B, tid, 0); R.c1 = mk_code(A, B, tid, 1); R.c2 = mk_code(A, B, tid, 2); R.c3 = mk_code(A, B, tid, 3); \
        mk_issue_values(A, sm, tid, 0, R.c0); mk_issue_values(A, sm, tid, 1, R.c1); mk_issue_ops(A, B, sm, tid, 0);     \
        mk_commit(); WAIT_NODES);                                                                                      \
    RUN(mk_phase_begin(A, B, sm, R, tid));                                                                             \
    for (int p = 0, rs0 = 0, rs1 = 1, rs2 = 2; p < (B).Lx; ++p) {                                                      \
      RUN(WAIT_NODES; WAIT_OPS(p));                                                                                    \
      RUN(mk_issue_values(A, sm, tid, rs2, R.c2); mk_issue_ops(A, B, sm, tid, p + 1); mk_commit();                     \
          R.m = R.c0 >= 0 ? A.mass[R.c0 & MK_IDMASK] : 1.0; mk_phase_element(A, B, sm, R, tid, p, rs1));               \
      RUNNS(mk_phase_flush(A, B, sm, tid, p, rs0, R.c0, R.m);                                                          \
            R.c0 = R.c1; R.c1 = R.c2; R.c2 = R.c3; R.c3 = mk_code(A, B, tid, p + 4));                                  \
      { const int t_ = rs0; rs0 = rs1; rs1 = rs2; rs2 = t_; }                                                          \
    }                                                                                                                  \
    RUN(WAIT_NODES);                                                                                                   \
    RUN(mk_phase_last(B, sm, R, tid));                                                                                 \
    RUN(mk_phase_flush(A, B, sm, tid, (B).Lx, (B).Lx % 3, R.c0, 1.0));                                                 \
  } while (0)

}  // namespace eqd
