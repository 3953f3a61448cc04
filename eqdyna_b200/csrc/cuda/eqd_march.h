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

constexpr int MK_BZ = 8, MK_BY = 16;                 // columns of a full bundle cross-section (z, y)
constexpr int MK_NCOL = MK_BZ * MK_BY;               // = threads per CTA = element slots per plane
constexpr int MK_NT = MK_NCOL;
constexpr int MK_PN = (MK_BZ + 1) * (MK_BY + 1);     // nodes of a plane
constexpr int MK_PNP = (MK_PN + 1) & ~1;             // padded to an even count (16-byte rows)
constexpr int MK_OPROWS = 15;                        // a_x a_y a_z ss1 ss4 ss6 lam mu det stress(6)
constexpr int MK_FUSED = 0x40000000;                 // node code: updated by the bundle itself
constexpr int MK_IDMASK = 0x3fffffff;
constexpr int MK_MINLX = 4;                          // shortest bundle the planner makes

struct MarchBundle {
  int e0;      // first element slot of the class SoA (multiple of MK_NCOL): slot = e0 + p*MK_NCOL + cz*MK_BY + cy
  int n0;      // first node slot: slot = n0 + p*MK_PN + iz*(MK_BY+1) + iy
  int Lx;      // elements along x (node planes 0..Lx)
  int shape;   // bz | by << 8: active columns cz < bz, cy < by
};

struct MarchArgs {
  const MarchBundle* rec;
  const int* ctaFirst;      // [grid+1]: CTA b marches bundles rec[ctaFirst[b]] .. rec[ctaFirst[b+1]-1]
  const int* code;          // [PFS] per node slot: -1 = no node, else node id | MK_FUSED
  size_t S, NnS, PFS;
  const double* a;          // [3][S] a_x, a_y, a_z  (eleshp rows 3, 7, 14)
  const double* ss;         // [3][S] ss1, ss4, ss6
  const double* lam; const double* mu; const double* det;
  double* stress;           // [6][S]
  double* vel; double* disp;   // [3][NnS]
  const double* mass;       // [Nn]
  double* pf;               // [3][PFS] partial force of every non-fused node slot
  double* force;            // [3][NnS] complete force of the fused nodes when update == 0
  double dt, rdampk, w;
  int update;               // 1: fused nodes are updated in place; 0: their force goes to `force` (last step of a run)
  StepState* st;
};

struct MarchShared {
  double ops[2][MK_OPROWS][MK_NCOL];   // operator stage of element plane p in ops[p & 1]
  double ring[3][7][MK_PNP];           // node plane pl in ring[pl % 3]: v(3) d(3) m
  double frc[4][3][MK_PNP];            // force plane, one buffer per (dy,dz) corner
  int ids[4][MK_PNP];                  // node codes of plane pl in ids[pl & 3]
  unsigned long long bar[2];           // mbarrier of each operator stage
};

// what a thread carries from one element to the next
struct MarchRegs {
  double wv[3][3];   // [S0 | Sy | Sz][component] of the x- face velocities
  double wl[3][3];   // [Sy | Sz | Syz][component] of the x- face l = d + rdampk v
  double cf[4][3];   // x+ face forces of the previous element, transformed: [1 | sy | sz | sy sz][component]
};

// ---- asynchronous copies: real on the device, immediate in the host reading
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ unsigned mk_s32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mk_async8(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(mk_s32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void mk_async4(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(mk_s32(dst)), "l"(src) : "memory");
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
EQD_HD void mk_async4(void* dst, const void* src) { *(int*)dst = *(const int*)src; }
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
EQD_HD const double* mk_op_row(const MarchArgs& A, int r) {
  if (r < 3) return A.a + (size_t)r * A.S;
  if (r < 6) return A.ss + (size_t)(r - 3) * A.S;
  if (r == 6) return A.lam;
  if (r == 7) return A.mu;
  if (r == 8) return A.det;
  return A.stress + (size_t)(r - 9) * A.S;
}

// ---- issue: node codes of plane pl -> ids[pl & 3]
EQD_HD void mk_issue_ids(const MarchArgs& A, const MarchBundle& B, MarchShared& sm, int tid, int pl) {
  if (pl > B.Lx) return;
  const int* src = A.code + (size_t)B.n0 + (size_t)pl * MK_PN;
  for (int i = tid; i < MK_PN; i += MK_NT) mk_async4(&sm.ids[pl & 3][i], src + i);
}
// ---- issue: v, d, m of plane pl -> ring[pl % 3] (its codes must have landed)
EQD_HD void mk_issue_values(const MarchArgs& A, const MarchBundle& B, MarchShared& sm, int tid, int pl) {
  if (pl > B.Lx) return;
  double(*dst)[MK_PNP] = sm.ring[pl % 3];
  for (int i = tid; i < MK_PN; i += MK_NT) {
    const int code = sm.ids[pl & 3][i];
    if (code >= 0) {
      const size_t n = (size_t)(code & MK_IDMASK);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        mk_async8(&dst[c][i], A.vel + c * A.NnS + n);
        mk_async8(&dst[3 + c][i], A.disp + c * A.NnS + n);
      }
      mk_async8(&dst[6][i], A.mass + n);
    } else {
#pragma unroll
      for (int c = 0; c < 6; ++c) dst[c][i] = 0.0;
      dst[6][i] = 1.0;
    }
  }
}
// ---- issue: operator rows of element plane p -> ops[p & 1]
EQD_HD void mk_issue_ops(const MarchArgs& A, const MarchBundle& B, MarchShared& sm, int tid, int p) {
  if (p >= B.Lx) return;
  unsigned long long* bar = &sm.bar[p & 1];
  if (tid == 0) mk_bar_expect(bar, (unsigned)(MK_OPROWS * MK_NCOL * sizeof(double)));
#if defined(__CUDA_ARCH__)
  __syncwarp();
#endif
  if (tid < MK_OPROWS)
    mk_bulk(sm.ops[p & 1][tid], mk_op_row(A, tid) + (size_t)B.e0 + (size_t)p * MK_NCOL, (unsigned)(MK_NCOL * sizeof(double)), bar);
}

// 4-point transform of a face: a[q], q = dy + 2 dz
#define MK_FACE_SUMS(a0, a1, a2, a3)          \
  const double s0_ = (a0) + (a1), s1_ = (a2) + (a3), d0_ = (a1) - (a0), d1_ = (a3) - (a2);

// ---- phase: bundle start.  Clears the force buffers and takes the transformed x- face of the first element.
EQD_HD void mk_phase_begin(const MarchArgs& A, const MarchBundle& B, MarchShared& sm, MarchRegs& R, int tid) {
  for (int i = tid; i < MK_PNP; i += MK_NT)
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int c = 0; c < 3; ++c) sm.frc[q][c][i] = 0.0;
  const int cz = tid / MK_BY, cy = tid - cz * MK_BY;
  if (cz >= mk_bz(B) || cy >= mk_by(B)) return;
  const int l0 = cz * (MK_BY + 1) + cy;
  double(*pl)[MK_PNP] = sm.ring[0];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const double v0 = pl[c][l0], v1 = pl[c][l0 + 1], v2 = pl[c][l0 + MK_BY + 1], v3 = pl[c][l0 + MK_BY + 2];
    {
      MK_FACE_SUMS(v0, v1, v2, v3)
      R.wv[0][c] = s0_ + s1_; R.wv[1][c] = d0_ + d1_; R.wv[2][c] = s1_ - s0_;
    }
    const double m0 = pl[3 + c][l0] + A.rdampk * v0, m1 = pl[3 + c][l0 + 1] + A.rdampk * v1;
    const double m2 = pl[3 + c][l0 + MK_BY + 1] + A.rdampk * v2, m3 = pl[3 + c][l0 + MK_BY + 2] + A.rdampk * v3;
    {
      MK_FACE_SUMS(m0, m1, m2, m3)
      R.wl[0][c] = d0_ + d1_; R.wl[1][c] = s1_ - s0_; R.wl[2][c] = d1_ - d0_;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) R.cf[k][c] = 0.0;
  }
}

// inverse face transform of G[k], k = 1 | sy | sz | sy sz, stored to the four corner buffers
EQD_HD void mk_store_face(MarchShared& sm, int c, int l0, double G0, double G1, double G2, double G3) {
  const double um = G0 - G1, up = G0 + G1, wm = G2 - G3, wp = G2 + G3;
  sm.frc[0][c][l0] = um - wm;                       // (y-, z-)
  sm.frc[1][c][l0 + 1] = up - wp;                   // (y+, z-)
  sm.frc[2][c][l0 + MK_BY + 1] = um + wm;           // (y-, z+)
  sm.frc[3][c][l0 + MK_BY + 2] = up + wp;           // (y+, z+)
}

// ---- phase: element p of this thread's column
EQD_HD void mk_phase_element(const MarchArgs& A, const MarchBundle& B, MarchShared& sm, MarchRegs& R, int tid, int p) {
  const int cz = tid / MK_BY, cy = tid - cz * MK_BY;
  if (cz >= mk_bz(B) || cy >= mk_by(B)) return;
  const int l0 = cz * (MK_BY + 1) + cy;
  double(*pl)[MK_PNP] = sm.ring[(p + 1) % 3];
  const double(*op)[MK_NCOL] = sm.ops[p & 1];
  double gx[3], gy[3], gz[3], P[4][3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const double v0 = pl[c][l0], v1 = pl[c][l0 + 1], v2 = pl[c][l0 + MK_BY + 1], v3 = pl[c][l0 + MK_BY + 2];
    {
      MK_FACE_SUMS(v0, v1, v2, v3)
      const double n0 = s0_ + s1_, ny = d0_ + d1_, nz = s1_ - s0_;
      // g[c][d] = sum_i sign_d(i) v_i[c]: x+ face minus x- face, y and z differences of both faces
      gx[c] = n0 - R.wv[0][c]; gy[c] = ny + R.wv[1][c]; gz[c] = nz + R.wv[2][c];
      R.wv[0][c] = n0; R.wv[1][c] = ny; R.wv[2][c] = nz;
    }
    // hrglss.f90:20-27: l = d + rdampk v
    const double m0 = pl[3 + c][l0] + A.rdampk * v0, m1 = pl[3 + c][l0 + 1] + A.rdampk * v1;
    const double m2 = pl[3 + c][l0 + MK_BY + 1] + A.rdampk * v2, m3 = pl[3 + c][l0 + MK_BY + 2] + A.rdampk * v3;
    {
      MK_FACE_SUMS(m0, m1, m2, m3)
      const double ny = d0_ + d1_, nz = s1_ - s0_, nyz = d1_ - d0_;
      // phid(m) = sum_i ha(m,i) l_i, ha = sy sz | sx sz | sx sy | sx sy sz (assembleGlobalMass.f90:336-339)
      P[0][c] = nyz + R.wl[2][c]; P[1][c] = nz - R.wl[1][c]; P[2][c] = ny - R.wl[0][c]; P[3][c] = nyz - R.wl[2][c];
      R.wl[0][c] = ny; R.wl[1][c] = nz; R.wl[2][c] = nyz;
    }
  }
  const double ax = op[0][tid], ay = op[1][tid], az = op[2][tid];
  // calcElemKU.f90:44-60 with eleshp(d,i) = sign_d(i) a_d
  double sr[6];
  sr[0] = ax * gx[0];
  sr[1] = ay * gy[1];
  sr[2] = az * gz[2];
  sr[3] = az * gz[1] + ay * gy[2];
  sr[4] = az * gz[0] + ax * gx[2];
  sr[5] = ay * gy[0] + ax * gx[1];
  const double lam = op[6][tid], mu = op[7][tid], l2m = lam + 2 * mu;
  double rate[6];   // calcElemKU.f90:63-70
  rate[0] = 0.0 + l2m * sr[0] + lam * sr[1] + lam * sr[2];
  rate[1] = 0.0 + lam * sr[0] + l2m * sr[1] + lam * sr[2];
  rate[2] = 0.0 + lam * sr[0] + lam * sr[1] + l2m * sr[2];
  rate[3] = mu * sr[3];
  rate[4] = mu * sr[4];
  rate[5] = mu * sr[5];
  const double temp = (-op[8][tid]) * A.w;   // calcElemKU.f90:169-173, constk = -eledet
  const size_t e = (size_t)B.e0 + (size_t)p * MK_NCOL + tid;
  double t[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const double sg = op[9 + k][tid] + rate[k] * A.dt;   // :72-76
    A.stress[(size_t)k * A.S + e] = sg;
    t[k] = temp * (sg + A.rdampk * rate[k]);
  }
  const double ssd[3] = {op[3][tid], op[4][tid], op[5][tid]};
  // B^T t (calcElemKU.f90:175-189): f_c = sx X + sy Y + sz Z ; hourglass (hrglss.f90:35-54): - sum_m ha(m) ss_c P_m
  const double X[3] = {ax * t[0], ax * t[5], ax * t[4]};
  const double Y[3] = {ay * t[5], ay * t[1], ay * t[3]};
  const double Z[3] = {az * t[4], az * t[3], az * t[2]};
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const double H0 = ssd[c] * P[0][c], H1 = ssd[c] * P[1][c], H2 = ssd[c] * P[2][c], H3 = ssd[c] * P[3][c];
    // x- face of this element (sx = -1) + the carried x+ face of the previous one
    const double G0 = R.cf[0][c] - X[c], G1 = R.cf[1][c] + (Y[c] + H2), G2 = R.cf[2][c] + (Z[c] + H1), G3 = R.cf[3][c] + (H3 - H0);
    mk_store_face(sm, c, l0, G0, G1, G2, G3);
    R.cf[0][c] = X[c]; R.cf[1][c] = Y[c] - H2; R.cf[2][c] = Z[c] - H1; R.cf[3][c] = 0.0 - H0 - H3;
  }
}

// ---- phase: the last node plane receives the carried x+ faces only
EQD_HD void mk_phase_last(const MarchBundle& B, MarchShared& sm, MarchRegs& R, int tid) {
  const int cz = tid / MK_BY, cy = tid - cz * MK_BY;
  if (cz >= mk_bz(B) || cy >= mk_by(B)) return;
  const int l0 = cz * (MK_BY + 1) + cy;
#pragma unroll
  for (int c = 0; c < 3; ++c) mk_store_face(sm, c, l0, R.cf[0][c], R.cf[1][c], R.cf[2][c], R.cf[3][c]);
}

// ---- phase: node plane pl is complete
EQD_HD void mk_phase_flush(const MarchArgs& A, const MarchBundle& B, MarchShared& sm, int tid, int pl) {
  double(*rg)[MK_PNP] = sm.ring[pl % 3];
  for (int i = tid; i < MK_PN; i += MK_NT) {
    const int code = sm.ids[pl & 3][i];
    if (code < 0) continue;
    double F[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) F[c] = ((sm.frc[0][c][i] + sm.frc[1][c][i]) + sm.frc[2][c][i]) + sm.frc[3][c][i];
    const size_t n = (size_t)(code & MK_IDMASK);
    if (code & MK_FUSED) {
      if (A.update) {
        const double m = rg[6][i];
        bool bad = false;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          double v = rg[c][i], d = rg[3 + c][i];
          v = v + (F[c] / m) * A.dt;            // driver.f90:29,102
          d = d + v * A.dt;                     // :104
          bad |= (v != v);
          A.vel[c * A.NnS + n] = v; A.disp[c * A.NnS + n] = d;
        }
        if (bad) mk_nan(A.st, (int)n);
      } else {
#pragma unroll
        for (int c = 0; c < 3; ++c) A.force[c * A.NnS + n] = F[c];
      }
    } else {
      const size_t slot = (size_t)B.n0 + (size_t)pl * MK_PN + i;
#pragma unroll
      for (int c = 0; c < 3; ++c) A.pf[c * A.PFS + slot] = F[c];
    }
  }
}

// The schedule.  RUN(body) executes `body` for every thread of the CTA and then synchronises: a statement +
// __syncthreads() on the device, a loop over the thread ids on the host; RUNNS(body) does the same without
// the barrier.  WAIT_NODES / WAIT_OPS(p) are the cp.async / mbarrier waits (nothing on the host, where copies
// complete at issue).
//   step p:  [plane p+1 and operator stage p have landed] barrier
//            issue plane p+2 values, plane p+3 codes, operator stage p+1 ; element p -> force buffers ; barrier
//            flush node plane p (update or partial)
// Ring slot (p+2)%3 was last read by the flush of plane p-1, the code slot (p+3)&3 by that flush too, the
// operator stage (p+1)&1 by element p-1: all before the barrier that opens step p.
#define MARCH_BUNDLE(RUN, RUNNS, WAIT_NODES, WAIT_OPS, A, B, sm, R)                                                           \
  do {                                                                                                                 \
    RUN(mk_issue_ids(A, B, sm, tid, 0); mk_issue_ids(A, B, sm, tid, 1); mk_issue_ids(A, B, sm, tid, 2); mk_commit(); WAIT_NODES); \
    RUN(mk_issue_values(A, B, sm, tid, 0); mk_issue_values(A, B, sm, tid, 1); mk_issue_ops(A, B, sm, tid, 0);           \
        mk_commit(); WAIT_NODES);                                                                                      \
    RUN(mk_phase_begin(A, B, sm, R, tid));                                                                             \
    for (int p = 0; p < (B).Lx; ++p) {                                                                                 \
      RUN(WAIT_NODES; WAIT_OPS(p));                                                                                    \
      RUN(mk_issue_values(A, B, sm, tid, p + 2); mk_issue_ids(A, B, sm, tid, p + 3); mk_issue_ops(A, B, sm, tid, p + 1); \
          mk_commit(); mk_phase_element(A, B, sm, R, tid, p));                                                         \
      RUNNS(mk_phase_flush(A, B, sm, tid, p));                                                                         \
    }                                                                                                                  \
    RUN(WAIT_NODES);                                                                                                   \
    RUN(mk_phase_last(B, sm, R, tid));                                                                                 \
    RUN(mk_phase_flush(A, B, sm, tid, (B).Lx));                                                                        \
  } while (0)

}  // namespace eqd
