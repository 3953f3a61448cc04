// PROTOTYPE (round-2 preparation; not part of the product, nothing under eqdyna_b200/ includes it).
//
// The marching kernel for bundles of axis-aligned hexahedra, written ONCE in "phase style":
// the body is a sequence of phases separated by barriers, each phase a function of (thread id, that thread's
// private registers).  Compiled by nvcc the phases run on the CTA's threads with __syncthreads() between
// them (k_march_reg below); compiled by g++ (march_emul.cpp) a driver loop runs every phase over all thread
// ids in turn, which is equivalent as long as no phase has a race -- and lets the CPU test suite check the
// indexing, the carry logic and the arithmetic of the very source the GPU will execute.
//
// Bundle = BZ x BY element columns, Lx <= MK_LX elements long in x.  Elements of a bundle are stored
// plane-major in the class SoA (slot = e0 + p*BZ*BY + cz*BY + cy), its nodes plane by plane in (z, y) order
// (slot = n0 + p*PN + iz*(BY+1) + iy; the ids may be split-node masters: nothing assumes ascending order).
// Thread t: role = t / (BZ*BY) (0 = stress part, calcElemKU.f90:44-189; 1 = hourglass part, hrglss.f90:20-54),
// column (cz, cy) = the rest.  Elastic, C_hg = 1, no body force.
#pragma once
#include <cstddef>
#include <cstdint>

#include "../../eqdyna_b200/csrc/cuda/eqd_box.h"

namespace march {

using namespace eqd;

constexpr int BZ = 4, BY = 16, NCOL = BZ * BY, NT = 2 * NCOL, PN = (BZ + 1) * (BY + 1);
constexpr int MK_LX = 64;

struct Bundle { int e0, n0, Lx, pad; };   // first element slot, first node slot, length

struct Args {
  int nBundles;
  const Bundle* rec;
  const int* tnode;                 // [node slots] plane-ordered node ids of every bundle
  size_t S, NnS, PFS;               // row strides: class SoA, nodal SoA, partial-force rows
  const double* ax; const double* ay; const double* az;     // a_d = eleshp rows 3, 7, 14
  const double* ss0; const double* ss3; const double* ss5;  // diagonal of ss
  const double* lam; const double* mu; const double* det;
  double* stress;                   // [6][S]
  const double* vel; const double* disp;   // [3][NnS]
  double* pf;                       // [3][PFS] one partial force per (bundle, node)
  double dt, rdampk, w;
};

// what a thread keeps in registers from one step to the next
struct Regs {
  double um[4][3];    // values at the element's x- corners (0,3,4,7): v (role 0) or d + rdampk v (role 1)
  double fc[4][3];    // forces on the x+ corners of the previous element, waiting for this element's share
  double fx[4][3];    // completed x- corner forces of this step, consumed by the four assembly phases
  double pre[2][3];   // prefetch variant: v and d of the one node of plane p+2 this thread fetches (PN <= NT)
  double op[12];      // operator-prefetch variant: the next element's a_d, lam, mu, det, stress (role 0) or ss (role 1)
};

// shared memory of one CTA
struct Shared {
  double val[3][2][3][PN];   // [plane mod 3][v | d + rdampk v][component][node]: a ring of node planes
  double frc[2][3][PN];      // [role][component][node]: force plane being assembled
};

// x- corners 0,3,4,7 and their x+ partners 1,2,5,6 sit at (y, z) offsets:
EQD_HD constexpr int xm(int q) { return q == 0 ? 0 : q == 1 ? 3 : q == 2 ? 4 : 7; }
EQD_HD constexpr int xp(int q) { return q == 0 ? 1 : q == 1 ? 2 : q == 2 ? 5 : 6; }
EQD_HD constexpr int dy(int q) { return q & 1; }
EQD_HD constexpr int dz(int q) { return q >> 1; }

// ---- phase: all threads load node plane `p` of the bundle into val[p % 3] (v and d + rdampk v)
EQD_HD void phase_load_plane(const Args& A, const Bundle& B, Shared& sm, int tid, int p) {
  for (int i = tid; i < PN; i += NT) {
    const int n = A.tnode[(size_t)B.n0 + (size_t)p * PN + i];
    for (int c = 0; c < 3; ++c) {
      const double v = A.vel[c * A.NnS + n], d = A.disp[c * A.NnS + n];
      sm.val[p % 3][0][c][i] = v;
      sm.val[p % 3][1][c][i] = d + A.rdampk * v;     // hrglss.f90:20-27
    }
  }
}

// ---- phase: zero the force plane, take the x- corner values of the first element from plane 0
EQD_HD void phase_begin(Shared& sm, Regs& R, int tid) {
  for (int i = tid; i < PN; i += NT)
    for (int r = 0; r < 2; ++r)
      for (int c = 0; c < 3; ++c) sm.frc[r][c][i] = 0.0;
  const int role = tid / NCOL, col = tid - role * NCOL, cz = col / BY, cy = col - cz * BY;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int li = (cz + dz(q)) * (BY + 1) + cy + dy(q);
#pragma unroll
    for (int c = 0; c < 3; ++c) { R.um[q][c] = sm.val[0][role][c][li]; R.fc[q][c] = 0.0; }
  }
}

// ---- phase: element (step p, this thread's column): forces from the carried x- values and plane p+1
// operator value k of element slot e: role 0: a_x, a_y, a_z, lam, mu, det, stress 1..6; role 1: ss1, ss4, ss6
EQD_HD double op_value(const Args& A, size_t e, int role, int k) {
  if (role == 0) {
    switch (k) {
      case 0: return A.ax[e];
      case 1: return A.ay[e];
      case 2: return A.az[e];
      case 3: return A.lam[e];
      case 4: return A.mu[e];
      case 5: return A.det[e];
      default: return A.stress[(size_t)(k - 6) * A.S + e];
    }
  }
  return k == 0 ? A.ss0[e] : k == 1 ? A.ss3[e] : A.ss5[e];
}
EQD_HD void fetch_ops(const Args& A, const Bundle& B, int tid, int p, double op[12]) {
  const int role = tid / NCOL, col = tid - role * NCOL;
  const size_t e = (size_t)B.e0 + (size_t)p * NCOL + col;
#pragma unroll
  for (int k = 0; k < 12; ++k)
    if (role == 0 || k < 3) op[k] = op_value(A, e, role, k);
}
// operator-prefetch variant: issue the loads of step p (p < Lx) into R.op
EQD_HD void phase_ops_issue(const Args& A, const Bundle& B, Regs& R, int tid, int p) {
  if (p < B.Lx) fetch_ops(A, B, tid, p, R.op);
}

template <bool OPS_IN_REGS>
EQD_HD void phase_element_t(const Args& A, const Bundle& B, Shared& sm, Regs& R, int tid, int p) {
  const int role = tid / NCOL, col = tid - role * NCOL, cz = col / BY, cy = col - cz * BY;
  const size_t e = (size_t)B.e0 + (size_t)p * NCOL + col;
  double u[8][3], f[8][3], up[4][3];
  // fetched during the previous step (operator-prefetch schedule) or read where it is needed
#define MK_OP(k) (OPS_IN_REGS ? R.op[k] : op_value(A, e, role, k))
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int li = (cz + dz(q)) * (BY + 1) + cy + dy(q);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      up[q][c] = sm.val[(p + 1) % 3][role][c][li];
      u[xm(q)][c] = R.um[q][c];
      u[xp(q)][c] = up[q][c];
    }
  }
  if (role == 0) {
    const double ax = MK_OP(0), ay = MK_OP(1), az = MK_OP(2);
    double g[3][3], sr[6], t[6];
    box_grad(u, g);
    box_strain(g, ax, ay, az, sr);
    const double lam = MK_OP(3), mu = MK_OP(4), l2m = lam + 2 * mu;
    // calcElemKU.f90:63-76
    const double rate[6] = {0.0 + l2m * sr[0] + lam * sr[1] + lam * sr[2], 0.0 + lam * sr[0] + l2m * sr[1] + lam * sr[2],
                            0.0 + lam * sr[0] + lam * sr[1] + l2m * sr[2], mu * sr[3], mu * sr[4], mu * sr[5]};
    const double temp = (-MK_OP(5)) * A.w;     // calcElemKU.f90:169-173, constk = -eledet
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      const double sg = MK_OP(6 + k) + rate[k] * A.dt;
      A.stress[k * A.S + e] = sg;
      t[k] = temp * (sg + A.rdampk * rate[k]);
    }
    box_force(t, ax, ay, az, f);
  } else {
    box_hourglass(u, MK_OP(0), MK_OP(1), MK_OP(2), f);
  }
#undef MK_OP
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      R.fx[q][c] = R.fc[q][c] + f[xm(q)][c];   // the previous element's x+ share + this element's x- share
      R.fc[q][c] = f[xp(q)][c];
      R.um[q][c] = up[q][c];
    }
}

EQD_HD void phase_element(const Args& A, const Bundle& B, Shared& sm, Regs& R, int tid, int p) {
  phase_element_t<false>(A, B, sm, R, tid, p);
}

// ---- prefetch variant: the loads of node plane `p` are issued into registers (before the element work of
// the step, so that their latency hides behind it) ...
EQD_HD void phase_prefetch_issue(const Args& A, const Bundle& B, Regs& R, int tid, int p) {
  static_assert(PN <= NT, "one node of a plane per thread");
  if (tid < PN && p <= B.Lx) {
    const int n = A.tnode[(size_t)B.n0 + (size_t)p * PN + tid];
#pragma unroll
    for (int c = 0; c < 3; ++c) { R.pre[0][c] = A.vel[c * A.NnS + n]; R.pre[1][c] = A.disp[c * A.NnS + n]; }
  }
}
// ... and committed to the ring slot of plane p once the step's assembly is over (the slot's previous plane,
// p - 3, was last read two steps ago)
EQD_HD void phase_prefetch_commit(const Args& A, const Bundle& B, Shared& sm, const Regs& R, int tid, int p) {
  if (tid < PN && p <= B.Lx) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      sm.val[p % 3][0][c][tid] = R.pre[0][c];
      sm.val[p % 3][1][c][tid] = R.pre[1][c] + A.rdampk * R.pre[0][c];
    }
  }
}

// ---- phases q = 0..3: every column adds its completed forces to its (dy(q), dz(q)) node of the plane;
// within one phase no two columns of a role touch the same node, and the roles own separate rows
EQD_HD void phase_assemble(Shared& sm, const Regs& R, int tid, int q, bool last) {
  const int role = tid / NCOL, col = tid - role * NCOL, cz = col / BY, cy = col - cz * BY;
  const int li = (cz + dz(q)) * (BY + 1) + cy + dy(q);
#pragma unroll
  for (int c = 0; c < 3; ++c) sm.frc[role][c][li] += last ? R.fc[q][c] : R.fx[q][c];
}

// ---- phase: force plane p is complete: one partial per bundle node, both roles merged; clear it
EQD_HD void phase_flush(const Args& A, const Bundle& B, Shared& sm, int tid, int p) {
  for (int i = tid; i < PN; i += NT)
    for (int c = 0; c < 3; ++c) {
      A.pf[c * A.PFS + (size_t)B.n0 + (size_t)p * PN + i] = sm.frc[0][c][i] + sm.frc[1][c][i];
      sm.frc[0][c][i] = 0.0; sm.frc[1][c][i] = 0.0;
    }
}

// The schedule, as a list of (phase, barrier) pairs.  RUN(body) executes `body` for every thread of the CTA
// and then synchronises: one statement + __syncthreads() on the device, a loop over tid on the host.
#define MARCH_BUNDLE(RUN, A, B, sm, REGS)                                                 \
  do {                                                                                    \
    RUN(phase_load_plane(A, B, sm, tid, 0));                                              \
    RUN(phase_begin(sm, REGS, tid));                                                      \
    for (int p = 0; p < (B).Lx; ++p) {                                                    \
      RUN(phase_load_plane(A, B, sm, tid, p + 1));                                        \
      RUN(phase_element(A, B, sm, REGS, tid, p));                                         \
      for (int q = 0; q < 4; ++q) RUN(phase_assemble(sm, REGS, tid, q, false));           \
      RUN(phase_flush(A, B, sm, tid, p));                                                 \
    }                                                                                     \
    for (int q = 0; q < 4; ++q) RUN(phase_assemble(sm, REGS, tid, q, true));              \
    RUN(phase_flush(A, B, sm, tid, (B).Lx));                                              \
  } while (0)

// Prefetch variant of the schedule: one barrier less per step and no exposed load of the next node plane.
#define MARCH_BUNDLE_PF(RUN, A, B, sm, REGS)                                                                   \
  do {                                                                                                        \
    RUN(phase_load_plane(A, B, sm, tid, 0); phase_load_plane(A, B, sm, tid, 1));                              \
    RUN(phase_begin(sm, REGS, tid));                                                                          \
    for (int p = 0; p < (B).Lx; ++p) {                                                                        \
      RUN(phase_prefetch_issue(A, B, REGS, tid, p + 2); phase_element(A, B, sm, REGS, tid, p));               \
      for (int q = 0; q < 4; ++q) RUN(phase_assemble(sm, REGS, tid, q, false));                               \
      RUN(phase_flush(A, B, sm, tid, p); phase_prefetch_commit(A, B, sm, REGS, tid, p + 2));                  \
    }                                                                                                         \
    for (int q = 0; q < 4; ++q) RUN(phase_assemble(sm, REGS, tid, q, true));                                  \
    RUN(phase_flush(A, B, sm, tid, (B).Lx));                                                                  \
  } while (0)

// Third schedule: additionally the next element's operator values travel through registers, so that no global
// load is waited for inside a step (12 more doubles of state per thread: wants 3 CTAs per SM instead of 4).
#define MARCH_BUNDLE_PF2(RUN, A, B, sm, REGS)                                                                  \
  do {                                                                                                        \
    RUN(phase_load_plane(A, B, sm, tid, 0); phase_load_plane(A, B, sm, tid, 1); phase_ops_issue(A, B, REGS, tid, 0)); \
    RUN(phase_begin(sm, REGS, tid));                                                                          \
    for (int p = 0; p < (B).Lx; ++p) {                                                                        \
      RUN(phase_prefetch_issue(A, B, REGS, tid, p + 2); phase_element_t<true>(A, B, sm, REGS, tid, p);        \
          phase_ops_issue(A, B, REGS, tid, p + 1));                                                           \
      for (int q = 0; q < 4; ++q) RUN(phase_assemble(sm, REGS, tid, q, false));                               \
      RUN(phase_flush(A, B, sm, tid, p); phase_prefetch_commit(A, B, sm, REGS, tid, p + 2));                  \
    }                                                                                                         \
    for (int q = 0; q < 4; ++q) RUN(phase_assemble(sm, REGS, tid, q, true));                                  \
    RUN(phase_flush(A, B, sm, tid, (B).Lx));                                                                  \
  } while (0)

#ifdef __CUDACC__
// One CTA per bundle at a time, persistent over the bundle list.  (No software pipelining yet: the next
// plane is loaded synchronously; with < 8 KB of shared memory per CTA occupancy is set by registers.)
__global__ void __launch_bounds__(NT, 4) k_march_reg(Args A) {
  __shared__ Shared sm;
  Regs regs;
  const int tid = threadIdx.x;
  for (int b = blockIdx.x; b < A.nBundles; b += gridDim.x) {
    const Bundle B = A.rec[b];
#define MK_RUN(body) do { body; __syncthreads(); } while (0)
    MARCH_BUNDLE(MK_RUN, A, B, sm, regs);
#undef MK_RUN
  }
}
__global__ void __launch_bounds__(NT, 4) k_march_reg_pf(Args A) {
  __shared__ Shared sm;
  Regs regs;
  const int tid = threadIdx.x;
  for (int b = blockIdx.x; b < A.nBundles; b += gridDim.x) {
    const Bundle B = A.rec[b];
#define MK_RUN(body) do { body; __syncthreads(); } while (0)
    MARCH_BUNDLE_PF(MK_RUN, A, B, sm, regs);
#undef MK_RUN
  }
}
__global__ void __launch_bounds__(NT, 3) k_march_reg_pf2(Args A) {
  __shared__ Shared sm;
  Regs regs;
  const int tid = threadIdx.x;
  for (int b = blockIdx.x; b < A.nBundles; b += gridDim.x) {
    const Bundle B = A.rec[b];
#define MK_RUN(body) do { body; __syncthreads(); } while (0)
    MARCH_BUNDLE_PF2(MK_RUN, A, B, sm, regs);
#undef MK_RUN
  }
}
#endif

}  // namespace march
