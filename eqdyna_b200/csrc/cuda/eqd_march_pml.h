// Marching kernel for bundles of axis-aligned PML hexahedra: calcPMLElemKU (assembleGlobalKU.f90:70-346) +
// hrglss.f90:20-54 + the assembly of the twelve split-field force rows of a PML node (assembleGlobalKU.f90:328-344),
// in the same scheme as eqd_march.h (one CTA marches a block of element columns along x; thread t owns node t of
// every plane and the column whose (y-, z-) corner that node is; transformed x- faces travel in registers).
//
// What differs from the regular kernel:
//   * 21 stresses per element: 15 split ones, updated with the damping profile at the centroid (precomputed rows
//     `damps`, :130-213), and 6 regular ones that are read but never written (:320-325);
//   * 12 force rows per node.  The nine split rows take the shape derivative of ONE direction only
//     (force(i) = -det w  eleshp(d,i)  s), and eleshp(d,i) = sign_d(i) a_d: a row of x type is +V on the x+ face and
//     -V on the x- face of the element, the same for all four nodes of a face; a row of y (z) type is sy V (sz V) on
//     both faces.  So one number per row and element column describes the element's share of a node plane, the
//     carry from the previous element is one number too, and the flush applies the signs when it adds the four
//     columns around a node.  Rows 10-12 (regular part + hourglass) are handled as in eqd_march.h;
//   * the nodes are not updated here: a 12-dof node needs its twelve split velocities, which are not staged; every
//     node slot writes one partial of 12 rows, summed by k_node_update12 / k_node_update3 as the tile partials are.
#pragma once
#include "eqd_march.h"

namespace eqd {

constexpr int MP_OPROWS = 12 + 21;   // a_x a_y a_z ss1 ss4 ss6 lam mu det damps(3) | stress(21)
constexpr int MP_COLN = (MK_NZ + 1) * (MK_NY + 1);   // element columns with a border of empty ones all around

struct MarchPmlArgs {
  const MarchBundle* rec;
  const int* ctaFirstA;
  const int* ctaFirstB;
  const int* code;          // [node slots] -1 = no node, else node id (no flags: nothing is updated in place)
  size_t S, NnS, PFS;       // PFS = row stride of pf; slotBase = first partial slot of this class inside pf
  size_t slotBase;
  const double* a;          // [3][S]
  const double* ss;         // [3][S]
  const double* lam; const double* mu; const double* det;
  const double* damps;      // [3][S]
  double* stress;           // [21][S]
  const double* vel; const double* disp;
  double* pf;               // [12][PFS]
  double dt, rdampk, w;
};

struct MarchPmlShared {
  double ops[2][MP_OPROWS][MK_ES];
  double ring[3][6][MK_PN];
  double frc[4][3][MK_PN];          // rows 10-12: one buffer per (dy,dz) corner, as in eqd_march.h
  double col[9][MP_COLN];           // rows 1-9: one number per element column; the border stays zero
  unsigned long long bar[2];
};

struct MarchPmlRegs {
  double wv[3][3];
  double wl[3][3];
  double cf[4][3];   // rows 10-12, transformed x+ face of the previous element
  double cs[9];      // rows 1-9, the previous element's number
  int c0, c1, c2, c3;
};

EQD_HD const double* mp_op_row(const MarchPmlArgs& A, int r) {
  if (r < 3) return A.a + (size_t)r * A.S;
  if (r < 6) return A.ss + (size_t)(r - 3) * A.S;
  if (r == 6) return A.lam;
  if (r == 7) return A.mu;
  if (r == 8) return A.det;
  if (r < 12) return A.damps + (size_t)(r - 9) * A.S;
  return A.stress + (size_t)(r - 12) * A.S;
}
EQD_HD int mp_code(const MarchPmlArgs& A, const MarchBundle& B, int tid, int pl) {
  return pl <= B.Lx ? A.code[(size_t)B.n0 + (size_t)pl * MK_PN + tid] : -1;
}
EQD_HD void mp_issue_values(const MarchPmlArgs& A, MarchPmlShared& sm, int tid, int rs, int code) {
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
EQD_HD void mp_issue_ops(const MarchPmlArgs& A, const MarchBundle& B, MarchPmlShared& sm, int tid, int p) {
  if (p >= B.Lx) return;
  unsigned long long* bar = &sm.bar[p & 1];
  const int es = mk_es(B);
  if (tid == 0) mk_bar_expect(bar, (unsigned)(MP_OPROWS * es * sizeof(double)));
#if defined(__CUDA_ARCH__)
  __syncwarp();
#endif
  if (tid < MP_OPROWS)
    mk_bulk(sm.ops[p & 1][tid], mp_op_row(A, tid) + (size_t)B.e0 + (size_t)p * es, (unsigned)(es * sizeof(double)), bar);
}
// this thread's column in the bordered column array (ny = nodes per row of the bundle's node plane: 16 or 8)
EQD_HD int mp_col(int tid, int ny) { return (tid / ny + 1) * (ny + 1) + (tid % ny) + 1; }
EQD_HD bool mp_active(const MarchBundle& B, int tid, int ny) { return (tid / ny) < mk_bz(B) && (tid % ny) < mk_by(B); }

EQD_HD void mp_phase_begin(const MarchPmlArgs& A, const MarchBundle& B, MarchPmlShared& sm, MarchPmlRegs& R, int tid) {
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int c = 0; c < 3; ++c) sm.frc[q][c][tid] = 0.0;
  for (int i = tid; i < MP_COLN; i += MK_NT)
#pragma unroll
    for (int r = 0; r < 9; ++r) sm.col[r][i] = 0.0;
  const int ny = mk_rowlen(B);
  if (!mp_active(B, tid, ny)) return;
  double(*pl)[MK_PN] = sm.ring[0];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const double v0 = pl[c][tid], v1 = pl[c][tid + 1], v2 = pl[c][tid + ny], v3 = pl[c][tid + ny + 1];
    {
      MK_FACE_SUMS(v0, v1, v2, v3)
      R.wv[0][c] = s0_ + s1_; R.wv[1][c] = d0_ + d1_; R.wv[2][c] = s1_ - s0_;
    }
    const double m0 = pl[3 + c][tid] + A.rdampk * v0, m1 = pl[3 + c][tid + 1] + A.rdampk * v1;
    const double m2 = pl[3 + c][tid + ny] + A.rdampk * v2, m3 = pl[3 + c][tid + ny + 1] + A.rdampk * v3;
    {
      MK_FACE_SUMS(m0, m1, m2, m3)
      R.wl[0][c] = d0_ + d1_; R.wl[1][c] = s1_ - s0_; R.wl[2][c] = d1_ - d0_;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) R.cf[k][c] = 0.0;
  }
#pragma unroll
  for (int r = 0; r < 9; ++r) R.cs[r] = 0.0;
}

EQD_HD void mp_store_face(MarchPmlShared& sm, int c, int tid, int ny, double G0, double G1, double G2, double G3) {
  const double um = G0 - G1, up = G0 + G1, wm = G2 - G3, wp = G2 + G3;
  sm.frc[0][c][tid] = um - wm;
  sm.frc[1][c][tid + 1] = up - wp;
  sm.frc[2][c][tid + ny] = um + wm;
  sm.frc[3][c][tid + ny + 1] = up + wp;
}

// direction of the shape derivative in split row r (0-based dof): assembleGlobalKU.f90:328-344
//   dofs 1,5,9 = (x,sxx) (y,syy) (z,szz); dofs 2,3 = (y,sxy) (z,sxz); 4,6 = (x,sxy) (z,syz); 7,8 = (x,sxz) (y,syz)
EQD_HD constexpr int mp_dir(int r) { return r == 0 || r == 3 || r == 6 ? 0 : (r == 1 || r == 4 || r == 7 ? 1 : 2); }

EQD_HD void mp_phase_element(const MarchPmlArgs& A, const MarchBundle& B, MarchPmlShared& sm, MarchPmlRegs& R, int tid, int p, int rs1) {
  const int ny = mk_rowlen(B);
  if (!mp_active(B, tid, ny)) return;
  double(*pl)[MK_PN] = sm.ring[rs1];
  const double(*op)[MK_ES] = sm.ops[p & 1];
  double gx[3], gy[3], gz[3], P[4][3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const double v0 = pl[c][tid], v1 = pl[c][tid + 1], v2 = pl[c][tid + ny], v3 = pl[c][tid + ny + 1];
    {
      MK_FACE_SUMS(v0, v1, v2, v3)
      const double n0 = s0_ + s1_, vy = d0_ + d1_, vz = s1_ - s0_;
      gx[c] = n0 - R.wv[0][c]; gy[c] = vy + R.wv[1][c]; gz[c] = vz + R.wv[2][c];
      R.wv[0][c] = n0; R.wv[1][c] = vy; R.wv[2][c] = vz;
    }
    const double m0 = pl[3 + c][tid] + A.rdampk * v0, m1 = pl[3 + c][tid + 1] + A.rdampk * v1;
    const double m2 = pl[3 + c][tid + ny] + A.rdampk * v2, m3 = pl[3 + c][tid + ny + 1] + A.rdampk * v3;
    {
      MK_FACE_SUMS(m0, m1, m2, m3)
      const double ly = d0_ + d1_, lz = s1_ - s0_, lyz = d1_ - d0_;
      P[0][c] = lyz + R.wl[2][c]; P[1][c] = lz - R.wl[1][c]; P[2][c] = ly - R.wl[0][c]; P[3][c] = lyz - R.wl[2][c];
      R.wl[0][c] = ly; R.wl[1][c] = lz; R.wl[2][c] = lyz;
    }
  }
  const int es = (tid / ny) * mk_by(B) + (tid % ny);
  const double ax = op[0][es], ay = op[1][es], az = op[2][es];
  const double lam = op[6][es], mu = op[7][es], l2m = lam + 2.0 * mu;
  const double detw = op[8][es] * A.w;
  const double rdt = 1 / A.dt;
  // :277-311  s <- (coef*D + (1/dt - d/2) s) / (1/dt + d/2); the three denominators once per element, as reciprocals
  const double dh[3] = {op[9][es] / 2, op[10][es] / 2, op[11][es] / 2};
  const double rden[3] = {1.0 / (rdt + dh[0]), 1.0 / (rdt + dh[1]), 1.0 / (rdt + dh[2])};
  double* sp = A.stress + (size_t)B.e0 + (size_t)p * mk_es(B) + es;
#define MP_UPD(k, cf, D, a, out)                                          \
  {                                                                       \
    const double x_ = ((cf) * (D) + (rdt - dh[a]) * op[12 + (k)][es]) * rden[a]; \
    out = x_;                                                             \
    sp[(size_t)(k) * A.S] = x_;                                           \
  }
  // velocity gradients (assembleGlobalKU.f90:248-275) with eleshp(d,i) = sign_d(i) a_d
  const double g00 = ax * gx[0], g11 = ay * gy[1], g22 = az * gz[2];
  const double g01 = ax * gx[1], g10 = ay * gy[0], g02 = ax * gx[2], g20 = az * gz[0], g12 = ay * gy[2], g21 = az * gz[1];
  double s[15];
  MP_UPD(0, l2m, g00, 0, s[0]) MP_UPD(1, lam, g11, 1, s[1]) MP_UPD(2, lam, g22, 2, s[2])
  MP_UPD(3, lam, g00, 0, s[3]) MP_UPD(4, l2m, g11, 1, s[4]) MP_UPD(5, lam, g22, 2, s[5])
  MP_UPD(6, lam, g00, 0, s[6]) MP_UPD(7, lam, g11, 1, s[7]) MP_UPD(8, l2m, g22, 2, s[8])
  MP_UPD(9, mu, g01, 0, s[9]) MP_UPD(10, mu, g10, 1, s[10])
  MP_UPD(11, mu, g02, 0, s[11]) MP_UPD(12, mu, g20, 2, s[12])
  MP_UPD(13, mu, g12, 1, s[13]) MP_UPD(14, mu, g21, 2, s[14])
#undef MP_UPD
  const double sxx = s[0] + s[1] + s[2], syy = s[3] + s[4] + s[5], szz = s[6] + s[7] + s[8];
  const double sxy = s[9] + s[10], sxz = s[11] + s[12], syz = s[13] + s[14];
  // rows 1-9 (:328-341): the element's number per row, V = 0 - det w a_d s
  const double V[9] = {0.0 - detw * ax * sxx, 0.0 - detw * ay * sxy, 0.0 - detw * az * sxz,
                       0.0 - detw * ax * sxy, 0.0 - detw * ay * syy, 0.0 - detw * az * syz,
                       0.0 - detw * ax * sxz, 0.0 - detw * ay * syz, 0.0 - detw * az * szz};
  const int cb = mp_col(tid, ny);
#pragma unroll
  for (int r = 0; r < 9; ++r) {
    // x type: +V on the x+ face, -V on the x- face; y / z type: the same on both faces
    sm.col[r][cb] = mp_dir(r) == 0 ? R.cs[r] - V[r] : R.cs[r] + V[r];
    R.cs[r] = V[r];
  }
  // rows 10-12: regular part s(16:21) + rdampk * rate (read only, :215-246,320-325) and hourglass
  double sr[6];
  sr[0] = g00;
  sr[1] = g11;
  sr[2] = g22;
  sr[3] = az * gz[1] + ay * gy[2];
  sr[4] = az * gz[0] + ax * gx[2];
  sr[5] = ay * gy[0] + ax * gx[1];
  double rate[6];
  rate[0] = 0.0 + l2m * sr[0] + lam * sr[1] + lam * sr[2];
  rate[1] = 0.0 + lam * sr[0] + l2m * sr[1] + lam * sr[2];
  rate[2] = 0.0 + lam * sr[0] + lam * sr[1] + l2m * sr[2];
  rate[3] = mu * sr[3];
  rate[4] = mu * sr[4];
  rate[5] = mu * sr[5];
  double t[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) t[k] = (0.0 - detw) * (op[12 + 15 + k][es] + A.rdampk * rate[k]);
  const double ssd[3] = {op[3][es], op[4][es], op[5][es]};
  const double X[3] = {ax * t[0], ax * t[5], ax * t[4]};
  const double Y[3] = {ay * t[5], ay * t[1], ay * t[3]};
  const double Z[3] = {az * t[4], az * t[3], az * t[2]};
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const double H0 = ssd[c] * P[0][c], H1 = ssd[c] * P[1][c], H2 = ssd[c] * P[2][c], H3 = ssd[c] * P[3][c];
    const double G0 = R.cf[0][c] - X[c], G1 = R.cf[1][c] + (Y[c] + H2), G2 = R.cf[2][c] + (Z[c] + H1), G3 = R.cf[3][c] + (H3 - H0);
    mp_store_face(sm, c, tid, ny, G0, G1, G2, G3);
    R.cf[0][c] = X[c]; R.cf[1][c] = Y[c] - H2; R.cf[2][c] = Z[c] - H1; R.cf[3][c] = 0.0 - H0 - H3;
  }
}

EQD_HD void mp_phase_last(const MarchBundle& B, MarchPmlShared& sm, MarchPmlRegs& R, int tid) {
  const int ny = mk_rowlen(B);
  if (!mp_active(B, tid, ny)) return;
  const int cb = mp_col(tid, ny);
#pragma unroll
  for (int r = 0; r < 9; ++r) sm.col[r][cb] = R.cs[r];   // the x+ face of the last element: +V (x type) or V (y, z type)
#pragma unroll
  for (int c = 0; c < 3; ++c) mp_store_face(sm, c, tid, ny, R.cf[0][c], R.cf[1][c], R.cf[2][c], R.cf[3][c]);
}

EQD_HD void mp_phase_flush(const MarchPmlArgs& A, const MarchBundle& B, MarchPmlShared& sm, int tid, int pl, int code) {
  if (code < 0) return;
  const size_t slot = A.slotBase + (size_t)B.n0 + (size_t)pl * MK_PN + tid;
  // the four columns around node (iz, iy): c00 has it at its (y-, z-) corner, c01 at (y+, z-), c10 at (y-, z+), c11 at (y+, z+)
  const int ny = mk_rowlen(B);
  const int iz = tid / ny, iy = tid % ny;
  const int c11 = iz * (ny + 1) + iy, c10 = c11 + 1, c01 = c11 + (ny + 1), c00 = c01 + 1;
#pragma unroll
  for (int r = 0; r < 9; ++r) {
    const double g00 = sm.col[r][c00], g01 = sm.col[r][c01], g10 = sm.col[r][c10], g11 = sm.col[r][c11];
    double F;
    if (mp_dir(r) == 0) F = ((g00 + g01) + g10) + g11;
    else if (mp_dir(r) == 1) F = ((g01 - g00) - g10) + g11;
    else F = ((g10 - g00) - g01) + g11;
    A.pf[(size_t)r * A.PFS + slot] = F;
  }
#pragma unroll
  for (int c = 0; c < 3; ++c)
    A.pf[(size_t)(9 + c) * A.PFS + slot] = ((sm.frc[0][c][tid] + sm.frc[1][c][tid]) + sm.frc[2][c][tid]) + sm.frc[3][c][tid];
}

// same schedule as MARCH_BUNDLE (eqd_march.h)
#define MARCH_PML_BUNDLE(RUN, RUNNS, WAIT_NODES, WAIT_OPS, A, B, sm, R)                                                 \
  do {                                                                                                                 \
    RUN(R.c0 = mp_code(A, B, tid, 0); R.c1 = mp_code(A, B, tid, 1); R.c2 = mp_code(A, B, tid, 2); R.c3 = mp_code(A, B, tid, 3); \
        mp_issue_values(A, sm, tid, 0, R.c0); mp_issue_values(A, sm, tid, 1, R.c1); mp_issue_ops(A, B, sm, tid, 0);     \
        mk_commit(); WAIT_NODES);                                                                                      \
    RUN(mp_phase_begin(A, B, sm, R, tid));                                                                             \
    for (int p = 0, rs0 = 0, rs1 = 1, rs2 = 2; p < (B).Lx; ++p) {                                                      \
      RUN(WAIT_NODES; WAIT_OPS(p));                                                                                    \
      RUN(mp_issue_values(A, sm, tid, rs2, R.c2); mp_issue_ops(A, B, sm, tid, p + 1); mk_commit();                     \
          mp_phase_element(A, B, sm, R, tid, p, rs1));                                                                 \
      RUNNS(mp_phase_flush(A, B, sm, tid, p, R.c0);                                                                    \
            R.c0 = R.c1; R.c1 = R.c2; R.c2 = R.c3; R.c3 = mp_code(A, B, tid, p + 4));                                  \
      { const int t_ = rs0; rs0 = rs1; rs1 = rs2; rs2 = t_; }                                                          \
    }                                                                                                                  \
    RUN(WAIT_NODES);                                                                                                   \
    RUN(mp_phase_last(B, sm, R, tid));                                                                                 \
    RUN(mp_phase_flush(A, B, sm, tid, (B).Lx, R.c0));                                                                  \
  } while (0)

}  // namespace eqd
