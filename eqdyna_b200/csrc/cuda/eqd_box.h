// Operators of an axis-aligned ("box") hexahedron in closed form.
//
// EQdyna's built-in mesh is rectilinear wherever the fault is planar and
// vertical (meshgen.f90:64-107: x = xline(ix), y = yline(iy), z = zline(iz)); the
// nodes of such an element are (xlo|xhi, ylo|yhi, zlo|zhi) in the corner order of
// calcLocalShapeFunc.f90:19-25.  For it the one-point operators that the
// reference precomputes per element (assembleGlobalMass.f90:283-374) collapse to
//   eleshp(d,i) = sign_d(i) * a_d ,  a_d = 1/(4 h_d)        (calcGlobalShapeFunc.f90:19-75)
//   phi(i,m)    = ha(m,i) = +-1                             (assembleGlobalMass.f90:357-372)
//   ss          = diag(ss1, ss4, ss6), off-diagonals 0      (assembleGlobalMass.f90:345-356)
// up to the rounding of the reference's own arithmetic (measured on the shipped
// meshes: <= 4e-14 relative, tests/test_host_and_abi.py::test_box_operators).
// The tile kernel uses these forms for tiles whose elements are all boxes: it
// then streams 15 operator rows per element instead of 71 (a_x, a_y, a_z are
// rows 3, 7, 14 of eleshp = node 2's x-, node 3's y-, node 5's z-derivative).
// Host-callable so that the CPU tests can compare it with the general formulas.
#pragma once
#include <cstddef>

#if defined(__CUDACC__)
#define EQD_HD __host__ __device__ __forceinline__
#else
#define EQD_HD inline
#endif

namespace eqd {

// operator rows a box tile still needs (row numbering of k_tile_reg: RR_*)
enum { BOX_AX = 3, BOX_AY = 7, BOX_AZ = 14, BOX_SS0 = 56, BOX_SS3 = 59, BOX_SS5 = 61, BOX_FIRST_SCALAR = 62, BOX_ROWS = 15 };
EQD_HD constexpr bool box_row(int r) {
  return r == BOX_AX || r == BOX_AY || r == BOX_AZ || r == BOX_SS0 || r == BOX_SS3 || r == BOX_SS5 || r >= BOX_FIRST_SCALAR;
}

// slot of row r in a stage buffer that holds only the box rows
EQD_HD constexpr int box_slot(int r) {
  return r == BOX_AX ? 0 : r == BOX_AY ? 1 : r == BOX_AZ ? 2 : r == BOX_SS0 ? 3 : r == BOX_SS3 ? 4 : r == BOX_SS5 ? 5 : r - BOX_FIRST_SCALAR + 6;
}

// corner signs of calcLocalShapeFunc.f90:19-25: true = +1
EQD_HD constexpr bool box_px(int i) { return ((i ^ (i >> 1)) & 1) != 0; }  // - + + - - + + -
EQD_HD constexpr bool box_py(int i) { return ((i >> 1) & 1) != 0; }        // - - + + - - + +
EQD_HD constexpr bool box_pz(int i) { return ((i >> 2) & 1) != 0; }        // - - - - + + + +
// hourglass base vectors ha(m,i) of assembleGlobalMass.f90:336-339: true = +1
EQD_HD constexpr bool box_hp(int m, int i) {
  return (((m == 0 ? 0xC3 : m == 1 ? 0x69 : m == 2 ? 0x55 : 0x5A) >> i) & 1) != 0;
}
EQD_HD double box_sadd(double a, double v, bool pos) { return pos ? a + v : a - v; }

// g[c][d] = sum_i sign_d(i) * u[i][c]   (component c of the nodal field, direction d)
EQD_HD void box_grad(const double u[8][3], double g[3][3]) {
#pragma unroll
  for (int c = 0; c < 3; ++c) { g[c][0] = 0.0; g[c][1] = 0.0; g[c][2] = 0.0; }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      g[c][0] = box_sadd(g[c][0], u[i][c], box_px(i));
      g[c][1] = box_sadd(g[c][1], u[i][c], box_py(i));
      g[c][2] = box_sadd(g[c][2], u[i][c], box_pz(i));
    }
  }
}

// strain (rate) in the order of calcElemKU.f90:44-60 (engineering shear)
EQD_HD void box_strain(const double g[3][3], double ax, double ay, double az, double s[6]) {
  s[0] = ax * g[0][0];
  s[1] = ay * g[1][1];
  s[2] = az * g[2][2];
  s[3] = az * g[1][2] + ay * g[2][1];
  s[4] = az * g[0][2] + ax * g[2][0];
  s[5] = ay * g[0][1] + ax * g[1][0];
}

// nodal forces B^T t of calcElemKU.f90:175-189
EQD_HD void box_force(const double t[6], double ax, double ay, double az, double f[8][3]) {
  const double x0 = ax * t[0], z4 = az * t[4], y5 = ay * t[5];
  const double y1 = ay * t[1], z3 = az * t[3], x5 = ax * t[5];
  const double z2 = az * t[2], y3 = ay * t[3], x4 = ax * t[4];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const bool px = box_px(i), py = box_py(i), pz = box_pz(i);
    f[i][0] = box_sadd(box_sadd(px ? x0 : -x0, z4, pz), y5, py);
    f[i][1] = box_sadd(box_sadd(py ? y1 : -y1, z3, pz), x5, px);
    f[i][2] = box_sadd(box_sadd(pz ? z2 : -z2, y3, py), x4, px);
  }
}

// Kosloff-Frazier hourglass forces of hrglss.f90:20-54 with phi = ha, ss diagonal;
// l = d + rdampk*v of the 8 nodes, f = the (negative) hourglass resistance
EQD_HD void box_hourglass(const double l[8][3], double ss0, double ss3, double ss5, double f[8][3]) {
  double hv[4][3];
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    double p0 = 0.0, p1 = 0.0, p2 = 0.0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      p0 = box_sadd(p0, l[j][0], box_hp(m, j));
      p1 = box_sadd(p1, l[j][1], box_hp(m, j));
      p2 = box_sadd(p2, l[j][2], box_hp(m, j));
    }
    hv[m][0] = ss0 * p0; hv[m][1] = ss3 * p1; hv[m][2] = ss5 * p2;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    double h0 = 0.0, h1 = 0.0, h2 = 0.0;
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      h0 = box_sadd(h0, hv[m][0], !box_hp(m, i));
      h1 = box_sadd(h1, hv[m][1], !box_hp(m, i));
      h2 = box_sadd(h2, hv[m][2], !box_hp(m, i));
    }
    f[i][0] = h0; f[i][1] = h1; f[i][2] = h2;
  }
}

// exact geometric test on the reference's coordinates: conn = the element's 8 node
// ids (0-based, reference corner order), coor = meshCoor(3,Nn)
inline bool box_element(const int* conn, const double* coor) {
  static const int hiNode[3] = {1, 2, 4};
  for (int d = 0; d < 3; ++d) {
    const double lo = coor[d + 3 * (size_t)conn[0]], hi = coor[d + 3 * (size_t)conn[hiNode[d]]];
    if (!(hi > lo)) return false;
    for (int i = 0; i < 8; ++i) {
      const bool pos = d == 0 ? box_px(i) : d == 1 ? box_py(i) : box_pz(i);
      if (coor[d + 3 * (size_t)conn[i]] != (pos ? hi : lo)) return false;
    }
  }
  return true;
}

}  // namespace eqd
