// One-time operator precompute on the device:
// Jacobian determinant, one-point shape-function derivatives, Kosloff-Frazier
// hourglass operators and the lumped mass, from the mesh eqd_set_mesh already
// holds.  Replaces the host's assembleGlobalMass (src/assembleGlobalMass.f90:3-56,
// 283-406), calcGlobalShapeFunc (src/calcGlobalShapeFunc.f90:19-75),
// calcLocalShapeFunc (src/calcLocalShapeFunc.f90:19-25) and vlm
// (src/library.f90:60-93) plus the upload of their ~600 B per element.
//
// This file is compiled with --fmad=false: the reference's build has no FMA
// contraction, and with the same operation order every per-element operator
// (det, shp, ss, phi, element mass) is bit-identical to the Fortran / host values.
// Only the nodal lumped mass differs in its last bits: it is summed tile by tile
// in a fixed order instead of in ascending element order.
#include <cuda_runtime.h>

#include "eqd_dev.cuh"
#include "eqd_kernels.h"

namespace eqd {

namespace {

// calcLocalShapeFunc.f90:19-25: N_i,xi = acoor/8
__constant__ double c_acoor[8][3] = {{-1, -1, -1}, {1, -1, -1}, {1, 1, -1}, {-1, 1, -1},
                                     {-1, -1, 1},  {1, -1, 1},  {1, 1, 1},  {-1, 1, 1}};
// assembleGlobalMass.f90:336-339 (hourglass base vectors)
__constant__ int c_ha[4][8] = {{1, 1, -1, -1, -1, -1, 1, 1},
                               {1, -1, -1, 1, -1, 1, 1, -1},
                               {1, -1, 1, -1, 1, -1, 1, -1},
                               {-1, 1, -1, 1, 1, -1, 1, -1}};
// library.f90:66-73, it(a,i) stored as [i-1][a-1], 0-based node ids
__constant__ int c_it[8][8] = {{0, 1, 2, 3, 4, 5, 6, 7}, {1, 2, 3, 0, 5, 6, 7, 4}, {2, 3, 0, 1, 6, 7, 4, 5},
                               {3, 0, 1, 2, 7, 4, 5, 6}, {4, 7, 6, 5, 0, 3, 2, 1}, {5, 4, 7, 6, 1, 0, 3, 2},
                               {6, 5, 4, 7, 2, 1, 0, 3}, {7, 6, 5, 4, 3, 2, 1, 0}};

// vlm, library.f90:60-93 (Belytschko et al. 1984)
__device__ double vlm(const double xl[8][3]) {
  double volume = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
#define Y(a) xl[c_it[i][(a)-1]][1]
#define Z(a) xl[c_it[i][(a)-1]][2]
    const double bb = Y(2) * (Z(6) - Z(3) + Z(5) - Z(4)) + Y(3) * (Z(2) - Z(4)) + Y(4) * (Z(3) - Z(8) + Z(2) - Z(5)) +
                      Y(5) * (Z(8) - Z(6) + Z(4) - Z(2)) + Y(6) * (Z(5) - Z(2)) + Y(8) * (Z(4) - Z(5));
#undef Y
#undef Z
    volume = volume + xl[i][0] * bb;
  }
  return volume / 12.0;
}

}  // namespace

// one thread per class slot
__global__ void __launch_bounds__(128) k_elem_ops(OpsArgs A) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= A.S) return;
  const int e = A.refId[s];
  if (e < 0) return;
  const size_t S = A.S;
  double xl[8][3];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int nd = A.conn[8 * (size_t)e + i];
#pragma unroll
    for (int j = 0; j < 3; ++j) xl[i][j] = A.coor[j + 3 * (size_t)nd];
  }
  // calcGlobalShapeFunc.f90:19-35: local derivatives, wedge degeneration (nodes 3=4, 7=8)
  const double cst = 1.0 / 8.0;
  double shg[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    shg[i][3] = cst;
#pragma unroll
    for (int j = 0; j < 3; ++j) shg[i][j] = cst * c_acoor[i][j];
  }
  const int et = A.etype[e];
  if (et == 11 || et == 12) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      shg[2][j] = shg[2][j] + shg[3][j];
      shg[3][j] = 0.0;
      shg[6][j] = shg[6][j] + shg[7][j];
      shg[7][j] = 0.0;
    }
  }
  double xs[3][3];  // xs(j,i) -> xs[j-1][i-1], calcGlobalShapeFunc.f90:37-45
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      double temp = 0.0;
#pragma unroll
      for (int k = 0; k < 8; ++k) temp = temp + shg[k][i] * xl[k][j];
      xs[j][i] = temp;
    }
  const double cof11 = xs[1][1] * xs[2][2] - xs[1][2] * xs[2][1];
  const double cof12 = xs[1][2] * xs[2][0] - xs[1][0] * xs[2][2];
  const double cof13 = xs[1][0] * xs[2][1] - xs[1][1] * xs[2][0];
  const double cof21 = xs[2][1] * xs[0][2] - xs[2][2] * xs[0][1];
  const double cof22 = xs[2][2] * xs[0][0] - xs[2][0] * xs[0][2];
  const double cof23 = xs[2][0] * xs[0][1] - xs[2][1] * xs[0][0];
  const double cof31 = xs[0][1] * xs[1][2] - xs[0][2] * xs[1][1];
  const double cof32 = xs[0][2] * xs[1][0] - xs[0][0] * xs[1][2];
  const double cof33 = xs[0][0] * xs[1][1] - xs[0][1] * xs[1][0];
  const double det = xs[0][0] * cof11 + xs[0][1] * cof12 + xs[0][2] * cof13;
  if (!(det > 0.0)) { atomicMin(A.badElem, e); return; }   // calcGlobalShapeFunc.f90:57-61
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const double t0 = shg[i][0], t1 = shg[i][1], t2 = shg[i][2];
    shg[i][0] = (t0 * cof11 + t1 * cof12 + t2 * cof13) / det;
    shg[i][1] = (t0 * cof21 + t1 * cof22 + t2 * cof23) / det;
    shg[i][2] = (t0 * cof31 + t1 * cof32 + t2 * cof33) / det;
  }
  A.det[s] = det;
  if (A.compact) {
    // eleshp(1,2), eleshp(2,3), eleshp(3,5): the three numbers a box element's 24 derivatives consist of (eqd_box.h)
    A.shp[s] = shg[1][0]; A.shp[S + s] = shg[2][1]; A.shp[2 * S + s] = shg[4][2];
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) A.shp[(size_t)(3 * i + j) * S + s] = shg[i][j];
  }
  // material: mat(Ne,5) = vp, vs, rho, lam, mu with the element index fastest
  const double rho = A.mat[(size_t)e + (size_t)A.Ne * 2];
  const double lam = A.mat[(size_t)e + (size_t)A.Ne * 3], miu = A.mat[(size_t)e + (size_t)A.Ne * 4];
  A.lam[s] = lam; A.mu[s] = miu;
  if (A.rho) { A.rho[s] = rho; A.vp[s] = A.mat[e]; }
  // contm, assembleGlobalMass.f90:376-406 (row-sum lumping; the three dofs of a node carry one value)
  {
    const double totmas = rho * A.w * det;
    double dsum = 0.0, work[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const double temp2 = totmas * shg[j][3] * shg[j][3];
      dsum = dsum + temp2;
      work[j] = 0.0 + temp2;
    }
    const double temp1 = totmas / dsum;
#pragma unroll
    for (int j = 0; j < 8; ++j) A.em[(size_t)j * S + s] = temp1 * work[j];
  }
  // calcSSPhi4Hrgls, assembleGlobalMass.f90:328-374
  {
    const double vol = vlm(xl);
    double ce = miu * (3 * lam + 2 * miu) / (lam + miu);
    ce = 16.0 * ce / 15.0;
    const double co = ce * vol / 48.0;
    // xs = reshape((/cof11,cof12,cof13, cof21,.../),(/3,3/))/det : xs(1,1)=cof11, xs(2,1)=cof12, ...
    double x[3][3];  // x[a-1][b-1] = xs(a,b)
    x[0][0] = cof11 / det; x[1][0] = cof12 / det; x[2][0] = cof13 / det;
    x[0][1] = cof21 / det; x[1][1] = cof22 / det; x[2][1] = cof23 / det;
    x[0][2] = cof31 / det; x[1][2] = cof32 / det; x[2][2] = cof33 / det;
    if (A.compact) {
      A.ss[0 * S + s] = co * (x[0][0] * x[0][0] + x[1][0] * x[1][0] + x[2][0] * x[2][0]);
      A.ss[1 * S + s] = co * (x[0][1] * x[0][1] + x[1][1] * x[1][1] + x[2][1] * x[2][1]);
      A.ss[2 * S + s] = co * (x[0][2] * x[0][2] + x[1][2] * x[1][2] + x[2][2] * x[2][2]);
      return;
    }
    A.ss[0 * S + s] = co * (x[0][0] * x[0][0] + x[1][0] * x[1][0] + x[2][0] * x[2][0]);
    A.ss[1 * S + s] = co * (x[0][0] * x[0][1] + x[1][0] * x[1][1] + x[2][0] * x[2][1]);
    A.ss[2 * S + s] = co * (x[0][0] * x[0][2] + x[1][0] * x[1][2] + x[2][0] * x[2][2]);
    A.ss[3 * S + s] = co * (x[0][1] * x[0][1] + x[1][1] * x[1][1] + x[2][1] * x[2][1]);
    A.ss[4 * S + s] = co * (x[0][1] * x[0][2] + x[1][1] * x[1][2] + x[2][1] * x[2][2]);
    A.ss[5 * S + s] = co * (x[0][2] * x[0][2] + x[1][2] * x[1][2] + x[2][2] * x[2][2]);
    for (int i = 0; i < 4; ++i) {
      double ph[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        double v = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k)
          v = v + c_ha[i][k] * (xl[k][0] * shg[j][0] + xl[k][1] * shg[j][1] + xl[k][2] * shg[j][2]);
        ph[j] = c_ha[i][j] - v;
      }
      double v = 0.0;
#pragma unroll
      for (int j = 0; j < 8; ++j) v = v + ph[j] * ph[j];
      v = sqrt(v / 8.0);
#pragma unroll
      for (int j = 0; j < 8; ++j) A.phi[(size_t)(8 * i + j) * S + s] = ph[j] / v;
    }
  }
}

// Lumped mass of a tile's nodes from its elements, in a fixed order (ascending
// element slot, then local node): one CTA per tile, one thread per tile node.
__global__ void __launch_bounds__(128) k_tile_mass(TileMassArgs A) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const int4 rec = A.tileRec[blockIdx.x];
  const int e0 = rec.x, ne = rec.y & 0xffff, nb = rec.z, ln = rec.w;
  double* em = (double*)smraw;                 // [8][ne]
  uint16_t* lc = (uint16_t*)(em + 8 * A.capE); // [8][ne]
  for (int k = threadIdx.x; k < 8 * ne; k += blockDim.x) {
    const int i = k / ne, le = k - i * ne;
    em[i * A.capE + le] = A.em[(size_t)i * A.S + e0 + le];
    lc[i * A.capE + le] = A.lconn[(size_t)i * A.S + e0 + le] & EQD_LN_MASK;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < ln; i += blockDim.x) {
    double sum = 0.0;
    for (int le = 0; le < ne; ++le)
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (lc[k * A.capE + le] == i) sum = sum + em[k * A.capE + le];
    A.pm[nb + i] = sum;
  }
}

// nodal lumped mass = ordered sum of the node's tile partials (class, ascending tile)
__global__ void k_node_mass(NodeMassArgs A) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= A.Nn) return;
  const int cnt = A.slotCnt[n];
  double m = 0.0;
  for (int k = 0; k < cnt; ++k) {
    const uint32_t u = A.slotTab[(size_t)k * A.NnS + n];
    m = m + A.pm[u & 3][u >> 2];
  }
  A.mass[n] = m;
}

void launch_elem_ops(const OpsArgs& A, cudaStream_t s) {
  if (A.S > 0) k_elem_ops<<<(A.S + 127) / 128, 128, 0, s>>>(A);
}
void launch_tile_mass(const TileMassArgs& A, int ntiles, cudaStream_t s) {
  if (ntiles <= 0) return;
  const size_t sm = (size_t)8 * A.capE * (sizeof(double) + sizeof(uint16_t));
  cudaFuncSetAttribute(k_tile_mass, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  k_tile_mass<<<ntiles, 128, sm, s>>>(A);
}
void launch_node_mass(const NodeMassArgs& A, cudaStream_t s) {
  if (A.Nn > 0) k_node_mass<<<(A.Nn + 255) / 256, 256, 0, s>>>(A);
}

}  // namespace eqd
