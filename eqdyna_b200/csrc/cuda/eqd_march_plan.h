// Host side of the marching kernel (eqd_march.h): planner output and launch wrappers (eqd_march.cu).
#pragma once
#include <cuda_runtime.h>

#include <vector>

#include "eqd_march.h"
#include "eqd_march_pml.h"
#include "eqd_par.h"

namespace eqd {

struct MarchPlan {
  int n = 0;        // elements placed in bundles
  int S = 0;        // element slots of the class SoA (planes packed: (bz*by rounded up to even) slots each)
  int PFS = 0;      // node slots ((Lx + 1) * MK_PN per bundle)
  int grid = 0;     // CTAs of the persistent launch
  long nFused = 0;  // nodes the bundles update themselves
  std::vector<MarchBundle> rec;
  std::vector<int> ctaFirstA;    // [grid + 1] boundary work list: bundles [0, nBundlesA) of rec, balanced over the CTAs
  std::vector<int> ctaFirstB;    // [grid + 1] interior work list: bundles [nBundlesA, rec.size())
  int nBundlesA = 0;
  raw_vector<int> refId;         // [S] slot -> reference element (0-based), -1 = padding
  raw_vector<unsigned char> owner;  // [S] 1: this slot is the element's own copy, 0: padding or a ghost copy (another strip owns the element)
  raw_vector<int> code;          // [PFS] -1 | node id | MK_FUSED
  std::vector<int> slotBundle;   // [PFS / MK_PN] node plane -> bundle, -1 = padding
  std::vector<int> leftover;     // candidates the bundles do not cover, ascending (they stay in the tile classes)
};

// elems: regular elements on 3-dof nodes, ascending; conn (8,Ne) 0-based; coor = meshCoor(3,Nn); info = node
// kinds (EQD_INFO_KIND); ny, nz from infer_grid; nxg = node planes of the grid in x (0 = unknown: no caps);
// grid = CTAs the launch will have; share = 1: neighbouring strips share an element row / column as ghost copies so
// that the nodes between them are interior to one of them (needs the caller to double-buffer v and d).
// pml = true: elems are PML elements (type 2) on nodes of either kind; no node is updated by the bundles then.
void plan_march(const int* conn, const int* etype, const double* coor, const int* info, const std::vector<int>& elems, int Nn, int ny,
                int nz, int nxg, int grid, int share, MarchPlan& out, bool pml = false);

size_t march_smem_bytes();
int march_ctas_per_sm();   // occupancy of k_march on the current device (0: not launchable)
void launch_march(const MarchArgs& A, int grid, cudaStream_t s);
size_t march_pml_smem_bytes();
int march_pml_ctas_per_sm();
void launch_march_pml(const MarchPmlArgs& A, int grid, cudaStream_t s);
void launch_march_mass(const MarchBundle* rec, int nBundles, const int* slotBundle, const int* code, const double* em, size_t S, int PFS,
                       double* pm, double* mass, cudaStream_t s);

}  // namespace eqd
