// Host-side tile planner of the step library: cuts one element class into
// bricks ("tiles") that one CTA sweeps and assembles in shared memory
// (eqd_kernels.cu: k_tile_reg / k_tile_pml).  Pure index work on the
// reference's connectivity (nodeElemIdRelation, meshgen.f90:702-741); it never
// changes which nodes an element touches, only where the element is stored.
#pragma once
#include <cstdint>
#include <vector>

#include "eqd_par.h"

namespace eqd {

struct TilePlan {
  int n = 0;        // elements of the class
  int S = 0;        // padded slot count (SoA row length)
  int nTiles = 0;
  int LS = 0;       // max nodes of a tile (shared-memory row stride)
  int PFS = 0;      // padded total of tile-node slots (row length of the partial buffer)
  raw_vector<int> refId;          // [S] slot -> reference element id (0-based), -1 = padding
  std::vector<int> tileElem;       // [nTiles] first slot (multiple of 32)
  std::vector<int> tileCnt;        // [nTiles] elements
  std::vector<int> tileNode;       // [nTiles+1] first tile-node slot (multiple of 4)
  std::vector<uint8_t> tileColours;  // [nTiles] colours per assembly phase (1 = conflict free)
  raw_vector<int> tnode;          // [PFS] node id per tile-node slot, ascending inside a tile, -1 = padding
  raw_vector<uint16_t> lconn;     // [8][S] tile-local node | colour << 12
  // modelled shared-memory wavefronts of the 8 corner accesses of every half-warp (16 lanes, 8-byte words: a
  // wavefront serves one word per bank pair): conflict-free, with ascending element order, with the chosen order
  long bankIdeal = 0, bankAscending = 0, bankChosen = 0;
};

// default bricks (elements along x, z, y), measured on TPV104 @ 100 m (tools/tune_tiles.py): a regular
// brick of 4x4x16 is halved by the 400-node cap into two 2x4x16 tiles = exactly one 128-element stage
// with 16-element y runs (bank-conflict free in shared memory); PML slabs are 6 elements thick in any
// direction, 6x7x6 fills four 64-element stages whatever the slab's orientation
constexpr int kRegBrick[3] = {4, 4, 16};
constexpr int kPmlBrick[3] = {6, 7, 6};

struct TileShape {
  int bx = 4, bz = 4, by = 32;  // target brick, in elements, along the x / z / y grid axes
  int capE = 640;               // hard cap on elements per tile
  int capN = 1280;              // hard cap on nodes per tile (< 4096: 12-bit local ids)
  int bankOrder = 0;            // 1: order the elements inside a tile so that a half-warp's 16 gathers / updates of
                                //    one corner fall into 16 different shared-memory bank pairs where a permutation of
                                //    the grid axes allows it (0 = ascending reference id: y fastest, then z, then x)
                                // 2: keep the element order, number the tile-local NODES of complete bricks by residue
                                //    (index mod 16 == (iy + ey*iz + ey*ez*ix) mod 16): conflict free by construction
};

// Node-grid strides of the structured part of the mesh, inferred from a plain
// brick's connectivity: node id = ix*ny*nz + iz*ny + iy (meshgen.f90:64-107).
// Returns false when no brick reveals them (tiles then follow storage order).
bool infer_grid(const int* conn, const int* etype, int Ne, int Nn, int& ny, int& nz);

// elems: the class's reference element ids, ascending.  conn: (8,Ne) 0-based.
void plan_tiles(const int* conn, const std::vector<int>& elems, int Nn, int ny, int nz, bool gridOk,
                const TileShape& shape, int threadsPerTile, TilePlan& out);

}  // namespace eqd
