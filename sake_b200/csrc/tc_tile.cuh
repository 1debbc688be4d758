// Tile geometry shared by the tcgen05 kernels: a tile is 128 atom pairs = whole receiver rows
// (i, all j) packed together when N <= 128, or a slice of one row when N > 128.
// Ragged batches (Dims.hdr != NULL, N <= 128): tiles come from the device table built by
// sake_ragged_prepare — rows of ONE n_real class per tile, so the segment length is uniform inside a tile.
#pragma once
#include "common.cuh"

namespace sake {

constexpr int TILE = 128;            // atom pairs per tile

struct TileGeom {
  int N, R, rpt, nseg, js, num_tiles;     // num_tiles: the host's (worst-case) count, sizes the grid
  const RaggedHdr* hdr;                   // ragged: actual counts live here
  const int4* rowinfo;
  const int4* tileinfo;
};
static inline TileGeom make_geom(const Dims& d) {
  TileGeom g;
  g.N = d.N; g.R = d.R;
  if (d.N <= TILE) { g.rpt = TILE / d.N; g.nseg = 1; g.js = d.N; g.num_tiles = (d.R + g.rpt - 1) / g.rpt; }
  else { g.rpt = 1; g.nseg = (d.N + TILE - 1) / TILE; g.js = (d.N + g.nseg - 1) / g.nseg; g.num_tiles = d.R * g.nseg; }
  g.hdr = d.hdr; g.rowinfo = d.rowinfo; g.tileinfo = d.tileinfo;
  // ragged worst case: every tile holds at least one row, and a class wastes at most one partial tile
  if (d.hdr) g.num_tiles = d.R;
  return g;
}
// tiles to process (device side)
__device__ __forceinline__ int geom_tiles(const TileGeom& g) { return g.hdr ? g.hdr->num_tiles : g.num_tiles; }

// One tile: rows [row0, row0 + nrows) with n senders each; pair slot of (row0 + lr, j) = pair0 + lr*n + j.
// nseg > 1 (N > 128, uniform only): ONE row, senders [j0, j0 + n).
struct TileDesc { int row0, nrows, n, j0; long long pair0; };
__device__ __forceinline__ TileDesc tile_desc(const TileGeom& g, int tile) {
  TileDesc t;
  if (g.tileinfo) {
    const int4 q = __ldg(g.tileinfo + tile);
    t.row0 = q.x; t.nrows = q.y; t.n = q.z; t.pair0 = q.w; t.j0 = 0;
  } else if (g.nseg == 1) {
    t.row0 = tile * g.rpt; t.nrows = min(g.rpt, g.R - t.row0); t.n = g.N; t.j0 = 0;
    t.pair0 = (long long)t.row0 * g.N;
  } else {
    t.row0 = tile / g.nseg;
    const int seg = tile - t.row0 * g.nseg;
    t.nrows = 1; t.j0 = seg * g.js; t.n = min(g.js, g.N - t.j0);
    t.pair0 = (long long)t.row0 * g.N + t.j0;
  }
  return t;
}
// pair handled by column/row p of a tile: valid, receiver row, sender index j within the molecule, pair slot
__device__ __forceinline__ void tile_pair(const TileDesc& t, int p, bool& valid, int& row, int& j, long long& prx) {
  const int lr = p / t.n;
  const int jj = p - lr * t.n;
  valid = lr < t.nrows;
  row = t.row0 + lr;
  j = t.j0 + jj;
  prx = t.pair0 + p;
}
// first row of the molecule that owns `row` (sender j of that molecule is row mol0 + j)
__device__ __forceinline__ int geom_mol0(const TileGeom& g, int row) {
  return g.rowinfo ? __ldg(g.rowinfo + row).x : (row / g.N) * g.N;
}

}  // namespace sake
