// Tile geometry shared by the tcgen05 kernels: a tile is 128 atom pairs = whole receiver rows
// (i, all j) packed together when N <= 128, or a slice of one row when N > 128.
#pragma once
#include "common.cuh"

namespace sake {

constexpr int TILE = 128;            // atom pairs per tile

struct TileGeom {
  int N, R, rpt, nseg, js, num_tiles;
};
static inline TileGeom make_geom(const Dims& d) {
  TileGeom g;
  g.N = d.N; g.R = d.R;
  if (d.N <= TILE) { g.rpt = TILE / d.N; g.nseg = 1; g.js = d.N; g.num_tiles = (d.R + g.rpt - 1) / g.rpt; }
  else { g.rpt = 1; g.nseg = (d.N + TILE - 1) / TILE; g.js = (d.N + g.nseg - 1) / g.nseg; g.num_tiles = d.R * g.nseg; }
  return g;
}
// pair handled by column/row p of a tile
__device__ __forceinline__ void tile_pair(const TileGeom& g, int tile, int p, bool& valid, int& row, int& j,
                                          bool& seg_end) {
  if (g.nseg == 1) {
    const int lr = p / g.N;
    j = p - lr * g.N;
    row = tile * g.rpt + lr;
    valid = lr < g.rpt && row < g.R;
    seg_end = valid && (j == g.N - 1);
  } else {
    row = tile / g.nseg;
    const int seg = tile - row * g.nseg;
    j = seg * g.js + p;
    const int nj = min(g.js, g.N - seg * g.js);
    valid = p < nj;
    seg_end = valid && (p == nj - 1);
  }
}


}  // namespace sake
