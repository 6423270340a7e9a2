// Who reduces what in the fused o_proj + all-reduce launch (oproj_allreduce.cu): the same few formulas on the device
// (phase 2 of the kernel) and on the host (hg_oproj_allreduce_plan, which the CPU tests use to check that every 16-byte
// vector of the [m, n] output is reduced exactly once, by exactly one rank, for any world size, shape and launch geometry).
//
//   tiles      128 x bn, numbered column-major over the m tiles: tile = nt * m_tiles + mt.  CTA c of a rank MULTIPLIES tiles
//              c, c + grid, ...; rank (tile % world) OWNS (reduces) the tile, whoever multiplied it.
//   slices     an owned tile is cut into slices of (u * rows_per_inst) rows; slice s of a rank = owned tile s / slices_per_tile,
//              part s % slices_per_tile; reduce warp w of CTA c walks slices w * grid + c, + grid * warps, ...
//   vectors    instruction j (< u) of a warp covers rows_per_inst rows of the slice: lane l -> row j * rows_per_inst + l / lanes_per_row,
//              16-byte vector l % lanes_per_row of the tile row.
#pragma once

#ifdef __CUDACC__
#define HG_HD __host__ __device__ __forceinline__
#else
#define HG_HD inline
#endif

namespace hg {

constexpr int kOprojBM = 128;

struct OprojGeom {
  int m, n, bn, u, world, rank;
  int m_tiles, n_tiles;                                   // n_tiles = all tiles of the product
  int lanes_per_row, rows_per_inst, unit_rows, units_per_tile;
  int n_own, n_units;                                     // tiles / slices this rank owns
};

HG_HD OprojGeom oproj_geom(int m, int n, int bn, int u, int world, int rank) {
  OprojGeom g;
  g.m = m; g.n = n; g.bn = bn; g.u = u; g.world = world; g.rank = rank;
  g.m_tiles = (m + kOprojBM - 1) / kOprojBM;
  g.n_tiles = g.m_tiles * ((n + bn - 1) / bn);
  g.lanes_per_row = bn / 8;
  g.rows_per_inst = 32 / g.lanes_per_row;
  g.unit_rows = u * g.rows_per_inst;
  g.units_per_tile = kOprojBM / g.unit_rows;
  g.n_own = g.n_tiles > rank ? (g.n_tiles - rank + world - 1) / world : 0;  // tiles rank, rank + world, ...
  g.n_units = g.n_own * g.units_per_tile;
  return g;
}
HG_HD int oproj_owner(int tile, int world) { return tile % world; }
HG_HD int oproj_unit_tile(const OprojGeom& g, int unit) { return g.rank + (unit / g.units_per_tile) * g.world; }
// row / first column of the 16-byte vector lane `lane` touches in instruction j of slice `unit`; false: outside [m, n]
HG_HD bool oproj_unit_vector(const OprojGeom& g, int unit, int j, int lane, int* row, int* col) {
  const int tile = oproj_unit_tile(g, unit);
  const int mt = tile % g.m_tiles, nt = tile / g.m_tiles;
  *row = mt * kOprojBM + (unit % g.units_per_tile) * g.unit_rows + j * g.rows_per_inst + lane / g.lanes_per_row;
  *col = nt * g.bn + (lane % g.lanes_per_row) * 8;
  return *row < g.m && *col < g.n;
}
// first slice of reduce warp w of CTA c, and the stride of its walk
HG_HD int oproj_first_unit(int cta, int warp, int grid) { return warp * grid + cta; }
HG_HD int oproj_unit_stride(int grid, int warps) { return grid * warps; }

}  // namespace hg
