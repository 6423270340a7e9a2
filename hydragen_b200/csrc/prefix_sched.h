// Work schedule of the persistent shared-prefix attention kernel (csrc/prefix_sm100.cu): plain integer arithmetic,
// shared by the device code and by the host (hg_prefix_schedule: the CPU tests walk it and check that every key
// block of every unit is covered exactly once).
//
// A UNIT is one (shared level, group, pair of 128-row query tiles, query head): up to 256 query rows against the
// keys of one shared sequence, i.e. one pass of the kernel's main loop over that sequence's 64-key blocks.  All shared
// levels of a hierarchy (hydragen/attention.py:250-341 runs one flash-attn call per level) are units of ONE launch.
//
// mode 0 (stream-K): the units are laid end to end on a cost axis (cost of a unit = its key blocks + a fixed c0 for
// its prologue / epilogue) and the axis is cut into n_ctas equal ranges, one per persistent CTA (one CTA per SM).  A
// cut that falls inside a unit splits it into PIECES; every piece of a split unit leaves an fp32 partial result
// (unnormalised O, running max, row sum) in the workspace and the pieces' CTAs merge them, each a slice of the rows.
// Cuts are snapped so that no piece is shorter than min_piece blocks.
// mode 1: whole units only, unit u -> CTA u % n_ctas (no workspace; the causal instantiation, whose units differ in
// length, orders its units heavy-first).
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#define HG_HD __host__ __device__ __forceinline__
#else
#define HG_HD inline
#endif

namespace hg {

constexpr int kMaxLevels = 4;       // shared levels per launch
constexpr int kMaxCtas = 160;       // persistent CTAs per launch (>= SMs of the device: 148)
constexpr int kMaxUnitPieces = 24;  // pieces of one unit the merge handles (the host sizes the grid accordingly)

struct SchedLevel {
  int n_groups, q_per_group, tiles_per_group;
  int nb_max;   // key blocks of the longest group (the schedule's cost of every unit of the level)
  int n_units;  // n_groups * tiles_per_group * hq
  int unit0;    // index of the level's first unit
  long long cost0;  // position of the level's first unit on the cost axis
};

struct UnitPos {
  int unit, blk;
};

struct SchedParams {
  SchedLevel lv[kMaxLevels];
  int n_levels, hq, n_ctas, mode;
  int c0, min_piece, total_units;
  int heavy_first;  // unit order: row tiles descending first (the causal instantiation: later tiles see more keys)
  long long total_cost;
  // mode 0: the snapped cuts, tabulated by the host (sched_fill_bounds) -- the device never divides 64-bit numbers
  // (r02f trace: computing them per CTA and per split piece cost microseconds of dependent divisions)
  UnitPos bounds[kMaxCtas + 1];
};
HG_HD bool operator<(const UnitPos& a, const UnitPos& b) { return a.unit < b.unit || (a.unit == b.unit && a.blk < b.blk); }
HG_HD bool operator==(const UnitPos& a, const UnitPos& b) { return a.unit == b.unit && a.blk == b.blk; }

struct SchedPiece {
  int unit, level;
  int head, grp, mt;  // decoded unit
  int b_lo, b_hi;     // key blocks [b_lo, b_hi) of the unit (before clipping to the group's actual length)
  int split;          // 1: the piece is not the whole unit -> partial result to the workspace
  int slot;           // workspace slot of this CTA the partial goes to (0: the CTA's first piece, 1: any later one)
};

HG_HD int sched_level_of_unit(const SchedParams& S, int unit) {
  int l = 0;
  while (l + 1 < S.n_levels && unit >= S.lv[l + 1].unit0) ++l;
  return l;
}

// Cut number j (0 .. n_ctas) of the cost axis, snapped: CTA j owns [boundary(j), boundary(j + 1)).
inline UnitPos sched_boundary_compute(const SchedParams& S, int j) {
  if (j <= 0) return UnitPos{0, 0};
  if (j >= S.n_ctas) return UnitPos{S.total_units, 0};
  const long long x = (long long)j * S.total_cost / S.n_ctas;
  int l = 0;
  while (l + 1 < S.n_levels && x >= S.lv[l + 1].cost0) ++l;
  const SchedLevel& L = S.lv[l];
  const long long w = L.nb_max + S.c0;
  long long uu = (x - L.cost0) / w;
  int b = (int)((x - L.cost0) % w) - S.c0;
  if (b < S.min_piece) {
    b = 0;
  } else if (L.nb_max - b < S.min_piece) {
    ++uu;
    b = 0;
  }
  return UnitPos{L.unit0 + (int)uu, b};
}

inline void sched_fill_bounds(SchedParams& S) {
  for (int j = 0; j <= S.n_ctas && j <= kMaxCtas; ++j) S.bounds[j] = sched_boundary_compute(S, j);
}
HG_HD UnitPos sched_boundary(const SchedParams& S, int j) { return S.bounds[j < 0 ? 0 : (j > S.n_ctas ? S.n_ctas : j)]; }

HG_HD void sched_decode_unit(const SchedParams& S, int unit, SchedPiece& p) {
  const int l = sched_level_of_unit(S, unit);
  const SchedLevel& L = S.lv[l];
  const int uu = unit - L.unit0;
  p.unit = unit;
  p.level = l;
  int r;
  if (S.heavy_first) {
    const int per = S.hq * L.n_groups;
    p.mt = L.tiles_per_group - 1 - uu / per;
    r = uu % per;
  } else {
    // consecutive units = the row tiles of one (head, group): they read the same K/V
    p.mt = uu % L.tiles_per_group;
    r = uu / L.tiles_per_group;
  }
  p.grp = r % L.n_groups;
  p.head = r / L.n_groups;
}

// Iterator over the pieces of CTA `cta`, in the order the kernel processes them.
struct SchedIter {
  UnitPos cur, end;
  int first;
};

HG_HD void sched_begin(const SchedParams& S, int cta, SchedIter& it) {
  if (S.mode == 0) {
    it.cur = sched_boundary(S, cta);
    it.end = sched_boundary(S, cta + 1);
  } else {
    it.cur = UnitPos{cta, 0};
    it.end = UnitPos{S.total_units, 0};
  }
  it.first = 1;
}

HG_HD bool sched_next(const SchedParams& S, SchedIter& it, SchedPiece& p) {
  if (!(it.cur < it.end)) return false;
  sched_decode_unit(S, it.cur.unit, p);
  const int nb = S.lv[p.level].nb_max;
  p.b_lo = it.cur.blk;
  if (S.mode == 0) {
    p.b_hi = (it.cur.unit == it.end.unit) ? it.end.blk : nb;
    it.cur = UnitPos{it.cur.unit + 1, 0};
    if (p.unit == it.end.unit) it.cur = it.end;
  } else {
    p.b_hi = nb;
    it.cur = UnitPos{it.cur.unit + S.n_ctas, 0};
  }
  p.split = !(p.b_lo == 0 && p.b_hi == nb);
  p.slot = it.first ? 0 : 1;
  it.first = 0;
  return true;
}

// The CTAs that hold a piece of split unit `unit`, in key order, with the workspace slot each one used.
// `near` = any CTA known to hold a piece of it (the search starts there).  Returns the number of pieces.
HG_HD int sched_unit_pieces(const SchedParams& S, int unit, int near, int* ctas, int* slots) {
  const int nb = S.lv[sched_level_of_unit(S, unit)].nb_max;
  const UnitPos u0{unit, 0}, u1{unit, nb};
  int jf = near;
  while (jf > 0 && u0 < sched_boundary(S, jf)) --jf;
  int n = 0;
  for (int j = jf; j < S.n_ctas; ++j) {
    const UnitPos a = sched_boundary(S, j), b = sched_boundary(S, j + 1);
    if (!(a < u1)) break;
    const UnitPos lo = a < u0 ? u0 : a, hi = u1 < b ? u1 : b;
    if (lo < hi) {
      if (n < kMaxUnitPieces) {
        ctas[n] = j;
        slots[n] = (a.unit == unit) ? 0 : 1;  // the piece is CTA j's first one iff j's range starts inside this unit
      }
      ++n;
    }
  }
  return n;
}

}  // namespace hg
